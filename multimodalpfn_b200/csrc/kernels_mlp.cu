// kernels_mlp.cu — the whole MLP sublayer of a PerFeatureEncoderLayer in ONE persistent kernel:
//
//     state <- LayerNorm(state + W2 gelu(W1 state))            (mlp.py:93-138, layer.py:437-455)
//
// The 768-wide hidden activation never leaves the SM.  Per 128-token tile the hidden axis is walked
// in 12 chunks of 64: GEMM1 (tcgen05, 128 x 64 x 192) leaves the chunk in TMEM, eight GELU warps
// turn it into a bf16 K-major operand tile in shared memory, GEMM2 (128 x 192 x 64) accumulates it
// into the output tile in TMEM, and four epilogue warps add the fp32 residual, normalise and write
// the fp32 state and its bf16 shadow.  HBM traffic per token is 192 x (2 + 4 + 4 + 2) B instead of
// the 192 x 28 B of the two-GEMM form (the 768 x 2 B hidden row written and read back).
//
// One CTA per SM, all 512 TMEM columns: two output accumulators (the epilogue of tile i runs under
// the MMAs of tile i+1) and two hidden-chunk accumulators (GEMM1 of chunk g+1 runs under the GELU
// of chunk g).
//
//   warp 0        TMA producer: activation tiles (2 buffers) and W1/W2 chunks (2 stages)
//   warp 1        MMA issue
//   warp 2        TMEM allocation
//   warps 4-11    GELU: TMEM lane quarter = warp % 4, column half = (warp - 4) / 4
//   warps 12-15   epilogue: residual + LayerNorm, thread = token row
//
// GELU: 0.5 x (1 + erf(x / sqrt 2)) with erf(z) = tanh(g(z)), g an odd polynomial fitted to
// atanh(erf) (max |error| of the GELU 3e-5 before the hardware tanh, whose 2^-11 relative error is
// a quarter of the bf16 rounding the hidden activation gets anyway); one MUFU.TANH and ~5 issue
// slots per element with packed f32x2 arithmetic, against ~30 for erff().
#include "tc_common.cuh"

namespace mmpfn {
namespace {

constexpr int F_BM = 128;                         // tokens per tile
constexpr int F_CH = 64;                          // hidden columns per chunk
constexpr int F_NCH = kHid / F_CH;                // 12
constexpr int F_A_BYTES = F_BM * kE * 2;          // 48 KB: 3 k-blocks [128][64] bf16, 128B swizzle
constexpr int F_W1_BYTES = F_CH * kE * 2;         // 24 KB: 3 k-blocks [64][64]
constexpr int F_W2_BYTES = kE * F_CH * 2;         // 24 KB: [192][64]
constexpr int F_WSTAGE = F_W1_BYTES + F_W2_BYTES;
constexpr int F_HS_BYTES = F_BM * F_CH * 2;       // 16 KB: [128][64] bf16, 128B swizzle
constexpr int F_OFF_A = 0;
constexpr int F_OFF_W = 2 * F_A_BYTES;
constexpr int F_OFF_HS = F_OFF_W + 2 * F_WSTAGE;
constexpr int F_OFF_BAR = F_OFF_HS + 2 * F_HS_BYTES;
constexpr int F_SMEM = F_OFF_BAR + 256 + 1024;
constexpr int F_THREADS = 512;
constexpr uint32_t F_TM_OUT = 0;                  // 2 x 192 columns
constexpr uint32_t F_TM_H = 384;                  // 2 x 64 columns

struct MlpArgs {
  int M, n_tiles;
  float* resid;
  uint16_t* ln_bf16;
};

__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint64_t mul_f32x2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
// gelu of two values; returns them packed as bf16x2 (lo = first)
__device__ __forceinline__ uint32_t gelu2_bf16(float x0, float x1) {
  const uint64_t x = pack_f32x2(x0, x1);
  float q0, q1;
  unpack_f32x2(mul_f32x2(x, x), q0, q1);
  // beyond |x| = 6 the polynomial is frozen: u = 1.67 x there, tanh(u) = +-1 in fp32
  const uint64_t x2 = pack_f32x2(fminf(q0, 36.0f), fminf(q1, 36.0f));
  const uint64_t a0 = pack_f32x2(0.7974578142166138f, 0.7974578142166138f);
  const uint64_t a1 = pack_f32x2(0.03705103322863579f, 0.03705103322863579f);
  const uint64_t a2 = pack_f32x2(-0.0003588674298953265f, -0.0003588674298953265f);
  const uint64_t half = pack_f32x2(0.5f, 0.5f);
  uint64_t p = fma_f32x2(a2, x2, a1);
  p = fma_f32x2(p, x2, a0);
  float u0, u1;
  unpack_f32x2(mul_f32x2(p, x), u0, u1);
  const uint64_t t = pack_f32x2(tanh_approx(u0), tanh_approx(u1));
  const uint64_t h = mul_f32x2(x, half);
  float y0, y1;
  unpack_f32x2(fma_f32x2(h, t, h), y0, y1);
  return pack_bf16x2(y0, y1);
}

__global__ void __launch_bounds__(F_THREADS, 1) tc_mlp_kernel(const __grid_constant__ CUtensorMap map_a,
                                                            const __grid_constant__ CUtensorMap map_w1,
                                                            const __grid_constant__ CUtensorMap map_w2,
                                                            const MlpArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + F_OFF_BAR);
  uint64_t* a_full = bars;             // [2] activation tile landed
  uint64_t* a_empty = bars + 2;        // [2] ... and all GEMM1 of its tile have read it
  uint64_t* w_full = bars + 4;         // [2] W1/W2 chunk landed
  uint64_t* w_empty = bars + 6;        // [2] ... and GEMM2 of its chunk has completed
  uint64_t* h_full = bars + 8;         // [2] hidden chunk accumulator complete in TMEM
  uint64_t* h_free = bars + 10;        // [2] ... and read out by the GELU warps (256 arrivals)
  uint64_t* hs_full = bars + 12;       // [2] gelu(chunk) is in shared memory (256 arrivals)
  uint64_t* hs_empty = bars + 14;      // [2] ... and GEMM2 has consumed it
  uint64_t* out_full = bars + 16;      // [2] output accumulator of a tile complete
  uint64_t* out_empty = bars + 18;     // [2] ... and drained by the epilogue (128 arrivals)
  uint32_t* tmem_slot = (uint32_t*)(bars + 20);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // tiles of this CTA: blockIdx.x, blockIdx.x + gridDim.x, ...
  const int my_tiles = (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int G = my_tiles * F_NCH;      // chunks this CTA walks, flat over its tiles

  if (threadIdx.x == 0) {
    prefetch_tmap(&map_a);
    prefetch_tmap(&map_w1);
    prefetch_tmap(&map_w2);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 1);
      mbar_init(&w_full[s], 1);
      mbar_init(&w_empty[s], 1);
      mbar_init(&h_full[s], 1);
      mbar_init(&h_free[s], 256);
      mbar_init(&hs_full[s], 256);
      mbar_init(&hs_empty[s], 1);
      mbar_init(&out_full[s], 1);
      mbar_init(&out_empty[s], 128);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      auto load_a = [&](int i) {
        const int tile = (int)blockIdx.x + i * (int)gridDim.x;
        const int ab = i & 1;
        mbar_wait(&a_empty[ab], ((i >> 1) & 1) ^ 1);
        mbar_expect_tx(&a_full[ab], F_A_BYTES);
        uint8_t* a_dst = smem + F_OFF_A + ab * F_A_BYTES;
#pragma unroll
        for (int kb = 0; kb < 3; ++kb) tma_load_2d(a_dst + kb * (F_BM * 128), &map_a, &a_full[ab], kb * 64, tile * F_BM);
      };
      if (my_tiles > 0) load_a(0);
      for (int i = 0; i < my_tiles; ++i) {
        for (int c = 0; c < F_NCH; ++c) {
          // the next tile's activations are requested ten chunks before its first GEMM1
          if (c == 2 && i + 1 < my_tiles) load_a(i + 1);
          const int g = i * F_NCH + c;
          const int ws = g & 1;
          mbar_wait(&w_empty[ws], ((g >> 1) & 1) ^ 1);
          mbar_expect_tx(&w_full[ws], F_WSTAGE);
          uint8_t* w_dst = smem + F_OFF_W + ws * F_WSTAGE;
#pragma unroll
          for (int kb = 0; kb < 3; ++kb) tma_load_2d(w_dst + kb * (F_CH * 128), &map_w1, &w_full[ws], kb * 64, c * F_CH);
          tma_load_2d(w_dst + F_W1_BYTES, &map_w2, &w_full[ws], c * F_CH, 0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc1 = make_idesc(F_BM, F_CH);
      constexpr uint32_t idesc2 = make_idesc(F_BM, kE);
      const uint32_t sbase = smem_u32(smem);
      // GEMM1(g): hidden chunk g = A(tile) W1c^T into TMEM H[g & 1]
      auto gemm1 = [&](int g) {
        const int i = g / F_NCH, c = g % F_NCH;
        const int hb = g & 1, ab = i & 1;
        if (c == 0) mbar_wait(&a_full[ab], (i >> 1) & 1);
        mbar_wait(&w_full[hb], (g >> 1) & 1);
        mbar_wait(&h_free[hb], ((g >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t a_addr = sbase + F_OFF_A + ab * F_A_BYTES;
        const uint32_t w_addr = sbase + F_OFF_W + hb * F_WSTAGE;
#pragma unroll
        for (int kb = 0; kb < 3; ++kb) {
          const uint64_t adesc = make_desc(a_addr + kb * (F_BM * 128), 1024, kSw128);
          const uint64_t bdesc = make_desc(w_addr + kb * (F_CH * 128), 1024, kSw128);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem + F_TM_H + hb * F_CH, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc1,
                      (kb | k) != 0);
        }
        umma_commit(&h_full[hb]);
        if (c == F_NCH - 1) umma_commit(&a_empty[ab]);     // last GEMM1 of the tile: A buffer is free
      };
      // GEMM2(g): OUT[tile & 1] (+)= gelu(chunk g) W2c^T
      auto gemm2 = [&](int g) {
        const int i = g / F_NCH, c = g % F_NCH;
        const int hb = g & 1, ob = i & 1;
        if (c == 0) mbar_wait(&out_empty[ob], ((i >> 1) & 1) ^ 1);
        mbar_wait(&hs_full[hb], (g >> 1) & 1);
        tc_fence_after();
        const uint64_t adesc = make_desc(sbase + F_OFF_HS + hb * F_HS_BYTES, 1024, kSw128);
        const uint64_t bdesc = make_desc(sbase + F_OFF_W + hb * F_WSTAGE + F_W1_BYTES, 1024, kSw128);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem + F_TM_OUT + ob * kE, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc2, (c | k) != 0);
        umma_commit(&w_empty[hb]);
        umma_commit(&hs_empty[hb]);
        if (c == F_NCH - 1) umma_commit(&out_full[ob]);
      };
      if (G > 0) gemm1(0);
      for (int g = 0; g < G; ++g) {
        if (g + 1 < G) gemm1(g + 1);     // runs under the GELU of chunk g
        gemm2(g);
      }
    }
  } else if (warp >= 4 && warp < 12) {
    // ---- GELU warps ----
    const int quarter = warp & 3, half = (warp - 4) >> 2;
    const int r = quarter * 32 + lane;
    const uint32_t trow = tmem + F_TM_H + ((uint32_t)(quarter * 32) << 16) + half * 32;
    const uint32_t hs_row = smem_u32(smem) + F_OFF_HS + r * 128;
    const int rsw = r & 7;
    uint32_t v[32];
    for (int g = 0; g < G; ++g) {
      const int hb = g & 1;
      const uint32_t ph = (g >> 1) & 1;
      mbar_wait(&h_full[hb], ph);
      tc_fence_after();
      tmem_ld32(trow + hb * F_CH, v);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&h_free[hb]);
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) pk[i] = gelu2_bf16(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
      mbar_wait(&hs_empty[hb], ph ^ 1);          // GEMM2 of chunk g-2 has released this buffer
      // 32 hidden columns = 64 B = chunks [half*4, half*4+4) of this row's 128 B, XOR-swizzled
      const uint32_t dst = hs_row + hb * F_HS_BYTES;
#pragma unroll
      for (int q = 0; q < 4; ++q)
        st_shared_v4(dst + (((half * 4 + q) ^ rsw) << 4), pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
      fence_proxy_async();
      mbar_arrive(&hs_full[hb]);
    }
  } else if (warp >= 12) {
    // ---- epilogue warps: state = LN(state + acc), fp32 state and bf16 shadow; thread = row ----
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    uint32_t v[32];
    for (int i = 0; i < my_tiles; ++i) {
      const int tile = (int)blockIdx.x + i * (int)gridDim.x;
      const int ob = i & 1;
      const uint32_t trow = tmem + F_TM_OUT + ob * kE + ((uint32_t)(quarter * 32) << 16);
      const long long m = (long long)tile * F_BM + r;
      const bool ok = m < p.M;
      float* res = p.resid + m * kE;
      mbar_wait(&out_full[ob], (i >> 1) & 1);
      tc_fence_after();
      float sum = 0.f, sq = 0.f;
#pragma unroll 1
      for (int c = 0; c < kE / 32; ++c) {
        tmem_ld32(trow + c * 32, v);
        tmem_ld_wait();
        if (ok) {
          const float4* r4 = reinterpret_cast<const float4*>(res + c * 32);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float4 rr = r4[k];
            const float a0 = __uint_as_float(v[4 * k]) + rr.x, a1 = __uint_as_float(v[4 * k + 1]) + rr.y,
                        a2 = __uint_as_float(v[4 * k + 2]) + rr.z, a3 = __uint_as_float(v[4 * k + 3]) + rr.w;
            sum += (a0 + a1) + (a2 + a3);
            sq = fmaf(a0, a0, sq); sq = fmaf(a1, a1, sq); sq = fmaf(a2, a2, sq); sq = fmaf(a3, a3, sq);
            v[4 * k] = __float_as_uint(a0); v[4 * k + 1] = __float_as_uint(a1);
            v[4 * k + 2] = __float_as_uint(a2); v[4 * k + 3] = __float_as_uint(a3);
          }
        }
        tmem_st32(trow + c * 32, v);
      }
      tmem_st_wait();
      const float mean = sum * (1.0f / kE);
      const float var = fmaxf(sq * (1.0f / kE) - mean * mean, 0.f);
      const float rstd = rsqrtf(var + kLnEps);
      uint16_t* lnb = p.ln_bf16 + m * kE;
#pragma unroll 1
      for (int c = 0; c < kE / 32; ++c) {
        tmem_ld32(trow + c * 32, v);
        tmem_ld_wait();
        if (ok) {
          float y[32];
#pragma unroll
          for (int k = 0; k < 32; ++k) y[k] = (__uint_as_float(v[k]) - mean) * rstd;
          float4* o4 = reinterpret_cast<float4*>(res + c * 32);
#pragma unroll
          for (int k = 0; k < 8; ++k) o4[k] = make_float4(y[4 * k], y[4 * k + 1], y[4 * k + 2], y[4 * k + 3]);
          uint4* b4 = reinterpret_cast<uint4*>(lnb + c * 32);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            b4[k] = make_uint4(pack_bf16x2(y[8 * k], y[8 * k + 1]), pack_bf16x2(y[8 * k + 2], y[8 * k + 3]),
                               pack_bf16x2(y[8 * k + 4], y[8 * k + 5]), pack_bf16x2(y[8 * k + 6], y[8 * k + 7]));
        }
      }
      tc_fence_before();
      mbar_arrive(&out_empty[ob]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

}  // namespace

int launch_tc_mlp(const TcMlp& p, cudaStream_t st) {
  if (p.M <= 0) return MMPFN_OK;
  if (!p.state_b || !p.w1 || !p.w2 || !p.resid_f32) { set_error("tc_mlp: null argument"); return MMPFN_EINVAL; }
  CUtensorMap ma, mw1, mw2;
  {
    const cuuint64_t dims[2] = {(cuuint64_t)kE, (cuuint64_t)p.M};
    const cuuint64_t strides[1] = {(cuuint64_t)kE * 2};
    const cuuint32_t box[2] = {64, F_BM};
    MMPFN_TRY(encode_map(&ma, p.state_b, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B));
  }
  {
    const cuuint64_t dims[2] = {(cuuint64_t)kE, (cuuint64_t)kHid};
    const cuuint64_t strides[1] = {(cuuint64_t)kE * 2};
    const cuuint32_t box[2] = {64, F_CH};
    MMPFN_TRY(encode_map(&mw1, p.w1, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B));
  }
  {
    const cuuint64_t dims[2] = {(cuuint64_t)kHid, (cuuint64_t)kE};
    const cuuint64_t strides[1] = {(cuuint64_t)kHid * 2};
    const cuuint32_t box[2] = {F_CH, kE};
    MMPFN_TRY(encode_map(&mw2, p.w2, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B));
  }
  static int n_sm = 0;
  if (!n_sm) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    cudaFuncSetAttribute(tc_mlp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, F_SMEM);
  }
  MlpArgs a{};
  a.M = p.M;
  a.n_tiles = (p.M + F_BM - 1) / F_BM;
  a.resid = p.resid_f32;
  a.ln_bf16 = p.state_b_out;
  const int grid = a.n_tiles < n_sm ? a.n_tiles : n_sm;
  tc_mlp_kernel<<<grid, F_THREADS, F_SMEM, st>>>(ma, mw1, mw2, a);
  return count_launch();
}

}  // namespace mmpfn
