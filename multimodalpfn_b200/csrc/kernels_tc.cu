// kernels_tc.cu — the tensor-core path (bf16 operands, fp32 accumulation) written directly against
// sm_100a: TMA (cp.async.bulk.tensor) stages 128B/64B-swizzled tiles in shared memory, a single
// elected thread issues tcgen05.mma with accumulators in TMEM, epilogue / softmax warps read them
// back with tcgen05.ld.  Two kernels:
//
//   tc_gemm_kernel       out = epi(A W^T), 128 x 192 x 64 tiles.  Serves the QKV / output
//                        projections (multi_head_attention.py:430, :513-517) and the MLP
//                        (mlp.py:93-104).  Epilogues: bf16 store, exact GELU, residual + LayerNorm
//                        (layer.py:437-455), and the item-attention QKV scatter.
//   tc_item_attn_kernel  flash attention across items (layer.py:341-379), d = 32: S = Q K^T and
//                        O += P V on tcgen05, online softmax in fp32 by four warps.
//
// Every mbarrier wait is bounded (trap after ~2 s) so that a protocol bug cannot hang the GPU.
#include "tc_common.cuh"

namespace mmpfn {
namespace {

// ---------------------------------------------------------------------------------------------
// GEMM
// ---------------------------------------------------------------------------------------------
constexpr int G_BM = 128, G_BN = 192, G_BK = 64, G_STAGES = 2;
constexpr int G_A_BYTES = G_BM * G_BK * 2;           // 16 KB
constexpr int G_B_BYTES = G_BN * G_BK * 2;           // 24 KB
constexpr int G_STAGE_BYTES = G_A_BYTES + G_B_BYTES;
constexpr int G_SMEM = G_STAGES * G_STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
constexpr int G_TMEM_COLS = 256;
constexpr int G_THREADS = 192;

struct GemmArgs {
  int M, N, K;
  int items, B, S, T, tiles_s;
  uint16_t* out_bf16;
  float* resid;
  uint16_t* ln_bf16;
  uint16_t *q_out, *k_out, *vt_out, *k0_out, *vt0_out;
  int S_pad;
};

template <int EPI>
__global__ void __launch_bounds__(G_THREADS) tc_gemm_kernel(const __grid_constant__ CUtensorMap map_a,
                                                            const __grid_constant__ CUtensorMap map_w,
                                                            const GemmArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = (uint64_t*)(smem + G_STAGES * G_STAGE_BYTES);
  uint64_t* empty = full + G_STAGES;
  uint64_t* acc_full = empty + G_STAGES;
  uint32_t* tmem_slot = (uint32_t*)(acc_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int KB = p.K / G_BK;
  const int n_tiles = p.N / G_BN;
  const int n0 = (blockIdx.x % n_tiles) * G_BN;
  const int m_tile = blockIdx.x / n_tiles;

  // tile coordinates
  int m0 = 0, tb = 0, tt = 0, s0 = 0;
  if (p.items) {
    const int mt = m_tile;
    const int per_b = p.T * p.tiles_s;
    tb = mt / per_b;
    const int r = mt % per_b;
    tt = r / p.tiles_s;
    s0 = (r % p.tiles_s) * G_BM;
  } else {
    m0 = m_tile * G_BM;
  }

  if (threadIdx.x == 0) {
    prefetch_tmap(&map_a);
    prefetch_tmap(&map_w);
    for (int s = 0; s < G_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, G_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < KB; ++kb) {
        const int s = kb % G_STAGES;
        mbar_wait(&empty[s], ((kb / G_STAGES) & 1) ^ 1);
        uint8_t* a_dst = smem + s * G_STAGE_BYTES;
        uint8_t* b_dst = a_dst + G_A_BYTES;
        mbar_expect_tx(&full[s], G_STAGE_BYTES);
        if (p.items) tma_load_4d(a_dst, &map_a, &full[s], kb * G_BK, tt, s0, tb);
        else tma_load_2d(a_dst, &map_a, &full[s], kb * G_BK, m0);
        tma_load_2d(b_dst, &map_w, &full[s], kb * G_BK, n0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(G_BM, G_BN);
      for (int kb = 0; kb < KB; ++kb) {
        const int s = kb % G_STAGES;
        mbar_wait(&full[s], (kb / G_STAGES) & 1);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + s * G_STAGE_BYTES);
        const uint64_t adesc = make_desc(a_addr, 1024, kSw128);
        const uint64_t bdesc = make_desc(a_addr + G_A_BYTES, 1024, kSw128);
#pragma unroll
        for (int k = 0; k < G_BK / 16; ++k)
          umma_bf16(tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0);
        umma_commit(&empty[s]);
      }
      umma_commit(acc_full);
    }
  } else {
    // ---- epilogue: 4 warps, warp w owns TMEM lanes [32*(w%4), +32) = tile rows ----
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t trow = tmem + ((uint32_t)(quarter * 32) << 16);
    mbar_wait(acc_full, 0);
    tc_fence_after();
    uint32_t v[32];
    if (EPI == TC_EPI_BF16 || EPI == TC_EPI_GELU_BF16) {
      const long long m = (long long)m0 + r;
      const bool ok = m < p.M;
      uint16_t* dst = p.out_bf16 + m * p.N + n0;
#pragma unroll 1
      for (int c = 0; c < G_BN / 32; ++c) {
        tmem_ld32(trow + c * 32, v);
        tmem_ld_wait();
        if (ok) {
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float a = __uint_as_float(v[2 * i]), b = __uint_as_float(v[2 * i + 1]);
            if (EPI == TC_EPI_GELU_BF16) { a = gelu_exact(a); b = gelu_exact(b); }
            pk[i] = pack_bf16x2(a, b);
          }
          uint4* d4 = reinterpret_cast<uint4*>(dst + c * 32);
#pragma unroll
          for (int i = 0; i < 4; ++i) d4[i] = make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
        }
      }
    } else if (EPI == TC_EPI_RESID_LN) {
      // state = LN(state + acc): pass 1 adds the residual, keeps v in TMEM and accumulates the
      // statistics; pass 2 normalises and writes the fp32 state and its bf16 shadow.
      const long long m = (long long)m0 + r;
      const bool ok = m < p.M;
      float* res = p.resid + m * kE;
      float sum = 0.f, sq = 0.f;
#pragma unroll 1
      for (int c = 0; c < kE / 32; ++c) {
        tmem_ld32(trow + c * 32, v);
        tmem_ld_wait();
        if (ok) {
          const float4* r4 = reinterpret_cast<const float4*>(res + c * 32);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 rr = r4[i];
            const float a0 = __uint_as_float(v[4 * i]) + rr.x, a1 = __uint_as_float(v[4 * i + 1]) + rr.y,
                        a2 = __uint_as_float(v[4 * i + 2]) + rr.z, a3 = __uint_as_float(v[4 * i + 3]) + rr.w;
            sum += (a0 + a1) + (a2 + a3);
            sq = fmaf(a0, a0, sq); sq = fmaf(a1, a1, sq); sq = fmaf(a2, a2, sq); sq = fmaf(a3, a3, sq);
            v[4 * i] = __float_as_uint(a0); v[4 * i + 1] = __float_as_uint(a1);
            v[4 * i + 2] = __float_as_uint(a2); v[4 * i + 3] = __float_as_uint(a3);
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
          for (int i = 0; i < 32; ++i) y[i] = (__uint_as_float(v[i]) - mean) * rstd;
          float4* o4 = reinterpret_cast<float4*>(res + c * 32);
#pragma unroll
          for (int i = 0; i < 8; ++i) o4[i] = make_float4(y[4 * i], y[4 * i + 1], y[4 * i + 2], y[4 * i + 3]);
          uint4* b4 = reinterpret_cast<uint4*>(lnb + c * 32);
#pragma unroll
          for (int i = 0; i < 4; ++i)
            b4[i] = make_uint4(pack_bf16x2(y[8 * i], y[8 * i + 1]), pack_bf16x2(y[8 * i + 2], y[8 * i + 3]),
                               pack_bf16x2(y[8 * i + 4], y[8 * i + 5]), pack_bf16x2(y[8 * i + 6], y[8 * i + 7]));
        }
      }
    } else {  // TC_EPI_QKV_ITEMS: n-tile j in {q,k,v}; 32-column chunk c = head
      const int s = s0 + r;
      const bool ok = s < p.S;
      const int j = blockIdx.x % n_tiles;
      const long long bt = (long long)tb * p.T + tt;
#pragma unroll 1
      for (int h = 0; h < kH; ++h) {
        tmem_ld32(trow + h * 32, v);
        tmem_ld_wait();
        if (!ok) continue;
        const long long plane = bt * kH + h;
        if (j < 2) {
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) pk[i] = pack_bf16x2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
          uint16_t* dst = (j == 0 ? p.q_out : p.k_out) + (plane * p.S_pad + s) * kD;
          uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
          for (int i = 0; i < 4; ++i) d4[i] = make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
          if (j == 1 && h == 0 && p.k0_out) {
            uint4* c4 = reinterpret_cast<uint4*>(p.k0_out + (bt * p.S_pad + s) * kD);
#pragma unroll
            for (int i = 0; i < 4; ++i) c4[i] = make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
          }
        } else {
          // V is stored transposed ([d][s]) so that P V is a K-major x K-major MMA; consecutive
          // lanes hold consecutive s -> 64 B coalesced per d
          uint16_t* dst = p.vt_out + plane * kD * p.S_pad + s;
          uint16_t* dst0 = (h == 0 && p.vt0_out) ? p.vt0_out + bt * kD * p.S_pad + s : nullptr;
#pragma unroll
          for (int d = 0; d < kD; ++d) {
            __nv_bfloat16 bv = __float2bfloat16_rn(__uint_as_float(v[d]));
            const uint16_t bits = *reinterpret_cast<uint16_t*>(&bv);
            dst[(long long)d * p.S_pad] = bits;
            if (dst0) dst0[(long long)d * p.S_pad] = bits;
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, G_TMEM_COLS);
  }
}

template <int EPI>
int launch_gemm_t(const CUtensorMap& ma, const CUtensorMap& mw, const GemmArgs& a, dim3 grid, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(tc_gemm_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, G_SMEM);
    configured = true;
  }
  tc_gemm_kernel<EPI><<<grid, G_THREADS, G_SMEM, st>>>(ma, mw, a);
  return count_launch();
}

}  // namespace

int launch_tc_gemm(const TcGemm& p, cudaStream_t st) {
  if (p.K % G_BK != 0 || p.N % G_BN != 0) {
    set_error("tc_gemm: N=%d must be a multiple of %d and K=%d of %d", p.N, G_BN, p.K, G_BK);
    return MMPFN_EUNSUPPORTED;
  }
  if (p.epi == TC_EPI_RESID_LN && p.N != kE) { set_error("tc_gemm: LN epilogue needs N=%d", kE); return MMPFN_EINVAL; }
  GemmArgs a{};
  a.M = p.M; a.N = p.N; a.K = p.K; a.items = p.items; a.B = p.B; a.S = p.S; a.T = p.T;
  a.out_bf16 = p.out_bf16; a.resid = p.resid_f32; a.ln_bf16 = p.ln_bf16;
  a.q_out = p.q_out; a.k_out = p.k_out; a.vt_out = p.vt_out; a.k0_out = p.k0_out; a.vt0_out = p.vt0_out;
  a.S_pad = p.S_pad;
  CUtensorMap ma, mw;
  dim3 grid;
  if (p.items) {
    if (p.epi != TC_EPI_QKV_ITEMS) { set_error("tc_gemm: item tiles need the QKV epilogue"); return MMPFN_EINVAL; }
    a.tiles_s = (p.S + G_BM - 1) / G_BM;
    const cuuint64_t dims[4] = {(cuuint64_t)p.K, (cuuint64_t)p.T, (cuuint64_t)p.S, (cuuint64_t)p.B};
    const cuuint64_t strides[3] = {(cuuint64_t)p.K * 2, (cuuint64_t)p.T * p.K * 2, (cuuint64_t)p.S * p.T * p.K * 2};
    const cuuint32_t box[4] = {G_BK, 1, G_BM, 1};
    MMPFN_TRY(encode_map(&ma, p.A, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B));
    grid = dim3((unsigned)(p.B * p.T * a.tiles_s) * (p.N / G_BN));
  } else {
    if (p.M <= 0) return MMPFN_OK;
    const cuuint64_t dims[2] = {(cuuint64_t)p.K, (cuuint64_t)p.M};
    const cuuint64_t strides[1] = {(cuuint64_t)p.K * 2};
    const cuuint32_t box[2] = {G_BK, G_BM};
    MMPFN_TRY(encode_map(&ma, p.A, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B));
    grid = dim3((unsigned)((p.M + G_BM - 1) / G_BM) * (p.N / G_BN));
  }
  {
    const cuuint64_t dims[2] = {(cuuint64_t)p.K, (cuuint64_t)p.N};
    const cuuint64_t strides[1] = {(cuuint64_t)p.K * 2};
    const cuuint32_t box[2] = {G_BK, G_BN};
    MMPFN_TRY(encode_map(&mw, p.W, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B));
  }
  switch (p.epi) {
    case TC_EPI_BF16: return launch_gemm_t<TC_EPI_BF16>(ma, mw, a, grid, st);
    case TC_EPI_GELU_BF16: return launch_gemm_t<TC_EPI_GELU_BF16>(ma, mw, a, grid, st);
    case TC_EPI_RESID_LN: return launch_gemm_t<TC_EPI_RESID_LN>(ma, mw, a, grid, st);
    case TC_EPI_QKV_ITEMS: return launch_gemm_t<TC_EPI_QKV_ITEMS>(ma, mw, a, grid, st);
  }
  set_error("tc_gemm: bad epilogue %d", p.epi);
  return MMPFN_EINVAL;
}

// ---------------------------------------------------------------------------------------------
// Item attention
// ---------------------------------------------------------------------------------------------
namespace {
// One CTA = one 128-query tile of one (b, t, head) plane; two CTAs per SM.  Key tiles hold 112 keys so
// that TWO S accumulators (2 x 112 fp32 columns) and O (32) fit the 256 TMEM columns a CTA may take
// with two CTAs per SM.  With S double buffered, S(j+2) is issued as soon as the softmax has pulled
// S(j) into registers, a whole tile before it is needed: the mbarrier round trips (try_wait wake-up,
// tcgen05.commit arrival: ~1900 cycles per tile measured with the math knocked out) leave the
// softmax path, whose warps then run tile after tile without waiting.
//   warps 0-3  softmax, thread = query row (TMEM lane quarter = warp)
//   warp 4     TMA producer (K and V^T on separate rings) + TMEM allocation
//   warp 5     MMA issue: S = Q K^T into TMEM, O += P V with P read from shared memory
constexpr int A_BQ = 128, A_BK = 112;
constexpr int A_Q_BYTES = A_BQ * kD * 2;          // 8 KB  (64B rows, 64B swizzle)
constexpr int A_K_TX = A_BK * kD * 2;             // 7 KB per K tile
constexpr int A_K_BYTES = 8192;                   // slot stride (1024-aligned)
constexpr int A_VT_BYTES = 8192;                  // 2 k-blocks x [32 d][64 keys = 128 B], 128B swizzle (48 keys used in the 2nd)
constexpr int A_P_BYTES = A_BQ * 128 * 2;         // 32 KB = 2 k-blocks x [128 rows][128 B], 128B swizzle
constexpr int A_KV_STAGES = 2;
constexpr int A_OFF_K = A_Q_BYTES;
constexpr int A_OFF_VT = A_OFF_K + A_KV_STAGES * A_K_BYTES;
constexpr int A_OFF_P = A_OFF_VT + A_KV_STAGES * A_VT_BYTES;
constexpr int A_P_STAGES = 2;
constexpr int A_OFF_BAR = A_OFF_P + A_P_STAGES * A_P_BYTES;
constexpr int A_SMEM = A_OFF_BAR + 256 + 1024;
constexpr int A_TMEM_COLS = 256;                  // S0: [0,112)  S1: [112,224)  O: [224,256)
constexpr int A_THREADS = 192;
constexpr int kAttnPolyDefault = 4;

struct AttnArgs {
  uint16_t* out;
  int T, n_q, n_kv, shared_kv, q_tiles;
};

// DBG != 0: knock-out timing experiments (results are wrong): 1 no exp, 2 no S load from TMEM,
// 4 no P store, 8 no P V MMA, 16 no S MMA, 32 no row maximum; 64 = clock64 trace of one CTA.
__device__ long long g_attn_trace[4096];
#define ATTN_TRACE(slot)                                                          \
  do {                                                                            \
    if ((DBG & 64) && blockIdx.x == 5001) g_attn_trace[(slot)] = clock64();       \
  } while (0)
template <int PN, int DBG = 0>
__global__ void __launch_bounds__(A_THREADS, 2) tc_item_attn_kernel(const __grid_constant__ CUtensorMap map_q,
                                                                 const __grid_constant__ CUtensorMap map_k,
                                                                 const __grid_constant__ CUtensorMap map_vt,
                                                                 const AttnArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + A_OFF_BAR);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;         // [2]  K slot of tile j is free once S(j) has been computed,
  uint64_t* k_empty = bars + 3;        // [2]  V^T slot once P V(j) has
  uint64_t* v_full = bars + 5;         // [2]
  uint64_t* v_empty = bars + 7;        // [2]
  uint64_t* s_full = bars + 9;         // [2]  S(j) is in TMEM buffer j & 1
  uint64_t* s_free = bars + 11;        // [2]  ... and has been pulled into the softmax registers
  uint64_t* p_full = bars + 13;        // [2]  P(j) is in shared memory buffer j & 1.  Per buffer: a warp may run one
                                       //      tile ahead of the slowest one, and its arrival must not count for it
  uint64_t* pv_done = bars + 15;       // [2]  P V(j) has completed -> pv_done[j & 1].  The softmax only looks at it when
                                       //      it has to (rescale, final read); with one barrier per parity of j a
                                       //      parity wait stays unambiguous although phases go unobserved
  uint32_t* tmem_slot = (uint32_t*)(bars + 17);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int plane = blockIdx.x / p.q_tiles;     // (b*T + t)*kH + h; q tiles of a plane are adjacent CTAs
  const int q0 = (blockIdx.x % p.q_tiles) * A_BQ;
  const int h = plane % kH;
  const int bt = plane / kH;
  const int kv_plane = p.shared_kv ? bt : plane;
  const int nkt = (p.n_kv + A_BK - 1) / A_BK;

  if (threadIdx.x == 0) {
    prefetch_tmap(&map_q);
    prefetch_tmap(&map_k);
    prefetch_tmap(&map_vt);
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
      mbar_init(&s_full[s], 1);
      mbar_init(&s_free[s], 128);
      mbar_init(&p_full[s], 128);
      mbar_init(&pv_done[s], 1);
    }
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(tmem_slot, A_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tmem_o = tmem + 2 * A_BK;

  if (warp == 4) {
    if (lane == 0) {
      mbar_expect_tx(q_full, A_Q_BYTES);
      tma_load_3d(smem, &map_q, q_full, 0, q0, plane);
      auto load_k = [&](int j) {
        const int s = j & 1;
        mbar_wait(&k_empty[s], ((j >> 1) & 1) ^ 1);
        mbar_expect_tx(&k_full[s], A_K_TX);
        tma_load_3d(smem + A_OFF_K + s * A_K_BYTES, &map_k, &k_full[s], 0, j * A_BK, kv_plane);
      };
      auto load_v = [&](int j) {
        const int s = j & 1;
        mbar_wait(&v_empty[s], ((j >> 1) & 1) ^ 1);
        mbar_expect_tx(&v_full[s], A_VT_BYTES);
        tma_load_3d(smem + A_OFF_VT + s * A_VT_BYTES, &map_vt, &v_full[s], j * A_BK, 0, kv_plane);
        tma_load_3d(smem + A_OFF_VT + s * A_VT_BYTES + A_VT_BYTES / 2, &map_vt, &v_full[s], j * A_BK + 64, 0, kv_plane);
      };
      // issue order = the order in which the slots become free: S(j) is issued two tiles ahead of P V(j)
      load_k(0);
      if (nkt > 1) load_k(1);
      for (int j = 0; j < nkt; ++j) {
        if (j + 2 < nkt) load_k(j + 2);
        load_v(j);
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc(A_BQ, A_BK);
      constexpr uint32_t idesc_o = make_idesc(A_BQ, kD);
      const uint32_t sbase = smem_u32(smem);
      const uint64_t qdesc = make_desc(sbase, 512, kSw64);
      // S(j) = Q K(j)^T into TMEM buffer j & 1; completion arrives on s_full and frees the K slot
      auto issue_s = [&](int j) {
        const int s = j & 1;
        if (!(DBG & 16)) {
          const uint64_t kdesc = make_desc(sbase + A_OFF_K + s * A_K_BYTES, 512, kSw64);
#pragma unroll
          for (int k = 0; k < kD / 16; ++k)
            umma_bf16(tmem + s * A_BK, qdesc + (uint64_t)(k * 2), kdesc + (uint64_t)(k * 2), idesc_s, k != 0);
        }
        umma_commit(&k_empty[s]);
      };
      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      issue_s(0);
      umma_commit(&s_full[0]);
      if (nkt > 1) {
        mbar_wait(&k_full[1], 0);
        tc_fence_after();
        issue_s(1);
        umma_commit(&s_full[1]);
      }
      for (int j = 0; j < nkt; ++j) {
        const int s = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        if (j + 2 < nkt) {
          // the TMA wait first: it is long satisfied and must not sit behind the softmax hand-off
          mbar_wait(&k_full[s], ph ^ 1);         // K(j+2): completion (j+2)/2 of slot s
          mbar_wait(&s_free[s], ph);             // S(j) is in the softmax registers: its columns are free
          tc_fence_after();
          ATTN_TRACE(2048 + j * 4 + 0);
          issue_s(j + 2);
          ATTN_TRACE(2048 + j * 4 + 1);
        }
        mbar_wait(&v_full[s], ph);
        mbar_wait(&p_full[s], ph);               // P(j) is in shared memory
        tc_fence_after();
        ATTN_TRACE(2048 + j * 4 + 2);
        if (!(DBG & 8)) {
          const uint64_t pdesc = make_desc(sbase + A_OFF_P + s * A_P_BYTES, 1024, kSw128);
          const uint64_t vdesc = make_desc(sbase + A_OFF_VT + s * A_VT_BYTES, 1024, kSw128);
#pragma unroll
          for (int k = 0; k < A_BK / 16; ++k)
            umma_bf16(tmem_o, pdesc + (uint64_t)((k / 4) * (A_P_BYTES >> 5) + (k % 4) * 2),
                      vdesc + (uint64_t)((k / 4) * (A_VT_BYTES >> 5) + (k % 4) * 2), idesc_o, (j | k) != 0);
        }
        // ONE arrival tells the softmax both that S(j+2) is in TMEM buffer s and that P V(j) has
        // released P buffer s (tcgen05 ops complete in issue order), so its loop waits once per tile
        if (j + 2 < nkt) umma_commit(&s_full[s]);
        umma_commit(&v_empty[s]);
        umma_commit(&pv_done[s]);
        ATTN_TRACE(2048 + j * 4 + 3);
      }
    }
  } else {
    // ---- softmax warps: thread = one query row ----
    const int r = warp * 32 + lane;
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
    const float c = 0.17677669529663687f * 1.4426950408889634f;   // log2(e)/sqrt(d)
    const int rsw = r & 7;
    const uint32_t prow_s = smem_u32(smem) + A_OFF_P + r * 128;
    // m_ref is the score the exponent of this row is measured from.  It only follows the running
    // maximum when that has grown by more than kTau (log2 units): p = 2^((s - m_ref) c) then stays
    // below 2^kTau, which fp32 sums and bf16 P hold without loss, and the round trip that rescales O
    // in TMEM (needed on nearly every tile otherwise) becomes rare after the first tiles.
    constexpr float kTau = 8.0f;
    constexpr int kNP = A_BK / 2;                // 56 pairs of keys per row and tile
    constexpr int kAhead = 4;                    // pairs whose scaled argument is ready ahead of their ex2
    constexpr int kBehind = 5;                   // pairs whose ex2 is in flight before the first consumer reads one
    float m_ref = -INFINITY, l_run = 0.f;
    uint32_t s0[32], s1[32], s2[32], s3[16];
    // element e of the row (compile-time index after unrolling)
#define S_AT(e) ((e) < 32 ? s0[(e) & 31] : (e) < 64 ? s1[(e) & 31] : (e) < 96 ? s2[(e) & 31] : s3[(e) & 15])
    for (int j = 0; j < nkt; ++j) {
      const int sb = j & 1;
      const uint32_t tmem_s = tmem + sb * A_BK + lane_off;
      const uint32_t pbuf = prow_s + sb * A_P_BYTES;
      const int valid = p.n_kv - j * A_BK;       // keys of this tile that exist
      if (threadIdx.x == 0) ATTN_TRACE(j * 8 + 0);
      mbar_wait(&s_full[sb], (j >> 1) & 1);      // S(j) is in TMEM and P buffer sb is free (P V(j-2) done)
      tc_fence_after();
      if (threadIdx.x == 0) ATTN_TRACE(j * 8 + 1);
      // the whole 112-key row of S into registers (masking the keys a partial last tile does not have)
      auto load_row = [&]() {
        tmem_ld32(tmem_s + 0, s0);
        tmem_ld32(tmem_s + 32, s1);
        tmem_ld32(tmem_s + 64, s2);
        tmem_ld16(tmem_s + 96, s3);
        tmem_ld_wait();
        if (valid < A_BK) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            if (i >= valid) s0[i] = 0xff800000u;
            if (32 + i >= valid) s1[i] = 0xff800000u;
            if (64 + i >= valid) s2[i] = 0xff800000u;
            if (i < 16 && 96 + i >= valid) s3[i & 15] = 0xff800000u;
          }
        }
      };
      // One software-pipelined sweep over the row, written so that a lone warp keeps the MUFU pipe fed:
      // step k scales pair k+kAhead (FFMA2, and folds it into the row maximum, FMNMX3), starts the two
      // ex2 of pair k, and retires pair k-kBehind (row sum FADD2, bf16 pack F2FP, every fourth pair a
      // 16-byte store into the 128B-swizzled A tile of the PV MMA: 8 keys = one chunk of the row's
      // 128 B k-block line, chunk index XOR (row & 7)).  Values are transformed in place: s -> x -> p.
      float mxa, mxb;
      const uint64_t c2 = pack_f32x2(c, c);
      auto sweep = [&](float mc) -> float {
        const uint64_t nmc2 = pack_f32x2(-mc, -mc);
        uint64_t lsum2 = 0ull;                                // (0.f, 0.f)
        uint32_t pk[4];
        mxa = -INFINITY;
        mxb = -INFINITY;
        auto scale = [&](int k) {
          const float sa = __uint_as_float(S_AT(2 * k)), sb2 = __uint_as_float(S_AT(2 * k + 1));
          if (!(DBG & 32)) {
            if (k & 1) mxb = fmaxf(mxb, fmaxf(sa, sb2));
            else mxa = fmaxf(mxa, fmaxf(sa, sb2));
          }
          float xa, xb;
          unpack_f32x2(fma_f32x2(pack_f32x2(sa, sb2), c2, nmc2), xa, xb);
          S_AT(2 * k) = __float_as_uint(xa);
          S_AT(2 * k + 1) = __float_as_uint(xb);
        };
        auto expo = [&](int k) {
          const float xa = __uint_as_float(S_AT(2 * k)), xb = __uint_as_float(S_AT(2 * k + 1));
          const float a = (DBG & 1) ? xa : poly_sel((2 * k) & 31, PN) ? poly_exp2(xa) : fast_exp2(xa);
          const float b = (DBG & 1) ? xb : poly_sel((2 * k + 1) & 31, PN) ? poly_exp2(xb) : fast_exp2(xb);
          S_AT(2 * k) = __float_as_uint(a);
          S_AT(2 * k + 1) = __float_as_uint(b);
        };
        auto retire = [&](int k) {
          const float a = __uint_as_float(S_AT(2 * k)), b = __uint_as_float(S_AT(2 * k + 1));
          lsum2 = add_f32x2(lsum2, pack_f32x2(a, b));
          pk[k & 3] = pack_bf16x2(a, b);
          if ((k & 3) == 3) {
            const int c16 = k >> 2;                           // 16-byte chunk of the row: keys 8*c16 .. +7
            const uint32_t kb = pbuf + (c16 >> 3) * (A_P_BYTES / 2);
            if (!(DBG & 4)) st_shared_v4(kb + (((c16 & 7) ^ rsw) << 4), pk[0], pk[1], pk[2], pk[3]);
            else if (pk[0] == 0x12345678u) l_run += 1.f;     // keep the values alive
          }
        };
#pragma unroll
        for (int k = 0; k < kAhead; ++k) scale(k);
#pragma unroll
        for (int k = 0; k < kNP + kBehind; ++k) {
          if (k + kAhead < kNP) scale(k + kAhead);
          if (k < kNP) expo(k);
          if (k >= kBehind) retire(k - kBehind);
        }
        float lsum0, lsum1;
        unpack_f32x2(lsum2, lsum0, lsum1);
        return lsum0 + lsum1;
      };
      if (!(DBG & 2)) load_row();
      if (threadIdx.x == 0) ATTN_TRACE(j * 8 + 2);
      if (j == 0) {
        // first tile: the reference is the true maximum of the tile
        float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          mx0 = fmaxf(mx0, fmaxf(__uint_as_float(s0[i]), __uint_as_float(s0[i + 1])));
          mx1 = fmaxf(mx1, fmaxf(__uint_as_float(s1[i]), __uint_as_float(s1[i + 1])));
          mx2 = fmaxf(mx2, fmaxf(__uint_as_float(s2[i]), __uint_as_float(s2[i + 1])));
          if (i < 16) mx3 = fmaxf(mx3, fmaxf(__uint_as_float(s3[i & 15]), __uint_as_float(s3[(i + 1) & 15])));
        }
        m_ref = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
      }
      if (threadIdx.x == 0) ATTN_TRACE(j * 8 + 3);
      // later tiles: exponentials are taken against the reference of the previous tile while the
      // maximum is still being found; only when a row's maximum then turns out to have grown by more
      // than kTau is the sweep repeated (rare: the running maximum of n keys moves ~ log n times).
      // S was consumed in place, so the repeat reads it from TMEM again: the S columns are handed
      // back to the MMA warp only after the decision (S is double buffered: no one is waiting).
      float lsum = sweep(m_ref * c);
      const float mx = fmaxf(mxa, mxb);
      const bool moved = (mx - m_ref) * c > kTau;
      const bool any_moved = __any_sync(0xffffffffu, moved);   // tcgen05.ld is warp-collective
      float alpha = 1.0f;
      if (any_moved) {
        if (moved) {
          alpha = fast_exp2((m_ref - mx) * c);
          m_ref = mx;
        }
        load_row();
        lsum = sweep(m_ref * c);
      }
      tc_fence_before();
      mbar_arrive(&s_free[sb]);
      l_run = l_run * alpha + lsum;
      if (threadIdx.x == 0) ATTN_TRACE(j * 8 + 4);
      // rescale the running output when some row of this warp moved its reference: needs P V(j-1)
      if (any_moved) {
        mbar_wait(&pv_done[(j - 1) & 1], ((j - 1) >> 1) & 1);   // j >= 1 here: tile 0 never moves
        tc_fence_after();
        tmem_ld32(tmem_o + lane_off, s0);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) s0[i] = __float_as_uint(__uint_as_float(s0[i]) * alpha);
        tmem_st32(tmem_o + lane_off, s0);
        tmem_st_wait();
      }
      if (threadIdx.x == 0) ATTN_TRACE(j * 8 + 6);
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(&p_full[sb]);
      if (threadIdx.x == 0) ATTN_TRACE(j * 8 + 7);
    }
#undef S_AT
    uint32_t (&v)[32] = s0;
    mbar_wait(&pv_done[(nkt - 1) & 1], ((nkt - 1) >> 1) & 1);   // tcgen05 ops complete in order: covers all P V
    tc_fence_after();
    tmem_ld32(tmem_o + lane_off, v);
    tmem_ld_wait();
    const int qi = q0 + r;
    if (qi < p.n_q) {
      const float inv = 1.0f / l_run;
      const int b = bt / p.T, t = bt % p.T;
      uint16_t* dst = p.out + (((long long)b * p.n_q + qi) * p.T + t) * kE + h * kD;
      uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
      for (int i = 0; i < 4; ++i)
        d4[i] = make_uint4(pack_bf16x2(__uint_as_float(v[8 * i]) * inv, __uint_as_float(v[8 * i + 1]) * inv),
                           pack_bf16x2(__uint_as_float(v[8 * i + 2]) * inv, __uint_as_float(v[8 * i + 3]) * inv),
                           pack_bf16x2(__uint_as_float(v[8 * i + 4]) * inv, __uint_as_float(v[8 * i + 5]) * inv),
                           pack_bf16x2(__uint_as_float(v[8 * i + 6]) * inv, __uint_as_float(v[8 * i + 7]) * inv));
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem, A_TMEM_COLS);
  }
}
}  // namespace

int launch_tc_item_attn(const TcItemAttn& p, cudaStream_t st) {
  if (p.n_q <= 0 || p.B <= 0) return MMPFN_OK;
  if (p.n_kv <= 0) { set_error("item attention: empty key set"); return MMPFN_EINVAL; }
  const long long planes_q = (long long)p.B * p.T * kH;
  const long long planes_kv = p.shared_kv ? (long long)p.B * p.T : planes_q;
  const int q_tiles = (p.n_q + A_BQ - 1) / A_BQ;
  if (planes_q * q_tiles > 2147483647LL) { set_error("item attention: grid too large"); return MMPFN_EUNSUPPORTED; }
  CUtensorMap mq, mk, mvt;
  {
    const cuuint64_t dims[3] = {(cuuint64_t)kD, (cuuint64_t)p.n_q, (cuuint64_t)planes_q};
    const cuuint64_t strides[2] = {(cuuint64_t)kD * 2, (cuuint64_t)p.Sq_pad * kD * 2};
    const cuuint32_t box[3] = {kD, A_BQ, 1};
    MMPFN_TRY(encode_map(&mq, p.q, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_64B));
  }
  {
    const cuuint64_t dims[3] = {(cuuint64_t)kD, (cuuint64_t)p.n_kv, (cuuint64_t)planes_kv};
    const cuuint64_t strides[2] = {(cuuint64_t)kD * 2, (cuuint64_t)p.Skv_pad * kD * 2};
    const cuuint32_t box[3] = {kD, A_BK, 1};
    MMPFN_TRY(encode_map(&mk, p.k, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_64B));
  }
  {
    const cuuint64_t dims[3] = {(cuuint64_t)p.n_kv, (cuuint64_t)kD, (cuuint64_t)planes_kv};
    const cuuint64_t strides[2] = {(cuuint64_t)p.Skv_pad * 2, (cuuint64_t)p.Skv_pad * kD * 2};
    const cuuint32_t box[3] = {64, kD, 1};
    MMPFN_TRY(encode_map(&mvt, p.vt, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B));
  }
  // MMPFN_ATTN_POLY (0..16, read once): how many of every 32 exponentials leave the MUFU pipe for
  // the FMA-pipe polynomial; the default is the measured optimum.
  static int poly = -1, dbg = 0;
  if (poly < 0) {
    const char* e = getenv("MMPFN_ATTN_POLY");
    poly = e ? atoi(e) : kAttnPolyDefault;
    e = getenv("MMPFN_ATTN_DBG");
    dbg = e ? atoi(e) : 0;
  }
  AttnArgs a{p.out, p.T, p.n_q, p.n_kv, p.shared_kv, q_tiles};
  const dim3 grid((unsigned)(planes_q * q_tiles));
#define MMPFN_ATTN_LAUNCH(PN, DBG)                                                                          \
  do {                                                                                                      \
    cudaFuncSetAttribute(tc_item_attn_kernel<PN, DBG>, cudaFuncAttributeMaxDynamicSharedMemorySize, A_SMEM); \
    tc_item_attn_kernel<PN, DBG><<<grid, A_THREADS, A_SMEM, st>>>(mq, mk, mvt, a);                            \
  } while (0)
  if (dbg) {
    switch (dbg) {
      case 1: MMPFN_ATTN_LAUNCH(0, 1); break;
      case 2: MMPFN_ATTN_LAUNCH(0, 2); break;
      case 4: MMPFN_ATTN_LAUNCH(0, 4); break;
      case 8: MMPFN_ATTN_LAUNCH(0, 8); break;
      case 16: MMPFN_ATTN_LAUNCH(0, 16); break;
      case 32: MMPFN_ATTN_LAUNCH(0, 32); break;
      case 33: MMPFN_ATTN_LAUNCH(0, 33); break;
      case 39: MMPFN_ATTN_LAUNCH(0, 39); break;
      case 24: MMPFN_ATTN_LAUNCH(0, 24); break;
      case 63: MMPFN_ATTN_LAUNCH(0, 63); break;
      case 64: MMPFN_ATTN_LAUNCH(0, 64); break;
      default: set_error("unknown MMPFN_ATTN_DBG"); return MMPFN_EINVAL;
    }
    return count_launch();
  }
  switch (poly) {
    case 4: MMPFN_ATTN_LAUNCH(4, 0); break;
    case 8: MMPFN_ATTN_LAUNCH(8, 0); break;
    case 12: MMPFN_ATTN_LAUNCH(12, 0); break;
    default: MMPFN_ATTN_LAUNCH(0, 0); break;
  }
#undef MMPFN_ATTN_LAUNCH
  return count_launch();
}

}  // namespace mmpfn

// debug: copy the clock64 trace of the traced CTA to the host (MMPFN_ATTN_DBG=64 runs)
extern "C" int mmpfn_debug_attn_trace(long long* host_out, int n) {
  if (n > 4096) n = 4096;
  return cudaMemcpyFromSymbol(host_out, mmpfn::g_attn_trace, sizeof(long long) * n) == cudaSuccess ? 0 : MMPFN_ECUDA;
}
