// kernels_attn.cu — attention across items (layer.py:341-379, multi_head_attention.py:547-736) as a
// flash-attention kernel written directly against sm_100a, d = 32:
//
//   S = Q K^T and O += P V on tcgen05 (accumulators in TMEM, operands staged by TMA), online softmax
//   in fp32 by four warps (thread = query row).  P goes back into tensor memory over the S columns it
//   came from (bf16 pairs) and the second MMA reads its A operand there (PT = true, the default); the
//   older hand-off through a 128B-swizzled shared-memory tile is kept as PT = false for A/B timing.
//
// At d = 32 every exponential buys only 128 tensor FLOP, so the kernel lives on the MUFU/FMA pipes
// (16 ex2/clk/SM = 595 TFLOP/s at 1965 MHz), not on the tensor pipe: everything here is arranged so
// that the softmax warps never wait (S double buffered in TMEM, no row maximum in the inner loop,
// per-parity barriers) and that enough of them are resident to keep the MUFU queue full.
//
// Two tile shapes (template BK = keys per tile):
//   BK = 48,  4 CTAs/SM   TMEM: S0 [0,48)  S1 [48,96)   O [96,128)   — the default: four softmax warps
//                         per scheduler (460 TFLOP/s at the cfg2 train shape)
//   BK = 112, 2 CTAs/SM   TMEM: S0 [0,112) S1 [112,224) O [224,256)  — MMPFN_ATTN_BK=112 (411 TFLOP/s)
//
// Test pass (every query head of a column reads that column's head-0 K/V): the six heads are stacked on
// the tile's row axis, so 6 x 300 query rows fill 15 tiles of 128 instead of 18.
//
// Every mbarrier wait is bounded (trap after ~2 s) so that a protocol bug cannot hang the GPU.
#include "tc_common.cuh"

namespace mmpfn {
namespace {

constexpr int A_BQ = 128;
constexpr int A_Q_BYTES = A_BQ * kD * 2;          // 8 KB  (64B rows, 64B swizzle)
constexpr int A_THREADS = 192;
constexpr int kAttnPolyDefault = 4;

template <int BK, bool PT = false>
struct AttnCfg {
  static constexpr int kNKB = (BK + 63) / 64;                 // 64-key k-blocks of the P and V^T tiles
  static constexpr int kKTx = BK * kD * 2;                    // bytes of one K tile (64 B per key)
  static constexpr int kKSlot = BK > 64 ? 8192 : kKTx;          // slot stride: a multiple of 512 (the 64B-swizzle atom)
  static constexpr int kVtBytes = kNKB * kD * 128;            // [32 d][64 keys = 128 B] per k-block, 128B swizzle
  static constexpr int kPBytes = PT ? 0 : kNKB * A_BQ * 128;  // [128 rows][128 B] per k-block, 128B swizzle (P in TMEM: none)
  static constexpr int kOffK = A_Q_BYTES;
  static constexpr int kOffVt = kOffK + 2 * kKSlot;
  static constexpr int kOffP = (kOffVt + 2 * kVtBytes + 1023) / 1024 * 1024;
  static constexpr int kOffBar = kOffP + 2 * kPBytes;
  static constexpr int kSmem = kOffBar + 256 + 1024;
  static constexpr int kTmemCols = 2 * BK + kD <= 128 ? 128 : 256;
  static constexpr int kMinCtas = 512 / kTmemCols;
  static_assert(2 * BK + kD <= 256, "two S accumulators and O must fit the CTA's TMEM share");
  static_assert(BK % 16 == 0 && kKSlot % 512 == 0, "tile shape");
};

struct AttnArgs {
  uint16_t* out;
  int T, n_q, n_kv, shared_kv, q_tiles;
  int stack_rows;   // > 0: the six query heads of a (b, t) column are stacked on the tile's row axis, Sq_pad rows each
};

// DBG != 0: knock-out timing experiments (results are wrong): 1 no exp, 2 no S load from TMEM,
// 4 no P store, 8 no P V MMA, 16 no S MMA, 32 no row maximum; 64 = clock64 trace of one CTA.
__device__ long long g_attn_trace[4096];
#define ATTN_TRACE(slot)                                                          \
  do {                                                                            \
    if ((DBG & 64) && blockIdx.x == 5001) g_attn_trace[(slot)] = clock64();       \
  } while (0)

// n consecutive fp32 columns of this warp's 32 TMEM lanes into v[0..n)
template <int N>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t* v) {
  static_assert(N % 16 == 0, "columns come in chunks of 16");
#pragma unroll
  for (int c = 0; c + 32 <= N; c += 32) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[c + 0]), "=r"(v[c + 1]), "=r"(v[c + 2]), "=r"(v[c + 3]), "=r"(v[c + 4]), "=r"(v[c + 5]),
          "=r"(v[c + 6]), "=r"(v[c + 7]), "=r"(v[c + 8]), "=r"(v[c + 9]), "=r"(v[c + 10]), "=r"(v[c + 11]),
          "=r"(v[c + 12]), "=r"(v[c + 13]), "=r"(v[c + 14]), "=r"(v[c + 15]), "=r"(v[c + 16]), "=r"(v[c + 17]),
          "=r"(v[c + 18]), "=r"(v[c + 19]), "=r"(v[c + 20]), "=r"(v[c + 21]), "=r"(v[c + 22]), "=r"(v[c + 23]),
          "=r"(v[c + 24]), "=r"(v[c + 25]), "=r"(v[c + 26]), "=r"(v[c + 27]), "=r"(v[c + 28]), "=r"(v[c + 29]),
          "=r"(v[c + 30]), "=r"(v[c + 31])
        : "r"(taddr + c)
        : "memory");
  }
  if (N % 32 == 16) {
    constexpr int c = N - 16;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[c + 0]), "=r"(v[c + 1]), "=r"(v[c + 2]), "=r"(v[c + 3]), "=r"(v[c + 4]), "=r"(v[c + 5]),
          "=r"(v[c + 6]), "=r"(v[c + 7]), "=r"(v[c + 8]), "=r"(v[c + 9]), "=r"(v[c + 10]), "=r"(v[c + 11]),
          "=r"(v[c + 12]), "=r"(v[c + 13]), "=r"(v[c + 14]), "=r"(v[c + 15])
        : "r"(taddr + c)
        : "memory");
  }
}

// n consecutive 32-bit columns of this warp's 32 TMEM lanes from v[0..n), n a multiple of 8
template <int N>
__device__ __forceinline__ void tmem_st_cols(uint32_t taddr, const uint32_t* v) {
  static_assert(N % 8 == 0, "columns go in chunks of 8");
  int c = 0;
#pragma unroll
  for (; c + 16 <= N; c += 16) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr + c), "r"(v[c + 0]), "r"(v[c + 1]), "r"(v[c + 2]), "r"(v[c + 3]), "r"(v[c + 4]), "r"(v[c + 5]),
          "r"(v[c + 6]), "r"(v[c + 7]), "r"(v[c + 8]), "r"(v[c + 9]), "r"(v[c + 10]), "r"(v[c + 11]), "r"(v[c + 12]),
          "r"(v[c + 13]), "r"(v[c + 14]), "r"(v[c + 15])
        : "memory");
  }
  if (N % 16 == 8) {
    constexpr int c8 = N - 8;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr + c8), "r"(v[c8 + 0]), "r"(v[c8 + 1]), "r"(v[c8 + 2]), "r"(v[c8 + 3]), "r"(v[c8 + 4]),
                   "r"(v[c8 + 5]), "r"(v[c8 + 6]), "r"(v[c8 + 7])
                 : "memory");
  }
}

//   warps 0-3  softmax, thread = query row (TMEM lane quarter = warp)
//   warp 4     TMA producer (K and V^T on separate rings) + TMEM allocation
//   warp 5     MMA issue: S = Q K^T into TMEM, O += P V with P read from shared memory
// With S double buffered, S(j+2) is issued as soon as the softmax has pulled S(j) into registers, a
// whole tile before it is needed: the mbarrier round trips (try_wait wake-up, tcgen05.commit arrival:
// ~1900 cycles per tile measured with the math knocked out) leave the softmax path.
// PT = true: P never touches shared memory.  The softmax writes it (bf16 pairs, 24 columns for 48 keys)
// over the S columns it came from and the P V MMA reads its A operand from tensor memory: no st.shared,
// no generic->async proxy fence (a full MEMBAR per thread and tile), and the S buffer of tile j simply
// stays occupied until P V(j) has run — S(j+2) is issued right behind it by the same thread.
template <int BK, int PN, int DBG, bool PT>
__global__ void __launch_bounds__(A_THREADS, AttnCfg<BK, PT>::kMinCtas)
    tc_item_attn_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                        const __grid_constant__ CUtensorMap map_vt, const AttnArgs p) {
  using C = AttnCfg<BK, PT>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + C::kOffBar);
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
  // Train pass: plane = (b*T + t)*kH + h, q tiles of a plane are adjacent CTAs.  Test pass (all six query
  // heads read the head-0 K/V of their column, multi_head_attention.py:436-445): the heads are stacked on
  // the row axis — row = h * Sq_pad + s of column bt — so that 6 x 300 rows fill 15 tiles instead of 18.
  const int plane = blockIdx.x / p.q_tiles;
  const int q0 = (blockIdx.x % p.q_tiles) * A_BQ;
  const int h = p.stack_rows ? 0 : plane % kH;
  const int bt = p.stack_rows ? plane : plane / kH;
  const int kv_plane = p.shared_kv ? bt : plane;
  const int nkt = (p.n_kv + BK - 1) / BK;

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
      mbar_init(&s_free[s], 4);
      mbar_init(&p_full[s], 4);
      mbar_init(&pv_done[s], 1);
    }
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(tmem_slot, C::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tmem_o = tmem + 2 * BK;

  if (warp == 4) {
    if (elect_one()) {
      mbar_expect_tx(q_full, A_Q_BYTES);
      tma_load_3d(smem, &map_q, q_full, 0, q0, plane);
      auto load_k = [&](int j) {
        const int s = j & 1;
        mbar_wait(&k_empty[s], ((j >> 1) & 1) ^ 1);
        mbar_expect_tx(&k_full[s], C::kKTx);
        tma_load_3d(smem + C::kOffK + s * C::kKSlot, &map_k, &k_full[s], 0, j * BK, kv_plane);
      };
      auto load_v = [&](int j) {
        const int s = j & 1;
        mbar_wait(&v_empty[s], ((j >> 1) & 1) ^ 1);
        mbar_expect_tx(&v_full[s], C::kVtBytes);
#pragma unroll
        for (int kb = 0; kb < C::kNKB; ++kb)
          tma_load_3d(smem + C::kOffVt + s * C::kVtBytes + kb * (kD * 128), &map_vt, &v_full[s], j * BK + kb * 64, 0,
                      kv_plane);
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
    if (elect_one()) {
      constexpr uint32_t idesc_s = make_idesc(A_BQ, BK);
      constexpr uint32_t idesc_o = make_idesc(A_BQ, kD);
      const uint32_t sbase = smem_u32(smem);
      const uint64_t qdesc = make_desc(sbase, 512, kSw64);
      // S(j) = Q K(j)^T into TMEM buffer j & 1; completion arrives on s_full and frees the K slot
      auto issue_s = [&](int j) {
        const int s = j & 1;
        if (!(DBG & 16)) {
          const uint64_t kdesc = make_desc(sbase + C::kOffK + s * C::kKSlot, 512, kSw64);
#pragma unroll
          for (int k = 0; k < kD / 16; ++k)
            umma_bf16(tmem + s * BK, qdesc + (uint64_t)(k * 2), kdesc + (uint64_t)(k * 2), idesc_s, k != 0);
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
        if (!PT && j + 2 < nkt) {
          // the TMA wait first: it is long satisfied and must not sit behind the softmax hand-off
          // K(j+2): completion (j+2)/2 of slot s; S(j) is in the softmax registers: its columns are free
          mbar_wait2(&k_full[s], ph ^ 1, &s_free[s], ph);
          tc_fence_after();
          ATTN_TRACE(2048 + j * 4 + 0);
          issue_s(j + 2);
          ATTN_TRACE(2048 + j * 4 + 1);
        }
        mbar_wait2(&v_full[s], ph, &p_full[s], ph);      // V^T(j) landed; P(j) is written (shared memory / TMEM)
        tc_fence_after();
        ATTN_TRACE(2048 + j * 4 + 2);
        if (!(DBG & 8)) {
          const uint64_t vdesc = make_desc(sbase + C::kOffVt + s * C::kVtBytes, 1024, kSw128);
          if (PT) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_bf16_ts(tmem_o, tmem + s * BK + k * 8,
                           vdesc + (uint64_t)((k / 4) * ((kD * 128) >> 4) + (k % 4) * 2), idesc_o, (j | k) != 0);
          } else {
            const uint64_t pdesc = make_desc(sbase + C::kOffP + s * C::kPBytes, 1024, kSw128);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_bf16(tmem_o, pdesc + (uint64_t)((k / 4) * ((A_BQ * 128) >> 4) + (k % 4) * 2),
                        vdesc + (uint64_t)((k / 4) * ((kD * 128) >> 4) + (k % 4) * 2), idesc_o, (j | k) != 0);
          }
        }
        if (PT && j + 2 < nkt) {
          // S(j+2) overwrites the columns P(j) sits in: issued behind P V(j) (tcgen05 ops of one thread
          // execute in order)
          mbar_wait(&k_full[s], ph ^ 1);
          tc_fence_after();
          issue_s(j + 2);
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
    const uint32_t prow_s = smem_u32(smem) + C::kOffP + r * 128;
    // m_ref is the score the exponent of this row is measured from.  It only follows the running
    // maximum when that has grown by more than kTau (log2 units): p = 2^((s - m_ref) c) then stays
    // below 2^kTau, which fp32 sums and bf16 P hold without loss, and the round trip that rescales O
    // in TMEM (needed on nearly every tile otherwise) becomes rare after the first tiles.
    constexpr float kSumLimit = 256.0f;          // 2^kTau, kTau = 8
    constexpr int kNP = BK / 2;                  // pairs of keys per row and tile
    constexpr int kAhead = 4;                    // pairs whose scaled argument is ready ahead of their ex2
    constexpr int kBehind = 5;                   // pairs whose ex2 is in flight before the first consumer reads one
    float m_ref = -INFINITY, l_run = 0.f;
    bool next_ready = false;
    uint32_t sv[BK];
    for (int j = 0; j < nkt; ++j) {
      const int sb = j & 1;
      const uint32_t tmem_s = tmem + sb * BK + lane_off;
      const uint32_t pbuf = prow_s + sb * C::kPBytes;
      const int valid = p.n_kv - j * BK;         // keys of this tile that exist
      if (threadIdx.x == 0) ATTN_TRACE(j * 8 + 0);
      // S(j) is in TMEM and P buffer sb is free (P V(j-2) done).  Usually the probe made at the end of
      // the previous tile has already said so.
      if (!next_ready) mbar_wait(&s_full[sb], (j >> 1) & 1);
      tc_fence_after();
      if (threadIdx.x == 0) ATTN_TRACE(j * 8 + 1);
      // the whole row of S into registers (masking the keys a partial last tile does not have)
      auto load_row = [&]() {
        tmem_ld_cols<BK>(tmem_s, sv);
        tmem_ld_wait();
        if (valid < BK) {
#pragma unroll
          for (int i = 0; i < BK; ++i)
            if (i >= valid) sv[i] = 0xff800000u;
        }
      };
      // One software-pipelined sweep over the row, written so that a lone warp keeps the MUFU pipe fed:
      // step k scales pair k+kAhead (FFMA2), starts the two ex2 of pair k, and retires pair k-kBehind
      // (row sum FADD2, bf16 pack F2FP, every fourth pair a 16-byte store into the 128B-swizzled A
      // tile of the PV MMA: 8 keys = one chunk of the row's 128 B k-block line, chunk index XOR
      // (row & 7)).  Values are transformed in place: s -> x -> p.  No row maximum is tracked here.
      const uint64_t c2 = pack_f32x2(c, c);
      auto sweep = [&](float mc) -> float {
        const uint64_t nmc2 = pack_f32x2(-mc, -mc);
        uint64_t lsum2 = 0ull;                                // (0.f, 0.f)
        uint32_t pk[4];
        auto scale = [&](int k) {
          const float sa = __uint_as_float(sv[2 * k]), sb2 = __uint_as_float(sv[2 * k + 1]);
          float xa, xb;
          unpack_f32x2(fma_f32x2(pack_f32x2(sa, sb2), c2, nmc2), xa, xb);
          sv[2 * k] = __float_as_uint(xa);
          sv[2 * k + 1] = __float_as_uint(xb);
        };
        auto expo = [&](int k) {
          const float xa = __uint_as_float(sv[2 * k]), xb = __uint_as_float(sv[2 * k + 1]);
          const float a = (DBG & 1) ? xa : poly_sel((2 * k) & 31, PN) ? poly_exp2(xa) : fast_exp2(xa);
          const float b = (DBG & 1) ? xb : poly_sel((2 * k + 1) & 31, PN) ? poly_exp2(xb) : fast_exp2(xb);
          sv[2 * k] = __float_as_uint(a);
          sv[2 * k + 1] = __float_as_uint(b);
        };
        auto retire = [&](int k) {
          const float a = __uint_as_float(sv[2 * k]), b = __uint_as_float(sv[2 * k + 1]);
          lsum2 = add_f32x2(lsum2, pack_f32x2(a, b));
          if (PT) {
            // packed in place: slot k belonged to pair k / 2, retired no later than this one
            sv[k] = pack_bf16x2(a, b);
            return;
          }
          pk[k & 3] = pack_bf16x2(a, b);
          if ((k & 3) == 3) {
            const int c16 = k >> 2;                           // 16-byte chunk of the row: keys 8*c16 .. +7
            const uint32_t kb = pbuf + (c16 >> 3) * (A_BQ * 128);
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
      // the true maximum of the row of S held in sv (first tile, and the rare repeat below)
      auto row_max = [&]() -> float {
        float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
        for (int i = 0; i < BK; i += 8) {
          mx0 = fmaxf(mx0, fmaxf(__uint_as_float(sv[i]), __uint_as_float(sv[i + 1])));
          mx1 = fmaxf(mx1, fmaxf(__uint_as_float(sv[i + 2]), __uint_as_float(sv[i + 3])));
          mx2 = fmaxf(mx2, fmaxf(__uint_as_float(sv[i + 4]), __uint_as_float(sv[i + 5])));
          mx3 = fmaxf(mx3, fmaxf(__uint_as_float(sv[i + 6]), __uint_as_float(sv[i + 7])));
        }
        return fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
      };
      if (!(DBG & 2)) load_row();
      if (threadIdx.x == 0) ATTN_TRACE(j * 8 + 2);
      if (j == 0) m_ref = row_max();             // first tile: the reference is the true maximum of the tile
      if (threadIdx.x == 0) ATTN_TRACE(j * 8 + 3);
      // Later tiles: exponentials are taken against the reference of the earlier tiles, and no maximum
      // is computed at all: a score more than kTau (log2 units) above the reference shows up as a row
      // sum above 2^kTau (p <= sum p; inf and NaN fail the comparison too).  Only then — rare: the
      // running maximum of n keys moves ~ log n times, and a sum that merely crowds the limit just
      // refreshes the reference — is the true maximum taken and the sweep repeated.  S was consumed in
      // place, so the repeat reads it from TMEM again: the S columns are handed back to the MMA warp
      // only after the decision (S is double buffered: no one is waiting).
      float lsum = sweep(m_ref * c);
      // early probe of the next tile's S: consumed at the top of the next iteration
      const bool probe = j + 1 < nkt && mbar_test(&s_full[sb ^ 1], ((j + 1) >> 1) & 1);
      const bool moved = !(lsum <= kSumLimit) && !(DBG & 32);
      const bool any_moved = __any_sync(0xffffffffu, moved);   // tcgen05.ld is warp-collective
      float alpha = 1.0f;
      if (any_moved) {
        load_row();
        if (moved) {
          const float mx = fmaxf(row_max(), m_ref);
          alpha = fast_exp2((m_ref - mx) * c);
          m_ref = mx;
        }
        lsum = sweep(m_ref * c);
      }
      if (PT) {
        tmem_st_cols<BK / 2>(tmem_s, sv);          // P(j) over the first half of the S(j) columns
      } else {
        tc_fence_before();
        mbar_arrive_warp(&s_free[sb]);
      }
      l_run = l_run * alpha + lsum;
      if (threadIdx.x == 0) ATTN_TRACE(j * 8 + 4);
      // rescale the running output when some row of this warp moved its reference: needs P V(j-1)
      if (any_moved) {
        mbar_wait(&pv_done[(j - 1) & 1], ((j - 1) >> 1) & 1);   // j >= 1 here: tile 0 never moves
        tc_fence_after();
        uint32_t o[32];
        tmem_ld32(tmem_o + lane_off, o);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
        tmem_st32(tmem_o + lane_off, o);
        tmem_st_wait();
      }
      if (threadIdx.x == 0) ATTN_TRACE(j * 8 + 6);
      if (PT) tmem_st_wait();
      else fence_proxy_async();
      tc_fence_before();
      mbar_arrive_warp(&p_full[sb]);
      next_ready = __all_sync(0xffffffffu, probe);
      if (threadIdx.x == 0) ATTN_TRACE(j * 8 + 7);
    }
    uint32_t v[32];
    mbar_wait(&pv_done[(nkt - 1) & 1], ((nkt - 1) >> 1) & 1);   // tcgen05 ops complete in order: covers all P V
    tc_fence_after();
    tmem_ld32(tmem_o + lane_off, v);
    tmem_ld_wait();
    int qi = q0 + r, hh = h;
    if (p.stack_rows) {
      hh = qi / p.stack_rows;
      qi = hh < kH ? qi - hh * p.stack_rows : p.n_q;      // rows past the last head: nothing to store
    }
    if (qi < p.n_q) {
      const float inv = 1.0f / l_run;
      const int b = bt / p.T, t = bt % p.T;
      uint16_t* dst = p.out + (((long long)b * p.n_q + qi) * p.T + t) * kE + hh * kD;
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
    tmem_dealloc(tmem, C::kTmemCols);
  }
}

template <int BK, int PN, int DBG, bool PT>
void launch_attn_t(const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mvt, const AttnArgs& a, dim3 grid,
                   cudaStream_t st) {
  using C = AttnCfg<BK, PT>;
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(tc_item_attn_kernel<BK, PN, DBG, PT>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmem);
    configured = true;
  }
  tc_item_attn_kernel<BK, PN, DBG, PT><<<grid, A_THREADS, C::kSmem, st>>>(mq, mk, mvt, a);
}

template <int BK, bool PT>
int launch_attn_bk(const TcItemAttn& p, int poly, int dbg, cudaStream_t st) {
  const long long planes_q = (long long)p.B * p.T * kH;
  const long long planes_kv = p.shared_kv ? (long long)p.B * p.T : planes_q;
  // stacking pays when it saves tiles; the rows between n_q and Sq_pad of every head are then computed
  // (and dropped), so they should hold finite values (the QKV projection writes zeros there)
  const bool stack = p.shared_kv && (kH * p.Sq_pad + A_BQ - 1) / A_BQ < kH * ((p.n_q + A_BQ - 1) / A_BQ);
  const int q_tiles = stack ? (kH * p.Sq_pad + A_BQ - 1) / A_BQ : (p.n_q + A_BQ - 1) / A_BQ;
  const long long grid_planes = stack ? planes_kv : planes_q;
  if (grid_planes * q_tiles > 2147483647LL) { set_error("item attention: grid too large"); return MMPFN_EUNSUPPORTED; }
  CUtensorMap mq, mk, mvt;
  {
    const cuuint64_t dims[3] = {(cuuint64_t)kD, (cuuint64_t)(stack ? kH * p.Sq_pad : p.n_q),
                                (cuuint64_t)(stack ? planes_kv : planes_q)};
    const cuuint64_t strides[2] = {(cuuint64_t)kD * 2, (cuuint64_t)(stack ? kH : 1) * p.Sq_pad * kD * 2};
    const cuuint32_t box[3] = {kD, A_BQ, 1};
    MMPFN_TRY(encode_map(&mq, p.q, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_64B));
  }
  {
    const cuuint64_t dims[3] = {(cuuint64_t)kD, (cuuint64_t)p.n_kv, (cuuint64_t)planes_kv};
    const cuuint64_t strides[2] = {(cuuint64_t)kD * 2, (cuuint64_t)p.Skv_pad * kD * 2};
    const cuuint32_t box[3] = {kD, BK, 1};
    MMPFN_TRY(encode_map(&mk, p.k, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_64B));
  }
  {
    const cuuint64_t dims[3] = {(cuuint64_t)p.n_kv, (cuuint64_t)kD, (cuuint64_t)planes_kv};
    const cuuint64_t strides[2] = {(cuuint64_t)p.Skv_pad * 2, (cuuint64_t)p.Skv_pad * kD * 2};
    const cuuint32_t box[3] = {64, kD, 1};
    MMPFN_TRY(encode_map(&mvt, p.vt, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B));
  }
  AttnArgs a{p.out, p.T, p.n_q, p.n_kv, p.shared_kv, q_tiles, stack ? p.Sq_pad : 0};
  const dim3 grid((unsigned)(grid_planes * q_tiles));
  if (dbg) {
    switch (dbg) {
      case 1: launch_attn_t<BK, 0, 1, PT>(mq, mk, mvt, a, grid, st); break;
      case 8: launch_attn_t<BK, 0, 8, PT>(mq, mk, mvt, a, grid, st); break;
      case 63: launch_attn_t<BK, 0, 63, PT>(mq, mk, mvt, a, grid, st); break;
      case 64: launch_attn_t<BK, 0, 64, PT>(mq, mk, mvt, a, grid, st); break;
      default: set_error("unknown MMPFN_ATTN_DBG"); return MMPFN_EINVAL;
    }
    return count_launch();
  }
  switch (poly) {
    case 4: launch_attn_t<BK, 4, 0, PT>(mq, mk, mvt, a, grid, st); break;
    case 8: launch_attn_t<BK, 8, 0, PT>(mq, mk, mvt, a, grid, st); break;
    case 12: launch_attn_t<BK, 12, 0, PT>(mq, mk, mvt, a, grid, st); break;
    default: launch_attn_t<BK, 0, 0, PT>(mq, mk, mvt, a, grid, st); break;
  }
  return count_launch();
}

}  // namespace

int launch_tc_item_attn(const TcItemAttn& p, cudaStream_t st) {
  if (p.n_q <= 0 || p.B <= 0) return MMPFN_OK;
  if (p.n_kv <= 0) { set_error("item attention: empty key set"); return MMPFN_EINVAL; }
  // MMPFN_ATTN_POLY (0..16, read once): how many of every 32 exponentials leave the MUFU pipe for
  // the FMA-pipe polynomial; MMPFN_ATTN_BK: keys per tile (112 or 48).  Defaults = measured optimum.
  static int poly = -1, dbg = 0, bk = 48, pt = 1;
  if (poly < 0) {
    const char* e = getenv("MMPFN_ATTN_POLY");
    poly = e ? atoi(e) : kAttnPolyDefault;
    e = getenv("MMPFN_ATTN_DBG");
    dbg = e ? atoi(e) : 0;
    e = getenv("MMPFN_ATTN_BK");
    bk = e ? atoi(e) : 48;
    e = getenv("MMPFN_ATTN_PT");          // 0: P through shared memory (the older hand-off), A/B timing only
    pt = e ? atoi(e) : 1;
  }
  if (bk == 112) return pt ? launch_attn_bk<112, true>(p, poly, dbg, st) : launch_attn_bk<112, false>(p, poly, dbg, st);
  if (!pt) return launch_attn_bk<48, false>(p, poly, dbg, st);
  return launch_attn_bk<48, true>(p, poly, dbg, st);
}

}  // namespace mmpfn

// debug: copy the clock64 trace of the traced CTA to the host (MMPFN_ATTN_DBG=64 runs)
extern "C" int mmpfn_debug_attn_trace(long long* host_out, int n) {
  if (n > 4096) n = 4096;
  return cudaMemcpyFromSymbol(host_out, mmpfn::g_attn_trace, sizeof(long long) * n) == cudaSuccess ? 0 : MMPFN_ECUDA;
}
