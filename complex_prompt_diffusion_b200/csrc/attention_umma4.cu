// Fused flash-style attention with a SPLIT-ROW softmax (sm_100a): the hot self-attention of SD-1.x (d = 40, 4096 keys).
//
// Same data flow as attention_umma2.cu (two 128-row query tiles per CTA; S = Q K^T by SS-mode tcgen05.mma into TMEM; P
// packed to 16 bit in TMEM; O += P V by TS-mode tcgen05.mma; row sums through a ones row of V^T; lazy rescaling), with the
// three changes the clock64 timelines of that kernel asked for (profiles/r01_attn_timelines.txt):
//   * TWO threads per query row (each owns 64 of the 128 key columns of a block): 16 softmax warps = 4 per SM
//     sub-partition.  A lone warp in its exp phase issues a MUFU.EX2 only every ~10.7 cycles (8 would saturate the pipe);
//     with two warps per tile on the same sub-partition the pipe is shared at ~8.6 cycles per instruction, and the other
//     tile's two warps fill the gaps.  The two halves of a row agree on the block maximum through shared memory and a
//     64-thread named barrier (the warps with the same TMEM lane quarter).
//   * P in its own TMEM columns: S(j+1) is issued as soon as both halves of S(j) sit in registers, so the
//     P -> P V -> Q K^T -> S round trip (~1150 cycles, 35 % of a softmax warp's time in the aliased kernel) leaves the
//     critical path.  TMEM per tile: S 128 | P 64 | O 64 columns, i.e. head dims <= 63.
//   * ONE MMA issuer warp PER TILE: a single issuer serialises ~1500 cycles of barrier / fence / commit latency per
//     (tile, key block) and cannot keep up once the softmax no longer waits for it.
//
// Replaces the sliced einsum / softmax / einsum of cpd/models/attention.py:283-348 for head dims <= 63 (default) and, opt-in
// with P aliased over S, for head dims <= 111.
#include <stdlib.h>

#include "../../include/cpd_b200.h"
#include "common.cuh"

namespace {

constexpr int BQ = 128;
constexpr int BKV = 128;
constexpr int ATOM_BYTES = 128 * 128;  // 128 rows x 64 16-bit elements
constexpr int SPLIT = 2;               // threads per query row
constexpr int CW = BKV / SPLIT;        // key columns of a block per thread
constexpr int NUM_THREADS = 64 + 2 * SPLIT * 128 + 32;  // warp 0 TMA, warp 1 issuer of tile 0, warps 2-17 softmax (tile, half, lane quarter), warp 18 issuer of tile 1
constexpr int ISSUER1_WARP = 2 + 2 * SPLIT * 4;
constexpr int MAX_STAGES = 4;
constexpr float RESCALE_TAU = 8.0f;    // lazy O rescale threshold (log2 domain)

#ifdef CPD_TIMELINE
// Debug build (make TIMELINE=1): CTA 0 stamps clock64 at the phases of key blocks 8..15 (tools/attn4_timeline.py):
// softmax warps [tile][half][block][phase] (lane 0 of lane quarter 0), issuer warps [tile][block][phase]
__device__ long long cpd_dbg_attn4[2][2][8][8];
__device__ long long cpd_dbg_mma4[2][8][8];
#define A4_STAMP(ph)                                                                                              \
  do {                                                                                                            \
    if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && qd == 0 && lane == 0 && j >= 8 && j < 16)       \
      cpd_dbg_attn4[t][hf][j - 8][ph] = clock64();                                                                \
  } while (0)
#define M4_STAMP(ph)                                                                                              \
  do {                                                                                                            \
    if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0 && j >= 8 && j < 16)                   \
      cpd_dbg_mma4[t][j - 8][ph] = clock64();                                                                     \
  } while (0)
#else
#define A4_STAMP(ph) \
  do {               \
  } while (0)
#define M4_STAMP(ph) \
  do {               \
  } while (0)
#endif

struct Attn4Args {
  CUtensorMap map_q, map_k, map_vt;
  bf16* o;
  int ldo;
  int batch, heads, nq, nk, nk_pad, kv_batch;
  int dqk;      // padded head dim of Q / K / O columns (multiple of 16)
  int d;        // real head dim (rows of V^T loaded by TMA)
  int dv;       // MMA N of P V: round16(d + 1) (ones row + zero rows follow the d real rows)
  int datoms;   // ceil(dqk / 64)
  int stages;
  int fp16;
  float scale_log2;
};

__device__ __forceinline__ void umma_f16_ts4(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st32_4(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
      "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
      "r"(r[31])
      : "memory");
}

// SEP = true : P in its own TMEM columns (per tile S 128 | P 64 | O 64: head dims <= 63), S(j+1) runs ahead of the softmax.
// SEP = false: P overwrites S in place (per tile S/P 128, O 128 in the upper half: head dims <= 111); S(j+1) follows P V(j)
//              on the in-order tensor pipe - the split rows and the per-tile issuers still apply.
template <bool SEP, bool F16>  // F16: the activation format (fp16 / bf16) as a compile-time constant: one F2FP per packed pair
__global__ void __launch_bounds__(NUM_THREADS, 1) attention4_kernel(const __grid_constant__ Attn4Args a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int datoms = a.datoms;
  const int stages = a.stages;
  const int q_tile_bytes = datoms * ATOM_BYTES;
  const int k_stage_bytes = datoms * ATOM_BYTES;
  const int vt_atom_bytes = a.dv * 128;
  const int v_stage_bytes = 2 * vt_atom_bytes;
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + 2 * q_tile_bytes;
  uint8_t* sV = sK + stages * k_stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + stages * v_stage_bytes);
  uint64_t* q_full = bars;                   // [1]
  uint64_t* k_full = bars + 1;               // [MAX_STAGES]
  uint64_t* k_empty = k_full + MAX_STAGES;
  uint64_t* v_full = k_empty + MAX_STAGES;
  uint64_t* v_empty = v_full + MAX_STAGES;
  uint64_t* s_full = v_empty + MAX_STAGES;   // [2]
  uint64_t* p_full = s_full + 2;             // [2]
  uint64_t* pv_done = p_full + 2;            // [2]  P V_t(j) complete: P_t may be overwritten, O_t may be rescaled / read
  uint64_t* s_free = pv_done + 2;            // [2]  S_t(j) has been read into registers (separate-P mode)
  uint64_t* stagger = s_free + 2;            // [1]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(stagger + 1);
  float* xch = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 512);  // [2 parities][2 tiles][SPLIT][128] block maxima

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * (2 * BQ);
  const int head = blockIdx.y;
  const int b = blockIdx.z;
  const int bkv = b % a.kv_batch;
  const int nblk = (a.nk + BKV - 1) / BKV;
  constexpr bool f16 = F16;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&a.map_q);
    tma_prefetch_desc(&a.map_k);
    tma_prefetch_desc(&a.map_vt);
    mbar_init(q_full, 1);
    for (int s = 0; s < stages; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 2);  // one tcgen05.commit per tile issuer
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 2);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&p_full[t], 128 * SPLIT);
      mbar_init(&pv_done[t], 1);
      mbar_init(&s_free[t], 128 * SPLIT);
    }
    mbar_init(stagger, 128 * SPLIT);
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_smem, 512);
    tmem_relinquish();
  }
  // rows d .. dv-1 of every V^T stage atom: a row of ones (-> O[:, d] = sum of P), then zeros.  The TMA box never
  // touches them.  Every 16-byte chunk of a row holds the same value, so the 128B swizzle does not matter.
  {
    const uint32_t one2 = f16 ? 0x3C003C00u : 0x3F803F80u;
    const int pad_rows = a.dv - a.d;
    const int chunks = stages * 2 * pad_rows * 8;  // 16-byte chunks
    for (int i = threadIdx.x; i < chunks; i += NUM_THREADS) {
      const int c16 = i & 7;
      const int rr = (i >> 3) % pad_rows;
      const int at = (i >> 3) / pad_rows;  // stage * 2 + atom
      const uint32_t v = (rr == 0) ? one2 : 0u;
      *reinterpret_cast<uint4*>(sV + at * vt_atom_bytes + (a.d + rr) * 128 + c16 * 16) = make_uint4(v, v, v, v);
    }
    fence_proxy_async_smem();
  }
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();
  // TMEM columns.  Separate-P mode (dv <= 64, i.e. SD-1.x d = 40): per tile [S 128 | P 64 | O 64]; S(j+1) is issued as
  // soon as the softmax threads have READ S(j), so the Q K^T round trip never stalls the exp stream.  Aliased mode
  // (larger head dims): P overwrites S in place, O lives in the upper half; S(j+1) follows P V(j) on the in-order pipe.
  constexpr bool sep = SEP;
  const uint32_t colS0 = 0, colS1 = sep ? 256u : 128u;
  const uint32_t offP = sep ? 128u : 0u;                 // P_t relative to S_t
  const uint32_t colO0 = sep ? 192u : 256u, colO1 = sep ? 448u : 384u;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, 2 * q_tile_bytes);
      for (int t = 0; t < 2; ++t)
        for (int dd = 0; dd < datoms; ++dd)
          tma_load_2d(sQ + t * q_tile_bytes + dd * ATOM_BYTES, &a.map_q, q_full, head * a.dqk + dd * 64, b * a.nq + q0 + t * BQ);
      int st = 0;
      uint32_t ph = 0;
      for (int j = 0; j < nblk; ++j) {
        mbar_wait(&k_empty[st], ph ^ 1, 10);
        mbar_arrive_expect_tx(&k_full[st], k_stage_bytes);
        for (int dd = 0; dd < datoms; ++dd)
          tma_load_2d(sK + st * k_stage_bytes + dd * ATOM_BYTES, &a.map_k, &k_full[st], head * a.dqk + dd * 64,
                      bkv * a.nk_pad + j * BKV);
        mbar_wait(&v_empty[st], ph ^ 1, 11);
        mbar_arrive_expect_tx(&v_full[st], 2 * a.d * 128);
        for (int t = 0; t < 2; ++t)
          tma_load_2d(sV + st * v_stage_bytes + t * vt_atom_bytes, &a.map_vt, &v_full[st], bkv * a.nk_pad + j * BKV + t * 64,
                      head * a.dqk);
        if (++st == stages) {
          st = 0;
          ph ^= 1;
        }
      }
    }
  } else if (warp == 1 || warp == ISSUER1_WARP) {
    // ================= separate-P mode: one MMA issuer warp PER TILE =================
    // A single issuer serialises ~1500 cycles of barrier / fence / elect / commit latencies per (tile, key block) - as long as
    // a whole softmax block - and becomes the bottleneck once S(j+1) is taken off the critical path.  Each tile's issuer runs
    //   S_t(0);  for j: [S_t(j+1) once the softmax has read S_t(j)]  [P V_t(j) once P_t(j) is stored]
    // and both commit to the K / V stage-release barriers (count 2).
    const int t = warp == 1 ? 0 : 1;
    const uint32_t idesc_s = umma_idesc_f16(BQ, BKV, f16, f16);
    const uint32_t idesc_o = umma_idesc_f16(BQ, a.dv, f16, f16);
    const int ksteps_s = a.dqk / 16;
    const uint32_t qa = smem_u32(sQ) + t * q_tile_bytes, k_addr = smem_u32(sK), v_addr = smem_u32(sV);
    const uint32_t dS = tmem_base + (t ? colS1 : colS0);
    const uint32_t dO = tmem_base + (t ? colO1 : colO0);
    const uint32_t aP = dS + offP;
    auto issue_s = [&](int st) {
      if (elect_one()) {
        const uint32_t ka = k_addr + st * k_stage_bytes;
        for (int kk = 0; kk < ksteps_s; ++kk) {
          const uint32_t off = (uint32_t)((kk >> 2) * ATOM_BYTES + (kk & 3) * 32);
          umma_bf16(dS, umma_desc_sw128(qa + off), umma_desc_sw128(ka + off), idesc_s, kk > 0 ? 1u : 0u);
        }
        umma_commit(&s_full[t]);
        umma_commit(&k_empty[st]);
      }
      __syncwarp();
    };
    mbar_wait(q_full, 0, 21);
    mbar_wait(&k_full[0], 0, 20);
    if (t == 1 && nblk > 1) mbar_wait(stagger, 0, 24);
    tc_fence_after();
    issue_s(0);
    int st = 0;
    uint32_t ph = 0;
    for (int j = 0; j < nblk; ++j) {
      int st_n = st + 1;
      uint32_t ph_n = ph;
      if (st_n == stages) {
        st_n = 0;
        ph_n ^= 1;
      }
      M4_STAMP(0);
      if (sep && j + 1 < nblk) {  // S_t(j+1) as soon as both halves of S_t(j) sit in registers
        mbar_wait(&s_free[t], (uint32_t)(j & 1), 25);
        M4_STAMP(1);
        mbar_wait(&k_full[st_n], ph_n, 20);
        tc_fence_after();
        issue_s(st_n);
        M4_STAMP(2);
      }
      mbar_wait(&v_full[st], ph, 23);
      mbar_wait(&p_full[t], (uint32_t)(j & 1), 22);
      tc_fence_after();
      M4_STAMP(3);
      if (elect_one()) {
        const int kv_valid = min(BKV, a.nk - j * BKV);
        const int ksteps_o = (kv_valid + 15) / 16;
        const uint32_t va = v_addr + st * v_stage_bytes;
        for (int kk = 0; kk < ksteps_o; ++kk) {
          const uint32_t offv = (uint32_t)((kk >> 2) * vt_atom_bytes + (kk & 3) * 32);
          umma_f16_ts4(dO, aP + kk * 8, umma_desc_sw128(va + offv), idesc_o, (j > 0 || kk > 0) ? 1u : 0u);
        }
        umma_commit(&pv_done[t]);
        umma_commit(&v_empty[st]);
      }
      __syncwarp();
      M4_STAMP(4);
      if (!sep && j + 1 < nblk) {  // aliased: S_t(j+1) overwrites P_t(j), so it follows P V_t(j) on the in-order tensor pipe
        mbar_wait(&k_full[st_n], ph_n, 20);
        tc_fence_after();
        issue_s(st_n);
      }
      st = st_n;
      ph = ph_n;
    }
  } else if (warp >= 2 && warp < ISSUER1_WARP) {
    // ================= softmax: SPLIT threads per query row, each owning CW = 64 key columns of a block =================
    const int w2 = warp - 2;
    const int t = w2 / (4 * SPLIT);       // tile
    const int hf = (w2 >> 2) % SPLIT;     // which half of the key columns
    const int qd = warp & 3;              // TMEM lane quarter of this warp
    const int r = qd * 32 + lane;
    const uint32_t lane_off = (uint32_t)(qd * 32) << 16;
    const uint32_t tS = tmem_base + (t ? colS1 : colS0) + lane_off;
    const uint32_t tP = tS + offP;
    const uint32_t tO = tmem_base + (t ? colO1 : colO0) + lane_off;
    const int bar_id = 1 + t * 4 + qd;    // the SPLIT warps that hold the same 32 rows (same SM sub-partition)
    const int col0 = hf * CW;             // first key column (inside the block) of this thread
    const uint32_t xch_addr = smem_u32(xch);
    float m_used = -INFINITY;
    for (int j = 0; j < nblk; ++j) {
      const int kv_valid = min(BKV, a.nk - j * BKV);
      A4_STAMP(0);
      mbar_wait(&s_full[t], (uint32_t)(j & 1), 30);
      tc_fence_after();
      A4_STAMP(1);
      uint32_t s[CW];
      tmem_ld32(tS + col0, reinterpret_cast<uint32_t(&)[32]>(s[0]));
      tmem_ld32(tS + col0 + 32, reinterpret_cast<uint32_t(&)[32]>(s[32]));
      tmem_ld_wait();
      A4_STAMP(2);
      if (sep) {  // this half of the S row now lives in registers: once both halves do, the tensor pipe may write S_t(j+1)
        tc_fence_before();
        mbar_arrive(&s_free[t]);
      }
      const bool full = (kv_valid == BKV);
      float mx0 = -INFINITY, mx1 = -INFINITY;
      if (full) {
#pragma unroll
        for (int e = 0; e < CW; e += 4) {  // FMNMX3: two new elements per instruction, two independent chains
          mx0 = fmax3(mx0, __uint_as_float(s[e]), __uint_as_float(s[e + 1]));
          mx1 = fmax3(mx1, __uint_as_float(s[e + 2]), __uint_as_float(s[e + 3]));
        }
      } else {
#pragma unroll
        for (int e = 0; e < CW; ++e)
          if (col0 + e < kv_valid) mx0 = fmaxf(mx0, __uint_as_float(s[e]));
      }
      if (t == 0 && j == 0) mbar_arrive(stagger);
      // block maximum of the whole row: exchange the half maxima (double-buffered by block parity: a thread can be at
      // most one barrier ahead of its partner)
      float mx = fmaxf(mx0, mx1);
      {  // explicit shared-space accesses (the carved-up dynamic buffer is a generic pointer to the compiler)
        const uint32_t slot = xch_addr + (uint32_t)(((((j & 1) * 2 + t) * SPLIT) * 128 + r) * 4);
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(slot + hf * 512), "f"(mx) : "memory");
        named_bar_sync(bar_id, 32 * SPLIT);
        float other;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(other) : "r"(slot + (hf ^ 1) * 512) : "memory");
        mx = fmaxf(mx, other);
      }
      A4_STAMP(3);
      const float m_blk = mx * a.scale_log2;
      float alpha = 1.0f;
      bool need = false;
      if (j == 0) {
        m_used = m_blk;
      } else if (m_blk > m_used + RESCALE_TAU) {  // both halves see the same m_blk and m_used: same decision
        alpha = exp2f(m_used - m_blk);
        m_used = m_blk;
        need = true;
      }
      // P V_t(j-1) must be complete before O_t is rescaled (rare) and before P_t is overwritten - the latter only after the
      // exp phase, so the P V round trip hides behind it.
      bool pv_waited = !sep || j == 0;  // aliased mode: s_full(j) already implies that P V_t(j-1) has completed
      if (__any_sync(0xffffffffu, need)) {  // rare: rescale this warp's 32 rows of O (every other 16-column group per half)
        if (!pv_waited) {
          mbar_wait(&pv_done[t], (uint32_t)((j - 1) & 1), 31);
          tc_fence_after();
          pv_waited = true;
        }
        for (int c = hf * 16; c < a.dv; c += 16 * SPLIT) {
          uint32_t o16[16];
          tmem_ld16(tO + c, o16);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 16; ++e) o16[e] = __float_as_uint(__uint_as_float(o16[e]) * alpha);
          tmem_st16(tO + c, o16);
        }
      }
      // p = 2^(s * scale_log2 - m): one FFMA feeding MUFU.EX2; packed pairs overwrite s[] in place (s[e/2] <- e, e+1).
      // (One FFMA2 - sm_100 packed fp32 - per pair instead of two FFMA was measured SLOWER: 706.7 / 707.9 vs 692.5 / 691.4 us on
      // one box, 16 x 8 x 4096^2 d = 40.)
      const float neg_m = -m_used;
      if (full) {
#pragma unroll
        for (int e = 0; e < CW; e += 2) {
          const float x0 = fmaf(__uint_as_float(s[e]), a.scale_log2, neg_m);
          const float x1 = fmaf(__uint_as_float(s[e + 1]), a.scale_log2, neg_m);
          s[e >> 1] = pack_act2(fast_ex2(x0), fast_ex2(x1), f16);
        }
      } else {
#pragma unroll
        for (int e = 0; e < CW; e += 2) {
          float p0 = fast_ex2(fmaf(__uint_as_float(s[e]), a.scale_log2, neg_m));
          float p1 = fast_ex2(fmaf(__uint_as_float(s[e + 1]), a.scale_log2, neg_m));
          if (col0 + e >= kv_valid) p0 = 0.f;
          if (col0 + e + 1 >= kv_valid) p1 = 0.f;
          s[e >> 1] = pack_act2(p0, p1, f16);
        }
      }
      A4_STAMP(4);
      if (!pv_waited) {
        mbar_wait(&pv_done[t], (uint32_t)((j - 1) & 1), 31);
        tc_fence_after();
      }
      A4_STAMP(5);
      tmem_st32_4(tP + hf * (CW / 2), &s[0]);  // 16-bit P: two elements per TMEM column
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_full[t]);
      A4_STAMP(6);
    }
    // ---- epilogue: O / l -> global (l = O[:, d], accumulated by the ones row of V^T) ----
    mbar_wait(&pv_done[t], (uint32_t)((nblk - 1) & 1), 32);
    tc_fence_after();
    const int qrow = q0 + t * BQ + r;
    float inv_l;
    {
      uint32_t o16[16];
      tmem_ld16(tO + (a.d & ~15), o16);
      tmem_ld_wait();
      float l = 1.f;
#pragma unroll
      for (int e = 0; e < 16; ++e)
        if (e == (a.d & 15)) l = __uint_as_float(o16[e]);
      inv_l = 1.0f / l;
    }
    for (int c = hf * 16; c < a.dqk; c += 16 * SPLIT) {  // every other 16-column group per half
      uint32_t o16[16];
      tmem_ld16(tO + c, o16);
      tmem_ld_wait();
      if (qrow < a.nq) {
        uint32_t o[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float v0 = (c + 2 * e < a.d) ? __uint_as_float(o16[2 * e]) * inv_l : 0.f;
          const float v1 = (c + 2 * e + 1 < a.d) ? __uint_as_float(o16[2 * e + 1]) * inv_l : 0.f;
          o[e] = pack_act2(v0, v1, f16);
        }
        uint4* dst = reinterpret_cast<uint4*>(a.o + ((int64_t)b * a.nq + qrow) * a.ldo + head * a.dqk + c);
        dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
        dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

#ifdef CPD_TIMELINE
extern "C" int cpd_debug_attn4_timeline(long long* soft256, long long* mma128) {
  int rc = (int)cudaMemcpyFromSymbol(soft256, cpd_dbg_attn4, sizeof(long long) * 256);
  if (rc) return rc;
  return (int)cudaMemcpyFromSymbol(mma128, cpd_dbg_mma4, sizeof(long long) * 128);
}
#endif

// Returns CPD_ERR_UNSUPPORTED when the shape is outside this kernel's domain (the caller falls back).
cpd_status cpd_attention_split(const cpd_attn_params* p, void* stream) {
  const int d = p->d_head;
  if (d <= 0 || d > p->dpad || p->nq <= BQ) return CPD_ERR_UNSUPPORTED;
  const int dv = (d + 1 + 15) / 16 * 16;
  if (dv > 128) return CPD_ERR_UNSUPPORTED;          // O_t has at most 128 TMEM columns
  const bool sep = dv <= 64;                         // room for P next to S and O: S(j+1) runs ahead
  if (p->nk <= 2 * BKV) return CPD_ERR_UNSUPPORTED;  // few key blocks (cross-attention): the persistent kernel is the better fit
  // The aliased variant (head dims 64..111) is correct but measured SLOWER than the two-tile kernels (1024^2 d80: 93.0 vs
  // 84.6 us, 9216^2 d64: 1449 vs 1359 us): with P over S the Q K^T round trip is back on every tile's critical path and the
  // extra warps only add barrier traffic.  Opt-in: CPD_ATTN_SPLIT_ALIASED=1.
  static int alias_ok = -1;
  if (alias_ok < 0) {
    const char* e = getenv("CPD_ATTN_SPLIT_ALIASED");
    alias_ok = (e && e[0] == '1') ? 1 : 0;
  }
  if (!sep && !alias_ok) return CPD_ERR_UNSUPPORTED;
  Attn4Args a;
  a.o = (bf16*)p->o;
  a.ldo = p->ldo;
  a.batch = p->batch; a.heads = p->heads; a.nq = p->nq; a.nk = p->nk; a.nk_pad = p->nk_pad;
  a.kv_batch = p->kv_batch > 0 ? p->kv_batch : p->batch;
  a.dqk = p->dpad;
  a.d = d;
  a.dv = dv;
  a.datoms = (p->dpad + 63) / 64;
  a.fp16 = p->act_fp16;
  a.scale_log2 = p->scale * 1.4426950408889634f;
  const int q_bytes = 2 * a.datoms * ATOM_BYTES;
  const int per_stage = a.datoms * ATOM_BYTES + 2 * dv * 128;
  constexpr int XCH_BYTES = 2 * 2 * SPLIT * 128 * 4;  // block-maximum exchange buffer behind the barriers
  int stages = (227 * 1024 - 1024 - 512 - XCH_BYTES - q_bytes) / per_stage;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  if (stages < 2) return CPD_ERR_UNSUPPORTED;
  a.stages = stages;
  int rc;
  {
    uint64_t dims[2] = {(uint64_t)p->heads * p->dpad, (uint64_t)p->batch * p->nq};
    uint64_t str[1] = {(uint64_t)p->ldq * 2};
    uint32_t box[2] = {64, BQ};
    if ((rc = cpd_make_tmap_bf16(&a.map_q, p->q, 2, dims, str, box))) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)p->heads * p->dpad, (uint64_t)a.kv_batch * p->nk_pad};
    uint64_t str[1] = {(uint64_t)p->ldk * 2};
    uint32_t box[2] = {64, BKV};
    if ((rc = cpd_make_tmap_bf16(&a.map_k, p->k, 2, dims, str, box))) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)a.kv_batch * p->nk_pad, (uint64_t)p->heads * p->dpad};
    uint64_t str[1] = {(uint64_t)p->ldvt * 2};
    uint32_t box[2] = {64, (uint32_t)d};
    if ((rc = cpd_make_tmap_bf16(&a.map_vt, p->vt, 2, dims, str, box))) return rc;
  }
  const size_t shm = (size_t)q_bytes + (size_t)stages * per_stage + 512 + XCH_BYTES + 1024;
  // (Moving a share of the exponentials to an FMA-pipe polynomial was measured in round 1 - 747 / 771 / 842 us against 724 us
  // with 25 / 33 / 50 % moved - and removed: the softmax warps are bound by their own instruction stream, not by the MUFU pipe.)
  const int lim = (int)(227 * 1024);
  CPD_SMEM_OPTIN((attention4_kernel<true, true>), lim);
  CPD_SMEM_OPTIN((attention4_kernel<true, false>), lim);
  CPD_SMEM_OPTIN((attention4_kernel<false, true>), lim);
  CPD_SMEM_OPTIN((attention4_kernel<false, false>), lim);
  dim3 grid((p->nq + 2 * BQ - 1) / (2 * BQ), p->heads, p->batch);
  const dim3 block(NUM_THREADS);
  cudaStream_t st = (cudaStream_t)stream;
  const bool f16 = p->act_fp16 != 0;
  if (sep && f16) CPD_CUDA_CHECK(cpd_launch(attention4_kernel<true, true>, grid, block, shm, st, a));
  else if (sep) CPD_CUDA_CHECK(cpd_launch(attention4_kernel<true, false>, grid, block, shm, st, a));
  else if (f16) CPD_CUDA_CHECK(cpd_launch(attention4_kernel<false, true>, grid, block, shm, st, a));
  else CPD_CUDA_CHECK(cpd_launch(attention4_kernel<false, false>, grid, block, shm, st, a));
  CPD_CUDA_CHECK(cudaGetLastError());
  return CPD_OK;
}
