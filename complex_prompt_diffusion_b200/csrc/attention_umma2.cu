// Fused flash-style attention, two query tiles per CTA (sm_100a).
//
// One CTA owns 256 query rows of one (batch, head): two 128-row tiles, each with its own softmax warpgroup, sharing
// the K / V^T stream (TMA, 128B-swizzled shared memory, multi-stage).  Per tile and per 128-key block:
//   S = Q K^T            tcgen05.mma (A, B from shared memory) -> fp32 S in TMEM (128 columns)
//   softmax warpgroup    one thread per query row: ONE pass over the S row held in registers (row max, exp2 via MUFU
//                        fed by a single FFMA), P packed to 16 bit and stored back to TMEM over S (tcgen05.st)
//   O += P V             tcgen05.mma with the A operand (P) read from TENSOR MEMORY, B = V^T tile from shared memory
// While one tile's softmax runs on the MUFU/FMA pipes the tensor pipe computes the other tile's S and P V, so the
// exp2 stream (the roofline of this kernel: d = 40 gives only 160 MMA flops per exp) never waits for an MMA.
// The row sum needs no instructions: V^T carries a row of ones right after its d real rows (written once into the
// shared-memory stage buffers; the TMA box covers only the d real rows), so O[:, d] accumulates sum(P) in the MMA and
// is rescaled with O.  Lazy rescaling (threshold 2^8) keeps O corrections rare.
//
// Replaces the sliced einsum / softmax / einsum of cpd/models/attention.py:283-348 for head dims <= 112.
#include <stdlib.h>

#include "../../include/cpd_b200.h"
#include "common.cuh"

namespace {

constexpr int BQ = 128;
constexpr int BKV = 128;
constexpr int ATOM_BYTES = 128 * 128;  // 128 rows x 64 16-bit elements
constexpr int NUM_THREADS = 352;       // warp 0 TMA, warp 1 MMA, warps 2-5 softmax tile 0, warps 6-9 softmax tile 1, warp 10 second MMA issuer (separate-P mode)
constexpr int MAX_STAGES = 4;
constexpr float RESCALE_TAU = 8.0f;    // lazy O rescale threshold (log2 domain)

struct Attn2Args {
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
  int xu_gate;  // separate-P mode: alternate the two tiles' exp phases (CPD_ATTN_GATE=0 turns it off)
  float scale_log2;
};

__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* r) {
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

#ifdef CPD_TIMELINE
// Debug build (make TIMELINE=1): CTA 0, lane 0 of the first softmax warp of each tile stamps clock64 at the phases of key
// blocks 8..15 (tools/attn_timeline.py): [tile][block - 8][phase]
__device__ long long cpd_dbg_attn[2][8][8];
#define ATT_STAMP(ph)                                                                                              \
  do {                                                                                                             \
    if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && qd == 2 && lane == 0 && j >= 8 && j < 16)         \
      cpd_dbg_attn[t][j - 8][ph] = clock64();                                                                      \
  } while (0)
#else
#define ATT_STAMP(ph) \
  do {                \
  } while (0)
#endif

template <bool SEP>  // SEP: P in its own TMEM columns (dv <= 64); otherwise P overwrites S in place
__global__ void __launch_bounds__(NUM_THREADS, 1) attention2_kernel(const __grid_constant__ Attn2Args a) {
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
  uint64_t* xu_done = stagger + 1;           // [2][4] exp phase of (tile, lane quarter) finished: the MUFU pipe is free for the other tile
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(xu_done + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * (2 * BQ);
  const int head = blockIdx.y;
  const int b = blockIdx.z;
  const int bkv = b % a.kv_batch;
  const int nblk = (a.nk + BKV - 1) / BKV;
  const bool f16 = a.fp16 != 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&a.map_q);
    tma_prefetch_desc(&a.map_k);
    tma_prefetch_desc(&a.map_vt);
    mbar_init(q_full, 1);
    for (int s = 0; s < stages; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], SEP ? 2 : 1);  // separate-P mode: one tcgen05.commit per tile issuer
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], SEP ? 2 : 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&p_full[t], 128);
      mbar_init(&pv_done[t], 1);
      mbar_init(&s_free[t], 128);
    }
    mbar_init(stagger, 128);
    for (int i = 0; i < 8; ++i) mbar_init(&xu_done[i], 1);
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
  } else if (SEP && (warp == 1 || warp == 10)) {
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
      if (j + 1 < nblk) {
        mbar_wait(&s_free[t], (uint32_t)(j & 1), 25);
        mbar_wait(&k_full[st_n], ph_n, 20);
        tc_fence_after();
        issue_s(st_n);
      }
      mbar_wait(&v_full[st], ph, 23);
      mbar_wait(&p_full[t], (uint32_t)(j & 1), 22);
      tc_fence_after();
      if (elect_one()) {
        const int kv_valid = min(BKV, a.nk - j * BKV);
        const int ksteps_o = (kv_valid + 15) / 16;
        const uint32_t va = v_addr + st * v_stage_bytes;
        for (int kk = 0; kk < ksteps_o; ++kk) {
          const uint32_t offv = (uint32_t)((kk >> 2) * vt_atom_bytes + (kk & 3) * 32);
          umma_f16_ts(dO, aP + kk * 8, umma_desc_sw128(va + offv), idesc_o, (j > 0 || kk > 0) ? 1u : 0u);
        }
        umma_commit(&pv_done[t]);
        umma_commit(&v_empty[st]);
      }
      __syncwarp();
      st = st_n;
      ph = ph_n;
    }
  } else if (warp == 1) {
    // ================= MMA issuer (whole warp runs the loop; one elected lane issues) =================
    const uint32_t idesc_s = umma_idesc_f16(BQ, BKV, f16, f16);
    const uint32_t idesc_o = umma_idesc_f16(BQ, a.dv, f16, f16);
    const int ksteps_s = a.dqk / 16;
    const uint32_t q_addr = smem_u32(sQ), k_addr = smem_u32(sK), v_addr = smem_u32(sV);
    auto issue_s = [&](int t, int st) {  // S_t = Q_t K(st)^T
      if (elect_one()) {
        const uint32_t qa = q_addr + t * q_tile_bytes, ka = k_addr + st * k_stage_bytes;
        const uint32_t dS = tmem_base + (t ? colS1 : colS0);
        for (int kk = 0; kk < ksteps_s; ++kk) {
          const uint32_t off = (uint32_t)((kk >> 2) * ATOM_BYTES + (kk & 3) * 32);
          umma_bf16(dS, umma_desc_sw128(qa + off), umma_desc_sw128(ka + off), idesc_s, kk > 0 ? 1u : 0u);
        }
        umma_commit(&s_full[t]);
      }
      __syncwarp();
    };
    mbar_wait(q_full, 0, 21);
    mbar_wait(&k_full[0], 0, 20);
    tc_fence_after();
    // Stagger the two tiles by about half a block: tile 1 starts when tile 0 has finished the row-max phase of its first
    // block, so one tile's MUFU-free phases line up with the other tile's exp phase.
    issue_s(0, 0);
    if (nblk > 1) mbar_wait(stagger, 0, 24);  // a single block (cross-attention) has nothing to interleave with
    issue_s(1, 0);
    if (elect_one()) umma_commit(&k_empty[0]);
    __syncwarp();
    int st = 0;
    uint32_t ph = 0;
    for (int j = 0; j < nblk; ++j) {
      int st_n = st + 1;
      uint32_t ph_n = ph;
      if (st_n == stages) {
        st_n = 0;
        ph_n ^= 1;
      }
      const bool more = j + 1 < nblk;
      if (sep && more) {
        // S_t(j+1) as soon as S_t(j) sits in the softmax threads' registers
        for (int t = 0; t < 2; ++t) {
          mbar_wait(&s_free[t], (uint32_t)(j & 1), 25);
          if (t == 0) mbar_wait(&k_full[st_n], ph_n, 20);
          tc_fence_after();
          issue_s(t, st_n);
        }
        if (elect_one()) umma_commit(&k_empty[st_n]);
        __syncwarp();
      }
      mbar_wait(&v_full[st], ph, 23);
      const int kv_valid = min(BKV, a.nk - j * BKV);
      const int ksteps_o = (kv_valid + 15) / 16;
      for (int t = 0; t < 2; ++t) {
        mbar_wait(&p_full[t], (uint32_t)(j & 1), 22);  // P_t(j) is in TMEM
        tc_fence_after();
        if (elect_one()) {
          const uint32_t va = v_addr + st * v_stage_bytes;
          const uint32_t dO = tmem_base + (t ? colO1 : colO0);
          const uint32_t aP = tmem_base + (t ? colS1 : colS0) + offP;
          for (int kk = 0; kk < ksteps_o; ++kk) {
            const uint32_t offv = (uint32_t)((kk >> 2) * vt_atom_bytes + (kk & 3) * 32);
            // A = P_t: 16-bit elements packed two per TMEM column -> 8 columns per K = 16 step
            umma_f16_ts(dO, aP + kk * 8, umma_desc_sw128(va + offv), idesc_o, (j > 0 || kk > 0) ? 1u : 0u);
          }
          // separate-P mode: every block (the softmax waits for it before overwriting P_t / rescaling O_t); aliased mode:
          // only the last block (s_full of the next block already implies it on the in-order pipe)
          if (sep || !more) umma_commit(&pv_done[t]);
          if (t == 1) umma_commit(&v_empty[st]);
        }
        __syncwarp();
        if (!sep && more) {
          if (t == 0) {
            mbar_wait(&k_full[st_n], ph_n, 20);
            tc_fence_after();
          }
          issue_s(t, st_n);  // executes after P V_t(j) on the in-order tensor pipe: safe to overwrite S_t / P_t
          if (t == 1) {
            if (elect_one()) umma_commit(&k_empty[st_n]);
            __syncwarp();
          }
        }
      }
      st = st_n;
      ph = ph_n;
    }
  } else if (warp < 10) {
    // ================= softmax warpgroups: thread <-> query row =================
    const int t = (warp - 2) >> 2;  // tile
    const int qd = warp & 3;        // TMEM lane quarter of this warp
    const int r = qd * 32 + lane;
    const uint32_t lane_off = (uint32_t)(qd * 32) << 16;
    const uint32_t tS = tmem_base + (t ? colS1 : colS0) + lane_off;
    const uint32_t tP = tS + offP;
    const uint32_t tO = tmem_base + (t ? colO1 : colO0) + lane_off;
    float m_used = -INFINITY;
    for (int j = 0; j < nblk; ++j) {
      const int kv_valid = min(BKV, a.nk - j * BKV);
      ATT_STAMP(0);
      mbar_wait(&s_full[t], (uint32_t)(j & 1), 30);
      tc_fence_after();
      ATT_STAMP(1);
      uint32_t s[128];
      tmem_ld32(tS + 0, reinterpret_cast<uint32_t(&)[32]>(s[0]));
      tmem_ld32(tS + 32, reinterpret_cast<uint32_t(&)[32]>(s[32]));
      tmem_ld32(tS + 64, reinterpret_cast<uint32_t(&)[32]>(s[64]));
      tmem_ld32(tS + 96, reinterpret_cast<uint32_t(&)[32]>(s[96]));
      tmem_ld_wait();
      ATT_STAMP(2);
      if (sep) {  // the S row now lives in registers: the tensor pipe may overwrite S_t with the next block
        tc_fence_before();
        mbar_arrive(&s_free[t]);
      }
      const bool full = (kv_valid == BKV);
      float mx0 = -INFINITY, mx1 = -INFINITY;
      if (full) {
#pragma unroll
        for (int e = 0; e < 128; e += 2) {
          mx0 = fmaxf(mx0, __uint_as_float(s[e]));
          mx1 = fmaxf(mx1, __uint_as_float(s[e + 1]));
        }
      } else {
#pragma unroll
        for (int e = 0; e < 128; ++e)
          if (e < kv_valid) mx0 = fmaxf(mx0, __uint_as_float(s[e]));
      }
      ATT_STAMP(3);
      if (t == 0 && j == 0) mbar_arrive(stagger);
      const float m_blk = fmaxf(mx0, mx1) * a.scale_log2;
      float alpha = 1.0f;
      bool need = false;
      if (j == 0) {
        m_used = m_blk;
      } else if (m_blk > m_used + RESCALE_TAU) {
        alpha = exp2f(m_used - m_blk);
        m_used = m_blk;
        need = true;
      }
      // Separate-P mode: P V_t(j-1) must be complete before O_t is rescaled (rare) and before P_t is overwritten - the
      // latter only after the exp phase, so the P V round trip hides behind it.
      bool pv_waited = !sep || j == 0;
      if (__any_sync(0xffffffffu, need)) {  // rare: rescale this warp's 32 rows of O
        if (!pv_waited) {
          mbar_wait(&pv_done[t], (uint32_t)((j - 1) & 1), 31);
          tc_fence_after();
          pv_waited = true;
        }
        for (int c = 0; c < a.dv; c += 16) {
          uint32_t o16[16];
          tmem_ld16(tO + c, o16);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 16; ++e) o16[e] = __float_as_uint(__uint_as_float(o16[e]) * alpha);
          tmem_st16(tO + c, o16);
        }
      }
      // Separate-P mode: nothing couples the two tiles any more, and they drift into lockstep (both in the exp phase,
      // each at half MUFU rate, then both in the MUFU-free phases).  The two warps that share an SM sub-partition
      // (tile 0 / tile 1, same lane quarter) therefore alternate their exp phases: T0(j), T1(j), T0(j+1), ...
      if (sep && a.xu_gate) {
        if (t == 0) {
          if (j > 0) mbar_wait(&xu_done[4 + qd], (uint32_t)((j - 1) & 1), 33);
        } else {
          mbar_wait(&xu_done[qd], (uint32_t)(j & 1), 34);
        }
      }
      ATT_STAMP(4);
      // p = 2^(s * scale_log2 - m): one FFMA feeding MUFU.EX2; packed pairs overwrite s[] in place (s[e/2] <- e, e+1)
      const float neg_m = -m_used;
      if (full) {
#pragma unroll
        for (int e = 0; e < 64; e += 2) {
          const float p0 = fast_ex2(fmaf(__uint_as_float(s[e]), a.scale_log2, neg_m));
          const float p1 = fast_ex2(fmaf(__uint_as_float(s[e + 1]), a.scale_log2, neg_m));
          s[e >> 1] = pack_act2(p0, p1, f16);
        }
        if (sep && a.xu_gate == 2) {  // half-way: the other tile may start its exp phase (half-period offset, overlap allowed)
          __syncwarp();
          if (lane == 0) mbar_arrive(&xu_done[t * 4 + qd]);
        }
#pragma unroll
        for (int e = 64; e < 128; e += 2) {
          const float p0 = fast_ex2(fmaf(__uint_as_float(s[e]), a.scale_log2, neg_m));
          const float p1 = fast_ex2(fmaf(__uint_as_float(s[e + 1]), a.scale_log2, neg_m));
          s[e >> 1] = pack_act2(p0, p1, f16);
        }
      } else {
        if (sep && a.xu_gate == 2) {
          __syncwarp();
          if (lane == 0) mbar_arrive(&xu_done[t * 4 + qd]);
        }
#pragma unroll
        for (int e = 0; e < 128; e += 2) {
          float p0 = fast_ex2(fmaf(__uint_as_float(s[e]), a.scale_log2, neg_m));
          float p1 = fast_ex2(fmaf(__uint_as_float(s[e + 1]), a.scale_log2, neg_m));
          if (e >= kv_valid) p0 = 0.f;
          if (e + 1 >= kv_valid) p1 = 0.f;
          s[e >> 1] = pack_act2(p0, p1, f16);
        }
      }
      ATT_STAMP(5);
      if (sep && a.xu_gate == 1) {
        __syncwarp();
        if (lane == 0) mbar_arrive(&xu_done[t * 4 + qd]);
      }
      if (!pv_waited) {
        mbar_wait(&pv_done[t], (uint32_t)((j - 1) & 1), 31);
        tc_fence_after();
      }
      tmem_st32(tP + 0, &s[0]);
      tmem_st32(tP + 32, &s[32]);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_full[t]);
      ATT_STAMP(6);
    }
    // ---- epilogue: O / l -> global (l = O[:, d], accumulated by the ones row of V^T) ----
    mbar_wait(&pv_done[t], sep ? (uint32_t)((nblk - 1) & 1) : 0u, 32);
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
    for (int c = 0; c < a.dqk; c += 16) {
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
extern "C" int cpd_debug_attn_timeline(long long* host128) {
  return (int)cudaMemcpyFromSymbol(host128, cpd_dbg_attn, sizeof(long long) * 128);
}
#endif

// Returns CPD_ERR_UNSUPPORTED when the shape is outside this kernel's domain (the caller falls back to the one-tile kernel).
cpd_status cpd_attention_2tile(const cpd_attn_params* p, void* stream) {
  const int d = p->d_head;
  if (d <= 0 || d > p->dpad || p->nq <= BQ) return CPD_ERR_UNSUPPORTED;
  const int dv = (d + 1 + 15) / 16 * 16;
  if (dv > 128) return CPD_ERR_UNSUPPORTED;
  Attn2Args a;
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
  int stages = (227 * 1024 - 1024 - 512 - q_bytes) / per_stage;
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
  const size_t shm = (size_t)q_bytes + (size_t)stages * per_stage + 512 + 1024;
  static bool configured = false;
  if (!configured) {
    CPD_CUDA_CHECK(cudaFuncSetAttribute(attention2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)));
    CPD_CUDA_CHECK(cudaFuncSetAttribute(attention2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)));
    configured = true;
  }
  dim3 grid((p->nq + 2 * BQ - 1) / (2 * BQ), p->heads, p->batch);
  // The separate-P layout (S(j+1) issued as soon as S(j) is in registers; one MMA issuer warp per tile; optional exp-phase
  // gate between the tiles) removes the wait for S but measures no faster than aliasing P over S (839-870 vs 829-838 us on
  // 16 x 8 x 4096^2, d = 40): a lone warp in its exp phase issues a MUFU.EX2 only every ~11-13 cycles (8 would saturate the
  // pipe), so with two softmax warps per SM sub-partition the MUFU pipe stays ~60 % busy either way
  // (profiles/r01_attn_timelines.txt).  Opt-in (CPD_ATTN_SEP=1) for experiments.
  static int sep_env = -1;
  if (sep_env < 0) {
    const char* e = getenv("CPD_ATTN_SEP");
    sep_env = (e && e[0] == '1') ? 1 : 0;
  }
  static int gate_env = -1;
  if (gate_env < 0) {
    const char* e = getenv("CPD_ATTN_GATE");
    gate_env = e ? (e[0] - '0') : 2;  // 0 none, 1 strict alternation, 2 half-period offset
  }
  a.xu_gate = gate_env;
  if (dv <= 64 && sep_env)
    CPD_CUDA_CHECK(cpd_launch(attention2_kernel<true>, dim3(grid), dim3(NUM_THREADS), shm, (cudaStream_t)stream, a));
  else
    CPD_CUDA_CHECK(cpd_launch(attention2_kernel<false>, dim3(grid), dim3(NUM_THREADS), shm, (cudaStream_t)stream, a));
  CPD_CUDA_CHECK(cudaGetLastError());
  return CPD_OK;
}
