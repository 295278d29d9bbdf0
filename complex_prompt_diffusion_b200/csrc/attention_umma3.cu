// Persistent fused flash-style attention (sm_100a): the two-query-tile design of attention_umma2.cu (S = Q K^T by SS-mode
// tcgen05.mma, one-pass register softmax, P stored back to TMEM over S, O += P V by TS-mode tcgen05.mma, row sums through
// a ones row of V^T, lazy rescaling, staggered tiles) wrapped in a per-CTA loop over work items
// (batch, head, 256-query-row pair):
//   * one CTA per SM for the whole launch: barrier init, TMEM allocation and the V^T pad rows are paid once, not per item;
//   * Q is double-buffered and the K / V^T stage ring runs across items, so item i+1's loads and its first Q K^T are in
//     flight while item i's last P V, normalisation and stores drain (cross-attention with 77 keys is ONE key block per
//     item: without this overlap each CTA is a serial load -> MMA -> softmax -> MMA -> store latency chain);
//   * O_t of item i is handed back to the tensor pipe by an o_free barrier as soon as the softmax threads have read it.
// Replaces cpd/models/attention.py:283-348 for head dims <= 111 and more than 128 query rows.
#include <stdlib.h>

#include "../../include/cpd_b200.h"
#include "common.cuh"

namespace {

constexpr int BQ = 128;
constexpr int ATOM_BYTES = 128 * 128;  // 128 rows x 64 16-bit elements
// threads: warp 0 TMA, warp 1 MMA, then one softmax warpgroup (4 warps) per query tile: 64 + NT * 128
constexpr int MAX_STAGES = 4;
constexpr float RESCALE_TAU = 8.0f;    // lazy O rescale threshold (log2 domain)

struct Attn3Args {
  CUtensorMap map_q, map_k, map_vt;
  bf16* o;
  int ldo;
  int batch, heads, nq, nk, nk_pad, kv_batch;
  int dqk;      // padded head dim of Q / K / O columns (multiple of 16)
  int d;        // real head dim (rows of V^T loaded by TMA)
  int dv;       // MMA N of P V: round16(d + 1)
  int datoms;   // ceil(dqk / 64)
  int stages;
  int qbufs;    // Q buffers (2 = the next item's Q is loaded while this item runs; 1 when shared memory is short)
  int fp16;
  int nqp;      // work items (NT x 128 query rows) per (batch, head)
  int items;    // batch * heads * nqp
  float scale_log2;
};

__device__ __forceinline__ void umma_f16_ts3(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st32_3(uint32_t taddr, const uint32_t* r) {
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
// Debug build (make TIMELINE=1): CTA 0 stamps clock64 at the phases of key blocks 8..15 of its first item
// (tools/attn_timeline.py --persistent): softmax [tile][block - 8][phase], MMA issuer [block - 8][tile][PV issued, S issued]
__device__ long long cpd_dbg_attn3[4][8][8];
__device__ long long cpd_dbg_mma3[8][4][4];
#define ATT3_STAMP(ph)                                                                       \
  do {                                                                                       \
    if (blockIdx.x == 0 && it == 0 && qd == 2 && lane == 0 && j >= 8 && j < 16)              \
      cpd_dbg_attn3[t][j - 8][ph] = clock64();                                               \
  } while (0)
#define MMA3_STAMP(ph)                                                                       \
  do {                                                                                       \
    if (blockIdx.x == 0 && it == 0 && lane == 0 && j >= 8 && j < 16)                         \
      cpd_dbg_mma3[j - 8][t][ph] = clock64();                                                \
  } while (0)
#else
#define ATT3_STAMP(ph) \
  do {                 \
  } while (0)
#define MMA3_STAMP(ph) \
  do {                 \
  } while (0)
#endif

// NT query tiles of 128 rows per work item, key blocks of BKV keys.  <2, 128>: the general shape.  <4, 64>: four softmax
// warpgroups (4 warps per SM sub-partition) for head dims <= 63 - a lone warp in its exp phase reaches only ~75 % of the
// MUFU rate (every MUFU.EX2 carries an 8-cycle issue stall, the FFMA / F2FP around it add to it), two or more interleave
// to ~93 %; with four tiles in flight the tensor-pipe round trip of one tile also hides behind the other three.
template <int NT, int BKV>
__global__ void __launch_bounds__(64 + NT * 128, 1) attention3_kernel(const __grid_constant__ Attn3Args a) {
  constexpr int NUM_THREADS = 64 + NT * 128;
  constexpr int K_ATOM = BKV * 128;  // BKV rows x 64 16-bit elements
  constexpr int VATOMS = BKV / 64;   // 64-key atoms of a V^T stage
  constexpr int OCOLS = 256 / NT;    // TMEM columns reserved per O_t
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int datoms = a.datoms;
  const int stages = a.stages;
  const int q_tile_bytes = datoms * ATOM_BYTES;
  const int q_buf_bytes = NT * q_tile_bytes;  // all tiles of an item
  const int k_stage_bytes = datoms * K_ATOM;
  const int vt_atom_bytes = a.dv * 128;
  const int v_stage_bytes = VATOMS * vt_atom_bytes;
  uint8_t* sQ = smem;  // [2 buffers][2 tiles]
  uint8_t* sK = sQ + a.qbufs * q_buf_bytes;
  uint8_t* sV = sK + stages * k_stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + stages * v_stage_bytes);
  uint64_t* q_full = bars;                    // [2]
  uint64_t* q_empty = q_full + 2;             // [2]
  uint64_t* k_full = q_empty + 2;             // [MAX_STAGES]
  uint64_t* k_empty = k_full + MAX_STAGES;
  uint64_t* v_full = k_empty + MAX_STAGES;
  uint64_t* v_empty = v_full + MAX_STAGES;
  uint64_t* s_full = v_empty + MAX_STAGES;    // [NT]  S_t of the current block is in TMEM
  uint64_t* p_full = s_full + NT;             // [NT]  P_t of the current block is in TMEM (128 arrivals)
  uint64_t* pv_last = p_full + NT;            // [NT]  the last P V_t of the item has completed: O_t is final
  uint64_t* o_free = pv_last + NT;            // [NT]  O_t has been read by its softmax warpgroup (128 arrivals)
  uint64_t* stagger = o_free + NT;            // [1]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(stagger + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nblk = (a.nk + BKV - 1) / BKV;
  const bool f16 = a.fp16 != 0;
  const int first = blockIdx.x, step = gridDim.x;

  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&a.map_q);
    tma_prefetch_desc(&a.map_k);
    tma_prefetch_desc(&a.map_vt);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&q_full[s], 1);
      mbar_init(&q_empty[s], 1);
    }
    for (int t = 0; t < NT; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&p_full[t], 128);
      mbar_init(&pv_last[t], 1);
      mbar_init(&o_free[t], 128);
    }
    for (int s = 0; s < stages; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    mbar_init(stagger, 128);
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_smem, 512);
    tmem_relinquish();
  }
  {  // rows d .. dv-1 of every V^T stage atom: ones row (-> O[:, d] = sum of P), then zero rows (never touched by TMA)
    const uint32_t one2 = f16 ? 0x3C003C00u : 0x3F803F80u;
    const int pad_rows = a.dv - a.d;
    const int chunks = stages * VATOMS * pad_rows * 8;
    for (int i = threadIdx.x; i < chunks; i += NUM_THREADS) {
      const int c16 = i & 7;
      const int rr = (i >> 3) % pad_rows;
      const int at = (i >> 3) / pad_rows;
      const uint32_t v = (rr == 0) ? one2 : 0u;
      *reinterpret_cast<uint4*>(sV + at * vt_atom_bytes + (a.d + rr) * 128 + c16 * 16) = make_uint4(v, v, v, v);
    }
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();
  // TMEM columns: S_t / P_t (aliased) at t * BKV, O_t at 256 + t * OCOLS

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int st = 0;
      uint32_t ph = 0;
      int it = 0;
      for (int w = first; w < a.items; w += step, ++it) {
        const int qp = w % a.nqp;
        const int bh = w / a.nqp;
        const int head = bh % a.heads;
        const int b = bh / a.heads;
        const int bkv = b % a.kv_batch;
        const int qb = a.qbufs == 2 ? (it & 1) : 0;
        mbar_wait(&q_empty[qb], (uint32_t)(((a.qbufs == 2 ? (it >> 1) : it) & 1) ^ 1), 12);
        mbar_arrive_expect_tx(&q_full[qb], q_buf_bytes);
        for (int t = 0; t < NT; ++t)
          for (int dd = 0; dd < datoms; ++dd)
            tma_load_2d(sQ + qb * q_buf_bytes + t * q_tile_bytes + dd * ATOM_BYTES, &a.map_q, &q_full[qb], head * a.dqk + dd * 64,
                        b * a.nq + qp * (NT * BQ) + t * BQ);
        for (int j = 0; j < nblk; ++j) {
          mbar_wait(&k_empty[st], ph ^ 1, 10);
          mbar_arrive_expect_tx(&k_full[st], k_stage_bytes);
          for (int dd = 0; dd < datoms; ++dd)
            tma_load_2d(sK + st * k_stage_bytes + dd * K_ATOM, &a.map_k, &k_full[st], head * a.dqk + dd * 64,
                        bkv * a.nk_pad + j * BKV);
          mbar_wait(&v_empty[st], ph ^ 1, 11);
          mbar_arrive_expect_tx(&v_full[st], VATOMS * a.d * 128);
          for (int t = 0; t < VATOMS; ++t)
            tma_load_2d(sV + st * v_stage_bytes + t * vt_atom_bytes, &a.map_vt, &v_full[st], bkv * a.nk_pad + j * BKV + t * 64,
                        head * a.dqk);
          if (++st == stages) {
            st = 0;
            ph ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (whole warp runs the loops; one elected lane issues) =================
    const uint32_t idesc_s = umma_idesc_f16(BQ, BKV, f16, f16);
    const uint32_t idesc_o = umma_idesc_f16(BQ, a.dv, f16, f16);
    const int ksteps_s = a.dqk / 16;
    const uint32_t q_addr = smem_u32(sQ), k_addr = smem_u32(sK), v_addr = smem_u32(sV);
    auto issue_s = [&](int t, int qb, int st) {  // S_t = Q_t K(st)^T
      if (elect_one()) {
        const uint32_t qa = q_addr + qb * q_buf_bytes + t * q_tile_bytes, ka = k_addr + st * k_stage_bytes;
        for (int kk = 0; kk < ksteps_s; ++kk) {
          const uint32_t offq = (uint32_t)((kk >> 2) * ATOM_BYTES + (kk & 3) * 32);
          const uint32_t offk = (uint32_t)((kk >> 2) * K_ATOM + (kk & 3) * 32);
          umma_bf16(tmem_base + t * BKV, umma_desc_sw128(qa + offq), umma_desc_sw128(ka + offk), idesc_s, kk > 0 ? 1u : 0u);
        }
        umma_commit(&s_full[t]);
      }
      __syncwarp();
    };
    // One flat pipeline over (item, key block): right after P V_t of a block the NEXT block's S_t is issued - also when it
    // belongs to the next work item - so a tile never waits for the other tile or for an item boundary.  On the in-order
    // tensor pipe S_t(next) runs after P V_t(current), which is what makes overwriting the aliased P_t safe.
    int st = 0;         // K / V stage of the current block
    uint32_t ph = 0;
    uint32_t blk = 0;   // blocks finished so far (parity of p_full)
    const int n_my = first < a.items ? (a.items - first + step - 1) / step : 0;  // items of this CTA
    auto q_buf = [&](int it) { return a.qbufs == 2 ? (it & 1) : 0; };
    auto q_par = [&](int it) { return (uint32_t)((a.qbufs == 2 ? (it >> 1) : it) & 1); };
    if (n_my > 0) {
      mbar_wait(&q_full[0], 0, 21);
      mbar_wait(&k_full[0], 0, 20);
      tc_fence_after();
      issue_s(0, 0, 0);
      if (NT == 2 && nblk > 1) mbar_wait(stagger, 0, 24);  // stagger tile 1 behind tile 0's row-max phase (self-attention)
      for (int t = 1; t < NT; ++t) issue_s(t, 0, 0);
      if (elect_one()) {
        umma_commit(&k_empty[0]);
        if (nblk == 1) umma_commit(&q_empty[0]);
      }
      __syncwarp();
    }
    for (int it = 0; it < n_my; ++it) {
      for (int j = 0; j < nblk; ++j, ++blk) {
        int st_n = st + 1;
        uint32_t ph_n = ph;
        if (st_n == stages) {
          st_n = 0;
          ph_n ^= 1;
        }
        const bool last_j = j + 1 == nblk;
        const bool has_next = !last_j || it + 1 < n_my;
        const int it_n = last_j ? it + 1 : it;       // item of the next block
        const int j_n = last_j ? 0 : j + 1;
        mbar_wait(&v_full[st], ph, 23);
        const int kv_valid = min(BKV, a.nk - j * BKV);
        const int ksteps_o = (kv_valid + 15) / 16;
        for (int t = 0; t < NT; ++t) {
          MMA3_STAMP(0);
          mbar_wait(&p_full[t], blk & 1u, 22);  // P_t(j) is in TMEM
          MMA3_STAMP(1);
          if (j == 0 && it > 0) mbar_wait(&o_free[t], (uint32_t)((it - 1) & 1), 26);  // previous item's O_t has been read
          tc_fence_after();
          if (elect_one()) {
            const uint32_t va = v_addr + st * v_stage_bytes;
            for (int kk = 0; kk < ksteps_o; ++kk) {
              const uint32_t offv = (uint32_t)((kk >> 2) * vt_atom_bytes + (kk & 3) * 32);
              umma_f16_ts3(tmem_base + 256 + t * OCOLS, tmem_base + t * BKV + kk * 8, umma_desc_sw128(va + offv), idesc_o,
                           (j > 0 || kk > 0) ? 1u : 0u);
            }
            if (last_j) umma_commit(&pv_last[t]);
            if (t == NT - 1) umma_commit(&v_empty[st]);
          }
          __syncwarp();
          if (has_next) {
            if (t == 0) {
              if (last_j) mbar_wait(&q_full[q_buf(it_n)], q_par(it_n), 21);
              mbar_wait(&k_full[st_n], ph_n, 20);
              tc_fence_after();
            }
            MMA3_STAMP(2);
            issue_s(t, q_buf(it_n), st_n);
            MMA3_STAMP(3);
            if (t == NT - 1) {
              if (elect_one()) {
                umma_commit(&k_empty[st_n]);
                if (j_n + 1 == nblk) umma_commit(&q_empty[q_buf(it_n)]);  // the last Q K^T of that item has been issued
              }
              __syncwarp();
            }
          }
        }
        st = st_n;
        ph = ph_n;
      }
    }
  } else {
    // ================= softmax warpgroups: thread <-> query row =================
    const int t = (warp - 2) >> 2;  // tile
    const int qd = warp & 3;        // TMEM lane quarter of this warp
    const int r = qd * 32 + lane;
    const uint32_t lane_off = (uint32_t)(qd * 32) << 16;
    const uint32_t tS = tmem_base + t * BKV + lane_off;
    const uint32_t tO = tmem_base + 256 + t * OCOLS + lane_off;
    uint32_t blk = 0;
    int it = 0;
    for (int w = first; w < a.items; w += step, ++it) {
      const int qp = w % a.nqp;
      const int bh = w / a.nqp;
      const int head = bh % a.heads;
      const int b = bh / a.heads;
      float m_used = -INFINITY;
      for (int j = 0; j < nblk; ++j, ++blk) {
        const int kv_valid = min(BKV, a.nk - j * BKV);
        ATT3_STAMP(0);
        mbar_wait(&s_full[t], blk & 1u, 30);  // also: P V_t of the previous block has completed (in-order pipe)
        tc_fence_after();
        ATT3_STAMP(1);
        uint32_t s[BKV];
        const bool full = (kv_valid == BKV);
#pragma unroll
        for (int c32 = 0; c32 < BKV / 32; ++c32)
          if (c32 == 0 || kv_valid > c32 * 32) tmem_ld32(tS + c32 * 32, reinterpret_cast<uint32_t(&)[32]>(s[c32 * 32]));
        tmem_ld_wait();
        ATT3_STAMP(2);
        float mx0 = -INFINITY, mx1 = -INFINITY;
        if (full) {
#pragma unroll
          for (int e = 0; e < BKV; e += 2) {
            mx0 = fmaxf(mx0, __uint_as_float(s[e]));
            mx1 = fmaxf(mx1, __uint_as_float(s[e + 1]));
          }
        } else {
#pragma unroll
          for (int e = 0; e < BKV; ++e)
            if (e < kv_valid) mx0 = fmaxf(mx0, __uint_as_float(s[e]));
        }
        if (t == 0 && blk == 0) mbar_arrive(stagger);
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
        if (__any_sync(0xffffffffu, need)) {  // rare: rescale this warp's 32 rows of O (P V_t(j-1) is complete)
          for (int c = 0; c < a.dv; c += 16) {
            uint32_t o16[16];
            tmem_ld16(tO + c, o16);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 16; ++e) o16[e] = __float_as_uint(__uint_as_float(o16[e]) * alpha);
            tmem_st16(tO + c, o16);
          }
        }
        ATT3_STAMP(3);
        // p = 2^(s * scale_log2 - m): one FFMA feeding MUFU.EX2; packed pairs overwrite s[] in place (s[e/2] <- e, e+1)
        const float neg_m = -m_used;
        if (full) {
#pragma unroll
          for (int e = 0; e < BKV; e += 2) {
            const float p0 = fast_ex2(fmaf(__uint_as_float(s[e]), a.scale_log2, neg_m));
            const float p1 = fast_ex2(fmaf(__uint_as_float(s[e + 1]), a.scale_log2, neg_m));
            s[e >> 1] = pack_act2(p0, p1, f16);
          }
        } else {
#pragma unroll
          for (int c32 = 0; c32 < BKV / 32; ++c32) {
            if (c32 * 32 < kv_valid) {  // warp-uniform: whole 32-column groups beyond the valid keys cost nothing
#pragma unroll
              for (int e = c32 * 32; e < c32 * 32 + 32; e += 2) {
                float p0 = fast_ex2(fmaf(__uint_as_float(s[e]), a.scale_log2, neg_m));
                float p1 = fast_ex2(fmaf(__uint_as_float(s[e + 1]), a.scale_log2, neg_m));
                if (e >= kv_valid) p0 = 0.f;
                if (e + 1 >= kv_valid) p1 = 0.f;
                s[e >> 1] = pack_act2(p0, p1, f16);
              }
            } else {
#pragma unroll
              for (int e = c32 * 32; e < c32 * 32 + 32; e += 2) s[e >> 1] = 0u;
            }
          }
        }
        ATT3_STAMP(4);
#pragma unroll
        for (int c32 = 0; c32 < BKV / 64; ++c32) tmem_st32_3(tS + c32 * 32, &s[c32 * 32]);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&p_full[t]);
        ATT3_STAMP(5);
      }
      // ---- item epilogue: O / l -> global (l = O[:, d], accumulated by the ones row of V^T) ----
      mbar_wait(&pv_last[t], (uint32_t)(it & 1), 32);
      tc_fence_after();
      const int qrow = qp * (NT * BQ) + t * BQ + r;
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
        if (c + 16 >= a.dqk) {  // last TMEM read of O_t: hand it back to the tensor pipe before the global stores
          tc_fence_before();
          mbar_arrive(&o_free[t]);
        }
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
extern "C" int cpd_debug_attn3_timeline(long long* soft256, long long* mma128) {
  int rc = (int)cudaMemcpyFromSymbol(soft256, cpd_dbg_attn3, sizeof(long long) * 256);
  if (rc) return rc;
  return (int)cudaMemcpyFromSymbol(mma128, cpd_dbg_mma3, sizeof(long long) * 128);
}
#endif

template <int NT, int BKV>
static cpd_status launch_attention3(const cpd_attn_params* p, int dv, void* stream) {
  const int d = p->d_head;
  Attn3Args a;
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
  a.nqp = (p->nq + NT * BQ - 1) / (NT * BQ);
  a.items = p->batch * p->heads * a.nqp;
  const int per_stage = a.datoms * BKV * 128 + (BKV / 64) * dv * 128;
  a.qbufs = 2;
  int q_bytes = a.qbufs * NT * a.datoms * ATOM_BYTES;  // (double-buffered) tiles of an item
  int stages = (227 * 1024 - 1024 - 512 - q_bytes) / per_stage;
  if (stages < 2) {  // large head dims: one Q buffer
    a.qbufs = 1;
    q_bytes = NT * a.datoms * ATOM_BYTES;
    stages = (227 * 1024 - 1024 - 512 - q_bytes) / per_stage;
  }
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
  CPD_SMEM_OPTIN((attention3_kernel<NT, BKV>), 227 * 1024);
  const int ctas = a.items < 148 ? a.items : 148;
  CPD_CUDA_CHECK(cpd_launch(attention3_kernel<NT, BKV>, dim3(ctas), dim3(64 + NT * 128), shm, (cudaStream_t)stream, a));
  return CPD_OK;
}

// Returns CPD_ERR_UNSUPPORTED when the shape is outside this kernel's domain (the caller falls back).
cpd_status cpd_attention_persistent(const cpd_attn_params* p, void* stream) {
  const int d = p->d_head;
  if (d <= 0 || d > p->dpad || p->nq <= BQ) return CPD_ERR_UNSUPPORTED;
  const int dv = (d + 1 + 15) / 16 * 16;
  if (dv > 128) return CPD_ERR_UNSUPPORTED;
  // (A four-tile x 64-key-block instantiation was measured in round 1 - 1254 vs 838 us on 16 x 8 x 4096^2, d = 40: the single
  // MMA issuer warp cannot feed four tiles - and removed; profiles/r01_attn_timelines.txt.)
  return launch_attention3<2, 128>(p, dv, stream);
}
