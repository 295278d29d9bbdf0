// Cross-attention (ONE key block: the 77 text tokens) as a persistent, two-CTAs-per-SM tcgen05 kernel (sm_100a).
//
// Replaces cpd/models/attention.py:283-348 for attn2 (context K / V) at head dims <= 80.  The persistent two-tile kernel
// (attention_umma3.cu) ran this shape at 73.6 us per 16 x 8 x 4096 x 77 launch where the Q-in / O-out traffic is ~15 us:
// with one key block per work item nothing overlaps inside an item - load -> Q K^T -> softmax -> P V -> normalise -> store is a
// serial latency chain per tile - and only two tiles (8 softmax warps) were in flight per SM.  Here
//   * S is an 128 x 80 tile (N = round16(tokens), not 128): 80 exponentials and 80 TMEM columns per row instead of 128;
//   * a CTA needs only 256 TMEM columns (per tile S 80 | O dv) and <= 113 KB of shared memory, so TWO CTAs are resident per
//     SM: four tiles / 16 softmax warps in flight, two MMA issuers, two TMA producers;
//   * K and V^T of a (context row, head) stay resident in shared memory across all the query tiles that use them: the work
//     items of a CTA are contiguous in (head, context row, image, query tile) order, K / V^T are reloaded only when
//     (head, context row) changes;
//   * Q of item i+1 is prefetched (double buffer) and its Q K^T is issued right behind P V of item i, so S is waiting in
//     TMEM when the softmax warps come back from their epilogue; the MMA issuer POLLS the tiles (mbarrier.test_wait) instead
//     of serving them in a fixed order, so a tile never waits behind the other tile's softmax;
//   * head dim as a template parameter (40 / 64 / 80): the epilogue reads the whole O row with ONE tcgen05.wait::ld, no
//     per-column selects; work items are walked incrementally (no integer divisions inside the loops);
//   * no running maximum / lazy rescale (one key block), row sums through the ones row of V^T, activation format as a
//     template parameter (one F2FP per packed pair), FMNMX3 for the row maximum;
//   * the normalised O tile leaves through shared memory and ONE TMA store per tile.  A thread owns a query row, so direct
//     stores are 16-byte pieces at a 768-byte stride: 21 K store transactions per SM and launch, which the clock64 timeline
//     (tools/attn5_timeline.py, profiles/r02_attn5_timeline.txt) showed as 2500-4800 cycles per item - the whole kernel.  The
//     staging tile reuses the item's own Q tile (dead once S = Q K^T has completed); the Q buffer goes back to the TMA producer
//     when the store has read it.
#include <stdlib.h>

#include "../../include/cpd_b200.h"
#include "common.cuh"

namespace {

constexpr int BQ = 128;
constexpr int ATOM_BYTES = 128 * 128;  // 128 rows x 64 16-bit elements

struct Attn5Args {
  CUtensorMap map_q, map_k, map_vt;
  CUtensorMap map_o;  // (column, row) view of O, box = dqk columns x 128 rows; 128B-swizzled when a row is exactly 128 bytes
  int batch, heads, nq, nk, nk_pad, kv_batch;
  int kv_stages;  // 1 or 2 resident K / V^T sets
  int nqp;        // work items (NT x 128 query rows) per (image, head)
  int reps;       // images per context row (batch / kv_batch)
  int items;      // heads * kv_batch * reps * nqp
  int ipc;        // items per CTA (contiguous)
  float scale_log2;
  long long* dbg;  // optional clock64 timeline (cpd_debug_attention_cross_timeline): [2 tiles][8 items][8 phases] softmax stamps of
                   // CTA 0, then [8 items][2 tiles][4 phases] of its MMA issuer; NULL in production
};

#define A5_STAMP(ph)                                                                                   \
  do {                                                                                                 \
    if (a.dbg && blockIdx.x == 0 && qd == 0 && lane == 0 && i < 8) a.dbg[(t * 8 + i) * 8 + (ph)] = clock64(); \
  } while (0)
#define A5_MMA_STAMP(ph)                                                                               \
  do {                                                                                                 \
    if (a.dbg && blockIdx.x == 0 && lane == 0 && it[t] < 8) a.dbg[128 + (it[t] * 2 + (t & 1)) * 4 + (ph)] = clock64(); \
  } while (0)

__device__ __forceinline__ void umma_f16_ts5(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
// non-blocking probe of an mbarrier phase (the issuer polls several barriers)
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}

// Work items in (head, context row, image, query-tile group) order: items that share K / V^T are contiguous.  One division
// chain at the start of a CTA's range, increments afterwards.
struct Cursor5 {
  int g;  // head * kv_batch + bkv: the K / V^T set
  int head, bkv, rep, qp;
  __device__ __forceinline__ void init(const Attn5Args& a, int w) {
    const int per_g = a.reps * a.nqp;
    g = w / per_g;
    const int r = w - g * per_g;
    head = g / a.kv_batch;
    bkv = g - head * a.kv_batch;
    rep = r / a.nqp;
    qp = r - rep * a.nqp;
  }
  __device__ __forceinline__ void advance(const Attn5Args& a) {
    if (++qp == a.nqp) {
      qp = 0;
      if (++rep == a.reps) {
        rep = 0;
        ++g;
        if (++bkv == a.kv_batch) {
          bkv = 0;
          ++head;
        }
      }
    }
  }
  __device__ __forceinline__ int b(const Attn5Args& a) const { return bkv + rep * a.kv_batch; }
};

// NT query tiles of 128 rows per work item; NG = key columns / 16 (77 tokens -> 5); DH = head dim.
// TMEM (256 columns): S_t / P_t (aliased) at t * 16 NG, O_t at NT * 16 NG + t * DV.
template <int NT, int NG, int DH, bool F16>
__global__ void __launch_bounds__(64 + NT * 128, 2) attention5_kernel(const __grid_constant__ Attn5Args a) {
  constexpr int BKN = NG * 16;
  constexpr int DQK = (DH + 15) / 16 * 16;       // padded head dim of the Q / K / O columns
  constexpr int DV = (DH + 1 + 15) / 16 * 16;    // MMA N of P V: ones row + zero rows follow the DH real rows of V^T
  constexpr int DATOMS = (DQK + 63) / 64;
  constexpr int NUM_THREADS = 64 + NT * 128;
  constexpr int K_ATOM = BKN * 128;        // BKN rows x 64 16-bit elements
  constexpr int VATOMS = (BKN + 63) / 64;  // 64-key atoms of V^T
  constexpr int Q_TILE = DATOMS * ATOM_BYTES;
  constexpr int Q_BUF = NT * Q_TILE;
  constexpr int K_SET = DATOMS * K_ATOM;
  constexpr int VT_ATOM = DV * 128;
  constexpr int V_SET = VATOMS * VT_ATOM;
  constexpr uint32_t COL_O = NT * BKN;
  constexpr int STAGE_TILE = 128 * DQK * 2;  // normalised O tile on its way to the TMA store
  // Own staging tiles when they fit next to the Q double buffer and one K / V^T set in a CTA's half of the SM (head dims 40,
  // 64); otherwise (head dim 80) the item's own Q tile is reused - then Q(i+2) can only be requested after item i has been
  // stored, which the timeline showed as a ~3000-cycle wait for S on every other item (profiles/r02_attn5_timeline_v3.txt).
  constexpr bool OWN_STAGE = 2 * Q_BUF + K_SET + V_SET + NT * STAGE_TILE <= 113 * 1024 - 1024 - 256;
  constexpr bool f16 = F16;
  static_assert(NT * (BKN + DV) <= 256, "TMEM budget of a CTA that shares its SM");
  static_assert(128 * DQK * 2 <= Q_TILE, "the O staging tile reuses the Q tile");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;  // [2 buffers][NT tiles]
  uint8_t* sK = sQ + 2 * Q_BUF;
  uint8_t* sV = sK + a.kv_stages * K_SET;
  uint8_t* sO = sV + a.kv_stages * V_SET;  // [NT] staging tiles (OWN_STAGE)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sO + (OWN_STAGE ? NT * STAGE_TILE : 0));
  uint64_t* q_full = bars;            // [2]
  uint64_t* q_empty = q_full + 2;     // [2]   NT Q K^T commits + NT arrivals once the O staging (in the Q tiles) has been stored
  uint64_t* kv_full = q_empty + 2;    // [2]
  uint64_t* kv_empty = kv_full + 2;   // [2]   NT commits: every tile's last P V on the set
  uint64_t* s_full = kv_empty + 2;    // [NT]  S_t of the item is in TMEM
  uint64_t* p_full = s_full + NT;     // [NT]  P_t is in TMEM (128 arrivals)
  uint64_t* pv_done = p_full + NT;    // [NT]  P V_t has completed: O_t is final
  uint64_t* o_free = pv_done + NT;    // [NT]  O_t has been read by its softmax warpgroup (128 arrivals)
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(o_free + NT);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int w0 = blockIdx.x * a.ipc;
  const int n_my = max(0, min(a.items, w0 + a.ipc) - w0);

  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&a.map_q);
    tma_prefetch_desc(&a.map_k);
    tma_prefetch_desc(&a.map_vt);
    tma_prefetch_desc(&a.map_o);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&q_full[s], 1);
      mbar_init(&q_empty[s], OWN_STAGE ? NT : 2 * NT);
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], NT);
    }
    for (int t = 0; t < NT; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&p_full[t], 128);
      mbar_init(&pv_done[t], 1);
      mbar_init(&o_free[t], 128);
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_smem, 256);
    tmem_relinquish();
  }
  {  // rows DH .. DV-1 of every V^T atom: ones row (-> O[:, DH] = sum of P), then zero rows (never touched by TMA)
    const uint32_t one2 = f16 ? 0x3C003C00u : 0x3F803F80u;
    constexpr int pad_rows = DV - DH;
    const int chunks = a.kv_stages * VATOMS * pad_rows * 8;
    for (int i = threadIdx.x; i < chunks; i += NUM_THREADS) {
      const int c16 = i & 7;
      const int rr = (i >> 3) % pad_rows;
      const int at = (i >> 3) / pad_rows;
      const uint32_t v = (rr == 0) ? one2 : 0u;
      *reinterpret_cast<uint4*>(sV + at * VT_ATOM + (DH + rr) * 128 + c16 * 16) = make_uint4(v, v, v, v);
    }
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0 && n_my > 0) {
      Cursor5 cu;
      cu.init(a, w0);
      int cur_g = -1, gi = -1;
      for (int i = 0; i < n_my; ++i, cu.advance(a)) {
        if (cu.g != cur_g) {  // new (head, context row): its K / V^T set
          cur_g = cu.g;
          ++gi;
          const int st = gi % a.kv_stages;
          mbar_wait(&kv_empty[st], (uint32_t)(((gi / a.kv_stages) & 1) ^ 1), 10);
          mbar_arrive_expect_tx(&kv_full[st], K_SET + VATOMS * DH * 128);
#pragma unroll
          for (int dd = 0; dd < DATOMS; ++dd)
            tma_load_2d(sK + st * K_SET + dd * K_ATOM, &a.map_k, &kv_full[st], cu.head * DQK + dd * 64, cu.bkv * a.nk_pad);
#pragma unroll
          for (int t = 0; t < VATOMS; ++t)
            tma_load_2d(sV + st * V_SET + t * VT_ATOM, &a.map_vt, &kv_full[st], cu.bkv * a.nk_pad + t * 64, cu.head * DQK);
        }
        const int qb = i & 1;
        mbar_wait(&q_empty[qb], (uint32_t)(((i >> 1) & 1) ^ 1), 12);
        mbar_arrive_expect_tx(&q_full[qb], Q_BUF);
#pragma unroll
        for (int t = 0; t < NT; ++t)
#pragma unroll
          for (int dd = 0; dd < DATOMS; ++dd)
            tma_load_2d(sQ + qb * Q_BUF + t * Q_TILE + dd * ATOM_BYTES, &a.map_q, &q_full[qb], cu.head * DQK + dd * 64,
                        cu.b(a) * a.nq + cu.qp * (NT * BQ) + t * BQ);
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (whole warp runs the loop; one elected lane issues) =================
    // Per tile a two-state machine, polled: [P] P_t(i) stored -> issue P V_t(i); [S] Q(i+1) (and its K / V^T set) landed ->
    // issue S_t(i+1) = Q K^T right behind it (in-order pipe: P_t(i), which S_t(i+1) overwrites, has been consumed).
    const uint32_t idesc_s = umma_idesc_f16(BQ, BKN, f16, f16);
    const uint32_t idesc_o = umma_idesc_f16(BQ, DV, f16, f16);
    const uint32_t q_addr = smem_u32(sQ), k_addr = smem_u32(sK), v_addr = smem_u32(sV);
    auto issue_s = [&](int t, int qb, int st) {  // S_t = Q_t K^T; frees its share of the Q buffer when the MMAs complete
      if (elect_one()) {
        const uint32_t qa = q_addr + qb * Q_BUF + t * Q_TILE, ka = k_addr + st * K_SET;
#pragma unroll
        for (int kk = 0; kk < DQK / 16; ++kk) {
          const uint32_t offq = (uint32_t)((kk >> 2) * ATOM_BYTES + (kk & 3) * 32);
          const uint32_t offk = (uint32_t)((kk >> 2) * K_ATOM + (kk & 3) * 32);
          umma_bf16(tmem_base + t * BKN, umma_desc_sw128(qa + offq), umma_desc_sw128(ka + offk), idesc_s, kk > 0 ? 1u : 0u);
        }
        umma_commit(&s_full[t]);
        umma_commit(&q_empty[qb]);
      }
      __syncwarp();
    };
    if (n_my > 0) {
      Cursor5 cu[NT];   // item it[t] of tile t
      int it[NT], gi[NT], st_ph[NT];  // item index, index of its K / V^T set, state (0 = [P], 1 = [S], 2 = done)
      mbar_wait(&kv_full[0], 0, 20);
      mbar_wait(&q_full[0], 0, 21);
      tc_fence_after();
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        cu[t].init(a, w0);
        it[t] = 0;
        gi[t] = 0;
        st_ph[t] = 0;
        issue_s(t, 0, 0);
      }
      int live = NT;
      while (live > 0) {
        bool progress = false;
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          if (st_ph[t] == 0) {
            const int i = it[t];
            bool ready = mbar_test(&p_full[t], (uint32_t)(i & 1)) && (i == 0 || mbar_test(&o_free[t], (uint32_t)((i - 1) & 1)));
            ready = __shfl_sync(0xffffffffu, ready, 0);
            if (!ready) continue;
            progress = true;
            A5_MMA_STAMP(1);
            tc_fence_after();
            Cursor5 nx = cu[t];
            nx.advance(a);
            const bool next = i + 1 < n_my;
            const bool chg = next && nx.g != cu[t].g;
            const int st = gi[t] % a.kv_stages;
            if (elect_one()) {
              const uint32_t va = v_addr + st * V_SET;
#pragma unroll
              for (int kk = 0; kk < NG; ++kk) {
                const uint32_t offv = (uint32_t)((kk >> 2) * VT_ATOM + (kk & 3) * 32);
                umma_f16_ts5(tmem_base + COL_O + t * DV, tmem_base + t * BKN + kk * 8, umma_desc_sw128(va + offv), idesc_o, kk > 0 ? 1u : 0u);
              }
              umma_commit(&pv_done[t]);
              if (chg || !next) umma_commit(&kv_empty[st]);  // this tile's last use of the set
            }
            __syncwarp();
            A5_MMA_STAMP(2);
            if (!next) {
              st_ph[t] = 2;
              --live;
              continue;
            }
            st_ph[t] = 1;
          }
          if (st_ph[t] == 1) {
            const int i1 = it[t] + 1;
            Cursor5 nx = cu[t];
            nx.advance(a);
            const bool chg = nx.g != cu[t].g;
            const int gi1 = gi[t] + (chg ? 1 : 0);
            bool ready = mbar_test(&q_full[i1 & 1], (uint32_t)((i1 >> 1) & 1)) &&
                         (!chg || mbar_test(&kv_full[gi1 % a.kv_stages], (uint32_t)((gi1 / a.kv_stages) & 1)));
            ready = __shfl_sync(0xffffffffu, ready, 0);
            if (!ready) continue;
            progress = true;
            tc_fence_after();
            issue_s(t, i1 & 1, gi1 % a.kv_stages);
            A5_MMA_STAMP(3);
            cu[t] = nx;
            it[t] = i1;
            gi[t] = gi1;
            st_ph[t] = 0;
          }
        }
        if (!progress) __nanosleep(20);  // nothing ready: leave the issue slots of this SM sub-partition to its softmax warps
      }
    }
  } else {
    // ================= softmax warpgroups: thread <-> query row =================
    const int t = (warp - 2) >> 2;  // tile
    const int qd = warp & 3;        // TMEM lane quarter of this warp
    const int r = qd * 32 + lane;
    const uint32_t lane_off = (uint32_t)(qd * 32) << 16;
    const uint32_t tS = tmem_base + t * BKN + lane_off;
    const uint32_t tO = tmem_base + COL_O + t * DV + lane_off;
    constexpr int ROW_BYTES = DQK * 2;
    const uint32_t swz = ROW_BYTES == 128 ? (uint32_t)(r & 7) : 0u;
    Cursor5 cu;
    cu.init(a, w0);
    for (int i = 0; i < n_my; ++i, cu.advance(a)) {
      A5_STAMP(0);
      mbar_wait(&s_full[t], (uint32_t)(i & 1), 30);
      tc_fence_after();
      A5_STAMP(1);
      uint32_t s[BKN];
#pragma unroll
      for (int g = 0; g < NG; ++g) tmem_ld16(tS + g * 16, reinterpret_cast<uint32_t(&)[16]>(s[g * 16]));
      tmem_ld_wait();
      A5_STAMP(2);
      // keys >= nk inside the last 16-column group (77 tokens -> 3 pad columns): -inf before the maximum, p = 0 after it
#pragma unroll
      for (int e = BKN - 16; e < BKN; ++e)
        if (e >= a.nk) s[e] = 0xff800000u;
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int e = 0; e < BKN; e += 4) {
        mx0 = fmax3(mx0, __uint_as_float(s[e]), __uint_as_float(s[e + 1]));
        mx1 = fmax3(mx1, __uint_as_float(s[e + 2]), __uint_as_float(s[e + 3]));
      }
      const float neg_m = -fmaxf(mx0, mx1) * a.scale_log2;
#pragma unroll
      for (int e = 0; e < BKN; e += 2) {  // p = 2^(s * scale_log2 - m); 2^(-inf) = 0 for the pad keys
        const float p0 = fast_ex2(fmaf(__uint_as_float(s[e]), a.scale_log2, neg_m));
        const float p1 = fast_ex2(fmaf(__uint_as_float(s[e + 1]), a.scale_log2, neg_m));
        s[e >> 1] = pack_act2(p0, p1, f16);
      }
      A5_STAMP(3);
#pragma unroll
      for (int c = 0; c + 16 <= BKN / 2; c += 16) tmem_st16(tS + c, reinterpret_cast<const uint32_t(&)[16]>(s[c]));
      if constexpr ((BKN / 2) % 16 == 8) tmem_st8(tS + (BKN / 2 - 8), &s[BKN / 2 - 8]);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_full[t]);
      A5_STAMP(4);
      // ---- epilogue: O / l (l = O[:, DH], accumulated by the ones row of V^T) -> staging tile -> one TMA store ----
      mbar_wait(&pv_done[t], (uint32_t)(i & 1), 32);
      tc_fence_after();
      A5_STAMP(5);
      uint32_t o[DV];
#pragma unroll
      for (int g = 0; g < DV / 16; ++g) tmem_ld16(tO + g * 16, reinterpret_cast<uint32_t(&)[16]>(o[g * 16]));
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&o_free[t]);  // O_t goes back to the tensor pipe
      A5_STAMP(6);
      const float inv_l = 1.0f / __uint_as_float(o[DH]);
      // staging tile = this item's own Q tile (S = Q K^T completed long ago): [128 rows][DQK] 16-bit; rows of exactly 128 bytes
      // are 128B-swizzled (16-byte chunk ^= row % 8) like the tensor map, so a quarter warp's st.shared spread over the banks
      uint8_t* stage = OWN_STAGE ? sO + t * STAGE_TILE : sQ + (i & 1) * Q_BUF + t * Q_TILE;
      uint8_t* my_row = stage + r * ROW_BYTES;
      if (OWN_STAGE && i > 0) {  // the previous item's store must have read the staging tile before it is overwritten
        if (warp == 2 + 4 * t && lane == 0) bulk_wait_group_read<0>();  // (the thread that committed the store group)
        named_bar_sync(1 + t, 128);
      }
#pragma unroll
      for (int c = 0; c < DQK; c += 8) {
        uint32_t w[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float v0 = (c + 2 * e < DH) ? __uint_as_float(o[c + 2 * e]) * inv_l : 0.f;
          const float v1 = (c + 2 * e + 1 < DH) ? __uint_as_float(o[c + 2 * e + 1]) * inv_l : 0.f;
          w[e] = pack_act2(v0, v1, f16);
        }
        sts128(smem_u32(my_row) + ((((uint32_t)(c >> 3)) ^ swz) << 4), make_uint4(w[0], w[1], w[2], w[3]));
      }
      fence_proxy_async_smem();          // generic-proxy writes -> visible to the TMA engine
      named_bar_sync(1 + t, 128);        // the tile's 128 rows are staged
      if (warp == 2 + 4 * t && lane == 0) {
        tma_store_2d(&a.map_o, stage, cu.head * DQK, cu.b(a) * a.nq + cu.qp * (NT * BQ) + t * BQ);
        bulk_commit_group();
        if (!OWN_STAGE) {
          bulk_wait_group_read<0>();     // the engine has read the tile: the Q buffer may be refilled
          mbar_arrive(&q_empty[i & 1]);
        }
      }
      __syncwarp();
      A5_STAMP(7);
    }
    if (OWN_STAGE && warp == 2 + 4 * t && lane == 0) bulk_wait_group_read<0>();  // shared memory must outlive the engine's reads
    __syncwarp();
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

long long* g_attn5_dbg = nullptr;

template <int NT, int NG, int DH, bool F16>
cpd_status launch_attention5(const cpd_attn_params* p, void* stream) {
  constexpr int BKN = NG * 16;
  constexpr int DQK = (DH + 15) / 16 * 16;
  constexpr int DV = (DH + 1 + 15) / 16 * 16;
  constexpr int DATOMS = (DQK + 63) / 64;
  Attn5Args a;
  a.batch = p->batch; a.heads = p->heads; a.nq = p->nq; a.nk = p->nk; a.nk_pad = p->nk_pad;
  a.kv_batch = p->kv_batch > 0 ? p->kv_batch : p->batch;
  a.scale_log2 = p->scale * 1.4426950408889634f;
  a.dbg = g_attn5_dbg;
  a.nqp = p->nq / (NT * BQ);
  a.reps = p->batch / a.kv_batch;
  a.items = p->heads * a.kv_batch * a.reps * a.nqp;
  constexpr int q_bytes = 2 * NT * DATOMS * ATOM_BYTES;
  constexpr int kv_set = DATOMS * BKN * 128 + ((BKN + 63) / 64) * DV * 128;
  constexpr int budget = 113 * 1024 - 1024 - 256;  // two CTAs per SM
  static_assert(q_bytes + kv_set <= budget, "shared memory of a CTA that shares its SM");
  constexpr int stage_bytes = (q_bytes + kv_set + NT * 128 * DQK * 2 <= budget) ? NT * 128 * DQK * 2 : 0;  // OWN_STAGE of the kernel
  a.kv_stages = (q_bytes + stage_bytes + 2 * kv_set <= budget) ? 2 : 1;
  const int ctas = a.items < 2 * 148 ? a.items : 2 * 148;
  a.ipc = (a.items + ctas - 1) / ctas;
  const int grid = (a.items + a.ipc - 1) / a.ipc;
  int rc;
  {
    uint64_t dims[2] = {(uint64_t)p->heads * DQK, (uint64_t)p->batch * p->nq};
    uint64_t str[1] = {(uint64_t)p->ldq * 2};
    uint32_t box[2] = {64, BQ};
    if ((rc = cpd_make_tmap_bf16(&a.map_q, p->q, 2, dims, str, box))) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)p->heads * DQK, (uint64_t)a.kv_batch * p->nk_pad};
    uint64_t str[1] = {(uint64_t)p->ldk * 2};
    uint32_t box[2] = {64, BKN};
    if ((rc = cpd_make_tmap_bf16(&a.map_k, p->k, 2, dims, str, box))) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)a.kv_batch * p->nk_pad, (uint64_t)p->heads * DQK};
    uint64_t str[1] = {(uint64_t)p->ldvt * 2};
    uint32_t box[2] = {64, (uint32_t)DH};
    if ((rc = cpd_make_tmap_bf16(&a.map_vt, p->vt, 2, dims, str, box))) return rc;
  }
  {  // O: one box = the DQK columns of a head x 128 query rows (rows of 128 bytes use the 128B swizzle, see the epilogue)
    uint64_t dims[2] = {(uint64_t)p->heads * DQK, (uint64_t)p->batch * p->nq};
    uint64_t str[1] = {(uint64_t)p->ldo * 2};
    uint32_t box[2] = {(uint32_t)DQK, BQ};
    if ((rc = cpd_make_tmap16(&a.map_o, p->o, 2, dims, str, box, DQK * 2 == 128 ? 128 : 0))) return rc;
  }
  const size_t shm = (size_t)q_bytes + (size_t)a.kv_stages * kv_set + stage_bytes + 256 + 1024;
  CPD_SMEM_OPTIN((attention5_kernel<NT, NG, DH, F16>), 113 * 1024);
  CPD_CUDA_CHECK(cpd_launch(attention5_kernel<NT, NG, DH, F16>, dim3(grid), dim3(64 + NT * 128), shm, (cudaStream_t)stream, a));
  return CPD_OK;
}

}  // namespace

// Debug aid (tools/attn5_timeline.py): clock64 stamps of CTA 0's first 8 work items go to `dev_buf` (192 long longs on
// the device) for every following cross-attention launch; NULL switches it off.
extern "C" void cpd_debug_attention_cross_timeline(long long* dev_buf) { g_attn5_dbg = dev_buf; }

// Returns CPD_ERR_UNSUPPORTED when the shape is outside this kernel's domain (the caller falls back).
cpd_status cpd_attention_cross(const cpd_attn_params* p, void* stream) {
  const int d = p->d_head;
  if (d <= 0 || p->dpad != (d + 15) / 16 * 16 || p->nq % BQ) return CPD_ERR_UNSUPPORTED;  // whole 128-row tiles (TMA-stored)
  const int bkn = (p->nk + 15) / 16 * 16;
  if (bkn != 80 || bkn > p->nk_pad) return CPD_ERR_UNSUPPORTED;  // instantiated for the 77-token context (5 x 16 key columns)
  const int kvb = p->kv_batch > 0 ? p->kv_batch : p->batch;
  if (p->batch % kvb) return CPD_ERR_UNSUPPORTED;
  const bool f16 = p->act_fp16 != 0;
  // instantiated head dims: 40 (SD-1.x at 64 x 64: two tiles per CTA), 64 (SD-2.x / SDXL), 80 (SD-1.x at 32 x 32)
  if (d == 40 && p->nq % (2 * BQ) == 0) return f16 ? launch_attention5<2, 5, 40, true>(p, stream) : launch_attention5<2, 5, 40, false>(p, stream);
  if (d == 40) return f16 ? launch_attention5<1, 5, 40, true>(p, stream) : launch_attention5<1, 5, 40, false>(p, stream);
  if (d == 64) return f16 ? launch_attention5<1, 5, 64, true>(p, stream) : launch_attention5<1, 5, 64, false>(p, stream);
  if (d == 80) return f16 ? launch_attention5<1, 5, 80, true>(p, stream) : launch_attention5<1, 5, 80, false>(p, stream);
  return CPD_ERR_UNSUPPORTED;
}
