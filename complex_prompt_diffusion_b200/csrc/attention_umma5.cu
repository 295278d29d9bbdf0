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
//     TMEM when the softmax warps come back from their epilogue;
//   * no running maximum / lazy rescale (one key block), row sums through the ones row of V^T, activation format as a
//     template parameter (one F2FP per packed pair), FMNMX3 for the row maximum.
#include <stdlib.h>

#include "../../include/cpd_b200.h"
#include "common.cuh"

namespace {

constexpr int BQ = 128;
constexpr int ATOM_BYTES = 128 * 128;  // 128 rows x 64 16-bit elements

struct Attn5Args {
  CUtensorMap map_q, map_k, map_vt;
  bf16* o;
  int ldo;
  int batch, heads, nq, nk, nk_pad, kv_batch;
  int dqk;        // padded head dim of Q / K / O columns (multiple of 16)
  int d;          // real head dim (rows of V^T loaded by TMA)
  int dv;         // MMA N of P V: round16(d + 1) (ones row + zero rows follow the d real rows)
  int datoms;     // ceil(dqk / 64)
  int kv_stages;  // 1 or 2 resident K / V^T sets
  int nqp;        // work items (NT x 128 query rows) per (image, head)
  int reps;       // images per context row (batch / kv_batch)
  int items;      // heads * kv_batch * reps * nqp
  int ipc;        // items per CTA (contiguous)
  float scale_log2;
};

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

// item -> (head, context row, image, query-tile group): items that share K / V^T are contiguous
struct Item5 {
  int g;     // head * kv_batch + bkv
  int head, bkv, b, qp;
};
__device__ __forceinline__ Item5 decode_item(const Attn5Args& a, int w) {
  Item5 it;
  const int per_g = a.reps * a.nqp;
  it.g = w / per_g;
  const int r = w - it.g * per_g;
  it.head = it.g / a.kv_batch;
  it.bkv = it.g - it.head * a.kv_batch;
  const int rep = r / a.nqp;
  it.qp = r - rep * a.nqp;
  it.b = it.bkv + rep * a.kv_batch;
  return it;
}

// NT query tiles of 128 rows per work item; NG = key columns / 16 (77 tokens -> 5).  TMEM: S_t at t * 16 NG, O_t behind them.
template <int NT, int NG, bool F16>
__global__ void __launch_bounds__(64 + NT * 128, 2) attention5_kernel(const __grid_constant__ Attn5Args a) {
  constexpr int BKN = NG * 16;
  constexpr int NUM_THREADS = 64 + NT * 128;
  constexpr int K_ATOM = BKN * 128;       // BKN rows x 64 16-bit elements
  constexpr int VATOMS = (BKN + 63) / 64;  // 64-key atoms of V^T
  constexpr bool f16 = F16;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int datoms = a.datoms;
  const int q_tile_bytes = datoms * ATOM_BYTES;
  const int q_buf_bytes = NT * q_tile_bytes;
  const int k_set_bytes = datoms * K_ATOM;
  const int vt_atom_bytes = a.dv * 128;
  const int v_set_bytes = VATOMS * vt_atom_bytes;
  uint8_t* sQ = smem;  // [2 buffers][NT tiles]
  uint8_t* sK = sQ + 2 * q_buf_bytes;
  uint8_t* sV = sK + a.kv_stages * k_set_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + a.kv_stages * v_set_bytes);
  uint64_t* q_full = bars;            // [2]
  uint64_t* q_empty = q_full + 2;     // [2]
  uint64_t* kv_full = q_empty + 2;    // [2]
  uint64_t* kv_empty = kv_full + 2;   // [2]
  uint64_t* s_full = kv_empty + 2;    // [NT]  S_t of the item is in TMEM
  uint64_t* p_full = s_full + NT;     // [NT]  P_t is in TMEM (128 arrivals)
  uint64_t* pv_done = p_full + NT;    // [NT]  P V_t has completed: O_t is final
  uint64_t* o_free = pv_done + NT;    // [NT]  O_t has been read by its softmax warpgroup (128 arrivals)
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(o_free + NT);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int w0 = blockIdx.x * a.ipc;
  const int w1 = min(a.items, w0 + a.ipc);
  const int n_my = max(0, w1 - w0);

  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&a.map_q);
    tma_prefetch_desc(&a.map_k);
    tma_prefetch_desc(&a.map_vt);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&q_full[s], 1);
      mbar_init(&q_empty[s], 1);
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
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
  {  // rows d .. dv-1 of every V^T atom: ones row (-> O[:, d] = sum of P), then zero rows (never touched by TMA)
    const uint32_t one2 = f16 ? 0x3C003C00u : 0x3F803F80u;
    const int pad_rows = a.dv - a.d;
    const int chunks = a.kv_stages * VATOMS * pad_rows * 8;
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
  const uint32_t colO = NT * BKN;  // O_t at colO + t * dv

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int cur_g = -1, gi = -1;
      for (int i = 0; i < n_my; ++i) {
        const Item5 it = decode_item(a, w0 + i);
        if (it.g != cur_g) {  // new (head, context row): its K / V^T set
          cur_g = it.g;
          ++gi;
          const int st = gi % a.kv_stages;
          mbar_wait(&kv_empty[st], (uint32_t)(((gi / a.kv_stages) & 1) ^ 1), 10);
          mbar_arrive_expect_tx(&kv_full[st], k_set_bytes + VATOMS * a.d * 128);
          for (int dd = 0; dd < datoms; ++dd)
            tma_load_2d(sK + st * k_set_bytes + dd * K_ATOM, &a.map_k, &kv_full[st], it.head * a.dqk + dd * 64, it.bkv * a.nk_pad);
          for (int t = 0; t < VATOMS; ++t)
            tma_load_2d(sV + st * v_set_bytes + t * vt_atom_bytes, &a.map_vt, &kv_full[st], it.bkv * a.nk_pad + t * 64,
                        it.head * a.dqk);
        }
        const int qb = i & 1;
        mbar_wait(&q_empty[qb], (uint32_t)(((i >> 1) & 1) ^ 1), 12);
        mbar_arrive_expect_tx(&q_full[qb], q_buf_bytes);
        for (int t = 0; t < NT; ++t)
          for (int dd = 0; dd < datoms; ++dd)
            tma_load_2d(sQ + qb * q_buf_bytes + t * q_tile_bytes + dd * ATOM_BYTES, &a.map_q, &q_full[qb], it.head * a.dqk + dd * 64,
                        it.b * a.nq + it.qp * (NT * BQ) + t * BQ);
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (whole warp runs the loop; one elected lane issues) =================
    const uint32_t idesc_s = umma_idesc_f16(BQ, BKN, f16, f16);
    const uint32_t idesc_o = umma_idesc_f16(BQ, a.dv, f16, f16);
    const int ksteps_s = a.dqk / 16;
    const uint32_t q_addr = smem_u32(sQ), k_addr = smem_u32(sK), v_addr = smem_u32(sV);
    auto issue_s = [&](int t, int qb, int st) {  // S_t = Q_t K^T
      if (elect_one()) {
        const uint32_t qa = q_addr + qb * q_buf_bytes + t * q_tile_bytes, ka = k_addr + st * k_set_bytes;
        for (int kk = 0; kk < ksteps_s; ++kk) {
          const uint32_t offq = (uint32_t)((kk >> 2) * ATOM_BYTES + (kk & 3) * 32);
          const uint32_t offk = (uint32_t)((kk >> 2) * K_ATOM + (kk & 3) * 32);
          umma_bf16(tmem_base + t * BKN, umma_desc_sw128(qa + offq), umma_desc_sw128(ka + offk), idesc_s, kk > 0 ? 1u : 0u);
        }
        umma_commit(&s_full[t]);
      }
      __syncwarp();
    };
    if (n_my > 0) {
      int g_cur = decode_item(a, w0).g;
      int gi = 0;  // index of the current K / V^T set
      mbar_wait(&kv_full[0], 0, 20);
      mbar_wait(&q_full[0], 0, 21);
      tc_fence_after();
      for (int t = 0; t < NT; ++t) issue_s(t, 0, 0);
      if (elect_one()) umma_commit(&q_empty[0]);
      __syncwarp();
      for (int i = 0; i < n_my; ++i) {
        const bool next = i + 1 < n_my;
        const int g_next = next ? decode_item(a, w0 + i + 1).g : g_cur;
        const bool chg = next && g_next != g_cur;
        const bool defer = chg && a.kv_stages == 1;  // one resident set: the next K may only be loaded once every P V of this item is done
        const int st = gi % a.kv_stages;
        const int st_n = chg ? (gi + 1) % a.kv_stages : st;
        const int qb_n = (i + 1) & 1;
        for (int t = 0; t < NT; ++t) {
          mbar_wait(&p_full[t], (uint32_t)(i & 1), 22);                         // P_t is in TMEM
          if (i > 0) mbar_wait(&o_free[t], (uint32_t)((i - 1) & 1), 26);         // the previous item's O_t has been read
          tc_fence_after();
          if (elect_one()) {
            const uint32_t va = v_addr + st * v_set_bytes;
#pragma unroll
            for (int kk = 0; kk < NG; ++kk) {
              const uint32_t offv = (uint32_t)((kk >> 2) * vt_atom_bytes + (kk & 3) * 32);
              umma_f16_ts5(tmem_base + colO + t * a.dv, tmem_base + t * BKN + kk * 8, umma_desc_sw128(va + offv), idesc_o, kk > 0 ? 1u : 0u);
            }
            umma_commit(&pv_done[t]);
            if (t == NT - 1 && (chg || !next)) umma_commit(&kv_empty[st]);  // last use of this K / V^T set
          }
          __syncwarp();
          if (next && !defer) {  // S_t of the next item right behind P V_t (in-order pipe: P_t has been consumed)
            if (t == 0) {
              if (chg) mbar_wait(&kv_full[st_n], (uint32_t)((((gi + 1) / a.kv_stages)) & 1), 20);
              mbar_wait(&q_full[qb_n], (uint32_t)(((i + 1) >> 1) & 1), 21);
              tc_fence_after();
            }
            issue_s(t, qb_n, st_n);
            if (t == NT - 1) {
              if (elect_one()) umma_commit(&q_empty[qb_n]);
              __syncwarp();
            }
          }
        }
        if (next && defer) {
          mbar_wait(&kv_full[st_n], (uint32_t)((((gi + 1) / a.kv_stages)) & 1), 20);
          mbar_wait(&q_full[qb_n], (uint32_t)(((i + 1) >> 1) & 1), 21);
          tc_fence_after();
          for (int t = 0; t < NT; ++t) issue_s(t, qb_n, st_n);
          if (elect_one()) umma_commit(&q_empty[qb_n]);
          __syncwarp();
        }
        if (chg) {
          ++gi;
          g_cur = g_next;
        }
      }
    }
  } else {
    // ================= softmax warpgroups: thread <-> query row =================
    const int t = (warp - 2) >> 2;  // tile
    const int qd = warp & 3;        // TMEM lane quarter of this warp
    const int r = qd * 32 + lane;
    const uint32_t lane_off = (uint32_t)(qd * 32) << 16;
    const uint32_t tS = tmem_base + t * BKN + lane_off;
    const uint32_t tO = tmem_base + colO + t * a.dv + lane_off;
    for (int i = 0; i < n_my; ++i) {
      const Item5 it = decode_item(a, w0 + i);
      mbar_wait(&s_full[t], (uint32_t)(i & 1), 30);
      tc_fence_after();
      uint32_t s[BKN];
#pragma unroll
      for (int g = 0; g < NG; ++g) tmem_ld16(tS + g * 16, reinterpret_cast<uint32_t(&)[16]>(s[g * 16]));
      tmem_ld_wait();
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
#pragma unroll
      for (int c = 0; c + 16 <= BKN / 2; c += 16) tmem_st16(tS + c, reinterpret_cast<const uint32_t(&)[16]>(s[c]));
      if constexpr ((BKN / 2) % 16 == 8) tmem_st8(tS + (BKN / 2 - 8), &s[BKN / 2 - 8]);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_full[t]);
      // ---- epilogue: O / l -> global (l = O[:, d], accumulated by the ones row of V^T) ----
      mbar_wait(&pv_done[t], (uint32_t)(i & 1), 32);
      tc_fence_after();
      const int qrow = it.qp * (NT * BQ) + t * BQ + r;
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
          uint4* dst = reinterpret_cast<uint4*>(a.o + ((int64_t)it.b * a.nq + qrow) * a.ldo + it.head * a.dqk + c);
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
    tmem_dealloc(tmem_base, 256);
  }
}

template <int NT, int NG, bool F16>
cpd_status launch_attention5(const cpd_attn_params* p, int dv, void* stream) {
  constexpr int BKN = NG * 16;
  const int d = p->d_head;
  Attn5Args a;
  a.o = (bf16*)p->o;
  a.ldo = p->ldo;
  a.batch = p->batch; a.heads = p->heads; a.nq = p->nq; a.nk = p->nk; a.nk_pad = p->nk_pad;
  a.kv_batch = p->kv_batch > 0 ? p->kv_batch : p->batch;
  a.dqk = p->dpad;
  a.d = d;
  a.dv = dv;
  a.datoms = (p->dpad + 63) / 64;
  a.scale_log2 = p->scale * 1.4426950408889634f;
  a.nqp = (p->nq + NT * BQ - 1) / (NT * BQ);
  a.reps = p->batch / a.kv_batch;
  a.items = p->heads * a.kv_batch * a.reps * a.nqp;
  const int q_bytes = 2 * NT * a.datoms * ATOM_BYTES;
  const int kv_set = a.datoms * BKN * 128 + ((BKN + 63) / 64) * dv * 128;
  const int budget = 113 * 1024 - 1024 - 256;  // two CTAs per SM
  if (q_bytes + kv_set > budget) return CPD_ERR_UNSUPPORTED;
  a.kv_stages = (q_bytes + 2 * kv_set <= budget) ? 2 : 1;
  const int ctas = a.items < 2 * 148 ? a.items : 2 * 148;
  a.ipc = (a.items + ctas - 1) / ctas;
  const int grid = (a.items + a.ipc - 1) / a.ipc;
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
    uint32_t box[2] = {64, BKN};
    if ((rc = cpd_make_tmap_bf16(&a.map_k, p->k, 2, dims, str, box))) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)a.kv_batch * p->nk_pad, (uint64_t)p->heads * p->dpad};
    uint64_t str[1] = {(uint64_t)p->ldvt * 2};
    uint32_t box[2] = {64, (uint32_t)d};
    if ((rc = cpd_make_tmap_bf16(&a.map_vt, p->vt, 2, dims, str, box))) return rc;
  }
  const size_t shm = (size_t)q_bytes + (size_t)a.kv_stages * kv_set + 256 + 1024;
  CPD_SMEM_OPTIN((attention5_kernel<NT, NG, F16>), 113 * 1024);
  CPD_CUDA_CHECK(cpd_launch(attention5_kernel<NT, NG, F16>, dim3(grid), dim3(64 + NT * 128), shm, (cudaStream_t)stream, a));
  return CPD_OK;
}

}  // namespace

// Returns CPD_ERR_UNSUPPORTED when the shape is outside this kernel's domain (the caller falls back).
cpd_status cpd_attention_cross(const cpd_attn_params* p, void* stream) {
  const int d = p->d_head;
  if (d <= 0 || d > p->dpad || p->nq < BQ) return CPD_ERR_UNSUPPORTED;
  const int bkn = (p->nk + 15) / 16 * 16;
  if (bkn != 80 || bkn > p->nk_pad) return CPD_ERR_UNSUPPORTED;  // instantiated for the 77-token context (5 x 16 key columns)
  const int kvb = p->kv_batch > 0 ? p->kv_batch : p->batch;
  if (p->batch % kvb) return CPD_ERR_UNSUPPORTED;
  const int dv = (d + 1 + 15) / 16 * 16;
  const bool f16 = p->act_fp16 != 0;
  if (2 * (bkn + dv) <= 256 && p->nq >= 2 * BQ)  // two tiles per CTA (SD-1.x d = 40)
    return f16 ? launch_attention5<2, 5, true>(p, dv, stream) : launch_attention5<2, 5, false>(p, dv, stream);
  if (bkn + dv <= 256)  // one tile per CTA (head dims 64 / 80)
    return f16 ? launch_attention5<1, 5, true>(p, dv, stream) : launch_attention5<1, 5, false>(p, dv, stream);
  return CPD_ERR_UNSUPPORTED;
}
