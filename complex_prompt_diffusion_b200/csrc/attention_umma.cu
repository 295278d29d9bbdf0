// Fused flash-style attention for sm_100a: S = Q K^T and O += P V on tcgen05 with both accumulators in
// TMEM, Q/K/V^T tiles staged by TMA into 128B-swizzled shared memory, online softmax in registers
// (one thread owns one query row = one TMEM lane, so row max / row sum need no shuffles), lazy rescaling
// of the O accumulator in TMEM.  The [heads*B, N, N] score matrix of the reference
// (attention.py:327-340) is never materialised.
//
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner, warps 2-5 = softmax / correction /
// epilogue (128 threads <-> 128 query rows).
#include <stdlib.h>

#include "../../include/cpd_b200.h"
#include "common.cuh"

namespace {

constexpr int BQ = 128;    // query rows per CTA
constexpr int BKV = 128;   // keys per block
constexpr int ATOM_BYTES = 128 * 128;  // 128 rows x 64 bf16
constexpr int NUM_THREADS = 192;
constexpr float RESCALE_TAU = 8.0f;    // lazy O rescale threshold (log2 domain)

struct AttnArgs {
  CUtensorMap map_q, map_k, map_vt;
  bf16* o;
  int ldo;
  int batch, heads, nq, nk, nk_pad, dpad, kv_batch;
  int datoms;      // ceil(dpad / 64)
  int kv_stages;   // 1 or 2
  int fp16;        // q, k, vt, o element type: 1 = fp16, 0 = bf16
  float scale_log2;
};

__global__ void __launch_bounds__(NUM_THREADS, 2) attention_kernel(const __grid_constant__ AttnArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int datoms = a.datoms;
  const int dpad = a.dpad;
  const int vt_atom_bytes = dpad * 128;
  const int k_stage_bytes = datoms * ATOM_BYTES;
  const int v_stage_bytes = 2 * vt_atom_bytes;
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + datoms * ATOM_BYTES;
  uint8_t* sV = sK + a.kv_stages * k_stage_bytes;
  uint8_t* sP = sV + a.kv_stages * v_stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * ATOM_BYTES);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;    // [2]
  uint64_t* k_empty = bars + 3;   // [2]
  uint64_t* v_full = bars + 5;    // [2]
  uint64_t* v_empty = bars + 7;   // [2]
  uint64_t* s_full = bars + 9;
  uint64_t* p_full = bars + 10;
  uint64_t* o_done = bars + 11;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * BQ;
  const int head = blockIdx.y;
  const int b = blockIdx.z;
  const int bkv = b % a.kv_batch;
  const int nblk = (a.nk + BKV - 1) / BKV;
  const uint32_t tmem_cols = (dpad <= 128) ? 256u : 512u;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&a.map_q);
    tma_prefetch_desc(&a.map_k);
    tma_prefetch_desc(&a.map_vt);
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(o_done, 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_smem, tmem_cols);
    tmem_relinquish();
  }
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();
  const uint32_t tmem_S = tmem_base;
  const uint32_t tmem_O = tmem_base + 128;

  if (warp == 0) {
    if (lane == 0) {
      // ---- TMA producer ----
      mbar_arrive_expect_tx(q_full, datoms * ATOM_BYTES);
      for (int d = 0; d < datoms; ++d)
        tma_load_2d(sQ + d * ATOM_BYTES, &a.map_q, q_full, head * dpad + d * 64, b * a.nq + q0);
      for (int j = 0; j < nblk; ++j) {
        const int st = j % a.kv_stages;
        const uint32_t ph = (uint32_t)((j / a.kv_stages) & 1);
        mbar_wait(&k_empty[st], ph ^ 1, 10);
        mbar_arrive_expect_tx(&k_full[st], k_stage_bytes);
        for (int d = 0; d < datoms; ++d)
          tma_load_2d(sK + st * k_stage_bytes + d * ATOM_BYTES, &a.map_k, &k_full[st], head * dpad + d * 64,
                      bkv * a.nk_pad + j * BKV);
        mbar_wait(&v_empty[st], ph ^ 1, 11);
        mbar_arrive_expect_tx(&v_full[st], v_stage_bytes);
        for (int t = 0; t < 2; ++t)
          tma_load_2d(sV + st * v_stage_bytes + t * vt_atom_bytes, &a.map_vt, &v_full[st], bkv * a.nk_pad + j * BKV + t * 64,
                      head * dpad);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ---- MMA issuer ----
      const uint32_t idesc_s = umma_idesc_f16(BQ, BKV, a.fp16 != 0, a.fp16 != 0);
      const uint32_t idesc_o = umma_idesc_f16(BQ, dpad, a.fp16 != 0, a.fp16 != 0);
      const int ksteps_s = dpad / 16;
      auto issue_s = [&](int j) {
        const int st = j % a.kv_stages;
        mbar_wait(&k_full[st], (uint32_t)((j / a.kv_stages) & 1), 20);
        tc_fence_after();
        const uint32_t qa = smem_u32(sQ), ka = smem_u32(sK + st * k_stage_bytes);
        for (int kk = 0; kk < ksteps_s; ++kk) {
          const uint32_t off = (uint32_t)((kk >> 2) * ATOM_BYTES + (kk & 3) * 32);
          umma_bf16(tmem_S, umma_desc_sw128(qa + off), umma_desc_sw128(ka + off), idesc_s, kk > 0 ? 1u : 0u);
        }
        umma_commit(s_full);
        umma_commit(&k_empty[st]);
      };
      mbar_wait(q_full, 0, 21);
      issue_s(0);
      for (int j = 0; j < nblk; ++j) {
        const int st = j % a.kv_stages;
        mbar_wait(p_full, (uint32_t)(j & 1), 22);  // P_j written, S_j consumed
        tc_fence_after();
        if (j + 1 < nblk) issue_s(j + 1);
        mbar_wait(&v_full[st], (uint32_t)((j / a.kv_stages) & 1), 23);
        tc_fence_after();
        const int kv_valid = min(BKV, a.nk - j * BKV);
        const int ksteps_o = (kv_valid + 15) / 16;
        const uint32_t pa = smem_u32(sP), va = smem_u32(sV + st * v_stage_bytes);
        for (int kk = 0; kk < ksteps_o; ++kk) {
          const uint32_t offp = (uint32_t)((kk >> 2) * ATOM_BYTES + (kk & 3) * 32);
          const uint32_t offv = (uint32_t)((kk >> 2) * vt_atom_bytes + (kk & 3) * 32);
          umma_bf16(tmem_O, umma_desc_sw128(pa + offp), umma_desc_sw128(va + offv), idesc_o, (j > 0 || kk > 0) ? 1u : 0u);
        }
        umma_commit(o_done);
        umma_commit(&v_empty[st]);
      }
    }
  } else {
    // ---- softmax / correction / epilogue: thread <-> query row ----
    const int qd = warp & 3;
    const int r = qd * 32 + lane;
    const uint32_t lane_off = (uint32_t)(qd * 32) << 16;
    float m_used = -INFINITY;
    float l = 0.f;
    for (int j = 0; j < nblk; ++j) {
      const int kv_valid = min(BKV, a.nk - j * BKV);
      mbar_wait(s_full, (uint32_t)(j & 1), 30);
      tc_fence_after();
      // pass 1 over the S row (TMEM reads are cheap: the row is re-read in pass 2 instead of living in 128 registers)
      const bool full = (kv_valid == BKV);
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t t[32];
        tmem_ld32(tmem_S + lane_off + c * 32, t);
        tmem_ld_wait();
        if (full) {
#pragma unroll
          for (int e = 0; e < 32; ++e) mx = fmaxf(mx, __uint_as_float(t[e]));
        } else {
#pragma unroll
          for (int e = 0; e < 32; ++e)
            if (c * 32 + e < kv_valid) mx = fmaxf(mx, __uint_as_float(t[e]));
        }
      }
      const float m_blk = mx * a.scale_log2;
      float alpha = 1.0f;
      bool need = false;
      if (j == 0) {
        m_used = m_blk;
      } else if (m_blk > m_used + RESCALE_TAU) {
        alpha = exp2f(m_used - m_blk);
        m_used = m_blk;
        need = true;
      }
      // pass 2: p = 2^(s * scale_log2 - m) with the MUFU ex2 fed by one FFMA; packed to 16-bit in registers
      float psum0 = 0.f, psum1 = 0.f;
      uint32_t pk[64];
      const float neg_m = -m_used;
      const bool f16 = a.fp16 != 0;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t t[32];
        tmem_ld32(tmem_S + lane_off + c * 32, t);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          float p0 = fast_ex2(fmaf(__uint_as_float(t[e]), a.scale_log2, neg_m));
          float p1 = fast_ex2(fmaf(__uint_as_float(t[e + 1]), a.scale_log2, neg_m));
          if (!full) {
            if (c * 32 + e >= kv_valid) p0 = 0.f;
            if (c * 32 + e + 1 >= kv_valid) p1 = 0.f;
          }
          psum0 += p0;
          psum1 += p1;
          pk[c * 16 + (e >> 1)] = pack_act2(p0, p1, f16);
        }
      }
      l = l * alpha + (psum0 + psum1);
      if (j > 0) {
        // previous P V must be complete before P is overwritten / O is rescaled
        mbar_wait(o_done, (uint32_t)((j - 1) & 1), 31);
        tc_fence_after();
        if (__any_sync(0xffffffffu, need)) {
          for (int c = 0; c < dpad; c += 16) {
            uint32_t t[16];
            tmem_ld16(tmem_O + lane_off + c, t);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 16; ++e) t[e] = __float_as_uint(__uint_as_float(t[e]) * alpha);
            tmem_st16(tmem_O + lane_off + c, t);
          }
          tmem_st_wait();
        }
      }
      // P row -> shared memory, K-major 128B-swizzled (16-byte chunk c of row r lives at chunk c ^ (r & 7))
#pragma unroll
      for (int ch = 0; ch < 16; ++ch) {
        uint8_t* dst = sP + (ch >> 3) * ATOM_BYTES + r * 128 + (((ch & 7) ^ (r & 7)) << 4);
        *reinterpret_cast<uint4*>(dst) = make_uint4(pk[ch * 4], pk[ch * 4 + 1], pk[ch * 4 + 2], pk[ch * 4 + 3]);
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(p_full);
    }
    mbar_wait(o_done, (uint32_t)((nblk - 1) & 1), 32);
    tc_fence_after();
    const float inv_l = 1.0f / l;
    const int qrow = q0 + r;
    for (int c = 0; c < dpad; c += 16) {
      uint32_t t[16];
      tmem_ld16(tmem_O + lane_off + c, t);
      tmem_ld_wait();
      if (qrow < a.nq) {
        uint32_t o[8];
#pragma unroll
        for (int e = 0; e < 8; ++e)
          o[e] = pack_act2(__uint_as_float(t[2 * e]) * inv_l, __uint_as_float(t[2 * e + 1]) * inv_l, a.fp16 != 0);
        uint4* dst = reinterpret_cast<uint4*>(a.o + ((int64_t)b * a.nq + qrow) * a.ldo + head * dpad + c);
        dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
        dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

}  // namespace

cpd_status cpd_attention_persistent(const cpd_attn_params* p, void* stream);  // attention_umma3.cu
cpd_status cpd_attention_split(const cpd_attn_params* p, void* stream);       // attention_umma4.cu
cpd_status cpd_attention_cross(const cpd_attn_params* p, void* stream);       // attention_umma5.cu

static bool env_off(const char* name) {
  const char* e = getenv(name);
  return e && e[0] == '0';
}

extern "C" cpd_status cpd_attention(const cpd_attn_params* p, void* stream) {
  CPD_REQUIRE(p && p->q && p->k && p->vt && p->o, "cpd_attention: null pointer");
  CPD_REQUIRE(p->dpad >= 16 && p->dpad % 16 == 0 && p->dpad <= 160, "cpd_attention: dpad=%d must be a multiple of 16 in [16,160]", p->dpad);
  CPD_REQUIRE(p->batch > 0 && p->heads > 0 && p->nq > 0 && p->nk > 0 && p->nk_pad >= p->nk, "cpd_attention: bad sizes");
  CPD_REQUIRE(p->ldq % 8 == 0 && p->ldk % 8 == 0 && p->ldvt % 8 == 0 && p->ldo % 8 == 0, "cpd_attention: leading dims must be multiples of 8");
  CPD_REQUIRE(((uintptr_t)p->o & 15) == 0, "cpd_attention: o must be 16-byte aligned");
  CPD_REQUIRE(p->d_head >= 0 && p->d_head <= p->dpad, "cpd_attention: d_head=%d must be in [0, dpad=%d]", p->d_head, p->dpad);
  if (p->d_head > 0) {
    // Dispatch by shape (each kernel returns CPD_ERR_UNSUPPORTED outside its domain): one key block (the 77-token context)
    // -> attention5; head dims <= 63 with many key blocks (the 4096^2 d = 40 self-attention) -> attention4 (split rows);
    // everything else with more than 128 query rows -> attention3 (persistent two-tile); the rest -> the one-tile kernel below.
    // CPD_ATTN_CROSS=0 / CPD_ATTN_SPLIT=0 / CPD_ATTN_PERSIST=0 take a kernel out of the chain (A/B measurements).
    static const bool cross = !env_off("CPD_ATTN_CROSS"), split = !env_off("CPD_ATTN_SPLIT"), persist = !env_off("CPD_ATTN_PERSIST");
    if (cross) {
      const cpd_status st5 = cpd_attention_cross(p, stream);
      if (st5 != CPD_ERR_UNSUPPORTED) return st5;
    }
    if (split) {
      const cpd_status st4 = cpd_attention_split(p, stream);
      if (st4 != CPD_ERR_UNSUPPORTED) return st4;
    }
    if (persist) {
      const cpd_status st3 = cpd_attention_persistent(p, stream);
      if (st3 != CPD_ERR_UNSUPPORTED) return st3;
    }
  }
  AttnArgs a;
  a.o = (bf16*)p->o;
  a.ldo = p->ldo;
  a.batch = p->batch; a.heads = p->heads; a.nq = p->nq; a.nk = p->nk; a.nk_pad = p->nk_pad; a.dpad = p->dpad;
  a.kv_batch = p->kv_batch > 0 ? p->kv_batch : p->batch;
  a.datoms = (p->dpad + 63) / 64;
  a.kv_stages = (p->dpad <= 80) ? 2 : 1;
  a.fp16 = p->act_fp16;
  a.scale_log2 = p->scale * 1.4426950408889634f;
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
    uint32_t box[2] = {64, (uint32_t)p->dpad};
    if ((rc = cpd_make_tmap_bf16(&a.map_vt, p->vt, 2, dims, str, box))) return rc;
  }
  const size_t shm = (size_t)a.datoms * ATOM_BYTES + (size_t)a.kv_stages * (a.datoms * ATOM_BYTES + 2 * p->dpad * 128) +
                     2 * ATOM_BYTES + 256 + 1024;
  CPD_REQUIRE(shm <= 227 * 1024, "cpd_attention: shared memory %zu exceeds 227 KB", shm);
  CPD_SMEM_OPTIN(attention_kernel, 227 * 1024);
  dim3 grid((p->nq + BQ - 1) / BQ, p->heads, p->batch);
  CPD_CUDA_CHECK(cpd_launch(attention_kernel, dim3(grid), dim3(NUM_THREADS), shm, (cudaStream_t)stream, a));
  CPD_CUDA_CHECK(cudaGetLastError());
  return CPD_OK;
}
