// GroupNorm(32)(+SiLU) and LayerNorm over NHWC bf16 activations: HBM-bound, 16-byte vectorised,
// fp32 statistics (models/util.py:95-105; attention.py:89-90,476-478 of the reference).
#include "../../include/cpd_b200.h"
#include "common.cuh"

namespace {

constexpr int GROUPS = 32;

// ---- GroupNorm pass 1: per-(image, pixel chunk, group) sum / sum of squares ----------------------------------------
// grid = (chunks, n_img).  Thread t owns channel vector (t % vec_per_px) and pixel lane (t / vec_per_px).  No atomics
// anywhere: per-thread partials go to shared memory, are summed over the pixel lanes in a fixed order, reduced per group in
// fp64 and written to partial[n][chunk][group][2]; pass 2 sums the chunks in a fixed order.  Results are bit-reproducible.
__global__ void __launch_bounds__(512, 2) gn_stats_kernel(const bf16* __restrict__ a0, const bf16* __restrict__ a1, int c0,
                                                          int c1, int hw, int px_per_block, int f16, double* __restrict__ partial) {
  extern __shared__ float sh[];  // [lanes][2][C]
  pdl_launch_dependents();
  pdl_wait();
  const int C = c0 + c1;
  const int vec_per_px = C / 8;
  const int lanes = blockDim.x / vec_per_px;
  const int n = blockIdx.y;
  const int cv = threadIdx.x % vec_per_px;
  const int pl = threadIdx.x / vec_per_px;
  if (pl < lanes) {
    const int ch = cv * 8;
    const bf16* src;
    int cs, coff;
    if (ch < c0) { src = a0; cs = c0; coff = ch; } else { src = a1; cs = c1; coff = ch - c0; }
    float s[8], ss[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) s[e] = ss[e] = 0.f;
    const int p0 = blockIdx.x * px_per_block;
    const int p1 = min(p0 + px_per_block, hw);
    const bf16* base = src + (int64_t)n * hw * cs + coff;
    int p = p0 + pl;
    for (; p + 7 * lanes < p1; p += 8 * lanes) {  // 8 independent 16-byte loads in flight per thread
      uint4 v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = __ldg(reinterpret_cast<const uint4*>(base + (int64_t)(p + j * lanes) * cs));
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t u[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = unpack_act2(u[e], f16 != 0);
          s[2 * e] += f.x; ss[2 * e] += f.x * f.x;
          s[2 * e + 1] += f.y; ss[2 * e + 1] += f.y * f.y;
        }
      }
    }
    for (; p < p1; p += lanes) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(base + (int64_t)p * cs));
      const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = unpack_act2(u[e], f16 != 0);
        s[2 * e] += f.x; ss[2 * e] += f.x * f.x;
        s[2 * e + 1] += f.y; ss[2 * e + 1] += f.y * f.y;
      }
    }
    float* mine = sh + (size_t)pl * 2 * C;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      mine[ch + e] = s[e];
      mine[C + ch + e] = ss[e];
    }
  }
  __syncthreads();
  const int cpg = C / GROUPS;
  if (threadIdx.x < GROUPS) {
    double S = 0.0, SS = 0.0;
    for (int c = threadIdx.x * cpg; c < (threadIdx.x + 1) * cpg; ++c) {
      for (int l = 0; l < lanes; ++l) {  // fixed order: lane 0, 1, ...
        S += (double)sh[(size_t)l * 2 * C + c];
        SS += (double)sh[(size_t)l * 2 * C + C + c];
      }
    }
    double* out = partial + (((int64_t)n * gridDim.x + blockIdx.x) * GROUPS + threadIdx.x) * 2;
    out[0] = S;
    out[1] = SS;
  }
}

// ---- GroupNorm pass 2: normalise, affine, optional SiLU, write bf16 ---------------------------------
__global__ void __launch_bounds__(512, 2) gn_apply_kernel(const bf16* __restrict__ a0, const bf16* __restrict__ a1, int c0,
                                                       int c1, int hw, int px_per_block, const double* __restrict__ partial,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       float eps, int silu, int f16, bf16* __restrict__ out) {
  extern __shared__ float sh[];  // scale[C], shift[C]
  pdl_launch_dependents();
  pdl_wait();
  const int C = c0 + c1;
  const int vec_per_px = C / 8;
  const int lanes = blockDim.x / vec_per_px;
  const int n = blockIdx.y;
  const int cpg = C / GROUPS;
  const double cnt = (double)hw * cpg;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / cpg;
    double S = 0.0, SS = 0.0;
    for (int k = 0; k < (int)gridDim.x; ++k) {  // chunk partials of pass 1, fixed order
      const double* pp = partial + (((int64_t)n * gridDim.x + k) * GROUPS + g) * 2;
      S += pp[0];
      SS += pp[1];
    }
    const double mean = S / cnt;
    double var = SS / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    const float rstd = (float)(1.0 / sqrt(var + (double)eps));
    const float sc = rstd * gamma[c];
    sh[c] = sc;
    sh[C + c] = beta[c] - (float)mean * sc;
  }
  __syncthreads();
  const int cv = threadIdx.x % vec_per_px;
  const int pl = threadIdx.x / vec_per_px;
  if (pl >= lanes) return;
  const int ch = cv * 8;
  const bf16* src;
  int cs, coff;
  if (ch < c0) { src = a0; cs = c0; coff = ch; } else { src = a1; cs = c1; coff = ch - c0; }
  float sc[8], sf[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { sc[e] = sh[ch + e]; sf[e] = sh[C + ch + e]; }
  const int p0 = blockIdx.x * px_per_block;
  const int p1 = min(p0 + px_per_block, hw);
  const bf16* base = src + (int64_t)n * hw * cs + coff;
  bf16* obase = out + (int64_t)n * hw * C + ch;
  auto apply_one = [&](const uint4& v, int p) {
    const uint32_t u[4] = {v.x, v.y, v.z, v.w};
    uint32_t o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f = unpack_act2(u[e], f16 != 0);
      float y0 = f.x * sc[2 * e] + sf[2 * e];
      float y1 = f.y * sc[2 * e + 1] + sf[2 * e + 1];
      if (silu) { y0 = silu_f(y0); y1 = silu_f(y1); }
      o[e] = pack_act2(y0, y1, f16 != 0);
    }
    *reinterpret_cast<uint4*>(obase + (int64_t)p * C) = make_uint4(o[0], o[1], o[2], o[3]);
  };
  int p = p0 + pl;
  for (; p + 7 * lanes < p1; p += 8 * lanes) {  // 8 independent 16-byte loads in flight per thread
    uint4 v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = __ldg(reinterpret_cast<const uint4*>(base + (int64_t)(p + j * lanes) * cs));
#pragma unroll
    for (int j = 0; j < 8; ++j) apply_one(v[j], p + j * lanes);
  }
  for (; p < p1; p += lanes) apply_one(__ldg(reinterpret_cast<const uint4*>(base + (int64_t)p * cs)), p);
}

// ---- LayerNorm: one warp per R rows (all loads of the R rows issued before the first use), rows in registers -----
template <int MAXV, int R>  // max 16-byte vectors per lane, rows per warp
__global__ void __launch_bounds__(256, 4) layernorm_kernel(const bf16* __restrict__ x, int rows, int c,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        float eps, int f16, bf16* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int row0 = warp * R;
  if (row0 >= rows) return;
  const int nvec = c / 8;
  uint4 raw[R][MAXV];
#pragma unroll
  for (int rr = 0; rr < R; ++rr)
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int vi = lane + i * 32;
      raw[rr][i] = make_uint4(0, 0, 0, 0);
      if (vi < nvec && row0 + rr < rows) raw[rr][i] = __ldg(reinterpret_cast<const uint4*>(x + (int64_t)(row0 + rr) * c + vi * 8));
    }
#pragma unroll
  for (int rr = 0; rr < R; ++rr) {
    if (row0 + rr >= rows) break;
    float v[MAXV][8];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const uint32_t w[4] = {raw[rr][i].x, raw[rr][i].y, raw[rr][i].z, raw[rr][i].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = unpack_act2(w[e], f16 != 0);
        v[i][2 * e] = f.x; v[i][2 * e + 1] = f.y;
        sum += f.x + f.y;  // lanes beyond nvec hold zeros
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum / (float)c;
    float var = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int vi = lane + i * 32;
      if (vi < nvec) {
#pragma unroll
        for (int e = 0; e < 8; ++e) { const float d = v[i][e] - mean; var += d * d; }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
    const float rstd = rsqrtf(var / (float)c + eps);
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int vi = lane + i * 32;
      if (vi < nvec) {
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + vi * 8));
        const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + vi * 8 + 4));
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + vi * 8));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + vi * 8 + 4));
        const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        uint32_t o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e)
          o[e] = pack_act2((v[i][2 * e] - mean) * rstd * gg[2 * e] + bb[2 * e],
                           (v[i][2 * e + 1] - mean) * rstd * gg[2 * e + 1] + bb[2 * e + 1], f16 != 0);
        *reinterpret_cast<uint4*>(out + (int64_t)(row0 + rr) * c + vi * 8) = make_uint4(o[0], o[1], o[2], o[3]);
      }
    }
  }
}

}  // namespace

extern "C" cpd_status cpd_groupnorm(const void* a0, const void* a1, int c0, int c1, int n_img, int hw, const float* gamma,
                                    const float* beta, float eps, int silu, int act_fp16, double* stats, void* out, void* stream) {
  CPD_REQUIRE(a0 && gamma && beta && stats && out, "cpd_groupnorm: null pointer");
  const int C = c0 + c1;
  CPD_REQUIRE(c0 > 0 && c1 >= 0 && c0 % 8 == 0 && c1 % 8 == 0 && C % GROUPS == 0 && C <= 4096,
              "cpd_groupnorm: bad channel counts c0=%d c1=%d", c0, c1);
  CPD_REQUIRE(c1 == 0 || a1, "cpd_groupnorm: c1 > 0 needs a1");
  CPD_REQUIRE(n_img > 0 && hw > 0, "cpd_groupnorm: empty input");
  cudaStream_t s = (cudaStream_t)stream;
  const int vec_per_px = C / 8;
  // <= 256 threads per block (4 blocks per SM at <= 64 registers), 8 x 16-byte loads in flight per thread
  int lanes = 256 / vec_per_px;
  if (lanes < 1) lanes = 1;
  if (lanes > hw) lanes = hw;
  const int threads = ((lanes * vec_per_px + 31) / 32) * 32;
  CPD_REQUIRE(threads <= 512, "cpd_groupnorm: C=%d too large", C);
  // one wave of ~4 blocks per SM over all images, but at least 8 pixels per pixel lane
  int chunks = (148 * 4 + n_img - 1) / n_img;
  int px_per_block = (hw + chunks - 1) / chunks;
  if (px_per_block < 8 * lanes) px_per_block = 8 * lanes;
  chunks = (hw + px_per_block - 1) / px_per_block;
  if (chunks > CPD_GN_MAX_CHUNKS) {  // the scratch holds CPD_GN_MAX_CHUNKS partials per image
    px_per_block = (hw + CPD_GN_MAX_CHUNKS - 1) / CPD_GN_MAX_CHUNKS;
    chunks = (hw + px_per_block - 1) / px_per_block;
  }
  const size_t shm = sizeof(float) * 2 * C;
  const size_t shm_stats = sizeof(float) * 2 * C * lanes;
  CPD_REQUIRE(shm_stats <= 160 * 1024, "cpd_groupnorm: C=%d needs %zu bytes of shared memory", C, shm_stats);
  static bool cfg = false;
  if (!cfg) {
    CPD_CUDA_CHECK(cudaFuncSetAttribute(gn_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    cfg = true;
  }
  CPD_CUDA_CHECK(cpd_launch(gn_stats_kernel, dim3(dim3(chunks, n_img)), dim3(threads), shm_stats, s, (const bf16*)a0, (const bf16*)a1, c0, c1, hw, px_per_block, act_fp16, stats));
  CPD_CUDA_CHECK(cudaGetLastError());
  CPD_CUDA_CHECK(cpd_launch(gn_apply_kernel, dim3(dim3(chunks, n_img)), dim3(threads), shm, s, (const bf16*)a0, (const bf16*)a1, c0, c1, hw, px_per_block, stats,
                                                           gamma, beta, eps, silu, act_fp16, (bf16*)out));
  CPD_CUDA_CHECK(cudaGetLastError());
  return CPD_OK;
}

extern "C" cpd_status cpd_layernorm(const void* x, int rows, int c, const float* gamma, const float* beta, float eps,
                                    int act_fp16, void* out, void* stream) {
  CPD_REQUIRE(x && gamma && beta && out, "cpd_layernorm: null pointer");
  CPD_REQUIRE(c > 0 && c % 8 == 0 && c <= 2048, "cpd_layernorm: c=%d must be a multiple of 8 and <= 2048", c);
  CPD_REQUIRE(rows >= 0, "cpd_layernorm: rows=%d", rows);
  if (rows == 0) return CPD_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const int nvec = c / 8;
  auto blocks = [&](int r_per_warp) { return (rows + 8 * r_per_warp - 1) / (8 * r_per_warp); };
  if (nvec <= 64) CPD_CUDA_CHECK(cpd_launch(layernorm_kernel<2, 4>, dim3(blocks(4)), dim3(256), 0, s, (const bf16*)x, rows, c, gamma, beta, eps, act_fp16, (bf16*)out));
  else if (nvec <= 160) CPD_CUDA_CHECK(cpd_launch(layernorm_kernel<5, 2>, dim3(blocks(2)), dim3(256), 0, s, (const bf16*)x, rows, c, gamma, beta, eps, act_fp16, (bf16*)out));
  else CPD_CUDA_CHECK(cpd_launch(layernorm_kernel<8, 1>, dim3(blocks(1)), dim3(256), 0, s, (const bf16*)x, rows, c, gamma, beta, eps, act_fp16, (bf16*)out));
  CPD_CUDA_CHECK(cudaGetLastError());
  return CPD_OK;
}
