// Thresholding extensions on the device: exact per-image percentile of |x| (radix select) + clamp.
// Replaces the CPU round trip of cpd/samplers/extension/threshold.py:65-88 (np.percentile on x.cpu() every step) and
// :47-63 (static clamp); both return x.half(), i.e. values rounded through fp16.
#include <math.h>

#include "../../include/cpd_b200.h"
#include "common.cuh"

namespace {

// One CTA per image.  |x| as a non-negative float orders like its bit pattern, so the k-th smallest value is found by a
// 4-pass most-significant-digit radix select (256-bin shared-memory histogram of the elements that match the prefix found
// so far; integer atomics: the counts are exact and order-independent).  The two order statistics around the virtual index
// (n - 1) * q / 100 are interpolated in fp32 the way numpy's _lerp does.
__global__ void __launch_bounds__(1024) abs_percentile_kernel(const float* __restrict__ x, int L, int k_lo, int k_hi, float t,
                                                              float* __restrict__ bound) {
  __shared__ unsigned hist[256];
  __shared__ unsigned s_prefix, s_rank;
  pdl_launch_dependents();
  pdl_wait();
  const float* xi = x + (int64_t)blockIdx.x * L;
  float stat[2];
  for (int which = 0; which < 2; ++which) {
    if (which == 1 && k_hi == k_lo) {
      stat[1] = stat[0];
      break;
    }
    if (threadIdx.x == 0) {
      s_prefix = 0u;
      s_rank = (unsigned)(which ? k_hi : k_lo);
    }
    unsigned mask = 0u;
    for (int pass = 3; pass >= 0; --pass) {
      for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0u;
      __syncthreads();
      const unsigned prefix = s_prefix;
      for (int i = threadIdx.x; i < L; i += blockDim.x) {
        const unsigned key = __float_as_uint(xi[i]) & 0x7fffffffu;
        if ((key & mask) == prefix) atomicAdd(&hist[(key >> (8 * pass)) & 255u], 1u);
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        unsigned rank = s_rank, cum = 0u;
        int bkt = 0;
        for (; bkt < 255; ++bkt) {
          if (rank < cum + hist[bkt]) break;
          cum += hist[bkt];
        }
        s_rank = rank - cum;
        s_prefix = prefix | ((unsigned)bkt << (8 * pass));
      }
      mask |= 0xffu << (8 * pass);
      __syncthreads();
    }
    stat[which] = __uint_as_float(s_prefix);
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float a = stat[0], b = stat[1];
    const float diff = __fsub_rn(b, a);
    float r = __fadd_rn(a, __fmul_rn(diff, t));                                  // numpy _lerp
    if (t >= 0.5f) r = __fsub_rn(b, __fmul_rn(diff, __fsub_rn(1.0f, t)));
    bound[blockIdx.x] = fmaxf(r, 1.0f);                                           // np.max(np.append(s, 1.0))
  }
}

__global__ void fill_bound_kernel(float* bound, int n, float v) {
  pdl_launch_dependents();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) bound[i] = v;
}

// x <- half(clamp(x.float(), -s, s)), kept as fp32 values (every later op of the loop promotes to fp32 anyway)
__global__ void __launch_bounds__(256) clamp_half_kernel(float* __restrict__ x, int L4, const float* __restrict__ bound) {
  pdl_launch_dependents();
  pdl_wait();
  const float s = __ldg(bound + blockIdx.y);
  float4* xi = reinterpret_cast<float4*>(x) + (int64_t)blockIdx.y * L4;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < L4; i += gridDim.x * blockDim.x) {
    float4 v = xi[i];
    v.x = __half2float(__float2half_rn(fminf(fmaxf(v.x, -s), s)));
    v.y = __half2float(__float2half_rn(fminf(fmaxf(v.y, -s), s)));
    v.z = __half2float(__float2half_rn(fminf(fmaxf(v.z, -s), s)));
    v.w = __half2float(__float2half_rn(fminf(fmaxf(v.w, -s), s)));
    xi[i] = v;
  }
}

}  // namespace

extern "C" cpd_status cpd_threshold(float* x, int n_images, int L, int alg, float threshold, int clamp_inplace, float* bound,
                                    void* stream) {
  CPD_REQUIRE(x && bound, "cpd_threshold: null pointer");
  CPD_REQUIRE(n_images >= 0 && L > 0 && L % 4 == 0, "cpd_threshold: n_images=%d L=%d (L must be a positive multiple of 4)", n_images, L);
  CPD_REQUIRE(alg == CPD_THRESH_DYNAMIC || alg == CPD_THRESH_STATIC, "cpd_threshold: unknown algorithm %d", alg);
  if (n_images == 0) return CPD_OK;
  cudaStream_t s = (cudaStream_t)stream;
  if (alg == CPD_THRESH_DYNAMIC) {
    CPD_REQUIRE(threshold >= 0.f && threshold <= 100.f, "cpd_threshold: percentile %f outside [0, 100]", (double)threshold);
    // numpy >= 2 on a float32 array: q = fl32(fl32(threshold) / 100f), virtual index = fl32((n - 1) * q), both in fp32
    // (volatile: no double-precision or fused evaluation by the host compiler)
    volatile float q32 = threshold / 100.0f;
    volatile float vi = (float)(L - 1) * q32;
    int k_lo = (int)floorf(vi);
    int k_hi = k_lo + 1;
    if (vi >= (float)(L - 1)) k_lo = k_hi = L - 1;  // _get_indexes: at or above the last index -> the maximum
    const float t = vi - (float)(int)floorf(vi);
    CPD_CUDA_CHECK(cpd_launch(abs_percentile_kernel, dim3(n_images), dim3(1024), 0, s, (const float*)x, L, k_lo, k_hi, t, bound));
  } else {
    CPD_CUDA_CHECK(cpd_launch(fill_bound_kernel, dim3((n_images + 127) / 128), dim3(128), 0, s, bound, n_images, threshold));
  }
  CPD_CUDA_CHECK(cudaGetLastError());
  if (clamp_inplace) {
    int bx = (L / 4 + 255) / 256;
    if (bx > 64) bx = 64;
    CPD_CUDA_CHECK(cpd_launch(clamp_half_kernel, dim3(bx, n_images), dim3(256), 0, s, x, L / 4, (const float*)bound));
    CPD_CUDA_CHECK(cudaGetLastError());
  }
  return CPD_OK;
}
