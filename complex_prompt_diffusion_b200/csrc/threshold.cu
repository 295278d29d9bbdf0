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
// ABS = false: the percentile of the SIGNED values (attention-guidance saliency mask, denoiser.py:411): keys are the
// order-preserving unsigned image of the float bits (negative: all bits flipped, non-negative: sign bit set); the result is
// the interpolated percentile itself (no max with 1).
template <bool ABS>
__device__ __forceinline__ unsigned pct_key(float v) {
  const unsigned u = __float_as_uint(v);
  if (ABS) return u & 0x7fffffffu;
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
template <bool ABS>
__device__ __forceinline__ float pct_value(unsigned key) {
  if (ABS) return __uint_as_float(key);
  return __uint_as_float((key & 0x80000000u) ? (key & 0x7fffffffu) : ~key);
}

template <bool ABS>
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
        const unsigned key = pct_key<ABS>(xi[i]);
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
    stat[which] = pct_value<ABS>(s_prefix);
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float a = stat[0], b = stat[1];
    const float diff = __fsub_rn(b, a);
    float r = __fadd_rn(a, __fmul_rn(diff, t));                                  // numpy _lerp
    if (t >= 0.5f) r = __fsub_rn(b, __fmul_rn(diff, __fsub_rn(1.0f, t)));
    bound[blockIdx.x] = ABS ? fmaxf(r, 1.0f) : r;                                 // np.max(np.append(s, 1.0)) for the clamp bound
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

// ---------------------------------------------------------------------------------------------------------------------
// The remaining registered variants (threshold.py:87-286): min-max rescaling to [-1, 1], torch.quantile / np.percentile
// bounds, RMS ("norm") and per-pixel channel-RMS ("spatial norm") rescaling.  One CTA per image does all of it: the image
// (16 K - 64 K floats) stays in L2 between the passes.  Every elementwise operation is a separately rounded fp32 op in
// the order the eager reference executes it; the result is rounded through fp16 (the reference returns x.half(), D10).
struct ExParams {
  int C, hw;       // image = [C][hw] floats
  int alg;
  float p0;        // quantile variants: unused; norm variants: fl32(threshold / 100) (scaled) or fl32(threshold) (spatial_norm)
  int k_lo, k_hi;  // order statistics around the virtual index
  float w;         // interpolation weight
};

__device__ __forceinline__ float block_reduce_minmax(float v, bool is_max, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float u = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? fmaxf(v, u) : fminf(v, u);
  }
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    float r = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : red[0];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float u = __shfl_xor_sync(0xffffffffu, r, o);
      r = is_max ? fmaxf(r, u) : fminf(r, u);
    }
    if (threadIdx.x == 0) red[32] = r;
  }
  __syncthreads();
  const float out = red[32];
  __syncthreads();
  return out;
}

__global__ void __launch_bounds__(1024) threshold_ex_kernel(float* __restrict__ x, ExParams p, float* __restrict__ bound) {
  __shared__ unsigned hist[256];
  __shared__ unsigned s_prefix, s_rank;
  __shared__ float red[33];
  __shared__ double dred[32];
  pdl_launch_dependents();
  pdl_wait();
  const int L = p.C * p.hw;
  float* xi = x + (int64_t)blockIdx.x * L;
  const bool scaled = p.alg == CPD_THRESH_SCALED_DYNAMIC_PERC || p.alg == CPD_THRESH_RENORM || p.alg == CPD_THRESH_SCALED_NORM ||
                      p.alg == CPD_THRESH_SCALED_SPATIAL_NORM;
  float mn = 0.f, mx = 0.f, rng = 1.f;
  if (scaled) {  // x_max, x_min = x.max(), x.min() (threshold.py:131,160,220,269)
    float lo = INFINITY, hi = -INFINITY;
    for (int i = threadIdx.x; i < L; i += blockDim.x) {
      const float v = xi[i];
      lo = fminf(lo, v);
      hi = fmaxf(hi, v);
    }
    mn = block_reduce_minmax(lo, false, red);
    mx = block_reduce_minmax(hi, true, red);
    rng = __fsub_rn(mx, mn);
  }
  // y = 2 * ((x - x_min) / (x_max - x_min)) - 1 (:132-133), or x itself
  auto val = [&](int i) -> float {
    const float v = xi[i];
    return scaled ? __fsub_rn(__fmul_rn(2.f, __fdiv_rn(__fsub_rn(v, mn), rng)), 1.f) : v;
  };
  auto unscale = [&](float y) -> float {  // x = (y + 1) / 2; x = (x_max - x_min) * x + x_min (:144-145)
    return scaled ? __fadd_rn(__fmul_rn(rng, __fmul_rn(__fadd_rn(y, 1.f), 0.5f)), mn) : y;
  };
  const bool quant = p.alg == CPD_THRESH_DYNANORMIC || p.alg == CPD_THRESH_SCALED_DYNAMIC_PERC || p.alg == CPD_THRESH_RENORM;
  if (quant) {
    float stat[2];
    for (int which = 0; which < 2; ++which) {
      if (which == 1 && p.k_hi == p.k_lo) {
        stat[1] = stat[0];
        break;
      }
      if (threadIdx.x == 0) {
        s_prefix = 0u;
        s_rank = (unsigned)(which ? p.k_hi : p.k_lo);
      }
      unsigned mask = 0u;
      for (int pass = 3; pass >= 0; --pass) {
        for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0u;
        __syncthreads();
        const unsigned prefix = s_prefix;
        for (int i = threadIdx.x; i < L; i += blockDim.x) {
          const unsigned key = __float_as_uint(val(i)) & 0x7fffffffu;
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
    const float a = stat[0], b = stat[1], diff = __fsub_rn(b, a);
    float s;
    if (p.alg == CPD_THRESH_SCALED_DYNAMIC_PERC) {  // np.percentile: numpy's _lerp, separately rounded
      s = __fadd_rn(a, __fmul_rn(diff, p.w));
      if (p.w >= 0.5f) s = __fsub_rn(b, __fmul_rn(diff, __fsub_rn(1.0f, p.w)));
    } else {  // torch.quantile: values_below.lerp_(values_above, weights), fused multiply-add on both CPU paths
      s = p.w < 0.5f ? fmaf(p.w, diff, a) : fmaf(__fsub_rn(p.w, 1.0f), diff, b);
    }
    s = fmaxf(s, 1.0f);  // maximum(s, 1) / np.max(np.append(s, 1.0)) / clamp_(min=1.0)
    if (threadIdx.x == 0) bound[blockIdx.x] = s;
    for (int i = threadIdx.x; i < L; i += blockDim.x) {
      float y = fminf(fmaxf(val(i), -s), s);
      if (p.alg == CPD_THRESH_DYNANORMIC) y = __fdiv_rn(y, s);  // x = clamp(x, -s, s); x = x / s (:113-114)
      xi[i] = __half2float(__float2half_rn(unscale(y)));
    }
    return;
  }
  // norm variants: thr = fl32(threshold / 100) * x_max (:224-225, 274-275) or the raw threshold (spatial_norm, :250)
  const float thr = (p.alg == CPD_THRESH_SPATIAL_NORM) ? p.p0 : __fmul_rn(p.p0, mx);
  if (p.alg == CPD_THRESH_SCALED_NORM) {
    // s = sqrt(mean(y^2)) over the image: fp64 accumulation in a fixed order (the eager fp32 sum order of the reference
    // is not reproducible; this is the correctly rounded value it approximates)
    double acc = 0.0;
    for (int i = threadIdx.x; i < L; i += blockDim.x) {
      const float y = val(i);
      acc += (double)__fmul_rn(y, y);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) dred[threadIdx.x >> 5] = acc;
    __syncthreads();
    double tot = 0.0;
    for (int j = 0; j < (int)(blockDim.x >> 5); ++j) tot += dred[j];
    const float s = fmaxf(sqrtf((float)(tot / (double)L)), thr);
    const float ratio = __fdiv_rn(thr, s);
    if (threadIdx.x == 0) bound[blockIdx.x] = s;
    for (int i = threadIdx.x; i < L; i += blockDim.x)
      xi[i] = __half2float(__float2half_rn(unscale(__fmul_rn(val(i), ratio))));
    return;
  }
  // spatial variants: s[pixel] = max(sqrt(mean_c y^2), thr), channels summed in order (x.pow(2).mean(1, keepdim=True))
  float smax = 0.f;
  const float inv_c = 1.0f / (float)p.C;
  const bool pow2_c = (p.C & (p.C - 1)) == 0;
  for (int px = threadIdx.x; px < p.hw; px += blockDim.x) {
    float acc = 0.f;
    for (int c = 0; c < p.C; ++c) {
      const float y = val(c * p.hw + px);
      acc = c == 0 ? __fmul_rn(y, y) : __fadd_rn(acc, __fmul_rn(y, y));
    }
    const float mean = pow2_c ? __fmul_rn(acc, inv_c) : __fdiv_rn(acc, (float)p.C);
    const float s = fmaxf(sqrtf(mean), thr);
    // scaled variant: thr is a 0-dim tensor, thr / s is a true division; spatial_norm: a Python float divided by a tensor
    // is Tensor.__rtruediv__ = s.reciprocal() * threshold, two roundings
    const float ratio = (p.alg == CPD_THRESH_SPATIAL_NORM) ? __fmul_rn(__frcp_rn(s), thr) : __fdiv_rn(thr, s);
    smax = fmaxf(smax, s);
    for (int c = 0; c < p.C; ++c) {
      const int i = c * p.hw + px;
      xi[i] = __half2float(__float2half_rn(unscale(__fmul_rn(val(i), ratio))));
    }
  }
  smax = block_reduce_minmax(smax, true, red);
  if (threadIdx.x == 0) bound[blockIdx.x] = smax;
}

}  // namespace

extern "C" cpd_status cpd_threshold_ex(float* x, int n_images, int channels, int hw, int alg, double threshold, float* bound,
                                       void* stream) {
  CPD_REQUIRE(x && bound, "cpd_threshold_ex: null pointer");
  CPD_REQUIRE(n_images >= 0 && channels > 0 && hw > 0, "cpd_threshold_ex: n_images=%d channels=%d hw=%d", n_images, channels, hw);
  CPD_REQUIRE(alg >= CPD_THRESH_DYNANORMIC && alg <= CPD_THRESH_SCALED_SPATIAL_NORM, "cpd_threshold_ex: unknown algorithm %d", alg);
  if (n_images == 0) return CPD_OK;
  const int L = channels * hw;
  ExParams p;
  p.C = channels;
  p.hw = hw;
  p.alg = alg;
  p.p0 = 0.f;
  p.k_lo = p.k_hi = 0;
  p.w = 0.f;
  if (alg == CPD_THRESH_DYNANORMIC || alg == CPD_THRESH_RENORM) {
    // torch.quantile(|x|, q): q = threshold / 100 when 1 < threshold <= 100 (a Python double), cast to the input dtype;
    // ranks = q32 * (n - 1) in fp32; below = trunc, above = ceil, weight = ranks - below (ATen quantile_compute)
    double q = threshold;
    if (q > 1.0 && q <= 100.0) q = q / 100.0;
    CPD_REQUIRE(q >= 0.0 && q <= 1.0, "cpd_threshold_ex: quantile %f outside [0, 1]", q);
    volatile float q32 = (float)q;
    volatile float rank = q32 * (float)(L - 1);
    p.k_lo = (int)floorf(rank);
    p.k_hi = (int)ceilf(rank);
    if (p.k_hi > L - 1) p.k_hi = L - 1;
    p.w = rank - (float)p.k_lo;
  } else if (alg == CPD_THRESH_SCALED_DYNAMIC_PERC) {
    CPD_REQUIRE(threshold >= 0.0 && threshold <= 100.0, "cpd_threshold_ex: percentile %f outside [0, 100]", threshold);
    volatile float q32 = (float)threshold / 100.0f;  // numpy >= 2 on float32 data, as in cpd_threshold
    volatile float vi = (float)(L - 1) * q32;
    p.k_lo = (int)floorf(vi);
    p.k_hi = p.k_lo + 1;
    if (vi >= (float)(L - 1)) p.k_lo = p.k_hi = L - 1;
    p.w = vi - (float)(int)floorf(vi);
  } else if (alg == CPD_THRESH_SPATIAL_NORM) {
    p.p0 = (float)threshold;
  } else {
    p.p0 = (float)(threshold / 100.0);
  }
  CPD_CUDA_CHECK(cpd_launch(threshold_ex_kernel, dim3(n_images), dim3(1024), 0, (cudaStream_t)stream, x, p, bound));
  CPD_CUDA_CHECK(cudaGetLastError());
  return CPD_OK;
}

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
    CPD_CUDA_CHECK(cpd_launch(abs_percentile_kernel<true>, dim3(n_images), dim3(1024), 0, s, (const float*)x, L, k_lo, k_hi, t, bound));
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

// np.percentile(x, q) of n signed fp32 values (one CTA; linear interpolation, virtual index in fp32 like numpy >= 2 on a float32
// array): out[0].  The saliency-mask threshold of the attention guidance (denoiser.py:411).
extern "C" cpd_status cpd_percentile(const float* x, int n, float q, float* out, void* stream) {
  CPD_REQUIRE(x && out, "cpd_percentile: null pointer");
  CPD_REQUIRE(n > 0 && q >= 0.f && q <= 100.f, "cpd_percentile: n=%d q=%f", n, (double)q);
  volatile float q32 = q / 100.0f;
  volatile float vi = (float)(n - 1) * q32;
  int k_lo = (int)floorf(vi);
  int k_hi = k_lo + 1;
  if (vi >= (float)(n - 1)) k_lo = k_hi = n - 1;
  const float t = vi - (float)(int)floorf(vi);
  CPD_CUDA_CHECK(cpd_launch(abs_percentile_kernel<false>, dim3(1), dim3(1024), 0, (cudaStream_t)stream, x, n, k_lo, k_hi, t, out));
  CPD_CUDA_CHECK(cudaGetLastError());
  return CPD_OK;
}
