// Fused multi-prompt CFG combine + eps/v -> denoised + Euler / Euler-ancestral / DPM++ 2M update.
// One HBM-streaming pass over the latents per sampler step (SURVEY.md section 8, rows A4(post-UNet)-A10).
//
// Arithmetic contract (bit-exact with the torch ops of the reference for fp32 eps):
//   * weighted delta in fp16 with every intermediate rounded (denoiser.py:450-460):
//       sum = sum_k (half(m_k) * half(w_k)) * (half(e_k) - half(e_u)), left to right
//   * scaled = half(float(sum) * s)                               (denoiser.py:514)
//   * e_t = e_u + scaled  (fp16+fp16 -> fp16, otherwise fp32)      (denoiser.py:515)
//   * denoised / update in fp32 with separately rounded mul/add/div - no FMA contraction - in the
//     reference's operation order (denoiser.py:540-542, euler.py:49-54,85-92, dpmpp.py:42-54).
#include "../../include/cpd_b200.h"
#include "common.cuh"

namespace {

struct StepArgs {
  cpd_step_params p;
};

__device__ __forceinline__ float h_round(float v) { return __half2float(__float2half_rn(v)); }

template <int DT>
__device__ __forceinline__ void load4(const void* base, int64_t idx, float (&out)[4]) {
  if (DT == CPD_F32) {
    float4 v = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + idx));
    out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
  } else if (DT == CPD_F16) {
    uint2 v = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __half*>(base) + idx));
    float2 a = __half22float2(*reinterpret_cast<__half2*>(&v.x));
    float2 b = __half22float2(*reinterpret_cast<__half2*>(&v.y));
    out[0] = a.x; out[1] = a.y; out[2] = b.x; out[3] = b.y;
  } else {
    uint2 v = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const bf16*>(base) + idx));
    float2 a = unpack_bf16x2(v.x);
    float2 b = unpack_bf16x2(v.y);
    out[0] = a.x; out[1] = a.y; out[2] = b.x; out[3] = b.y;
  }
}

template <int DT>
__global__ void __launch_bounds__(256, 4) sampler_step_kernel(const __grid_constant__ StepArgs args) {
  pdl_launch_dependents();
  pdl_wait();
  const cpd_step_params& p = args.p;
  const int L = 4 * p.hw;
  const int vec_per_img = L / 4;
  const int64_t total = (int64_t)p.n_images * vec_per_img;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < total; v += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(v / vec_per_img);
    const int i = (int)(v - (int64_t)b * vec_per_img) * 4;  // element offset inside the image
    const int pix = i % p.hw;
    const int64_t ebase = (int64_t)b * p.eps_image_stride + i;
    // Issue every independent load of this vector before the first use (memory-level parallelism: the kernel is a pure
    // HBM stream): x, the 2M history / ancestral noise, the unconditional row and the sub-prompt rows four at a time.
    const float4 xv = *reinterpret_cast<const float4*>(p.x + (int64_t)b * L + i);
    float4 aux = make_float4(0.f, 0.f, 0.f, 0.f);  // old_denoised (2M) or noise (ancestral)
    if (p.sampler == CPD_DPMPP_2M) {
      if (!p.dpm_first) aux = *reinterpret_cast<const float4*>(p.old_denoised + (int64_t)b * L + i);
    } else if (p.sampler == CPD_EULER_ANCESTRAL) {
      aux = __ldg(reinterpret_cast<const float4*>(p.noise + (int64_t)b * L + i));
    }
    // The fp16 delta (denoiser.py:450-460) runs on packed half2: for fp16 operands HSUB2 / HMUL2 / HADD2 (one rounding) give
    // bit-identical results to "compute in fp32, round to fp16" (products of two halves are exact in fp32; sums are exact
    // unless the smaller operand is below a quarter ulp of the larger, where both roundings return the larger).  The _rn
    // intrinsics keep ptxas from contracting mul + add into a single-rounding HFMA2.
    float eu[4];
    load4<DT>(p.eps, ebase, eu);
    const __half2 hu01 = __floats2half2_rn(eu[0], eu[1]), hu23 = __floats2half2_rn(eu[2], eu[3]);
    __half2 sum01 = __floats2half2_rn(0.f, 0.f), sum23 = sum01;
    for (int k0 = 0; k0 < p.n_sub; k0 += 4) {
      float ek4[4][4];
      float4 m4[4];
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const int k = k0 + kk;
        if (k < p.n_sub) {
          load4<DT>(p.eps, ebase + (int64_t)(k + 1) * p.eps_row_stride, ek4[kk]);
          if (p.masks[k] != nullptr) m4[kk] = __ldg(reinterpret_cast<const float4*>(p.masks[k] + pix));
        }
      }
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const int k = k0 + kk;
        if (k < p.n_sub) {
          // half(m) * half(w), rounded to fp16 (denoiser.py:451-452); weights live in the kernel-parameter constant bank
          const __half w_hk = __float2half_rn(p.weights[k]);
          __half2 mw01, mw23;
          if (p.masks[k] != nullptr) {
            const __half2 wk2 = __half2half2(w_hk);
            mw01 = __hmul2_rn(__floats2half2_rn(m4[kk].x, m4[kk].y), wk2);
            mw23 = __hmul2_rn(__floats2half2_rn(m4[kk].z, m4[kk].w), wk2);
          } else {
            mw01 = mw23 = __half2half2(__hmul_rn(__float2half_rn(p.mask_scalar[k]), w_hk));
          }
          const __half2 t01 = __hmul2_rn(mw01, __hsub2_rn(__floats2half2_rn(ek4[kk][0], ek4[kk][1]), hu01));
          const __half2 t23 = __hmul2_rn(mw23, __hsub2_rn(__floats2half2_rn(ek4[kk][2], ek4[kk][3]), hu23));
          if (k == 0) {
            sum01 = t01;
            sum23 = t23;
          } else {
            sum01 = __hadd2_rn(sum01, t01);
            sum23 = __hadd2_rn(sum23, t23);
          }
        }
      }
    }
    const float2 s01 = __half22float2(sum01), s23 = __half22float2(sum23);
    const float sum[4] = {s01.x, s01.y, s23.x, s23.y};
    float x[4] = {xv.x, xv.y, xv.z, xv.w};
    float et[4], den[4], xn[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float scaled = h_round(__fmul_rn(sum[j], p.guidance));
      if (DT == CPD_F16) et[j] = h_round(__fadd_rn(eu[j], scaled));
      else et[j] = __fadd_rn(eu[j], scaled);
      if (p.pred_type == CPD_PRED_EPSILON) den[j] = __fsub_rn(x[j], __fmul_rn(p.sigma_hat, et[j]));
      else den[j] = __fadd_rn(__fmul_rn(et[j], p.v_c_eps), __fdiv_rn(x[j], p.v_c_x_div));
    }
    if (p.sampler == CPD_DPMPP_2M) {
      const float od[4] = {aux.x, aux.y, aux.z, aux.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float dd = den[j];
        if (!p.dpm_first) dd = __fsub_rn(__fmul_rn(p.dpm_c1, den[j]), __fmul_rn(p.dpm_c2, od[j]));
        xn[j] = __fsub_rn(__fmul_rn(p.dpm_ratio, x[j]), __fmul_rn(p.dpm_expm1, dd));
      }
      if (p.write_old)
        *reinterpret_cast<float4*>(p.old_denoised + (int64_t)b * L + i) = make_float4(den[0], den[1], den[2], den[3]);
    } else {
      const float nz[4] = {aux.x, aux.y, aux.z, aux.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float d = __fdiv_rn(__fsub_rn(x[j], den[j]), p.sigma_hat);
        xn[j] = __fadd_rn(x[j], __fmul_rn(d, p.dt));
        if (p.sampler == CPD_EULER_ANCESTRAL) xn[j] = __fadd_rn(xn[j], __fmul_rn(nz[j], p.sigma_up));
      }
    }
    if (p.sampler != CPD_DENOISE_ONLY)
      *reinterpret_cast<float4*>(p.x + (int64_t)b * L + i) = make_float4(xn[0], xn[1], xn[2], xn[3]);
    if (p.denoised_out)
      *reinterpret_cast<float4*>(p.denoised_out + (int64_t)b * L + i) = make_float4(den[0], den[1], den[2], den[3]);
    if (p.eps_out)
      *reinterpret_cast<float4*>(p.eps_out + (int64_t)b * L + i) = make_float4(et[0], et[1], et[2], et[3]);
  }
}

}  // namespace

extern "C" cpd_status cpd_sampler_step(const cpd_step_params* p, void* stream) {
  CPD_REQUIRE(p != nullptr, "cpd_sampler_step: null params");
  CPD_REQUIRE(p->eps && p->x, "cpd_sampler_step: eps and x must be non-null");
  CPD_REQUIRE(p->n_sub >= 1 && p->n_sub <= CPD_MAX_SUBPROMPTS, "cpd_sampler_step: n_sub=%d out of range [1,%d]", p->n_sub,
              CPD_MAX_SUBPROMPTS);
  CPD_REQUIRE(p->hw > 0 && p->hw % 4 == 0, "cpd_sampler_step: hw=%d must be a positive multiple of 4", p->hw);
  CPD_REQUIRE(p->n_images >= 0, "cpd_sampler_step: n_images=%d", p->n_images);
  CPD_REQUIRE(p->eps_row_stride % 4 == 0 && p->eps_image_stride % 4 == 0, "cpd_sampler_step: eps strides must be multiples of 4");
  CPD_REQUIRE(p->sampler >= CPD_EULER && p->sampler <= CPD_DENOISE_ONLY, "cpd_sampler_step: unknown sampler %d", p->sampler);
  CPD_REQUIRE(p->pred_type == CPD_PRED_EPSILON || p->pred_type == CPD_PRED_VELOCITY, "cpd_sampler_step: unknown pred_type %d",
              p->pred_type);
  CPD_REQUIRE(p->sampler != CPD_EULER_ANCESTRAL || p->noise, "cpd_sampler_step: ancestral sampler needs noise");
  CPD_REQUIRE(p->sampler != CPD_DPMPP_2M || p->old_denoised || (p->dpm_first && !p->write_old),
              "cpd_sampler_step: DPM++ 2M needs old_denoised");
  if (p->n_images == 0) return CPD_OK;  // empty batch: nothing to do
  StepArgs args;
  args.p = *p;
  const int64_t total = (int64_t)p->n_images * p->hw;
  int blocks = (int)((total + 255) / 256);  // total = number of 4-element vectors
  const int max_blocks = 148 * 8;
  if (blocks > max_blocks) blocks = max_blocks;
  cudaStream_t s = (cudaStream_t)stream;
  switch (p->eps_dtype) {
    case CPD_F32: CPD_CUDA_CHECK(cpd_launch(sampler_step_kernel<CPD_F32>, dim3(blocks), dim3(256), 0, s, args)); break;
    case CPD_F16: CPD_CUDA_CHECK(cpd_launch(sampler_step_kernel<CPD_F16>, dim3(blocks), dim3(256), 0, s, args)); break;
    case CPD_BF16: CPD_CUDA_CHECK(cpd_launch(sampler_step_kernel<CPD_BF16>, dim3(blocks), dim3(256), 0, s, args)); break;
    default: cpd_set_error("cpd_sampler_step: unknown eps_dtype %d", p->eps_dtype); return CPD_ERR_INVALID;
  }
  CPD_CUDA_CHECK(cudaGetLastError());
  return CPD_OK;
}
