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

// VEC consecutive elements of one eps row, kept as raw 32-bit words (fp32: one element per word, 16-bit: two per word) so a
// row costs VEC or VEC/2 registers until it is used: one 16-byte load for fp32 x 4 and for 16-bit x 8.
template <int DT, int VEC>
struct EpsRow {
  static constexpr int W = (DT == CPD_F32) ? VEC : VEC / 2;
  uint32_t w[W];
  __device__ __forceinline__ void load(const void* base, int64_t idx) {
    if (DT == CPD_F32) {
#pragma unroll
      for (int q = 0; q < VEC / 4; ++q) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(base) + idx) + q);
        w[4 * q] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
      }
    } else if (VEC == 8) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(base) + idx));
      w[0] = v.x; w[1] = v.y; w[W - 2] = v.z; w[W - 1] = v.w;
    } else {
      const uint2 v = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const uint16_t*>(base) + idx));
      w[0] = v.x; w[1] = v.y;
    }
  }
  __device__ __forceinline__ __half2 half2_at(int q) const {  // elements 2q, 2q+1 rounded to fp16
    if (DT == CPD_F32) return __floats2half2_rn(__uint_as_float(w[2 * q]), __uint_as_float(w[2 * q + 1]));
    if (DT == CPD_F16) { uint32_t u = w[q]; return *reinterpret_cast<__half2*>(&u); }
    const float2 f = unpack_bf16x2(w[q]);
    return __floats2half2_rn(f.x, f.y);
  }
  __device__ __forceinline__ float at(int j) const {
    if (DT == CPD_F32) return __uint_as_float(w[j]);
    if (DT == CPD_F16) { uint32_t u = w[j >> 1]; const __half2 h = *reinterpret_cast<__half2*>(&u); return (j & 1) ? __high2float(h) : __low2float(h); }
    const float2 f = unpack_bf16x2(w[j >> 1]);
    return (j & 1) ? f.y : f.x;
  }
};

template <int VEC>
__device__ __forceinline__ void load_f32(const float* p, float (&out)[VEC]) {
#pragma unroll
  for (int q = 0; q < (VEC >= 4 ? VEC / 4 : 0); ++q) {
    const float4 v = *(reinterpret_cast<const float4*>(p) + q);
    out[4 * q] = v.x; out[4 * q + 1] = v.y; out[4 * q + 2] = v.z; out[4 * q + 3] = v.w;
  }
}
template <int VEC>
__device__ __forceinline__ void store_f32(float* p, const float (&v)[VEC]) {
#pragma unroll
  for (int q = 0; q < VEC / 4; ++q) *(reinterpret_cast<float4*>(p) + q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
}

// One thread owns VEC consecutive latent elements of one image (VEC = 4 for fp32 eps, 8 for 16-bit eps when hw % 8 == 0, so
// every eps row is fetched with 16-byte loads).
// FULL = false is the hot instantiation (Euler / Euler ancestral / DPM++ 2M / denoise-only on the UNet input, in place): the
// two-stage / multistep / thresholding operands (x_base, x_out, d_out, d_prev, clip_scaled, scaled_out, scaled_in) are
// compiled out, which is worth ~5 % of the HBM rate of this pure streaming kernel.
template <int DT, int VEC, bool FULL>
__global__ void __launch_bounds__(256, VEC == 8 ? 3 : 4) sampler_step_kernel(const __grid_constant__ StepArgs args) {
  pdl_launch_dependents();
  pdl_wait();
  const cpd_step_params& p = args.p;
  // per-step scalars: by value, or the "current" row of a device table (a captured graph replayed once per step)
  cpd_step_scalars sc;
  if (p.dyn) {
    sc = *p.dyn;  // uniform 64-byte load
  } else {
    sc.guidance = p.guidance; sc.sigma_hat = p.sigma_hat; sc.v_c_eps = p.v_c_eps; sc.v_c_x_div = p.v_c_x_div; sc.dt = p.dt;
    sc.sigma_up = p.sigma_up; sc.dpm_ratio = p.dpm_ratio; sc.dpm_expm1 = p.dpm_expm1; sc.dpm_c1 = p.dpm_c1; sc.dpm_c2 = p.dpm_c2;
    sc.dpm_first = p.dpm_first; sc.write_old = p.write_old; sc.noise_mul = p.noise_mul;
  }
  const int L = 4 * p.hw;
  const int vec_per_img = L / VEC;
  const int64_t total = (int64_t)p.n_images * vec_per_img;
  constexpr int H2 = VEC / 2;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < total; v += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(v / vec_per_img);
    const int i = (int)(v - (int64_t)b * vec_per_img) * VEC;  // element offset inside the image
    const int pix = i % p.hw;
    const int64_t ebase = (int64_t)b * p.eps_image_stride + i;
    // Issue every independent load of this vector before the first use (memory-level parallelism: the kernel is a pure
    // HBM stream): x, the 2M history / ancestral noise, the unconditional row and the sub-prompt rows four at a time.
    float x[VEC], aux[VEC];  // x: the UNet input; aux: old_denoised (2M), noise (ancestral / 2S-a) or d of stage 1 (Heun)
    load_f32<VEC>(p.x + (int64_t)b * L + i, x);
#pragma unroll
    for (int j = 0; j < VEC; ++j) aux[j] = 0.f;
    if (p.sampler == CPD_DPMPP_2M) {
      if (!sc.dpm_first) load_f32<VEC>(p.old_denoised + (int64_t)b * L + i, aux);
      else if (FULL && p.noise) load_f32<VEC>(p.noise + (int64_t)b * L + i, aux);  // DPM++ 2S ancestral: noise after the update
    } else if (p.sampler == CPD_EULER_ANCESTRAL) {
      load_f32<VEC>(p.noise + (int64_t)b * L + i, aux);
    } else if (FULL && p.sampler == CPD_HEUN2) {
      load_f32<VEC>(p.d_prev[0] + (int64_t)b * L + i, aux);
    }
    // The fp16 delta (denoiser.py:450-460) runs on packed half2: for fp16 operands HSUB2 / HMUL2 / HADD2 (one rounding) give
    // bit-identical results to "compute in fp32, round to fp16" (products of two halves are exact in fp32; sums are exact
    // unless the smaller operand is below a quarter ulp of the larger, where both roundings return the larger).  The _rn
    // intrinsics keep ptxas from contracting mul + add into a single-rounding HFMA2.
    EpsRow<DT, VEC> eu;
    eu.load(p.eps, ebase);
    __half2 sum[H2];
#pragma unroll
    for (int q = 0; q < H2; ++q) sum[q] = __floats2half2_rn(0.f, 0.f);
    for (int k0 = 0; k0 < p.n_sub; k0 += 4) {
      EpsRow<DT, VEC> ek4[4];
      float m4[4][VEC == 4 ? 4 : 1];  // spatial masks: prefetched with the rows (VEC 4) or read at use (VEC 8: registers)
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const int k = k0 + kk;
        if (k < p.n_sub) {
          ek4[kk].load(p.eps, ebase + (int64_t)(k + 1) * p.eps_row_stride);
          if (VEC == 4 && p.masks[k] != nullptr) load_f32<VEC == 4 ? 4 : 1>(p.masks[k] + pix, m4[kk]);
        }
      }
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const int k = k0 + kk;
        if (k < p.n_sub) {
          // half(m) * half(w), rounded to fp16 (denoiser.py:451-452); weights live in the kernel-parameter constant bank
          const __half w_hk = __float2half_rn(p.weights[k]);
          const __half2 wk2 = __half2half2(w_hk);
          const __half2 mw_s = __half2half2(__hmul_rn(__float2half_rn(p.mask_scalar[k]), w_hk));
#pragma unroll
          for (int q = 0; q < H2; ++q) {
            __half2 mw = mw_s;
            if (p.masks[k] != nullptr) {
              const float2 mm = (VEC == 4) ? make_float2(m4[kk][(2 * q) % 4], m4[kk][(2 * q + 1) % 4])
                                           : __ldg(reinterpret_cast<const float2*>(p.masks[k] + pix + 2 * q));
              mw = __hmul2_rn(__floats2half2_rn(mm.x, mm.y), wk2);
            }
            const __half2 t = __hmul2_rn(mw, __hsub2_rn(ek4[kk].half2_at(q), eu.half2_at(q)));
            sum[q] = (k == 0) ? t : __hadd2_rn(sum[q], t);
          }
        }
      }
    }
    float et[VEC], den[VEC], xn[VEC];
    const float clip = (FULL && p.clip_scaled) ? __ldg(p.clip_scaled + b) : 0.f;
    const float nmul = sc.noise_mul;  // used as given: 0 adds no noise (temperature = 0, dpmpp.py:111)
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const float sj = (j & 1) ? __high2float(sum[j >> 1]) : __low2float(sum[j >> 1]);
      float scaled = h_round(__fmul_rn(sj, sc.guidance));
      if (FULL && p.scaled_out) p.scaled_out[(int64_t)b * L + i + j] = scaled;
      if (FULL && p.clip_scaled) scaled = h_round(fminf(fmaxf(scaled, -clip), clip));  // x.float() -> clamp_ -> .half()
      if (FULL && p.scaled_in) scaled = __ldg(p.scaled_in + (int64_t)b * L + i + j);  // thresholded by cpd_threshold_ex
      if (DT == CPD_F16) et[j] = h_round(__fadd_rn(eu.at(j), scaled));
      else et[j] = __fadd_rn(eu.at(j), scaled);
      if (p.pred_type == CPD_PRED_EPSILON) den[j] = __fsub_rn(x[j], __fmul_rn(sc.sigma_hat, et[j]));
      else den[j] = __fadd_rn(__fmul_rn(et[j], sc.v_c_eps), __fdiv_rn(x[j], sc.v_c_x_div));
    }
    // xb: the sample the update starts from (the UNet input unless this is the second stage of a two-stage sampler)
    float xb[VEC];
    if (FULL && p.x_base) {
      load_f32<VEC>(p.x_base + (int64_t)b * L + i, xb);
    } else {
#pragma unroll
      for (int j = 0; j < VEC; ++j) xb[j] = x[j];
    }
    if (p.sampler == CPD_DPMPP_2M) {
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        float dd = den[j];
        if (!sc.dpm_first) dd = __fsub_rn(__fmul_rn(sc.dpm_c1, den[j]), __fmul_rn(sc.dpm_c2, aux[j]));
        xn[j] = __fsub_rn(__fmul_rn(sc.dpm_ratio, xb[j]), __fmul_rn(sc.dpm_expm1, dd));
        if (FULL && sc.dpm_first && p.noise) xn[j] = __fadd_rn(xn[j], __fmul_rn(__fmul_rn(aux[j], nmul), sc.sigma_up));  // dpmpp.py:111
      }
      if (sc.write_old) store_f32<VEC>(p.old_denoised + (int64_t)b * L + i, den);
    } else {
      float d[VEC];
#pragma unroll
      for (int j = 0; j < VEC; ++j) d[j] = __fdiv_rn(__fsub_rn(x[j], den[j]), sc.sigma_hat);  // to_ode
      if (FULL && p.d_out) store_f32<VEC>(p.d_out + (int64_t)b * L + i, d);
      if (FULL && p.sampler == CPD_HEUN2) {
#pragma unroll
        for (int j = 0; j < VEC; ++j) xn[j] = __fadd_rn(xb[j], __fmul_rn(__fdiv_rn(__fadd_rn(aux[j], d[j]), 2.0f), sc.dt));
      } else if (FULL && p.sampler == CPD_LMS) {
        float acc[VEC];  // sum(coeff * d ...) of lms.py:52: 0 + c0 * d_i, then + c1 * d_{i-1}, ... left to right
#pragma unroll
        for (int j = 0; j < VEC; ++j) acc[j] = __fmul_rn(p.lms_coeff[0], d[j]);
        for (int k = 1; k < p.lms_order; ++k) {
          float dk[VEC];
          load_f32<VEC>(p.d_prev[k - 1] + (int64_t)b * L + i, dk);
#pragma unroll
          for (int j = 0; j < VEC; ++j) acc[j] = __fadd_rn(acc[j], __fmul_rn(p.lms_coeff[k], dk[j]));
        }
#pragma unroll
        for (int j = 0; j < VEC; ++j) xn[j] = __fadd_rn(xb[j], acc[j]);
      } else {
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          xn[j] = __fadd_rn(xb[j], __fmul_rn(d[j], sc.dt));
          if (p.sampler == CPD_EULER_ANCESTRAL) xn[j] = __fadd_rn(xn[j], __fmul_rn(__fmul_rn(aux[j], nmul), sc.sigma_up));
        }
      }
    }
    if (p.sampler != CPD_DENOISE_ONLY) store_f32<VEC>(((FULL && p.x_out) ? p.x_out : p.x) + (int64_t)b * L + i, xn);
    if (p.denoised_out) store_f32<VEC>(p.denoised_out + (int64_t)b * L + i, den);
    if (p.eps_out) store_f32<VEC>(p.eps_out + (int64_t)b * L + i, et);
  }
}

__global__ void __launch_bounds__(256) add_noise_kernel(float* __restrict__ x, const float* __restrict__ noise, float noise_mul, float scale,
                                                        int64_t n4) {
  pdl_launch_dependents();
  pdl_wait();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 xv = reinterpret_cast<float4*>(x)[i];
    const float4 nv = __ldg(reinterpret_cast<const float4*>(noise) + i);
    xv.x = __fadd_rn(xv.x, __fmul_rn(__fmul_rn(nv.x, noise_mul), scale));
    xv.y = __fadd_rn(xv.y, __fmul_rn(__fmul_rn(nv.y, noise_mul), scale));
    xv.z = __fadd_rn(xv.z, __fmul_rn(__fmul_rn(nv.z, noise_mul), scale));
    xv.w = __fadd_rn(xv.w, __fmul_rn(__fmul_rn(nv.w, noise_mul), scale));
    reinterpret_cast<float4*>(x)[i] = xv;
  }
}

}  // namespace

extern "C" cpd_status cpd_add_noise(float* x, const float* noise, float noise_mul, float scale, int64_t n, void* stream) {
  CPD_REQUIRE(x && noise, "cpd_add_noise: null pointer");
  CPD_REQUIRE(n >= 0 && n % 4 == 0, "cpd_add_noise: n=%lld must be a non-negative multiple of 4", (long long)n);
  if (n == 0) return CPD_OK;
  int64_t blocks = (n / 4 + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  CPD_CUDA_CHECK(cpd_launch(add_noise_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, x, noise, noise_mul, scale, n / 4));
  CPD_CUDA_CHECK(cudaGetLastError());
  return CPD_OK;
}

__global__ void step_select_kernel(const cpd_step_scalars* __restrict__ table, int n_steps, int* counter, cpd_step_scalars* current) {
  pdl_launch_dependents();
  pdl_wait();
  if (threadIdx.x == 0) {
    const int i = *counter;
    *current = table[i < n_steps ? i : n_steps - 1];
    *counter = i + 1;
  }
}

extern "C" cpd_status cpd_step_select(const cpd_step_scalars* table, int n_steps, int* counter, cpd_step_scalars* current, void* stream) {
  CPD_REQUIRE(table && counter && current && n_steps > 0, "cpd_step_select: null pointer or empty table");
  CPD_CUDA_CHECK(cpd_launch(step_select_kernel, dim3(1), dim3(32), 0, (cudaStream_t)stream, table, n_steps, counter, current));
  return CPD_OK;
}

extern "C" cpd_status cpd_sampler_step(const cpd_step_params* p, void* stream) {
  CPD_REQUIRE(p != nullptr, "cpd_sampler_step: null params");
  CPD_REQUIRE(p->eps && p->x, "cpd_sampler_step: eps and x must be non-null");
  // n_sub = 0: eps holds one row per image that already IS e_t (after a score corrector rewrote it, denoiser.py:517-518)
  CPD_REQUIRE(p->n_sub >= 0 && p->n_sub <= CPD_MAX_SUBPROMPTS, "cpd_sampler_step: n_sub=%d out of range [0,%d]", p->n_sub,
              CPD_MAX_SUBPROMPTS);
  CPD_REQUIRE(p->hw > 0 && p->hw % 4 == 0, "cpd_sampler_step: hw=%d must be a positive multiple of 4", p->hw);
  CPD_REQUIRE(p->n_images >= 0, "cpd_sampler_step: n_images=%d", p->n_images);
  CPD_REQUIRE(p->eps_row_stride % 4 == 0 && p->eps_image_stride % 4 == 0, "cpd_sampler_step: eps strides must be multiples of 4");
  CPD_REQUIRE(p->sampler >= CPD_EULER && p->sampler <= CPD_LMS, "cpd_sampler_step: unknown sampler %d", p->sampler);
  CPD_REQUIRE(p->sampler != CPD_HEUN2 || p->d_prev[0], "cpd_sampler_step: the second Heun stage needs d_prev[0]");
  CPD_REQUIRE(p->sampler != CPD_LMS || (p->lms_order >= 1 && p->lms_order <= 4), "cpd_sampler_step: lms_order=%d", p->lms_order);
  if (p->sampler == CPD_LMS)
    for (int k = 1; k < p->lms_order; ++k) CPD_REQUIRE(p->d_prev[k - 1], "cpd_sampler_step: LMS order %d needs d_prev[%d]", p->lms_order, k - 1);
  CPD_REQUIRE(p->pred_type == CPD_PRED_EPSILON || p->pred_type == CPD_PRED_VELOCITY, "cpd_sampler_step: unknown pred_type %d",
              p->pred_type);
  CPD_REQUIRE(p->sampler != CPD_EULER_ANCESTRAL || p->noise, "cpd_sampler_step: ancestral sampler needs noise");
  CPD_REQUIRE(p->sampler != CPD_DPMPP_2M || p->old_denoised || (!p->dyn && p->dpm_first && !p->write_old),
              "cpd_sampler_step: DPM++ 2M needs old_denoised");
  if (p->n_images == 0) return CPD_OK;  // empty batch: nothing to do
  StepArgs args;
  args.p = *p;
  const bool wide = p->eps_dtype != CPD_F32 && p->hw % 8 == 0 && p->eps_row_stride % 8 == 0 && p->eps_image_stride % 8 == 0 &&
                    ((uintptr_t)p->eps & 15) == 0;
  const int64_t total = (int64_t)p->n_images * p->hw * 4 / (wide ? 8 : 4);  // number of per-thread vectors
  int blocks = (int)((total + 255) / 256);
  // grid-stride over whole waves of resident blocks (launch bounds: 3 per SM for the 8-wide kernel, 4 for the 4-wide one):
  // 148 * 8 blocks of the 8-wide kernel were 2.67 waves, i.e. a third wave that left a third of the SMs idle
  const int max_blocks = 148 * (wide ? 3 : 4) * 2;
  if (blocks > max_blocks) blocks = max_blocks;
  cudaStream_t s = (cudaStream_t)stream;
  const bool full = p->x_base || p->x_out || p->d_out || p->clip_scaled || p->scaled_out || p->scaled_in ||
                    p->sampler == CPD_HEUN2 || p->sampler == CPD_LMS || (p->sampler == CPD_DPMPP_2M && p->noise);
  const dim3 grid(blocks), block(256);
#define CPD_STEP_LAUNCH(DT)                                                                                        \
  do {                                                                                                             \
    if (wide && full) CPD_CUDA_CHECK(cpd_launch(sampler_step_kernel<DT, 8, true>, grid, block, 0, s, args));        \
    else if (wide) CPD_CUDA_CHECK(cpd_launch(sampler_step_kernel<DT, 8, false>, grid, block, 0, s, args));          \
    else if (full) CPD_CUDA_CHECK(cpd_launch(sampler_step_kernel<DT, 4, true>, grid, block, 0, s, args));           \
    else CPD_CUDA_CHECK(cpd_launch(sampler_step_kernel<DT, 4, false>, grid, block, 0, s, args));                    \
  } while (0)
  switch (p->eps_dtype) {
    case CPD_F32: CPD_STEP_LAUNCH(CPD_F32); break;  // wide is false for fp32 rows
    case CPD_F16: CPD_STEP_LAUNCH(CPD_F16); break;
    case CPD_BF16: CPD_STEP_LAUNCH(CPD_BF16); break;
    default: cpd_set_error("cpd_sampler_step: unknown eps_dtype %d", p->eps_dtype); return CPD_ERR_INVALID;
  }
#undef CPD_STEP_LAUNCH
  CPD_CUDA_CHECK(cudaGetLastError());
  return CPD_OK;
}
