// Device kernels of the Denoiser's optional guidance branches: the unconditional blur
// (cpd/samplers/extension/denoiser.py:333-337,441-442) and attention guidance (:341-350,404-435,461-462).  They are active
// on the last few steps of a schedule only; everything stays on the device (the reference round-trips through np.percentile
// on the CPU).  Every elementwise operation is a separately rounded fp32 op in the order the eager reference executes it.
#include "../../include/cpd_b200.h"
#include "common.cuh"

namespace {

constexpr int MAX_TAPS = 63;
struct BlurTaps {
  float k[MAX_TAPS];
};

__device__ __forceinline__ int reflect(int i, int n) {  // torch "reflect" padding (no edge repeat), |offset| < n
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}

// torchvision.transforms.functional.gaussian_blur: conv2d of the reflect-padded plane with kernel2d = mm(k1[:, None], k1[None, :])
// (each 2-D tap is the fp32-rounded product of two 1-D taps); fp32 accumulation, taps in row-major order.
__global__ void __launch_bounds__(256) gaussian_blur_kernel(const float* __restrict__ src, float* __restrict__ dst, int planes_per_img,
                                                            int64_t img_stride_src, int64_t img_stride_dst, int h, int w, int ksize,
                                                            const BlurTaps taps, int64_t total) {
  pdl_launch_dependents();
  pdl_wait();
  const int r = ksize >> 1;
  const int hw = h * w;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t plane = i / hw;
    const int p = (int)(i - plane * hw);
    const int y = p / w, x = p - y * w;
    const int64_t img = plane / planes_per_img, pc = plane - img * planes_per_img;
    const float* sp = src + img * img_stride_src + pc * hw;
    float acc = 0.f;
    for (int a = 0; a < ksize; ++a) {
      const float* row = sp + (int64_t)reflect(y + a - r, h) * w;
      const float ka = taps.k[a];
      for (int b = 0; b < ksize; ++b) acc = fmaf(__fmul_rn(ka, taps.k[b]), row[reflect(x + b - r, w)], acc);
    }
    dst[img * img_stride_dst + pc * hw + p] = acc;
  }
}

// mean over the channels of an NHWC activation tensor [rows][hw][C] -> fp32 [rows][hw]  (attn.mean(1, keepdims=True), :408)
template <bool F16>
__global__ void __launch_bounds__(256) channel_mean_kernel(const uint4* __restrict__ a, int64_t pixels, int c8, float inv_c, float* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t px = warp; px < pixels; px += nwarps) {  // one warp per pixel, lanes over 16-byte vectors, fixed butterfly
    float s = 0.f;
    for (int v = lane; v < c8; v += 32) {
      const uint4 q = a[px * c8 + v];
      const uint32_t u[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = unpack_act2(u[e], F16);
        s += f.x + f.y;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[px] = s * inv_c;
  }
}

__device__ __forceinline__ float eps_at(const void* eps, int dtype, int64_t i) {
  if (dtype == CPD_F32) return reinterpret_cast<const float*>(eps)[i];
  if (dtype == CPD_F16) return __half2float(reinterpret_cast<const __half*>(eps)[i]);
  return __bfloat162float(reinterpret_cast<const bf16*>(eps)[i]);
}

// stage 0: sample = x - sigma_hat * e_u (:419).  stage 1: the guided latent (:421-429).  stage 2: the mixed, scaled guidance
// term (:461-462, :514).  One launch handles n_images images of L = 4 * hw elements; e_u of image b starts at eps_u + b * eps_stride.
struct GuideArgs {
  int stage, n_images, hw, eps_dtype, mode;
  int64_t eps_stride;
  const float* x;
  const void* eps_u;
  const float* mask_mean;   // [n_images][rows_per_image][hw]: channel means of the saliency source; row 0 of an image is used
  int64_t mask_img_stride;  // rows_per_image * hw
  const float* pct;         // [n_images] percentile of each image's mask values
  const float* blur;        // stage 1: blurred sample [n_images][L]
  const float* sum16;       // stage 2: the fp16 weighted delta sum_e_t as fp32 values [n_images][L]
  const void* e_attn;       // stage 2: UNet output of the guided latent [n_images][L], eps_dtype
  float sigma_hat, c_in, scale, guidance;
  float* out;
};

__global__ void __launch_bounds__(256) attn_guide_kernel(const GuideArgs g) {
  pdl_launch_dependents();
  pdl_wait();
  const int L = 4 * g.hw;
  const int64_t total = (int64_t)g.n_images * L;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / L;
    const int j = (int)(i - b * L);
    if (g.stage == 0) {
      const float eu = eps_at(g.eps_u, g.eps_dtype, b * g.eps_stride + j);
      g.out[i] = __fsub_rn(g.x[i], __fmul_rn(g.sigma_hat, eu));
    } else if (g.stage == 1) {
      const float eu = eps_at(g.eps_u, g.eps_dtype, b * g.eps_stride + j);
      const float mm = g.mask_mean[b * g.mask_img_stride + (j % g.hw)];
      const float s = g.pct[b];
      const float m = mm > s ? 1.f : (mm < s ? 0.f : mm);                       // mask[mask > s] = 1; mask[mask < s] = 0
      const float blur_x = __fadd_rn(g.blur[i], __fdiv_rn(g.sigma_hat, eu));    // blur_sample + (sigma_hat / out[0]), as written
      float masked = __fmul_rn(blur_x, m);
      if (g.mode == 2) masked = __fmul_rn(masked, g.c_in);
      float gx = __fadd_rn(masked, __fmul_rn(g.x[i], __fsub_rn(1.f, m)));
      if (g.mode == 1) gx = __fmul_rn(gx, g.c_in);
      g.out[i] = gx;
    } else {
      const float ea = eps_at(g.e_attn, g.eps_dtype, i);
      const float mixed = __fadd_rn(ea, __fmul_rn(g.scale, __fsub_rn(g.sum16[i], ea)));  // e_attn + scale * (sum_e_t - e_attn)
      g.out[i] = __fmul_rn(g.guidance, mixed);                                          // uc_scale * sum_e_t (:514), fp32
    }
  }
}

}  // namespace

extern "C" cpd_status cpd_gaussian_blur(const float* src, float* dst, int n_images, int planes_per_image, int64_t img_stride_src,
                                        int64_t img_stride_dst, int h, int w, const float* taps_host, int ksize, void* stream) {
  CPD_REQUIRE(src && dst && taps_host, "cpd_gaussian_blur: null pointer");
  CPD_REQUIRE(ksize >= 1 && ksize <= MAX_TAPS && (ksize & 1), "cpd_gaussian_blur: kernel size %d must be odd and <= %d", ksize, MAX_TAPS);
  CPD_REQUIRE(h > ksize / 2 && w > ksize / 2, "cpd_gaussian_blur: reflect padding of %d needs a plane larger than that (%d x %d)", ksize / 2, h, w);
  CPD_REQUIRE(n_images >= 0 && planes_per_image > 0, "cpd_gaussian_blur: bad plane counts");
  if (n_images == 0) return CPD_OK;
  BlurTaps t;
  for (int i = 0; i < MAX_TAPS; ++i) t.k[i] = i < ksize ? taps_host[i] : 0.f;
  const int64_t total = (int64_t)n_images * planes_per_image * h * w;
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  CPD_CUDA_CHECK(cpd_launch(gaussian_blur_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, src, dst, planes_per_image,
                            img_stride_src, img_stride_dst, h, w, ksize, t, total));
  return CPD_OK;
}

extern "C" cpd_status cpd_channel_mean(const void* a, int64_t pixels, int c, int act_fp16, float* out, void* stream) {
  CPD_REQUIRE(a && out, "cpd_channel_mean: null pointer");
  CPD_REQUIRE(pixels >= 0 && c > 0 && c % 8 == 0, "cpd_channel_mean: pixels=%lld c=%d (c must be a multiple of 8)", (long long)pixels, c);
  if (pixels == 0) return CPD_OK;
  int64_t blocks = (pixels * 32 + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (act_fp16)
    CPD_CUDA_CHECK(cpd_launch(channel_mean_kernel<true>, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, (const uint4*)a, pixels, c / 8,
                              1.0f / (float)c, out));
  else
    CPD_CUDA_CHECK(cpd_launch(channel_mean_kernel<false>, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, (const uint4*)a, pixels, c / 8,
                              1.0f / (float)c, out));
  return CPD_OK;
}

extern "C" cpd_status cpd_attn_guide(int stage, int n_images, int hw, const float* x, const void* eps_u, int eps_dtype, int64_t eps_stride,
                                     const float* mask_mean, int64_t mask_img_stride, const float* pct, const float* blur, const float* sum16,
                                     const void* e_attn, float sigma_hat, float c_in, int mode, float scale, float guidance, float* out,
                                     void* stream) {
  CPD_REQUIRE(stage >= 0 && stage <= 2 && out, "cpd_attn_guide: stage=%d", stage);
  CPD_REQUIRE(n_images >= 0 && hw > 0, "cpd_attn_guide: n_images=%d hw=%d", n_images, hw);
  CPD_REQUIRE(eps_dtype == CPD_F32 || eps_dtype == CPD_F16 || eps_dtype == CPD_BF16, "cpd_attn_guide: eps_dtype=%d", eps_dtype);
  if (stage <= 1) CPD_REQUIRE(x && eps_u, "cpd_attn_guide: stage %d needs x and eps_u", stage);
  if (stage == 1) CPD_REQUIRE(mask_mean && pct && blur && (mode == 1 || mode == 2), "cpd_attn_guide: stage 1 needs mask_mean, pct, blur and mode 1 / 2");
  if (stage == 2) CPD_REQUIRE(sum16 && e_attn, "cpd_attn_guide: stage 2 needs sum16 and e_attn");
  if (n_images == 0) return CPD_OK;
  GuideArgs g;
  g.stage = stage; g.n_images = n_images; g.hw = hw; g.eps_dtype = eps_dtype; g.mode = mode; g.eps_stride = eps_stride;
  g.x = x; g.eps_u = eps_u; g.mask_mean = mask_mean; g.mask_img_stride = mask_img_stride; g.pct = pct; g.blur = blur; g.sum16 = sum16;
  g.e_attn = e_attn; g.sigma_hat = sigma_hat; g.c_in = c_in; g.scale = scale; g.guidance = guidance; g.out = out;
  const int64_t total = (int64_t)n_images * 4 * hw;
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  CPD_CUDA_CHECK(cpd_launch(attn_guide_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, g));
  return CPD_OK;
}
