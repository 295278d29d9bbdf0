/*
 * cpd_b200.h - C ABI of libcpd_b200.so: the B200 (sm_100a) kernels behind the denoising-loop hot path
 * of milesgray/complex_prompt_diffusion (reference paths below are relative to /root/reference).
 *
 * The reference is pure Python/PyTorch and has no FFI; the boundary these entry points replace is the
 * set of torch calls made by
 *   cpd/samplers/extension/denoiser.py:450-463,508-515,528-544   (CFG combine, eps/v -> denoised)
 *   cpd/samplers/euler.py:47-54,83-92 and cpd/samplers/dpmpp.py:39-54 (sampler updates)
 *   cpd/models/unet.py:765-831 + attention.py + models/util.py    (UNet forward)
 * and the binding a reference maintainer adds is the ctypes stub shown in INTEGRATION.md.
 *
 * Conventions: plain pointers and sizes only (no torch types); every pointer is a DEVICE pointer unless
 * stated otherwise; `stream` is a cudaStream_t passed as void*; every function returns 0 on success or a
 * non-zero cpd_status and never throws; cpd_last_error() describes the last failure of the calling thread.
 * No function allocates device memory or synchronises the device.
 */
#ifndef CPD_B200_H
#define CPD_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int cpd_status; /* 0 = ok, 1 = invalid argument, 2 = CUDA error, 3 = unsupported */

const char* cpd_last_error(void);
int cpd_abi_version(void);

/* dtypes of eps buffers */
enum { CPD_F32 = 0, CPD_F16 = 1, CPD_BF16 = 2 };
/* sampler kinds (registry names "Euler", "Euler Ancestral", "DPM++ 2m") */
enum { CPD_EULER = 0, CPD_EULER_ANCESTRAL = 1, CPD_DPMPP_2M = 2,
       CPD_DENOISE_ONLY = 3, /* Denoiser.forward alone: x is not updated, denoised_out receives the sample */
       CPD_HEUN2 = 4,        /* second stage of Heun (huen.py:52-57): x_base + ((d_prev[0] + d_2) / 2) * dt */
       CPD_LMS = 5 };        /* linear multistep (lms.py:40-52): x + sum_j lms_coeff[j] * d_{i-j} */
/* prediction type (denoiser.py:537-542) */
enum { CPD_PRED_EPSILON = 0, CPD_PRED_VELOCITY = 1 };

#define CPD_MAX_SUBPROMPTS 16

/*
 * Fused multi-prompt CFG combine + eps/v -> denoised + sampler update: ONE pass over the latents.
 * Replaces denoiser.py:450-460 (fp16 weighted delta), :510-515 (e_t = e_u + s * sum), :533-542
 * (denoised), and the update of euler.py:49-54 / euler.py:85-92 / dpmpp.py:42-54.
 *
 * eps holds the UNet outputs for `n_images` images x (1 + n_sub) rows x L elements
 * (L = 4 * hw); row 0 of each image is the unconditional row (n_sub = 0: the single row already is e_t, e.g. after a score
 * corrector rewrote it, denoiser.py:517-518; the combine is skipped).  Element (b, r, i) lives at
 * eps + (b * eps_image_stride + r * eps_row_stride + i) elements.
 * All scalars are the fp32 values the reference computes on 0-dim tensors (host-side, see
 * complex_prompt_diffusion_b200/samplers/k_diffusion.py).
 */
/* The per-step scalars of a sampler step as ONE row of a device table (cpd_step_select): a host fills a table for the whole
 * schedule at the start of sample() - the same fp32 values it would pass by value - and a captured CUDA graph of
 * [cpd_step_select -> cpd_unet_forward -> cpd_sampler_step] then replays once per step with no host work in between. */
typedef struct {
  float c_in;               /* UNet input scale 1 / sqrt(sigma^2 + 1) (denoiser.py:390) */
  float t;                  /* timestep, rounded to the model dtype (denoiser.py:393) */
  float guidance, sigma_hat, v_c_eps, v_c_x_div, dt, sigma_up;  /* as in cpd_step_params */
  float dpm_ratio, dpm_expm1, dpm_c1, dpm_c2;
  int dpm_first, write_old;
  float noise_mul;
  int reserved;
} cpd_step_scalars;

typedef struct {
  const void* eps;
  int eps_dtype;            /* CPD_F32 / CPD_F16 / CPD_BF16 */
  int64_t eps_image_stride; /* elements */
  int64_t eps_row_stride;   /* elements */
  float* x;                 /* [n_images][L] in/out, fp32 (k_diffusion.py:73-74 keeps x fp32) */
  float* old_denoised;      /* [n_images][L] in/out, DPM++ 2M history; may be NULL otherwise */
  const float* noise;       /* [n_images][L] ancestral noise (euler.py:92); NULL otherwise */
  float* denoised_out;      /* optional [n_images][L] copy of the denoised sample (callback 'eps'), or NULL */
  float* eps_out;           /* optional [n_images][L] combined e_t (fp32), or NULL */
  int n_images;
  int n_sub;                /* N weighted sub-prompts (conjunctions + negations), 1..CPD_MAX_SUBPROMPTS */
  int hw;                   /* latent pixels per channel; L = 4 * hw */
  float weights[CPD_MAX_SUBPROMPTS];     /* scale (negated for "not"), ALREADY rounded to the model dtype (P4) */
  float mask_scalar[CPD_MAX_SUBPROMPTS]; /* scalar mask value (1 when the sub-prompt has no mask) */
  const float* masks[CPD_MAX_SUBPROMPTS];/* optional spatial mask [hw] per sub-prompt (broadcast over the 4 channels, */
                                         /* values already rounded to the model dtype) or NULL */
  float guidance;           /* unconditional_guidance_scale after optional decay (denoiser.py:475-494) */
  int sampler;              /* CPD_EULER / CPD_EULER_ANCESTRAL / CPD_DPMPP_2M */
  int pred_type;            /* CPD_PRED_EPSILON / CPD_PRED_VELOCITY */
  float sigma_hat;          /* sigma_i * (gamma + 1), gamma = 0 */
  float v_c_eps;            /* velocity: -sigma / (sigma^2 + 1)^0.5 */
  float v_c_x_div;          /* velocity: sigma^2 + 1 (x is DIVIDED by it) */
  float dt;                 /* Euler: sigma_{i+1} - sigma_hat; ancestral: sigma_down - sigma_i */
  float sigma_up;           /* ancestral noise scale */
  float dpm_ratio;          /* 2M: sigma_fn(t_next) / sigma_fn(t) */
  float dpm_expm1;          /* 2M: (-h).expm1() */
  float dpm_c1, dpm_c2;     /* 2M: (1 + 1/(2r)), 1/(2r) */
  int dpm_first;            /* 2M: 1 when old_denoised is None or sigma_{i+1} == 0 (dpmpp.py:44) */
  int write_old;            /* 2M: store denoised into old_denoised */
  /* Two-stage and multistep samplers (Heun huen.py:24-58, DPM2 / DPM2-a dpm2.py:24-108, DPM++ 2S-a dpmpp.py:70-113, LMS
   * lms.py:28-52) reuse the modes above with the UNet input, the tensor the update starts from and the destination
   * taken apart: stage 1 = CPD_EULER (or CPD_DPMPP_2M with dpm_first) writing x_2 to x_out; stage 2 evaluates x = x_2
   * and updates x_base. */
  const float* x_base;      /* sample the update starts from; NULL = x (the UNet input) */
  float* x_out;             /* destination of the updated sample; NULL = x (in place) */
  float* d_out;             /* optional [n_images][L]: d = (x - denoised) / sigma_hat of this evaluation (to_ode) */
  const float* d_prev[3];   /* CPD_HEUN2: d of stage 1 in d_prev[0]; CPD_LMS: d_{i-1}, d_{i-2}, d_{i-3} */
  float lms_coeff[4];       /* CPD_LMS: coefficients of d_i, d_{i-1}, d_{i-2}, d_{i-3} (linear_multistep_coeff) */
  int lms_order;            /* CPD_LMS: min(i + 1, order), 1..4 */
  float noise_mul;          /* noise is multiplied by this (s_noise / temperature) before sigma_up; used as given (set 1 for plain noise) */
  /* Thresholding extensions (samplers/extension/threshold.py:65-88 "dynamic_thresholding", 47-63 "static_thresholding"),
   * with the clamp bound s = max(percentile(|.|), 1) computed ON THE DEVICE by cpd_abs_percentile_max1: */
  const float* clip_scaled; /* optional DEVICE array [n_images] (from cpd_threshold): the scaled guidance term s * sum_e_t of
                               image b is clamped to [-c[b], c[b]] and rounded to fp16 before it is added to e_u
                               (denoiser.py:510-515, scaled_clip) */
  float* scaled_out;        /* optional [n_images][L] fp32 copy of half(s * sum_e_t) BEFORE clamping (input of the
                               percentile pass); with sampler = CPD_DENOISE_ONLY nothing else needs to be written */
  const float* scaled_in;   /* optional [n_images][L]: the scaled guidance term already processed by a thresholding
                               extension that is not a clamp (cpd_threshold_ex on a scaled_out copy); replaces
                               s * sum_e_t in e_t = e_u + scaled (denoiser.py:510-515) */
  const cpd_step_scalars* dyn; /* optional DEVICE pointer: guidance, sigma_hat, v_c_*, dt, sigma_up, dpm_*, write_old and noise_mul are
                               read from it by the kernel instead of from the fields above (the "current" row cpd_step_select
                               wrote), so one captured launch serves every step of a schedule */
} cpd_step_params;

cpd_status cpd_sampler_step(const cpd_step_params* p, void* stream);

/* current[0] = table[min(*counter, n_steps - 1)]; *counter += 1 - one tiny kernel at the head of a captured sampler-step graph.
 * table, counter and current are device memory; &current->c_in and &current->t feed cpd_unet_forward's c_in / t, current feeds
 * cpd_step_params.dyn. */
cpd_status cpd_step_select(const cpd_step_scalars* table, int n_steps, int* counter, cpd_step_scalars* current, void* stream);

/* Stochastic churn of Karras et al. Algorithm 2 before the denoiser call (euler.py:43-46, huen.py:40-43, dpm2.py:40-43):
 * x[i] = x[i] + (noise[i] * noise_mul) * scale with separately rounded fp32 operations (noise_mul = s_noise,
 * scale = sqrt(sigma_hat^2 - sigma^2)); n elements, n % 4 == 0. */
cpd_status cpd_add_noise(float* x, const float* noise, float noise_mul, float scale, int64_t n, void* stream);

/*
 * Thresholding extensions on the device (samplers/extension/threshold.py; hooks denoiser.py:510-512, euler.py:55-56,
 * dpmpp.py:51-52,92-93).  The reference copies the tensor to the CPU for np.percentile every step.
 *   alg = CPD_THRESH_DYNAMIC ("dynamic_thresholding", threshold.py:65-88): for each of the n_images rows of
 *     x [n_images][L] (fp32) the `threshold`-th percentile (0..100, numpy's default linear interpolation between order
 *     statistics, evaluated in fp32 like np.percentile on a float32 array) of |x| is found by an exact 4-pass radix select;
 *     bound[b] = max(percentile_b, 1.0).  Images are independent trajectories here, so the bound is per
 *     image; with n_images = 1 this is exactly np.max(np.append(s, 1.0)).
 *   alg = CPD_THRESH_STATIC ("static_thresholding", threshold.py:47-63): bound[b] = threshold.
 * bound: [n_images] fp32 device array (always written).  If clamp_inplace != 0, x is then clamped to
 * [-bound[b], bound[b]] and rounded through fp16 (the reference returns x.half()).
 */
enum { CPD_THRESH_DYNAMIC = 0, CPD_THRESH_STATIC = 1 };
cpd_status cpd_threshold(float* x, int n_images, int L, int alg, float threshold, int clamp_inplace, float* bound, void* stream);

/*
 * The other runnable thresholding extensions of samplers/extension/threshold.py, in place on x [n_images][channels][hw]
 * (fp32 values; the result is rounded through fp16 because the reference returns x.half()).  One image = one independent
 * trajectory (x.max() / x.min() / the quantile are per image; identical to the reference at its batch size of 1).
 *   CPD_THRESH_DYNANORMIC            "dynanormic_thresholding" (:87-116): s = max(torch.quantile(|x|, q), 1); x = clamp(x, -s, s) / s
 *   CPD_THRESH_SCALED_DYNAMIC_PERC   "scaled_dynamic_perc_thresholding" (:118-146): y = 2 (x - min) / (max - min) - 1;
 *                                    s = max(np.percentile(|y|, threshold), 1); clamp; x = (max - min) (y + 1) / 2 + min
 *   CPD_THRESH_RENORM                "renorm_thresholding" (:148-180): the same with torch.quantile
 *   CPD_THRESH_SCALED_NORM           "scaled_norm_thresholding" (:207-237): thr = fl32(threshold / 100) * max;
 *                                    s = max(sqrt(mean(y^2)), thr); y *= thr / s; unscale
 *   CPD_THRESH_SPATIAL_NORM          "spatial_norm_thresholding" (:239-254): per pixel s = max(sqrt(mean_c x^2), threshold); x *= threshold / s
 *   CPD_THRESH_SCALED_SPATIAL_NORM   "scaled_spatial_norm_thresholding" (:256-286): min-max rescaled, thr as in SCALED_NORM
 * ("norm_thresholding", :182-205, reads an undefined x_max and cannot run in the reference.)
 * threshold is the reference's Python float (a double): quantiles accept 0..1 or 1..100 (divided by 100 like :100-101).
 * bound [n_images] receives the per-image s (the maximum over pixels for the spatial variants).
 * The quantile variants are bit-exact against torch.quantile / np.percentile on the CPU.  The RMS variants agree to one
 * fp16 ulp on rare rounding-boundary elements: the device uses the IEEE sqrt where torch's CPU sqrt goes through MKL VML
 * (below 1 ulp, not correctly rounded), and SCALED_NORM's per-image mean is an fp64 fixed-order sum (the eager fp32
 * summation order is not a contract).
 */
enum {
  CPD_THRESH_DYNANORMIC = 2,
  CPD_THRESH_SCALED_DYNAMIC_PERC = 3,
  CPD_THRESH_RENORM = 4,
  CPD_THRESH_SCALED_NORM = 5,
  CPD_THRESH_SPATIAL_NORM = 6,
  CPD_THRESH_SCALED_SPATIAL_NORM = 7
};
cpd_status cpd_threshold_ex(float* x, int n_images, int channels, int hw, int alg, double threshold, float* bound, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * UNet building blocks (all activations NHWC bf16 = row-major [pixels, channels]).  The Python host
 * (complex_prompt_diffusion_b200/models/unet.py) sequences them exactly like unet.py:765-831.
 * --------------------------------------------------------------------------------------------------------- */

/* Epilogue flags for cpd_gemm_conv */
enum { CPD_EPI_NONE = 0, CPD_EPI_GEGLU = 1 };

/*
 * Implicit-GEMM convolution / GEMM on tcgen05 (TMEM accumulators, TMA-fed):
 *   D[m, n] = sum_{tap, c} A[pixel(m) shifted by tap, c] * Wt[n, tap * C + c]  (+ bias[n]) (+ rowvec[img(m), n])
 *             (+ residual[m, n])
 * A is one or two NHWC bf16 tensors concatenated along channels (the UNet skip concat, unet.py:814):
 * channels [0, c0) come from a0, [c0, c0 + c1) from a1 (c1 may be 0).  c0, c1 multiples of 64.
 * ksize = 1 (plain GEMM / 1x1 conv, unet.py:247, attention.py:183-190,508-524) or 3 (pad 1; stride 1 or 2,
 * unet.py:105,153-160,210,236).  Output pixels = n_img * (h_in/stride) * (w_in/stride).
 * For a plain GEMM use n_img = 1, h_in = 1, w_in = M.
 * wt: [n_out][ksize*ksize][c0 + c1] bf16 (K-major).  bias: fp32 [n_out] or NULL.
 * rowvec: fp32, element (img, n) at rowvec[img * rowvec_stride + n] (time-embedding add, unet.py:266-274) or NULL.
 * residual: bf16 [pixels][ld_res] or NULL.  d: bf16 [pixels][ldd].
 * CPD_EPI_GEGLU: wt rows are interleaved per `geglu_block`-column tile (128 or 256) as
 * [geglu_block/2 value rows | geglu_block/2 gate rows]; d gets n_out/2 columns: value * gelu(gate)
 * (attention.py:92-100).
 *
 * variant 0 (auto) runs the persistent CTA-pair kernel (tcgen05 cta_group::2, 256 x BN tiles, double-buffered TMEM
 * accumulators, gemm_umma2.cu); variants 1 / 2 run the one-tile-per-CTA 128 x 128 / 128 x 256 kernel
 * (gemm_umma.cu); variant 32..256 forces the pair kernel's tile width BN = variant (a multiple of 32); 2000 + BN: 256 x 2BN tiles
 * (two sub-tiles, one TMEM accumulator stage); 1000 + BN: two-pair cluster with TMA-multicast A; + 10000 * S: split-K over S
 * K ranges (small-M layers that cannot fill 74 SM pairs), partial sums reduced through splitk_ws by a second kernel.
 */
typedef struct {
  const void* a0; const void* a1;
  int c0, c1;
  int n_img, h_in, w_in;
  int ksize, stride;
  const void* wt;
  int n_out;
  const float* bias;
  const float* rowvec; int rowvec_stride;
  const void* residual; int ld_res;
  void* d; int ldd;
  int epilogue;
  int variant; /* 0 = auto (CTA-pair kernel); 1 = 128x128 1-CTA tile, 2 = 128x256 1-CTA tile; >= 32: pair kernel, BN = variant */
  int m_valid; /* plain GEMM only: number of valid rows (<= w_in); 0 = all */
  int a_fp16;  /* element type of a0/a1: 1 = fp16, 0 = bf16 (16-bit either way; tcgen05 kind::f16 takes both) */
  int b_fp16;  /* element type of wt */
  int out_fp16;/* element type of d and residual */
  int geglu_block; /* CPD_EPI_GEGLU: interleave block of wt rows (128 or 256; 0 = 128) */
  float* splitk_ws;         /* optional fp32 scratch for split-K launches (variant >= 10000): >= S * pixels * n_out floats */
  int64_t splitk_ws_floats; /* (one slice per K range, summed in a fixed order by the finalize kernel: no atomics) */
  void* tune_scratch;       /* optional scratch the size of d: with variant 0 an unseen problem shape is first TIMED on every */
  int64_t tune_scratch_bytes; /* applicable tile variant (output redirected here, CUDA events, synchronous) and the fastest is */
                            /* remembered per shape; NULL = no timing (tuned table if the shape is known, else the cost model) */
  /* --- LayerNorm folded into the neighbouring GEMMs (attention.py:476-490: x + attn(norm(x)), x + ff(norm(x))) ---------------
   * LN is affine per row, so LN(x) W^T + b = rstd_m * (x (gamma . W)^T - mean_m * g) + (W beta + b) with g[n] = sum_k (gamma . W)[n][k]:
   * the CONSUMER GEMM runs on the raw residual stream with pre-scaled weights and applies the per-row scale / shift in its
   * epilogue; the PRODUCER GEMM that wrote the residual stream emits per-row partial (sum, sum of squares) of its output, one
   * float2 per (column tile, epilogue group) and row, laid out [part][ln_ld] (rows contiguous: coalesced), summed by the consumer
   * in a fixed order.  Pair kernel only (no split-K, no one-tile variants); n_out % 32 == 0 for a producer. */
  float* ln_sums_out;       /* producer: float2 [parts][ln_ld] or NULL */
  int* ln_parts_out;        /* producer: HOST int that receives the number of parts this launch writes */
  const float* ln_sums;     /* consumer: the producer's partials, or NULL */
  int ln_parts;             /* consumer: number of parts to sum */
  int64_t ln_ld;            /* rows per part (>= output rows), both sides */
  const float* ln_g;        /* consumer: fp32 [n_out], row sums of the folded weights (wt = gamma . W; bias = W beta + b is required) */
  int ln_c;                 /* consumer: normalised channels (the K of the GEMM) */
  float ln_eps;
  /* --- transposed tail: columns [dt_col0, n_out) of D are stored TRANSPOSED into d_t[(n - dt_col0)][m] (leading dimension ldd_t
   * elements, m = output row) instead of d: the V^T operand of the self-attention comes out of the fused Q | K | V projection.
   * Plain GEMMs (ksize 1, n_img = h_in = 1) on the pair kernel only; dt_col0 % 32 == 0.  NULL = off. */
  void* d_t;
  int dt_col0;
  int64_t ldd_t;
  /* --- GroupNorm statistics of D from the epilogue that writes it (models/util.py:95-105: the next GroupNorm then only
   * normalises - cpd_groupnorm_apply - instead of reading the tensor twice): per (image, channel) sum and sum of squares in
   * 64-bit FIXED POINT (sum * 2^24, sum of squares * 2^12), added with integer atomics - exact and order-independent, so the
   * result is bit-reproducible.  [n_img][n_out][2]; the caller zeroes it before the launch.  Pair kernel only (no split-K, no
   * one-tile variants, no GEGLU); every 32 consecutive rows of a CTA tile must belong to one image (th * tw % 32 == 0). */
  long long* gn_sums_out;
} cpd_gemm_params;

cpd_status cpd_gemm_conv(const cpd_gemm_params* p, void* stream);

/* Runtime switch of the per-shape timing of cpd_gemm_conv's variant 0 (initial state: CPD_GEMM_AUTOTUNE, default on) and a reset of
 * the tuned table.  With the timing off variant 0 resolves through the deterministic cost model. */
void cpd_gemm_set_autotune(int on);
void cpd_gemm_tune_clear(void);
/* The per-shape variant table behind variant 0.  Timing picks between candidates that are often within a few per cent, so two
 * processes can choose differently (different fp32 summation order at the 1e-7 level): export the table of one process and
 * import it into the others (ranks of a job) - or set CPD_GEMM_AUTOTUNE=0 - when bit-identical results across processes
 * matter.  Export writes a NUL-terminated text ("k0,k1,...=variant" lines) and returns the bytes needed (call with cap = 0
 * to size the buffer); import merges a text into the table and returns the number of entries read or -1. */
int64_t cpd_gemm_tune_export(char* buf, int64_t cap);
int cpd_gemm_tune_import(const char* text);

/*
 * GroupNorm(32 groups) [+ SiLU] over NHWC bf16, fp32 statistics (models/util.py:95-105, attention.py:89-90).
 * Input channels may come from two tensors (skip concat).  stats: fp64 scratch of n_img * CPD_GN_MAX_CHUNKS * 64
 * doubles (per-chunk partial sums of the two-launch path, reduced in a fixed order; the one-launch kernel reduces in shared /
 * distributed shared memory in a fixed order: no atomics on either path, bit-reproducible).  out: 16-bit [n_img*hw][c0 + c1].
 */
#define CPD_GN_MAX_CHUNKS 64
cpd_status cpd_groupnorm(const void* a0, const void* a1, int c0, int c1, int n_img, int hw, const float* gamma,
                         const float* beta, float eps, int silu, int act_fp16, double* stats, void* out, void* stream);
/* How many kernels cpd_groupnorm launches for c = c0 + c1 channels: 1 when the (image, group slab) tiles fit the shared-memory /
 * cluster kernel (one read + one write of the tensor; `stats` is then unused), 2 (statistics + apply) otherwise; 0 = bad shape. */
int cpd_groupnorm_launches(int c, int n_img, int hw);

/* GroupNorm(32) (+ SiLU) from per-(image, channel) fixed-point statistics emitted by the producing GEMM (cpd_gemm_params.gn_sums_out):
 * one read + one write of the tensor.  x, out: NHWC 16-bit [n_img][hw][c]. */
cpd_status cpd_groupnorm_apply(const void* x, int c, int n_img, int hw, const float* gamma, const float* beta, float eps, int silu,
                               int act_fp16, const long long* chan_sums, void* out, void* stream);
/* LayerNorm over the last dim of [rows][c] (attention.py:476-478), eps 1e-5.
 * act_fp16 (here and below): activation tensors are 16-bit, 1 = fp16, 0 = bf16.  Weights are always bf16. */
cpd_status cpd_layernorm(const void* x, int rows, int c, const float* gamma, const float* beta, float eps, int act_fp16,
                         void* out, void* stream);

/* Sinusoidal timestep embedding (models/util.py:65-85): t [rows] fp32 -> out bf16 [rows][dim], cos first.
 * t is first rounded to the model dtype (bf16) as denoiser.py:393 does. */
cpd_status cpd_timestep_embedding(const float* t, int rows, int dim, int round_t_bf16, void* out, void* stream);

/* Small-M linear: out[m][n] = sum_k act(x[m][k]) * w[n][k] + b[n]; x bf16 [m][k], w bf16 [n][k], m <= 32.
 * silu_in applies SiLU to x (unet.py:223-229 emb_layers).  out_f32 != NULL -> fp32 [m][ld_out];
 * out_bf16 != NULL -> bf16 [m][ld_out]. */
cpd_status cpd_small_linear(const void* x, int m, int k, const void* w, const float* b, int n, int silu_in,
                            float* out_f32, void* out_bf16, int ld_out, void* stream);

/* Input conv 3x3 (unet.py:548) fused with the Denoiser's input scaling and row broadcast (denoiser.py:390-391):
 * x fp32 NCHW [n][cin][h][w]; every image is multiplied by `scale` (c_in, fp32), cast to bf16 (unet.py:794) and
 * convolved; the result is written `rows_per_image` times (one copy per conditioning row), image-major:
 * out bf16 NHWC [n * rows_per_image][h][w][cout]; w bf16 [cout][3][3][cin]; cin <= 8.
 * scale_ptr: optional DEVICE pointer to the scale (overrides `scale`), so a captured CUDA graph replays with a new c_in. */
cpd_status cpd_conv_in(const float* x, int n, int cin, int h, int w, const void* wt, const float* bias, int cout,
                       float scale, const float* scale_ptr, int rows_per_image, int act_fp16, void* out, void* stream);

/* Output conv 3x3 (unet.py:729-733): a bf16 NHWC [n][h][w][cin] (already GroupNorm+SiLU) -> out NCHW
 * [n][cout][h][w] in out_dtype (CPD_BF16 or CPD_F32); cout <= 8. */
cpd_status cpd_conv_out(const void* a, int n, int h, int w, int cin, const void* wt, const float* bias, int cout,
                        void* out, int out_dtype, int act_fp16, void* stream);

/* Nearest 2x upsample of NHWC bf16 (unet.py:116). */
cpd_status cpd_upsample2x(const void* a, int n, int h, int w, int c, void* out, void* stream);

/*
 * Fused flash-style attention on tcgen05 (attention.py:283-296,327,337,340,345-348), one launch for all
 * (batch, head) pairs:  O = softmax(Q K^T * scale) V.
 * q : bf16, row (b * nq + i) at q + row * ldq, head h at columns [h*dpad, h*dpad + dpad)   (pad columns zero)
 * k : bf16, row (b * nk_pad + j) at k + row * ldk, same column layout
 * vt: bf16 V transposed: row (h * dpad + c) at vt + row * ldvt, column (b * nk_pad + j)
 * o : bf16, row (b * nq + i) at o + row * ldo, head h at columns [h*dpad, ...)
 * nk = valid keys per batch (<= nk_pad); dpad multiple of 16, <= 160.
 */
typedef struct {
  const void* q; int ldq;
  const void* k; int ldk;
  const void* vt; int ldvt;
  void* o; int ldo;
  int batch, heads, nq, nk, nk_pad, dpad;
  float scale; /* dim_head ** -0.5 */
  int kv_batch; /* number of distinct K/V batches: query batch b reads K/V batch (b % kv_batch); 0 = batch */
  int act_fp16; /* q, k, vt, o element type: 1 = fp16, 0 = bf16 */
  int d_head;   /* real head dim (<= dpad); > 0 enables the two-query-tile kernel (attention_umma2.cu) for nq > 128 and
                   d_head <= 111, which loads only d_head rows of V^T per head; 0 = one-tile kernel */
} cpd_attn_params;

cpd_status cpd_attention(const cpd_attn_params* p, void* stream);

/* ---- first-stage decoder (VAE decode, cpd/models/autoencoder.py:380-509,825-828: SURVEY.md 8-f row 3) ------------------
 * The decoder reuses cpd_gemm_conv / cpd_groupnorm / cpd_upsample2x / cpd_conv_in / cpd_conv_out; two small ops are its own: */

/* Row softmax of a 16-bit row-major matrix: out[r][c] = softmax_c(scale * x[r][c]) (fp32 arithmetic), the attention weights
 * of AttnBlock (autoencoder.py:250-255; one head of width C, so Q K^T and P V are plain cpd_gemm_conv GEMMs).
 * cols: multiple of 8, <= 16384; ld / ldo in elements; scale > 0; out may alias x. */
cpd_status cpd_softmax_rows(const void* x, int rows, int cols, int64_t ld, float scale, int act_fp16, void* out, int64_t ldo,
                            void* stream);

/* 1x1 convolution of an fp32 NCHW tensor with <= 8 channels (post_quant_conv, autoencoder.py:800,826), with the
 * 1 / scale_factor of decode_first_stage folded in: out[n][o][p] = b[o] + sum_c w[o][c] * (x[n][c][p] * scale). */
cpd_status cpd_pointwise_small(const float* x, int n, int cin, int cout, int64_t hw, const float* w, const float* b, float scale,
                               float* out, void* stream);

/* ---- optional guidance branches of the Denoiser, all on the device -----------------------------------
 * Depthwise Gaussian blur of fp32 planes with reflect padding = torchvision.transforms.GaussianBlur as used for the
 * unconditional blur (denoiser.py:337,441-442) and the attention-guidance blur (:350,420): the 2-D kernel is the outer product of
 * the 1-D taps (each product rounded to fp32, like torch.mm), fp32 accumulation.  Plane p of image b lives at
 * src + b * img_stride_src + p * h * w.  taps: HOST array of ksize fp32 values (odd ksize <= 63; h, w > ksize / 2). */
cpd_status cpd_gaussian_blur(const float* src, float* dst, int n_images, int planes_per_image, int64_t img_stride_src,
                             int64_t img_stride_dst, int h, int w, const float* taps, int ksize, void* stream);
/* out[px] = mean over the c channels of an NHWC activation tensor [pixels][c] (attn.mean(1, keepdims=True), denoiser.py:408). */
cpd_status cpd_channel_mean(const void* a, int64_t pixels, int c, int act_fp16, float* out, void* stream);
/* out[0] = np.percentile(x[0..n), q) of SIGNED fp32 values (linear interpolation; the saliency threshold, denoiser.py:411). */
cpd_status cpd_percentile(const float* x, int n, float q, float* out, void* stream);
/* The elementwise stages of attention guidance over n_images images of 4 * hw elements (denoiser.py:412-429,461-462,514):
 *   stage 0: out = x - sigma_hat * e_u                                          (the denoised sample that gets blurred)
 *   stage 1: m = mean > pct ? 1 : (mean < pct ? 0 : mean);  bx = blur + sigma_hat / e_u;
 *            out = [bx * m (* c_in if mode 2)] + x * (1 - m)   (* c_in if mode 1)            (the guided latent)
 *   stage 2: out = guidance * (e_attn + scale * (sum16 - e_attn))               (the mixed, scaled guidance term)
 * e_u of image b starts at eps_u + b * eps_stride elements (eps_dtype); mask_mean holds the channel means of the saliency
 * source, the row of image b at mask_mean + b * mask_img_stride; pct[b] its percentile. */
cpd_status cpd_attn_guide(int stage, int n_images, int hw, const float* x, const void* eps_u, int eps_dtype, int64_t eps_stride,
                          const float* mask_mean, int64_t mask_img_stride, const float* pct, const float* blur, const float* sum16,
                          const void* e_attn, float sigma_hat, float c_in, int mode, float scale, float guidance, float* out,
                          void* stream);

/* The latents -> images tail after the decoder (cpd/embeddings/prompts.py:472-475):
 * out[n][p][ch] = uint8(clamp((x[n][ch][p] + 1) / 2, 0, 1) * 255), i.e. NCHW fp32 in (channel stride of an image = hw, image
 * stride = ld_c * hw: the decoder's 4-channel output buffer holds 3-channel images), NHWC uint8 out; c <= 8. */
cpd_status cpd_images_to_uint8(const float* x, int n, int c, int64_t hw, int ld_c, uint8_t* out, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Plan-level UNet entry points: ONE call evaluates the whole UNet of cpd/models/unet.py:765-831.
 * The plan owns the packed weights, the text-context K / V^T cache, every workspace buffer and (optionally) a CUDA graph of
 * the ~850 kernel launches of an evaluation.  A non-Python host needs nothing else:
 *     cpd_unet_plan_create -> cpd_pack_weights (once per state_dict entry) -> cpd_cache_context_kv (once per prompt)
 *     -> cpd_unet_forward (once per sampler step) -> cpd_unet_plan_destroy.
 * One plan per device (the device current at creation); re-entrant per plan, not thread-safe on one plan.
 * Device memory is allocated at creation, while packing, when a context is cached and the FIRST time a new
 * (rows, h, w) shape is evaluated (workspace + optional graph capture) - never inside a steady-state forward.
 * --------------------------------------------------------------------------------------------------------- */
typedef struct cpd_unet_plan cpd_unet_plan;

#define CPD_UNET_MAX_LEVELS 8
typedef struct {
  /* constructor arguments of UNetModel (unet.py:415-470; config-1.49.yaml:27-42, v2-inference.yaml:20-37) */
  int in_channels, out_channels, model_channels, num_res_blocks;
  int n_levels;
  int channel_mult[CPD_UNET_MAX_LEVELS];
  int n_attention_resolutions;
  int attention_resolutions[CPD_UNET_MAX_LEVELS]; /* downsample factors (1, 2, 4, ...) that carry a SpatialTransformer */
  int num_heads;          /* used when num_head_channels == -1 */
  int num_head_channels;  /* -1 or the fixed head width (SD-2.x / SDXL: 64) */
  int transformer_depth[CPD_UNET_MAX_LEVELS];     /* per level; the middle block uses the last level's (the reference has one int) */
  int context_dim;
  int use_linear_in_transformer; /* accepted for the state_dict shapes: a 1x1 conv and a Linear are the same GEMM here */
  int adm_in_channels;    /* > 0: vector conditioning y through label_emb (SDXL extension), 0 otherwise */
  int act_fp16;           /* inter-kernel activations / tensor-core operands: 1 = fp16 (default of the Python host), 0 = bf16 */
  int eps_dtype;          /* CPD_F32 or CPD_BF16: element type of the eps output */
  int use_cuda_graph;     /* 1: every (shape, rows) is captured once into a CUDA graph owned by the plan and replayed */
} cpd_unet_config;

cpd_status cpd_unet_plan_create(const cpd_unet_config* cfg, cpd_unet_plan** plan);
void cpd_unet_plan_destroy(cpd_unet_plan* plan);

/* One state_dict entry under its REFERENCE name ("input_blocks.1.0.in_layers.2.weight", ...; cpd/models/unet.py,
 * attention.py).  `data` holds `numel` elements of dtype CPD_F32 / CPD_F16 / CPD_BF16 on the host (on_device = 0) or the
 * device (1).  Every value is first rounded to the model dtype (bf16, manager.py:25-36), then packed for the kernels:
 * conv weights K-major [Cout][ky][kx][Cin], head dims padded to multiples of 16, Q | K of self-attention fused, GEGLU
 * rows interleaved per 256-column tile, the 22 emb_layers concatenated.  Unknown names return CPD_ERR_INVALID. */
cpd_status cpd_pack_weights(cpd_unet_plan* plan, const char* name, const void* data, int dtype, int64_t numel, int on_device);
/* Number of state_dict entries the plan still misses (0 = ready); the first missing name is left in cpd_last_error(). */
int cpd_unet_plan_missing_weights(cpd_unet_plan* plan);

/* Text context [rows][tokens][context_dim] (device, dtype CPD_F32 / CPD_F16 / CPD_BF16): rounded to the model dtype
 * (denoiser.py:373-385) and projected ONCE through every cross-attention layer's to_k / to_v (attention.py:283-296); the
 * K / V^T buffers are persistent (stable addresses), so captured graphs stay valid across prompts of the same layout.
 * UNet row r attends to context row (r % rows). */
cpd_status cpd_cache_context_kv(cpd_unet_plan* plan, const void* context, int dtype, int rows, int tokens, void* stream);
/* Vector conditioning y [rows][adm_in_channels] (device) of UNets with adm_in_channels > 0; UNet row r uses y row (r % rows). */
cpd_status cpd_unet_set_vector(cpd_unet_plan* plan, const void* y, int dtype, int rows, void* stream);

typedef struct {
  const float* x;        /* device fp32 NCHW [n_images][in_channels][h][w]: the UNSCALED latents (k_diffusion.py:73-74) */
  int n_images, h, w;
  int rows_per_image;    /* every image is evaluated on this many conditioning rows sharing x * c_in (denoiser.py:383-391) */
  const float* c_in;     /* DEVICE pointer to the input scale 1 / sqrt(sigma^2 + 1) (denoiser.py:390), or NULL for 1 */
  const float* t;        /* DEVICE pointer to the timestep(s), already rounded to the model dtype (denoiser.py:393) */
  int t_count;           /* 1: shared by all rows (the Denoiser's call); n_images * rows_per_image: one per row */
  void* eps;             /* device output [n_images * rows_per_image][out_channels][h][w] (image-major rows) in the plan's
                            eps_dtype, or NULL: the result stays in the plan's own buffer (cpd_unet_plan_buffer "eps") */
  /* unet.py:806-813 (feature / skip injection) and :802-804,816-817 (return_attn / return_feat); all optional: */
  const void* const* inject_skips; /* [n output blocks] NHWC activations replacing the popped skip tensor, or NULL entries */
  const void* const* inject_feats; /* [n output blocks] NHWC activations replacing h before the block, or NULL entries */
  int no_graph;          /* 1: launch the kernels directly even if the plan uses CUDA graphs (also implied by injection and
                            by a stream that is being captured by the caller) */
} cpd_unet_io;

/* eps = UNet(x * c_in, t, cached context) for n_images * rows_per_image rows: unet.py:765-831 with the blocks of
 * :249-280, attention.py:280-348,485-490,526-537 and models/util.py:65-85,103-105. */
cpd_status cpd_unet_forward(cpd_unet_plan* plan, const cpd_unet_io* io, void* stream);

/* A named buffer of the plan after a forward: "eps", "<block>.out" (e.g. "input_blocks.4.1.out", NHWC activations: the skip
 * tensors of return_attn and the features of return_feat), packed weights under their packed names.  *ptr / *numel (elements)
 * receive the device pointer and size; the buffer with the largest size is returned when several shapes were evaluated. */
cpd_status cpd_unet_plan_buffer(cpd_unet_plan* plan, const char* name, void** ptr, int64_t* numel);
/* Kernels launched by the last cpd_unet_forward / cpd_cache_context_kv of this plan (graph replays count their nodes). */
int64_t cpd_unet_plan_launches(cpd_unet_plan* plan);
/* Output of block `index` of the LAST forward as an NHWC activation tensor [rows][h][w][channels]: which = 0 input blocks (the
 * 12 skip tensors `return_attn` hands back, unet.py:802-804), 1 the middle block, 2 output blocks (`return_feat`, :816-817).
 * cpd_unet_plan_blocks(plan, which) = how many blocks there are. */
cpd_status cpd_unet_plan_tap(cpd_unet_plan* plan, int which, int index, void** ptr, int* channels, int* h, int* w);
int cpd_unet_plan_blocks(cpd_unet_plan* plan, int which);
/* Per-launch CUDA-event timing of eager forwards (bench.py's roofline leg, tools/profile_layers.py): on = 1 clears the records
 * and brackets every following kernel-level call with events (graphs are bypassed while it is on); the dump synchronises the
 * device and writes "kind\tlabel\tmicroseconds\tflops" lines (NUL-terminated), returning the bytes needed. */
void cpd_unet_plan_set_profile(cpd_unet_plan* plan, int on);
int64_t cpd_unet_plan_profile_dump(cpd_unet_plan* plan, char* buf, int64_t cap);

/* ---- debug aids (not part of the product path; NULL / never called in production) ---------------------------------------- */
/* clock64 stamps of CTA 0's first 8 work items of every following cross-attention launch go to dev_buf (192 int64 on the
 * device); NULL switches it off (tools/attn5_timeline.py). */
void cpd_debug_attention_cross_timeline(long long* dev_buf);

#ifdef __cplusplus
}
#endif
#endif /* CPD_B200_H */
