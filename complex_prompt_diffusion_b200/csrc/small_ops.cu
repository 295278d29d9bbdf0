// Small CUDA-core kernels around the UNet GEMMs: timestep embedding, small-M linear (time-embedding MLP and
// the per-ResBlock emb projections), the 4->C input conv, the C->4 output conv and nearest 2x upsampling.
// None of them is GEMM-shaped enough for tensor cores (K = 36, N = 4 or M <= 32); all are memory bound.
#include "../../include/cpd_b200.h"
#include "common.cuh"

namespace {

// models/util.py:65-85: emb = cat[cos(t * f), sin(t * f)], f_k = exp(-ln(10000) * k / half), fp32.
__global__ void timestep_embedding_kernel(const float* __restrict__ t, int rows, int dim, int round_t, bf16* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  const int half = dim / 2;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * half) return;
  const int r = idx / half, k = idx - r * half;
  float tv = t[r];
  if (round_t) tv = __bfloat162float(__float2bfloat16_rn(tv));  // denoiser.py:393 casts t to the model dtype (P3)
  const float f = expf(-logf(10000.0f) * (float)k / (float)half);
  const float a = tv * f;
  out[(int64_t)r * dim + k] = __float2bfloat16_rn(cosf(a));
  out[(int64_t)r * dim + half + k] = __float2bfloat16_rn(sinf(a));
}

// out[m][n] = sum_k act(x[m][k]) * w[n][k] + b[n].  One warp per output column n, all m rows at once.
template <int MT>
__global__ void __launch_bounds__(256) small_linear_kernel(const bf16* __restrict__ x, int m, int k, const bf16* __restrict__ w,
                                                           const float* __restrict__ b, int n, int silu_in,
                                                           float* __restrict__ out_f32, bf16* __restrict__ out_bf16, int ld_out) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ uint8_t sh_raw[];
  bf16* xs = reinterpret_cast<bf16*>(sh_raw);  // [m][k] (activated)
  for (int i = threadIdx.x; i < m * k; i += blockDim.x) {
    float v = __bfloat162float(x[i]);
    if (silu_in) v = __bfloat162float(__float2bfloat16_rn(silu_f(v)));
    xs[i] = __float2bfloat16_rn(v);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int col = blockIdx.x * (blockDim.x >> 5) + warp;
  if (col >= n) return;
  float acc[MT];
#pragma unroll
  for (int i = 0; i < MT; ++i) acc[i] = 0.f;
  for (int kk = lane * 8; kk < k; kk += 256) {
    const uint4 wv = __ldg(reinterpret_cast<const uint4*>(w + (int64_t)col * k + kk));
    const uint32_t wu[4] = {wv.x, wv.y, wv.z, wv.w};
    float wf[8];
#pragma unroll
    for (int e = 0; e < 4; ++e) { const float2 f = unpack_bf16x2(wu[e]); wf[2 * e] = f.x; wf[2 * e + 1] = f.y; }
#pragma unroll
    for (int i = 0; i < MT; ++i) {
      if (i < m) {
        const uint4 xv = *reinterpret_cast<const uint4*>(xs + i * k + kk);
        const uint32_t xu[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = unpack_bf16x2(xu[e]);
          acc[i] += f.x * wf[2 * e] + f.y * wf[2 * e + 1];
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < MT; ++i) {
    float v = acc[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0 && i < m) {
      v += b ? b[col] : 0.f;
      v = __bfloat162float(__float2bfloat16_rn(v));  // the reference's Linear output is in the model dtype
      if (out_f32) out_f32[(int64_t)i * ld_out + col] = v;
      if (out_bf16) out_bf16[(int64_t)i * ld_out + col] = __float2bfloat16_rn(v);
    }
  }
}

// Input conv: x fp32 NCHW -> (x * scale) cast to the 16-bit model dtype -> 3x3 pad 1 -> NHWC 16-bit, written `rpi` times
// (one copy per conditioning row of the image: denoiser.py:390-391 broadcasts x_in over the rows).  One thread = one pixel
// of one SOURCE image x 8 output channels; weights are staged transposed in shared memory ([tap][cout] fp32) so a warp
// reads consecutive 32-byte groups, the 4 x 9 input taps are warp-broadcast loads.  `scale_ptr` (device, optional)
// overrides `scale` so that a captured CUDA graph can be replayed with a new c_in.
__global__ void __launch_bounds__(256) conv_in_kernel(const float* __restrict__ x, int n, int cin, int h, int w,
                                                      const bf16* __restrict__ wt, const float* __restrict__ bias, int cout,
                                                      float scale, const float* __restrict__ scale_ptr, int rpi, int f16,
                                                      bf16* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ uint8_t sh_raw[];
  float* ws = reinterpret_cast<float*>(sh_raw);  // [9 * cin][cout]
  const int taps = 9 * cin;
  for (int i = threadIdx.x; i < taps * cout; i += blockDim.x) {
    const int co = i / taps, tp = i - co * taps;  // wt is [cout][9][cin]
    ws[tp * cout + co] = __bfloat162float(wt[i]);
  }
  __syncthreads();
  if (scale_ptr) scale = __ldg(scale_ptr);
  const int cvecs = cout / 8;
  const int64_t total = (int64_t)n * h * w * cvecs;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int cv = (int)(idx % cvecs);
    const int64_t pix = idx / cvecs;
    const int xx = (int)(pix % w);
    const int yy = (int)((pix / w) % h);
    const int nn = (int)(pix / ((int64_t)w * h));
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = bias ? __ldg(bias + cv * 8 + e) : 0.f;
    for (int ky = 0; ky < 3; ++ky) {
      const int iy = yy + ky - 1;
      if (iy < 0 || iy >= h) continue;
      for (int kx = 0; kx < 3; ++kx) {
        const int ix = xx + kx - 1;
        if (ix < 0 || ix >= w) continue;
        for (int c = 0; c < cin; ++c) {
          const float xm = __fmul_rn(__ldg(x + (((int64_t)nn * cin + c) * h + iy) * w + ix), scale);
          const float xv = round_act(xm, f16 != 0);
          const float4 w0 = *reinterpret_cast<const float4*>(ws + ((ky * 3 + kx) * cin + c) * cout + cv * 8);
          const float4 w1 = *reinterpret_cast<const float4*>(ws + ((ky * 3 + kx) * cin + c) * cout + cv * 8 + 4);
          acc[0] += xv * w0.x; acc[1] += xv * w0.y; acc[2] += xv * w0.z; acc[3] += xv * w0.w;
          acc[4] += xv * w1.x; acc[5] += xv * w1.y; acc[6] += xv * w1.z; acc[7] += xv * w1.w;
        }
      }
    }
    const uint4 o = make_uint4(pack_act2(acc[0], acc[1], f16 != 0), pack_act2(acc[2], acc[3], f16 != 0),
                               pack_act2(acc[4], acc[5], f16 != 0), pack_act2(acc[6], acc[7], f16 != 0));
    const int64_t hw = (int64_t)h * w;
    const int64_t p_in = pix - (int64_t)nn * hw;
    for (int rr = 0; rr < rpi; ++rr)
      *reinterpret_cast<uint4*>(out + (((int64_t)nn * rpi + rr) * hw + p_in) * cout + cv * 8) = o;
  }
}

// Input conv, four pixels per thread (cin = 4, w % 4 == 0: the latent input of every UNet here).  The kernel above reads its
// weights from shared memory once per pixel (2 x LDS.128 per 8 multiply-adds: shared-memory-pipe bound, 43 us for 42 MB of
// output); here a thread owns FOUR horizontally adjacent pixels x 8 output channels, so every weight vector is read once
// per four pixels and every input value once per three taps.  Same accumulation order (ky, kx, c) as above: identical bits.
__global__ void __launch_bounds__(256) conv_in4_kernel(const float* __restrict__ x, int n, int h, int w,
                                                       const bf16* __restrict__ wt, const float* __restrict__ bias, int cout,
                                                       float scale, const float* __restrict__ scale_ptr, int rpi, int f16,
                                                       bf16* __restrict__ out) {
  constexpr int CIN = 4;
  pdl_launch_dependents();
  extern __shared__ uint8_t sh_raw[];
  float* ws = reinterpret_cast<float*>(sh_raw);  // [9 * CIN][cout]
  const int taps = 9 * CIN;
  for (int i = threadIdx.x; i < taps * cout; i += blockDim.x) {
    const int co = i / taps, tp = i - co * taps;  // wt is [cout][9][cin]
    ws[tp * cout + co] = __bfloat162float(wt[i]);
  }
  pdl_wait();
  __syncthreads();
  if (scale_ptr) scale = __ldg(scale_ptr);
  const int cvecs = cout / 8, wq = w / 4;
  const int total = n * h * wq * cvecs;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int cv = idx % cvecs;
    const int g = idx / cvecs;
    const int x0 = (g % wq) * 4;
    const int yy = (g / wq) % h;
    const int nn = g / (wq * h);
    float acc[4][8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float b = bias ? __ldg(bias + cv * 8 + e) : 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[q][e] = b;
    }
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int iy = yy + ky - 1;
      if (iy < 0 || iy >= h) continue;
      float xs[CIN][6];  // x0 - 1 .. x0 + 4 of this input row, scaled and rounded to the model dtype
      bool ok[6];
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        const int ix = x0 - 1 + j;
        ok[j] = ix >= 0 && ix < w;
#pragma unroll
        for (int c = 0; c < CIN; ++c)
          xs[c][j] = ok[j] ? round_act(__fmul_rn(__ldg(x + (((int64_t)nn * CIN + c) * h + iy) * w + ix), scale), f16 != 0) : 0.f;
      }
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
#pragma unroll
        for (int c = 0; c < CIN; ++c) {
          const float4 w0 = *reinterpret_cast<const float4*>(ws + ((ky * 3 + kx) * CIN + c) * cout + cv * 8);
          const float4 w1 = *reinterpret_cast<const float4*>(ws + ((ky * 3 + kx) * CIN + c) * cout + cv * 8 + 4);
          const float wf[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (!ok[q + kx]) continue;  // the per-pixel kernel skips taps outside the image (no + 0 * w: keeps -0 / NaN behaviour)
            const float xv = xs[c][q + kx];
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[q][e] += xv * wf[e];
          }
        }
      }
    }
    const int64_t hw = (int64_t)h * w;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint4 o = make_uint4(pack_act2(acc[q][0], acc[q][1], f16 != 0), pack_act2(acc[q][2], acc[q][3], f16 != 0),
                                 pack_act2(acc[q][4], acc[q][5], f16 != 0), pack_act2(acc[q][6], acc[q][7], f16 != 0));
      const int64_t p_in = (int64_t)yy * w + x0 + q;
      for (int rr = 0; rr < rpi; ++rr)
        *reinterpret_cast<uint4*>(out + (((int64_t)nn * rpi + rr) * hw + p_in) * cout + cv * 8) = o;
    }
  }
}

// Output conv: NHWC bf16 (cin) -> 3x3 pad 1 -> NCHW (cout <= 8).  One warp per output pixel.
template <int COUT>
__global__ void __launch_bounds__(256) conv_out_kernel(const bf16* __restrict__ a, int n, int h, int w, int cin,
                                                       const bf16* __restrict__ wt, const float* __restrict__ bias,
                                                       void* __restrict__ out, int out_dtype, int f16) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ uint8_t sh_raw[];
  bf16* ws = reinterpret_cast<bf16*>(sh_raw);  // [COUT][9][cin]
  for (int i = threadIdx.x * 8; i < COUT * 9 * cin; i += blockDim.x * 8)
    *reinterpret_cast<uint4*>(ws + i) = __ldg(reinterpret_cast<const uint4*>(wt + i));
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nvec = cin / 8;
  const int64_t total = (int64_t)n * h * w;
  // persistent blocks: the COUT x 9 x cin weights are staged once per block, then every warp walks over output pixels
  for (int64_t pix = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp; pix < total; pix += (int64_t)gridDim.x * (blockDim.x >> 5)) {
  const int xx = (int)(pix % w);
  const int yy = (int)((pix / w) % h);
  const int nn = (int)(pix / ((int64_t)w * h));
  float acc[COUT];
#pragma unroll
  for (int o = 0; o < COUT; ++o) acc[o] = 0.f;
  for (int v = lane; v < nvec; v += 32) {
    // all nine taps of this channel vector are loaded before the first use (nine 16-byte loads in flight per lane)
    uint4 av9[9];
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int iy = yy + tap / 3 - 1, ix = xx + tap % 3 - 1;
      av9[tap] = make_uint4(0u, 0u, 0u, 0u);  // zero padding
      if (iy >= 0 && iy < h && ix >= 0 && ix < w)
        av9[tap] = __ldg(reinterpret_cast<const uint4*>(a + (((int64_t)nn * h + iy) * w + ix) * cin + v * 8));
    }
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const uint32_t au[4] = {av9[tap].x, av9[tap].y, av9[tap].z, av9[tap].w};
      float af[8];
#pragma unroll
      for (int e = 0; e < 4; ++e) { const float2 f = unpack_act2(au[e], f16 != 0); af[2 * e] = f.x; af[2 * e + 1] = f.y; }
#pragma unroll
      for (int o = 0; o < COUT; ++o) {
        const uint4 wv = *reinterpret_cast<const uint4*>(ws + ((int64_t)o * 9 + tap) * cin + v * 8);
        const uint32_t wu[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = unpack_bf16x2(wu[e]);
          acc[o] += af[2 * e] * f.x + af[2 * e + 1] * f.y;
        }
      }
    }
  }
#pragma unroll
  for (int o = 0; o < COUT; ++o) {
    float v = acc[o];
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    if (lane == 0) {
      v += bias ? bias[o] : 0.f;
      const int64_t oi = (((int64_t)nn * COUT + o) * h + yy) * w + xx;
      if (out_dtype == CPD_BF16) reinterpret_cast<bf16*>(out)[oi] = __float2bfloat16_rn(v);
      else reinterpret_cast<float*>(out)[oi] = v;
    }
  }
  }
}

// Output conv, tiled (the hot form: cout = 4).  The one-warp-per-pixel kernel above re-reads its weights from shared
// memory for every pixel (4 wavefronts per 32 FMAs: shared-memory-pipe bound, 217 us at 16 x 64 x 64 x 320) and fetches
// every input pixel nine times through L1.  Here a block owns a strip of up to 64 output pixels of one image row:
//   * the 3 x (strip + 2) input pixels are staged once in shared memory, TRANSPOSED to [row][channel vector][pixel] so that
//     the 32 lanes of a warp (= 32 consecutive pixels) read 512 contiguous bytes per tap;
//   * the weights sit in shared memory as fp32 [tap][vector][cout][8] and are read as warp-wide BROADCASTS (one wavefront),
//     each thread reusing them for its two pixels (lane and lane + 32);
//   * the 8 warps split the channel vectors; their partial sums meet in shared memory and 256 threads write the
//     4 x 64 outputs (NCHW, coalesced).
// Persistent: the fp32 weights are staged once per block.
constexpr int CO_TP = 64;  // strip length
template <int NW>  // warps per block: they split the channel vectors
__global__ void __launch_bounds__(32 * NW, 1) conv_out_tiled_kernel(const bf16* __restrict__ a, int n, int h, int w, int cin,
                                                                 const bf16* __restrict__ wt, const float* __restrict__ bias,
                                                                 void* __restrict__ out, int out_dtype, int f16, int pp) {
  constexpr int COUT = 4;
  pdl_launch_dependents();
  extern __shared__ uint8_t sh_raw[];
  const int nvec = cin / 8;
  float* wsm = reinterpret_cast<float*>(sh_raw);                                   // [9][nvec][COUT][8] fp32
  float* red = wsm + 9 * nvec * COUT * 8;                                           // [NW warps][COUT][CO_TP]
  uint4* tile = reinterpret_cast<uint4*>(red + NW * COUT * CO_TP);                   // [3][nvec][pp] 16-byte channel vectors
  // weights: global [COUT][9][cin] bf16 -> shared [tap][vec][o][8] fp32 (independent of the previous kernel: before the wait)
  for (int i = threadIdx.x; i < COUT * 9 * nvec; i += blockDim.x) {
    const int v = i % nvec, tap = (i / nvec) % 9, o = i / (nvec * 9);
    const uint4 wv = __ldg(reinterpret_cast<const uint4*>(wt + ((int64_t)o * 9 + tap) * cin + v * 8));
    const float2 f0 = unpack_bf16x2(wv.x), f1 = unpack_bf16x2(wv.y), f2 = unpack_bf16x2(wv.z), f3 = unpack_bf16x2(wv.w);
    float4* dst = reinterpret_cast<float4*>(wsm + (((int64_t)tap * nvec + v) * COUT + o) * 8);
    dst[0] = make_float4(f0.x, f0.y, f1.x, f1.y);
    dst[1] = make_float4(f2.x, f2.y, f3.x, f3.y);
  }
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int strips = (w + CO_TP - 1) / CO_TP;
  const int64_t tiles = (int64_t)n * h * strips;
  for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
    const int sx = (int)(t % strips);
    const int yy = (int)((t / strips) % h);
    const int nn = (int)(t / ((int64_t)strips * h));
    const int x0 = sx * CO_TP;
    const int tp = min(CO_TP, w - x0);  // pixels of this strip
    __syncthreads();                    // the previous strip's tile / partial sums are no longer read
    // stage rows yy - 1 .. yy + 1, pixels x0 - 1 .. x0 + tp (zero outside the image) with 16-byte cp.async (zero-fill for the
    // padding): a warp takes a pixel, its lanes the channel vectors (coalesced global reads); the transposed destination has
    // a lane stride of pp * 16 bytes, pp odd: no bank conflicts.  All copies of a thread are in flight together (a plain
    // load -> store loop serialised ~30 L2 round trips per strip and cost more than the arithmetic).
    for (int pidx = warp; pidx < 3 * (tp + 2); pidx += NW) {
      const int r = pidx / (tp + 2), px = pidx - r * (tp + 2);
      const int iy = yy + r - 1, ix = x0 + px - 1;
      const bool inside = iy >= 0 && iy < h && ix >= 0 && ix < w;
      const bf16* src = inside ? a + (((int64_t)nn * h + iy) * w + ix) * cin : a;
      for (int v = lane; v < nvec; v += 32) {
        const uint32_t dst = smem_u32(tile + ((int64_t)r * nvec + v) * pp + px);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src + v * 8), "r"(inside ? 16 : 0) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    float acc[2][COUT];
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
      for (int o = 0; o < COUT; ++o) acc[q][o] = 0.f;
    const int px0 = (lane < tp) ? lane : 0;               // lanes past a short strip recompute pixel 0 (never written out)
    const int px1 = (lane + 32 < tp) ? lane + 32 : px0;   // second pixel of this lane (a duplicate when the strip is short)
    for (int v = warp; v < nvec; v += NW) {
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const int r = tap / 3, kx = tap % 3;
        const uint4* trow = tile + ((int64_t)r * nvec + v) * pp + kx;
        const uint4 a0 = trow[px0], a1 = trow[px1];
        float f0[8], f1[8];
        {
          const uint32_t u0[4] = {a0.x, a0.y, a0.z, a0.w}, u1[4] = {a1.x, a1.y, a1.z, a1.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 g0 = unpack_act2(u0[e], f16 != 0), g1 = unpack_act2(u1[e], f16 != 0);
            f0[2 * e] = g0.x; f0[2 * e + 1] = g0.y;
            f1[2 * e] = g1.x; f1[2 * e + 1] = g1.y;
          }
        }
        const float4* wrow = reinterpret_cast<const float4*>(wsm + ((int64_t)tap * nvec + v) * COUT * 8);
#pragma unroll
        for (int o = 0; o < COUT; ++o) {
          const float4 wa = wrow[2 * o], wb = wrow[2 * o + 1];  // same address in every lane: broadcast
          const float wf[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            acc[0][o] = fmaf(f0[e], wf[e], acc[0][o]);
            acc[1][o] = fmaf(f1[e], wf[e], acc[1][o]);
          }
        }
      }
    }
#pragma unroll
    for (int o = 0; o < COUT; ++o) {
      red[(warp * COUT + o) * CO_TP + lane] = acc[0][o];
      red[(warp * COUT + o) * CO_TP + lane + 32] = acc[1][o];
    }
    __syncthreads();
    if (threadIdx.x < COUT * CO_TP) {
      const int o = threadIdx.x / CO_TP, px = threadIdx.x % CO_TP;  // 256 threads = COUT x CO_TP outputs
      if (px < tp) {
        float v = bias ? __ldg(bias + o) : 0.f;
#pragma unroll
        for (int k = 0; k < NW; ++k) v += red[(k * COUT + o) * CO_TP + px];
        const int64_t oi = (((int64_t)nn * COUT + o) * h + yy) * w + x0 + px;
        if (out_dtype == CPD_BF16) reinterpret_cast<bf16*>(out)[oi] = __float2bfloat16_rn(v);
        else reinterpret_cast<float*>(out)[oi] = v;
      }
    }
  }
}

__global__ void __launch_bounds__(256) upsample2x_kernel(const bf16* __restrict__ a, int n, int h, int w, int c, bf16* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  // one thread per INPUT channel vector: read once, written to its 2 x 2 output pixels (a thread per output vector read
  // every input four times and spent more instructions on 64-bit index divisions than on the copy)
  const int cvecs = c / 8;
  const int64_t total = (int64_t)n * h * w * cvecs;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t pix = idx / cvecs;  // (nn * h + iy) * w + ix
    const int cv = (int)(idx - pix * cvecs);
    const int64_t row = pix / w;      // nn * h + iy
    const int ix = (int)(pix - row * w);
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(a + pix * c + cv * 8));
    bf16* o = out + ((2 * row) * (2 * w) + 2 * ix) * (int64_t)c + cv * 8;  // output rows 2 * (nn * h + iy), + 1
    *reinterpret_cast<uint4*>(o) = v;
    *reinterpret_cast<uint4*>(o + c) = v;
    *reinterpret_cast<uint4*>(o + (int64_t)2 * w * c) = v;
    *reinterpret_cast<uint4*>(o + (int64_t)2 * w * c + c) = v;
  }
}


// Row softmax over a 16-bit matrix (VAE mid-block attention, autoencoder.py:252-255): out[r][c] = softmax_c(scale * x[r][c]),
// fp32 arithmetic, one block per row, the row cached in registers (cols <= 256 threads x 8 vectors x 8 elements).
template <int MAXV>
__global__ void __launch_bounds__(256) softmax_rows_kernel(const bf16* __restrict__ x, int cols, int64_t ld, float scale_log2,
                                                           int f16, bf16* __restrict__ out, int64_t ldo) {
  __shared__ float red[8];
  pdl_launch_dependents();
  pdl_wait();
  const bf16* xr = x + (int64_t)blockIdx.x * ld;
  bf16* orow = out + (int64_t)blockIdx.x * ldo;
  const int nvec = cols / 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint4 raw[MAXV];
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int v = threadIdx.x + i * 256;
    raw[i] = make_uint4(0u, 0u, 0u, 0u);
    if (v < nvec) {
      raw[i] = __ldg(reinterpret_cast<const uint4*>(xr + v * 8));
      const uint32_t u[4] = {raw[i].x, raw[i].y, raw[i].z, raw[i].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = unpack_act2(u[e], f16 != 0);
        mx = fmaxf(mx, fmaxf(f.x, f.y));
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  mx = red[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) mx = fmaxf(mx, red[i]);
  __syncthreads();
  // scale > 0: max(scale * x) = scale * max(x); p = 2^((x - max) * scale * log2 e)
  float sum = 0.f;
  float p[MAXV][8];
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int v = threadIdx.x + i * 256;
    const uint32_t u[4] = {raw[i].x, raw[i].y, raw[i].z, raw[i].w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f = unpack_act2(u[e], f16 != 0);
      p[i][2 * e] = v < nvec ? exp2f((f.x - mx) * scale_log2) : 0.f;
      p[i][2 * e + 1] = v < nvec ? exp2f((f.y - mx) * scale_log2) : 0.f;
      sum += p[i][2 * e] + p[i][2 * e + 1];
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  sum = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) sum += red[i];  // fixed order: bit-reproducible
  const float inv = 1.0f / sum;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int v = threadIdx.x + i * 256;
    if (v < nvec)
      *reinterpret_cast<uint4*>(orow + v * 8) =
          make_uint4(pack_act2(p[i][0] * inv, p[i][1] * inv, f16 != 0), pack_act2(p[i][2] * inv, p[i][3] * inv, f16 != 0),
                     pack_act2(p[i][4] * inv, p[i][5] * inv, f16 != 0), pack_act2(p[i][6] * inv, p[i][7] * inv, f16 != 0));
  }
}

// 1x1 conv over an fp32 NCHW tensor with a handful of channels (post_quant_conv, autoencoder.py:800,826):
// out[n][o][p] = b[o] + sum_c w[o][c] * (x[n][c][p] * scale); cin, cout <= 8.
__global__ void __launch_bounds__(256) pointwise_small_kernel(const float* __restrict__ x, int n, int cin, int cout, int64_t hw,
                                                              const float* __restrict__ w, const float* __restrict__ b, float scale,
                                                              float* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t total = (int64_t)n * hw;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t nn = idx / hw, p = idx - nn * hw;
    float xv[8];
    for (int c = 0; c < cin; ++c) xv[c] = __fmul_rn(x[(nn * cin + c) * hw + p], scale);
    for (int o = 0; o < cout; ++o) {
      float acc = b ? b[o] : 0.f;
      for (int c = 0; c < cin; ++c) acc = fmaf(w[o * cin + c], xv[c], acc);
      out[(nn * cout + o) * hw + p] = acc;
    }
  }
}

}  // namespace

extern "C" cpd_status cpd_timestep_embedding(const float* t, int rows, int dim, int round_t_bf16, void* out, void* stream) {
  CPD_REQUIRE(t && out && rows > 0 && dim > 0 && dim % 2 == 0, "cpd_timestep_embedding: bad arguments (rows=%d dim=%d)", rows, dim);
  const int total = rows * (dim / 2);
  CPD_CUDA_CHECK(cpd_launch(timestep_embedding_kernel, dim3((total + 255) / 256), dim3(256), 0, (cudaStream_t)stream, t, rows, dim, round_t_bf16, (bf16*)out));
  CPD_CUDA_CHECK(cudaGetLastError());
  return CPD_OK;
}

extern "C" cpd_status cpd_small_linear(const void* x, int m, int k, const void* w, const float* b, int n, int silu_in,
                                       float* out_f32, void* out_bf16, int ld_out, void* stream) {
  CPD_REQUIRE(x && w && (out_f32 || out_bf16), "cpd_small_linear: null pointer");
  CPD_REQUIRE(m >= 1 && m <= 32, "cpd_small_linear: m=%d must be in [1,32]", m);
  CPD_REQUIRE(k > 0 && k % 8 == 0 && (size_t)m * k * 2 <= 160 * 1024, "cpd_small_linear: k=%d unsupported for m=%d", k, m);
  CPD_REQUIRE(n > 0 && ld_out >= n, "cpd_small_linear: n=%d ld_out=%d", n, ld_out);
  cudaStream_t s = (cudaStream_t)stream;
  const size_t shm = (size_t)m * k * 2;
  const int blocks = (n + 7) / 8;
#define LAUNCH_SL(MT)                                                                                                         \
  do {                                                                                                                        \
    CPD_SMEM_OPTIN(small_linear_kernel<MT>, 160 * 1024);                                                                      \
    CPD_CUDA_CHECK(cpd_launch(small_linear_kernel<MT>, dim3(blocks), dim3(256), shm, s, (const bf16*)x, m, k, (const bf16*)w, b, n, silu_in, out_f32,           \
                                                      (bf16*)out_bf16, ld_out));                                              \
  } while (0)
  if (m <= 4) LAUNCH_SL(4);
  else if (m <= 16) LAUNCH_SL(16);
  else LAUNCH_SL(32);
#undef LAUNCH_SL
  CPD_CUDA_CHECK(cudaGetLastError());
  return CPD_OK;
}

extern "C" cpd_status cpd_conv_in(const float* x, int n, int cin, int h, int w, const void* wt, const float* bias, int cout,
                                  float scale, const float* scale_ptr, int rows_per_image, int act_fp16, void* out, void* stream) {
  CPD_REQUIRE(x && wt && out, "cpd_conv_in: null pointer");
  CPD_REQUIRE(n > 0 && cin > 0 && cin <= 8 && h > 0 && w > 0 && cout % 8 == 0, "cpd_conv_in: bad shape n=%d cin=%d h=%d w=%d cout=%d", n, cin, h, w, cout);
  CPD_REQUIRE(rows_per_image >= 1, "cpd_conv_in: rows_per_image=%d", rows_per_image);
  const int64_t total = (int64_t)n * h * w * (cout / 8);
  const size_t shm = (size_t)9 * cin * cout * sizeof(float);
  CPD_REQUIRE(shm <= 100 * 1024, "cpd_conv_in: cin=%d x cout=%d weights do not fit 100 KB of shared memory", cin, cout);
  CPD_SMEM_OPTIN(conv_in_kernel, 100 * 1024);
  if (cin == 4 && w % 4 == 0 && (int64_t)n * h * w * (cout / 8) < (int64_t)1 << 31) {
    CPD_SMEM_OPTIN(conv_in4_kernel, 100 * 1024);
    int64_t blocks4 = (total / 4 + 255) / 256;
    if (blocks4 > 148 * 2) blocks4 = 148 * 2;
    CPD_CUDA_CHECK(cpd_launch(conv_in4_kernel, dim3((unsigned)blocks4), dim3(256), shm, (cudaStream_t)stream, x, n, h, w, (const bf16*)wt, bias, cout, scale,
                              scale_ptr, rows_per_image, act_fp16, (bf16*)out));
    CPD_CUDA_CHECK(cudaGetLastError());
    return CPD_OK;
  }
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 2) blocks = 148 * 2;  // each block stages the 9 x cin x cout weights once: few, long-lived blocks
  CPD_CUDA_CHECK(cpd_launch(conv_in_kernel, dim3((unsigned)blocks), dim3(256), shm, (cudaStream_t)stream, x, n, cin, h, w, (const bf16*)wt, bias, cout, scale, scale_ptr,
                                                                      rows_per_image, act_fp16, (bf16*)out));
  CPD_CUDA_CHECK(cudaGetLastError());
  return CPD_OK;
}

extern "C" cpd_status cpd_conv_out(const void* a, int n, int h, int w, int cin, const void* wt, const float* bias, int cout,
                                   void* out, int out_dtype, int act_fp16, void* stream) {
  CPD_REQUIRE(a && wt && out, "cpd_conv_out: null pointer");
  CPD_REQUIRE(cout == 4 || cout == 8, "cpd_conv_out: cout=%d unsupported (4 or 8)", cout);
  CPD_REQUIRE(cin % 8 == 0 && (size_t)cout * 9 * cin * 2 <= 200 * 1024, "cpd_conv_out: cin=%d unsupported", cin);
  CPD_REQUIRE(out_dtype == CPD_BF16 || out_dtype == CPD_F32, "cpd_conv_out: out_dtype=%d", out_dtype);
  const int64_t pixels = (int64_t)n * h * w;
  const size_t shm = (size_t)cout * 9 * cin * 2;
  cudaStream_t s = (cudaStream_t)stream;
  if (cout == 4) {  // tiled kernel when its strip (3 rows x 66 pixels x cin) and the fp32 weights fit in shared memory
    const int tp = w < CO_TP ? w : CO_TP;
    const int pp = (tp + 2) | 1;  // odd pixel pitch: conflict-free transposed stores
    static int tiled = -1, nw = 16;
    if (tiled < 0) {
      const char* e = getenv("CPD_CONV_OUT_TILED");
      tiled = (e && e[0] == '0') ? 0 : 1;
      const char* e2 = getenv("CPD_CONV_OUT_WARPS");
      if (e2 && atoi(e2) == 8) nw = 8;
    }
    const size_t need = (size_t)9 * (cin / 8) * 4 * 8 * 4 + (size_t)nw * 4 * CO_TP * 4 + (size_t)3 * (cin / 8) * pp * 16;
    if (tiled && need <= 227 * 1024) {
      CPD_SMEM_OPTIN(conv_out_tiled_kernel<8>, 227 * 1024);
      CPD_SMEM_OPTIN(conv_out_tiled_kernel<16>, 227 * 1024);
      const int64_t tiles = (int64_t)n * h * ((w + CO_TP - 1) / CO_TP);
      const unsigned blocks = (unsigned)(tiles < 148 ? tiles : 148);
      if (nw == 8)
        CPD_CUDA_CHECK(cpd_launch(conv_out_tiled_kernel<8>, dim3(blocks), dim3(256), need, s, (const bf16*)a, n, h, w, cin, (const bf16*)wt, bias,
                                  out, out_dtype, act_fp16, pp));
      else
        CPD_CUDA_CHECK(cpd_launch(conv_out_tiled_kernel<16>, dim3(blocks), dim3(512), need, s, (const bf16*)a, n, h, w, cin, (const bf16*)wt, bias,
                                  out, out_dtype, act_fp16, pp));
      CPD_CUDA_CHECK(cudaGetLastError());
      return CPD_OK;
    }
  }
  unsigned blocks = (unsigned)((pixels + 7) / 8);
  if (blocks > 148 * 8) blocks = 148 * 8;  // persistent: weights staged once per block, full occupancy (64 warps per SM)
  if (cout == 4) {
    CPD_SMEM_OPTIN(conv_out_kernel<4>, 200 * 1024);
    CPD_CUDA_CHECK(cpd_launch(conv_out_kernel<4>, dim3(blocks), dim3(256), shm, s, (const bf16*)a, n, h, w, cin, (const bf16*)wt, bias, out, out_dtype, act_fp16));
  } else {
    CPD_SMEM_OPTIN(conv_out_kernel<8>, 200 * 1024);
    CPD_CUDA_CHECK(cpd_launch(conv_out_kernel<8>, dim3(blocks), dim3(256), shm, s, (const bf16*)a, n, h, w, cin, (const bf16*)wt, bias, out, out_dtype, act_fp16));
  }
  CPD_CUDA_CHECK(cudaGetLastError());
  return CPD_OK;
}

extern "C" cpd_status cpd_upsample2x(const void* a, int n, int h, int w, int c, void* out, void* stream) {
  CPD_REQUIRE(a && out && n > 0 && h > 0 && w > 0 && c % 8 == 0, "cpd_upsample2x: bad arguments");
  const int64_t total = (int64_t)n * h * w * (c / 8);
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  CPD_CUDA_CHECK(cpd_launch(upsample2x_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, (const bf16*)a, n, h, w, c, (bf16*)out));
  CPD_CUDA_CHECK(cudaGetLastError());
  return CPD_OK;
}

extern "C" cpd_status cpd_softmax_rows(const void* x, int rows, int cols, int64_t ld, float scale, int act_fp16, void* out, int64_t ldo,
                                       void* stream) {
  CPD_REQUIRE(x && out, "cpd_softmax_rows: null pointer");
  CPD_REQUIRE(rows >= 0 && cols > 0 && cols % 8 == 0 && cols <= 256 * 8 * 8, "cpd_softmax_rows: cols=%d must be a multiple of 8, <= 16384", cols);
  CPD_REQUIRE(ld % 8 == 0 && ldo % 8 == 0 && ld >= cols && ldo >= cols, "cpd_softmax_rows: bad leading dimensions");
  CPD_REQUIRE(scale > 0.f, "cpd_softmax_rows: scale must be positive");
  if (rows == 0) return CPD_OK;
  const float sl2 = scale * 1.4426950408889634f;
  cudaStream_t s = (cudaStream_t)stream;
  if (cols <= 256 * 8 * 2)
    CPD_CUDA_CHECK(cpd_launch(softmax_rows_kernel<2>, dim3(rows), dim3(256), 0, s, (const bf16*)x, cols, ld, sl2, act_fp16, (bf16*)out, ldo));
  else
    CPD_CUDA_CHECK(cpd_launch(softmax_rows_kernel<8>, dim3(rows), dim3(256), 0, s, (const bf16*)x, cols, ld, sl2, act_fp16, (bf16*)out, ldo));
  CPD_CUDA_CHECK(cudaGetLastError());
  return CPD_OK;
}

extern "C" cpd_status cpd_pointwise_small(const float* x, int n, int cin, int cout, int64_t hw, const float* w, const float* b, float scale,
                                          float* out, void* stream) {
  CPD_REQUIRE(x && w && out, "cpd_pointwise_small: null pointer");
  CPD_REQUIRE(n >= 0 && cin >= 1 && cin <= 8 && cout >= 1 && cout <= 8 && hw > 0, "cpd_pointwise_small: n=%d cin=%d cout=%d", n, cin, cout);
  if (n == 0) return CPD_OK;
  int64_t blocks = ((int64_t)n * hw + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  CPD_CUDA_CHECK(cpd_launch(pointwise_small_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, x, n, cin, cout, hw, w, b, scale, out));
  CPD_CUDA_CHECK(cudaGetLastError());
  return CPD_OK;
}

// ---- latents -> images tail (cpd/embeddings/prompts.py:472-475): imgs = clamp((x + 1) / 2, 0, 1).permute(0, 2, 3, 1).mul(255).to(uint8)
__global__ void __launch_bounds__(256) images_to_uint8_kernel(const float* __restrict__ x, int n, int c, int64_t hw, int ld_c,
                                                              uint8_t* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t total = (int64_t)n * hw;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / hw, p = i - b * hw;
    for (int ch = 0; ch < c; ++ch) {
      float v = __fdiv_rn(__fadd_rn(x[(b * ld_c + ch) * hw + p], 1.0f), 2.0f);
      v = fminf(fmaxf(v, 0.0f), 1.0f);
      out[i * c + ch] = (uint8_t)__fmul_rn(v, 255.0f);  // float -> uint8 truncates like torch's .to(torch.uint8)
    }
  }
}

extern "C" cpd_status cpd_images_to_uint8(const float* x, int n, int c, int64_t hw, int ld_c, uint8_t* out, void* stream) {
  CPD_REQUIRE(x && out, "cpd_images_to_uint8: null pointer");
  CPD_REQUIRE(n >= 0 && c >= 1 && c <= 8 && hw > 0 && ld_c >= c, "cpd_images_to_uint8: n=%d c=%d ld_c=%d", n, c, ld_c);
  if (n == 0) return CPD_OK;
  int64_t blocks = ((int64_t)n * hw + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  CPD_CUDA_CHECK(cpd_launch(images_to_uint8_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, x, n, c, hw, ld_c, out));
  CPD_CUDA_CHECK(cudaGetLastError());
  return CPD_OK;
}
