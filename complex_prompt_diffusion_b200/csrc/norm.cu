// GroupNorm(32)(+SiLU) and LayerNorm over NHWC bf16 activations: HBM-bound, 16-byte vectorised,
// fp32 statistics (models/util.py:95-105; attention.py:89-90,476-478 of the reference).
#include <stdlib.h>

#include "../../include/cpd_b200.h"
#include "common.cuh"

namespace {

constexpr int GROUPS = 32;
// SiLU with one MUFU op: x * sigmoid(x) = h + h * tanh(h), h = x / 2 (tanh.approx.f32: |error| ~ 2^-11, the size of the
// 16-bit output rounding; exp + rcp would keep the MUFU pipe busy twice as long and bound the apply pass).
__device__ __forceinline__ float silu_tanh_f(float x) {
  const float h = 0.5f * x;
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}
constexpr int RING = 8;  // per-thread cp.async ring depth of the GroupNorm kernels (16-byte slots)

// Per-thread asynchronous 16-byte copies global -> shared (LDGSTS): RING loads in flight per thread at no register cost.  A
// thread only ever reads the slots it filled itself, so no barrier is involved - cp.async.wait_group orders the read.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---- GroupNorm pass 1: per-(image, pixel chunk, group) sum / sum of squares ----------------------------------------
// grid = (chunks, n_img).  Thread t owns channel vector (t % vec_per_px) and pixel lane (t / vec_per_px).  No atomics
// anywhere: per-thread partials go to shared memory, are summed over the pixel lanes in a fixed order, reduced per group in
// fp64 and written to partial[n][chunk][group][2]; pass 2 sums the chunks in a fixed order.  Results are bit-reproducible.
template <bool F16>
__global__ void __launch_bounds__(512, 2) gn_stats_kernel(const bf16* __restrict__ a0, const bf16* __restrict__ a1, int c0,
                                                          int c1, int hw, int px_per_block, double* __restrict__ partial) {
  extern __shared__ __align__(16) float sh_raw[];  // ring[RING][blockDim.x] of uint4, then partial sums [lanes][2][C]
  uint4* ring = reinterpret_cast<uint4*>(sh_raw) + threadIdx.x;
  float* sh = sh_raw + (size_t)RING * blockDim.x * 4;
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
    const bf16* base = src + ((int64_t)n * hw + p0 + pl) * cs + coff;
    const int64_t step = (int64_t)lanes * cs;
    const int np = (p1 - p0 - pl + lanes - 1) / lanes;  // pixels of this thread
#pragma unroll
    for (int j = 0; j < RING; ++j) {
      if (j < np) cp_async16(ring + j * blockDim.x, base + j * step);
      cp_async_commit();
    }
    for (int j0 = 0; j0 < np; j0 += RING) {
#pragma unroll
      for (int jj = 0; jj < RING; ++jj) {
        const int j = j0 + jj;
        cp_async_wait<RING - 1>();
        if (j < np) {
          const uint4 v = ring[jj * blockDim.x];
          if (j + RING < np) cp_async16(ring + jj * blockDim.x, base + (j + RING) * step);
          const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 f = unpack_act2(u[e], F16);
            s[2 * e] += f.x; ss[2 * e] += f.x * f.x;
            s[2 * e + 1] += f.y; ss[2 * e + 1] += f.y * f.y;
          }
        }
        cp_async_commit();
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
  // Block reduction in a fixed order: (1) fp32 sum over the pixel lanes per (statistic, channel), (2) fp64 sum over the
  // channels of a group by one thread per (group, statistic).
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
    float acc = sh[i];
    for (int l = 1; l < lanes; ++l) acc += sh[(size_t)l * 2 * C + i];
    sh[i] = acc;
  }
  __syncthreads();
  if (threadIdx.x < 2 * GROUPS) {
    const int cpg = C / GROUPS;
    const int g = threadIdx.x >> 1, st = threadIdx.x & 1;
    const float* v = sh + st * C + g * cpg;
    double acc = 0.0;
    for (int c = 0; c < cpg; ++c) acc += (double)v[c];
    partial[(((int64_t)n * gridDim.x + blockIdx.x) * GROUPS + g) * 2 + st] = acc;
  }
}

// ---- GroupNorm pass 2: normalise, affine, optional SiLU, write bf16 ---------------------------------
template <bool F16>
__global__ void __launch_bounds__(512, 2) gn_apply_kernel(const bf16* __restrict__ a0, const bf16* __restrict__ a1, int c0,
                                                       int c1, int hw, int px_per_block, const double* __restrict__ partial,
                                                       const long long* __restrict__ chan_sums,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       float eps, int silu, bf16* __restrict__ out) {
  extern __shared__ __align__(16) float sh_raw[];  // ring[RING][blockDim.x] of uint4, scale[C], shift[C], group stats
  uint4* ring = reinterpret_cast<uint4*>(sh_raw) + threadIdx.x;
  float* sh = sh_raw + (size_t)RING * blockDim.x * 4;
  pdl_launch_dependents();
  pdl_wait();
  const int C = c0 + c1;
  const int vec_per_px = C / 8;
  const int lanes = blockDim.x / vec_per_px;
  const int n = blockIdx.y;
  // The data loads do not depend on the statistics: fill the ring first, reduce the chunk partials while they fly.
  const int cv = threadIdx.x % vec_per_px;
  const int pl = threadIdx.x / vec_per_px;
  const int ch = cv * 8;
  const bf16* src;
  int cs, coff;
  if (ch < c0) { src = a0; cs = c0; coff = ch; } else { src = a1; cs = c1; coff = ch - c0; }
  const int p0 = blockIdx.x * px_per_block;
  const int p1 = min(p0 + px_per_block, hw);
  const bf16* base = src + ((int64_t)n * hw + p0 + pl) * cs + coff;
  const int64_t step = (int64_t)lanes * cs;
  const int np = pl < lanes ? (p1 - p0 - pl + lanes - 1) / lanes : 0;  // pixels of this thread
#pragma unroll
  for (int j = 0; j < RING; ++j) {
    if (j < np) cp_async16(ring + j * blockDim.x, base + j * step);
    cp_async_commit();
  }
  const int cpg = C / GROUPS;
  const double cnt = (double)hw * cpg;
  {  // chunk partials of pass 1: one warp per group at a time, lanes over chunks, fixed butterfly (bit-reproducible)
    float* gstat = sh + 2 * C;  // [GROUPS][2]: mean, rstd
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int g = warp; g < GROUPS; g += nwarps) {
      double S = 0.0, SS = 0.0;
      if (chan_sums) {
        // statistics emitted by the producing GEMM (cpd_gemm_params.gn_sums_out): exact fixed-point sums per (image, channel)
        long long s1 = 0, s2 = 0;
        for (int c = lane; c < cpg; c += 32) {
          const longlong2 pp = *reinterpret_cast<const longlong2*>(chan_sums + ((int64_t)n * C + g * cpg + c) * 2);
          s1 += pp.x;
          s2 += pp.y;
        }
        S = (double)s1 * (1.0 / 16777216.0);
        SS = (double)s2 * (1.0 / 4096.0);
      } else {
        for (int k = lane; k < (int)gridDim.x; k += 32) {
          const double2 pp = *reinterpret_cast<const double2*>(partial + (((int64_t)n * gridDim.x + k) * GROUPS + g) * 2);
          S += pp.x;
          SS += pp.y;
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        S += __shfl_xor_sync(0xffffffffu, S, o);
        SS += __shfl_xor_sync(0xffffffffu, SS, o);
      }
      if (lane == 0) {
        const double mean = S / cnt;
        double var = SS / cnt - mean * mean;
        if (var < 0.0) var = 0.0;
        gstat[2 * g] = (float)mean;
        gstat[2 * g + 1] = (float)(1.0 / sqrt(var + (double)eps));
      }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      const int g = (int)__fdividef((float)c + 0.5f, (float)cpg);  // exact for c < 2^20
      const float sc = gstat[2 * g + 1] * gamma[c];
      sh[c] = sc;
      sh[C + c] = beta[c] - gstat[2 * g] * sc;
    }
  }
  __syncthreads();
  if (pl >= lanes) return;
  float sc[8], sf[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { sc[e] = sh[ch + e]; sf[e] = sh[C + ch + e]; }
  bf16* obase = out + ((int64_t)n * hw + p0 + pl) * C + ch;
  const int64_t ostep = (int64_t)lanes * C;
  for (int j0 = 0; j0 < np; j0 += RING) {
#pragma unroll
    for (int jj = 0; jj < RING; ++jj) {
      const int j = j0 + jj;
      cp_async_wait<RING - 1>();
      if (j < np) {
        const uint4 v = ring[jj * blockDim.x];
        if (j + RING < np) cp_async16(ring + jj * blockDim.x, base + (j + RING) * step);
        const uint32_t u[4] = {v.x, v.y, v.z, v.w};
        uint32_t o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = unpack_act2(u[e], F16);
          float y0 = f.x * sc[2 * e] + sf[2 * e];
          float y1 = f.y * sc[2 * e + 1] + sf[2 * e + 1];
          if (silu) { y0 = silu_tanh_f(y0); y1 = silu_tanh_f(y1); }
          o[e] = pack_act2(y0, y1, F16);
        }
        *reinterpret_cast<uint4*>(obase + j * ostep) = make_uint4(o[0], o[1], o[2], o[3]);
      }
      cp_async_commit();
    }
  }
}

// ---- GroupNorm, single pass with the slab in SHARED memory (round 2, end): one block owns a pixel range of (image, slab of whole
// groups); the slab's pixels are split over a cluster of CL blocks when they do not fit one block.  Every thread fetches its
// 16-byte vectors with cp.async into slots only it reads (no register cost: any number in flight, two to four blocks per SM, so
// one block's loads overlap another's arithmetic and stores - an SM ingests ~46 B / clk, 1/148 of what the L2 delivers, so every
// SM has to be loading all the time).  The first version of this kernel executed 20 instructions per element (ncu: issue slots
// 54 % busy, DRAM 19 %), 43 % of them outside the two passes over the data - so this one is built around the instruction count:
// * grid = (cluster rank, slab, image) and multiply-shift divisions by the host's magic numbers: no integer division;
// * packed fp32 arithmetic (FADD2 / FFMA2, sm_100) in both passes;
// * a vector of 8 channels lies in at most two groups (an even number of channels per group), so a thread ends with four sums
//   (S, SS of its first and of its second group); lanes that hold the same channel vector are folded with shuffles, one lane per
//   (vector, warp) writes them to shared memory and 2 x (groups of the slab) threads add them up in a fixed order in fp64;
// * cluster blocks PUSH their per-group partials into every block of the cluster (distributed shared memory stores), meet at ONE
//   cluster barrier and add the ranks' partials locally in rank order: every block gets the same bits, nothing depends on timing
//   (bit-reproducible), and nobody reads a peer's memory after the barrier, so blocks leave independently;
// * the loads are two cp.async groups: the sums of the first half run while the second half lands.
// Slabs are 32-byte-sector aligned where the channel count allows (80 channels at 10 / 20 / 40 channels per group), so no sector
// is fetched by two blocks.
template <bool F16, bool SILU>
__global__ void __launch_bounds__(256, 4)
    gn_slab_kernel(const bf16* __restrict__ a0, const bf16* __restrict__ a1, int c0, int c1, int hw, int slab_c, int lanes, int ppc,
                   uint32_t magic_vps, uint32_t magic_cpg, int hdr, int half, const float* __restrict__ gamma,
                   const float* __restrict__ beta, float eps, bf16* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char slab_raw[];
  double* part = reinterpret_cast<double*>(slab_raw);        // [cluster rank][groups of the slab][2]: every block's (S, SS), pushed by its owner
  float* gstat = reinterpret_cast<float*>(slab_raw + 1024);  // [groups of the slab][2]: mean, rstd
  float4* red = reinterpret_cast<float4*>(slab_raw + 1088);  // [vector of the slab][warp]: (S, SS) of its two groups
  pdl_launch_dependents();
  const int tid = threadIdx.x, T = blockDim.x;
  const int C = c0 + c1;
  const int cpg = C >> 5;
  const int vps = slab_c >> 3;
  const int gslab = (int)(((uint32_t)slab_c * magic_cpg) >> 16);
  const int cl = gridDim.x, rank = blockIdx.x;  // the cluster spans the grid's x dimension
  // "I am running": a peer may store into this block's shared memory only once the block executes; the matching wait sits in front
  // of the first such store, by which time every block has long arrived (split arrive / wait: nobody stalls here)
  if (cl > 1) asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
  const int ch0 = blockIdx.y * slab_c;
  const int n = blockIdx.z;
  const int p_begin = rank * ppc, p_end = min(hw, p_begin + ppc);
  const int pl = (int)(((uint32_t)tid * magic_vps) >> 16);  // tid / vps
  const int cv = tid - pl * vps;
  const int ch = ch0 + cv * 8;
  // pairs [0, split2) of the vector's four channel pairs belong to group gA of the slab, the rest to gA + 1
  const int gA = (int)(((uint32_t)(cv * 8) * magic_cpg) >> 16);
  const int split2 = min(4, ((gA + 1) * cpg - cv * 8) >> 1);
  // the affine parameters are weights: not written by the previous kernel, fetched in front of the dependency wait
  const float4 gm0 = __ldg(reinterpret_cast<const float4*>(gamma + ch)), gm1 = __ldg(reinterpret_cast<const float4*>(gamma + ch + 4));
  const float4 bt0 = __ldg(reinterpret_cast<const float4*>(beta + ch)), bt1 = __ldg(reinterpret_cast<const float4*>(beta + ch + 4));
  const bf16* src;
  int cs, coff;
  if (ch < c0) { src = a0; cs = c0; coff = ch; } else { src = a1; cs = c1; coff = ch - c0; }
  const uint32_t slot0 = smem_u32(slab_raw + hdr) + tid * 16, sstep = T * 16;
  int np = 0;  // pixels of this thread
  pdl_wait();
  int np1 = 0;  // ... of which in the first of the two cp.async groups (the sums of the first half run while the second half lands)
  if (pl < lanes) {
    const char* gp = reinterpret_cast<const char*>(src + ((int64_t)n * hw + p_begin + pl) * cs + coff);
    const int64_t gstep = (int64_t)lanes * cs * 2;
    uint32_t sa = slot0;
    int p = p_begin + pl;
    for (; p < p_end && np < half; p += lanes, ++np, sa += sstep, gp += gstep)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gp) : "memory");
    np1 = np;
    cp_async_commit();
    for (; p < p_end; p += lanes, ++np, sa += sstep, gp += gstep)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gp) : "memory");
  } else {
    cp_async_commit();
  }
  cp_async_commit();
  float2 s[4], q[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) s[e] = q[e] = make_float2(0.f, 0.f);
  auto accumulate = [&](int j) {
    const uint4 v = lds128(slot0 + j * sstep);
    const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f = unpack_act2(u[e], F16);
      s[e] = __fadd2_rn(s[e], f);
      q[e] = __ffma2_rn(f, f, q[e]);
    }
  };
  cp_async_wait<1>();
#pragma unroll 4
  for (int j = 0; j < np1; ++j) accumulate(j);
  cp_async_wait<0>();
#pragma unroll 4
  for (int j = np1; j < np; ++j) accumulate(j);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);  // (S, SS) of group gA, (S, SS) of group gA + 1
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float a = s[e].x + s[e].y, b = q[e].x + q[e].y;
    if (e < split2) { acc.x += a; acc.y += b; } else { acc.z += a; acc.w += b; }
  }
  const int warp = tid >> 5, lane = tid & 31, nwarps = T >> 5;
  const float4 own = acc;  // lanes l, l + vps, l + 2 vps, ... of a warp hold the same channel vector: lane l < vps adds them up
  for (int k = vps; k < 32; k += vps) {
    const float x = __shfl_down_sync(0xffffffffu, own.x, k), y = __shfl_down_sync(0xffffffffu, own.y, k);
    const float z = __shfl_down_sync(0xffffffffu, own.z, k), w = __shfl_down_sync(0xffffffffu, own.w, k);
    if (lane + k < 32) { acc.x += x; acc.y += y; acc.z += z; acc.w += w; }
  }
  if (lane < vps) red[cv * nwarps + warp] = acc;
  __syncthreads();
  if (cl > 1) asm volatile("barrier.cluster.wait.aligned;" ::: "memory");  // every block of the cluster has started
  if (tid < 2 * gslab) {  // fixed order over the vectors of the group and the warps
    const int g = tid >> 1, st = tid & 1;
    const int cv_lo = (g * cpg) >> 3, cv_hi = ((g + 1) * cpg - 1) >> 3;
    const float* redf = reinterpret_cast<const float*>(red);
    double a = 0.0;
    for (int c = cv_lo; c <= cv_hi; ++c) {
      const int gc = (int)(((uint32_t)(c * 8) * magic_cpg) >> 16);
      const float* col = redf + (size_t)c * nwarps * 4 + (gc == g ? st : 2 + st);
      for (int w = 0; w < nwarps; ++w) a += (double)col[w * 4];
    }
    // PUSH the partial into every block of the cluster (distributed shared memory stores, fire and forget): after ONE cluster
    // barrier each block holds all partials and reduces them locally - nobody reads a peer's memory afterwards, so no second
    // cluster barrier has to keep the blocks alive for each other.
    if (cl > 1) {
      const uint32_t laddr = smem_u32(part + rank * 16 + tid);
      for (int r = 0; r < cl; ++r) {
        uint32_t raddr;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(laddr), "r"(r));
        asm volatile("st.shared::cluster.b64 [%0], %1;" ::"r"(raddr), "l"(__double_as_longlong(a)) : "memory");
      }
    } else {
      part[tid] = a;
    }
  }
  if (cl > 1) cluster_sync_all(); else __syncthreads();  // every block's partials have landed here (release / acquire at cluster scope)
  if (tid < gslab) {
    const int g = tid;
    double S = 0.0, SS = 0.0;
    for (int r = 0; r < cl; ++r) {  // fixed order over the cluster ranks: the same bits in every block
      S += part[r * 16 + 2 * g];
      SS += part[r * 16 + 2 * g + 1];
    }
    const double cnt = (double)hw * cpg;
    const double mean = S / cnt;
    double var = SS / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    gstat[2 * g] = (float)mean;
    gstat[2 * g + 1] = (float)(1.0 / sqrt(var + (double)eps));
  }
  __syncthreads();
  if (np == 0) return;
  float2 sc[4], sf[4];
  {
    const int gB = min(gA + 1, gslab - 1);
    const float mA = gstat[2 * gA], rA = gstat[2 * gA + 1], mB = gstat[2 * gB], rB = gstat[2 * gB + 1];
    const float gm[8] = {gm0.x, gm0.y, gm0.z, gm0.w, gm1.x, gm1.y, gm1.z, gm1.w};
    const float bt[8] = {bt0.x, bt0.y, bt0.z, bt0.w, bt1.x, bt1.y, bt1.z, bt1.w};
    // SiLU(y) = h + h tanh(h) with h = y / 2: the halving is folded into scale and shift (exact)
    const float half = SILU ? 0.5f : 1.0f;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float r = (e < split2 ? rA : rB) * half, m = e < split2 ? mA : mB;
      sc[e].x = r * gm[2 * e];
      sc[e].y = r * gm[2 * e + 1];
      sf[e].x = fmaf(-m, sc[e].x, half * bt[2 * e]);
      sf[e].y = fmaf(-m, sc[e].y, half * bt[2 * e + 1]);
    }
  }
  char* op = reinterpret_cast<char*>(out + ((int64_t)n * hw + p_begin + pl) * C + ch);
  const int64_t ostep = (int64_t)lanes * C * 2;
#pragma unroll 2
  for (int j = 0; j < np; ++j, op += ostep) {
    const uint4 v = lds128(slot0 + j * sstep);
    const uint32_t u[4] = {v.x, v.y, v.z, v.w};
    uint32_t o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float2 y = __ffma2_rn(unpack_act2(u[e], F16), sc[e], sf[e]);
      if (SILU) {
        float2 t;
        asm("tanh.approx.f32 %0, %1;" : "=f"(t.x) : "f"(y.x));
        asm("tanh.approx.f32 %0, %1;" : "=f"(t.y) : "f"(y.y));
        y = __ffma2_rn(y, t, y);
      }
      o[e] = pack_act2(y.x, y.y, F16);
    }
    *reinterpret_cast<uint4*>(op) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// Launch with an optional cluster dimension (the cluster size is a runtime choice of the shape) and programmatic dependent launch.
template <typename... KArgs, typename... Args>
cudaError_t launch_clustered(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, int cl, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[2];
  int na = 0;
  if (cl > 1) {
    at[na].id = cudaLaunchAttributeClusterDimension;
    at[na].val.clusterDim.x = cl;
    at[na].val.clusterDim.y = 1;
    at[na].val.clusterDim.z = 1;
    ++na;
  }
  if (cpd_pdl_enabled()) {
    at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = at;
  cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e && e[0] ? atoi(e) : dflt;
}

// Shape -> (slab channels, cluster size, pixel lanes, threads, slots per thread) of gn_slab_kernel; false = the shape does not fit.
struct SlabChoice { int slab_c, cl, lanes, threads, slots, ppc, hdr; uint32_t magic_vps, magic_cpg; size_t smem; };
// x / d == (x * magic) >> 16 for every x < limit?  (the kernel divides thread ids by the vectors per pixel and channels by the group size)
bool magic_div(int d, int limit, uint32_t* magic) {
  const uint32_t m = (65536u + d - 1) / d;
  for (int x = 0; x < limit; ++x)
    if ((int)(((uint32_t)x * m) >> 16) != x / d) return false;
  *magic = m;
  return true;
}
bool choose_slab(int C, int hw, SlabChoice* o) {
  const int cpg = C / GROUPS;
  if (cpg < 4 || cpg % 2 || (cpg < 8 && cpg != 4)) return false;  // a vector of 8 channels: at most two groups, split at an even channel
  auto lcm = [](int a, int b) { int x = a, y = b; while (y) { const int t = x % y; x = y; y = t; } return a / x * b; };
  // Measured on the SD-1.5 / SD-2.1 shapes (profiles/r02_gn_slab_bench.txt): blocks of <= 48 KB and 256 threads, four per SM, else
  // <= 100 KB, two per SM; 512-thread blocks, unaligned slabs first and one 200 KB block per SM were all slower.
  const int small_kb = 48, big_kb = 100, tmax = 256;
  int slabs[2] = {lcm(cpg, 16), lcm(cpg, 8)};  // [0]: sector-aligned (32 bytes), [1]: the smallest slab of whole groups and vectors
  while (slabs[0] * 2 < 128 && slabs[0] * 2 <= C && (slabs[0] * 2) / cpg <= 8) slabs[0] *= 2;  // at least a 128-byte line per pixel
  const int order[4][2] = {{0, small_kb}, {0, big_kb}, {1, small_kb}, {1, big_kb}};
  for (const auto& cand : order) {
    const int slab_c = slabs[cand[0]], limit_kb = cand[1];
    if (C % slab_c || slab_c / cpg > 8) continue;
    const int vps = slab_c / 8;
    if (vps > 32 || vps > tmax) continue;
    uint32_t mv, mc;
    if (!magic_div(vps, 512, &mv) || !magic_div(cpg, slab_c + 1, &mc)) continue;
    for (int cl : {1, 2, 4, 8}) {
      if (cl > hw) break;
      const int ppc = (hw + cl - 1) / cl;
      if ((int64_t)ppc * slab_c * 2 > (int64_t)limit_kb * 1024) continue;
      int lanes = tmax / vps;
      if (lanes > ppc) lanes = ppc;
      const int slots = (ppc + lanes - 1) / lanes;
      lanes = (ppc + slots - 1) / slots;
      const int threads = ((lanes * vps + 31) / 32) * 32;
      if (threads < 2 * (slab_c / cpg)) continue;  // the reduction's (group, statistic) threads
      const int hdr = 1088 + vps * (threads / 32) * 16;
      const size_t smem = (size_t)hdr + (size_t)slots * threads * 16;
      if (smem > 220 * 1024) continue;
      *o = {slab_c, cl, lanes, threads, slots, ppc, hdr, mv, mc, smem};
      return true;
    }
  }
  return false;
}

bool slab_enabled() {
  static int slab_on = -1;
  if (slab_on < 0) slab_on = env_int("CPD_GN_SLAB", 1);
  return slab_on != 0;
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

// ---- LayerNorm, sub-warp layout: LPR lanes share one row, VPL 16-byte vectors per lane (c = 8 * LPR * VPL), 32 / LPR rows
// per warp and pass, R passes whose loads are all issued before the first use (R * VPL * 16 bytes in flight per lane, every
// lane busy - the one-warp-per-row layout leaves 3/4 of the lanes idle on the second vector at c = 320).
template <int LPR, int VPL, int R>
__global__ void __launch_bounds__(256, 3) layernorm_sub_kernel(const bf16* __restrict__ x, int rows, int c,
                                                               const float* __restrict__ gamma, const float* __restrict__ beta,
                                                               float eps, int f16, bf16* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int RPW = 32 / LPR;  // rows per warp and pass
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int sub = lane % LPR;
  const int row0 = warp * (RPW * R) + lane / LPR;
  if (warp * (RPW * R) >= rows) return;
  uint4 raw[R][VPL];
#pragma unroll
  for (int rr = 0; rr < R; ++rr) {
    const int row = row0 + rr * RPW;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      raw[rr][i] = make_uint4(0, 0, 0, 0);
      if (row < rows) raw[rr][i] = __ldg(reinterpret_cast<const uint4*>(x + (int64_t)row * c + (sub + i * LPR) * 8));
    }
  }
  float mean[R], rstd[R];
  const float inv_c = 1.0f / (float)c;
#pragma unroll
  for (int rr = 0; rr < R; ++rr) {
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const uint32_t w[4] = {raw[rr][i].x, raw[rr][i].y, raw[rr][i].z, raw[rr][i].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = unpack_act2(w[e], f16 != 0);
        sum += f.x + f.y;
      }
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    mean[rr] = sum * inv_c;
    float var = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const uint32_t w[4] = {raw[rr][i].x, raw[rr][i].y, raw[rr][i].z, raw[rr][i].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = unpack_act2(w[e], f16 != 0);
        const float d0 = f.x - mean[rr], d1 = f.y - mean[rr];
        var += d0 * d0 + d1 * d1;
      }
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
    rstd[rr] = rsqrtf(var * inv_c + eps);
  }
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int col = (sub + i * LPR) * 8;
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + col));
    const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + col + 4));
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + col));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + col + 4));
    const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int rr = 0; rr < R; ++rr) {
      const int row = row0 + rr * RPW;
      if (row >= rows) continue;
      const uint32_t w[4] = {raw[rr][i].x, raw[rr][i].y, raw[rr][i].z, raw[rr][i].w};
      uint32_t o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = unpack_act2(w[e], f16 != 0);
        o[e] = pack_act2((f.x - mean[rr]) * rstd[rr] * gg[2 * e] + bb[2 * e], (f.y - mean[rr]) * rstd[rr] * gg[2 * e + 1] + bb[2 * e + 1],
                         f16 != 0);
      }
      *reinterpret_cast<uint4*>(out + (int64_t)row * c + col) = make_uint4(o[0], o[1], o[2], o[3]);
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
  {  // single pass with the slab in shared memory, split over a cluster where needed (CPD_GN_SLAB=0 or a shape that does not fit: the two kernels below)
    SlabChoice sc;
    if (slab_enabled() && n_img <= 65535 && choose_slab(C, hw, &sc)) {
      const dim3 grid(sc.cl, C / sc.slab_c, n_img);  // the cluster is the grid's x dimension
#define CPD_GN_SLAB_LAUNCH(F, S)                                                                                                  \
  do {                                                                                                                            \
    CPD_SMEM_OPTIN((gn_slab_kernel<F, S>), 220 * 1024);                                                                           \
    CPD_CUDA_CHECK(launch_clustered(gn_slab_kernel<F, S>, grid, dim3(sc.threads), sc.smem, s, sc.cl, (const bf16*)a0,             \
                                    (const bf16*)a1, c0, c1, hw, sc.slab_c, sc.lanes, sc.ppc, sc.magic_vps, sc.magic_cpg, sc.hdr,  \
                                    (sc.slots + 1) / 2, gamma, beta, eps, (bf16*)out));                                                               \
  } while (0)
      if (act_fp16) { if (silu) CPD_GN_SLAB_LAUNCH(true, true); else CPD_GN_SLAB_LAUNCH(true, false); }
      else { if (silu) CPD_GN_SLAB_LAUNCH(false, true); else CPD_GN_SLAB_LAUNCH(false, false); }
#undef CPD_GN_SLAB_LAUNCH
      CPD_CUDA_CHECK(cudaGetLastError());
      return CPD_OK;
    }
  }
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
  const size_t ring_bytes = (size_t)RING * threads * 16;
  const size_t shm = ring_bytes + sizeof(float) * (2 * C + 2 * GROUPS);
  const size_t shm_stats = ring_bytes + sizeof(float) * 2 * C * lanes;
  CPD_REQUIRE(shm_stats <= 160 * 1024 && shm <= 160 * 1024, "cpd_groupnorm: C=%d needs %zu bytes of shared memory", C, shm_stats);
  CPD_SMEM_OPTIN(gn_stats_kernel<true>, 160 * 1024);
  CPD_SMEM_OPTIN(gn_stats_kernel<false>, 160 * 1024);
  CPD_SMEM_OPTIN(gn_apply_kernel<true>, 160 * 1024);
  CPD_SMEM_OPTIN(gn_apply_kernel<false>, 160 * 1024);
  const dim3 grid(chunks, n_img);
  if (act_fp16) {
    CPD_CUDA_CHECK(cpd_launch(gn_stats_kernel<true>, grid, dim3(threads), shm_stats, s, (const bf16*)a0, (const bf16*)a1, c0, c1, hw, px_per_block, stats));
    CPD_CUDA_CHECK(cpd_launch(gn_apply_kernel<true>, grid, dim3(threads), shm, s, (const bf16*)a0, (const bf16*)a1, c0, c1, hw, px_per_block,
                              (const double*)stats, (const long long*)nullptr, gamma, beta, eps, silu, (bf16*)out));
  } else {
    CPD_CUDA_CHECK(cpd_launch(gn_stats_kernel<false>, grid, dim3(threads), shm_stats, s, (const bf16*)a0, (const bf16*)a1, c0, c1, hw, px_per_block, stats));
    CPD_CUDA_CHECK(cpd_launch(gn_apply_kernel<false>, grid, dim3(threads), shm, s, (const bf16*)a0, (const bf16*)a1, c0, c1, hw, px_per_block,
                              (const double*)stats, (const long long*)nullptr, gamma, beta, eps, silu, (bf16*)out));
  }
  CPD_CUDA_CHECK(cudaGetLastError());
  return CPD_OK;
}

// Kernel launches cpd_groupnorm makes for this shape: 1 (shared-memory slab kernel) or 2 (statistics + apply).
extern "C" int cpd_groupnorm_launches(int c, int n_img, int hw) {
  SlabChoice sc;
  if (c <= 0 || c % GROUPS || n_img <= 0 || hw <= 0) return 0;
  return slab_enabled() && n_img <= 65535 && choose_slab(c, hw, &sc) ? 1 : 2;
}

// GroupNorm from the producer's statistics: the apply pass alone (one read + one write of the tensor).
extern "C" cpd_status cpd_groupnorm_apply(const void* x, int c, int n_img, int hw, const float* gamma, const float* beta, float eps,
                                          int silu, int act_fp16, const long long* chan_sums, void* out, void* stream) {
  CPD_REQUIRE(x && gamma && beta && chan_sums && out, "cpd_groupnorm_apply: null pointer");
  CPD_REQUIRE(c > 0 && c % 8 == 0 && c % GROUPS == 0 && c <= 4096, "cpd_groupnorm_apply: bad channel count %d", c);
  CPD_REQUIRE(n_img > 0 && hw > 0, "cpd_groupnorm_apply: empty input");
  cudaStream_t s = (cudaStream_t)stream;
  const int vec_per_px = c / 8;
  int lanes = 256 / vec_per_px;
  if (lanes < 1) lanes = 1;
  if (lanes > hw) lanes = hw;
  const int threads = ((lanes * vec_per_px + 31) / 32) * 32;
  CPD_REQUIRE(threads <= 512, "cpd_groupnorm_apply: C=%d too large", c);
  int chunks = (148 * 4 + n_img - 1) / n_img;  // one wave of ~4 blocks per SM over all images, >= 8 pixels per pixel lane
  int px_per_block = (hw + chunks - 1) / chunks;
  if (px_per_block < 8 * lanes) px_per_block = 8 * lanes;
  chunks = (hw + px_per_block - 1) / px_per_block;
  const size_t shm = (size_t)RING * threads * 16 + sizeof(float) * (2 * c + 2 * GROUPS);
  CPD_REQUIRE(shm <= 160 * 1024, "cpd_groupnorm_apply: C=%d needs %zu bytes of shared memory", c, shm);
  CPD_SMEM_OPTIN(gn_apply_kernel<true>, 160 * 1024);
  CPD_SMEM_OPTIN(gn_apply_kernel<false>, 160 * 1024);
  const dim3 grid(chunks, n_img);
  if (act_fp16)
    CPD_CUDA_CHECK(cpd_launch(gn_apply_kernel<true>, grid, dim3(threads), shm, s, (const bf16*)x, (const bf16*)nullptr, c, 0, hw, px_per_block,
                              (const double*)nullptr, chan_sums, gamma, beta, eps, silu, (bf16*)out));
  else
    CPD_CUDA_CHECK(cpd_launch(gn_apply_kernel<false>, grid, dim3(threads), shm, s, (const bf16*)x, (const bf16*)nullptr, c, 0, hw, px_per_block,
                              (const double*)nullptr, chan_sums, gamma, beta, eps, silu, (bf16*)out));
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
#define CPD_LN_SUB(LPR, VPL, R)                                                                                                  \
  CPD_CUDA_CHECK(cpd_launch(layernorm_sub_kernel<LPR, VPL, R>, dim3(blocks((32 / LPR) * R)), dim3(256), 0, s, (const bf16*)x, rows, c, \
                            gamma, beta, eps, act_fp16, (bf16*)out))
  if (nvec == 40) CPD_LN_SUB(8, 5, 2);        // c = 320
  else if (nvec == 80) CPD_LN_SUB(16, 5, 2);  // c = 640
  else if (nvec == 160) CPD_LN_SUB(32, 5, 2);  // c = 1280
  else if (nvec <= 64) CPD_CUDA_CHECK(cpd_launch(layernorm_kernel<2, 4>, dim3(blocks(4)), dim3(256), 0, s, (const bf16*)x, rows, c, gamma, beta, eps, act_fp16, (bf16*)out));
  else if (nvec <= 160) CPD_CUDA_CHECK(cpd_launch(layernorm_kernel<5, 2>, dim3(blocks(2)), dim3(256), 0, s, (const bf16*)x, rows, c, gamma, beta, eps, act_fp16, (bf16*)out));
  else CPD_CUDA_CHECK(cpd_launch(layernorm_kernel<8, 1>, dim3(blocks(1)), dim3(256), 0, s, (const bf16*)x, rows, c, gamma, beta, eps, act_fp16, (bf16*)out));
#undef CPD_LN_SUB
  CPD_CUDA_CHECK(cudaGetLastError());
  return CPD_OK;
}
