// Persistent CTA-pair implicit-GEMM convolution / GEMM for sm_100a.
//
//   * tcgen05.mma.cta_group::2: the two CTAs of a cluster (one SM pair) compute one 256 x BN output tile.  Each
//     CTA stages its own 128 rows of A and HALF of the B tile (BN/2 weight rows), so the shared-memory operand
//     traffic per MMA is half that of a 1-CTA 128 x BN tile (which is shared-memory-bound at BN = 128).
//   * persistent: one cluster per SM pair loops over output tiles; the fp32 accumulator is double-buffered in
//     TMEM (2 x 256 columns), so the epilogue of tile i (TMEM -> registers -> bias / time-embedding / residual /
//     GEGLU -> 16-bit stores) overlaps the TMA + MMA main loop of tile i + 1.
//   * warp roles: warp 0 = TMA producer (both CTAs), warp 1 = MMA issuer (leader CTA only), warp 2 = TMEM
//     allocator, warps 4-11 = epilogue (two warps per TMEM lane quarter, each taking half of the columns).
//   * BN is a RUNTIME multiple of 32 (<= 256) chosen per layer so that BN divides N (320 -> 160, 640 -> 160/128,
//     1280 -> 256/160) and the tile count fills whole waves of 74 SM pairs.
//
// Same operand addressing as gemm_umma.cu (gemm_geom.cuh): 3x3 taps are shifted TMA boxes with hardware zero fill,
// the skip concat is two tensor maps.  Replaces the cuDNN/cuBLAS calls behind nn.Conv2d / nn.Linear in
// cpd/models/unet.py:105,153-160,210,236,247 and cpd/models/attention.py:92-118,183-190,508-524.
#include "gemm_geom.cuh"

namespace {
using namespace cpd_gemm;

constexpr int EPI_WARPS = 8;
constexpr int FIRST_EPI_WARP = 4;
constexpr int NUM_THREADS2 = 32 * (FIRST_EPI_WARP + EPI_WARPS);
constexpr int ACC_COLS = 256;   // TMEM columns per accumulator stage
constexpr int MAX_STAGES = 8;
constexpr int A_BYTES = BM * BK * 2;
constexpr int NUM_SM_PAIRS = 74;

struct Gemm2Args {
  CUtensorMap map_a0, map_a1, map_b;
  ConvGeom g;
  const float* bias;
  const float* rowvec;
  const bf16* residual;
  bf16* d;
  int bn;           // tile N
  int stages;
  int stage_bytes;  // A_BYTES + bn/2 * 128
  int m_tiles2;     // 256-row tiles
  int n_tiles;
  int num_k;        // taps * 64-channel blocks
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS2, 1)
    gemm2_kernel(const __grid_constant__ Gemm2Args args) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int stages = args.stages;
  const int stage_bytes = args.stage_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + stages * stage_bytes);
  uint64_t* empty_bar = full_bar + MAX_STAGES;
  uint64_t* tmem_full = empty_bar + MAX_STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const ConvGeom& g = args.g;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;
  const int total_tiles = args.m_tiles2 * args.n_tiles;
  const int num_k = args.num_k;
  const int cbt = g.cb0 + g.cb1;

  cluster_sync_all();  // both CTAs of the pair are resident before the pair-wide TMEM allocation
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&args.map_a0);
    tma_prefetch_desc(&args.map_b);
    if (g.cb1 > 0) tma_prefetch_desc(&args.map_a1);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full_bar[s], 2);   // leader: its own arrive.expect_tx + the peer's remote arrive
      mbar_init(&empty_bar[s], 1);  // tcgen05.commit multicast to both CTAs
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);               // commit multicast to both CTAs
      mbar_init(&tmem_empty[s], 2 * EPI_WARPS);  // leader: one arrive per epilogue warp of both CTAs
    }
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc_pair(tmem_ptr_smem, 512);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ================= TMA producer (both CTAs) =================
    // One thread; the loop body must stay far below the MMA time of a stage (BN/2 * 4 cycles), so everything that
    // needs a division is hoisted to once per tile and the (tap, channel-block) walk is incremental.
    {
      const int box_bytes = g.tw * g.th * g.nb * 128;
      const uint32_t tx_bytes = 2u * (uint32_t)stage_bytes;
      const int bn_half = args.bn >> 1;
      int stage = 0;
      uint32_t phase = 0;
      for (int t = cluster_id; t < total_tiles; t += num_clusters) {
        const int m2 = t / args.n_tiles;
        const int n_tile = t - m2 * args.n_tiles;
        const int m_tile = m2 * 2 + (int)rank;
        const int b_row = n_tile * args.bn + (int)rank * bn_half;
        // box 0 of this tile at tap (0, 0) (the only box when nbox == 1, the common case)
        const BoxCoord b0 = box_coord(g, m_tile, 0, 4, 0);  // tap 4 = centre: no shift
        int tap = 0, cb = 0;
        int dy = 0, dx = 0, py = 0, px = 0;
        if (g.taps == 9) {
          if (g.stride == 1) { dy = -1; dx = -1; } else { dy = -1; py = 1; dx = -1; px = 1; }
        }
        for (int kt = 0; kt < num_k; ++kt) {
          mbar_wait(&empty_bar[stage], phase ^ 1, 1);
          uint8_t* sa = smem + stage * stage_bytes;
          const bool src1 = cb >= g.cb0;
          const CUtensorMap* ma = src1 ? &args.map_a1 : &args.map_a0;
          const int cc = (src1 ? (cb - g.cb0) : cb) * BK + px * (src1 ? g.c1 : g.c0);
          if (elect_one()) {
            if (rank == 0)
              mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
            else
              mbar_arrive_cluster(&full_bar[stage], 0);
            if (g.nbox == 1) {
              tma_load_5d_pair(sa, ma, &full_bar[stage], cc, b0.x + dx, py, b0.y + dy, b0.n);
            } else {
              for (int j = 0; j < g.nbox; ++j) {
                const BoxCoord bc = box_coord(g, m_tile, j, 4, 0);
                tma_load_5d_pair(sa + j * box_bytes, ma, &full_bar[stage], cc, bc.x + dx, py, bc.y + dy, bc.n);
              }
            }
            tma_load_2d_pair(sa + A_BYTES, &args.map_b, &full_bar[stage], kt * BK, b_row);
          }
          __syncwarp();
          if (++stage == stages) {
            stage = 0;
            phase ^= 1;
          }
          if (++cb == cbt) {  // next filter tap
            cb = 0;
            ++tap;
            const int ky = tap / 3, kx = tap - ky * 3;
            if (g.stride == 1) {
              dy = ky - 1;
              dx = kx - 1;
            } else {  // input row = 2*oy + ky - 1 = 2*(oy + dy) + py
              dy = (ky == 0) ? -1 : 0;
              py = (ky == 0) ? 1 : ky - 1;
              dx = (kx == 0) ? -1 : 0;
              px = (kx == 0) ? 1 : kx - 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (leader CTA only) =================
    if (rank == 0) {
      const uint32_t idesc = g.idesc;
      const uint64_t desc0 = umma_desc_sw128(smem_u32(smem));   // stage 0, A operand
      const uint64_t stage_step = (uint64_t)(stage_bytes >> 4);  // descriptor address field is in 16-byte units
      const uint64_t b_off = (uint64_t)(A_BYTES >> 4);
      int stage = 0;
      uint32_t phase = 0;
      uint64_t da = desc0;
      int it = 0;
      for (int t = cluster_id; t < total_tiles; t += num_clusters, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (uint32_t)((it >> 1) & 1);
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1, 4);  // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * ACC_COLS);
        uint32_t accum = 0;
        for (int kt = 0; kt < num_k; ++kt) {
          mbar_wait(&full_bar[stage], phase, 2);
          tc_fence_after();
          const uint64_t db = da + b_off;
          if (elect_one()) {
            // +2 in the address field = 32 bytes = 16 elements along K inside the 128-byte swizzle atom
            umma_f16_pair(d_tmem, da, db, idesc, accum);
            umma_f16_pair(d_tmem, da + 2, db + 2, idesc, 1u);
            umma_f16_pair(d_tmem, da + 4, db + 4, idesc, 1u);
            umma_f16_pair(d_tmem, da + 6, db + 6, idesc, 1u);
            umma_commit_pair(&empty_bar[stage], 3);  // frees the stage in both CTAs
          }
          __syncwarp();
          accum = 1u;
          da += stage_step;
          if (++stage == stages) {
            stage = 0;
            phase ^= 1;
            da = desc0;
          }
        }
        if (elect_one()) umma_commit_pair(&tmem_full[acc], 3);  // accumulator complete, in both CTAs
        __syncwarp();
      }
    }
  } else if (warp >= FIRST_EPI_WARP) {
    // ================= epilogue (warps 4..11 of both CTAs) =================
    const bool of16 = g.out_fp16 != 0;
    const int q = warp & 3;                           // TMEM lane quarter this warp may access
    const int half = (warp - FIRST_EPI_WARP) >> 2;    // which half of the tile's columns
    const int r = q * 32 + lane;                      // row within this CTA's 128-row tile
    const int bn = args.bn;
    int it = 0;
    for (int t = cluster_id; t < total_tiles; t += num_clusters, ++it) {
      const int m2 = t / args.n_tiles;
      const int n_tile = t - m2 * args.n_tiles;
      const RowCoord rc = row_coord(g, m2 * 2 + (int)rank, r);
      const bool valid = rc.valid;
      const int64_t row = rc.row;
      const int acc = it & 1;
      const uint32_t acc_phase = (uint32_t)((it >> 1) & 1);
      mbar_wait(&tmem_full[acc], acc_phase, 3);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * ACC_COLS);

      if (g.epilogue == CPD_EPI_GEGLU) {
        // tile columns [0, bn/2) = value, [bn/2, bn) = gate  ->  bn/2 output columns
        const int hb = bn >> 1;
        const int nv = hb >> 4;  // 16-column value chunks
        const int c_lo = half == 0 ? 0 : (nv + 1) / 2, c_hi = half == 0 ? (nv + 1) / 2 : nv;
        const int ncol0 = n_tile * hb;
        const float* bias_v = args.bias ? args.bias + n_tile * bn : nullptr;
#pragma unroll 1
        for (int ch = c_lo; ch < c_hi; ++ch) {
          const int c = ch * 16;
          uint32_t va[16], vg[16];
          tmem_ld16(taddr + c, va);
          tmem_ld16(taddr + hb + c, vg);
          tmem_ld_wait();
          if (valid && ncol0 + c < g.n_store) {
            uint32_t packed[8];
#pragma unroll
            for (int e = 0; e < 16; e += 2) {
              float a0 = __uint_as_float(va[e]), a1 = __uint_as_float(va[e + 1]);
              float g0 = __uint_as_float(vg[e]), g1 = __uint_as_float(vg[e + 1]);
              if (bias_v) {
                a0 += __ldg(bias_v + c + e);
                a1 += __ldg(bias_v + c + e + 1);
                g0 += __ldg(bias_v + hb + c + e);
                g1 += __ldg(bias_v + hb + c + e + 1);
              }
              // the reference rounds the projection to the model dtype before x * gelu(gate) (attention.py:98-100)
              a0 = round_act(a0, of16);
              a1 = round_act(a1, of16);
              g0 = round_act(gelu_erf_f(round_act(g0, of16)), of16);
              g1 = round_act(gelu_erf_f(round_act(g1, of16)), of16);
              packed[e / 2] = pack_act2(a0 * g0, a1 * g1, of16);
            }
            uint4* dst = reinterpret_cast<uint4*>(args.d + row * g.ldd + ncol0 + c);
            dst[0] = make_uint4(packed[0], packed[1], packed[2], packed[3]);
            dst[1] = make_uint4(packed[4], packed[5], packed[6], packed[7]);
          }
        }
      } else {
        const int nch = bn >> 5;  // 32-column chunks
        const int c_lo = half == 0 ? 0 : (nch + 1) / 2, c_hi = half == 0 ? (nch + 1) / 2 : nch;
        const int ncol0 = n_tile * bn;
        const float* rv = (args.rowvec && valid) ? args.rowvec + (int64_t)rc.n * g.rowvec_stride : nullptr;
#pragma unroll 1
        for (int ch = c_lo; ch < c_hi; ++ch) {
          uint32_t v[32];
          tmem_ld32(taddr + ch * 32, v);
          tmem_ld_wait();
          if (valid) {
#pragma unroll
            for (int h8 = 0; h8 < 4; ++h8) {  // 4 groups of 8 columns = one 16-byte store each
              const int col = ncol0 + ch * 32 + h8 * 8;
              if (col < g.n_store) {
                float f[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[h8 * 8 + e]);
                if (args.bias) {
                  const float4 b0 = __ldg(reinterpret_cast<const float4*>(args.bias + col));
                  const float4 b1 = __ldg(reinterpret_cast<const float4*>(args.bias + col + 4));
                  f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w;
                  f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
                }
                if (rv) {
                  const float4 b0 = __ldg(reinterpret_cast<const float4*>(rv + col));
                  const float4 b1 = __ldg(reinterpret_cast<const float4*>(rv + col + 4));
                  f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w;
                  f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
                }
                if (args.residual) {
                  const uint4 rr = *reinterpret_cast<const uint4*>(args.residual + row * g.ld_res + col);
                  const float2 r0 = unpack_act2(rr.x, of16), r1 = unpack_act2(rr.y, of16), r2 = unpack_act2(rr.z, of16),
                               r3 = unpack_act2(rr.w, of16);
                  f[0] += r0.x; f[1] += r0.y; f[2] += r1.x; f[3] += r1.y;
                  f[4] += r2.x; f[5] += r2.y; f[6] += r3.x; f[7] += r3.y;
                }
                *reinterpret_cast<uint4*>(args.d + row * g.ldd + col) =
                    make_uint4(pack_act2(f[0], f[1], of16), pack_act2(f[2], f[3], of16), pack_act2(f[4], f[5], of16),
                               pack_act2(f[6], f[7], of16));
              }
            }
          }
        }
      }
      // all TMEM reads of this warp are complete (tcgen05.wait::ld above): hand the accumulator back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(&tmem_empty[acc], 0);
    }
  }

  __syncwarp();
  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
  }
}

// Tile width: a multiple of 32 that (a) wastes few padded columns and (b) fills whole waves of SM pairs.
int pick_bn(int m_tiles2, int n_out, bool geglu) {
  static const int cand_plain[] = {256, 224, 192, 160, 128, 96, 64};
  static const int cand_geglu[] = {256, 128};
  const int* cand = geglu ? cand_geglu : cand_plain;
  const int ncand = geglu ? 2 : 7;
  int best = cand[0];
  double best_cost = 1e30;
  for (int i = 0; i < ncand; ++i) {
    const int bn = cand[i];
    if (geglu && n_out % bn) continue;
    const int n_tiles = (n_out + bn - 1) / bn;
    const long tiles = (long)m_tiles2 * n_tiles;
    const long waves = (tiles + NUM_SM_PAIRS - 1) / NUM_SM_PAIRS;
    // per-tile time ~ MMA cycles (bn/2 per K=16 step, floored by the shared-memory operand feed) + fixed overhead
    const double cost = (double)waves * ((bn > 96 ? bn : 96) + 16);
    if (cost < best_cost - 1e-9) {
      best_cost = cost;
      best = bn;
    }
  }
  return best;
}

}  // namespace

cpd_status cpd_gemm_conv_2cta(const cpd_gemm_params* p, void* stream) {
  Gemm2Args args;
  ConvGeom& g = args.g;
  int m_tiles_cta = 0;
  int rc = fill_geometry(p, &g, &args.map_a0, &args.map_a1, &m_tiles_cta);
  if (rc) return rc;
  const bool geglu = p->epilogue == CPD_EPI_GEGLU;
  args.m_tiles2 = (m_tiles_cta + 1) / 2;
  int bn = (p->variant >= 32) ? p->variant : pick_bn(args.m_tiles2, p->n_out, geglu);
  if (geglu) bn = p->geglu_block ? p->geglu_block : 128;  // fixed by the weight interleave
  CPD_REQUIRE(bn % 32 == 0 && bn >= 32 && bn <= 256, "cpd_gemm_conv: tile width %d must be a multiple of 32 in [32, 256]", bn);
  if (geglu) {
    CPD_REQUIRE((bn == 128 || bn == 256) && p->n_out % bn == 0,
                "cpd_gemm_conv: GEGLU needs geglu_block 128 or 256 dividing n_out (geglu_block=%d, n_out=%d)", bn, p->n_out);
    g.n_store = p->n_out / 2;
  } else {
    g.n_store = p->n_out;
  }
  args.bn = bn;
  args.n_tiles = (p->n_out + bn - 1) / bn;
  args.num_k = g.taps * (g.cb0 + g.cb1);
  args.stage_bytes = A_BYTES + (bn / 2) * 128;
  int stages = (227 * 1024 - 1024 - 512) / args.stage_bytes;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  args.stages = stages;
  g.idesc = umma_idesc_f16(256, bn, p->a_fp16 != 0, p->b_fp16 != 0);
  rc = make_b_map(p, g.taps, bn / 2, &args.map_b);
  if (rc) return rc;
  args.bias = p->bias;
  args.rowvec = p->rowvec;
  args.residual = reinterpret_cast<const bf16*>(p->residual);
  args.d = reinterpret_cast<bf16*>(p->d);

  const int smem_bytes = stages * args.stage_bytes + 512 + 1024;
  static bool configured = false;
  if (!configured) {
    CPD_CUDA_CHECK(cudaFuncSetAttribute(gemm2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  const long total_tiles = (long)args.m_tiles2 * args.n_tiles;
  const int clusters = (int)(total_tiles < NUM_SM_PAIRS ? total_tiles : NUM_SM_PAIRS);
  gemm2_kernel<<<dim3(2 * clusters), NUM_THREADS2, smem_bytes, (cudaStream_t)stream>>>(args);
  CPD_CUDA_CHECK(cudaGetLastError());
  return CPD_OK;
}
