// Implicit-GEMM convolution / GEMM for sm_100a: tcgen05.mma with the accumulator in TMEM, operands staged
// into 128B-swizzled shared memory by TMA (cp.async.bulk.tensor), mbarrier producer/consumer pipeline,
// warp-specialised roles (warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner, warps 2-5 = epilogue).
//
// One CTA computes a 128 x BN output tile.  The A operand (activations, NHWC bf16) is addressed through a
// rank-5 tensor map (c, x, parity, y, n): a 3x3 tap is just a shifted TMA box whose out-of-bounds part is
// zero-filled by the TMA unit (= the conv's zero padding), so no im2col buffer ever exists.  Stride-2 convs
// use the parity view [n][h/2][2][w/2][2*C] of the same buffer.  The skip-concat input of the UNet output
// blocks (unet.py:814) is read through two tensor maps without materialising the concatenation.
//
// Replaces the cuDNN/cuBLAS calls behind nn.Conv2d / nn.Linear in cpd/models/unet.py:105,153-160,210,236,247
// and cpd/models/attention.py:92-118,183-190,508-524.
#include "gemm_geom.cuh"

namespace {
using namespace cpd_gemm;

constexpr int NUM_THREADS = 192;

struct GemmArgs {
  CUtensorMap map_a0, map_a1, map_b;
  ConvGeom g;
  const float* bias;
  const float* rowvec;
  const bf16* residual;
  bf16* d;
};

template <int BN, int STAGES>
struct SmemLayout {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + 256 + 1024;  // + barriers + alignment slack
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(NUM_THREADS, 1) gemm_conv_kernel(const __grid_constant__ GemmArgs args) {
  using L = SmemLayout<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const ConvGeom& g = args.g;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_tile = blockIdx.x;
  const int n_tile = blockIdx.y;
  const int cbt = g.cb0 + g.cb1;
  const int num_k = g.taps * cbt;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&args.map_a0);
    tma_prefetch_desc(&args.map_b);
    if (g.cb1 > 0) tma_prefetch_desc(&args.map_a1);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_smem, BN);
    tmem_relinquish();
  }
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      const int box_bytes = g.tw * g.th * g.nb * 128;
      int stage = 0;
      uint32_t phase = 0;
      for (int kt = 0; kt < num_k; ++kt) {
        const int tap = kt / cbt;
        const int cb = kt - tap * cbt;
        mbar_wait(&empty_bar[stage], phase ^ 1, 1);
        uint8_t* sa = smem + stage * L::STAGE_BYTES;
        uint8_t* sb = sa + L::A_BYTES;
        mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)(g.nbox * box_bytes + (L::STAGE_BYTES - L::A_BYTES)));  // boxes may cover < 128 rows
        for (int j = 0; j < g.nbox; ++j) {
          const BoxCoord bc = box_coord(g, m_tile, j, tap, cb);
          tma_load_5d(sa + j * box_bytes, bc.src1 ? &args.map_a1 : &args.map_a0, &full_bar[stage], bc.c, bc.x, bc.p, bc.y, bc.n);
        }
        tma_load_2d(sb, &args.map_b, &full_bar[stage], kt * BK, n_tile * BN);
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      const uint32_t idesc = g.idesc;
      int stage = 0;
      uint32_t phase = 0;
      for (int kt = 0; kt < num_k; ++kt) {
        mbar_wait(&full_bar[stage], phase, 2);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * L::STAGE_BYTES);
        const uint32_t sb = sa + L::A_BYTES;
        const uint64_t da = umma_desc_sw128(sa);
        const uint64_t db = umma_desc_sw128(sb);
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {
          // advance 32 bytes (16 bf16) along K inside the 128-byte swizzle atom: +2 in the >>4 address field
          umma_bf16(tmem_base, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kt > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[stage]);  // frees this smem stage when the MMAs above have read it
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      umma_commit(tmem_full_bar);  // accumulator complete
    }
  } else {
    // ================= epilogue (warps 2..5) =================
    const bool of16 = g.out_fp16 != 0;
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int r = q * 32 + lane;       // row within the tile
    const RowCoord rc = row_coord(g, m_tile, r);
    const bool valid = rc.valid;
    const int64_t row = rc.row;
    const int n = rc.n;

    mbar_wait(tmem_full_bar, 0, 3);
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);

    if (g.epilogue == CPD_EPI_GEGLU) {
      // tile columns [0, BN/2) = value, [BN/2, BN) = gate  ->  BN/2 output columns
      constexpr int HALF = BN / 2;
      const int ncol0 = n_tile * HALF;
#pragma unroll 1
      for (int c = 0; c < HALF; c += 16) {
        uint32_t va[16], vg[16];
        tmem_ld16(taddr + c, va);
        tmem_ld16(taddr + HALF + c, vg);
        tmem_ld_wait();
        if (valid && ncol0 + c < g.n_store) {
          uint32_t packed[8];
#pragma unroll
          for (int e = 0; e < 16; e += 2) {
            float a0 = __uint_as_float(va[e]), a1 = __uint_as_float(va[e + 1]);
            float g0 = __uint_as_float(vg[e]), g1 = __uint_as_float(vg[e + 1]);
            if (args.bias) {
              a0 += __ldg(args.bias + n_tile * BN + c + e);
              a1 += __ldg(args.bias + n_tile * BN + c + e + 1);
              g0 += __ldg(args.bias + n_tile * BN + HALF + c + e);
              g1 += __ldg(args.bias + n_tile * BN + HALF + c + e + 1);
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
      const int ncol0 = n_tile * BN;
      const float* rv = (args.rowvec && valid) ? args.rowvec + (int64_t)n * g.rowvec_stride : nullptr;
#pragma unroll 1
      for (int c = 0; c < BN; c += 32) {
        uint32_t v[32];
        tmem_ld32(taddr + c, v);
        tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int h8 = 0; h8 < 4; ++h8) {  // 4 groups of 8 columns = one 16-byte store each
            const int col = ncol0 + c + h8 * 8;
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
                const uint4 rr = __ldg(reinterpret_cast<const uint4*>(args.residual + row * g.ld_res + col));
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
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BN);
  }
}

template <int BN, int STAGES>
cpd_status launch(const GemmArgs& args, int m_tiles, int n_tiles, cudaStream_t s) {
  using L = SmemLayout<BN, STAGES>;
  CPD_SMEM_OPTIN((gemm_conv_kernel<BN, STAGES>), L::TOTAL);
  CPD_CUDA_CHECK(cpd_launch(gemm_conv_kernel<BN, STAGES>, dim3(dim3(m_tiles, n_tiles)), dim3(NUM_THREADS), L::TOTAL, s, args));
  CPD_CUDA_CHECK(cudaGetLastError());
  return CPD_OK;
}

}  // namespace

namespace cpd_gemm {

// Box of output pixels covered by one TMA load: nb images x th rows x tw columns, tw | w, a multiple of 8 rows (one swizzle
// atom), at most 128 rows; a CTA tile takes floor(128 / rows) boxes.  nb > 1 only when the box spans whole images.
// th need NOT divide h and the rows need not divide 128 (round 2): the last box of an image then hangs over its bottom edge
// (TMA zero-fills the loads and clips the stores) and the tile's remaining rows are padding (row_coord: valid = false).  What
// this buys: a 12 x 12 image (SD-2.1 at 96 x 96, level 3) is two boxes of 12 x 10 - ONE TMA load per k-iteration and CTA - where
// the exact tiling needed eight 4 x 4 boxes (2700 instead of ~600 cycles per k-iteration: the single-thread TMA issue was the
// limiter, 65 us per 3x3 conv).  Cost model: tiles x (1 + 0.35 per extra box), ties to the fuller tile.
static bool choose_box(int h, int w, int n_img, int* tw, int* th, int* nb) {
  double best_cost = 1e30;
  int best_px = 0;
  *tw = *th = *nb = 0;
  for (int a = 1; a <= 128 && a <= w; ++a) {
    if (w % a) continue;
    for (int b = 1; a * b <= 128 && b <= h; ++b) {
      const int max_n = (a == w && b == h) ? 128 / (a * b) : 1;
      for (int c = 1; c <= max_n; c *= 2) {
        const int px = a * b * c;
        if (px % 8 || px > 128) continue;
        if (c > 1 && c > 2 * n_img) continue;  // do not pad tiny batches with many empty images
        const int nbox = 128 / px;
        const long boxes = (long)((n_img + c - 1) / c) * (w / a) * ((h + b - 1) / b);
        const long tiles = (boxes + nbox - 1) / nbox;
        const double cost = (double)tiles * (1.0 + 0.35 * (nbox - 1));
        const int fill = px * nbox;
        if (cost < best_cost - 1e-9 || (cost < best_cost + 1e-9 && (fill > best_px || (fill == best_px && a > *tw)))) {
          best_cost = cost;
          best_px = fill;
          *tw = a;
          *th = b;
          *nb = c;
        }
      }
    }
  }
  return best_px > 0;
}

// NHWC activation view (c, x, parity, y, n) with a box of nb images x th rows x tw columns x 64 channels
static int make_a_map(const cpd_gemm_params* p, int tw, int th, int nb, const void* base, int csrc, CUtensorMap* m) {
  uint64_t dims[5], str[4];
  uint32_t box[5] = {64, (uint32_t)tw, 1, (uint32_t)th, (uint32_t)nb};
  if (p->stride == 1) {
    dims[0] = csrc; dims[1] = p->w_in; dims[2] = 1; dims[3] = p->h_in; dims[4] = p->n_img;
    str[0] = (uint64_t)csrc * 2;
    str[1] = (uint64_t)p->w_in * csrc * 2;  // parity dim (size 1)
    str[2] = (uint64_t)p->w_in * csrc * 2;
    str[3] = (uint64_t)p->h_in * p->w_in * csrc * 2;
  } else {
    dims[0] = 2 * (uint64_t)csrc; dims[1] = p->w_in / 2; dims[2] = 2; dims[3] = p->h_in / 2; dims[4] = p->n_img;
    str[0] = (uint64_t)csrc * 2 * 2;
    str[1] = (uint64_t)p->w_in * csrc * 2;
    str[2] = (uint64_t)p->w_in * csrc * 2 * 2;
    str[3] = (uint64_t)p->h_in * p->w_in * csrc * 2;
  }
  return cpd_make_tmap_bf16(m, base, 5, dims, str, box);
}

int fill_geometry(const cpd_gemm_params* p, ConvGeom* gp, CUtensorMap* map_a0, CUtensorMap* map_a1, int* m_tiles_cta) {
  CPD_REQUIRE(p != nullptr, "cpd_gemm_conv: null params");
  CPD_REQUIRE(p->a0 && p->wt && p->d, "cpd_gemm_conv: a0, wt and d must be non-null");
  CPD_REQUIRE(p->c0 > 0 && p->c0 % 64 == 0 && p->c1 >= 0 && p->c1 % 64 == 0,
              "cpd_gemm_conv: channel counts must be multiples of 64 (c0=%d c1=%d)", p->c0, p->c1);
  CPD_REQUIRE(p->c1 == 0 || p->a1, "cpd_gemm_conv: c1 > 0 needs a1");
  CPD_REQUIRE(p->ksize == 1 || p->ksize == 3, "cpd_gemm_conv: ksize must be 1 or 3 (got %d)", p->ksize);
  CPD_REQUIRE(p->stride == 1 || (p->stride == 2 && p->ksize == 3 && p->h_in % 2 == 0 && p->w_in % 2 == 0),
              "cpd_gemm_conv: stride must be 1, or 2 with a 3x3 kernel and even h/w");
  CPD_REQUIRE(p->n_img > 0 && p->h_in > 0 && p->w_in > 0, "cpd_gemm_conv: empty input (%d x %d x %d)", p->n_img, p->h_in, p->w_in);
  CPD_REQUIRE(p->n_out > 0 && p->n_out % 8 == 0, "cpd_gemm_conv: n_out=%d must be a positive multiple of 8", p->n_out);
  CPD_REQUIRE(p->ldd % 8 == 0 && (p->residual == nullptr || p->ld_res % 8 == 0), "cpd_gemm_conv: ldd/ld_res must be multiples of 8");
  CPD_REQUIRE(((uintptr_t)p->d & 15) == 0, "cpd_gemm_conv: d must be 16-byte aligned");
  CPD_REQUIRE(p->stride == 1 || p->c1 == 0, "cpd_gemm_conv: stride 2 supports a single source");
  CPD_REQUIRE((p->a_fp16 != 0) == (p->b_fp16 != 0),
              "cpd_gemm_conv: A and B must have the same 16-bit format (tcgen05 kind::f16 traps on mixed fp16 x bf16)");
  ConvGeom& g = *gp;
  g.taps = p->ksize * p->ksize;
  g.cb0 = p->c0 / 64;
  g.cb1 = p->c1 / 64;
  g.c0 = p->c0;
  g.c1 = p->c1;
  g.stride = p->stride;
  g.n_img = p->n_img;
  g.h_out = p->h_in / p->stride;
  g.w_out = p->w_in / p->stride;
  g.m_valid = p->m_valid;
  g.n_out = p->n_out;
  g.epilogue = p->epilogue;
  g.rowvec_stride = p->rowvec_stride;
  g.ld_res = p->ld_res;
  g.ldd = p->ldd;
  g.out_fp16 = p->out_fp16;
  const bool plain = (p->h_in == 1 && p->n_img == 1 && p->ksize == 1);
  if (plain) {
    g.tw = 128;
    g.th = 1;
    g.nb = 1;
    g.nbox = 1;
    g.bx_count = (g.w_out + 127) / 128;
    g.by_count = 1;
  } else {
    int tw = 0, th = 0, nb = 0;
    CPD_REQUIRE(choose_box(g.h_out, g.w_out, g.n_img, &tw, &th, &nb), "cpd_gemm_conv: no 8..128-row box tiles the %d x %d output",
                g.h_out, g.w_out);
    g.tw = tw;
    g.th = th;
    g.nb = nb;
    g.nbox = 128 / (tw * th * nb);            // floor: the rows beyond nbox boxes are padding
    g.bx_count = g.w_out / tw;
    g.by_count = (g.h_out + th - 1) / th;     // the last box of an image may hang over its bottom edge
  }
  g.inv_box_rows = 1.0f / (float)(g.tw * g.th * g.nb);
  g.inv_bx = 1.0f / (float)g.bx_count;
  g.inv_by = 1.0f / (float)g.by_count;
  g.inv_img_px = 1.0f / (float)(g.tw * g.th);
  g.inv_tw = 1.0f / (float)g.tw;
  const int64_t n_groups = (g.n_img + g.nb - 1) / g.nb;
  const int64_t total_boxes = n_groups * g.bx_count * g.by_count;
  *m_tiles_cta = (int)((total_boxes + g.nbox - 1) / g.nbox);

  int rc = make_a_map(p, g.tw, g.th, g.nb, p->a0, p->c0, map_a0);
  if (rc) return rc;
  if (p->c1 > 0) {
    rc = make_a_map(p, g.tw, g.th, g.nb, p->a1, p->c1, map_a1);
    if (rc) return rc;
  } else {
    *map_a1 = *map_a0;
  }
  return CPD_OK;
}

// Half-box views for the multicast kernel: the box is split in two along its outermost dimension of extent > 1
// (images, then rows, then columns), so the first half of the box is the first 64 rows of the shared-memory tile.
int make_a_half_maps(const cpd_gemm_params* p, const ConvGeom& g, CUtensorMap* map_a0h, CUtensorMap* map_a1h, int* half_x,
                     int* half_y, int* half_n) {
  if (g.nbox != 1) return CPD_ERR_UNSUPPORTED;
  int tw = g.tw, th = g.th, nb = g.nb;
  *half_x = *half_y = *half_n = 0;
  if (nb > 1 && nb % 2 == 0) {
    nb /= 2;
    *half_n = nb;
  } else if (nb == 1 && th > 1 && th % 2 == 0) {
    th /= 2;
    *half_y = th;
  } else if (nb == 1 && th == 1 && tw % 2 == 0) {
    tw /= 2;
    *half_x = tw;
  } else {
    return CPD_ERR_UNSUPPORTED;
  }
  if ((tw * th * nb) % 8 != 0) return CPD_ERR_UNSUPPORTED;
  int rc = make_a_map(p, tw, th, nb, p->a0, p->c0, map_a0h);
  if (rc) return rc;
  if (p->c1 > 0) {
    rc = make_a_map(p, tw, th, nb, p->a1, p->c1, map_a1h);
    if (rc) return rc;
  } else {
    *map_a1h = *map_a0h;
  }
  return CPD_OK;
}

int make_b_map(const cpd_gemm_params* p, int taps, int box_rows, CUtensorMap* map_b) {
  const int C = p->c0 + p->c1;
  uint64_t dims[2] = {(uint64_t)taps * C, (uint64_t)p->n_out};
  uint64_t str[1] = {(uint64_t)taps * C * 2};
  uint32_t box[2] = {64, (uint32_t)box_rows};
  return cpd_make_tmap_bf16(map_b, p->wt, 2, dims, str, box);
}

}  // namespace cpd_gemm

// 1-CTA tiles (variant 1: 128 x 128, variant 2: 128 x 256), one tile per CTA.  Kept as the simple baseline the
// persistent CTA-pair kernel (gemm_umma2.cu) is checked against; variant 0 (auto) dispatches to the pair kernel.
static cpd_status gemm_conv_1cta(const cpd_gemm_params* p, void* stream) {
  CPD_REQUIRE(!p->ln_sums && !p->ln_sums_out && !p->d_t && !p->gn_sums_out,
              "cpd_gemm_conv: the one-tile kernels (variants 1, 2) have neither the folded LayerNorm, the transposed tail nor the GroupNorm statistics");
  GemmArgs args;
  ConvGeom& g = args.g;
  int m_tiles = 0;
  int rc = fill_geometry(p, &g, &args.map_a0, &args.map_a1, &m_tiles);
  if (rc) return rc;
  const int variant = p->epilogue == CPD_EPI_GEGLU ? 1 : p->variant;
  const int BN = variant == 2 ? 256 : 128;
  g.idesc = umma_idesc_f16(BM, BN, p->a_fp16 != 0, p->b_fp16 != 0);
  if (p->epilogue == CPD_EPI_GEGLU) {
    CPD_REQUIRE(p->n_out % 128 == 0 && (p->geglu_block == 0 || p->geglu_block == 128),
                "cpd_gemm_conv: the 1-CTA GEGLU epilogue needs n_out %% 128 == 0 and weights interleaved per 128 columns");
    g.n_store = p->n_out / 2;
  } else {
    g.n_store = p->n_out;
  }
  const int n_tiles = (p->n_out + BN - 1) / BN;
  rc = make_b_map(p, g.taps, BN, &args.map_b);
  if (rc) return rc;
  args.bias = p->bias;
  args.rowvec = p->rowvec;
  args.residual = reinterpret_cast<const bf16*>(p->residual);
  args.d = reinterpret_cast<bf16*>(p->d);
  cudaStream_t s = (cudaStream_t)stream;
  if (variant == 2) return launch<256, 4>(args, m_tiles, n_tiles, s);
  return launch<128, 3>(args, m_tiles, n_tiles, s);
}

// variant -> kernel (gemm_tune.cu resolves variant 0 through the per-shape table first)
cpd_status cpd_gemm_conv_dispatch(const cpd_gemm_params* p, void* stream) {
  CPD_REQUIRE(p != nullptr, "cpd_gemm_conv: null params");
  if (p->variant == 1 || p->variant == 2) return gemm_conv_1cta(p, stream);
  return cpd_gemm_conv_2cta(p, stream);
}
