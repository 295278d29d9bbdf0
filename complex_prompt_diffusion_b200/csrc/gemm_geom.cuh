// Geometry shared by the implicit-GEMM conv / GEMM kernels (gemm_umma.cu: 1-CTA tiles, gemm_umma2.cu: persistent
// 2-CTA tiles): how the 128 rows of a CTA tile map to boxes of output pixels, and the TMA tensor maps of the
// NHWC activations (rank 5: c, x, parity, y, n) and of the K-major weights.
#pragma once
#include "../../include/cpd_b200.h"
#include "common.cuh"

namespace cpd_gemm {

constexpr int BM = 128;     // output rows per CTA (one TMEM lane each)
constexpr int BK = 64;      // K elements per pipeline stage (one 128-byte swizzle atom of 16-bit elements)
constexpr int UMMA_K = 16;  // K per tcgen05.mma for 16-bit inputs

struct ConvGeom {
  int taps;        // 1 or 9
  int cb0, cb1;    // 64-channel blocks from source 0 / 1
  int c0;          // channels of source 0 (stride-2 parity offset)
  int c1;
  int stride;      // 1 or 2
  int tw, th, nb;  // box = nb images x th x tw output pixels
  int nbox;        // boxes per 128-row CTA tile
  int bx_count, by_count;
  int n_img, h_out, w_out;
  int m_valid;     // plain GEMM: valid rows; 0 = all
  int n_out;       // GEMM N
  int n_store;     // columns of D (n_out, or n_out/2 for GEGLU)
  int epilogue;
  int rowvec_stride, ld_res, ldd;
  int out_fp16;    // D / residual element type: 1 = fp16, 0 = bf16
  uint32_t idesc;  // tcgen05 instruction descriptor (encodes the A / B element formats and the MMA shape)
  // reciprocals of the divisors of the per-tile coordinate math (filled by fill_geometry; see qdiv)
  float inv_box_rows, inv_bx, inv_by, inv_img_px, inv_tw;
};

// Integer division by a launch constant through its fp32 reciprocal: ~4 instructions instead of the ~40 of an IDIV sequence
// (the TMA producer ran ten of those between two tiles: a ~1000-cycle bubble per tile in a loop that is the limiter of the
// main loop).  Exact for n + d < 2^22 (|error| <= (n + d) * 2^-23 < 0.5 / d on (n + 0.5) / d); larger values divide normally.
__device__ __forceinline__ int qdiv(int n, int d, float inv) {
  if (n + d >= (1 << 22)) return n / d;
  return (int)(((float)n + 0.5f) * inv);
}

// Row r (0..127) of CTA tile `m_tile` -> output pixel.  Returns false when the row is padding.
struct RowCoord {
  int n, y, x;
  int64_t row;  // linear output row = (n * h_out + y) * w_out + x
  bool valid;
};
__device__ __forceinline__ RowCoord row_coord(const ConvGeom& g, int m_tile, int r) {
  const int box_rows = g.tw * g.th * g.nb;
  const int j = qdiv(r, box_rows, g.inv_box_rows);
  const int pidx = r - j * box_rows;
  const int s_idx = m_tile * g.nbox + j;
  const int t2 = qdiv(s_idx, g.bx_count, g.inv_bx);
  const int bx = s_idx - t2 * g.bx_count;
  const int ng = qdiv(t2, g.by_count, g.inv_by);
  const int by = t2 - ng * g.by_count;
  const int img_px = g.tw * g.th;
  const int ib = qdiv(pidx, img_px, g.inv_img_px);
  const int p2 = pidx - ib * img_px;
  const int py2 = qdiv(p2, g.tw, g.inv_tw);
  RowCoord rc;
  rc.n = ng * g.nb + ib;
  rc.y = by * g.th + py2;
  rc.x = bx * g.tw + (p2 - py2 * g.tw);
  rc.valid = (j < g.nbox) && (rc.n < g.n_img) && (rc.y < g.h_out) && (rc.x < g.w_out);  // rows beyond the tile's boxes are padding
  rc.row = ((int64_t)rc.n * g.h_out + rc.y) * g.w_out + rc.x;
  if (g.m_valid > 0 && rc.row >= g.m_valid) rc.valid = false;
  return rc;
}

// Coordinates of the TMA box j of CTA tile m_tile for filter tap `tap` (0..8) / channel block cb.
struct BoxCoord {
  int c, x, p, y, n;
  bool src1;
};
__device__ __forceinline__ BoxCoord box_coord(const ConvGeom& g, int m_tile, int j, int tap, int cb) {
  int dy = 0, dx = 0, py = 0, px = 0;
  if (g.taps == 9) {
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
  BoxCoord b;
  b.src1 = cb >= g.cb0;
  const int csrc = b.src1 ? g.c1 : g.c0;
  b.c = (b.src1 ? cb - g.cb0 : cb) * BK + px * csrc;
  const int s_idx = m_tile * g.nbox + j;
  const int t2 = qdiv(s_idx, g.bx_count, g.inv_bx);
  const int bx = s_idx - t2 * g.bx_count;
  const int ng = qdiv(t2, g.by_count, g.inv_by);  // beyond n_img -> fully out of bounds -> zero fill
  const int by = t2 - ng * g.by_count;
  b.x = bx * g.tw + dx;
  b.p = py;
  b.y = by * g.th + dy;
  b.n = ng * g.nb;
  return b;
}

// Host: validates p, fills the geometry (everything but idesc / n_store / tile counts) and the A tensor maps.
int fill_geometry(const cpd_gemm_params* p, ConvGeom* g, CUtensorMap* map_a0, CUtensorMap* map_a1, int* m_tiles_cta);
// Host: half-box A maps for the multicast kernel (CPD_ERR_UNSUPPORTED when the box cannot be halved).
int make_a_half_maps(const cpd_gemm_params* p, const ConvGeom& g, CUtensorMap* map_a0h, CUtensorMap* map_a1h, int* half_x,
                     int* half_y, int* half_n);
// Host: weights tensor map, box = 64 x box_rows.
int make_b_map(const cpd_gemm_params* p, int taps, int box_rows, CUtensorMap* map_b);

}  // namespace cpd_gemm

// 2-CTA persistent kernel entry (gemm_umma2.cu); returns CPD_ERR_UNSUPPORTED when the shape is outside its domain.
cpd_status cpd_gemm_conv_2cta(const cpd_gemm_params* p, void* stream);
