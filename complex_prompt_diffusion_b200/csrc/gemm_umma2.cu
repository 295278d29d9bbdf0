// Persistent CTA-pair implicit-GEMM convolution / GEMM for sm_100a.
//
//   * tcgen05.mma.cta_group::2: the two CTAs of a cluster (one SM pair) compute one 256 x BN output tile.  Each
//     CTA stages its own 128 rows of A and HALF of the B tile (BN/2 weight rows), so the shared-memory operand
//     traffic per MMA is half that of a 1-CTA 128 x BN tile (which is shared-memory-bound at BN = 128).
//   * persistent: one cluster per SM pair loops over output tiles; the fp32 accumulator is double-buffered in
//     TMEM (2 x 256 columns), so the epilogue of tile i overlaps the TMA + MMA main loop of tile i + 1.
//   * epilogue through shared memory: TMEM -> registers -> bias / time-embedding / residual / GEGLU -> 64B-swizzled
//     staging tile -> TMA store (full-line coalesced, asynchronous, clipped at the tensor edge by the TMA unit).  The
//     residual tile is PREFETCHED by TMA into the same staging buffer one chunk ahead, so no thread ever issues a
//     row-strided global load or store.
//   * warp roles: warp 0 = TMA producer (both CTAs), warp 1 = MMA issuer (leader CTA only), warp 2 = TMEM
//     allocator, warps 4-11 = epilogue (EPI_GROUPS groups of four warps, one warp per TMEM lane quarter, each group taking a contiguous share of the columns); warp 3 = TMA-store issuer.
//   * BN is a RUNTIME multiple of 32 (<= 256) chosen per layer so that BN divides N (320 -> 160, 640 -> 160/128,
//     1280 -> 256/160) and the tile count fills whole waves of 74 SM pairs.
//
// Same operand addressing as gemm_umma.cu (gemm_geom.cuh): 3x3 taps are shifted TMA boxes with hardware zero fill,
// the skip concat is two tensor maps.  Replaces the cuDNN/cuBLAS calls behind nn.Conv2d / nn.Linear in
// cpd/models/unet.py:105,153-160,210,236,247 and cpd/models/attention.py:92-118,183-190,508-524.
#include <stdlib.h>

#include "gemm_geom.cuh"

namespace {
using namespace cpd_gemm;

// Two epilogue groups.  More were measured twice and are slower: round 1, 12 warps under a 128-register cap: 240 vs 229.5 ms of
// GEMM time per generation; round 2, 12 warps with setmaxnreg (40 registers for warps 0-3, 152 for the epilogue, 80 bytes of
// spills): 352.2 vs 329.2 ms per generation, every shape class slower including the main-loop-bound 3x3 convs (1191 vs 1325
// TFLOP/s) - the extra staging buffers cost a pipeline stage and the extra warps issue slots the producer / MMA warps need.
constexpr int EPI_WARPS = 8;
constexpr int EPI_GROUPS = EPI_WARPS / 4;            // one warp per TMEM lane quarter in each group
constexpr int FIRST_EPI_WARP = 4;
constexpr int NUM_THREADS2 = 32 * (FIRST_EPI_WARP + EPI_WARPS);
constexpr int ACC_COLS = 256;   // TMEM columns per accumulator stage
constexpr int MAX_STAGES = 8;
constexpr int A_BYTES = BM * BK * 2;
constexpr int NUM_SM_PAIRS = 74;

constexpr int CHUNK_COLS = 32;                     // output columns per staging chunk (64-byte rows, SWIZZLE_64B)
constexpr int CHUNK_BYTES = BM * CHUNK_COLS * 2;   // 8 KB
constexpr int STAGING_BUFS = 4;                    // per epilogue group (3 left the residual prefetch one chunk short: ~2000-cycle waits, profiles/r02_gemm_timeline.txt)
constexpr int STAGING_BYTES = EPI_GROUPS * STAGING_BUFS * CHUNK_BYTES;
constexpr int BIAS_FLOATS = 512;                   // widest tile: 2 sub-tiles x 256 columns
constexpr int BIAS_BYTES = 2 * EPI_GROUPS * 2 * BIAS_FLOATS * 4;  // bias and (folded LayerNorm) g: per group, double-buffered by tile parity

#ifdef CPD_TIMELINE
// Debug build (make TIMELINE=1): CTA 0 stamps %globaltimer at the phases of its first tile (tools/gemm_timeline.py)
__device__ unsigned long long cpd_dbg_ts[16];
#define CPD_STAMP(i)                                                                  \
  do {                                                                                \
    if (blockIdx.x == 0) {                                                            \
      unsigned long long _t;                                                          \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(_t));                          \
      cpd_dbg_ts[i] = _t;                                                             \
    }                                                                                 \
  } while (0)
// MMA issuer of CTA 0: clock64 after the full-barrier wait and after the issue of each k-iteration of its first 3 tiles
__device__ long long cpd_dbg_mma[3][16][2];
#define MMA_STAMP(ph)                                                                          \
  do {                                                                                         \
    if (blockIdx.x == 0 && lane == 0 && it < 3 && (kt - k0) < 16) cpd_dbg_mma[it][kt - k0][ph] = clock64(); \
  } while (0)
// epilogue chunk phases of the first two tiles (group 0 leader of CTA 0): [tile][chunk][phase], clock64
__device__ long long cpd_dbg_epi[2][8][8];
#define EPI_STAMP(ph)                                                                          \
  do {                                                                                         \
    if (blockIdx.x == 0 && grp == 0 && leader && it < 2 && (ch - c_lo) < 8) cpd_dbg_epi[it][ch - c_lo][ph] = clock64(); \
  } while (0)
#else
#define CPD_STAMP(i) \
  do {               \
  } while (0)
#define EPI_STAMP(ph) \
  do {                \
  } while (0)
#define MMA_STAMP(ph) \
  do {                \
  } while (0)
#endif

struct Gemm2Args {
  CUtensorMap map_a0, map_a1, map_b;
  CUtensorMap map_d, map_res;  // (c, x, y, n) views of the output / residual, box = 32 columns x one pixel box
  CUtensorMap map_a0h, map_a1h;  // MC = 2: half-box views of the activations
  int half_x, half_y, half_n;    // MC = 2: coordinate offset of the second half box
  ConvGeom g;
  const float* bias;
  const float* rowvec;
  const bf16* residual;
  bf16* d;
  int bn;           // MMA N (columns per sub-tile)
  int nsub;         // sub-tiles per output tile (1 or 2): the tile is 256 x (nsub * bn); nsub * bn > 256 leaves room for only
                    // ONE accumulator stage in TMEM (no epilogue / main-loop overlap) but halves the A bytes per flop
  int stages;
  int stage_bytes;  // A_BYTES + nsub * bn/2 * 128
  int m_tiles2;     // 256-row tiles
  int n_tiles;
  int num_k;        // taps * 64-channel blocks
  int splits;       // split-K: each output tile is computed by `splits` work units over disjoint K ranges; unit s stores
                    // its fp32 partial tile into slice s of `ws` and a second kernel sums the slices in a fixed order
                    // (no atomics: bit-reproducible) and applies the epilogue
  float* ws;        // [splits][rows][n_store] fp32
  int64_t ws_slice; // rows * n_store
  int split_prod;   // 1: warp 2 issues the weight (B) loads, warp 0 only the activation (A) loads (CPD_GEMM_SPLIT_PROD)
  // LayerNorm folded into the GEMM (cpd_gemm_params.ln_*): consumer side (ln_in) and producer side (ln_out)
  const float2* ln_in;
  int ln_parts;
  int64_t ln_ld;
  const float* ln_g;
  float ln_inv_c, ln_eps;
  float2* ln_out;
  // transposed tail: chunks whose first column is >= dt_col0 go to d_t through map_dt (128 rows x 32 columns -> 32 x 128)
  CUtensorMap map_dt;
  int dt_col0;      // < 0: off
  long long* gn_out;  // [n_img][n_out][2] fixed-point sum / sum of squares of D (cpd_gemm_params.gn_sums_out)
};

// MC = 1: cluster = one CTA pair.  MC = 2: cluster = two CTA pairs working on the same 256 rows and adjacent column
// tiles; each CTA loads only HALF of its 128-row A tile and TMA-multicasts it to the same-rank CTA of the other pair,
// halving the L2 -> SM traffic of the A operand (the main loop is L2-bandwidth-bound).
// OF16: the element type of D / the residual (fp16 or bf16) as a compile-time constant - with a runtime flag every F2FP of
// the epilogue was emitted twice under complementary predicates.
// FEAT: epilogue features as compile-time bits (the plain instantiation must not pay for them: with runtime flags the
// epilogue-bound K = 320 linears lost 17 %): 1 = folded-LayerNorm consumer, 2 = folded-LayerNorm producer, 4 = transposed tail,
// 8 = GroupNorm statistics of D (per image and channel, fixed-point integer atomics).
constexpr int FEAT_LN_IN = 1, FEAT_LN_OUT = 2, FEAT_DT = 4, FEAT_GN_OUT = 8;
template <int MC, bool OF16, int FEAT>
__global__ void __cluster_dims__(2 * MC, 1, 1) __launch_bounds__(NUM_THREADS2, 1)
    gemm2_kernel(const __grid_constant__ Gemm2Args args) {
  constexpr bool ln_in = (FEAT & FEAT_LN_IN) != 0, ln_out = (FEAT & FEAT_LN_OUT) != 0, has_dt = (FEAT & FEAT_DT) != 0;
  constexpr bool gn_out = (FEAT & FEAT_GN_OUT) != 0;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int stages = args.stages;
  const int stage_bytes = args.stage_bytes;
  uint8_t* staging = smem + stages * stage_bytes;
  float* bias_s = reinterpret_cast<float*>(staging + STAGING_BYTES);  // [EPI_GROUPS][2][BIAS_FLOATS]
  float* g_s = bias_s + EPI_GROUPS * 2 * BIAS_FLOATS;                 // same shape: g of a folded LayerNorm
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + STAGING_BYTES + BIAS_BYTES);
  uint64_t* empty_bar = full_bar + MAX_STAGES;
  uint64_t* tmem_full = empty_bar + MAX_STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* res_bar = tmem_empty + 2;  // [EPI_GROUPS][STAGING_BUFS]  residual tile landed in the staging buffer (TMA tx)
  uint64_t* chunk_ready = res_bar + EPI_GROUPS * STAGING_BUFS;  // [groups][bufs]  128 epilogue threads have written the chunk
  uint64_t* buf_free = chunk_ready + EPI_GROUPS * STAGING_BUFS;  // [groups][bufs]  the TMA store has finished reading the buffer
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(buf_free + EPI_GROUPS * STAGING_BUFS);

  const ConvGeom& g = args.g;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();        // rank in the cluster (0 .. 2 * MC - 1)
  const uint32_t rank = crank & 1;                 // rank in the CTA pair (0 = leader: issues the MMAs)
  const int pair = (int)(crank >> 1);              // pair in the cluster
  const uint32_t leader_crank = crank & ~1u;
  const int cluster_id = blockIdx.x / (2 * MC);
  const int num_clusters = gridDim.x / (2 * MC);
  const int n_tiles_c = args.n_tiles / MC;         // column tiles per cluster step (host guarantees divisibility)
  const int splits = args.splits;
  const int total_tiles = args.m_tiles2 * n_tiles_c * splits;  // work units: (row block, column tile(s), K range)
  // Work units are dealt round-robin: the 74 clusters work on 74 neighbouring tiles at any time.  (A CONTIGUOUS range per cluster -
  // consecutive tiles sharing their 256 rows, so that per-row epilogue state is loaded once per row block - was measured and is
  // slower: 65536 x 768 x 320 61-65 -> 76 us, 65536 x 320 x 320 29.6 -> 31.4 us on the same box.)
  const int t_begin = cluster_id, t_end = total_tiles, t_step = num_clusters;
  // (integer divisions cost ~50 cycles each, 64-bit ones several hundred: the common splits == 1 / single-column-tile cases skip them)
  const float inv_splits = 1.0f / (float)splits, inv_ntc = 1.0f / (float)n_tiles_c;
  auto tile_mn = [&](int u, int& m2, int& n_tile) {
    const int t = splits == 1 ? u : qdiv(u, splits, inv_splits);
    m2 = n_tiles_c == 1 ? t : qdiv(t, n_tiles_c, inv_ntc);
    n_tile = (t - m2 * n_tiles_c) * MC + pair;
  };
  auto k_range = [&](int u, int& k0, int& k1) {
    if (splits == 1) {
      k0 = 0;
      k1 = args.num_k;
      return;
    }
    const unsigned sp = (unsigned)(u % splits);  // sp * num_k < 2^31: num_k <= a few thousand 64-channel blocks
    k0 = (int)(sp * (unsigned)args.num_k / (unsigned)splits);
    k1 = (int)((sp + 1u) * (unsigned)args.num_k / (unsigned)splits);
  };
  const int cbt = g.cb0 + g.cb1;
  const bool two_acc = args.bn * args.nsub <= ACC_COLS;  // room for two accumulator stages in the 512 TMEM columns

  if (threadIdx.x == 0) CPD_STAMP(0);
  pdl_launch_dependents();
  cluster_sync_all();  // both CTAs of the pair are resident before the pair-wide TMEM allocation
  if (threadIdx.x == 0) CPD_STAMP(1);
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&args.map_a0);
    tma_prefetch_desc(&args.map_b);
    if (g.cb1 > 0) tma_prefetch_desc(&args.map_a1);
    tma_prefetch_desc(&args.map_d);
    if (args.residual) tma_prefetch_desc(&args.map_res);
    if (has_dt) tma_prefetch_desc(&args.map_dt);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < EPI_GROUPS * STAGING_BUFS; ++s) {
      mbar_init(&res_bar[s], 1);
      mbar_init(&chunk_ready[s], 128);
      mbar_init(&buf_free[s], 1);
    }
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full_bar[s], 2);   // leader: its own arrive.expect_tx + the peer's remote arrive
      mbar_init(&empty_bar[s], MC);  // one tcgen05.commit (multicast to the whole cluster) per pair that reads it
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
  pdl_wait();  // everything above overlapped the previous kernel's tail; global memory is touched only below
  if (threadIdx.x == 0) CPD_STAMP(2);

  // warps 0-3: TMA producer, MMA issuer, TMEM allocator, store warp; warps 4-11: epilogue
  if (warp < FIRST_EPI_WARP) {
  if (warp == 0) {
    // ================= TMA producer (both CTAs) =================
    // One thread; the loop body must stay far below the MMA time of a stage (BN/2 * 4 cycles), so everything that
    // needs a division is hoisted to once per tile and the (tap, channel-block) walk is incremental.
    {
      const int box_bytes = g.tw * g.th * g.nb * 128;
      // both CTAs' bytes land on the leader's barrier; the boxes of a tile may cover fewer than 128 rows (gemm_umma.cu: choose_box)
      const uint32_t tx_bytes = 2u * (uint32_t)(stage_bytes - A_BYTES + g.nbox * box_bytes);
      const int bn_half = args.bn >> 1;
      const int tile_w = args.bn * args.nsub;
      int stage = 0;
      uint32_t phase = 0;
      const uint16_t mc_mask = (uint16_t)(0x5u << rank);  // same-rank CTAs of both pairs
      if (lane == 0) CPD_STAMP(13);
      for (int t = t_begin; t < t_end; t += t_step) {
        int m2, n_tile;
        tile_mn(t, m2, n_tile);
        const int m_tile = m2 * 2 + (int)rank;
        const int b_row = n_tile * tile_w + (int)rank * bn_half;
        // box 0 of this tile at tap (0, 0) (the only box when nbox == 1, the common case)
        const BoxCoord b0 = box_coord(g, m_tile, 0, 4, 0);  // tap 4 = centre: no shift
        int k0, k1;
        k_range(t, k0, k1);
        int tap = k0 ? k0 / cbt : 0, cb = k0 - tap * cbt;  // k0 == 0 unless split-K
        int dy = 0, dx = 0, py = 0, px = 0;
        if (g.taps == 9) {
          const int ky = tap / 3, kx = tap - ky * 3;
          if (g.stride == 1) {
            dy = ky - 1;
            dx = kx - 1;
          } else {
            dy = (ky == 0) ? -1 : 0;
            py = (ky == 0) ? 1 : ky - 1;
            dx = (kx == 0) ? -1 : 0;
            px = (kx == 0) ? 1 : kx - 1;
          }
        }
        if (t == t_begin && lane == 0) CPD_STAMP(11);
        for (int kt = k0; kt < k1; ++kt) {
          mbar_wait(&empty_bar[stage], phase ^ 1, 1);
          if (kt == k0 && t == t_begin && lane == 0) CPD_STAMP(12);
          uint8_t* sa = smem + stage * stage_bytes;
          const bool src1 = cb >= g.cb0;
          const CUtensorMap* ma = src1 ? &args.map_a1 : &args.map_a0;
          const int cc = (src1 ? (cb - g.cb0) : cb) * BK + px * (src1 ? g.c1 : g.c0);
          if (elect_one()) {
            if (rank == 0)
              mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
            else
              mbar_arrive_cluster(&full_bar[stage], leader_crank);
            if (MC == 2) {
              // this CTA's half of the box (split along the box's outermost non-unit dimension), multicast to both pairs
              const CUtensorMap* mh = src1 ? &args.map_a1h : &args.map_a0h;
              tma_load_5d_pair_mc(sa + pair * (A_BYTES / 2), mh, &full_bar[stage], cc, b0.x + dx + pair * args.half_x, py,
                                  b0.y + dy + pair * args.half_y, b0.n + pair * args.half_n, mc_mask);
            } else if (g.nbox == 1) {
              tma_load_5d_pair(sa, ma, &full_bar[stage], cc, b0.x + dx, py, b0.y + dy, b0.n);
            } else {
              for (int j = 0; j < g.nbox; ++j) {
                const BoxCoord bc = box_coord(g, m_tile, j, 4, 0);
                tma_load_5d_pair(sa + j * box_bytes, ma, &full_bar[stage], cc, bc.x + dx, py, bc.y + dy, bc.n);
              }
            }
            if (!args.split_prod)
              for (int sub = 0; sub < args.nsub; ++sub)  // this CTA's half of every sub-tile's weight rows
                tma_load_2d_pair(sa + A_BYTES + sub * (bn_half * 128), &args.map_b, &full_bar[stage], kt * BK, b_row + sub * args.bn);
          }
          __syncwarp();
          if (kt == k0 && t == t_begin && lane == 0) CPD_STAMP(3);
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
  } else if (warp == 2 && args.split_prod) {
    // ================= weight producer (both CTAs) =================
    // Opt-in experiment (CPD_GEMM_SPLIT_PROD=1).  The same stage walk as warp 0, issuing only the B loads, to test whether
    // the producer's issue slots are what the main loop waits for while the epilogue warps run
    // (profiles/r01_gemm2_timeline_v2.txt).  Measured: 361.9 ms per generation with it, 361.5 ms without - the producer
    // is NOT the limiter, so the default stays one producer warp.  The bytes land on the same full barrier; its expect_tx (warp 0) may arrive after these
    // complete_tx - the phase cannot complete before both pending arrivals, so a transiently negative tx-count is fine.
    const int bn_half = args.bn >> 1;
    const int tile_w = args.bn * args.nsub;
    int stage = 0;
    uint32_t phase = 0;
    for (int t = t_begin; t < t_end; t += t_step) {
      int m2, n_tile;
      tile_mn(t, m2, n_tile);
      const int b_row = n_tile * tile_w + (int)rank * bn_half;
      int k0, k1;
      k_range(t, k0, k1);
      for (int kt = k0; kt < k1; ++kt) {
        mbar_wait(&empty_bar[stage], phase ^ 1, 1);
        uint8_t* sb = smem + stage * stage_bytes + A_BYTES;
        if (elect_one())
          for (int sub = 0; sub < args.nsub; ++sub)
            tma_load_2d_pair(sb + sub * (bn_half * 128), &args.map_b, &full_bar[stage], kt * BK, b_row + sub * args.bn);
        __syncwarp();
        if (++stage == stages) {
          stage = 0;
          phase ^= 1;
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
      const uint64_t sub_step = (uint64_t)((args.bn >> 1) * 128 >> 4);  // bn/2 weight rows of 128 bytes
      int stage = 0;
      uint32_t phase = 0;
      uint64_t da = desc0;
      int it = 0;
      for (int t = t_begin; t < t_end; t += t_step, ++it) {
        const int acc = two_acc ? (it & 1) : 0;
        const uint32_t acc_phase = (uint32_t)((two_acc ? (it >> 1) : it) & 1);
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1, 4);  // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * ACC_COLS);
        uint32_t accum = 0;
        int k0, k1;
        k_range(t, k0, k1);
        for (int kt = k0; kt < k1; ++kt) {
          mbar_wait(&full_bar[stage], phase, 2);
          tc_fence_after();
          MMA_STAMP(0);
          if (kt == k0 && t == t_begin && lane == 0) CPD_STAMP(4);
          const uint64_t db = da + b_off;
          if (elect_one()) {
            // +2 in the address field = 32 bytes = 16 elements along K inside the 128-byte swizzle atom
            umma_f16_pair(d_tmem, da, db, idesc, accum);
            umma_f16_pair(d_tmem, da + 2, db + 2, idesc, 1u);
            umma_f16_pair(d_tmem, da + 4, db + 4, idesc, 1u);
            umma_f16_pair(d_tmem, da + 6, db + 6, idesc, 1u);
            if (args.nsub == 2) {  // second sub-tile: same A, next bn/2 weight rows, next bn accumulator columns
              const uint64_t db2 = db + sub_step;
              const uint32_t d2 = d_tmem + (uint32_t)args.bn;
              umma_f16_pair(d2, da, db2, idesc, accum);
              umma_f16_pair(d2, da + 2, db2 + 2, idesc, 1u);
              umma_f16_pair(d2, da + 4, db2 + 4, idesc, 1u);
              umma_f16_pair(d2, da + 6, db2 + 6, idesc, 1u);
            }
            umma_commit_pair(&empty_bar[stage], (uint16_t)((1u << (2 * MC)) - 1));  // frees the stage cluster-wide
          }
          __syncwarp();
          MMA_STAMP(1);
          accum = 1u;
          da += stage_step;
          if (++stage == stages) {
            stage = 0;
            phase ^= 1;
            da = desc0;
          }
        }
        if (elect_one()) umma_commit_pair(&tmem_full[acc], (uint16_t)(3u << (2 * pair)));  // accumulator complete, both CTAs of the pair
        __syncwarp();
        if (t == t_begin && lane == 0) CPD_STAMP(5);
      }
    }
  } else if (warp == 3 && lane < EPI_GROUPS && splits == 1) {
    // ================= store warp: lane g serves epilogue group g =================
    // The TMA store of a finished chunk (~650 cycles to issue), the wait for older stores to release their staging
    // buffers and the residual prefetch used to sit in the epilogue leader's path behind a 128-thread barrier, i.e. on
    // the critical path of every chunk (2200 cycles per 128 x 32 chunk, more than the main loop of any layer with
    // K <= 1280).  Here the 128 epilogue threads only arrive on chunk_ready and go on; this thread issues the store,
    // frees the buffer of the previous chunk once it has been read (buf_free, or by loading the next residual tile into
    // it) and keeps STAGING_BUFS residual tiles in flight.
    const int grp = lane;
    const bool geglu = g.epilogue == CPD_EPI_GEGLU;
    const int bn = args.bn;
    const int out_w = geglu ? (bn >> 1) : bn * args.nsub;
    const int nch = out_w / CHUNK_COLS;
    const int c_lo = grp * (nch / EPI_GROUPS) + min(grp, nch % EPI_GROUPS);  // contiguous, sizes differ by at most one
    const int c_hi = c_lo + nch / EPI_GROUPS + (grp < nch % EPI_GROUPS ? 1 : 0);
    const int box_rows = g.tw * g.th * g.nb;
    uint8_t* my_staging = staging + grp * STAGING_BUFS * CHUNK_BYTES;
    uint64_t* my_res_bar = res_bar + grp * STAGING_BUFS;
    uint64_t* my_ready = chunk_ready + grp * STAGING_BUFS;
    uint64_t* my_free = buf_free + grp * STAGING_BUFS;
    const bool has_res = args.residual != nullptr;
    if (c_lo < c_hi) {
      auto issue_residual = [&](int t, int ch, int buf) {
        int m2, n_tile;
        tile_mn(t, m2, n_tile);
        const int col0 = n_tile * out_w + ch * CHUNK_COLS;
        mbar_arrive_expect_tx(&my_res_bar[buf], (uint32_t)(g.nbox * box_rows * (CHUNK_COLS * 2)));
        for (int j = 0; j < g.nbox; ++j) {
          const BoxCoord bc = box_coord(g, m2 * 2 + (int)rank, j, 4, 0);
          tma_load_4d(my_staging + buf * CHUNK_BYTES + j * box_rows * (CHUNK_COLS * 2), &args.map_res, &my_res_bar[buf], col0, bc.x,
                      bc.y, bc.n);
        }
      };
      int t_r = t_begin, ch_r = c_lo, kc_r = 0;  // residual prefetch cursor
      auto advance_r = [&]() {
        ++kc_r;
        if (++ch_r == c_hi) {
          ch_r = c_lo;
          t_r += t_step;
        }
      };
      if (has_res)
        for (int i = 0; i < STAGING_BUFS && t_r < t_end; ++i) {
          issue_residual(t_r, ch_r, kc_r % STAGING_BUFS);
          advance_r();
        }
      int kc = 0;
      for (int t = t_begin; t < t_end; t += t_step) {
        int m2, n_tile;
        tile_mn(t, m2, n_tile);
        const int m_tile = m2 * 2 + (int)rank;
        const BoxCoord bc0 = box_coord(g, m_tile, 0, 4, 0);  // once per tile: the chunks differ only in their column
        for (int ch = c_lo; ch < c_hi; ++ch, ++kc) {
          const int buf = kc % STAGING_BUFS;
          const uint8_t* sbuf = my_staging + buf * CHUNK_BYTES;
          const int col0 = n_tile * out_w + ch * CHUNK_COLS;
          mbar_wait(&my_ready[buf], (uint32_t)((kc / STAGING_BUFS) & 1), 7);
          if (has_dt && col0 >= args.dt_col0) {
            // two [32 columns][64 rows] halves -> d_t[column][row].  Row = m_tile * 128 (plain GEMM: one 128-row box per CTA tile),
            // NOT bc0.x: for a CTA tile beyond the last row box_coord wraps x to 0 and moves the tile into the (out-of-bounds)
            // image coordinate, which this 2-D map does not have.
            tma_store_2d(&args.map_dt, sbuf, m_tile * BM, col0 - args.dt_col0);
            tma_store_2d(&args.map_dt, sbuf + CHUNK_BYTES / 2, m_tile * BM + 64, col0 - args.dt_col0);
          } else if (col0 < g.n_store) {
            tma_store_4d(&args.map_d, sbuf, col0, bc0.x, bc0.y, bc0.n);
            for (int j = 1; j < g.nbox; ++j) {
              const BoxCoord bc = box_coord(g, m_tile, j, 4, 0);
              tma_store_4d(&args.map_d, sbuf + j * box_rows * (CHUNK_COLS * 2), col0, bc.x, bc.y, bc.n);
            }
          }
          bulk_commit_group();
          if (kc >= 1) {
            bulk_wait_group_read<1>();  // every store but the one just issued has been read: chunk kc - 1's buffer is free
            const int prev = (kc - 1) % STAGING_BUFS;
            if (has_res) {
              if (t_r < t_end) {
                issue_residual(t_r, ch_r, prev);  // kc_r == kc + STAGING_BUFS - 1: this IS the buffer that chunk will use
                advance_r();
              }
            } else {
              mbar_arrive(&my_free[prev]);
            }
          }
        }
      }
      // the staging buffers must outlive the TMA engine's READS of them; the global writes themselves are complete (and
      // visible to the next kernel) at grid completion like any other store
      bulk_wait_group_read<0>();
    }
  }
  } else {
    // ================= epilogue (warps 4..11 of both CTAs) =================
    // EPI_GROUPS groups of 4 warps (one warp per TMEM lane quarter); group `grp` owns a contiguous share of the tile's
    // 32-column chunks.  Per chunk: TMEM load -> epilogue math -> wait until the staging buffer is free (or, with a
    // residual, until its tile has landed there) -> swizzled st.shared -> arrive on chunk_ready.  The store warp (warp 3)
    // issues the TMA store, recycles the buffers and prefetches the residual tiles.
    constexpr bool of16 = OF16;
    const int q = warp & 3;                         // TMEM lane quarter this warp may access
    const int grp = (warp - FIRST_EPI_WARP) >> 2;   // epilogue group
    const int r = q * 32 + lane;                    // row within this CTA's 128-row tile
    const bool leader = (q == 0 && lane == 0);
    const int bn = args.bn;
    const bool geglu = g.epilogue == CPD_EPI_GEGLU;
    const int out_w = geglu ? (bn >> 1) : bn * args.nsub;  // output columns per tile
    const int nch = out_w / CHUNK_COLS;
    const int c_lo = grp * (nch / EPI_GROUPS) + min(grp, nch % EPI_GROUPS);  // contiguous, sizes differ by at most one
    const int c_hi = c_lo + nch / EPI_GROUPS + (grp < nch % EPI_GROUPS ? 1 : 0);
    uint8_t* my_staging = staging + grp * STAGING_BUFS * CHUNK_BYTES;
    uint64_t* my_res_bar = res_bar + grp * STAGING_BUFS;
    const bool has_res = args.residual != nullptr && splits == 1;
    const uint32_t sw = (uint32_t)((r >> 1) & 3);   // SWIZZLE_64B: 16-byte chunk index ^= (row / 2) % 4
    uint8_t* my_row = nullptr;                      // set per buffer

    int kc = 0;  // chunks processed by this group so far (staging buffer = kc % 3, residual barrier parity = (kc / 3) & 1)
    int it = 0;
    float ln_rstd = 1.f, ln_nm = 0.f;  // folded LayerNorm: this thread's row of the current row block
    int ln_m2 = -1;
    float2 ln_pf0 = make_float2(0.f, 0.f), ln_pf1 = ln_pf0, ln_pf2 = ln_pf0, ln_pf3 = ln_pf0;  // prefetched parts of the next tile's row
    bool ln_pf_valid = false;
    for (int t = t_begin; t < t_end; t += t_step, ++it) {
      int m2, n_tile;
      tile_mn(t, m2, n_tile);
      const int m_tile = m2 * 2 + (int)rank;
      const RowCoord rc = row_coord(g, m_tile, r);
      const int n_img_row = rc.n < g.n_img ? rc.n : g.n_img - 1;
      const float* rv = args.rowvec ? args.rowvec + (int64_t)n_img_row * g.rowvec_stride : nullptr;
      // This tile's bias goes to shared memory while the main loop runs (one L2 round trip per tile instead of one per
      // 32-column chunk: prefetch.global.L1 did not keep the per-chunk __ldg from missing).  Double-buffered by tile parity:
      // with one group barrier per tile a warp is at most one tile ahead of the slowest reader.
      const int bias_w = geglu ? bn : out_w;  // floats of bias per tile
      float* my_bias = bias_s + (grp * 2 + (it & 1)) * BIAS_FLOATS;
      float* my_g = g_s + (grp * 2 + (it & 1)) * BIAS_FLOATS;
      // folded LayerNorm, consumer side: mean / rstd of this thread's row from the producer's partial sums (summed in part order).
      // The epilogue is the busy stage of the K <= 1280 linears, so an L2 round trip at the top of every tile is exposed
      // (measured: +20 % on 65536 x 1152 x 320): the first four parts of the NEXT tile's row are requested here, one tile ahead.
      if (ln_in && m2 != ln_m2) {
        ln_m2 = m2;
        float s1 = 0.f, s2 = 0.f;
        if (rc.valid) {
          const float2* lp = args.ln_in + rc.row;
          const float2 z = make_float2(0.f, 0.f);
          if (!ln_pf_valid) {  // first tile of this CTA
            ln_pf0 = __ldg(lp);
            ln_pf1 = args.ln_parts > 1 ? __ldg(lp + args.ln_ld) : z;
            ln_pf2 = args.ln_parts > 2 ? __ldg(lp + 2 * args.ln_ld) : z;
            ln_pf3 = args.ln_parts > 3 ? __ldg(lp + 3 * args.ln_ld) : z;
          }
          s1 = ((ln_pf0.x + ln_pf1.x) + ln_pf2.x) + ln_pf3.x;
          s2 = ((ln_pf0.y + ln_pf1.y) + ln_pf2.y) + ln_pf3.y;
          for (int pi = 4; pi < args.ln_parts; pi += 4) {  // wide rows (N >= 640 producers): four loads in flight at a time
            const float2 v0 = __ldg(lp + (int64_t)pi * args.ln_ld);
            const float2 v1 = pi + 1 < args.ln_parts ? __ldg(lp + (int64_t)(pi + 1) * args.ln_ld) : z;
            const float2 v2 = pi + 2 < args.ln_parts ? __ldg(lp + (int64_t)(pi + 2) * args.ln_ld) : z;
            const float2 v3 = pi + 3 < args.ln_parts ? __ldg(lp + (int64_t)(pi + 3) * args.ln_ld) : z;
            s1 += v0.x; s2 += v0.y;
            s1 += v1.x; s2 += v1.y;
            s1 += v2.x; s2 += v2.y;
            s1 += v3.x; s2 += v3.y;
          }
        }
        const float mean = s1 * args.ln_inv_c;
        const float var = fmaxf(fmaf(s2, args.ln_inv_c, -mean * mean), 0.f);
        ln_rstd = rsqrtf(var + args.ln_eps);
        ln_nm = -mean * ln_rstd;
      }
      if (ln_in) {
        ln_pf_valid = false;
        const int t_next = t + t_step;
        if (t_next < t_end) {
          int m2n, ntn;
          tile_mn(t_next, m2n, ntn);
          if (m2n != m2) {
            const RowCoord rn = row_coord(g, m2n * 2 + (int)rank, r);
            ln_pf_valid = true;  // (rows beyond the valid range: zeros, never used)
            const float2 z = make_float2(0.f, 0.f);
            const float2* lp = args.ln_in + rn.row;
            ln_pf0 = rn.valid ? __ldg(lp) : z;
            ln_pf1 = rn.valid && args.ln_parts > 1 ? __ldg(lp + args.ln_ld) : z;
            ln_pf2 = rn.valid && args.ln_parts > 2 ? __ldg(lp + 2 * args.ln_ld) : z;
            ln_pf3 = rn.valid && args.ln_parts > 3 ? __ldg(lp + 3 * args.ln_ld) : z;
          }
        }
      }
      float ln_s1 = 0.f, ln_s2 = 0.f;  // producer side: this thread's row over this group's columns of the tile
      if (args.bias) {
        const int col_first = n_tile * bias_w;
        const int idx = r * 4;
        if (idx < bias_w) {
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (col_first + idx + 3 < g.n_out) {
            v = __ldg(reinterpret_cast<const float4*>(args.bias + col_first + idx));
          } else {
            if (col_first + idx < g.n_out) v.x = __ldg(args.bias + col_first + idx);
            if (col_first + idx + 1 < g.n_out) v.y = __ldg(args.bias + col_first + idx + 1);
            if (col_first + idx + 2 < g.n_out) v.z = __ldg(args.bias + col_first + idx + 2);
          }
          sts128f(smem_u32(my_bias + idx), v);
          if (ln_in) {  // g[n] of the folded LayerNorm, same columns (n_out % 4 == 0 for folded launches)
            float4 gv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (col_first + idx + 3 < g.n_out) gv = __ldg(reinterpret_cast<const float4*>(args.ln_g + col_first + idx));
            sts128f(smem_u32(my_g + idx), gv);
          }
        }
        if (rv && (lane < ((bias_w + 31) >> 5)) && col_first + lane * 32 < g.n_out)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(rv + col_first + lane * 32));
        named_bar_sync(1 + grp, 128);
      }
      const int acc = two_acc ? (it & 1) : 0;
      const uint32_t acc_phase = (uint32_t)((two_acc ? (it >> 1) : it) & 1);
      mbar_wait(&tmem_full[acc], acc_phase, 3);
      tc_fence_after();
      if (t == t_begin && leader && grp == 0) CPD_STAMP(6);
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * ACC_COLS);

      if (splits > 1) {
        // split-K: store this unit's fp32 partial tile in its workspace slice; bias / residual / conversion happen in the
        // finalize kernel once every K range has been stored.  A thread owns a ROW of the tile: stored directly its 32 columns
        // are eight 16-byte pieces at a row stride (32 different lines per store instruction: 8.6 us of a 42 us launch in the
        // timeline of 1024 x 1280 x 11520).  So the warp's 32 x 32 block goes through the group's staging buffers - idle in
        // split-K launches, the store warp does not run - swizzled like the 16-bit chunks, and leaves as whole 128-byte lines:
        // instruction i writes columns 4 (lane % 8) .. + 3 of rows 4 i + lane / 8.
        const uint32_t wbuf = smem_u32(my_staging) + (uint32_t)q * 4096u;  // 32 rows x 128 bytes per warp (4 warps: 2 staging buffers)
        int rows_i[8];  // workspace row of this lane's row in store instruction i (-1: padding row)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int src = 4 * i + (lane >> 3);
          const int rr = __shfl_sync(0xffffffffu, (int)rc.row, src);
          const int vv = __shfl_sync(0xffffffffu, (int)rc.valid, src);
          rows_i[i] = vv ? rr : -1;
        }
        const int c16 = lane & 7;
        for (int ch = c_lo; ch < c_hi; ++ch) {
          uint32_t v[32];
          tmem_ld32(taddr + ch * CHUNK_COLS, v);
          tmem_ld_wait();
          const int col0 = n_tile * out_w + ch * CHUNK_COLS;
#pragma unroll
          for (int c = 0; c < 8; ++c)
            sts128(wbuf + (uint32_t)lane * 128u + (uint32_t)((c ^ (lane & 7)) << 4), make_uint4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]));
          __syncwarp();
          float* wbase = args.ws + (int64_t)(t % splits) * args.ws_slice + col0 + c16 * 4;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rl = 4 * i + (lane >> 3);
            const uint4 x = lds128(wbuf + (uint32_t)rl * 128u + (uint32_t)((c16 ^ (rl & 7)) << 4));
            if (rows_i[i] >= 0 && col0 + c16 * 4 < g.n_store) *reinterpret_cast<uint4*>(wbase + (int64_t)rows_i[i] * g.n_store) = x;
          }
          __syncwarp();  // the block is overwritten by the next chunk
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(&tmem_empty[acc], leader_crank);
        continue;
      }
      // (Issuing the TMEM load of chunk i + 1 before the math of chunk i - a software-pipelined epilogue - was measured on the same
      // box and is SLOWER: 27.3 vs 25.0 us on 65536 x 320 x 320, 335.8 vs 330.8 ms per generation; removed.)
      for (int ch = c_lo; ch < c_hi; ++ch, ++kc) {
        const int buf = kc % STAGING_BUFS;
        uint8_t* sbuf = my_staging + buf * CHUNK_BYTES;
        my_row = sbuf + r * (CHUNK_COLS * 2);
        EPI_STAMP(0);
        EPI_STAMP(1);
        const int col0 = n_tile * out_w + ch * CHUNK_COLS;  // first output column of this chunk
        float f[32];
        uint32_t packed[16];  // GEGLU produces packed 16-bit pairs directly
        if (geglu) {
          // tile columns [0, bn/2) = value, [bn/2, bn) = gate.  The reference rounds the projection to the model dtype,
          // applies gelu (rounded), then multiplies in the model dtype (attention.py:98-100): value and gelu(gate) are
          // packed to 16 bit and multiplied with one HMUL2 per pair.  (The gate is NOT rounded to 16 bit in front of the
          // gelu - three instructions per pair, and the fp32 value is the more exact one.)
          uint32_t va[32], vg[32];
          tmem_ld32(taddr + ch * CHUNK_COLS, va);
          tmem_ld32(taddr + (bn >> 1) + ch * CHUNK_COLS, vg);
          tmem_ld_wait();
          const float* bias_v = args.bias ? my_bias + ch * CHUNK_COLS : nullptr;  // shared memory (broadcast reads)
          if (ln_in) {  // value / gate = rstd * (acc - mean * g) + b'
            const float* g_v = my_g + ch * CHUNK_COLS;
#pragma unroll
            for (int e = 0; e < 32; e += 4) {
              const float4 bv = lds128f(smem_u32(bias_v + e)), bg = lds128f(smem_u32(bias_v + (bn >> 1) + e));
              const float4 gv = lds128f(smem_u32(g_v + e)), gg = lds128f(smem_u32(g_v + (bn >> 1) + e));
              // packed fp32 pairs (FFMA2 / FMUL2 / FADD2, sm_100): the same operations per lane - bit-identical - in about two
              // thirds of the issue slots (65536 x 2560 x 320: 139 -> 131 us).  The plain bias / residual / LayerNorm-sum epilogues
              // were measured 1-3 % SLOWER with packed arithmetic (register-pair moves) and keep scalar instructions.
              const float2 rs2 = make_float2(ln_rstd, ln_rstd), nm2 = make_float2(ln_nm, ln_nm);
              const float2 v01 = __ffma2_rn(make_float2(__uint_as_float(va[e]), __uint_as_float(va[e + 1])), rs2,
                                            __ffma2_rn(nm2, make_float2(gv.x, gv.y), make_float2(bv.x, bv.y)));
              const float2 v23 = __ffma2_rn(make_float2(__uint_as_float(va[e + 2]), __uint_as_float(va[e + 3])), rs2,
                                            __ffma2_rn(nm2, make_float2(gv.z, gv.w), make_float2(bv.z, bv.w)));
              const float2 g01 = __ffma2_rn(make_float2(__uint_as_float(vg[e]), __uint_as_float(vg[e + 1])), rs2,
                                            __ffma2_rn(nm2, make_float2(gg.x, gg.y), make_float2(bg.x, bg.y)));
              const float2 g23 = __ffma2_rn(make_float2(__uint_as_float(vg[e + 2]), __uint_as_float(vg[e + 3])), rs2,
                                            __ffma2_rn(nm2, make_float2(gg.z, gg.w), make_float2(bg.z, bg.w)));
              const float2 y01 = gelu_sig_f2(g01), y23 = gelu_sig_f2(g23);
              packed[e >> 1] = mul_act2(pack_act2(v01.x, v01.y, of16), pack_act2(y01.x, y01.y, of16), of16);
              packed[(e >> 1) + 1] = mul_act2(pack_act2(v23.x, v23.y, of16), pack_act2(y23.x, y23.y, of16), of16);
            }
          } else {
#pragma unroll
            for (int e = 0; e < 32; e += 4) {
              float4 bv = make_float4(0.f, 0.f, 0.f, 0.f), bg = bv;
              if (bias_v) {
                bv = lds128f(smem_u32(bias_v + e));
                bg = lds128f(smem_u32(bias_v + (bn >> 1) + e));
              }
              const float2 v01 = __fadd2_rn(make_float2(__uint_as_float(va[e]), __uint_as_float(va[e + 1])), make_float2(bv.x, bv.y));
              const float2 v23 = __fadd2_rn(make_float2(__uint_as_float(va[e + 2]), __uint_as_float(va[e + 3])), make_float2(bv.z, bv.w));
              const float2 g01 = __fadd2_rn(make_float2(__uint_as_float(vg[e]), __uint_as_float(vg[e + 1])), make_float2(bg.x, bg.y));
              const float2 g23 = __fadd2_rn(make_float2(__uint_as_float(vg[e + 2]), __uint_as_float(vg[e + 3])), make_float2(bg.z, bg.w));
              const float2 y01 = gelu_sig_f2(g01), y23 = gelu_sig_f2(g23);
              packed[e >> 1] = mul_act2(pack_act2(v01.x, v01.y, of16), pack_act2(y01.x, y01.y, of16), of16);
              packed[(e >> 1) + 1] = mul_act2(pack_act2(v23.x, v23.y, of16), pack_act2(y23.x, y23.y, of16), of16);
            }
          }
        } else {
          uint32_t v[32];
          tmem_ld32(taddr + ch * CHUNK_COLS, v);
          tmem_ld_wait();
          EPI_STAMP(2);
#pragma unroll
          for (int e = 0; e < 32; ++e) f[e] = __uint_as_float(v[e]);
          if (col0 + CHUNK_COLS <= g.n_store) {
            if (ln_in) {  // rstd * (acc - mean * g) + b'  (host: folded launches always carry the folded bias)
#pragma unroll
              for (int e = 0; e < 32; e += 4) {
                const float4 b4 = lds128f(smem_u32(my_bias + ch * CHUNK_COLS + e));
                const float4 g4 = lds128f(smem_u32(my_g + ch * CHUNK_COLS + e));
                f[e] = fmaf(f[e], ln_rstd, fmaf(ln_nm, g4.x, b4.x));
                f[e + 1] = fmaf(f[e + 1], ln_rstd, fmaf(ln_nm, g4.y, b4.y));
                f[e + 2] = fmaf(f[e + 2], ln_rstd, fmaf(ln_nm, g4.z, b4.z));
                f[e + 3] = fmaf(f[e + 3], ln_rstd, fmaf(ln_nm, g4.w, b4.w));
              }
            } else if (args.bias) {
#pragma unroll
              for (int e = 0; e < 32; e += 4) {
                const float4 b4 = lds128f(smem_u32(my_bias + ch * CHUNK_COLS + e));
                f[e] += b4.x; f[e + 1] += b4.y; f[e + 2] += b4.z; f[e + 3] += b4.w;
              }
            }
            if (rv) {
#pragma unroll
              for (int e = 0; e < 32; e += 4) {
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(rv + col0 + e));
                f[e] += b4.x; f[e + 1] += b4.y; f[e + 2] += b4.z; f[e + 3] += b4.w;
              }
            }
          } else {  // ragged last chunk (n_out is only a multiple of 8)
#pragma unroll
            for (int e = 0; e < 32; ++e) {
              if (col0 + e < g.n_store) {
                if (args.bias) f[e] += lds32f(smem_u32(my_bias + ch * CHUNK_COLS + e));
                if (rv) f[e] += __ldg(rv + col0 + e);
              }
            }
          }
        }
        EPI_STAMP(3);
        // the staging buffer is ours once its previous TMA store has been read out: with a residual the landed residual
        // tile implies it (the store warp loads it only into a freed buffer), otherwise the store warp says so
        if (!has_res) mbar_wait(&buf_free[grp * STAGING_BUFS + buf], (uint32_t)(((kc / STAGING_BUFS) & 1) ^ 1), 6);
        if (has_res && !geglu) {
          mbar_wait(&my_res_bar[buf], (uint32_t)((kc / STAGING_BUFS) & 1), 5);
#pragma unroll
          for (int c16 = 0; c16 < 4; ++c16) {
            const uint4 rr = lds128(smem_u32(my_row) + ((c16 ^ sw) << 4));
            const float2 r0 = unpack_act2(rr.x, of16), r1 = unpack_act2(rr.y, of16), r2 = unpack_act2(rr.z, of16),
                         r3 = unpack_act2(rr.w, of16);
            f[c16 * 8 + 0] += r0.x; f[c16 * 8 + 1] += r0.y; f[c16 * 8 + 2] += r1.x; f[c16 * 8 + 3] += r1.y;
            f[c16 * 8 + 4] += r2.x; f[c16 * 8 + 5] += r2.y; f[c16 * 8 + 6] += r3.x; f[c16 * 8 + 7] += r3.y;
          }
        }
        if (ln_out && !geglu) {  // folded LayerNorm, producer side (host: n_store % 32 == 0, so every chunk is full)
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            ln_s1 += f[e];
            ln_s2 = fmaf(f[e], f[e], ln_s2);
          }
        }
        if (!geglu) {
#pragma unroll
          for (int c16 = 0; c16 < 4; ++c16) {
            packed[c16 * 4 + 0] = pack_act2(f[c16 * 8 + 0], f[c16 * 8 + 1], of16);
            packed[c16 * 4 + 1] = pack_act2(f[c16 * 8 + 2], f[c16 * 8 + 3], of16);
            packed[c16 * 4 + 2] = pack_act2(f[c16 * 8 + 4], f[c16 * 8 + 5], of16);
            packed[c16 * 4 + 3] = pack_act2(f[c16 * 8 + 6], f[c16 * 8 + 7], of16);
          }
        }
        if (gn_out && !geglu && col0 + CHUNK_COLS <= g.n_store) {
          // GroupNorm statistics of this chunk: the warp's 32 rows x 32 columns are reduced over the ROWS by a transposing
          // butterfly (lane l ends up with column l: 31 shuffles per statistic, no shared memory, no barrier), the column sums go
          // to the per-(image, channel) accumulators as 64-bit fixed point - integer atomics are exact and order-independent.
          // f[] is dead after the packing above; the conv producers have the epilogue slack for these ~300 instructions.
          float q[32];
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            if (!rc.valid) f[e] = 0.f;
            q[e] = f[e] * f[e];
          }
#pragma unroll
          for (int sft = 16; sft >= 1; sft >>= 1) {
            const bool up = (lane & sft) != 0;
#pragma unroll
            for (int i = 0; i < sft; ++i) {
              const float send_f = up ? f[i] : f[i + sft], keep_f = up ? f[i + sft] : f[i];
              const float send_q = up ? q[i] : q[i + sft], keep_q = up ? q[i + sft] : q[i];
              f[i] = keep_f + __shfl_xor_sync(0xffffffffu, send_f, sft);
              q[i] = keep_q + __shfl_xor_sync(0xffffffffu, send_q, sft);
            }
          }
          const int n_w = __shfl_sync(0xffffffffu, rc.n, 0);  // host: the 32 rows of a warp belong to one image
          if (__shfl_sync(0xffffffffu, (int)rc.valid, 0)) {
            unsigned long long* acc = reinterpret_cast<unsigned long long*>(args.gn_out) + ((int64_t)n_w * g.n_store + col0 + lane) * 2;
            atomicAdd(acc, (unsigned long long)__float2ll_rn(f[0] * 16777216.0f));
            atomicAdd(acc + 1, (unsigned long long)__float2ll_rn(q[0] * 4096.0f));
          }
        }
        if (has_dt && col0 >= args.dt_col0) {
          // Staging = two [32 columns][64 rows] halves (4 KB each, one TMA store each), 128B-swizzled.  Lanes r, r ^ 1 exchange
          // their packed pairs so that every lane writes whole 32-bit words {row r & ~1, row r | 1} of one column: the even lane
          // column 2j, the odd lane column 2j' + 1 with j' = j ^ 2 - columns whose swizzle phases differ in bit 2, so the 32 words
          // of a store instruction fall into 32 different banks.
          const bool odd = (lane & 1) != 0;
          const int t = r & 63;                                     // row within the half
          const uint32_t half_base = smem_u32(sbuf) + (uint32_t)(r >> 6) * (CHUNK_BYTES / 2);
          const uint32_t c16 = (uint32_t)(t >> 3), in16 = (uint32_t)((t & 6) * 2);  // 16-byte piece of the row, byte inside it (word-aligned)
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const uint32_t send = odd ? packed[j] : packed[j ^ 2];
            const uint32_t got = __shfl_xor_sync(0xffffffffu, send, 1);
            // even: {mine[j].lo, partner[j].lo} -> column 2j; odd: {partner[j^2].hi, mine[j^2].hi} -> column 2(j^2) + 1
            const uint32_t word = odd ? __byte_perm(got, packed[j ^ 2], 0x7632) : __byte_perm(packed[j], got, 0x5410);
            const uint32_t col = odd ? (uint32_t)(2 * (j ^ 2) + 1) : (uint32_t)(2 * j);
            sts32(half_base + col * 128u + ((c16 ^ (col & 7u)) << 4) + in16, word);
          }
        } else {
#pragma unroll
          for (int c16 = 0; c16 < 4; ++c16)
            sts128(smem_u32(my_row) + ((c16 ^ sw) << 4), make_uint4(packed[c16 * 4], packed[c16 * 4 + 1], packed[c16 * 4 + 2], packed[c16 * 4 + 3]));
        }
        EPI_STAMP(4);
        fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the TMA (async proxy)
        EPI_STAMP(5);
        if (ch + 1 == c_hi) {
          // last TMEM read of this tile by this warp: hand the accumulator back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(&tmem_empty[acc], leader_crank);
        }
        mbar_arrive(&chunk_ready[grp * STAGING_BUFS + buf]);  // the store warp issues the TMA store: nobody waits here
        EPI_STAMP(6);
        if (t == t_begin && grp == 0 && ch == c_lo && leader) CPD_STAMP(7);
        EPI_STAMP(7);
      }
      if (ln_out && rc.valid) args.ln_out[(int64_t)(n_tile * EPI_GROUPS + grp) * args.ln_ld + rc.row] = make_float2(ln_s1, ln_s2);
      if (c_lo == c_hi) {  // this group has no chunk in the tile (single-chunk tiles): still release the accumulator
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(&tmem_empty[acc], leader_crank);
      }
    }
    if (leader && grp == 0) CPD_STAMP(8);
  }

  __syncwarp();
  tc_fence_before();
  cluster_sync_all();
  if (threadIdx.x == 0) CPD_STAMP(9);
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
    if (lane == 0) CPD_STAMP(10);
  }
}

// Split-K epilogue: d = convert(sum_s ws[s] + bias + rowvec + residual), slices summed in a fixed order.
__global__ void __launch_bounds__(256) splitk_finalize_kernel(const float* __restrict__ ws, int splits, int64_t ws_slice, int64_t rows,
                                                              int n_store, int rows_per_img,
                                                              const float* __restrict__ bias, const float* __restrict__ rowvec,
                                                              int rowvec_stride, const bf16* residual, int ld_res, bf16* d, int ldd,
                                                              int of16) {
  pdl_launch_dependents();
  pdl_wait();
  const int vec_per_row = n_store / 8;
  const int64_t total = rows * vec_per_row;
  const bool small = total < (1ll << 31);  // 32-bit index arithmetic (a 64-bit division costs ~100 instructions per item)
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = small ? (int64_t)((uint32_t)i / (uint32_t)vec_per_row) : i / vec_per_row;
    const int col = (int)(i - row * vec_per_row) * 8;
    float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const float* w0 = ws + row * n_store + col;
    int sp = 0;
    for (; sp + 4 <= splits; sp += 4) {  // four slices in flight, added in slice order (bit-identical to the one-by-one loop)
      float4 a[4][2];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float4* wp = reinterpret_cast<const float4*>(w0 + (sp + k) * ws_slice);
        a[k][0] = wp[0];
        a[k][1] = wp[1];
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        f[0] += a[k][0].x; f[1] += a[k][0].y; f[2] += a[k][0].z; f[3] += a[k][0].w;
        f[4] += a[k][1].x; f[5] += a[k][1].y; f[6] += a[k][1].z; f[7] += a[k][1].w;
      }
    }
    for (; sp < splits; ++sp) {
      const float4* wp = reinterpret_cast<const float4*>(w0 + sp * ws_slice);
      const float4 a0 = wp[0], a1 = wp[1];
      f[0] += a0.x; f[1] += a0.y; f[2] += a0.z; f[3] += a0.w;
      f[4] += a1.x; f[5] += a1.y; f[6] += a1.z; f[7] += a1.w;
    }
    if (bias) {
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] += __ldg(bias + col + e);
    }
    if (rowvec) {
      const float* rv = rowvec + (row / rows_per_img) * rowvec_stride + col;
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] += __ldg(rv + e);
    }
    if (residual) {
      const uint4 rr = *reinterpret_cast<const uint4*>(residual + row * ld_res + col);
      const float2 r0 = unpack_act2(rr.x, of16 != 0), r1 = unpack_act2(rr.y, of16 != 0), r2 = unpack_act2(rr.z, of16 != 0),
                   r3 = unpack_act2(rr.w, of16 != 0);
      f[0] += r0.x; f[1] += r0.y; f[2] += r1.x; f[3] += r1.y;
      f[4] += r2.x; f[5] += r2.y; f[6] += r3.x; f[7] += r3.y;
    }
    *reinterpret_cast<uint4*>(d + row * ldd + col) =
        make_uint4(pack_act2(f[0], f[1], of16 != 0), pack_act2(f[2], f[3], of16 != 0), pack_act2(f[4], f[5], of16 != 0),
                   pack_act2(f[6], f[7], of16 != 0));
  }
}

// Default tile shape from a cost model (the Python host additionally auto-tunes the variant per layer shape on first
// use, ops.gemm_conv).  Measured on B200 (profiles/r01_*): the main loop is bound by the bytes DELIVERED into
// each SM (~45 B/cycle/SM whatever the source - TMA multicast across two pairs did not help, B300_MICROARCH: "at
// csz <= 4, MC = UC"), so what matters is flops per byte per PAIR: a 256 x (2 x 160) tile moves 72 KB per 64-deep
// k-iteration for 640 MMA cycles where 256 x 160 moves 52 KB for 320.  Per k-iteration a tile costs
// max(MMA = 2 * bn * nsub cycles, bytes_per_pair / 90 B/cycle); a tile wider than 256 columns has a single TMEM
// accumulator stage, so its epilogue (~6.5 cycles per column) is not overlapped; plus a fixed per-tile overhead.
struct TileChoice {
  int bn, nsub, mc;
};
TileChoice pick_tiles(int m_tiles2, int n_out, int num_k, bool geglu, int geglu_block, int force_bn, int force_nsub, int force_mc,
                      bool mc2_ok) {
  static const int cand_plain[] = {256, 224, 192, 160, 128, 96, 64};
  TileChoice best = {force_bn > 0 ? force_bn : (geglu ? geglu_block : 256), force_nsub > 0 ? force_nsub : 1,
                     force_mc > 0 ? force_mc : 1};
  double best_cost = 1e30;
  for (int mc = 1; mc <= 2; ++mc) {
    if ((force_mc > 0 && mc != force_mc) || (mc == 2 && !mc2_ok)) continue;
    for (int nsub = 1; nsub <= 2; ++nsub) {
      if ((force_nsub > 0 && nsub != force_nsub) || (geglu && nsub == 2)) continue;
      for (int i = 0; i < 7; ++i) {
        const int bn = cand_plain[i];
        if ((force_bn > 0 && bn != force_bn) || (geglu && bn != geglu_block)) continue;
        const int w = bn * nsub;
        if (w > 512) continue;
        const int n_tiles = (n_out + w - 1) / w;
        if (mc == 2 && (n_tiles % 2)) continue;
        if (nsub == 2 && n_out <= bn) continue;  // the second sub-tile would be all padding
        if (nsub == 2 && num_k < 32) continue;  // short main loops cannot amortise the un-overlapped epilogue
        const long tiles = (long)m_tiles2 * n_tiles;
        const long per_wave = mc == 2 ? 66 : 74;
        const long waves = (tiles + per_wave - 1) / per_wave;
        const double bytes_pair = 2.0 * (16384.0 + 64.0 * w);
        const double mma = 2.0 * w, load = bytes_pair / 90.0;
        const double per_tile = num_k * (mma > load ? mma : load) + (w > 256 ? 6.5 * w : 0.0) + 2500.0;
        const double cost = waves * per_tile * (mc == 2 ? 1.03 : 1.0);  // no measured gain from the 4-CTA cluster
        if (cost < best_cost - 1e-9) {
          best_cost = cost;
          best.bn = bn;
          best.nsub = nsub;
          best.mc = mc;
        }
      }
    }
  }
  return best;
}

template <int MC, bool OF16, int FEAT>
cpd_status launch2(const Gemm2Args& args, int smem_bytes, cudaStream_t stream) {
  CPD_SMEM_OPTIN((gemm2_kernel<MC, OF16, FEAT>), 227 * 1024);
  const long steps = (long)args.m_tiles2 * (args.n_tiles / MC) * args.splits;
  const int max_clusters = MC == 2 ? 33 : NUM_SM_PAIRS;
  const int clusters = (int)(steps < max_clusters ? steps : max_clusters);
  CPD_CUDA_CHECK(cpd_launch(gemm2_kernel<MC, OF16, FEAT>, dim3(2 * MC * clusters), dim3(NUM_THREADS2), smem_bytes, stream, args));
  return CPD_OK;
}
template <bool OF16>
cpd_status launch2_feat(const Gemm2Args& args, int feat, int smem_bytes, cudaStream_t stream) {
  switch (feat) {
    case 0: return launch2<1, OF16, 0>(args, smem_bytes, stream);
    case FEAT_LN_IN: return launch2<1, OF16, FEAT_LN_IN>(args, smem_bytes, stream);
    case FEAT_LN_OUT: return launch2<1, OF16, FEAT_LN_OUT>(args, smem_bytes, stream);
    case FEAT_DT: return launch2<1, OF16, FEAT_DT>(args, smem_bytes, stream);
    case FEAT_LN_IN | FEAT_DT: return launch2<1, OF16, FEAT_LN_IN | FEAT_DT>(args, smem_bytes, stream);
    case FEAT_GN_OUT: return launch2<1, OF16, FEAT_GN_OUT>(args, smem_bytes, stream);
    default: break;
  }
  cpd_set_error("cpd_gemm_conv: unsupported combination of folded-LayerNorm / transposed-tail options (%d)", feat);
  return CPD_ERR_UNSUPPORTED;
}

}  // namespace

#ifdef CPD_TIMELINE
extern "C" int cpd_debug_gemm_timeline(unsigned long long* host16) {
  return (int)cudaMemcpyFromSymbol(host16, cpd_dbg_ts, sizeof(unsigned long long) * 16);
}
extern "C" int cpd_debug_gemm_mma(long long* host96) {
  return (int)cudaMemcpyFromSymbol(host96, cpd_dbg_mma, sizeof(long long) * 96);
}
extern "C" int cpd_debug_gemm_epilogue(long long* host128) {
  return (int)cudaMemcpyFromSymbol(host128, cpd_dbg_epi, sizeof(long long) * 128);
}
#endif

cpd_status cpd_gemm_conv_2cta(const cpd_gemm_params* p, void* stream) {
  Gemm2Args args;
  ConvGeom& g = args.g;
  int m_tiles_cta = 0;
  int rc = fill_geometry(p, &g, &args.map_a0, &args.map_a1, &m_tiles_cta);
  if (rc) return rc;
  const bool geglu = p->epilogue == CPD_EPI_GEGLU;
  args.m_tiles2 = (m_tiles_cta + 1) / 2;
  args.num_k = g.taps * (g.cb0 + g.cb1);
  // variant: 0 = auto; 32..256 = pair kernel, one sub-tile of BN = variant; 2000 + BN = two sub-tiles of BN (256 x 2BN tile);
  // 1000 + BN = two-pair multicast cluster with BN (1000 = auto BN)
  // + 10000 * S: split-K over S disjoint K ranges (needs p->splitk_ws)
  int force_bn = 0, force_mc = 0, force_nsub = 0;
  int variant = p->variant;
  int splits = 1;
  if (variant >= 10000) {
    splits = variant / 10000;
    variant %= 10000;
    if (variant < 32) variant = 160;  // default tile for split launches
  }
  if (variant >= 2000) {
    force_mc = 1;
    force_nsub = 2;
    force_bn = variant - 2000;
  } else if (variant >= 1000) {
    force_mc = 2;
    force_bn = variant - 1000;
  } else if (variant >= 32) {
    force_mc = 1;
    force_nsub = 1;
    force_bn = variant;
  }
  // Two round-1 experiments that measured no gain stay behind a build flag (make EXPERIMENTAL=1): the 4-CTA cluster with
  // TMA-multicast A (CPD_GEMM_MC=1 / variant 1000 + BN: equal to the CTA pair) and a second TMA producer warp for the weights
  // (CPD_GEMM_SPLIT_PROD=1: 361.9 vs 361.5 ms per generation).
  int mc_env = 0, split_prod_env = 0;
#ifdef CPD_EXPERIMENTAL
  {
    static int mc_e = -1, sp_e = -1;
    if (mc_e < 0) {
      const char* e = getenv("CPD_GEMM_MC");
      mc_e = (e && e[0] == '1') ? 1 : 0;
      e = getenv("CPD_GEMM_SPLIT_PROD");
      sp_e = (e && e[0] == '1') ? 1 : 0;
    }
    mc_env = mc_e;
    split_prod_env = sp_e;
  }
#endif
  args.split_prod = split_prod_env;
  args.half_x = args.half_y = args.half_n = 0;
  args.map_a0h = args.map_a0;
  args.map_a1h = args.map_a1;
  const bool mc2_ok = (force_mc == 2 || (force_mc == 0 && mc_env)) &&
                      make_a_half_maps(p, g, &args.map_a0h, &args.map_a1h, &args.half_x, &args.half_y, &args.half_n) == CPD_OK;
  CPD_REQUIRE(force_mc != 2 || mc2_ok, "cpd_gemm_conv: the multicast cluster needs a single box per tile that can be halved");
  const int gblock = p->geglu_block ? p->geglu_block : 128;
  const TileChoice tc = pick_tiles(args.m_tiles2, p->n_out, args.num_k, geglu, gblock, force_bn, force_nsub, force_mc, mc2_ok);
  const int bn = tc.bn;
  const int mc = tc.mc;
  const int tile_w = tc.bn * tc.nsub;
  args.nsub = tc.nsub;
  {
    static int dbg = -1;
    if (dbg < 0) dbg = getenv("CPD_GEMM_DEBUG") ? 1 : 0;
    if (dbg)
      fprintf(stderr, "cpd_gemm_conv: M-tiles(256)=%d N=%d K-iters=%d -> BN=%d nsub=%d mc=%d\n", args.m_tiles2, p->n_out, args.num_k,
              tc.bn, tc.nsub, tc.mc);
  }
  CPD_REQUIRE(bn % 32 == 0 && bn >= 32 && bn <= 256, "cpd_gemm_conv: tile width %d must be a multiple of 32 in [32, 256]", bn);
  CPD_REQUIRE(!(geglu && p->residual), "cpd_gemm_conv: the GEGLU epilogue takes no residual");
  if (geglu) {
    CPD_REQUIRE((bn == 128 || bn == 256) && bn == gblock && p->n_out % bn == 0,
                "cpd_gemm_conv: GEGLU needs geglu_block 128 or 256 dividing n_out (geglu_block=%d, n_out=%d)", gblock, p->n_out);
    g.n_store = p->n_out / 2;
  } else {
    g.n_store = p->n_out;
  }
  args.bn = bn;
  args.n_tiles = (p->n_out + tile_w - 1) / tile_w;
  CPD_REQUIRE(mc == 1 || args.n_tiles % 2 == 0, "cpd_gemm_conv: the multicast cluster needs an even number of column tiles (BN=%d)", bn);
  args.stage_bytes = A_BYTES + tc.nsub * (bn / 2) * 128;
  int stages = (227 * 1024 - 1024 - 512 - STAGING_BYTES - BIAS_BYTES) / args.stage_bytes;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  {
    static int cap = -1;  // CPD_GEMM_STAGES: pipeline-depth experiments
    if (cap < 0) {
      const char* e = getenv("CPD_GEMM_STAGES");
      cap = e ? atoi(e) : 0;
    }
    if (cap >= 2 && cap < stages) stages = cap;
  }
  args.stages = stages;
  g.idesc = umma_idesc_f16(256, bn, p->a_fp16 != 0, p->b_fp16 != 0);
  rc = make_b_map(p, g.taps, bn / 2, &args.map_b);
  if (rc) return rc;
  args.bias = p->bias;
  args.rowvec = p->rowvec;
  args.residual = reinterpret_cast<const bf16*>(p->residual);
  args.d = reinterpret_cast<bf16*>(p->d);
  const int64_t out_rows = (p->m_valid > 0 && g.h_out == 1 && g.n_img == 1) ? p->m_valid : (int64_t)g.n_img * g.h_out * g.w_out;
  if (splits > 1) {
    CPD_REQUIRE(!geglu && mc == 1, "cpd_gemm_conv: split-K supports neither the GEGLU epilogue nor the multicast cluster");
    CPD_REQUIRE(p->splitk_ws != nullptr && p->splitk_ws_floats >= (int64_t)splits * out_rows * g.n_store,
                "cpd_gemm_conv: split-K needs an fp32 workspace of %lld floats (got %lld)", (long long)(splits * out_rows * g.n_store),
                (long long)p->splitk_ws_floats);
    CPD_REQUIRE(g.n_store % 8 == 0 && splits <= args.num_k, "cpd_gemm_conv: bad split-K configuration (splits=%d, k-iterations=%d)",
                splits, args.num_k);
  }
  // LayerNorm folded into the GEMM / transposed tail (cpd_gemm_params.ln_*, d_t)
  const bool ln_any = p->ln_sums_out != nullptr || p->ln_sums != nullptr;
  CPD_REQUIRE(!(ln_any || p->d_t) || (splits == 1 && mc == 1), "cpd_gemm_conv: the folded LayerNorm / transposed tail need the plain pair kernel (no split-K, no multicast cluster)");
  args.ln_in = reinterpret_cast<const float2*>(p->ln_sums);
  args.ln_parts = p->ln_parts;
  args.ln_ld = p->ln_ld;
  args.ln_g = p->ln_g;
  args.ln_inv_c = p->ln_c > 0 ? 1.0f / (float)p->ln_c : 0.f;
  args.ln_eps = p->ln_eps;
  args.ln_out = reinterpret_cast<float2*>(p->ln_sums_out);
  if (p->ln_sums) {
    CPD_REQUIRE(p->ln_g && p->bias && p->ln_parts > 0 && p->ln_c > 0 && p->ln_ld >= out_rows && g.n_store % 32 == 0 && p->n_out % 4 == 0 && !p->rowvec,
                "cpd_gemm_conv: a folded-LayerNorm consumer needs ln_g, the folded bias, ln_parts / ln_c > 0, ln_ld >= rows, N %% 32 == 0 and no rowvec");
  }
  if (p->ln_sums_out) {
    CPD_REQUIRE(!geglu && p->ln_ld >= out_rows && g.n_store % 32 == 0, "cpd_gemm_conv: a folded-LayerNorm producer needs a plain epilogue, ln_ld >= rows and N %% 32 == 0");
    if (p->ln_parts_out) *p->ln_parts_out = args.n_tiles * EPI_GROUPS;
  }
  args.gn_out = p->gn_sums_out;
  if (p->gn_sums_out)
    CPD_REQUIRE(!geglu && splits == 1 && mc == 1 && g.n_store % 32 == 0 && (g.tw * g.th) % 32 == 0 && !ln_any && !p->d_t,
                "cpd_gemm_conv: GroupNorm statistics need the plain pair kernel, N %% 32 == 0 and 32-row groups inside one image");
  args.dt_col0 = -1;
  args.map_dt = args.map_a0;
  if (p->d_t) {
    CPD_REQUIRE(!geglu && !p->residual && g.taps == 1 && g.nbox == 1 && g.h_out == 1 && g.n_img == 1 && g.tw == BM && p->dt_col0 % 32 == 0 &&
                    p->dt_col0 >= 0 && p->dt_col0 < p->n_out && g.n_store % 32 == 0 && p->ldd_t >= out_rows && out_rows % 8 == 0,
                "cpd_gemm_conv: the transposed tail needs a plain GEMM without residual, dt_col0 %% 32 == 0, rows %% 8 == 0 (the TMA store "
                "clips in 16-byte pieces) and ldd_t >= rows");
    uint64_t dims[2] = {(uint64_t)out_rows, (uint64_t)(p->n_out - p->dt_col0)};
    uint64_t str[1] = {(uint64_t)p->ldd_t * 2};
    uint32_t box[2] = {(uint32_t)(BM / 2), (uint32_t)CHUNK_COLS};  // 64 rows = 128 bytes: one 128B-swizzle span
    rc = cpd_make_tmap16(&args.map_dt, p->d_t, 2, dims, str, box, 128);
    if (rc) return rc;
    args.dt_col0 = p->dt_col0;
  }
  args.splits = splits;
  args.ws = p->splitk_ws;
  args.ws_slice = out_rows * g.n_store;
  {  // output / residual views (c, x, y, n), 64-byte swizzle, box = 32 columns x one pixel box
    const uint64_t rows_x = (uint64_t)((p->m_valid > 0 && g.h_out == 1 && g.n_img == 1) ? p->m_valid : g.w_out);
    // (with a transposed tail the columns from dt_col0 on never go through this map: d may be only dt_col0 columns wide)
    uint64_t dims[4] = {(uint64_t)(p->d_t ? p->dt_col0 : g.n_store), rows_x, (uint64_t)g.h_out, (uint64_t)g.n_img};
    uint32_t box[4] = {(uint32_t)CHUNK_COLS, (uint32_t)g.tw, (uint32_t)g.th, (uint32_t)g.nb};
    uint64_t str[3] = {(uint64_t)p->ldd * 2, (uint64_t)g.w_out * p->ldd * 2, (uint64_t)g.h_out * g.w_out * p->ldd * 2};
    rc = cpd_make_tmap16(&args.map_d, p->d, 4, dims, str, box, 64);
    if (rc) return rc;
    if (p->residual) {
      uint64_t rstr[3] = {(uint64_t)p->ld_res * 2, (uint64_t)g.w_out * p->ld_res * 2, (uint64_t)g.h_out * g.w_out * p->ld_res * 2};
      rc = cpd_make_tmap16(&args.map_res, p->residual, 4, dims, rstr, box, 64);
      if (rc) return rc;
    } else {
      args.map_res = args.map_d;
    }
  }

  const int smem_bytes = stages * args.stage_bytes + STAGING_BYTES + BIAS_BYTES + 512 + 1024;
  const bool of16 = p->out_fp16 != 0;
#ifdef CPD_EXPERIMENTAL
  if (mc == 2) return of16 ? launch2<2, true, 0>(args, smem_bytes, (cudaStream_t)stream) : launch2<2, false, 0>(args, smem_bytes, (cudaStream_t)stream);
#else
  CPD_REQUIRE(mc == 1, "cpd_gemm_conv: the 4-CTA multicast cluster (variant 1000 + BN) is compiled only with -DCPD_EXPERIMENTAL");
#endif
  const int feat = (p->ln_sums ? FEAT_LN_IN : 0) | (p->ln_sums_out ? FEAT_LN_OUT : 0) | (p->d_t ? FEAT_DT : 0) | (p->gn_sums_out ? FEAT_GN_OUT : 0);
  const cpd_status st = of16 ? launch2_feat<true>(args, feat, smem_bytes, (cudaStream_t)stream) : launch2_feat<false>(args, feat, smem_bytes, (cudaStream_t)stream);
  if (st != CPD_OK || splits == 1) return st;
  const int64_t vecs = out_rows * (g.n_store / 8);
  int blocks = (int)((vecs + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  CPD_CUDA_CHECK(cpd_launch(splitk_finalize_kernel, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, (const float*)p->splitk_ws, splits, args.ws_slice, out_rows, g.n_store,
                            g.h_out * g.w_out, p->bias, p->rowvec, p->rowvec_stride, reinterpret_cast<const bf16*>(p->residual),
                            p->ld_res, reinterpret_cast<bf16*>(p->d), p->ldd, p->out_fp16));
  return CPD_OK;
}
