// Plan-level UNet executor behind the C ABI (include/cpd_b200.h: cpd_unet_plan_create / cpd_pack_weights /
// cpd_cache_context_kv / cpd_unet_forward / cpd_unet_plan_destroy).
//
// One cpd_unet_forward call = the whole forward of cpd/models/unet.py:765-831: timestep embedding + emb MLPs (models/util.py:65-85,
// unet.py:529-534,787-788), 12 input blocks, middle block, 12 output blocks with the skip concat (:795-815) and the output head
// (:729-733), sequenced over the hand-written kernels of this library:
//     3x3 / 1x1 convs and every Linear -> cpd_gemm_conv   (tcgen05 implicit GEMM; fused bias / time-embedding / residual / GEGLU)
//     self- and cross-attention         -> cpd_attention   (attention.py:283-348)
//     GroupNorm(+SiLU), LayerNorm       -> cpd_groupnorm / cpd_layernorm (models/util.py:95-105, attention.py:476-478)
//     emb projections of all ResBlocks  -> ONE cpd_small_linear launch (unet.py:223-229)
// The skip concat is never materialised (GroupNorm and the 1x1 skip conv read both sources), the text-context K / V^T are
// computed once per prompt, weights are packed once (cpd_pack_weights), every workspace buffer has a stable address and one
// evaluation is captured into a CUDA graph the plan owns.  Round 1 kept this sequencing in the Python host
// (models/unet.py); it is C++ now so that a host in any language can run the UNet through five C calls.
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/cpd_b200.h"
#include "common.cuh"

namespace {

// ------------------------------------------------------------------------------------------------ small device helpers
__device__ __forceinline__ float load_as_float(const void* src, int dtype, int64_t i) {
  if (dtype == CPD_F32) return reinterpret_cast<const float*>(src)[i];
  if (dtype == CPD_F16) return __half2float(reinterpret_cast<const __half*>(src)[i]);
  return __bfloat162float(reinterpret_cast<const bf16*>(src)[i]);
}
// every parameter / context value goes through the model dtype (bf16) first, like the reference's .half()/.to(dtype) casts
__device__ __forceinline__ void store_from_model_dtype(void* dst, int dst_type, int64_t i, float v, bool accumulate) {
  const float vb = __bfloat162float(__float2bfloat16_rn(v));
  if (dst_type == CPD_F32) {
    float* d = reinterpret_cast<float*>(dst);
    d[i] = accumulate ? d[i] + vb : vb;
  } else if (dst_type == CPD_F16) {
    reinterpret_cast<__half*>(dst)[i] = __float2half_rn(vb);
  } else {
    reinterpret_cast<bf16*>(dst)[i] = __float2bfloat16_rn(vb);
  }
}

enum { PACK_PLAIN = 0, PACK_CONV3 = 1, PACK_PAD_ROWS = 2, PACK_PAD_COLS = 3, PACK_GEGLU_ROWS = 4 };
struct PackJob {
  int kind;
  int64_t dst_rows, dst_cols, dst_ld, dst_off;  // dst element (r, c) at dst_off + r * dst_ld + c
  int64_t src_cols;
  int heads, d, dpad;   // PACK_PAD_*
  int cin;              // PACK_CONV3
  int inner4, hb;       // PACK_GEGLU_ROWS: rows of the value half, half block (128)
  int dst_type;
  int accumulate;       // fp32 destinations only: dst += value (sum of two biases)
};

__global__ void __launch_bounds__(256) pack_kernel(const void* __restrict__ src, int src_dtype, void* __restrict__ dst, PackJob j) {
  const int64_t total = j.dst_rows * j.dst_cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / j.dst_cols, c = i - r * j.dst_cols;
    int64_t si = 0;
    bool valid = true;
    switch (j.kind) {
      case PACK_PLAIN:
        si = r * j.src_cols + c;
        break;
      case PACK_CONV3: {  // src [O][Cin][3][3] -> dst [O][tap][Cin]
        const int64_t tap = c / j.cin, ci = c - tap * j.cin;
        si = (r * j.cin + ci) * 9 + tap;
        break;
      }
      case PACK_PAD_ROWS: {  // [heads * d][K] -> [heads * dpad][K], zero rows
        const int64_t h = r / j.dpad, jj = r - h * j.dpad;
        valid = jj < j.d;
        si = (h * j.d + jj) * j.src_cols + c;
        break;
      }
      case PACK_PAD_COLS: {  // [N][heads * d] -> [N][heads * dpad], zero columns
        const int64_t h = c / j.dpad, jj = c - h * j.dpad;
        valid = jj < j.d;
        si = r * j.src_cols + h * j.d + jj;
        break;
      }
      case PACK_GEGLU_ROWS: {  // per 2*hb-row block: [hb value rows | hb gate rows] (attention.py:92-100 chunks value | gate)
        const int64_t b = r / (2 * j.hb), wi = r - b * 2 * j.hb;
        const int64_t sr = wi < j.hb ? b * j.hb + wi : j.inner4 + b * j.hb + (wi - j.hb);
        si = sr * j.src_cols + c;
        break;
      }
    }
    const float v = valid ? load_as_float(src, src_dtype, si) : 0.f;
    store_from_model_dtype(dst, j.dst_type, j.dst_off + r * j.dst_ld + c, v, j.accumulate != 0);
  }
}

// context [rc][ntok][D] (any dtype) -> model dtype -> activation dtype, rows padded to nk_pad with zeros
__global__ void __launch_bounds__(256) ctx_pack_kernel(const void* __restrict__ src, int src_dtype, void* __restrict__ dst, int dst_type,
                                                       int rc, int ntok, int nk_pad, int D) {
  const int64_t total = (int64_t)rc * nk_pad * D;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / D, c = i - row * D;
    const int64_t b = row / nk_pad, t = row - b * nk_pad;
    const float v = t < ntok ? load_as_float(src, src_dtype, (b * ntok + t) * D + c) : 0.f;
    store_from_model_dtype(dst, dst_type, i, v, false);
  }
}

// dst[r][c] = src[r % src_rows][c]: the vector-conditioning rows repeated per image / a shared timestep broadcast to every row
__global__ void __launch_bounds__(256) repeat_rows_kernel(const void* __restrict__ src, void* __restrict__ dst, int rows, int src_rows,
                                                          int row_bytes4) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t total = (int64_t)rows * row_bytes4;
  const uint32_t* s = reinterpret_cast<const uint32_t*>(src);
  uint32_t* d = reinterpret_cast<uint32_t*>(dst);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / row_bytes4, c = i - r * row_bytes4;
    d[i] = s[(r % src_rows) * row_bytes4 + c];
  }
}

// dst[(b * reps + r)][i] = src[b][i]: an activation tensor computed once per IMAGE broadcast to the image's conditioning rows
// (16-byte vectors; per_image = 16-byte vectors per image)
__global__ void __launch_bounds__(256) expand_images_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int n_images, int reps,
                                                            int64_t per_image) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t total = (int64_t)n_images * per_image;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / per_image, k = i - b * per_image;
    const uint4 v = src[i];
    for (int r = 0; r < reps; ++r) dst[(b * reps + r) * per_image + k] = v;
  }
}

// ------------------------------------------------------------------------------------------------ host-side structures
// LayerNorm folded into the weights of the GEMM that consumes it (cpd_gemm_params.ln_*): one block per output row n,
//   wf[n][k] = round16(W[n][k] * gamma[k]),  g[n] = sum_k wf[n][k],  bf[n] = sum_k W[n][k] * beta[k] + bias[n]
// (g from the ROUNDED wf, so that rstd * (x wf^T - mean * g) is exactly rstd * ((x - mean) wf^T)).  Fixed-order reduction.
__global__ void __launch_bounds__(128) ln_fold_kernel(const uint16_t* __restrict__ w, int cols, const float* __restrict__ gamma,
                                                      const float* __restrict__ beta, const float* __restrict__ bias, int f16,
                                                      uint16_t* __restrict__ wf, float* __restrict__ g, float* __restrict__ bf) {
  const int n = blockIdx.x;
  const uint16_t* wr = w + (int64_t)n * cols;
  uint16_t* wo = wf + (int64_t)n * cols;
  float gs = 0.f, bs = 0.f;
  for (int k = threadIdx.x; k < cols; k += blockDim.x) {
    const float wv = f16 ? __half2float(__ushort_as_half(wr[k])) : __bfloat162float(__ushort_as_bfloat16(wr[k]));
    const float prod = wv * gamma[k];
    uint16_t o;
    float r;
    if (f16) {
      const __half h = __float2half_rn(prod);
      o = __half_as_ushort(h);
      r = __half2float(h);
    } else {
      const __nv_bfloat16 h = __float2bfloat16_rn(prod);
      o = __bfloat16_as_ushort(h);
      r = __bfloat162float(h);
    }
    wo[k] = o;
    gs += r;
    bs = fmaf(wv, beta[k], bs);
  }
  __shared__ float red[2][128];
  red[0][threadIdx.x] = gs;
  red[1][threadIdx.x] = bs;
  __syncthreads();
  for (int off = 64; off > 0; off >>= 1) {
    if ((int)threadIdx.x < off) {
      red[0][threadIdx.x] += red[0][threadIdx.x + off];
      red[1][threadIdx.x] += red[1][threadIdx.x + off];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    g[n] = red[0][0];
    bf[n] = red[1][0] + (bias ? bias[n] : 0.f);
  }
}

struct DevBuf {
  void* p = nullptr;
  int64_t numel = 0;
  int elt = 2;  // bytes per element
};

enum { L_CONV_IN, L_RES, L_ATTN, L_DOWN, L_UP };
struct Layer {
  int kind;
  int a = 0, b = 0, c = 0, d = 0;  // RES: cin, cout, c0, c1 (c1 = skip channels of an output block); ATTN: ch, depth; DOWN / UP: ch
};
struct Block {
  std::string prefix;
  std::vector<Layer> layers;
};

struct Tap {
  void* p = nullptr;
  int c = 0, h = 0, w = 0;
};

struct GraphKey {
  int n, h, w, rpi, t_count, rc, ntok, ry;
  const void *x, *c_in, *t, *eps;
  bool operator<(const GraphKey& o) const { return memcmp(this, &o, sizeof(GraphKey)) < 0; }
};
struct GraphEntry {
  cudaGraphExec_t exec = nullptr;
  int64_t launches = 0;
  uint64_t last_use = 0;
};

struct ProfRec {
  const char* kind;
  std::string label;
  double flops;
  cudaEvent_t e0, e1;
};

struct FoldJob {  // one ln_fold_kernel launch (buffer names in cpd_unet_plan::w)
  std::string w, gamma, beta, bias, wf, g, bf;
  int64_t rows = 0, dst_row0 = 0;
  int cols = 0;
};

struct WeightSpec {
  std::vector<PackJob> jobs;
  std::vector<std::string> dst;  // destination buffer of each job
  int64_t numel = 0;
  bool loaded = false;
};

inline int round16(int d) { return (d + 15) / 16 * 16; }

}  // namespace

struct cpd_unet_plan {
  cpd_unet_config cfg;
  int device = 0;
  int mc = 0, ted = 0, adm = 0, emb_total = 0;
  std::vector<Block> inputs, outputs;
  Block middle;
  std::map<std::string, int> emb_off;      // ResBlock prefix -> first column of its emb projection in emb_all
  std::map<std::string, DevBuf> w;         // packed weights
  std::map<std::string, WeightSpec> spec;  // reference state_dict name -> how it is packed
  std::map<std::string, DevBuf> ws;        // workspace, keyed by name + size
  // context cache
  int rc = 0, ntok = 0, nk_pad = 0;
  bool have_ctx = false;
  // vector conditioning
  DevBuf y;
  int ry = 0;
  // graphs
  std::map<GraphKey, GraphEntry> graphs;
  uint64_t use_counter = 0;
  // bookkeeping of the last forward
  int64_t launches = 0;
  bool capturing = false;  // the current forward_impl runs inside a stream capture: no allocation, no tuning
  bool allow_alloc = true;
  std::vector<Tap> tap_in, tap_out;
  Tap tap_mid;
  DevBuf tune_scratch, splitk;
  cudaStream_t cap_stream = nullptr;
  // LayerNorm folding (CPD_UNET_FOLD_LN=0: every LayerNorm as its own kernel, A/B measurements)
  std::vector<FoldJob> folds;
  bool fold_ln = true, fold_dirty = true;
  int fold_geglu_min_ch = 640;
  // Levels with fewer pixels PER IMAGE keep the LayerNorm kernels (their GEMMs want split-K / one-tile variants).  The decision
  // must not depend on the batch or on the rows per image: an image has to come out the same whatever it is batched with
  // (tools/check_rowshard.py: image-sharded and row-sharded runs are bit-identical to the single-GPU run).
  int fold_min_hw = 256;
  // GroupNorm statistics from the producing conv's epilogue (cpd_gemm_params.gn_sums_out): one zeroed arena of fixed-point
  // accumulators per forward, handed out in launch order.  OPT-IN (CPD_UNET_GN_STATS=1): measured on B200 (tools/bench_gnstats.py)
  // the ~300 extra epilogue instructions per 128 x 32 chunk are NOT hidden behind the main loop of the 3x3 convolutions (64 x 64:
  // 91 -> 103 us per conv against 33.8 -> 20.2 us for the GroupNorm; 32 x 32: +5 us against -6 us; 16 x 16: +2.4 against -0.3):
  // break-even at best, 15.8 vs 15.4 ms per 16-row evaluation.  The default keeps every GroupNorm's own statistics pass.
  bool gn_stats = false;
  int64_t gn_min_rows = 4096;
  DevBuf gn_arena;                 // int64 accumulators
  std::vector<void*> gn_retired;   // outgrown arenas: captured graphs may still point into them
  int64_t gn_used = 0, gn_need = 0;  // bytes handed out in the current forward / by the last complete one
  bool share_prefix = true;  // CPD_UNET_SHARE_PREFIX=0: evaluate every row through the whole network (A/B measurements)
  // profiling (eager launches bracketed by events)
  bool profile = false;
  std::vector<ProfRec> prof;
  std::string err;
};

namespace {

#define PLAN_CHECK(expr)                     \
  do {                                       \
    const cpd_status _s = (expr);            \
    if (_s != CPD_OK) return _s;             \
  } while (0)

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
    else prev = -1;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

void heads_of(const cpd_unet_config& c, int ch, int* nh, int* dh) {
  if (c.num_head_channels == -1) {
    *nh = c.num_heads;
    *dh = ch / c.num_heads;
  } else {
    *nh = ch / c.num_head_channels;
    *dh = c.num_head_channels;
  }
}

bool has_attention(const cpd_unet_config& c, int ds) {
  for (int i = 0; i < c.n_attention_resolutions; ++i)
    if (c.attention_resolutions[i] == ds) return true;
  return false;
}

// Block structure built by the reference constructor (unet.py:545-727)
void enumerate_blocks(cpd_unet_plan* P) {
  const cpd_unet_config& c = P->cfg;
  const int mc = c.model_channels;
  auto depth = [&](int level) { return c.transformer_depth[level < c.n_levels ? level : c.n_levels - 1]; };
  std::vector<int> chans;
  {
    Block b;
    b.prefix = "input_blocks.0.";
    Layer l;
    l.kind = L_CONV_IN;
    l.a = c.in_channels;
    l.b = mc;
    b.layers.push_back(l);
    P->inputs.push_back(b);
    chans.push_back(mc);
  }
  int ch = mc, ds = 1;
  for (int level = 0; level < c.n_levels; ++level) {
    const int mult = c.channel_mult[level];
    for (int i = 0; i < c.num_res_blocks; ++i) {
      Block b;
      b.prefix = "input_blocks." + std::to_string(P->inputs.size()) + ".";
      Layer r;
      r.kind = L_RES;
      r.a = ch;
      r.b = mult * mc;
      r.c = ch;
      r.d = 0;
      b.layers.push_back(r);
      ch = mult * mc;
      if (has_attention(c, ds)) {
        Layer a;
        a.kind = L_ATTN;
        a.a = ch;
        a.b = depth(level);
        b.layers.push_back(a);
      }
      P->inputs.push_back(b);
      chans.push_back(ch);
    }
    if (level != c.n_levels - 1) {
      Block b;
      b.prefix = "input_blocks." + std::to_string(P->inputs.size()) + ".";
      Layer d;
      d.kind = L_DOWN;
      d.a = ch;
      b.layers.push_back(d);
      P->inputs.push_back(b);
      chans.push_back(ch);
      ds *= 2;
    }
  }
  {
    P->middle.prefix = "middle_block.";
    Layer r;
    r.kind = L_RES;
    r.a = r.b = r.c = ch;
    Layer a;
    a.kind = L_ATTN;
    a.a = ch;
    a.b = depth(c.n_levels - 1);
    P->middle.layers = {r, a, r};
  }
  for (int level = c.n_levels - 1; level >= 0; --level) {
    const int mult = c.channel_mult[level];
    for (int i = 0; i < c.num_res_blocks + 1; ++i) {
      const int ich = chans.back();
      chans.pop_back();
      Block b;
      b.prefix = "output_blocks." + std::to_string(P->outputs.size()) + ".";
      Layer r;
      r.kind = L_RES;
      r.a = ch + ich;
      r.b = mc * mult;
      r.c = ch;
      r.d = ich;
      b.layers.push_back(r);
      ch = mc * mult;
      if (has_attention(c, ds)) {
        Layer a;
        a.kind = L_ATTN;
        a.a = ch;
        a.b = depth(level);
        b.layers.push_back(a);
      }
      if (level && i == c.num_res_blocks) {
        Layer u;
        u.kind = L_UP;
        u.a = ch;
        b.layers.push_back(u);
        ds /= 2;
      }
      P->outputs.push_back(b);
    }
  }
}

// ---- weights ------------------------------------------------------------------------------------------------
cpd_status alloc_dev(DevBuf* b, int64_t numel, int elt, bool zero) {
  b->numel = numel;
  b->elt = elt;
  CPD_CUDA_CHECK(cudaMalloc(&b->p, (size_t)(numel > 0 ? numel : 1) * elt));
  if (zero) CPD_CUDA_CHECK(cudaMemset(b->p, 0, (size_t)numel * elt));
  return CPD_OK;
}

struct SpecBuilder {
  cpd_unet_plan* P;
  cpd_status st = CPD_OK;
  int act_type() const { return P->cfg.act_fp16 ? CPD_F16 : CPD_BF16; }
  DevBuf* dest(const std::string& name, int64_t numel, int type, bool zero = false) {
    auto it = P->w.find(name);
    if (it != P->w.end()) return &it->second;
    DevBuf b;
    if (st == CPD_OK) st = alloc_dev(&b, numel, type == CPD_F32 ? 4 : 2, zero);
    P->w[name] = b;
    return &P->w[name];
  }
  void add(const std::string& ref, int64_t numel, const std::string& dst, PackJob j) {
    WeightSpec& s = P->spec[ref];
    s.numel = numel;
    s.jobs.push_back(j);
    s.dst.push_back(dst);
  }
  static PackJob plain(int64_t rows, int64_t cols, int type) {
    PackJob j;
    memset(&j, 0, sizeof(j));
    j.kind = PACK_PLAIN;
    j.dst_rows = rows;
    j.dst_cols = cols;
    j.dst_ld = cols;
    j.src_cols = cols;
    j.dst_type = type;
    return j;
  }
  // 1-D parameter used in fp32 (rounded to the model dtype first)
  void vec(const std::string& ref, const std::string& dst, int64_t n) {
    dest(dst, n, CPD_F32);
    add(ref, n, dst, plain(1, n, CPD_F32));
  }
  void mat(const std::string& ref, const std::string& dst, int64_t rows, int64_t cols, int type) {
    dest(dst, rows * cols, type);
    add(ref, rows * cols, dst, plain(rows, cols, type));
  }
  void conv3(const std::string& ref, const std::string& dst, int64_t cout, int cin, int type) {
    dest(dst, cout * 9 * cin, type);
    PackJob j = plain(cout, 9 * (int64_t)cin, type);
    j.kind = PACK_CONV3;
    j.cin = cin;
    add(ref, cout * 9 * (int64_t)cin, dst, j);
  }
  void pad_rows(const std::string& ref, const std::string& dst, int heads, int d, int dpad, int64_t K, int64_t dst_row0, int64_t dst_rows_total) {
    dest(dst, dst_rows_total * K, act_type());
    PackJob j = plain((int64_t)heads * dpad, K, act_type());
    j.kind = PACK_PAD_ROWS;
    j.heads = heads;
    j.d = d;
    j.dpad = dpad;
    j.dst_off = dst_row0 * K;
    add(ref, (int64_t)heads * d * K, dst, j);
  }
  void pad_cols(const std::string& ref, const std::string& dst, int heads, int d, int dpad, int64_t N) {
    dest(dst, N * heads * dpad, act_type());
    PackJob j = plain(N, (int64_t)heads * dpad, act_type());
    j.kind = PACK_PAD_COLS;
    j.heads = heads;
    j.d = d;
    j.dpad = dpad;
    j.src_cols = (int64_t)heads * d;
    add(ref, N * heads * d, dst, j);
  }
};

constexpr int GEGLU_BLOCK = 256;  // = the CTA-pair kernel's tile width (cpd_gemm_conv CPD_EPI_GEGLU, geglu_block)

cpd_status build_specs(cpd_unet_plan* P) {
  const cpd_unet_config& c = P->cfg;
  SpecBuilder B{P};
  const int act = B.act_type();
  const int mc = P->mc, ted = P->ted;
  // time embedding MLP (cuda-core kernels: weights stay bf16)
  B.mat("time_embed.0.weight", "te0.w", ted, mc, CPD_BF16);
  B.vec("time_embed.0.bias", "te0.b", ted);
  B.mat("time_embed.2.weight", "te2.w", ted, ted, CPD_BF16);
  B.vec("time_embed.2.bias", "te2.b", ted);
  if (P->adm) {
    // emb = time_embed(t_emb) + label_emb(y) (SDXL extension) is ONE small linear over [SiLU(e1) | SiLU(l1)]: the second layers
    // are concatenated along K and their biases summed
    B.mat("label_emb.0.0.weight", "lab0.w", ted, P->adm, CPD_BF16);
    B.vec("label_emb.0.0.bias", "lab0.b", ted);
    B.dest("te2lab2.w", (int64_t)ted * 2 * ted, CPD_BF16);
    B.dest("te2lab2.b", ted, CPD_F32, true);
    PackJob j = SpecBuilder::plain(ted, ted, CPD_BF16);
    j.dst_ld = 2 * ted;
    B.add("time_embed.2.weight", (int64_t)ted * ted, "te2lab2.w", j);
    j.dst_off = ted;
    B.add("label_emb.0.2.weight", (int64_t)ted * ted, "te2lab2.w", j);
    PackJob jb = SpecBuilder::plain(1, ted, CPD_F32);
    jb.accumulate = 1;
    B.add("time_embed.2.bias", ted, "te2lab2.b", jb);
    B.add("label_emb.0.2.bias", ted, "te2lab2.b", jb);
  }
  // first pass: the concatenated emb projection of every ResBlock (unet.py:223-229) needs its total width
  int off = 0;
  auto count_emb = [&](const Block& b) {
    for (size_t j = 0; j < b.layers.size(); ++j)
      if (b.layers[j].kind == L_RES) {
        P->emb_off[b.prefix + std::to_string(j) + "."] = off;
        off += b.layers[j].b;
      }
  };
  for (const Block& b : P->inputs) count_emb(b);
  count_emb(P->middle);
  for (const Block& b : P->outputs) count_emb(b);
  P->emb_total = off;
  B.dest("emb_all.w", (int64_t)off * ted, CPD_BF16);
  B.dest("emb_all.b", off, CPD_F32);

  auto res = [&](const std::string& p, int cin, int cout) {
    B.vec(p + "in_layers.0.weight", p + "gn1.g", cin);
    B.vec(p + "in_layers.0.bias", p + "gn1.b", cin);
    B.conv3(p + "in_layers.2.weight", p + "conv1.w", cout, cin, act);
    B.vec(p + "in_layers.2.bias", p + "conv1.b", cout);
    {
      PackJob j = SpecBuilder::plain(cout, ted, CPD_BF16);
      j.dst_off = (int64_t)P->emb_off[p] * ted;
      B.add(p + "emb_layers.1.weight", (int64_t)cout * ted, "emb_all.w", j);
      PackJob jb = SpecBuilder::plain(1, cout, CPD_F32);
      jb.dst_off = P->emb_off[p];
      B.add(p + "emb_layers.1.bias", cout, "emb_all.b", jb);
    }
    B.vec(p + "out_layers.0.weight", p + "gn2.g", cout);
    B.vec(p + "out_layers.0.bias", p + "gn2.b", cout);
    B.conv3(p + "out_layers.3.weight", p + "conv2.w", cout, cout, act);
    B.vec(p + "out_layers.3.bias", p + "conv2.b", cout);
    if (cin != cout) {
      B.mat(p + "skip_connection.weight", p + "skip.w", cout, cin, act);
      B.vec(p + "skip_connection.bias", p + "skip.b", cout);
    }
  };
  auto tblock = [&](const std::string& b, int ch) {
    int nh, dh;
    heads_of(c, ch, &nh, &dh);
    const int dpad = round16(dh), ip = nh * dpad;
    for (const char* n : {"norm1", "norm2", "norm3"}) {
      B.vec(b + n + ".weight", b + n + ".g", ch);
      B.vec(b + n + ".bias", b + n + ".b", ch);
    }
    B.pad_rows(b + "attn1.to_q.weight", b + "attn1.qk.w", nh, dh, dpad, ch, 0, 2 * ip);  // fused Q | K projection
    B.pad_rows(b + "attn1.to_k.weight", b + "attn1.qk.w", nh, dh, dpad, ch, ip, 2 * ip);
    B.pad_rows(b + "attn1.to_v.weight", b + "attn1.v.w", nh, dh, dpad, ch, 0, ip);
    B.pad_cols(b + "attn1.to_out.0.weight", b + "attn1.out.w", nh, dh, dpad, ch);
    B.vec(b + "attn1.to_out.0.bias", b + "attn1.out.b", ch);
    B.pad_rows(b + "attn2.to_q.weight", b + "attn2.q.w", nh, dh, dpad, ch, 0, ip);
    B.pad_rows(b + "attn2.to_k.weight", b + "attn2.k.w", nh, dh, dpad, c.context_dim, 0, ip);
    B.pad_rows(b + "attn2.to_v.weight", b + "attn2.v.w", nh, dh, dpad, c.context_dim, 0, ip);
    B.pad_cols(b + "attn2.to_out.0.weight", b + "attn2.out.w", nh, dh, dpad, ch);
    B.vec(b + "attn2.to_out.0.bias", b + "attn2.out.b", ch);
    {  // GEGLU: [128 value rows | 128 gate rows] per 256-column tile
      const int64_t inner4 = 4 * (int64_t)ch;
      B.dest(b + "ff1.w", 2 * inner4 * ch, act);
      PackJob j = SpecBuilder::plain(2 * inner4, ch, act);
      j.kind = PACK_GEGLU_ROWS;
      j.inner4 = (int)inner4;
      j.hb = GEGLU_BLOCK / 2;
      B.add(b + "ff.net.0.proj.weight", 2 * inner4 * ch, b + "ff1.w", j);
      B.dest(b + "ff1.b", 2 * inner4, CPD_F32);
      PackJob jb = SpecBuilder::plain(2 * inner4, 1, CPD_F32);
      jb.kind = PACK_GEGLU_ROWS;
      jb.inner4 = (int)inner4;
      jb.hb = GEGLU_BLOCK / 2;
      B.add(b + "ff.net.0.proj.bias", 2 * inner4, b + "ff1.b", jb);
    }
    B.mat(b + "ff.net.2.weight", b + "ff2.w", ch, 4 * (int64_t)ch, act);
    B.vec(b + "ff.net.2.bias", b + "ff2.b", ch);
    // folded copies (LayerNorm applied by the consumer GEMM's epilogue): filled by ensure_folded once the weights are loaded
    auto fold = [&](const std::string& w, const char* norm, const std::string& bias, const std::string& dst, int64_t rows, int64_t row0,
                    int64_t rows_total) {
      B.dest(dst + ".wf", rows_total * ch, act);
      B.dest(dst + ".g", rows_total, CPD_F32);
      B.dest(dst + ".bf", rows_total, CPD_F32);
      FoldJob j;
      j.w = w; j.gamma = b + norm + ".g"; j.beta = b + norm + ".b"; j.bias = bias;
      j.wf = dst + ".wf"; j.g = dst + ".g"; j.bf = dst + ".bf";
      j.rows = rows; j.dst_row0 = row0; j.cols = ch;
      P->folds.push_back(j);
    };
    fold(b + "attn1.qk.w", "norm1", "", b + "attn1.qkv", 2 * ip, 0, 3 * ip);   // fused Q | K | V projection of LN1(x)
    fold(b + "attn1.v.w", "norm1", "", b + "attn1.qkv", ip, 2 * ip, 3 * ip);
    fold(b + "attn2.q.w", "norm2", "", b + "attn2.q", ip, 0, ip);
    fold(b + "ff1.w", "norm3", b + "ff1.b", b + "ff1", 8 * (int64_t)ch, 0, 8 * (int64_t)ch);
  };
  auto attn = [&](const std::string& p, int ch, int depth) {
    B.vec(p + "norm.weight", p + "norm.g", ch);
    B.vec(p + "norm.bias", p + "norm.b", ch);
    B.mat(p + "proj_in.weight", p + "proj_in.w", ch, ch, act);  // 1x1 conv or Linear: the same [out][in] matrix
    B.vec(p + "proj_in.bias", p + "proj_in.b", ch);
    B.mat(p + "proj_out.weight", p + "proj_out.w", ch, ch, act);
    B.vec(p + "proj_out.bias", p + "proj_out.b", ch);
    for (int d = 0; d < depth; ++d) tblock(p + "transformer_blocks." + std::to_string(d) + ".", ch);
  };
  auto block = [&](const Block& b) {
    for (size_t j = 0; j < b.layers.size(); ++j) {
      const Layer& l = b.layers[j];
      const std::string p = b.prefix + std::to_string(j) + ".";
      if (l.kind == L_CONV_IN) {
        B.conv3(p + "weight", p + "w", l.b, l.a, CPD_BF16);
        B.vec(p + "bias", p + "b", l.b);
      } else if (l.kind == L_RES) {
        res(p, l.a, l.b);
      } else if (l.kind == L_ATTN) {
        if (l.a % 64) B.st = CPD_ERR_UNSUPPORTED;
        attn(p, l.a, l.b);
      } else if (l.kind == L_DOWN) {
        B.conv3(p + "op.weight", p + "w", l.a, l.a, act);
        B.vec(p + "op.bias", p + "b", l.a);
      } else if (l.kind == L_UP) {
        B.conv3(p + "conv.weight", p + "w", l.a, l.a, act);
        B.vec(p + "conv.bias", p + "b", l.a);
      }
    }
  };
  for (const Block& b : P->inputs) block(b);
  block(P->middle);
  for (const Block& b : P->outputs) block(b);
  B.vec("out.0.weight", "out.gn.g", mc);
  B.vec("out.0.bias", "out.gn.b", mc);
  B.conv3("out.2.weight", "out.w", c.out_channels, mc, CPD_BF16);
  B.vec("out.2.bias", "out.b", c.out_channels);
  if (B.st == CPD_ERR_UNSUPPORTED) cpd_set_error("cpd_unet_plan_create: attention widths must be multiples of 64 channels");
  return B.st;
}

// ---- workspace ----------------------------------------------------------------------------------------------
cpd_status ws_get(cpd_unet_plan* P, const std::string& name, int64_t numel, int elt, void** out) {
  const std::string key = name + "#" + std::to_string(numel) + "x" + std::to_string(elt);
  auto it = P->ws.find(key);
  if (it == P->ws.end()) {
    CPD_REQUIRE(P->allow_alloc,
                "cpd_unet_forward: workspace buffer %s (%lld elements) does not exist yet - evaluate this shape once outside "
                "stream capture first",
                name.c_str(), (long long)numel);
    DevBuf b;
    PLAN_CHECK(alloc_dev(&b, numel, elt, false));
    it = P->ws.emplace(key, b).first;
  }
  *out = it->second.p;
  return CPD_OK;
}

void* W(cpd_unet_plan* P, const std::string& name) {
  auto it = P->w.find(name);
  return it == P->w.end() ? nullptr : it->second.p;
}

// every kernel-level call goes through here: launch counting + optional event bracketing
struct OpScope {
  cpd_unet_plan* P;
  cudaStream_t st;
  bool on;
  ProfRec rec;
  OpScope(cpd_unet_plan* P_, cudaStream_t s, const char* kind, const std::string& label, double flops, int launches) : P(P_), st(s) {
    P->launches += launches;
    on = P->profile && !P->capturing;
    if (on) {
      rec.kind = kind;
      rec.label = label;
      rec.flops = flops;
      cudaEventCreate(&rec.e0);
      cudaEventCreate(&rec.e1);
      cudaEventRecord(rec.e0, st);
    }
  }
  ~OpScope() {
    if (on) {
      cudaEventRecord(rec.e1, st);
      P->prof.push_back(rec);
    }
  }
};

struct GemmOpt {
  const void* a1 = nullptr;
  int c1 = 0;
  int ksize = 1, stride = 1;
  const float* bias = nullptr;
  const float* rowvec = nullptr;
  int rowvec_stride = 0;
  const void* residual = nullptr;
  int ld_res = 0;
  int epilogue = CPD_EPI_NONE;
  // folded LayerNorm (cpd_gemm_params.ln_*): producer side / consumer side; transposed tail
  float* ln_out = nullptr;
  int* ln_parts_out = nullptr;
  const float* ln_in = nullptr;
  int ln_parts = 0;
  int64_t ln_ld = 0;
  const float* ln_g = nullptr;
  void* d_t = nullptr;
  int dt_col0 = 0;
  int64_t ldd_t = 0;
  int ldd = 0;  // 0: n_out (n_out / 2 for GEGLU)
  long long* gn_out = nullptr;  // GroupNorm statistics of the output (fixed-point accumulators, zeroed)
};

cpd_status ensure_scratch(cpd_unet_plan* P, int64_t bytes, cudaStream_t st) {
  if (P->capturing || (int64_t)P->tune_scratch.numel >= bytes) return CPD_OK;
  if (P->tune_scratch.p) {
    CPD_CUDA_CHECK(cudaStreamSynchronize(st));
    CPD_CUDA_CHECK(cudaFree(P->tune_scratch.p));
    P->tune_scratch.p = nullptr;
  }
  return alloc_dev(&P->tune_scratch, bytes, 1, false);
}

cpd_status gemm(cpd_unet_plan* P, cudaStream_t st, const void* a0, const void* wt, void* out, int n_img, int h, int w, int c0, int n_out,
                const GemmOpt& o = GemmOpt()) {
  cpd_gemm_params p;
  memset(&p, 0, sizeof(p));
  p.a0 = a0;
  p.a1 = o.a1;
  p.c0 = c0;
  p.c1 = o.c1;
  p.n_img = n_img;
  p.h_in = h;
  p.w_in = w;
  p.ksize = o.ksize;
  p.stride = o.stride;
  p.wt = wt;
  p.n_out = n_out;
  p.bias = o.bias;
  p.rowvec = o.rowvec;
  p.rowvec_stride = o.rowvec_stride;
  p.residual = o.residual;
  p.ld_res = o.ld_res;
  p.d = out;
  p.ldd = o.ldd ? o.ldd : (o.epilogue == CPD_EPI_GEGLU ? n_out / 2 : n_out);
  p.epilogue = o.epilogue;
  p.ln_sums_out = o.ln_out;
  p.ln_parts_out = o.ln_parts_out;
  p.ln_sums = o.ln_in;
  p.ln_parts = o.ln_parts;
  p.ln_ld = o.ln_ld;
  p.ln_g = o.ln_g;
  p.ln_c = o.ksize * o.ksize * (c0 + o.c1);
  p.ln_eps = 1e-5f;  // nn.LayerNorm default (attention.py:476-478)
  p.d_t = o.d_t;
  p.dt_col0 = o.dt_col0;
  p.ldd_t = o.ldd_t;
  p.gn_sums_out = o.gn_out;
  p.variant = 0;
  p.a_fp16 = p.b_fp16 = p.out_fp16 = P->cfg.act_fp16;
  p.geglu_block = o.epilogue == CPD_EPI_GEGLU ? GEGLU_BLOCK : 0;
  p.splitk_ws = reinterpret_cast<float*>(P->splitk.p);
  p.splitk_ws_floats = P->splitk.numel;
  const int64_t rows = (int64_t)n_img * (h / o.stride) * (w / o.stride);
  PLAN_CHECK(ensure_scratch(P, rows * p.ldd * 2, st));
  p.tune_scratch = P->capturing ? nullptr : P->tune_scratch.p;
  p.tune_scratch_bytes = P->tune_scratch.numel;
  const int K = o.ksize * o.ksize * (c0 + o.c1);
  char label[96];
  snprintf(label, sizeof(label), "M=%lld N=%d K=%d%s", (long long)rows, n_out, K, o.epilogue ? " geglu" : "");
  OpScope op(P, st, "gemm_conv", label, 2.0 * rows * n_out * K, 1);
  return cpd_gemm_conv(&p, st);
}

cpd_status groupnorm(cpd_unet_plan* P, cudaStream_t st, const void* a0, const void* a1, int c0, int c1, int n_img, int hw, const std::string& g,
                     const std::string& b, float eps, int silu, double* stats, void* out, const long long* chan_sums = nullptr) {
  char label[64];
  snprintf(label, sizeof(label), "%sn=%d hw=%d C=%d", chan_sums && !a1 ? "apply " : "", n_img, hw, c0 + c1);
  if (chan_sums && !a1) {  // the producer's epilogue already reduced the statistics: normalise only
    OpScope op(P, st, "groupnorm", label, 0.0, 1);
    return cpd_groupnorm_apply(a0, c0, n_img, hw, (const float*)W(P, g), (const float*)W(P, b), eps, silu, P->cfg.act_fp16, chan_sums, out, st);
  }
  OpScope op(P, st, "groupnorm", label, 0.0, cpd_groupnorm_launches(c0 + c1, n_img, hw));
  return cpd_groupnorm(a0, a1, c0, c1, n_img, hw, (const float*)W(P, g), (const float*)W(P, b), eps, silu, P->cfg.act_fp16, stats, out, st);
}

cpd_status layernorm(cpd_unet_plan* P, cudaStream_t st, const void* x, int rows, int c, const std::string& g, const std::string& b, void* out) {
  char label[64];
  snprintf(label, sizeof(label), "rows=%d C=%d", rows, c);
  OpScope op(P, st, "layernorm", label, 0.0, 1);
  return cpd_layernorm(x, rows, c, (const float*)W(P, g), (const float*)W(P, b), 1e-5f, P->cfg.act_fp16, out, st);
}

cpd_status attention(cpd_unet_plan* P, cudaStream_t st, const void* q, int ldq, const void* k, int ldk, const void* vt, int ldvt, void* o, int ldo,
                     int batch, int heads, int nq, int nk, int nk_pad, int dpad, int d_head, int kv_batch) {
  cpd_attn_params p;
  memset(&p, 0, sizeof(p));
  p.q = q; p.ldq = ldq; p.k = k; p.ldk = ldk; p.vt = vt; p.ldvt = ldvt; p.o = o; p.ldo = ldo;
  p.batch = batch; p.heads = heads; p.nq = nq; p.nk = nk; p.nk_pad = nk_pad; p.dpad = dpad;
  p.scale = (float)pow((double)d_head, -0.5);  // dim_head ** -0.5 (attention.py:176)
  p.kv_batch = kv_batch;
  p.act_fp16 = P->cfg.act_fp16;
  p.d_head = d_head;
  char label[96];
  snprintf(label, sizeof(label), "B=%d H=%d nq=%d nk=%d d=%d", batch, heads, nq, nk, d_head);
  OpScope op(P, st, "attention", label, 4.0 * batch * heads * (double)nq * nk * d_head, 1);  // algorithmic: the REAL head dim
  return cpd_attention(&p, st);
}

// ---- layers (unet.py:249-280, attention.py:469-537) ----------------------------------------------------------
struct Act {  // an NHWC activation tensor
  void* p = nullptr;
  int c = 0;
  const long long* gn = nullptr;  // its GroupNorm statistics, when the GEMM that produced it emitted them
};

// Accumulators for the GroupNorm statistics of one [n][hw][c] tensor, or NULL (level too small, arena not yet sized: the consumer
// then computes its own statistics).  The demand is recorded either way; cpd_unet_forward sizes the arena from it.
long long* gn_alloc(cpd_unet_plan* P, int n, int hw, int c) {
  if (!P->gn_stats || (int64_t)n * hw < P->gn_min_rows || hw % 32 != 0 || c % 32 != 0) return nullptr;
  const int64_t bytes = (int64_t)n * c * 2 * sizeof(long long);
  const int64_t off = P->gn_used;
  P->gn_used += bytes;
  if (!P->gn_arena.p || P->gn_used > P->gn_arena.numel) return nullptr;
  return reinterpret_cast<long long*>(reinterpret_cast<uint8_t*>(P->gn_arena.p) + off);
}

struct FwdCtx {
  cpd_unet_plan* P;
  cudaStream_t st;
  int R, h, w;
  const float* emb_all;
  int emb_stride;
  double* stats;
  // folded LayerNorm: per-row partial sums of the transformer's residual stream `tr.h`, written by the GEMM that produced it
  // (ln_parts == 0: none - the next LayerNorm runs as its own kernel)
  float* lnp = nullptr;
  int ln_parts = 0;
  int64_t ln_ld = 0;
};

// Folded copies of the weights behind LN1 / LN2 / LN3 (see ln_fold_kernel); after any cpd_pack_weights.
cpd_status ensure_folded(cpd_unet_plan* P, cudaStream_t st) {
  if (!P->fold_dirty) return CPD_OK;
  for (const FoldJob& j : P->folds) {
    const uint16_t* w = (const uint16_t*)W(P, j.w);
    uint16_t* wf = (uint16_t*)W(P, j.wf);
    float* g = (float*)W(P, j.g);
    float* bf = (float*)W(P, j.bf);
    CPD_REQUIRE(w && wf && g && bf && W(P, j.gamma) && W(P, j.beta), "cpd_unet_forward: folded-LayerNorm buffers of %s are missing", j.w.c_str());
    ln_fold_kernel<<<(unsigned)j.rows, 128, 0, st>>>(w, j.cols, (const float*)W(P, j.gamma), (const float*)W(P, j.beta),
                                                      j.bias.empty() ? nullptr : (const float*)W(P, j.bias), P->cfg.act_fp16,
                                                      wf + j.dst_row0 * j.cols, g + j.dst_row0, bf + j.dst_row0);
    CPD_CUDA_CHECK(cudaGetLastError());
  }
  P->fold_dirty = false;
  return CPD_OK;
}

// The GEMM that writes the residual stream also emits its rows' partial sums when the level is large enough to fold.
void ln_producer(FwdCtx& F, GemmOpt& g) {
  F.ln_parts = 0;
  if (!F.lnp) return;
  g.ln_out = F.lnp;
  g.ln_parts_out = &F.ln_parts;
  g.ln_ld = F.ln_ld;
}
void ln_consumer(FwdCtx& F, GemmOpt& g, const float* gvec, const float* bias_folded) {
  g.ln_in = F.lnp;
  g.ln_parts = F.ln_parts;
  g.ln_ld = F.ln_ld;
  g.ln_g = gvec;
  g.bias = bias_folded;
}
cpd_status ln_setup(FwdCtx& F, int64_t T, int ch) {
  F.lnp = nullptr;
  F.ln_parts = 0;
  if (!F.P->fold_ln || F.h * F.w < F.P->fold_min_hw) return CPD_OK;
  void* b;
  const int64_t max_parts = 2 * ((ch + 63) / 64);  // narrowest tile the tuner may pick: 64 columns x two epilogue groups
  PLAN_CHECK(ws_get(F.P, "tr.lnpart", max_parts * T * 2, 4, &b));
  F.lnp = (float*)b;
  F.ln_ld = T;
  return CPD_OK;
}

cpd_status res_block(FwdCtx& F, const std::string& p, Act x0, Act x1, int cout, Act* outp) {
  cpd_unet_plan* P = F.P;
  const int hw = F.h * F.w, cin = x0.c + x1.c;
  const int64_t T = (int64_t)F.R * hw;
  void *gn, *h1, *gn2, *out;
  PLAN_CHECK(ws_get(P, "gn", T * cin, 2, &gn));
  PLAN_CHECK(groupnorm(P, F.st, x0.p, x1.p, x0.c, x1.c, F.R, hw, p + "gn1.g", p + "gn1.b", 1e-5f, 1, F.stats, gn, x1.p ? nullptr : x0.gn));
  PLAN_CHECK(ws_get(P, "h1", T * cout, 2, &h1));
  long long* h1_gn = gn_alloc(P, F.R, hw, cout);
  {
    GemmOpt o;
    o.ksize = 3;
    o.gn_out = h1_gn;
    o.bias = (const float*)W(P, p + "conv1.b");
    o.rowvec = F.emb_all + P->emb_off[p];  // h + emb_out[:, :, None, None] (unet.py:266-274) in the epilogue
    o.rowvec_stride = F.emb_stride;
    PLAN_CHECK(gemm(P, F.st, gn, W(P, p + "conv1.w"), h1, F.R, F.h, F.w, cin, cout, o));
  }
  PLAN_CHECK(ws_get(P, "gn", T * cout, 2, &gn2));
  PLAN_CHECK(groupnorm(P, F.st, h1, nullptr, cout, 0, F.R, hw, p + "gn2.g", p + "gn2.b", 1e-5f, 1, F.stats, gn2, h1_gn));
  const void* skip = x0.p;
  if (cin != cout) {  // 1x1 skip conv over both concat sources (unet.py:247)
    void* sk;
    PLAN_CHECK(ws_get(P, "skip", T * cout, 2, &sk));
    GemmOpt o;
    o.a1 = x1.p;
    o.c1 = x1.c;
    o.bias = (const float*)W(P, p + "skip.b");
    PLAN_CHECK(gemm(P, F.st, x0.p, W(P, p + "skip.w"), sk, F.R, F.h, F.w, x0.c, cout, o));
    skip = sk;
  }
  PLAN_CHECK(ws_get(P, p + "out", T * cout, 2, &out));
  long long* out_gn = gn_alloc(P, F.R, hw, cout);  // the GroupNorm of the next ResBlock / SpatialTransformer reads this tensor
  {
    GemmOpt o;
    o.ksize = 3;
    o.gn_out = out_gn;
    o.bias = (const float*)W(P, p + "conv2.b");
    o.residual = skip;
    o.ld_res = cout;
    PLAN_CHECK(gemm(P, F.st, gn2, W(P, p + "conv2.w"), out, F.R, F.h, F.w, cout, cout, o));
  }
  outp->p = out;
  outp->c = cout;
  outp->gn = out_gn;
  return CPD_OK;
}

// SpatialTransformer (attention.py:506-537) in four pieces, so that the executor can run the part in front of the first
// cross-attention once per IMAGE (see forward_impl): pre = GroupNorm + proj_in; self = x + attn1(LN1(x)); rest = cross-attention +
// GEGLU feed-forward; post = proj_out + x_in.
struct AttnDims {
  int ch, hw, nh, dh, dpad, ip;
  int64_t T;
};
cpd_status attn_dims(FwdCtx& F, int ch, AttnDims* d) {
  d->ch = ch;
  d->hw = F.h * F.w;
  d->T = (int64_t)F.R * d->hw;
  CPD_REQUIRE(d->T < (1ll << 31), "cpd_unet_forward: %lld tokens per evaluation exceed the GEMM row limit", (long long)d->T);
  heads_of(F.P->cfg, ch, &d->nh, &d->dh);
  d->dpad = round16(d->dh);
  d->ip = d->nh * d->dpad;
  return CPD_OK;
}

cpd_status attn_pre(FwdCtx& F, const std::string& p, Act x, void** hcur_out) {
  cpd_unet_plan* P = F.P;
  AttnDims D;
  PLAN_CHECK(attn_dims(F, x.c, &D));
  void *gn, *hcur;
  PLAN_CHECK(ws_get(P, "gn", D.T * D.ch, 2, &gn));
  PLAN_CHECK(groupnorm(P, F.st, x.p, nullptr, D.ch, 0, F.R, D.hw, p + "norm.g", p + "norm.b", 1e-6f, 0, F.stats, gn, x.gn));
  PLAN_CHECK(ws_get(P, "tr.h", D.T * D.ch, 2, &hcur));
  GemmOpt g;
  g.bias = (const float*)W(P, p + "proj_in.b");
  PLAN_CHECK(ln_setup(F, D.T, D.ch));
  ln_producer(F, g);
  PLAN_CHECK(gemm(P, F.st, gn, W(P, p + "proj_in.w"), hcur, 1, 1, (int)D.T, D.ch, D.ch, g));
  *hcur_out = hcur;
  return CPD_OK;
}

// x = attn1(LN1(x)) + x (attention.py:485-487)
cpd_status tblock_self(FwdCtx& F, const std::string& b, void* hcur, int ch) {
  cpd_unet_plan* P = F.P;
  AttnDims D;
  PLAN_CHECK(attn_dims(F, ch, &D));
  const int64_t T = D.T;
  const int ip = D.ip;
  void *ln, *qk, *vt, *o;
  PLAN_CHECK(ws_get(P, "tr.qk", T * 2 * ip, 2, &qk));
  PLAN_CHECK(ws_get(P, "tr.vt", (int64_t)ip * T, 2, &vt));
  if (F.ln_parts > 0 && T % 8 == 0) {
    // ONE projection Q | K | V of the raw residual stream: LN1 is applied by the epilogue (folded weights), the V columns leave
    // transposed (V^T [ip][T] is the attention kernels' operand) - no LayerNorm kernel, no second (transposed) GEMM
    GemmOpt g;
    ln_consumer(F, g, (const float*)W(P, b + "attn1.qkv.g"), (const float*)W(P, b + "attn1.qkv.bf"));
    g.ldd = 2 * ip;
    g.d_t = vt;
    g.dt_col0 = 2 * ip;
    g.ldd_t = T;
    PLAN_CHECK(gemm(P, F.st, hcur, W(P, b + "attn1.qkv.wf"), qk, 1, 1, (int)T, ch, 3 * ip, g));
  } else {
    PLAN_CHECK(ws_get(P, "tr.ln", T * ch, 2, &ln));
    PLAN_CHECK(layernorm(P, F.st, hcur, (int)T, ch, b + "norm1.g", b + "norm1.b", ln));
    PLAN_CHECK(gemm(P, F.st, ln, W(P, b + "attn1.qk.w"), qk, 1, 1, (int)T, ch, 2 * ip));
    PLAN_CHECK(gemm(P, F.st, W(P, b + "attn1.v.w"), ln, vt, 1, 1, ip, ch, (int)T));  // V^T = Wv LN(x)^T
  }
  PLAN_CHECK(ws_get(P, "tr.o", T * ip, 2, &o));
  PLAN_CHECK(attention(P, F.st, qk, 2 * ip, (const uint16_t*)qk + ip, 2 * ip, vt, (int)T, o, ip, F.R, D.nh, D.hw, D.hw, D.hw, D.dpad, D.dh, 0));
  GemmOpt g;
  g.bias = (const float*)W(P, b + "attn1.out.b");
  g.residual = hcur;
  g.ld_res = ch;
  ln_producer(F, g);  // LN2 reads these rows next
  return gemm(P, F.st, o, W(P, b + "attn1.out.w"), hcur, 1, 1, (int)T, ip, ch, g);
}

// x = attn2(LN2(x), context) + x (K / V^T cached per prompt); x = ff(LN3(x)) + x (attention.py:488-490, 92-118)
cpd_status tblock_rest(FwdCtx& F, const std::string& b, void* hcur, int ch) {
  cpd_unet_plan* P = F.P;
  AttnDims D;
  PLAN_CHECK(attn_dims(F, ch, &D));
  const int64_t T = D.T;
  const int ip = D.ip;
  CPD_REQUIRE(P->have_ctx, "cpd_unet_forward: no text context cached (call cpd_cache_context_kv first)");
  void *ln, *o, *kc, *vtc, *q2, *ff;
  PLAN_CHECK(ws_get(P, "tr.ln", T * ch, 2, &ln));
  PLAN_CHECK(ws_get(P, "tr.o", T * ip, 2, &o));
  PLAN_CHECK(ws_get(P, b + "kc", (int64_t)P->rc * P->nk_pad * ip, 2, &kc));
  PLAN_CHECK(ws_get(P, b + "vt", (int64_t)ip * P->rc * P->nk_pad, 2, &vtc));
  PLAN_CHECK(ws_get(P, "tr.q2", T * ip, 2, &q2));
  if (F.ln_parts > 0) {
    GemmOpt g;
    ln_consumer(F, g, (const float*)W(P, b + "attn2.q.g"), (const float*)W(P, b + "attn2.q.bf"));
    PLAN_CHECK(gemm(P, F.st, hcur, W(P, b + "attn2.q.wf"), q2, 1, 1, (int)T, ch, ip, g));
  } else {
    PLAN_CHECK(layernorm(P, F.st, hcur, (int)T, ch, b + "norm2.g", b + "norm2.b", ln));
    PLAN_CHECK(gemm(P, F.st, ln, W(P, b + "attn2.q.w"), q2, 1, 1, (int)T, ch, ip));
  }
  PLAN_CHECK(attention(P, F.st, q2, ip, kc, ip, vtc, P->rc * P->nk_pad, o, ip, F.R, D.nh, D.hw, P->ntok, P->nk_pad, D.dpad, D.dh, P->rc));
  {
    GemmOpt g;
    g.bias = (const float*)W(P, b + "attn2.out.b");
    g.residual = hcur;
    g.ld_res = ch;
    ln_producer(F, g);  // LN3 reads these rows next
    PLAN_CHECK(gemm(P, F.st, o, W(P, b + "attn2.out.w"), hcur, 1, 1, (int)T, ip, ch, g));
  }
  PLAN_CHECK(ws_get(P, "tr.ff", T * 4 * ch, 2, &ff));
  // (the GEGLU epilogue is the busy stage of ff1 when K = ch is small: at 320 channels the two extra FMAs per element cost
  // more (65536 x 2560 x 320: 130 -> 164 us) than the LayerNorm kernel they replace (20 us))
  if (F.ln_parts > 0 && ch >= P->fold_geglu_min_ch) {
    GemmOpt g;
    ln_consumer(F, g, (const float*)W(P, b + "ff1.g"), (const float*)W(P, b + "ff1.bf"));
    g.epilogue = CPD_EPI_GEGLU;
    PLAN_CHECK(gemm(P, F.st, hcur, W(P, b + "ff1.wf"), ff, 1, 1, (int)T, ch, 8 * ch, g));
  } else {
    PLAN_CHECK(layernorm(P, F.st, hcur, (int)T, ch, b + "norm3.g", b + "norm3.b", ln));
    GemmOpt g;
    g.bias = (const float*)W(P, b + "ff1.b");
    g.epilogue = CPD_EPI_GEGLU;
    PLAN_CHECK(gemm(P, F.st, ln, W(P, b + "ff1.w"), ff, 1, 1, (int)T, ch, 8 * ch, g));
  }
  GemmOpt g;
  g.bias = (const float*)W(P, b + "ff2.b");
  g.residual = hcur;
  g.ld_res = ch;
  ln_producer(F, g);  // LN1 of the next transformer block of this SpatialTransformer (depth > 1), if any
  return gemm(P, F.st, ff, W(P, b + "ff2.w"), hcur, 1, 1, (int)T, 4 * ch, ch, g);
}

cpd_status attn_post(FwdCtx& F, const std::string& p, Act x, void* hcur, Act* outp) {
  cpd_unet_plan* P = F.P;
  const int64_t T = (int64_t)F.R * F.h * F.w;
  void* out;
  PLAN_CHECK(ws_get(P, p + "out", T * x.c, 2, &out));
  GemmOpt g;
  g.bias = (const float*)W(P, p + "proj_out.b");
  g.residual = x.p;  // x + x_in (attention.py:537)
  g.ld_res = x.c;
  PLAN_CHECK(gemm(P, F.st, hcur, W(P, p + "proj_out.w"), out, 1, 1, (int)T, x.c, x.c, g));
  outp->p = out;
  outp->c = x.c;
  outp->gn = nullptr;
  return CPD_OK;
}

cpd_status attn_block(FwdCtx& F, const std::string& p, Act x, int depth, Act* outp) {
  void* hcur;
  PLAN_CHECK(attn_pre(F, p, x, &hcur));
  for (int d = 0; d < depth; ++d) {
    const std::string b = p + "transformer_blocks." + std::to_string(d) + ".";
    PLAN_CHECK(tblock_self(F, b, hcur, x.c));
    PLAN_CHECK(tblock_rest(F, b, hcur, x.c));
  }
  return attn_post(F, p, x, hcur, outp);
}

cpd_status expand_images(cpd_unet_plan* P, cudaStream_t st, const void* src, void* dst, int n_images, int reps, int64_t elems_per_image) {
  OpScope op(P, st, "expand", "", 0.0, 1);
  const int64_t per_image = elems_per_image * 2 / 16;
  int64_t blocks = (n_images * per_image + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  CPD_CUDA_CHECK(cpd_launch(expand_images_kernel, dim3((unsigned)blocks), dim3(256), 0, st, (const uint4*)src, (uint4*)dst, n_images, reps, per_image));
  return CPD_OK;
}

cpd_status run_block(FwdCtx& F, const Block& blk, Act hcur, Act skip, Act* outp) {
  cpd_unet_plan* P = F.P;
  for (size_t j = 0; j < blk.layers.size(); ++j) {
    const Layer& l = blk.layers[j];
    const std::string p = blk.prefix + std::to_string(j) + ".";
    if (l.kind == L_RES) {
      Act x1;
      if (skip.p && j == 0) x1 = skip;
      CPD_REQUIRE(hcur.c + x1.c == l.a, "cpd_unet_forward: %s expects %d input channels, got %d + %d", p.c_str(), l.a, hcur.c, x1.c);
      PLAN_CHECK(res_block(F, p, hcur, x1, l.b, &hcur));
    } else if (l.kind == L_ATTN) {
      PLAN_CHECK(attn_block(F, p, hcur, l.b, &hcur));
    } else if (l.kind == L_DOWN) {
      void* out;
      PLAN_CHECK(ws_get(P, p + "out", (int64_t)F.R * (F.h / 2) * (F.w / 2) * l.a, 2, &out));
      GemmOpt o;
      o.ksize = 3;
      o.stride = 2;
      o.bias = (const float*)W(P, p + "b");
      o.gn_out = gn_alloc(P, F.R, (F.h / 2) * (F.w / 2), l.a);
      PLAN_CHECK(gemm(P, F.st, hcur.p, W(P, p + "w"), out, F.R, F.h, F.w, l.a, l.a, o));
      hcur.p = out;
      hcur.gn = o.gn_out;
      F.h /= 2;
      F.w /= 2;
    } else if (l.kind == L_UP) {
      void *up, *out;
      PLAN_CHECK(ws_get(P, "up", (int64_t)F.R * 4 * F.h * F.w * l.a, 2, &up));
      {
        OpScope op(P, F.st, "upsample", "", 0.0, 1);
        PLAN_CHECK(cpd_upsample2x(hcur.p, F.R, F.h, F.w, l.a, up, F.st));
      }
      F.h *= 2;
      F.w *= 2;
      PLAN_CHECK(ws_get(P, p + "out", (int64_t)F.R * F.h * F.w * l.a, 2, &out));
      GemmOpt o;
      o.ksize = 3;
      o.bias = (const float*)W(P, p + "b");
      PLAN_CHECK(gemm(P, F.st, up, W(P, p + "w"), out, F.R, F.h, F.w, l.a, l.a, o));
      hcur.p = out;
      hcur.gn = nullptr;
    }
  }
  *outp = hcur;
  return CPD_OK;
}

cpd_status small_linear(cpd_unet_plan* P, cudaStream_t st, const void* x, int m, int k, const std::string& w, const std::string& b, int n,
                        int silu_in, float* out_f32, void* out_bf16, int ld_out) {
  OpScope op(P, st, "small", "", 0.0, 1);
  return cpd_small_linear(x, m, k, W(P, w), (const float*)W(P, b), n, silu_in, out_f32, out_bf16, ld_out, st);
}

// t_rows: fp32 [m] on the device (already rounded to the model dtype); y_rows: bf16 [m][adm] or NULL.  Returns fp32 [m][emb_total].
cpd_status embeddings(cpd_unet_plan* P, cudaStream_t st, const float* t_rows, int m, const void* y_rows, float** emb_all) {
  const int mc = P->mc, ted = P->ted;
  void *temb, *emb, *ea;
  PLAN_CHECK(ws_get(P, "temb", (int64_t)m * mc, 2, &temb));
  {
    OpScope op(P, st, "small", "", 0.0, 1);
    PLAN_CHECK(cpd_timestep_embedding(t_rows, m, mc, 0, temb, st));
  }
  PLAN_CHECK(ws_get(P, "emb", (int64_t)m * ted, 2, &emb));
  if (P->adm) {
    void* e1l1;  // [m][SiLU inputs of time_embed.2 | label_emb.0.2]
    PLAN_CHECK(ws_get(P, "e1l1", (int64_t)m * 2 * ted, 2, &e1l1));
    PLAN_CHECK(small_linear(P, st, temb, m, mc, "te0.w", "te0.b", ted, 0, nullptr, e1l1, 2 * ted));
    PLAN_CHECK(small_linear(P, st, y_rows, m, P->adm, "lab0.w", "lab0.b", ted, 0, nullptr, (uint16_t*)e1l1 + ted, 2 * ted));
    PLAN_CHECK(small_linear(P, st, e1l1, m, 2 * ted, "te2lab2.w", "te2lab2.b", ted, 1, nullptr, emb, ted));
  } else {
    void* e1;
    PLAN_CHECK(ws_get(P, "e1", (int64_t)m * ted, 2, &e1));
    PLAN_CHECK(small_linear(P, st, temb, m, mc, "te0.w", "te0.b", ted, 0, nullptr, e1, ted));
    PLAN_CHECK(small_linear(P, st, e1, m, ted, "te2.w", "te2.b", ted, 1, nullptr, emb, ted));
  }
  PLAN_CHECK(ws_get(P, "emb_all", (int64_t)m * P->emb_total, 4, &ea));
  PLAN_CHECK(small_linear(P, st, emb, m, ted, "emb_all.w", "emb_all.b", P->emb_total, 1, (float*)ea, nullptr, P->emb_total));
  *emb_all = (float*)ea;
  return CPD_OK;
}

cpd_status forward_impl(cpd_unet_plan* P, const cpd_unet_io* io, void* eps_out, cudaStream_t st) {
  const cpd_unet_config& c = P->cfg;
  const int B = io->n_images, rpi = io->rows_per_image, R = B * rpi;
  int h = io->h, w = io->w;
  bool shared_t = io->t_count == 1;
  const float* t_rows = io->t;
  void* stats;
  PLAN_CHECK(ws_get(P, "gn.stats", (int64_t)R * 64 * CPD_GN_MAX_CHUNKS, 8, &stats));
  P->gn_used = 0;
  if (P->gn_arena.p) CPD_CUDA_CHECK(cudaMemsetAsync(P->gn_arena.p, 0, (size_t)P->gn_arena.numel, st));  // one node per forward
  float* emb_all = nullptr;
  int m_emb = shared_t ? 1 : R;
  if (P->adm) {
    // vector conditioning differs per conditioning row: one embedding row per UNet row (t repeated when shared)
    CPD_REQUIRE(P->ry > 0, "cpd_unet_forward: no vector conditioning set (call cpd_unet_set_vector first)");
    CPD_REQUIRE(R <= 16, "cpd_unet_forward: more than 16 UNet rows per evaluation with vector conditioning (shard the batch)");
    CPD_REQUIRE(R % P->ry == 0, "cpd_unet_forward: %d UNet rows are not a multiple of the %d vector-conditioning rows", R, P->ry);
    void* y_rows;
    PLAN_CHECK(ws_get(P, "y_rows", (int64_t)R * P->adm, 2, &y_rows));
    {
      OpScope op(P, st, "small", "", 0.0, 1);
      CPD_CUDA_CHECK(cpd_launch(repeat_rows_kernel, dim3(32), dim3(256), 0, st, (const void*)P->y.p, y_rows, R, P->ry, P->adm / 2));
    }
    if (shared_t) {
      void* t_all;
      PLAN_CHECK(ws_get(P, "t_rows", R, 4, &t_all));
      OpScope op(P, st, "small", "", 0.0, 1);
      CPD_CUDA_CHECK(cpd_launch(repeat_rows_kernel, dim3(1), dim3(32), 0, st, (const void*)t_rows, t_all, R, 1, 1));
      t_rows = (const float*)t_all;
      shared_t = false;
    }
    m_emb = R;
    PLAN_CHECK(embeddings(P, st, t_rows, m_emb, y_rows, &emb_all));
  } else {
    CPD_REQUIRE(m_emb <= 32, "cpd_unet_forward: more than 32 rows with distinct timesteps (use a shared timestep: t_count = 1)");
    PLAN_CHECK(embeddings(P, st, t_rows, m_emb, nullptr, &emb_all));
  }
  FwdCtx F{P, st, R, h, w, emb_all, shared_t ? 0 : P->emb_total, (double*)stats};
  const int mc = P->mc;
  void* h0;
  PLAN_CHECK(ws_get(P, "input_blocks.0.0.out", (int64_t)R * h * w * mc, 2, &h0));
  // The rows of an image share x * c_in and t (denoiser.py:383-393) and differ only in their text context, so every activation
  // in front of the FIRST cross-attention - the input conv, the ResBlock of input block 1 and, inside its SpatialTransformer,
  // GroupNorm, proj_in, LN1 and the whole self-attention (at 64 x 64 the most expensive attention of the network) - is
  // identical for the (1 + N) rows of an image.  It is evaluated ONCE per image and broadcast to the rows right before the
  // cross-attention; the result is bit-identical to evaluating every row.  (Not with per-row timesteps, vector conditioning or
  // injected tensors, where the rows do differ.)
  const bool share = rpi > 1 && shared_t && !P->adm && !io->inject_skips && !io->inject_feats && P->inputs.size() > 1 &&
                     P->inputs[1].layers.size() == 2 && P->inputs[1].layers[0].kind == L_RES && P->inputs[1].layers[1].kind == L_ATTN &&
                     P->share_prefix;
  if (!share) {
    OpScope op(P, st, "conv_in", "", 0.0, 1);
    PLAN_CHECK(cpd_conv_in(io->x, B, c.in_channels, h, w, W(P, "input_blocks.0.0.w"), (const float*)W(P, "input_blocks.0.0.b"), mc, 1.0f,
                           io->c_in, rpi, c.act_fp16, h0, st));
  }
  struct Skip {
    Act a;
    int h, w;
  };
  std::vector<Skip> hs;
  P->tap_in.assign(P->inputs.size(), Tap());
  P->tap_out.assign(P->outputs.size(), Tap());
  Act hcur{h0, mc};
  hs.push_back({hcur, h, w});
  P->tap_in[0] = Tap{h0, mc, h, w};
  size_t first_block = 1;
  if (share) {
    const Block& blk = P->inputs[1];
    const std::string pr = blk.prefix + "0.", pa = blk.prefix + "1.";
    const int cout = blk.layers[0].b, depth = blk.layers[1].b;
    const int64_t hw = (int64_t)h * w;
    void* h0b;
    PLAN_CHECK(ws_get(P, "input_blocks.0.0.out", (int64_t)B * hw * mc, 2, &h0b));
    {
      OpScope op(P, st, "conv_in", "", 0.0, 1);
      PLAN_CHECK(cpd_conv_in(io->x, B, c.in_channels, h, w, W(P, "input_blocks.0.0.w"), (const float*)W(P, "input_blocks.0.0.b"), mc, 1.0f,
                             io->c_in, 1, c.act_fp16, h0b, st));
    }
    // the R-row input-conv output (the skip tensor of the last output block) is the per-image one repeated: a copy, not a second conv
    PLAN_CHECK(expand_images(P, st, h0b, h0, B, rpi, hw * mc));
    FwdCtx Fb{P, st, B, h, w, emb_all, 0, (double*)stats};
    Act xres_b, xres;
    PLAN_CHECK(res_block(Fb, pr, Act{h0b, mc}, Act(), cout, &xres_b));
    void *hb, *hr, *xr;
    PLAN_CHECK(attn_pre(Fb, pa, xres_b, &hb));
    const std::string b0 = pa + "transformer_blocks.0.";
    PLAN_CHECK(tblock_self(Fb, b0, hb, cout));
    PLAN_CHECK(ws_get(P, pr + "out", (int64_t)R * hw * cout, 2, &xr));
    PLAN_CHECK(expand_images(P, st, xres_b.p, xr, B, rpi, hw * cout));
    PLAN_CHECK(ws_get(P, "tr.h", (int64_t)R * hw * cout, 2, &hr));
    PLAN_CHECK(expand_images(P, st, hb, hr, B, rpi, hw * cout));
    xres = Act{xr, cout};
    // the partial sums of the per-image rows are repeated like the rows themselves ([part][image] -> [part][image][row]), so that
    // LN2 of this block is folded exactly as it is when every row is evaluated (results independent of the sharing)
    PLAN_CHECK(ln_setup(F, (int64_t)R * hw, cout));
    if (F.lnp && Fb.lnp && Fb.ln_parts > 0) {
      PLAN_CHECK(expand_images(P, st, Fb.lnp, F.lnp, Fb.ln_parts * B, rpi, hw * 4));  // float2 per row = four 16-bit units
      F.ln_parts = Fb.ln_parts;
    }
    PLAN_CHECK(tblock_rest(F, b0, hr, cout));
    for (int d = 1; d < depth; ++d) {
      const std::string bd = pa + "transformer_blocks." + std::to_string(d) + ".";
      PLAN_CHECK(tblock_self(F, bd, hr, cout));
      PLAN_CHECK(tblock_rest(F, bd, hr, cout));
    }
    PLAN_CHECK(attn_post(F, pa, xres, hr, &hcur));
    hs.push_back({hcur, F.h, F.w});
    P->tap_in[1] = Tap{hcur.p, hcur.c, F.h, F.w};
    first_block = 2;
  }
  for (size_t i = first_block; i < P->inputs.size(); ++i) {
    PLAN_CHECK(run_block(F, P->inputs[i], hcur, Act(), &hcur));
    hs.push_back({hcur, F.h, F.w});
    P->tap_in[i] = Tap{hcur.p, hcur.c, F.h, F.w};
  }
  PLAN_CHECK(run_block(F, P->middle, hcur, Act(), &hcur));
  P->tap_mid = Tap{hcur.p, hcur.c, F.h, F.w};
  for (size_t i = 0; i < P->outputs.size(); ++i) {
    Skip s = hs.back();
    hs.pop_back();
    CPD_REQUIRE(s.h == F.h && s.w == F.w, "cpd_unet_forward: skip tensor %zu is %dx%d, the decoder is at %dx%d (h, w must be multiples of %d)", i, s.h,
                s.w, F.h, F.w, 1 << (c.n_levels - 1));
    if (io->inject_skips && io->inject_skips[i]) {  // unet.py:806-809
      s.a.p = const_cast<void*>(io->inject_skips[i]);
      s.a.gn = nullptr;
    }
    if (io->inject_feats && io->inject_feats[i]) {  // unet.py:810-813
      hcur.p = const_cast<void*>(io->inject_feats[i]);
      hcur.gn = nullptr;
    }
    PLAN_CHECK(run_block(F, P->outputs[i], hcur, s.a, &hcur));
    P->tap_out[i] = Tap{hcur.p, hcur.c, F.h, F.w};
  }
  void* gn;
  PLAN_CHECK(ws_get(P, "gn", (int64_t)R * F.h * F.w * hcur.c, 2, &gn));
  PLAN_CHECK(groupnorm(P, st, hcur.p, nullptr, hcur.c, 0, R, F.h * F.w, "out.gn.g", "out.gn.b", 1e-5f, 1, (double*)stats, gn, hcur.gn));
  P->gn_need = P->gn_used;
  {
    OpScope op(P, st, "conv_out", "", 0.0, 1);
    PLAN_CHECK(cpd_conv_out(gn, R, F.h, F.w, hcur.c, W(P, "out.w"), (const float*)W(P, "out.b"), c.out_channels, eps_out, c.eps_dtype,
                            c.act_fp16, st));
  }
  return CPD_OK;
}

bool stream_capturing(cudaStream_t st) {
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cs) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return cs != cudaStreamCaptureStatusNone;
}

int eps_elt(const cpd_unet_config& c) { return c.eps_dtype == CPD_F32 ? 4 : 2; }

}  // namespace

// ------------------------------------------------------------------------------------------------ C ABI
extern "C" cpd_status cpd_unet_plan_create(const cpd_unet_config* cfg, cpd_unet_plan** plan) {
  CPD_REQUIRE(cfg && plan, "cpd_unet_plan_create: null argument");
  CPD_REQUIRE(cfg->n_levels >= 1 && cfg->n_levels <= CPD_UNET_MAX_LEVELS && cfg->n_attention_resolutions >= 0 &&
                  cfg->n_attention_resolutions <= CPD_UNET_MAX_LEVELS,
              "cpd_unet_plan_create: n_levels / n_attention_resolutions out of range");
  CPD_REQUIRE(cfg->model_channels > 0 && cfg->model_channels % 64 == 0, "cpd_unet_plan_create: model_channels=%d must be a multiple of 64",
              cfg->model_channels);
  CPD_REQUIRE(cfg->in_channels > 0 && cfg->in_channels <= 8 && cfg->out_channels > 0 && cfg->out_channels <= 8,
              "cpd_unet_plan_create: in / out channels must be in 1..8");
  CPD_REQUIRE(cfg->num_res_blocks >= 1 && cfg->context_dim > 0 && cfg->context_dim % 64 == 0,
              "cpd_unet_plan_create: num_res_blocks >= 1 and context_dim a multiple of 64 are required");
  CPD_REQUIRE(cfg->num_head_channels == -1 ? cfg->num_heads > 0 : cfg->num_head_channels > 0, "cpd_unet_plan_create: bad head configuration");
  CPD_REQUIRE(cfg->eps_dtype == CPD_F32 || cfg->eps_dtype == CPD_BF16, "cpd_unet_plan_create: eps_dtype must be CPD_F32 or CPD_BF16");
  CPD_REQUIRE(cfg->adm_in_channels >= 0 && cfg->adm_in_channels % 2 == 0, "cpd_unet_plan_create: adm_in_channels must be even");
  cpd_unet_plan* P = new cpd_unet_plan();
  P->cfg = *cfg;
  CPD_CUDA_CHECK(cudaGetDevice(&P->device));
  P->mc = cfg->model_channels;
  P->ted = 4 * cfg->model_channels;
  P->adm = cfg->adm_in_channels;
  {
    const char* e = getenv("CPD_UNET_SHARE_PREFIX");
    P->share_prefix = !(e && e[0] == '0');
    e = getenv("CPD_UNET_FOLD_LN");
    P->fold_ln = !(e && e[0] == '0');
    e = getenv("CPD_UNET_FOLD_MIN_HW");
    if (e && atoi(e) > 0) P->fold_min_hw = atoi(e);
    e = getenv("CPD_UNET_GN_STATS");
    P->gn_stats = e && e[0] == '1';
    e = getenv("CPD_UNET_FOLD_GEGLU_MIN_CH");
    if (e && atoi(e) > 0) P->fold_geglu_min_ch = atoi(e);
  }
  enumerate_blocks(P);
  cpd_status st = build_specs(P);
  if (st == CPD_OK) st = alloc_dev(&P->splitk, 32ll * 1024 * 1024, 4, false);  // fp32 partial tiles of split-K launches
  if (st != CPD_OK) {
    cpd_unet_plan_destroy(P);
    return st;
  }
  *plan = P;
  return CPD_OK;
}

extern "C" void cpd_unet_plan_destroy(cpd_unet_plan* P) {
  if (!P) return;
  DeviceGuard g(P->device);
  cudaDeviceSynchronize();
  for (auto& kv : P->graphs)
    if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
  for (auto& kv : P->w)
    if (kv.second.p) cudaFree(kv.second.p);
  for (auto& kv : P->ws)
    if (kv.second.p) cudaFree(kv.second.p);
  if (P->y.p) cudaFree(P->y.p);
  if (P->tune_scratch.p) cudaFree(P->tune_scratch.p);
  if (P->splitk.p) cudaFree(P->splitk.p);
  if (P->gn_arena.p) cudaFree(P->gn_arena.p);
  for (void* q : P->gn_retired) cudaFree(q);
  if (P->cap_stream) cudaStreamDestroy(P->cap_stream);
  for (auto& r : P->prof) {
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  delete P;
}

extern "C" cpd_status cpd_pack_weights(cpd_unet_plan* P, const char* name, const void* data, int dtype, int64_t numel, int on_device) {
  CPD_REQUIRE(P && name && data, "cpd_pack_weights: null argument");
  CPD_REQUIRE(dtype == CPD_F32 || dtype == CPD_F16 || dtype == CPD_BF16, "cpd_pack_weights: dtype %d", dtype);
  DeviceGuard g(P->device);
  auto it = P->spec.find(name);
  CPD_REQUIRE(it != P->spec.end(), "cpd_pack_weights: '%s' is not a parameter of this UNet configuration", name);
  WeightSpec& s = it->second;
  CPD_REQUIRE(numel == s.numel, "cpd_pack_weights: '%s' has %lld elements, expected %lld", name, (long long)numel, (long long)s.numel);
  const void* src = data;
  void* tmp = nullptr;
  const size_t bytes = (size_t)numel * (dtype == CPD_F32 ? 4 : 2);
  if (!on_device) {
    CPD_CUDA_CHECK(cudaMalloc(&tmp, bytes));
    CPD_CUDA_CHECK(cudaMemcpy(tmp, data, bytes, cudaMemcpyHostToDevice));
    src = tmp;
  }
  cpd_status rc = CPD_OK;
  for (size_t i = 0; i < s.jobs.size() && rc == CPD_OK; ++i) {
    const PackJob& j = s.jobs[i];
    if (j.accumulate && s.loaded) continue;  // a reload must not add the bias twice (reloading summed biases is not supported)
    void* dst = W(P, s.dst[i]);
    const int64_t total = j.dst_rows * j.dst_cols;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    pack_kernel<<<blocks, 256>>>(src, dtype, dst, j);
    if (cudaGetLastError() != cudaSuccess) rc = CPD_ERR_CUDA;
  }
  if (cudaDeviceSynchronize() != cudaSuccess) rc = CPD_ERR_CUDA;
  if (tmp) cudaFree(tmp);
  if (rc != CPD_OK) {
    cpd_set_error("cpd_pack_weights: packing '%s' failed: %s", name, cudaGetErrorString(cudaGetLastError()));
    return rc;
  }
  s.loaded = true;
  P->fold_dirty = true;
  // the packed weights changed: captured graphs stay valid (same addresses), the context cache does not (its K / V^T were
  // computed from the old to_k / to_v)
  if (strstr(name, "attn2.to_k") || strstr(name, "attn2.to_v")) P->have_ctx = false;
  return CPD_OK;
}

extern "C" int cpd_unet_plan_missing_weights(cpd_unet_plan* P) {
  if (!P) return -1;
  int missing = 0;
  for (const auto& kv : P->spec)
    if (!kv.second.loaded) {
      if (!missing) cpd_set_error("missing state_dict entry '%s'", kv.first.c_str());
      ++missing;
    }
  return missing;
}

extern "C" cpd_status cpd_cache_context_kv(cpd_unet_plan* P, const void* context, int dtype, int rows, int tokens, void* stream) {
  CPD_REQUIRE(P && context, "cpd_cache_context_kv: null argument");
  CPD_REQUIRE(rows > 0 && tokens > 0, "cpd_cache_context_kv: rows=%d tokens=%d", rows, tokens);
  CPD_REQUIRE(dtype == CPD_F32 || dtype == CPD_F16 || dtype == CPD_BF16, "cpd_cache_context_kv: dtype %d", dtype);
  CPD_REQUIRE(cpd_unet_plan_missing_weights(P) == 0, "cpd_cache_context_kv: %s", cpd_last_error());
  DeviceGuard g(P->device);
  cudaStream_t st = (cudaStream_t)stream;
  CPD_REQUIRE(!stream_capturing(st), "cpd_cache_context_kv: must not run inside a stream capture");
  const int D = P->cfg.context_dim;
  const int nk_pad = round16(tokens);
  P->launches = 0;
  P->capturing = false;
  P->allow_alloc = true;
  void* ctx_pad;
  PLAN_CHECK(ws_get(P, "ctx_pad", (int64_t)rows * nk_pad * D, 2, &ctx_pad));
  {
    const int64_t total = (int64_t)rows * nk_pad * D;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    ctx_pack_kernel<<<blocks, 256, 0, st>>>(context, dtype, ctx_pad, P->cfg.act_fp16 ? CPD_F16 : CPD_BF16, rows, tokens, nk_pad, D);
    CPD_CUDA_CHECK(cudaGetLastError());
    P->launches += 1;
  }
  auto layer_kv = [&](const Block& blk) -> cpd_status {
    for (size_t j = 0; j < blk.layers.size(); ++j) {
      const Layer& l = blk.layers[j];
      if (l.kind != L_ATTN) continue;
      int nh, dh;
      heads_of(P->cfg, l.a, &nh, &dh);
      const int ip = nh * round16(dh);
      for (int d = 0; d < l.b; ++d) {
        const std::string b = blk.prefix + std::to_string(j) + ".transformer_blocks." + std::to_string(d) + ".";
        void *kc, *vt;  // persistent buffers (stable addresses: captured graphs stay valid across prompts)
        PLAN_CHECK(ws_get(P, b + "kc", (int64_t)rows * nk_pad * ip, 2, &kc));
        PLAN_CHECK(ws_get(P, b + "vt", (int64_t)ip * rows * nk_pad, 2, &vt));
        PLAN_CHECK(gemm(P, st, ctx_pad, W(P, b + "attn2.k.w"), kc, 1, 1, rows * nk_pad, D, ip));
        PLAN_CHECK(gemm(P, st, W(P, b + "attn2.v.w"), ctx_pad, vt, 1, 1, ip, D, rows * nk_pad));
      }
    }
    return CPD_OK;
  };
  for (const Block& b : P->inputs) PLAN_CHECK(layer_kv(b));
  PLAN_CHECK(layer_kv(P->middle));
  for (const Block& b : P->outputs) PLAN_CHECK(layer_kv(b));
  P->rc = rows;
  P->ntok = tokens;
  P->nk_pad = nk_pad;
  P->have_ctx = true;
  return CPD_OK;
}

extern "C" cpd_status cpd_unet_set_vector(cpd_unet_plan* P, const void* y, int dtype, int rows, void* stream) {
  CPD_REQUIRE(P && y, "cpd_unet_set_vector: null argument");
  CPD_REQUIRE(P->adm > 0, "cpd_unet_set_vector: this UNet has no vector conditioning (adm_in_channels = 0)");
  CPD_REQUIRE(rows > 0 && rows <= 16, "cpd_unet_set_vector: rows=%d must be in 1..16", rows);
  DeviceGuard g(P->device);
  cudaStream_t st = (cudaStream_t)stream;
  if (P->y.numel < (int64_t)16 * P->adm) {
    CPD_REQUIRE(!stream_capturing(st), "cpd_unet_set_vector: first call must not run inside a stream capture");
    PLAN_CHECK(alloc_dev(&P->y, (int64_t)16 * P->adm, 2, true));  // 16 rows: one stable address for every prompt layout
  }
  PackJob j = SpecBuilder::plain(rows, P->adm, CPD_BF16);
  pack_kernel<<<64, 256, 0, st>>>(y, dtype, P->y.p, j);
  CPD_CUDA_CHECK(cudaGetLastError());
  P->ry = rows;
  return CPD_OK;
}

extern "C" cpd_status cpd_unet_forward(cpd_unet_plan* P, const cpd_unet_io* io, void* stream) {
  CPD_REQUIRE(P && io && io->x && io->t, "cpd_unet_forward: null argument");
  CPD_REQUIRE(io->n_images > 0 && io->rows_per_image > 0 && io->h > 0 && io->w > 0, "cpd_unet_forward: bad shape");
  const int R = io->n_images * io->rows_per_image;
  CPD_REQUIRE(io->t_count == 1 || io->t_count == R, "cpd_unet_forward: t_count=%d must be 1 or n_images * rows_per_image = %d", io->t_count, R);
  const int div = 1 << (P->cfg.n_levels - 1);
  CPD_REQUIRE(io->h % div == 0 && io->w % div == 0, "cpd_unet_forward: h=%d, w=%d must be multiples of %d", io->h, io->w, div);
  CPD_REQUIRE(cpd_unet_plan_missing_weights(P) == 0, "cpd_unet_forward: %s", cpd_last_error());
  DeviceGuard g(P->device);
  cudaStream_t st = (cudaStream_t)stream;
  P->launches = 0;
  const bool outer_capture = stream_capturing(st);
  if (P->fold_dirty) {
    CPD_REQUIRE(!outer_capture, "cpd_unet_forward: the first evaluation after cpd_pack_weights must not run inside a stream capture");
    PLAN_CHECK(ensure_folded(P, st));
  }
  const bool inject = io->inject_skips != nullptr || io->inject_feats != nullptr;
  void* eps = io->eps;
  P->allow_alloc = !outer_capture;
  if (!eps) PLAN_CHECK(ws_get(P, "eps", (int64_t)R * P->cfg.out_channels * io->h * io->w, eps_elt(P->cfg), &eps));
  P->allow_alloc = true;
  const bool use_graph = P->cfg.use_cuda_graph && !io->no_graph && !inject && !outer_capture && !P->profile;
  // The arena of GroupNorm accumulators is sized from the demand the previous pass over this shape recorded; a pass that finds
  // it too small still computes correctly (its GroupNorms reduce their own statistics).
  auto grow_gn_arena = [&]() -> cpd_status {
    if (P->gn_need <= P->gn_arena.numel) return CPD_OK;
    if (P->gn_arena.p) P->gn_retired.push_back(P->gn_arena.p);  // captured graphs may still point into the old arena
    P->gn_arena.p = nullptr;
    return alloc_dev(&P->gn_arena, P->gn_need, 1, true);
  };
  if (!use_graph) {
    P->capturing = outer_capture;
    P->allow_alloc = !outer_capture;
    cpd_status rc = forward_impl(P, io, eps, st);
    if (rc == CPD_OK && !outer_capture && P->gn_need > P->gn_arena.numel) {
      // first evaluation of a larger shape: size the arena and evaluate again, so that the launches that use it are tuned before
      // a host captures this shape into a graph of its own (samplers/k_diffusion.py: one eager evaluation, then the capture)
      CPD_CUDA_CHECK(cudaStreamSynchronize(st));
      rc = grow_gn_arena();
      if (rc == CPD_OK) rc = forward_impl(P, io, eps, st);
    }
    P->capturing = false;
    P->allow_alloc = true;
    return rc;
  }
  // One CUDA graph per (shape, rows, context layout, operand addresses): the ~850 kernel launches of an evaluation are
  // replayed with a single cudaGraphLaunch.  x, c_in, t and eps are baked in by ADDRESS: hosts keep them in stable buffers
  // (the Python host's static x / scalar buffers, a sampler that updates x in place) and only their contents change.
  GraphKey key;
  memset(&key, 0, sizeof(key));
  key.n = io->n_images; key.h = io->h; key.w = io->w; key.rpi = io->rows_per_image; key.t_count = io->t_count;
  key.rc = P->rc; key.ntok = P->ntok; key.ry = P->ry;
  key.x = io->x; key.c_in = io->c_in; key.t = io->t; key.eps = eps;
  auto it = P->graphs.find(key);
  if (it == P->graphs.end()) {
    if (P->graphs.size() >= 16) {  // hosts that pass fresh pointers every call would otherwise leak graphs: drop the oldest
      auto old = P->graphs.begin();
      for (auto k = P->graphs.begin(); k != P->graphs.end(); ++k)
        if (k->second.last_use < old->second.last_use) old = k;
      cudaGraphExecDestroy(old->second.exec);
      P->graphs.erase(old);
    }
    // eager warm-up: allocates every workspace buffer, times the GEMM variants, opts the kernels in to their shared memory
    P->capturing = false;
    P->allow_alloc = true;
    PLAN_CHECK(forward_impl(P, io, eps, st));
    CPD_CUDA_CHECK(cudaStreamSynchronize(st));
    if (P->gn_need > P->gn_arena.numel) {  // first evaluation of a larger shape: size the arena, then tune the launches that use it
      PLAN_CHECK(grow_gn_arena());
      PLAN_CHECK(forward_impl(P, io, eps, st));
      CPD_CUDA_CHECK(cudaStreamSynchronize(st));
    }
    P->launches = 0;
    P->capturing = true;
    P->allow_alloc = false;
    cudaGraph_t graph = nullptr;
    // captured on a stream of the plan's own: the caller's stream may be the legacy default stream, which cannot capture
    if (!P->cap_stream) CPD_CUDA_CHECK(cudaStreamCreateWithFlags(&P->cap_stream, cudaStreamNonBlocking));
    CPD_CUDA_CHECK(cudaStreamBeginCapture(P->cap_stream, cudaStreamCaptureModeThreadLocal));
    const cpd_status rc = forward_impl(P, io, eps, P->cap_stream);
    const cudaError_t ce = cudaStreamEndCapture(P->cap_stream, &graph);
    P->capturing = false;
    P->allow_alloc = true;
    if (rc != CPD_OK) {
      if (graph) cudaGraphDestroy(graph);
      return rc;
    }
    CPD_CUDA_CHECK(ce);
    GraphEntry e;
    const cudaError_t ie = cudaGraphInstantiate(&e.exec, graph, 0);
    cudaGraphDestroy(graph);
    CPD_CUDA_CHECK(ie);
    e.launches = P->launches;
    it = P->graphs.emplace(key, e).first;
  }
  it->second.last_use = ++P->use_counter;
  CPD_CUDA_CHECK(cudaGraphLaunch(it->second.exec, st));
  P->launches = it->second.launches;
  return CPD_OK;
}

extern "C" cpd_status cpd_unet_plan_buffer(cpd_unet_plan* P, const char* name, void** ptr, int64_t* numel) {
  CPD_REQUIRE(P && name && ptr, "cpd_unet_plan_buffer: null argument");
  auto wi = P->w.find(name);
  if (wi != P->w.end()) {
    *ptr = wi->second.p;
    if (numel) *numel = wi->second.numel;
    return CPD_OK;
  }
  const std::string prefix = std::string(name) + "#";
  const DevBuf* best = nullptr;
  for (const auto& kv : P->ws)
    if (kv.first.compare(0, prefix.size(), prefix) == 0 && (!best || kv.second.numel > best->numel)) best = &kv.second;
  CPD_REQUIRE(best != nullptr, "cpd_unet_plan_buffer: no buffer named '%s'", name);
  *ptr = best->p;
  if (numel) *numel = best->numel;
  return CPD_OK;
}

extern "C" int64_t cpd_unet_plan_launches(cpd_unet_plan* P) { return P ? P->launches : 0; }

// Output of block `index` of the last forward: which = 0 input blocks (the skip tensors of return_attn, unet.py:802-804), 1 the
// middle block, 2 output blocks (the features of return_feat, :816-817).  NHWC activations [rows][h][w][channels].
extern "C" cpd_status cpd_unet_plan_tap(cpd_unet_plan* P, int which, int index, void** ptr, int* channels, int* h, int* w) {
  CPD_REQUIRE(P && ptr, "cpd_unet_plan_tap: null argument");
  const Tap* t = nullptr;
  if (which == 0 && index >= 0 && index < (int)P->tap_in.size()) t = &P->tap_in[index];
  if (which == 1) t = &P->tap_mid;
  if (which == 2 && index >= 0 && index < (int)P->tap_out.size()) t = &P->tap_out[index];
  CPD_REQUIRE(t && t->p, "cpd_unet_plan_tap: no such block output (which=%d index=%d); run a forward first", which, index);
  *ptr = t->p;
  if (channels) *channels = t->c;
  if (h) *h = t->h;
  if (w) *w = t->w;
  return CPD_OK;
}

extern "C" int cpd_unet_plan_blocks(cpd_unet_plan* P, int which) {
  if (!P) return 0;
  return which == 0 ? (int)P->inputs.size() : (which == 2 ? (int)P->outputs.size() : 1);
}

// Per-launch CUDA-event timing of eager forwards (bench.py's roofline leg, tools/profile_layers.py): on = 1 clears the
// record list and brackets every following kernel-level call (graphs are bypassed while it is on).
extern "C" void cpd_unet_plan_set_profile(cpd_unet_plan* P, int on) {
  if (!P) return;
  for (auto& r : P->prof) {
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  P->prof.clear();
  P->profile = on != 0;
}

// "kind\tlabel\tmicroseconds\tflops" lines of the recorded launches (synchronises the device); returns the bytes needed.
extern "C" int64_t cpd_unet_plan_profile_dump(cpd_unet_plan* P, char* buf, int64_t cap) {
  if (!P) return 0;
  DeviceGuard g(P->device);
  cudaDeviceSynchronize();
  std::string out;
  char line[256];
  for (auto& r : P->prof) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, r.e0, r.e1);
    snprintf(line, sizeof(line), "%s\t%s\t%.3f\t%.6g\n", r.kind, r.label.c_str(), ms * 1e3f, r.flops);
    out += line;
  }
  if (buf && cap > 0) {
    const size_t n = out.size() < (size_t)cap - 1 ? out.size() : (size_t)cap - 1;
    memcpy(buf, out.data(), n);
    buf[n] = 0;
  }
  return (int64_t)out.size() + 1;
}
