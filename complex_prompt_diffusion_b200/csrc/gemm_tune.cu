// cpd_gemm_conv: variant 0 ("auto") resolved through a per-problem-shape table of TIMED tile variants.
//
// The tile shape that runs a layer fastest depends on how its tile count fills the 74 SM pairs and on how its K depth
// amortises the epilogue; the cost model of gemm_umma2.cu (pick_tiles) gets most shapes right but not all (profiles/
// r01_tuned_variants.txt).  So the first time a shape is seen - outside stream capture, and only when the caller lends a scratch
// output (`tune_scratch`) - every applicable variant is launched on the real operands with the output redirected to the
// scratch (in-place residual launches stay idempotent), timed with CUDA events (minimum over interleaved rounds), and the
// winner is remembered.  Round 1 did this in the Python host (ops._tune_gemm); it lives behind the C ABI now so that the
// plan-level entry points (unet_plan.cu) and non-Python hosts get the same kernels.  CPD_GEMM_AUTOTUNE=0 switches timing off.
#include <stdlib.h>
#include <string.h>

#include <array>
#include <map>
#include <mutex>
#include <string>

#include "../../include/cpd_b200.h"
#include "common.cuh"

cpd_status cpd_gemm_conv_dispatch(const cpd_gemm_params* p, void* stream);  // gemm_umma.cu

namespace {

typedef std::array<int, 14> TuneKey;
std::map<TuneKey, int> g_table;
std::mutex g_mu;

TuneKey make_key(const cpd_gemm_params* p) {
  return TuneKey{p->n_img, p->h_in, p->w_in, p->c0, p->c1, p->n_out, p->ksize, p->stride, p->epilogue,
                 p->epilogue == CPD_EPI_GEGLU ? p->geglu_block : 0,
                 (p->residual != nullptr) | ((p->ln_sums_out != nullptr) << 1) | ((p->ln_sums != nullptr) << 2) | ((p->d_t != nullptr) << 3) |
                     ((p->gn_sums_out != nullptr) << 4),
                 p->rowvec != nullptr, p->a_fp16, p->m_valid};
}

int g_autotune = -1;  // -1: not yet read from the environment
bool autotune_on() {
  if (g_autotune < 0) {
    const char* e = getenv("CPD_GEMM_AUTOTUNE");
    g_autotune = (e && e[0] == '0') ? 0 : 1;
  }
  return g_autotune != 0;
}

// pair-kernel tile widths, two-sub-tile tiles (2000 + BN), the one-tile-per-CTA kernels (1, 2); split-K x tile for small-M layers
const int kCandidates[] = {160, 128, 96, 192, 224, 256, 64, 2160, 2128, 2256, 2096, 2192, 1, 2};
const int kSplitK[] = {20160, 30160, 40160, 22160, 32160, 42160, 20128, 40128};
constexpr int kRounds = 3, kReps = 6;  // (2 x 5 left near-ties to chance: the same shape picked different variants from run to run)

int tune(const cpd_gemm_params* p, cudaStream_t st) {
  cpd_gemm_params q = *p;
  q.d = p->tune_scratch;
  int cands[32];
  int n = 0;
  for (int v : kCandidates) cands[n++] = v;
  const int64_t rows = (int64_t)p->n_img * (p->h_in / p->stride) * (p->w_in / p->stride);
  const int k_iters = p->ksize * p->ksize * (p->c0 + p->c1) / 64;
  if (((rows + 255) / 256) * ((p->n_out + 159) / 160) <= 40 && k_iters >= 32 && p->splitk_ws)
    for (int v : kSplitK) cands[n++] = v;  // too few tiles for 74 SM pairs: also try split-K
  float best_ms[32];
  bool ok[32];
  for (int i = 0; i < n; ++i) {
    best_ms[i] = 1e30f;
    ok[i] = false;
  }
  cudaEvent_t e0, e1;
  if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) return 0;
  for (int rnd = 0; rnd < kRounds; ++rnd) {  // candidates are often within a few per cent: minimum over interleaved rounds
    for (int i = 0; i < n; ++i) {
      const int v = cands[i];
      if (v % 1000 >= p->n_out + 32 && v % 1000 > 64) continue;  // tile much wider than N: all padding
      if (rnd && !ok[i]) continue;
      q.variant = v;
      if (cpd_gemm_conv_dispatch(&q, st) != CPD_OK) continue;  // variant not applicable to this shape
      cudaEventRecord(e0, st);
      for (int r = 0; r < kReps; ++r) cpd_gemm_conv_dispatch(&q, st);
      cudaEventRecord(e1, st);
      if (cudaEventSynchronize(e1) != cudaSuccess) {
        cudaGetLastError();
        continue;
      }
      float ms = 0.f;
      cudaEventElapsedTime(&ms, e0, e1);
      ok[i] = true;
      if (ms / kReps < best_ms[i]) best_ms[i] = ms / kReps;
    }
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cpd_set_error("%s", "");  // inapplicable candidates left their messages behind
  int best = 0;
  float best_t = 1e30f;
  for (int i = 0; i < n; ++i)  // ties go to the earlier candidate
    if (ok[i] && best_ms[i] < best_t) {
      best_t = best_ms[i];
      best = cands[i];
    }
  static int dbg = -1;
  if (dbg < 0) dbg = getenv("CPD_GEMM_DEBUG") ? 1 : 0;
  if (dbg)
    fprintf(stderr, "tuned gemm rows=%lld N=%d K=%d epi=%d res=%d: variant %d (%.1f us)\n", (long long)rows, p->n_out,
            p->ksize * p->ksize * (p->c0 + p->c1), p->epilogue, p->residual != nullptr, best, best_t * 1e3f);
  return best;
}

}  // namespace

extern "C" cpd_status cpd_gemm_conv(const cpd_gemm_params* p, void* stream) {
  CPD_REQUIRE(p != nullptr, "cpd_gemm_conv: null params");
  if (p->variant != 0) return cpd_gemm_conv_dispatch(p, stream);
  cpd_gemm_params q = *p;
  if (p->epilogue == CPD_EPI_GEGLU) {  // the interleave block of the packed weights fixes the tile
    q.variant = p->geglu_block ? p->geglu_block : 128;
    return cpd_gemm_conv_dispatch(&q, stream);
  }
  const TuneKey key = make_key(p);
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_table.find(key);
    if (it != g_table.end()) {
      q.variant = it->second;
      return cpd_gemm_conv_dispatch(&q, stream);
    }
  }
  const int64_t rows = (int64_t)p->n_img * (p->h_in / (p->stride > 0 ? p->stride : 1)) * (p->w_in / (p->stride > 0 ? p->stride : 1));
  if (autotune_on() && p->tune_scratch && p->tune_scratch_bytes >= rows * (int64_t)p->ldd * 2) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing((cudaStream_t)stream, &cs) == cudaSuccess && cs == cudaStreamCaptureStatusNone) {
      // validate with the cost-model variant first (output in the scratch: an in-place residual launch must run exactly ONCE
      // on the real output): a bad problem must fail with its own message, not be "tuned"
      cpd_gemm_params v = *p;
      v.d = p->tune_scratch;
      const cpd_status st0 = cpd_gemm_conv_dispatch(&v, stream);
      if (st0 != CPD_OK) return st0;
      const int best = tune(p, (cudaStream_t)stream);
      if (p->gn_sums_out)  // every timed launch added its statistics to the accumulators: start the real launch from zero
        CPD_CUDA_CHECK(cudaMemsetAsync(p->gn_sums_out, 0, (size_t)p->n_img * p->n_out * 2 * sizeof(long long), (cudaStream_t)stream));
      std::lock_guard<std::mutex> lk(g_mu);
      g_table[key] = best;
      q.variant = best;
      return cpd_gemm_conv_dispatch(&q, stream);
    }
  }
  return cpd_gemm_conv_dispatch(p, stream);  // cost model (pick_tiles)
}

// Runtime switch of the per-shape timing (the environment variable CPD_GEMM_AUTOTUNE only sets the initial state) and a way to
// forget the table: jobs that need the SAME tile variants for every batch size (bit-identical images whatever they are batched
// with) switch the timing off and clear the table, or import one table everywhere.
extern "C" void cpd_gemm_set_autotune(int on) { g_autotune = on ? 1 : 0; }
extern "C" void cpd_gemm_tune_clear(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  g_table.clear();
}

extern "C" int64_t cpd_gemm_tune_export(char* buf, int64_t cap) {
  std::string out;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    for (const auto& kv : g_table) {
      for (size_t i = 0; i < kv.first.size(); ++i) {
        out += std::to_string(kv.first[i]);
        out += (i + 1 < kv.first.size()) ? ',' : '=';
      }
      out += std::to_string(kv.second);
      out += '\n';
    }
  }
  if (buf && cap > 0) {
    const size_t n = out.size() < (size_t)cap - 1 ? out.size() : (size_t)cap - 1;
    memcpy(buf, out.data(), n);
    buf[n] = 0;
  }
  return (int64_t)out.size() + 1;
}

extern "C" int cpd_gemm_tune_import(const char* text) {
  if (!text) return -1;
  int count = 0;
  const char* s = text;
  std::lock_guard<std::mutex> lk(g_mu);
  while (*s) {
    TuneKey key;
    char* end = nullptr;
    bool good = true;
    for (size_t i = 0; i < key.size() && good; ++i) {
      key[i] = (int)strtol(s, &end, 10);
      good = end != s && *end == (i + 1 < key.size() ? ',' : '=');
      s = good ? end + 1 : s;
    }
    if (!good) return -1;
    const int v = (int)strtol(s, &end, 10);
    if (end == s) return -1;
    g_table[key] = v;
    ++count;
    s = end;
    while (*s == '\n' || *s == '\r' || *s == ' ') ++s;
  }
  return count;
}
