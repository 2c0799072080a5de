// C-ABI glue: error reporting, device check and the bde_gemm dispatcher.
#include <stdarg.h>
#include "common.cuh"
#include <stdlib.h>

namespace bde {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

}  // namespace bde

#include <mutex>
#include <set>
#include <utility>
namespace bde {
bool first_use_on_device(const void* key) {
  static std::mutex mu;
  static std::set<std::pair<int, const void*>> seen;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(mu);
  return seen.insert(std::make_pair(dev, key)).second;
}
bool pdl_enabled() {
  const char* e = getenv("BDE2VID_PDL");
  return !(e != nullptr && e[0] == '0');
}

int device_sm_count() {
  static int cache[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return kNumSMs;
  if (cache[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = kNumSMs;
    cache[dev] = n;     // (a benign race: every writer stores the same value)
  }
  return cache[dev];
}
}  // namespace bde

using namespace bde;

extern "C" const char* bde_last_error(void) { return g_err; }

extern "C" int bde_abi_version(void) { return BDE_ABI_VERSION; }

extern "C" int bde_device_ok(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    set_error("cudaGetDevice: %s", cudaGetErrorString(e));
    return -2;
  }
  int major = 0, minor = 0;
  e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (e != cudaSuccess) {
    set_error("cudaDeviceGetAttribute: %s", cudaGetErrorString(e));
    return -2;
  }
  // the library holds sm_100a code only (arch-specific: not forward compatible, e.g. with sm_103)
  if (major == 10 && minor == 0) return 1;
  set_error("device is compute capability %d.%d; libbde2vid_sm100.so holds sm_100a code only", major, minor);
  return 0;
}

// ---- optional per-launch timing of the tcgen05 GEMM kernel (used by bench.py's roofline measurement) ----
// CUDA events are recorded on the launching stream right around the kernel, from C, so that the pair
// brackets nothing but the launch itself.
#include <vector>
namespace {
std::vector<cudaEvent_t> g_prof_ev;  // start/stop pairs
size_t g_prof_used = 0;
bool g_prof_on = false;
double g_prof_flops = 0.0;  // 2 M N K of the profiled launches (K without padding)
}  // namespace

extern "C" int bde_profile_begin(int max_launches) {
  for (cudaEvent_t e : g_prof_ev) cudaEventDestroy(e);
  g_prof_ev.clear();
  g_prof_used = 0;
  g_prof_on = false;
  g_prof_flops = 0.0;
  if (max_launches <= 0) return 0;
  g_prof_ev.resize((size_t)max_launches * 2);
  for (auto& e : g_prof_ev)
    if (cudaEventCreate(&e) != cudaSuccess) {
      set_error("bde_profile_begin: cudaEventCreate failed");
      return -2;
    }
  g_prof_on = true;
  return 0;
}

extern "C" int bde_profile_end(double* total_ms, int* n_launches) {
  g_prof_on = false;
  double tot = 0.0;
  const size_t n = g_prof_used / 2;
  for (size_t i = 0; i < n; ++i) {
    float ms = 0.f;
    cudaEventSynchronize(g_prof_ev[2 * i + 1]);
    if (cudaEventElapsedTime(&ms, g_prof_ev[2 * i], g_prof_ev[2 * i + 1]) != cudaSuccess) {
      set_error("bde_profile_end: cudaEventElapsedTime failed");
      return -2;
    }
    tot += ms;
  }
  if (total_ms != nullptr) *total_ms = tot;
  if (n_launches != nullptr) *n_launches = (int)n;
  return 0;
}

extern "C" int bde_profile_flops(double* flops) {
  if (flops != nullptr) *flops = g_prof_flops;
  return 0;
}

extern "C" int bde_gemm(const bde_gemm_desc* d, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  BDE_REQUIRE(d != nullptr, "bde_gemm: null descriptor");
  BDE_REQUIRE(d->dtype == BDE_F32 || d->dtype == BDE_BF16, "bde_gemm: bad dtype %d", d->dtype);
  BDE_REQUIRE((d->a0 != nullptr || d->ln_mode != 0) && d->w != nullptr && d->out != nullptr, "bde_gemm: null operand");
  BDE_REQUIRE(d->ln_mode == 0 || d->engine == BDE_ENGINE_TCGEN05, "bde_gemm: ln_mode is only implemented by the tcgen05 engine");
  BDE_REQUIRE(d->c0 > 0 && d->c1 >= 0 && (d->c1 == 0 || d->a1 != nullptr), "bde_gemm: bad sources");
  BDE_REQUIRE(d->ksize >= 1 && d->stride >= 1 && d->pad >= 0 && d->n > 0, "bde_gemm: bad conv geometry");
  BDE_REQUIRE(d->h_out == (d->h_in + 2 * d->pad - d->ksize) / d->stride + 1 &&
                  d->w_out == (d->w_in + 2 * d->pad - d->ksize) / d->stride + 1,
              "bde_gemm: output size %dx%d inconsistent with input %dx%d k%d s%d p%d", d->h_out, d->w_out, d->h_in,
              d->w_in, d->ksize, d->stride, d->pad);
  BDE_REQUIRE((size_t)d->n_img * d->h_out * d->w_out < ((size_t)1 << 31), "bde_gemm: M overflows int32");
  if (d->epi == BDE_EPI_LSTM) {
    BDE_REQUIRE(d->n % 4 == 0 && d->c_out != nullptr, "bde_gemm: LSTM epilogue needs n %% 4 == 0 and c_out");
  } else if (d->epi == BDE_EPI_SCATTER) {
    BDE_REQUIRE(d->row_map != nullptr, "bde_gemm: SCATTER epilogue needs row_map");
  } else if (d->epi == BDE_EPI_GRU_UR) {
    BDE_REQUIRE(d->n % 4 == 0 && d->c_out != nullptr, "bde_gemm: GRU_UR epilogue needs n %% 4 == 0 and c_out (u)");
  } else if (d->epi == BDE_EPI_GRU_OUT) {
    BDE_REQUIRE(d->c_out != nullptr && d->residual != nullptr, "bde_gemm: GRU_OUT epilogue needs c_out (h') and residual (u)");
  } else {
    BDE_REQUIRE(d->epi == BDE_EPI_STORE, "bde_gemm: unknown epilogue %d", d->epi);
  }
  BDE_REQUIRE(d->a0_ld == 0 || d->a0_ld >= d->c0, "bde_gemm: a0_ld smaller than c0");
  BDE_REQUIRE(d->a1_ld == 0 || d->a1_ld >= d->c1, "bde_gemm: a1_ld smaller than c1");
  const bool pitched = (d->a0_ld != 0 && d->a0_ld != d->c0) || (d->c1 > 0 && d->a1_ld != 0 && d->a1_ld != d->c1);
  BDE_REQUIRE(!pitched || d->engine == BDE_ENGINE_TCGEN05, "bde_gemm: pitched A operands need the tcgen05 engine");
  if (d->engine == BDE_ENGINE_SIMT) return gemm_simt(d, s);
  if (d->engine == BDE_ENGINE_TCGEN05) {
    const bool prof = g_prof_on && g_prof_used + 2 <= g_prof_ev.size();
    if (prof) cudaEventRecord(g_prof_ev[g_prof_used], s);
    const int rc = gemm_tcgen05(d, s);
    if (prof) {
      cudaEventRecord(g_prof_ev[g_prof_used + 1], s);
      g_prof_used += 2;
      g_prof_flops += 2.0 * d->n_img * d->h_out * d->w_out * (double)d->n * d->ksize * d->ksize * (d->c0 + d->c1);
    }
    return rc;
  }
  BDE_REQUIRE(false, "bde_gemm: unknown engine %d", d->engine);
}
