// C-ABI glue: error reporting, device check and the bde_gemm dispatcher.
#include <stdarg.h>
#include "common.cuh"

namespace bde {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
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
  int major = 0;
  e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess) {
    set_error("cudaDeviceGetAttribute: %s", cudaGetErrorString(e));
    return -2;
  }
  return major == 10 ? 1 : 0;
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
  } else {
    BDE_REQUIRE(d->epi == BDE_EPI_STORE, "bde_gemm: unknown epilogue %d", d->epi);
  }
  if (d->engine == BDE_ENGINE_SIMT) return gemm_simt(d, s);
  if (d->engine == BDE_ENGINE_TCGEN05) return gemm_tcgen05(d, s);
  BDE_REQUIRE(false, "bde_gemm: unknown engine %d", d->engine);
}
