// Library-level entry points and error plumbing of libvistracker_sm100a.so.
#include <stdarg.h>
#include "common.cuh"
#include "vt_internal.h"

namespace vt {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("%s: CUDA error %d (%s)", what, (int)e, cudaGetErrorString(e));
  return -3;
}

}  // namespace vt

extern "C" {

const char* vt_last_error(void) { return vt::g_err; }
int vt_version(void) { return 1; }
/* cudaMemsetAsync on the caller's stream (a memset node when captured into a CUDA graph) */
int vt_zero(void* p, long long bytes, void* stream) {
  if (bytes <= 0) return 0;
  cudaError_t e = cudaMemsetAsync(p, 0, (size_t)bytes, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : vt::cuda_fail(e, "vt_zero");
}
int vt_compiled_arch(void) {
#ifdef VT_ARCH
  return VT_ARCH;
#else
  return 0;
#endif
}

}  // extern "C"
