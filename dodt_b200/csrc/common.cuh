// Shared helpers for libdodt_fe.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "dodt_fe.h"

namespace dodt {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

void set_cuda_error(cudaError_t e, const char *what);
void count_launch(int n = 1);

inline cudaStream_t as_stream(dodt_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

inline int ceil_div(int64_t a, int64_t b) { return static_cast<int>((a + b - 1) / b); }

}  // namespace dodt

// Measurement knobs. The PRODUCT library (libdodt_fe.so) is built without DODT_DIAG: every knob is
// its compile-time default and no environment variable changes what a kernel computes or how it is
// launched. tools/ build a separate libdodt_fe_diag.so (python -m dodt_b200._build --diag) in which
// the knobs read the environment and the diagnostic kernel instantiations exist.
#ifdef DODT_DIAG
#include <stdlib.h>
namespace dodt {
inline int diag_knob(const char *name, int dflt) {
  const char *e = getenv(name);
  return e ? atoi(e) : dflt;
}
}  // namespace dodt
#define DODT_KNOB(name, dflt) ::dodt::diag_knob(name, dflt)
#else
#define DODT_KNOB(name, dflt) (dflt)
#endif

#define DODT_CUDA_TRY(expr)                       \
  do {                                            \
    cudaError_t _e = (expr);                      \
    if (_e != cudaSuccess) {                      \
      ::dodt::set_cuda_error(_e, #expr);          \
      return DODT_ECUDA;                          \
    }                                             \
  } while (0)

// after a <<<>>> launch: record it and pick up launch-configuration errors
#define DODT_AFTER_LAUNCH()                       \
  do {                                            \
    ::dodt::count_launch();                       \
    DODT_CUDA_TRY(cudaPeekAtLastError());         \
  } while (0)
