// Error bookkeeping and version entry points of libdodt_fe.so.
#include <stdio.h>
#include <string.h>

#include <atomic>

#include "common.cuh"

namespace dodt {

static thread_local char g_cuda_error[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_cuda_error(cudaError_t e, const char *what) {
  snprintf(g_cuda_error, sizeof(g_cuda_error), "%s: %s (%s)", what, cudaGetErrorName(e),
           cudaGetErrorString(e));
  // clear the sticky-less error so that the next call starts clean
  (void)cudaGetLastError();
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

}  // namespace dodt

extern "C" {

const char *dodt_strerror(int code) {
  switch (code) {
    case DODT_OK: return "ok";
    case DODT_EINVAL: return "invalid argument";
    case DODT_ESHAPE: return "inconsistent or empty shape";
    case DODT_ECAPACITY: return "workspace or key capacity exceeded";
    case DODT_ECUDA: return "CUDA runtime error";
    case DODT_EALIGN: return "misaligned pointer";
    default: return "unknown error";
  }
}

const char *dodt_last_cuda_error(void) { return dodt::g_cuda_error; }

int dodt_version(void) { return DODT_FE_VERSION; }

int64_t dodt_launch_count(void) { return dodt::g_launches.load(std::memory_order_relaxed); }

}  // extern "C"
