// TEST INFRASTRUCTURE (oracle/): extern "C" driver around the reference's OWN correlation kernels.
//
// oracle/build_oracle.py compiles /root/reference/avod/core/ops/correlation/{correlation_kernel,
// pad,correlation_grad_kernel}.cu.cc UNMODIFIED (where they lie, against oracle/ref_stubs/) and links
// them with this file into oracle/_ref/libcorr_ref.so. This file restates only the host-side shape
// and padding logic of the TensorFlow OpKernels, which cannot be compiled without TensorFlow:
//   forward : correlation_kernel.cc:26-124      (Pad a, Pad b -> Correlation)
//   backward: correlation_grad_kernel.cc:28-151 (Pad a, Pad b -> CorrelationGradA, CorrelationGradB)
// Buffers are HOST pointers; the driver owns its device memory. Kernel times (CUDA events on the
// launching stream) come back in ms[]: the survey's stated bar for S4 is this naive kernel.
// Only tests/, smoke() and bench.py's cpu/reference legs may load the library.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>

#include "correlation_kernel.h"
#include "pad.h"

using tensorflow::GPUDevice;

namespace {
struct Shape {
  int padded_h, padded_w, kernel_radius, border, out_h, out_w, grid_radius, grid_width, out_c;
};

// correlation_kernel.cc:39-57 / correlation_grad_kernel.cc:43-58
Shape shape_of(int H, int W, int ks, int md, int s1, int s2, int pad) {
  Shape s;
  s.padded_h      = H + 2 * pad;
  s.padded_w      = W + 2 * pad;
  s.kernel_radius = (ks - 1) / 2;
  s.border        = md + s.kernel_radius;
  s.out_h         = (int)ceil((float)(s.padded_h - s.border * 2) / (float)s1);
  s.out_w         = (int)ceil((float)(s.padded_w - s.border * 2) / (float)s1);
  s.grid_radius   = md / s2;
  s.grid_width    = s.grid_radius * 2 + 1;
  s.out_c         = s.grid_width * s.grid_width;
  return s;
}

#define CK(x)                                                                        \
  do {                                                                               \
    cudaError_t e_ = (x);                                                            \
    if (e_ != cudaSuccess) {                                                         \
      fprintf(stderr, "ref_corr_driver: %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      return -3;                                                                     \
    }                                                                                \
  } while (0)
}  // namespace

extern "C" int ref_correlation_out_shape(int H, int W, int ks, int md, int s1, int s2, int pad, int* out_hwc) {
  if (ks % 2 == 0) return -1;  // correlation_kernel.cc:23
  Shape s = shape_of(H, W, ks, md, s1, s2, pad);
  if (s.out_h < 1 || s.out_w < 1) return -2;  // correlation_kernel.cc:50-53
  out_hwc[0] = s.out_h;
  out_hwc[1] = s.out_w;
  out_hwc[2] = s.out_c;
  return 0;
}

// ms[0] = both PadData launches (+ their memsets), ms[1] = CorrelateData, averaged over `reps` runs
extern "C" int ref_correlation(const float* a, const float* b, int N, int H, int W, int C, int ks, int md, int s1,
                               int s2, int pad, float* out, float* ms, int reps) {
  int hwc[3];
  int rc = ref_correlation_out_shape(H, W, ks, md, s1, s2, pad, hwc);
  if (rc) return rc;
  Shape  s        = shape_of(H, W, ks, md, s1, s2, pad);
  size_t in_elems = (size_t)N * H * W * C, pad_elems = (size_t)N * s.padded_h * s.padded_w * C;
  size_t out_elems = (size_t)N * s.out_h * s.out_w * s.out_c;
  float *da, *db, *pa, *pb, *dout;
  CK(cudaMalloc(&da, in_elems * 4));
  CK(cudaMalloc(&db, in_elems * 4));
  CK(cudaMalloc(&pa, pad_elems * 4));
  CK(cudaMalloc(&pb, pad_elems * 4));
  CK(cudaMalloc(&dout, out_elems * 4));
  CK(cudaMemcpy(da, a, in_elems * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db, b, in_elems * 4, cudaMemcpyHostToDevice));
  GPUDevice   dev(0);
  cudaEvent_t e0, e1, e2;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  CK(cudaEventCreate(&e2));
  if (reps < 1) reps = 1;
  float t_pad = 0.f, t_corr = 0.f;
  for (int r = 0; r < reps; r++) {
    CK(cudaEventRecord(e0, dev.stream()));
    tensorflow::Pad(dev, da, N, H, W, C, s.padded_h, s.padded_w, pa);
    tensorflow::Pad(dev, db, N, H, W, C, s.padded_h, s.padded_w, pb);
    CK(cudaEventRecord(e1, dev.stream()));
    tensorflow::Correlation(dev, pa, pb, N, s.out_h, s.out_w, s.out_c, s.out_h * s.out_w * s.out_c, s.padded_h,
                            s.padded_w, C, md, s.grid_radius, s.grid_width, s.kernel_radius, ks, s1, s2, dout);
    CK(cudaEventRecord(e2, dev.stream()));
    CK(cudaEventSynchronize(e2));
    float x;
    CK(cudaEventElapsedTime(&x, e0, e1));
    t_pad += x;
    CK(cudaEventElapsedTime(&x, e1, e2));
    t_corr += x;
  }
  CK(cudaGetLastError());
  if (ms) {
    ms[0] = t_pad / reps;
    ms[1] = t_corr / reps;
  }
  CK(cudaMemcpy(out, dout, out_elems * 4, cudaMemcpyDeviceToHost));
  cudaFree(da); cudaFree(db); cudaFree(pa); cudaFree(pb); cudaFree(dout);
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2);
  return 0;
}

// ms[0] = pads, ms[1] = CorrelateDataBackward0 launches, ms[2] = CorrelateDataBackward1 launches
extern "C" int ref_correlation_grad(const float* g, const float* a, const float* b, int N, int H, int W, int C,
                                    int ks, int md, int s1, int s2, int pad, float* ga, float* gb, float* ms,
                                    int reps) {
  int hwc[3];
  int rc = ref_correlation_out_shape(H, W, ks, md, s1, s2, pad, hwc);
  if (rc) return rc;
  Shape  s        = shape_of(H, W, ks, md, s1, s2, pad);
  size_t in_elems = (size_t)N * H * W * C, pad_elems = (size_t)N * s.padded_h * s.padded_w * C;
  size_t out_elems = (size_t)N * s.out_h * s.out_w * s.out_c;
  float *da, *db, *pa, *pb, *dg, *dga, *dgb;
  CK(cudaMalloc(&da, in_elems * 4));
  CK(cudaMalloc(&db, in_elems * 4));
  CK(cudaMalloc(&pa, pad_elems * 4));
  CK(cudaMalloc(&pb, pad_elems * 4));
  CK(cudaMalloc(&dg, out_elems * 4));
  CK(cudaMalloc(&dga, in_elems * 4));
  CK(cudaMalloc(&dgb, in_elems * 4));
  CK(cudaMemcpy(da, a, in_elems * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db, b, in_elems * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dg, g, out_elems * 4, cudaMemcpyHostToDevice));
  GPUDevice   dev(0);
  cudaEvent_t e0, e1, e2, e3;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  CK(cudaEventCreate(&e2));
  CK(cudaEventCreate(&e3));
  if (reps < 1) reps = 1;
  float t[3] = {0.f, 0.f, 0.f};
  const int in_count = H * W * C;
  for (int r = 0; r < reps; r++) {
    CK(cudaEventRecord(e0, dev.stream()));
    tensorflow::Pad(dev, da, N, H, W, C, s.padded_h, s.padded_w, pa);
    tensorflow::Pad(dev, db, N, H, W, C, s.padded_h, s.padded_w, pb);
    CK(cudaEventRecord(e1, dev.stream()));
    tensorflow::CorrelationGradA(dev, N, s.out_w, s.out_h, s.out_c, md, s.grid_radius, s.grid_width, s.kernel_radius,
                                 s1, s2, W, H, s.padded_w, s.padded_h, C, in_count, pad, pb, dg, dga);
    CK(cudaEventRecord(e2, dev.stream()));
    tensorflow::CorrelationGradB(dev, N, s.out_w, s.out_h, s.out_c, md, s.grid_radius, s.grid_width, s.kernel_radius,
                                 s1, s2, W, H, s.padded_w, s.padded_h, C, in_count, pad, pa, dg, dgb);
    CK(cudaEventRecord(e3, dev.stream()));
    CK(cudaEventSynchronize(e3));
    float x;
    CK(cudaEventElapsedTime(&x, e0, e1)); t[0] += x;
    CK(cudaEventElapsedTime(&x, e1, e2)); t[1] += x;
    CK(cudaEventElapsedTime(&x, e2, e3)); t[2] += x;
  }
  CK(cudaGetLastError());
  if (ms)
    for (int i = 0; i < 3; i++) ms[i] = t[i] / reps;
  CK(cudaMemcpy(ga, dga, in_elems * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(gb, dgb, in_elems * 4, cudaMemcpyDeviceToHost));
  cudaFree(da); cudaFree(db); cudaFree(pa); cudaFree(pb); cudaFree(dg); cudaFree(dga); cudaFree(dgb);
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2); cudaEventDestroy(e3);
  return 0;
}
