// TEST INFRASTRUCTURE (oracle/): the three TensorFlow 1.3 helpers the reference's
// correlation_grad_kernel.cu.cc uses (CUDA_1D_KERNEL_LOOP, CudaLaunchConfig, GetCudaLaunchConfig),
// restated from their published definition (tensorflow/core/util/cuda_kernel_helper.h, v1.3):
// a grid-stride loop and "as many threads as the device holds, 1024 per block, at most one block
// per SM". The launch shape does not change any result: every loop index is independent.
#pragma once
#include <algorithm>
#include "third_party/eigen3/unsupported/Eigen/CXX11/Tensor"

#define CUDA_1D_KERNEL_LOOP(i, n) \
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < (n); i += blockDim.x * gridDim.x)

namespace tensorflow {
struct CudaLaunchConfig {
  int virtual_thread_count = -1;
  int thread_per_block     = -1;
  int block_count          = -1;
};

inline CudaLaunchConfig GetCudaLaunchConfig(int work_element_count, const Eigen::GpuDevice& d) {
  CudaLaunchConfig config;
  const int virtual_thread_count  = work_element_count;
  const int physical_thread_count = std::min(
    d.getNumCudaMultiProcessors() * d.maxCudaThreadsPerMultiProcessor(), virtual_thread_count);
  const int thread_per_block = std::min(1024, d.maxCudaThreadsPerBlock());
  const int block_count      = std::min(
    (physical_thread_count + thread_per_block - 1) / thread_per_block, d.getNumCudaMultiProcessors());
  config.virtual_thread_count = virtual_thread_count;
  config.thread_per_block     = thread_per_block;
  config.block_count          = block_count;
  return config;
}
}  // namespace tensorflow
