// Micro-benchmark: cost of LDS.128 when pairs of lanes of the same quarter-warp read the SAME
// 16-byte address (the "row-skew" idea for the correlation kernel).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o lds_broadcast lds_broadcast.cu
// mode 0: 32 distinct addresses (conflict-free)            -> 4 wavefronts expected
// mode 1: lanes l and l^2 share an address (16 unique, pairs inside a quarter-warp)
// mode 2: lanes l and l^16 share an address (16 unique, pairs in different quarter-warps)
// mode 3: lanes l and l^1 share (16 unique, neighbours)
// mode 4: all 32 lanes read one address
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k(int mode, int iters, float *out, long long *cycles) {
  __shared__ float4 buf[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) buf[i] = make_float4(i, i + 1, i + 2, i + 3);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int idx;
  switch (mode) {
    case 0: idx = lane; break;
    case 1: idx = lane & ~2; break;
    case 2: idx = lane & 15; break;
    case 3: idx = lane & ~1; break;
    default: idx = 0; break;
  }
  idx += warp * 32;
  float4 acc = make_float4(0, 0, 0, 0);
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const float4 v = buf[(idx + u * 64 + it) & 1023];
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

int main() {
  float *out; long long *cyc;
  cudaMalloc(&out, 148 * 256 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 4096;
  for (int mode = 0; mode < 5; ++mode) {
    k<<<148, 256>>>(mode, iters, out, cyc);
    k<<<148, 256>>>(mode, iters, out, cyc);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
    // 8 warps x iters x 16 LDS.128 per SM
    printf("mode %d: %.2f cycles per warp-level LDS.128 (SM-wide, 8 warps)\n", mode, avg / (8.0 * iters * 16));
  }
  return 0;
}
