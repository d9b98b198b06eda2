// Micro-benchmark: HBM read bandwidth of an NHWC fp32 map [700,800,32] when every pass reads one
// 32-byte sector of each 128-byte pixel (what the 8-channel chunks of the correlation kernel do)
// versus 64 or 128 bytes per pixel per pass. Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cuda_runtime.h>
constexpr int H = 700, W = 800, C = 32;
// mode: bytes per pixel per pass (32, 64, 128); tiles of 8 rows x 64 px per CTA-iteration like the kernel
template <int BYTES>
__global__ void reader(const float4 *__restrict__ map, float *__restrict__ sink, int n_tiles, int tiles_x) {
  constexpr int V = BYTES / 16;            // float4 per pixel per pass
  constexpr int PASSES = 128 / BYTES;
  float acc = 0.f;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int tx = tile % tiles_x, ty = tile / tiles_x;
    for (int pass = 0; pass < PASSES; ++pass) {
      // 8 rows x 64 px x V float4
      for (int i = threadIdx.x; i < 8 * 64 * V; i += blockDim.x) {
        const int v = i % V, px = (i / V) % 64, r = i / (V * 64);
        const int y = ty * 8 + r, x = tx * 64 + px;
        if (y < H && x < W) {
          const float4 t = __ldcg(map + (static_cast<size_t>(y) * W + x) * (C / 4) + pass * V + v);
          acc += t.x + t.y + t.z + t.w;
        }
      }
    }
  }
  if (acc == 123.456f) sink[0] = acc;
}
template <int BYTES>
float run(const float4 *map, float *sink, int grid, int threads) {
  const int tiles_x = (W + 63) / 64, n_tiles = tiles_x * ((H + 7) / 8);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 3; ++i) reader<BYTES><<<grid, threads>>>(map, sink, n_tiles, tiles_x);
  cudaEventRecord(a);
  const int reps = 20;
  for (int i = 0; i < reps; ++i) reader<BYTES><<<grid, threads>>>(map + (i % 3) * (size_t)H * W * C / 4, sink, n_tiles, tiles_x);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  return ms * 1e3f / reps;
}
int main() {
  float4 *map; float *sink;
  cudaMalloc(&map, 3 * sizeof(float) * H * W * C); cudaMalloc(&sink, 4);
  cudaMemset(map, 0, 3 * sizeof(float) * H * W * C);
  const double mb = 4.0 * H * W * C / 1e6;
  for (int grid : {148, 296, 592}) for (int threads : {128, 256}) {
    const float t32 = run<32>(map, sink, grid, threads), t64 = run<64>(map, sink, grid, threads), t128 = run<128>(map, sink, grid, threads);
    printf("grid %4d x %3d thr: 32B/pass %.1f us (%.0f GB/s)  64B/pass %.1f us (%.0f GB/s)  128B/pass %.1f us (%.0f GB/s)\n", grid, threads,
           t32, mb / t32 * 1e3, t64, mb / t64 * 1e3, t128, mb / t128 * 1e3);
  }
  return 0;
}
