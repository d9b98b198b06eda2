// Micro-benchmark (round 2): how fast can one SM be FED with tiles of an NHWC fp32 map
// [700,800,32] by the mechanisms the correlation kernel could use? Same bytes for every mode:
// 36 864-byte slabs (the B tile with halo + the A tile of an 8x64-pixel output tile = 6 slabs),
// three-slab ring per CTA, 128 threads, two CTAs per SM, nothing but the copies and one shared-
// memory read per thread and slab.
//   mode 0  cp.async 16 B, 32-byte pieces: an 8-channel chunk of 16 rows x 72 px  (round-1 loader)
//   mode 1  cp.async 16 B, 64-byte pieces: a 16-channel chunk of 8 rows x 72 px
//   mode 2  cp.async 16 B, whole 128-byte pixels: 4 rows x 72 px
//   mode 3  TMA tensor, 128-byte inner box (32 ch, 72 px, 4 rows), SWIZZLE_128B
//   mode 4  TMA tensor,  64-byte inner box (16 ch, 72 px, 8 rows), SWIZZLE_64B
//   mode 5  TMA tensor,  32-byte inner box ( 8 ch, 72 px, 16 rows), SWIZZLE_32B
//   mode 6  cp.async.bulk 1-D: four row segments of 72 px x 128 B
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o feed_bench feed_bench.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

constexpr int H = 700, W = 800, C = 32;
constexpr int kSlab = 36864, kNST = 3, kThreads = 128;
constexpr int kTilesX = 13, kTilesY = 88, kParts = 6;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, int c0, int c1, int c2, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

struct Slab { int x0, y0, c0; };   // first pixel column / row (may be negative), first channel

template <int MODE>
__device__ __forceinline__ Slab slab_of(int s) {
  const int tile = s / kParts, part = s % kParts;
  const int tx = tile % kTilesX, ty = tile / kTilesX;
  Slab r;
  r.x0 = tx * 64 - 4;
  if (MODE == 0 || MODE == 5) {          // 8-channel chunk, 16 rows: parts 0..3 = chunks of the B tile, 4..5 = A chunks (over-counted to 16 rows)
    r.y0 = ty * 8 - 4; r.c0 = (part & 3) * 8;
  } else if (MODE == 1 || MODE == 4) {   // 16-channel chunk, 8 rows
    r.y0 = ty * 8 - 4 + (part >> 1 & 1) * 8; r.c0 = (part & 1) * 16;
    if (part >= 4) r.y0 = ty * 8;
  } else {                               // all channels, 4 rows
    r.y0 = part < 4 ? ty * 8 - 4 + part * 4 : ty * 8 + (part - 4) * 4; r.c0 = 0;
  }
  return r;
}

template <int MODE>
__global__ void __launch_bounds__(kThreads, 2)
feed(const float *__restrict__ map, const __grid_constant__ CUtensorMap tmap, int n_slabs, float *sink) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t base = smem_u32(smem);
  const uint32_t bars = base + kNST * kSlab;
  constexpr bool kTma = MODE >= 3;
  if (kTma && threadIdx.x == 0) {
    for (int i = 0; i < kNST; ++i) mbar_init(bars + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int my = (n_slabs - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);

  auto issue = [&](int k) {
    const Slab sl = slab_of<MODE>(blockIdx.x + k * gridDim.x);
    const uint32_t dst = base + (k % kNST) * kSlab;
    if (kTma) {
      if (threadIdx.x == 0) {
        const uint32_t bar = bars + 8 * (k % kNST);
        mbar_expect_tx(bar, kSlab);
        if (MODE == 6) {
          // four row segments; rows / columns outside the image are clamped (a benchmark, not the kernel)
          for (int r = 0; r < 4; ++r) {
            int y = sl.y0 + r; y = y < 0 ? 0 : (y >= H ? H - 1 : y);
            int x = sl.x0 < 0 ? 0 : (sl.x0 + 72 > W ? W - 72 : sl.x0);
            bulk_load_1d(dst + r * 9216, map + (static_cast<size_t>(y) * W + x) * C, 9216, bar);
          }
        } else {
          tma_load_3d(dst, &tmap, sl.c0, sl.x0, sl.y0, bar);
        }
      }
    } else {
      constexpr int PB = MODE == 0 ? 32 : (MODE == 1 ? 64 : 128);   // bytes per pixel piece
      constexpr int V = PB / 16;
      constexpr int ROWS = kSlab / (72 * PB);
      for (int i = threadIdx.x; i < ROWS * 72 * V; i += kThreads) {
        const int v = i % V, px = (i / V) % 72, r = i / (V * 72);
        const int y = sl.y0 + r, x = sl.x0 + px;
        const bool ok = static_cast<unsigned>(y) < static_cast<unsigned>(H) && static_cast<unsigned>(x) < static_cast<unsigned>(W);
        const float *src = map + (static_cast<long long>(ok ? y : 0) * W + (ok ? x : 0)) * C + sl.c0 + v * 4;
        // XOR swizzle of the 16-byte piece index so that later float4 reads would be conflict-free
        const uint32_t d = dst + (r * 72 + px) * PB + ((v ^ (px & (V - 1))) << 4);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(ok ? 16 : 0) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
  };

  for (int k = 0; k < kNST - 1; ++k) {
    if (k < my) issue(k);
    else if (!kTma) asm volatile("cp.async.commit_group;" ::: "memory");
  }
  float acc = 0.f;
  for (int k = 0; k < my; ++k) {
    if (kTma) mbar_wait(bars + 8 * (k % kNST), (k / kNST) & 1);
    else asm volatile("cp.async.wait_group %0;" ::"n"(kNST - 2) : "memory");
    __syncthreads();   // slab k landed for everyone; everyone finished reading slab k-1
    if (k + kNST - 1 < my) issue(k + kNST - 1);
    else if (!kTma) asm volatile("cp.async.commit_group;" ::: "memory");
    acc += reinterpret_cast<const float *>(smem + (k % kNST) * kSlab)[threadIdx.x * 71 % (kSlab / 4)];
  }
  if (acc == 123.456f) sink[0] = acc;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static bool make_map(CUtensorMap *m, const float *ptr, int inner, int rows, CUtensorMapSwizzle sw) {
  void *p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) return false;
  const cuuint64_t dims[3] = {C, W, H};
  const cuuint64_t strides[2] = {C * 4, (cuuint64_t)W * C * 4};
  const cuuint32_t box[3] = {(cuuint32_t)inner, 72, (cuuint32_t)rows};
  const cuuint32_t estr[3] = {1, 1, 1};
  return reinterpret_cast<EncodeTiledFn>(p)(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(ptr), dims, strides, box,
                                            estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int MODE>
void run(const char *name, float *maps, float *sink, int grid) {
  constexpr int n_maps = 4;
  const size_t map_elems = (size_t)H * W * C;
  CUtensorMap tm[n_maps];
  for (int i = 0; i < n_maps; ++i) {
    if (MODE == 3 && !make_map(&tm[i], maps + i * map_elems, 32, 4, CU_TENSOR_MAP_SWIZZLE_128B)) { printf("%s: no tensor map\n", name); return; }
    if (MODE == 4 && !make_map(&tm[i], maps + i * map_elems, 16, 8, CU_TENSOR_MAP_SWIZZLE_64B)) { printf("%s: no tensor map\n", name); return; }
    if (MODE == 5 && !make_map(&tm[i], maps + i * map_elems, 8, 16, CU_TENSOR_MAP_SWIZZLE_32B)) { printf("%s: no tensor map\n", name); return; }
    if (MODE < 3 || MODE == 6) tm[i] = CUtensorMap{};
  }
  const int smem_bytes = kNST * kSlab + 64;
  cudaFuncSetAttribute(feed<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  const int n_slabs = kTilesX * kTilesY * kParts;
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 4; ++i) feed<MODE><<<grid, kThreads, smem_bytes>>>(maps + (i % n_maps) * map_elems, tm[i % n_maps], n_slabs, sink);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); exit(1); }
  const int reps = 40;
  cudaEventRecord(a);
  for (int i = 0; i < reps; ++i) feed<MODE><<<grid, kThreads, smem_bytes>>>(maps + (i % n_maps) * map_elems, tm[i % n_maps], n_slabs, sink);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  const double us = ms * 1e3 / reps, mb = (double)n_slabs * kSlab / 1e6;
  printf("%-44s grid %3d: %7.1f us per %.0f MB into shared memory = %6.0f GB/s = %5.1f B/clk/SM @1.965GHz\n", name, grid, us, mb,
         mb / us * 1e3, mb * 1e6 / (us * 1e-6) / 148 / 1.965e9);
}

int main() {
  float *maps, *sink;
  const size_t bytes = 4ull * sizeof(float) * H * W * C;
  cudaMalloc(&maps, bytes); cudaMalloc(&sink, 4);
  cudaMemset(maps, 0, bytes);
  for (int grid : {296, 148}) {
    run<0>("cp.async 32-byte pieces (8 ch x 16 rows)", maps, sink, grid);
    run<1>("cp.async 64-byte pieces (16 ch x 8 rows)", maps, sink, grid);
    run<2>("cp.async 128-byte pixels (32 ch x 4 rows)", maps, sink, grid);
    run<3>("TMA tensor 128-byte inner box, SW128", maps, sink, grid);
    run<4>("TMA tensor 64-byte inner box, SW64", maps, sink, grid);
    run<5>("TMA tensor 32-byte inner box, SW32", maps, sink, grid);
    run<6>("cp.async.bulk 1-D rows of 9216 B", maps, sink, grid);
  }
  return 0;
}
