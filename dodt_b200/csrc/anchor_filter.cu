// S2 — integral image of the occupancy grid + O(1) box-sum empty-anchor filter, sm_100a.
//
// Reference behaviour reproduced (paths relative to the Guoxs/DODT checkout):
//   wavedata/.../core/integral_image_2d.py:7-37   cumsum over both axes, zero-padded to (nx+1,nz+1)
//   wavedata/.../core/integral_image_2d.py:39-87  query: clamp to size-1, four-corner sum
//   wavedata/.../core/voxel_grid_2d.py:162-186    map_to_index: divide in the coordinate dtype,
//                                                  truncate toward zero, shift, clip to [0, ndiv]
//   avod/core/anchor_filter.py:64-119             get_empty_anchor_filter_2d
//
// The reference keeps the integral image in float64; every value is an integer <= nx*nz, so the
// int32 image written here is bit-for-bit the same numbers.
//
// Kernels
//   ii_band_scan     one CTA per band of kBand grid rows (x); one warp per row scans along z with
//                    shuffles (uchar4 loads, 128 cells per warp step), row prefixes are staged in
//                    shared memory as uint16, then the CTA scans the band along x and writes the
//                    band-local integral image plus the band's column totals.
//   ii_band_prefix   exclusive prefix of the band totals along the band axis (one thread per column).
//   ii_band_offsets  one thread per image element adds its band's offset (coalesced along z).
//   anchor_box_filter one thread per anchor; anchor rows are staged through shared memory so the
//                    48-byte rows are read with fully coalesced loads; four L2-resident gathers.
#include <string.h>

#include "common.cuh"

namespace dodt {
namespace {

constexpr int kBand = 16;              // grid rows (x) per CTA
constexpr int kBandThreads = kBand * 32;

__global__ void __launch_bounds__(kBandThreads)
ii_band_scan(const unsigned char *__restrict__ occ, int nx, int nz, int vec_ok,
             int *__restrict__ ii, int *__restrict__ bandsum, int *__restrict__ bandoff,
             unsigned int *__restrict__ counter) {
  extern __shared__ unsigned short rowp[];  // [kBand][nz] inclusive row prefixes along z
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int band = blockIdx.x;
  const int x = band * kBand + warp;
  const int ld = nz + 1;

  if (x < nx) {
    const unsigned char *row = occ + static_cast<size_t>(x) * nz;
    unsigned carry = 0;
    for (int z0 = lane * 4; z0 - lane * 4 < nz; z0 += 128) {
      unsigned v0 = 0, v1 = 0, v2 = 0, v3 = 0;
      if (vec_ok && z0 + 3 < nz) {
        const uchar4 t = __ldg(reinterpret_cast<const uchar4 *>(row + z0));
        v0 = t.x; v1 = t.y; v2 = t.z; v3 = t.w;
      } else {
        if (z0 < nz) v0 = __ldg(row + z0);
        if (z0 + 1 < nz) v1 = __ldg(row + z0 + 1);
        if (z0 + 2 < nz) v2 = __ldg(row + z0 + 2);
        if (z0 + 3 < nz) v3 = __ldg(row + z0 + 3);
      }
      // any non-zero byte counts as 1: leaf_layout + 1 is 0 or 1 (anchor_filter.py:83)
      v0 = v0 != 0; v1 = v1 != 0; v2 = v2 != 0; v3 = v3 != 0;
      const unsigned p1 = v0, p2 = p1 + v1, p3 = p2 + v2, p4 = p3 + v3;
      unsigned incl = p4;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const unsigned up = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += up;
      }
      const unsigned excl = carry + incl - p4;
      unsigned short *dst = rowp + static_cast<size_t>(warp) * nz;
      if (z0 < nz) dst[z0] = static_cast<unsigned short>(excl + p1);
      if (z0 + 1 < nz) dst[z0 + 1] = static_cast<unsigned short>(excl + p2);
      if (z0 + 2 < nz) dst[z0 + 2] = static_cast<unsigned short>(excl + p3);
      if (z0 + 3 < nz) dst[z0 + 3] = static_cast<unsigned short>(excl + p4);
      carry += __shfl_sync(0xffffffffu, incl, 31);
    }
  }
  __syncthreads();

  const int rows = min(kBand, nx - band * kBand);
  // zero column z = 0 of this band's rows, and row x = 0 of the image (integral_image_2d.py:34)
  if (threadIdx.x < rows) ii[static_cast<size_t>(band * kBand + threadIdx.x + 1) * ld] = 0;
  if (band == 0)
    for (int z = threadIdx.x; z < ld; z += kBandThreads) ii[z] = 0;
  for (int z = threadIdx.x; z < nz; z += kBandThreads) {
    int acc = 0;
    for (int r = 0; r < rows; ++r) {
      acc += rowp[static_cast<size_t>(r) * nz + z];
      ii[static_cast<size_t>(band * kBand + r + 1) * ld + z + 1] = acc;
    }
    bandsum[static_cast<size_t>(band) * nz + z] = acc;
  }
  if (!counter) return;
  // Banded form (dodt_integral_image_2d_banded): the LAST band to finish turns the band totals
  // into exclusive band offsets, so that no second launch is needed; consumers add
  // bandoff[band of the row][z] to the band-local image themselves (anchor_fused.cu).
  __shared__ int s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(counter, 1u) == gridDim.x - 1 ? 1 : 0;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const int bands = gridDim.x;
  for (int z = threadIdx.x; z < nz; z += kBandThreads) {
    int run = 0;
    for (int b0 = 0; b0 < bands; b0 += 16) {
      int v[16];
#pragma unroll
      for (int k = 0; k < 16; ++k)
        v[k] = b0 + k < bands ? __ldcg(bandsum + static_cast<size_t>(b0 + k) * nz + z) : 0;
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        if (b0 + k < bands) bandoff[static_cast<size_t>(b0 + k) * nz + z] = run;
        run += v[k];
      }
    }
  }
  if (threadIdx.x == 0) *counter = 0u;   // re-armed for the next call on this workspace
}

// exclusive prefix of the band totals along the band axis: bandoff[b][z] = sum of bandsum[0..b-1][z].
// One thread per column; separate in/out arrays so that the loads of a chunk are all in flight
// before the dependent adds start.
__global__ void __launch_bounds__(64)
ii_band_prefix(int bands, int nz, const int *__restrict__ bandsum, int *__restrict__ bandoff) {
  const int z = blockIdx.x * 64 + threadIdx.x;
  if (z >= nz) return;
  int run = 0;
  for (int b0 = 0; b0 < bands; b0 += 16) {
    int v[16];
#pragma unroll
    for (int k = 0; k < 16; ++k)
      v[k] = b0 + k < bands ? __ldg(bandsum + static_cast<size_t>(b0 + k) * nz + z) : 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      if (b0 + k < bands) bandoff[static_cast<size_t>(b0 + k) * nz + z] = run;
      run += v[k];
    }
  }
}

// add its band's offset to every integral-image element: a thread owns one column z of
// kOffRows consecutive rows (kOffRows divides kBand, so the rows share one band offset)
constexpr int kOffRows = 8;
static_assert(kBand % kOffRows == 0, "rows of a CTA lie in one band");
__global__ void __launch_bounds__(256)
ii_band_offsets(int nx, int nz, int *__restrict__ ii, const int *__restrict__ bandoff) {
  const int z = blockIdx.x * 256 + threadIdx.x;
  const int x0 = blockIdx.y * kOffRows;       // grid rows x0.. <-> image rows x0+1..
  if (z >= nz || x0 < kBand) return;          // band 0 needs no offset
  const int off = __ldg(bandoff + static_cast<size_t>(x0 / kBand) * nz + z);
  if (off == 0) return;
  const int x1 = min(nx, x0 + kOffRows);
  for (int x = x0; x < x1; ++x) ii[static_cast<size_t>(x + 1) * (nz + 1) + z + 1] += off;
}

template <typename T>
__device__ __forceinline__ int trunc_div_index(T v, T voxel);
template <>
__device__ __forceinline__ int trunc_div_index<float>(float v, float voxel) {
  return __float2int_rz(__fdiv_rn(v, voxel));
}
template <>
__device__ __forceinline__ int trunc_div_index<double>(double v, double voxel) {
  return __double2int_rz(__ddiv_rn(v, voxel));
}

__device__ __forceinline__ int clip_index(int trunc, int min_coord, int ndiv) {
  // np.int32(...) - min_voxel_coord is evaluated in float64 by NumPy, so it cannot wrap
  const long long v = static_cast<long long>(trunc) - min_coord;
  return v < 0 ? 0 : (v > ndiv ? ndiv : static_cast<int>(v));
}

template <typename T>
__global__ void __launch_bounds__(256)
map_to_index_kernel(const T *__restrict__ coords, long long n, T voxel, int min_x, int min_z,
                    int nx, int nz, int *__restrict__ idx) {
  const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;  // over 2n scalars
  if (i >= 2 * n) return;
  const bool is_z = i & 1;
  idx[i] = clip_index(trunc_div_index<T>(__ldg(coords + i), voxel), is_z ? min_z : min_x,
                      is_z ? nz : nx);
}

constexpr int kFilterBlock = 256;

template <typename T>
__global__ void __launch_bounds__(kFilterBlock)
anchor_box_filter(const T *__restrict__ anchors, long long n, const int *__restrict__ ii, int nx,
                  int nz, int min_x, int min_z, float voxel_f, double thr,
                  unsigned char *__restrict__ keep, int *__restrict__ scores) {
  __shared__ T rows[kFilterBlock * 6];
  const long long first = static_cast<long long>(blockIdx.x) * kFilterBlock;
  const long long remaining = n - first;
  const int count = remaining < kFilterBlock ? static_cast<int>(remaining) : kFilterBlock;
  const T *src = anchors + first * 6;
  for (int k = threadIdx.x; k < count * 6; k += kFilterBlock) rows[k] = __ldg(src + k);
  __syncthreads();
  if (static_cast<int>(threadIdx.x) >= count) return;
  const T *a = rows + threadIdx.x * 6;
  // anchor_filter.py:93-102 — corners in the anchors' dtype, stored to float32
  const T half = T(2);
  const T x = a[0], z = a[2], hx = a[3] / half, hz = a[5] / half;
  const float tlx = static_cast<float>(x - hx), tlz = static_cast<float>(z - hz);
  const float brx = static_cast<float>(x + hx), brz = static_cast<float>(z + hz);
  // map_to_index on float32 corners (voxel_grid_2d.py:182-184), then uint32 clamp of query()
  const int x1 = clip_index(trunc_div_index<float>(tlx, voxel_f), min_x, nx);
  const int z1 = clip_index(trunc_div_index<float>(tlz, voxel_f), min_z, nz);
  const int x2 = clip_index(trunc_div_index<float>(brx, voxel_f), min_x, nx);
  const int z2 = clip_index(trunc_div_index<float>(brz, voxel_f), min_z, nz);
  const int ld = nz + 1;
  const int s = __ldg(ii + static_cast<size_t>(x2) * ld + z2) + __ldg(ii + static_cast<size_t>(x1) * ld + z1) -
                __ldg(ii + static_cast<size_t>(x2) * ld + z1) - __ldg(ii + static_cast<size_t>(x1) * ld + z2);
  keep[first + threadIdx.x] = static_cast<double>(s) >= thr ? 1 : 0;
  if (scores) scores[first + threadIdx.x] = s;
}

}  // namespace
}  // namespace dodt

extern "C" {

size_t dodt_integral_workspace_bytes(int32_t nx, int32_t nz) {
  if (nx <= 0 || nz <= 0) return 0;
  const size_t bands = (static_cast<size_t>(nx) + dodt::kBand - 1) / dodt::kBand;
  return 2 * bands * nz * sizeof(int32_t);  // band totals + their exclusive prefix
}

int dodt_integral_image_2d(const uint8_t *occ, int32_t nx, int32_t nz, int32_t *ii,
                           void *workspace, size_t workspace_bytes, dodt_stream_t stream_) {
  using namespace dodt;
  if (!occ || !ii || nx <= 0 || nz <= 0) return DODT_EINVAL;
  if (nz > 65535) return DODT_ECAPACITY;  // row prefixes are staged as uint16
  const size_t need = dodt_integral_workspace_bytes(nx, nz);
  if (!workspace || workspace_bytes < need) return DODT_ECAPACITY;
  cudaStream_t stream = as_stream(stream_);
  const int bands = ceil_div(nx, kBand);
  const size_t smem = static_cast<size_t>(kBand) * nz * sizeof(unsigned short);
  if (smem > 200 * 1024) return DODT_ECAPACITY;
  if (smem > 48 * 1024)
    DODT_CUDA_TRY(cudaFuncSetAttribute(ii_band_scan, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem)));
  const int vec_ok = (reinterpret_cast<uintptr_t>(occ) % 4 == 0 && nz % 4 == 0) ? 1 : 0;
  int *bandsum = static_cast<int *>(workspace);
  ii_band_scan<<<bands, kBandThreads, smem, stream>>>(occ, nx, nz, vec_ok, ii, bandsum, nullptr, nullptr);
  DODT_AFTER_LAUNCH();
  if (bands > 1) {
    int *bandoff = bandsum + static_cast<size_t>(bands) * nz;
    ii_band_prefix<<<ceil_div(nz, 64), 64, 0, stream>>>(bands, nz, bandsum, bandoff);
    DODT_AFTER_LAUNCH();
    dim3 grid(ceil_div(nz, 256), ceil_div(nx, kOffRows));
    ii_band_offsets<<<grid, 256, 0, stream>>>(nx, nz, ii, bandoff);
    DODT_AFTER_LAUNCH();
  }
  return DODT_OK;
}

size_t dodt_integral_banded_workspace_bytes(int32_t nx, int32_t nz) {
  if (nx <= 0 || nz <= 0) return 0;
  return dodt_integral_workspace_bytes(nx, nz) + 256;   // + the finished-band counter
}

int32_t dodt_integral_band_rows(void) { return dodt::kBand; }

int dodt_integral_image_2d_banded(const uint8_t *occ, int32_t nx, int32_t nz, int32_t *ii_local,
                                  void *workspace, size_t workspace_bytes, int32_t **bandoff_out,
                                  dodt_stream_t stream_) {
  using namespace dodt;
  if (!occ || !ii_local || nx <= 0 || nz <= 0) return DODT_EINVAL;
  if (nz > 65535) return DODT_ECAPACITY;
  const size_t need = dodt_integral_banded_workspace_bytes(nx, nz);
  if (!workspace || workspace_bytes < need) return DODT_ECAPACITY;
  if (reinterpret_cast<uintptr_t>(workspace) % 4 != 0) return DODT_EALIGN;
  cudaStream_t stream = as_stream(stream_);
  const int bands = ceil_div(nx, kBand);
  const size_t smem = static_cast<size_t>(kBand) * nz * sizeof(unsigned short);
  if (smem > 200 * 1024) return DODT_ECAPACITY;
  if (smem > 48 * 1024)
    DODT_CUDA_TRY(cudaFuncSetAttribute(ii_band_scan, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem)));
  const int vec_ok = (reinterpret_cast<uintptr_t>(occ) % 4 == 0 && nz % 4 == 0) ? 1 : 0;
  int *bandsum = static_cast<int *>(workspace);
  int *bandoff = bandsum + static_cast<size_t>(bands) * nz;
  unsigned int *counter = reinterpret_cast<unsigned int *>(bandoff + static_cast<size_t>(bands) * nz);
  ii_band_scan<<<bands, kBandThreads, smem, stream>>>(occ, nx, nz, vec_ok, ii_local, bandsum, bandoff, counter);
  DODT_AFTER_LAUNCH();
  if (bandoff_out) *bandoff_out = bandoff;
  return DODT_OK;
}

int dodt_map_to_index(const void *coords, int32_t dtype, int64_t n, double voxel_size,
                      int32_t min_x, int32_t min_z, int32_t nx, int32_t nz, int32_t *idx,
                      dodt_stream_t stream_) {
  using namespace dodt;
  if (n < 0 || (n > 0 && (!coords || !idx)) || !(voxel_size > 0.0)) return DODT_EINVAL;
  if (n == 0) return DODT_OK;
  cudaStream_t stream = as_stream(stream_);
  const int blocks = ceil_div(2 * n, 256);
  if (dtype == DODT_F32)
    map_to_index_kernel<float><<<blocks, 256, 0, stream>>>(static_cast<const float *>(coords), n,
                                                           static_cast<float>(voxel_size), min_x,
                                                           min_z, nx, nz, idx);
  else if (dtype == DODT_F64)
    map_to_index_kernel<double><<<blocks, 256, 0, stream>>>(static_cast<const double *>(coords), n,
                                                            voxel_size, min_x, min_z, nx, nz, idx);
  else
    return DODT_EINVAL;
  DODT_AFTER_LAUNCH();
  return DODT_OK;
}

int dodt_anchor_filter_2d(const void *anchors, int32_t dtype, int64_t n, const int32_t *ii,
                          int32_t nx, int32_t nz, int32_t min_x, int32_t min_z, double voxel_size,
                          double density_threshold, uint8_t *keep, int32_t *scores,
                          dodt_stream_t stream_) {
  using namespace dodt;
  if (n < 0 || (n > 0 && (!anchors || !keep)) || !ii || nx <= 0 || nz <= 0 || !(voxel_size > 0.0))
    return DODT_EINVAL;
  if (n == 0) return DODT_OK;
  cudaStream_t stream = as_stream(stream_);
  const int blocks = ceil_div(n, kFilterBlock);
  const float voxel_f = static_cast<float>(voxel_size);
  if (dtype == DODT_F64)
    anchor_box_filter<double><<<blocks, kFilterBlock, 0, stream>>>(
        static_cast<const double *>(anchors), n, ii, nx, nz, min_x, min_z, voxel_f,
        density_threshold, keep, scores);
  else if (dtype == DODT_F32)
    anchor_box_filter<float><<<blocks, kFilterBlock, 0, stream>>>(
        static_cast<const float *>(anchors), n, ii, nx, nz, min_x, min_z, voxel_f,
        density_threshold, keep, scores);
  else
    return DODT_EINVAL;
  DODT_AFTER_LAUNCH();
  return DODT_OK;
}

}  // extern "C"
