// Glue kernels of the frame stream: ordered compaction of the anchor keep-mask and row gathers.
//
// In the reference these are NumPy boolean / fancy indexing on the host between the stages
// (avod/core/models/dt_rpn_model.py:952-958 `anchors[anchor_filter]`, :593-597 tf.gather of the
// NMS survivors). Keeping them on the device lets the whole front end of a frame run without a
// host round trip: the kept count stays in device memory and the following stages read it there.
#include "common.cuh"

namespace dodt {
namespace {

constexpr int kCompactBlock = 1024;
constexpr int kItems = 4;                               // mask bytes per thread (one uchar4)
constexpr int kTile = kCompactBlock * kItems;           // 4096 mask entries per CTA

__global__ void __launch_bounds__(kCompactBlock)
compact_count(const unsigned char *__restrict__ keep, long long n, int *__restrict__ tile_count) {
  const long long i0 = static_cast<long long>(blockIdx.x) * kTile + threadIdx.x * kItems;
  int c = 0;
#pragma unroll
  for (int k = 0; k < kItems; ++k) c += (i0 + k < n && keep[i0 + k]) ? 1 : 0;
  // block reduction
  __shared__ int warp_sum[kCompactBlock / 32];
  for (int d = 16; d > 0; d >>= 1) c += __shfl_down_sync(0xffffffffu, c, d);
  if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x < 32) {
    int v = warp_sum[threadIdx.x];
    for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
    if (threadIdx.x == 0) tile_count[blockIdx.x] = v;
  }
}

__global__ void __launch_bounds__(kCompactBlock)
compact_scatter(const unsigned char *__restrict__ keep, long long n,
                const int *__restrict__ tile_count, int n_tiles, int *__restrict__ idx,
                int *__restrict__ count) {
  __shared__ int warp_sum[kCompactBlock / 32];
  __shared__ int s_base;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // base of this tile = sum of the counts of all earlier tiles (n_tiles is a few dozen)
  int part = 0;
  for (int t = threadIdx.x; t < static_cast<int>(blockIdx.x); t += kCompactBlock) part += tile_count[t];
  for (int d = 16; d > 0; d >>= 1) part += __shfl_down_sync(0xffffffffu, part, d);
  if (lane == 0) warp_sum[warp] = part;
  __syncthreads();
  if (threadIdx.x < 32) {
    int v = warp_sum[threadIdx.x];
    for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
    if (threadIdx.x == 0) s_base = v;
  }
  __syncthreads();
  const int base = s_base;
  __syncthreads();

  const long long i0 = static_cast<long long>(blockIdx.x) * kTile + threadIdx.x * kItems;
  bool f[kItems];
  int c = 0;
#pragma unroll
  for (int k = 0; k < kItems; ++k) {
    f[k] = i0 + k < n && keep[i0 + k];
    c += f[k] ? 1 : 0;
  }
  // exclusive scan of per-thread counts: warp scan, then scan of warp totals
  int incl = c;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int up = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += up;
  }
  if (lane == 31) warp_sum[warp] = incl;
  __syncthreads();
  if (threadIdx.x < 32) {
    const int v = warp_sum[threadIdx.x];
    int w = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int up = __shfl_up_sync(0xffffffffu, w, d);
      if (lane >= d) w += up;
    }
    warp_sum[threadIdx.x] = w - v;  // exclusive
  }
  __syncthreads();
  int pos = base + warp_sum[warp] + incl - c;
#pragma unroll
  for (int k = 0; k < kItems; ++k)
    if (f[k]) idx[pos++] = static_cast<int>(i0 + k);
  if (blockIdx.x == static_cast<unsigned>(n_tiles - 1) && threadIdx.x == kCompactBlock - 1)
    *count = pos;  // last thread of the last tile ends at the total
}

__global__ void __launch_bounds__(256)
gather_rows(const float *__restrict__ src, int width, const int *__restrict__ idx,
            const int *__restrict__ count, long long n_max, float *__restrict__ dst) {
  const long long t = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  const long long row = t / width;
  const int col = static_cast<int>(t % width);
  if (row >= n_max || row >= __ldg(count)) return;
  dst[t] = __ldg(src + static_cast<long long>(__ldg(idx + row)) * width + col);
}

struct GatherMulti {
  const float *src[DODT_MAX_GATHER];
  float *dst[DODT_MAX_GATHER];
  int width[DODT_MAX_GATHER];
  int vec4[DODT_MAX_GATHER];      // width % 4 == 0 and both arrays 16-byte aligned
  int n_specs;
};

// one thread per gathered row: the row index is read once, 4-float rows (boxes) move as float4
__global__ void __launch_bounds__(256)
gather_rows_multi(const GatherMulti g, const int *__restrict__ idx, const int *__restrict__ count,
                  int n_max) {
  const int row = blockIdx.x * 256 + threadIdx.x;
  if (row >= n_max || row >= __ldg(count)) return;
  const size_t srow = static_cast<size_t>(__ldg(idx + row));
#pragma unroll
  for (int k = 0; k < DODT_MAX_GATHER; ++k) {
    if (k >= g.n_specs) break;
    const int w = g.width[k];
    if (g.vec4[k]) {
      const float4 *sp = reinterpret_cast<const float4 *>(g.src[k] + srow * w);
      float4 *dp = reinterpret_cast<float4 *>(g.dst[k] + static_cast<size_t>(row) * w);
      for (int c = 0; c < w / 4; ++c) dp[c] = __ldg(sp + c);
    } else {
      const float *sp = g.src[k] + srow * w;
      float *dp = g.dst[k] + static_cast<size_t>(row) * w;
      for (int c = 0; c < w; ++c) dp[c] = __ldg(sp + c);
    }
  }
}

// One frame's detection list appended to the shard-level block that is gathered once per shard
// (SURVEY 8(e)). The row of the block is taken from a device-side cursor, so the launch is
// replayable inside a CUDA graph; frames beyond the block's capacity are counted but dropped.
__global__ void __launch_bounds__(128)
emit_detections(const float *__restrict__ boxes, const float *__restrict__ scores,
                const int *__restrict__ keep, const int *__restrict__ n_keep, int max_det,
                const int *__restrict__ frame_id, float *__restrict__ rows, int *__restrict__ counts,
                int *__restrict__ frame_ids, int *__restrict__ cursor, int max_frames,
                int *__restrict__ row_io, int rewrite) {
  __shared__ int s_row;
  if (threadIdx.x == 0) {
    if (rewrite) {
      s_row = row_io[0];
    } else {
      s_row = atomicAdd(cursor, 1);
      if (row_io) row_io[0] = s_row;
    }
  }
  __syncthreads();
  const int row = s_row;
  if (row < 0 || row >= max_frames) return;
  const int n = min(max_det, __ldg(n_keep));
  for (int k = threadIdx.x; k < max_det; k += 128) {
    float *dst = rows + (static_cast<size_t>(row) * max_det + k) * 6;
    if (k < n) {
      const int src = __ldg(keep + k);
      const float4 b = __ldg(reinterpret_cast<const float4 *>(boxes) + src);
      dst[0] = b.x; dst[1] = b.y; dst[2] = b.z; dst[3] = b.w;
      dst[4] = __ldg(scores + src);
      dst[5] = static_cast<float>(src);
    } else {
      dst[0] = dst[1] = dst[2] = dst[3] = dst[4] = dst[5] = 0.0f;
    }
  }
  if (threadIdx.x == 0) {
    counts[row] = n;
    frame_ids[2 * row] = frame_id ? __ldg(frame_id) : -2;
    frame_ids[2 * row + 1] = frame_id ? __ldg(frame_id + 1) : row;
  }
}

}  // namespace
}  // namespace dodt

extern "C" {

int dodt_emit_detections(const float *boxes, const float *scores, const int32_t *keep,
                         const int32_t *n_keep, int32_t max_det, const int32_t *frame_id,
                         float *rows, int32_t *counts, int32_t *frame_ids, int32_t *cursor,
                         int32_t max_frames, int32_t *row_io, int32_t rewrite, dodt_stream_t stream_) {
  using namespace dodt;
  if (!boxes || !scores || !keep || !n_keep || !rows || !counts || !frame_ids || !cursor ||
      max_det <= 0 || max_frames <= 0 || (rewrite && !row_io))
    return DODT_EINVAL;
  if (reinterpret_cast<uintptr_t>(boxes) % 16 != 0) return DODT_EALIGN;
  emit_detections<<<1, 128, 0, as_stream(stream_)>>>(boxes, scores, keep, n_keep, max_det, frame_id,
                                                     rows, counts, frame_ids, cursor, max_frames, row_io,
                                                     rewrite);
  DODT_AFTER_LAUNCH();
  return DODT_OK;
}

int dodt_gather_rows_multi(const dodt_gather_spec *specs, int32_t n_specs, const int32_t *idx,
                           const int32_t *count, int64_t n_max, dodt_stream_t stream_) {
  using namespace dodt;
  if (!specs || n_specs <= 0 || n_specs > DODT_MAX_GATHER || n_max < 0 || !count) return DODT_EINVAL;
  if (n_max == 0) return DODT_OK;
  if (!idx) return DODT_EINVAL;
  if (n_max > 0x7FFFFFFF) return DODT_ECAPACITY;
  GatherMulti g;
  for (int k = 0; k < DODT_MAX_GATHER; ++k) {
    g.src[k] = nullptr; g.dst[k] = nullptr; g.width[k] = 0; g.vec4[k] = 0;
    if (k < n_specs) {
      if (!specs[k].src || !specs[k].dst || specs[k].width <= 0) return DODT_EINVAL;
      g.src[k] = specs[k].src;
      g.dst[k] = specs[k].dst;
      g.width[k] = specs[k].width;
      g.vec4[k] = specs[k].width % 4 == 0 && reinterpret_cast<uintptr_t>(specs[k].src) % 16 == 0 &&
                  reinterpret_cast<uintptr_t>(specs[k].dst) % 16 == 0;
    }
  }
  g.n_specs = n_specs;
  gather_rows_multi<<<ceil_div(n_max, 256), 256, 0, as_stream(stream_)>>>(g, idx, count,
                                                                          static_cast<int>(n_max));
  DODT_AFTER_LAUNCH();
  return DODT_OK;
}

size_t dodt_compact_workspace_bytes(int64_t n) {
  if (n < 0) return 0;
  const size_t tiles = (static_cast<size_t>(n) + dodt::kTile - 1) / dodt::kTile;
  return (tiles + 1) * sizeof(int32_t);
}

int dodt_compact_mask(const uint8_t *keep, int64_t n, int32_t *idx, int32_t *count,
                      void *workspace, size_t workspace_bytes, dodt_stream_t stream_) {
  using namespace dodt;
  if (n < 0 || n > 0x7FFFFFFF || !count || (n > 0 && (!keep || !idx))) return DODT_EINVAL;
  cudaStream_t stream = as_stream(stream_);
  if (n == 0) {
    DODT_CUDA_TRY(cudaMemsetAsync(count, 0, sizeof(int32_t), stream));
    return DODT_OK;
  }
  if (!workspace || workspace_bytes < dodt_compact_workspace_bytes(n)) return DODT_ECAPACITY;
  const int tiles = ceil_div(n, kTile);
  int *tile_count = static_cast<int *>(workspace);
  compact_count<<<tiles, kCompactBlock, 0, stream>>>(keep, n, tile_count);
  DODT_AFTER_LAUNCH();
  compact_scatter<<<tiles, kCompactBlock, 0, stream>>>(keep, n, tile_count, tiles, idx, count);
  DODT_AFTER_LAUNCH();
  return DODT_OK;
}

int dodt_gather_rows(const float *src, int32_t width, const int32_t *idx, const int32_t *count,
                     int64_t n_max, float *dst, dodt_stream_t stream_) {
  using namespace dodt;
  if (n_max < 0 || width <= 0 || !count || (n_max > 0 && (!src || !idx || !dst))) return DODT_EINVAL;
  if (n_max == 0) return DODT_OK;
  const long long total = static_cast<long long>(n_max) * width;
  const long long blocks = (total + 255) / 256;
  if (blocks > 0x7FFFFFFFll) return DODT_ECAPACITY;
  gather_rows<<<static_cast<unsigned>(blocks), 256, 0, as_stream(stream_)>>>(src, width, idx, count, n_max, dst);
  DODT_AFTER_LAUNCH();
  return DODT_OK;
}

}  // extern "C"
