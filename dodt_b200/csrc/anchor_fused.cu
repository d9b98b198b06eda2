// S2 fused for the frame runner (round 2): ONE kernel does what the reference does on the host
// between the occupancy grid and the RPN NMS, and what round 1 did in five launches
// (anchor_box_filter, compact_count, compact_scatter, gather_rows_multi, rpn_decode):
//
//   avod/core/anchor_filter.py:64-119        get_empty_anchor_filter_2d -> keep mask
//   avod/core/models/dt_rpn_model.py:952-958 anchors[anchor_filter] (ordered compaction)
//   avod/core/models/dt_rpn_model.py:975-985 BEV / image crop boxes of the kept anchors
//   avod/core/models/dt_rpn_model.py:573-591 regressed anchors of the kept anchors -> BEV boxes
//
// A CTA owns a tile of kFuseBlock consecutive anchors: box sums from the integral image (either the full
// image or the band-local image + band offsets that dodt_integral_image_2d_banded leaves, which
// saves the pass that adds the offsets), keep flags, block scan, decoupled look-back over the tile
// aggregates for the global position (ordered, so kept_idx ascends like NumPy boolean indexing),
// then the tile's kept anchors are processed with dense lanes: kept index, the two precomputed
// crop boxes, the RPN score and the decoded BEV box of the regressed anchor, all written at the
// compacted position. The last tile writes the kept count and re-arms the workspace.
#include <string.h>

#include "anchor_math.cuh"
#include "common.cuh"

namespace dodt {
namespace {

constexpr int kFuseBlock = 256;

__device__ __forceinline__ int trunc_index_f32(float v, float voxel) { return __float2int_rz(__fdiv_rn(v, voxel)); }
__device__ __forceinline__ int clip_index(int trunc, int min_coord, int ndiv) {
  // np.int32(...) - min_voxel_coord is evaluated in float64 by NumPy, so it cannot wrap
  const long long v = static_cast<long long>(trunc) - min_coord;
  return v < 0 ? 0 : (v > ndiv ? ndiv : static_cast<int>(v));
}

struct FusedArgs {
  const double *anchors;        // [n, 6], or null: the anchors are the grid `grid`, evaluated from the index
  GridGeom grid;
  long long n;
  const int *ii;                // (nx+1) x (nz+1): full integral image, or band-local (bandoff != null)
  const int *bandoff;           // [bands, nz] exclusive band offsets, or null
  int band_rows;
  int nx, nz, min_x, min_z;
  float voxel_f;
  double thr;
  const float *anchor_bev_boxes, *anchor_img_boxes;   // [n, 4] each (may be null)
  const float *rpn_scores;                            // [n] (may be null)
  const float *rpn_offsets;                           // [n, 6] (may be null: no decode)
  double x_min, x_max, z_min, z_max;
  int decode_f32;               // 1: the TF graph's float32 decode (anchor_math.cuh)
  unsigned char *keep;          // [n]
  int *kept_idx;                // [n]
  int *n_kept;                  // [1]
  float *k_bev_boxes, *k_img_boxes, *k_scores, *k_rpn_boxes;
  unsigned long long *status;   // [tiles]: flag << 32 | value; zero when idle
  unsigned int *ticket;         // tile counter, zero when idle
  unsigned int *done;           // finished-tile counter, zero when idle
};

__device__ __forceinline__ int ii_at(const FusedArgs &g, int x, int z) {
  // image row x <-> grid rows < x; band-local values belong to the band of grid row x - 1
  int v = __ldg(g.ii + static_cast<size_t>(x) * (g.nz + 1) + z);
  if (g.bandoff && x > 0 && z > 0)
    v += __ldg(g.bandoff + static_cast<size_t>((x - 1) / g.band_rows) * g.nz + (z - 1));
  return v;
}

__global__ void __launch_bounds__(kFuseBlock)
anchor_filter_fused(const FusedArgs g) {
  __shared__ int s_warp[kFuseBlock / 32];
  __shared__ int s_list[kFuseBlock];
  __shared__ int s_tile, s_base, s_total;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_tile = static_cast<int>(atomicAdd(g.ticket, 1u));
  __syncthreads();
  const int tile = s_tile;
  const int n_tiles = static_cast<int>((g.n + kFuseBlock - 1) / kFuseBlock);
  const long long i = static_cast<long long>(tile) * kFuseBlock + threadIdx.x;

  // ---- keep flag (anchor_filter.py:93-119): corners in float64, stored to float32, map_to_index
  bool keep = false;
  if (i < g.n) {
    double a[6];
    if (g.anchors) {
      a[0] = __ldg(g.anchors + i * 6); a[2] = __ldg(g.anchors + i * 6 + 2);
      a[3] = __ldg(g.anchors + i * 6 + 3); a[5] = __ldg(g.anchors + i * 6 + 5);
    } else {
      grid_anchor(g.grid, i, a);
    }
    const double x = a[0], z = a[2], hx = __ddiv_rn(a[3], 2.0), hz = __ddiv_rn(a[5], 2.0);
    const float tlx = __double2float_rn(__dsub_rn(x, hx)), tlz = __double2float_rn(__dsub_rn(z, hz));
    const float brx = __double2float_rn(__dadd_rn(x, hx)), brz = __double2float_rn(__dadd_rn(z, hz));
    const int x1 = clip_index(trunc_index_f32(tlx, g.voxel_f), g.min_x, g.nx);
    const int z1 = clip_index(trunc_index_f32(tlz, g.voxel_f), g.min_z, g.nz);
    const int x2 = clip_index(trunc_index_f32(brx, g.voxel_f), g.min_x, g.nx);
    const int z2 = clip_index(trunc_index_f32(brz, g.voxel_f), g.min_z, g.nz);
    const int s = ii_at(g, x2, z2) + ii_at(g, x1, z1) - ii_at(g, x2, z1) - ii_at(g, x1, z2);
    keep = static_cast<double>(s) >= g.thr;
    g.keep[i] = keep ? 1 : 0;
  }

  // ---- block scan of the flags
  const unsigned ballot = __ballot_sync(0xffffffffu, keep);
  const int in_warp = __popc(ballot & ((1u << lane) - 1u));
  if (lane == 0) s_warp[warp] = __popc(ballot);
  __syncthreads();
  if (warp == 0) {
    const int v = lane < kFuseBlock / 32 ? s_warp[lane] : 0;
    int w = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int up = __shfl_up_sync(0xffffffffu, w, d);
      if (lane >= d) w += up;
    }
    if (lane < kFuseBlock / 32) s_warp[lane] = w - v;   // exclusive
    if (lane == 31) s_total = w;
  }
  __syncthreads();
  const int local = s_warp[warp] + in_warp;
  const int total = s_total;
  if (keep) s_list[local] = threadIdx.x;

  // ---- decoupled look-back: exclusive prefix of the tile totals. Warp 0 inspects 32 predecessors
  // per step (tiles run in ticket order, so every predecessor has started): the chain of dependent
  // L2 reads is tile/32 long instead of tile.
  if (warp == 0) {
    if (lane == 0 && tile > 0) atomicExch(g.status + tile, (1ull << 32) | static_cast<unsigned>(total));
    int base = 0;
    for (int j = tile - 1; j >= 0; j -= 32) {
      const int idx = j - lane;
      unsigned long long st = 2ull << 32;            // before the first tile: inclusive prefix 0
      if (idx >= 0) {
        do {
          st = *reinterpret_cast<volatile unsigned long long *>(g.status + idx);
        } while ((st >> 32) == 0);
      }
      const unsigned inclusive = __ballot_sync(0xffffffffu, (st >> 32) == 2);
      const int stop = inclusive ? __ffs(inclusive) - 1 : 31;   // nearest predecessor with a full prefix
      int v = lane <= stop ? static_cast<int>(st & 0xffffffffu) : 0;
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
      base += v;
      if (inclusive) break;
    }
    if (lane == 0) {
      __threadfence();
      atomicExch(g.status + tile, (2ull << 32) | static_cast<unsigned>(base + total));
      s_base = base;
      if (tile == n_tiles - 1) *g.n_kept = base + total;
    }
  }
  __syncthreads();
  const int base = s_base;

  // ---- the tile's kept anchors, dense lanes: thread k < total handles the k-th kept anchor
  if (static_cast<int>(threadIdx.x) < total) {
    const long long src = static_cast<long long>(tile) * kFuseBlock + s_list[threadIdx.x];
    const size_t pos = static_cast<size_t>(base) + threadIdx.x;
    g.kept_idx[pos] = static_cast<int>(src);
    if (g.k_bev_boxes)
      reinterpret_cast<float4 *>(g.k_bev_boxes)[pos] = __ldg(reinterpret_cast<const float4 *>(g.anchor_bev_boxes) + src);
    if (g.k_img_boxes)
      reinterpret_cast<float4 *>(g.k_img_boxes)[pos] = __ldg(reinterpret_cast<const float4 *>(g.anchor_img_boxes) + src);
    if (g.k_scores) g.k_scores[pos] = __ldg(g.rpn_scores + src);
    if (g.k_rpn_boxes) {
      double a[6];
      if (g.anchors) load_anchor(g.anchors + src * 6, a); else grid_anchor(g.grid, src, a);
      if (g.decode_f32) {
        float r[6];
        decode_anchor_f32(a, g.rpn_offsets + src * 6, r);
        reinterpret_cast<float4 *>(g.k_rpn_boxes)[pos] =
            bev_box_of_f32(r, bev_extents_f32(g.x_min, g.x_max, g.z_min, g.z_max));
      } else {
        double r[6];
        decode_anchor(a, g.rpn_offsets + src * 6, r);
        reinterpret_cast<float4 *>(g.k_rpn_boxes)[pos] = bev_box_of(r, g.x_min, g.x_max, g.z_min, g.z_max);
      }
    }
  }

  // ---- re-arm the workspace: the last tile to finish zeroes the status words and counters
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    s_tile = atomicAdd(g.done, 1u) == static_cast<unsigned>(n_tiles - 1) ? 1 : 0;
  }
  __syncthreads();
  if (s_tile) {
    for (int t = threadIdx.x; t < n_tiles; t += kFuseBlock) g.status[t] = 0ull;
    if (threadIdx.x == 0) { *g.ticket = 0u; *g.done = 0u; }
  }
}

}  // namespace
}  // namespace dodt

extern "C" {

size_t dodt_anchor_filter_fused_workspace_bytes(int64_t n) {
  if (n < 0) return 0;
  const size_t tiles = (static_cast<size_t>(n) + dodt::kFuseBlock - 1) / dodt::kFuseBlock;
  return (tiles + 2) * sizeof(unsigned long long);
}

int dodt_anchor_filter_fused(const double *anchors, const dodt_anchor_grid *grid, int64_t n, const int32_t *ii,
                             const int32_t *bandoff,
                             int32_t band_rows, int32_t nx, int32_t nz, int32_t min_x, int32_t min_z,
                             double voxel_size, double density_threshold, const float *anchor_bev_boxes,
                             const float *anchor_img_boxes, const float *rpn_scores, const float *rpn_offsets,
                             const double bev_extents[4], int32_t decode_f32, uint8_t *keep, int32_t *kept_idx,
                             int32_t *n_kept, float *k_bev_boxes, float *k_img_boxes, float *k_scores,
                             float *k_rpn_boxes, void *workspace, size_t workspace_bytes, dodt_stream_t stream_) {
  using namespace dodt;
  if (n < 0 || n > 0x7FFFFFFF || !ii || nx <= 0 || nz <= 0 || !(voxel_size > 0.0) || !n_kept) return DODT_EINVAL;
  if (bandoff && band_rows <= 0) return DODT_EINVAL;
  cudaStream_t stream = as_stream(stream_);
  if (n == 0) {
    DODT_CUDA_TRY(cudaMemsetAsync(n_kept, 0, sizeof(int32_t), stream));
    return DODT_OK;
  }
  if ((!anchors && !grid) || !keep || !kept_idx) return DODT_EINVAL;
  FusedArgs g;
  memset(&g.grid, 0, sizeof(g.grid));
  if (!anchors) {   // anchors as a function of the index
    const int rc = fill_grid_geom(grid->extents, grid->sizes, grid->n_sizes, grid->stride, grid->plane, &g.grid);
    if (rc != DODT_OK) return rc;
    if (static_cast<int64_t>(g.grid.nx) * g.grid.nz * g.grid.n_sizes * 2 != n) return DODT_ESHAPE;
  }
  if ((k_bev_boxes && !anchor_bev_boxes) || (k_img_boxes && !anchor_img_boxes) || (k_scores && !rpn_scores) ||
      (k_rpn_boxes && (!rpn_offsets || !bev_extents)))
    return DODT_EINVAL;
  const uintptr_t al = reinterpret_cast<uintptr_t>(anchor_bev_boxes) | reinterpret_cast<uintptr_t>(anchor_img_boxes) |
                       reinterpret_cast<uintptr_t>(k_bev_boxes) | reinterpret_cast<uintptr_t>(k_img_boxes) |
                       reinterpret_cast<uintptr_t>(k_rpn_boxes);
  if (al % 16 != 0 || reinterpret_cast<uintptr_t>(workspace) % 8 != 0) return DODT_EALIGN;
  if (!workspace || workspace_bytes < dodt_anchor_filter_fused_workspace_bytes(n)) return DODT_ECAPACITY;
  const int tiles = ceil_div(n, kFuseBlock);
  g.anchors = anchors; g.n = n; g.ii = ii; g.bandoff = bandoff; g.band_rows = band_rows;
  g.nx = nx; g.nz = nz; g.min_x = min_x; g.min_z = min_z;
  g.voxel_f = static_cast<float>(voxel_size);
  g.thr = density_threshold;
  g.anchor_bev_boxes = anchor_bev_boxes; g.anchor_img_boxes = anchor_img_boxes;
  g.rpn_scores = rpn_scores; g.rpn_offsets = rpn_offsets;
  g.x_min = bev_extents ? bev_extents[0] : 0.0; g.x_max = bev_extents ? bev_extents[1] : 1.0;
  g.z_min = bev_extents ? bev_extents[2] : 0.0; g.z_max = bev_extents ? bev_extents[3] : 1.0;
  g.decode_f32 = decode_f32 ? 1 : 0;
  g.keep = keep; g.kept_idx = kept_idx; g.n_kept = n_kept;
  g.k_bev_boxes = k_bev_boxes; g.k_img_boxes = k_img_boxes; g.k_scores = k_scores; g.k_rpn_boxes = k_rpn_boxes;
  g.status = static_cast<unsigned long long *>(workspace);
  g.ticket = reinterpret_cast<unsigned int *>(g.status + tiles);
  g.done = g.ticket + 1;
  anchor_filter_fused<<<tiles, kFuseBlock, 0, stream>>>(g);
  DODT_AFTER_LAUNCH();
  return DODT_OK;
}

}  // extern "C"
