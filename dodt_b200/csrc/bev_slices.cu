// S1 — LiDAR points -> BEV height-slice maps + density map (+ the S2 occupancy grid), sm_100a.
//
// Reference behaviour reproduced (paths relative to the Guoxs/DODT checkout):
//   avod/core/bev_generators/bev_slices.py:33-150      BevSlices.generate_bev
//   avod/datasets/kitti/kitti_utils.py:81-109          create_slice_filter (xor of two filters)
//   wavedata/.../obj_detection/obj_utils.py:453-500    get_point_filter (open extents, plane test)
//   wavedata/.../core/voxel_grid_2d.py:43-160          voxelize_2d (lexsort + unique = "first point
//                                                       of the lowest y-bin in file order wins")
//   wavedata/.../core/geometry_utils.py:25-40          dist_to_plane
//   avod/core/bev_generators/bev_generator.py:23-41    _create_density_map
//
// Design. The reference sorts the cloud six times per frame. Here one pass over the points does
// everything with atomics straight into the OUTPUT maps, which double as the scratch space:
//
//   pass 0  cudaMemsetAsync(maps, 0)          — 0x00000000 is both "empty cell" and 0.0f, so the
//                                               compulsory write of the (mostly empty) maps is
//                                               also the initialisation of the atomics.
//   pass 1  bev_accumulate                    — per point: open-extent test, three voxel bins,
//            the S+2 slice predicates in fp64 (bit-exact with NumPy), then for every map the point
//            belongs to a warp-aggregated atomicMax of an inverted (y-bin, point index) key (height
//            maps) or a warp-aggregated atomicAdd (density counts) on the map cell itself, and a
//            byte store into the occupancy grid. Every atomic is fire-and-forget (RED): nothing
//            waits for an L2 round trip, a warp retires as soon as its reductions are issued.
//   pass 2  bev_resolve_scan                  — streams over the maps (uint4 loads); for every
//            non-empty cell it gathers the winning point, evaluates dist_to_plane / normalisation
//            in fp64 in NumPy's operation order and overwrites the key with the final float;
//            density counts go through a host-computed LUT of min(1, ln(n+1)/ln16) so they are
//            bit-exact with NumPy's log. The 0/1-point slice fallback of bev_slices.py:76-99 is
//            applied here. (An earlier version collected first-touch cells in a list so that this
//            pass visited only those; the returning atomics it needed serialised two L2 round
//            trips per point and cost more than re-reading the 13 MB of maps.)
//
// HBM traffic per frame: 12 B (fp32) per point, one write and one read of the maps; keys and
// counts never exist outside the maps.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace dodt {
namespace {

constexpr int kBlock = 256;

struct BevDev {
  double a, b, c, d, norm;
  double ext[6];
  double voxel;
  double hpd;
  double dhi[DODT_MAX_SLICES + 2];  // d - hi of height slices, then density, then occupancy
  double dlo[DODT_MAX_SLICES + 2];  // d - lo
  double lo[DODT_MAX_SLICES];       // lower bound of each height slice (normalisation)
  float origin_val[DODT_MAX_SLICES];
  float lut[DODT_MAX_DENSITY_LUT];
  int lut_len;
  int S;  // number of height slices
  int filter_mode;
  int height_from_y;
  int nx, nz, min_x, min_y_biased, min_z;
  int idx_bits;
  unsigned idx_mask;
  unsigned yb_max;
  int origin_cell;  // linear cell (row-major [nz][nx], rotated) of point (0,0,0), or -1
  int has_occ;
  unsigned touched_cap;
};

template <typename T, int VEC>
struct VecLoad;
template <>
struct VecLoad<float, 4> {
  static __device__ __forceinline__ void load(const float *p, float (&v)[4]) {
    float4 t = __ldg(reinterpret_cast<const float4 *>(p));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
};
template <>
struct VecLoad<float, 2> {
  static __device__ __forceinline__ void load(const float *p, float (&v)[2]) {
    float2 t = __ldg(reinterpret_cast<const float2 *>(p));
    v[0] = t.x; v[1] = t.y;
  }
};
template <>
struct VecLoad<float, 1> {
  static __device__ __forceinline__ void load(const float *p, float (&v)[1]) { v[0] = __ldg(p); }
};
template <>
struct VecLoad<double, 2> {
  static __device__ __forceinline__ void load(const double *p, double (&v)[2]) {
    double2 t = __ldg(reinterpret_cast<const double2 *>(p));
    v[0] = t.x; v[1] = t.y;
  }
};
template <>
struct VecLoad<double, 1> {
  static __device__ __forceinline__ void load(const double *p, double (&v)[1]) { v[0] = __ldg(p); }
};
template <>
struct VecLoad<double, 4> {
  static __device__ __forceinline__ void load(const double *p, double (&v)[4]) {
    double2 t0 = __ldg(reinterpret_cast<const double2 *>(p));
    double2 t1 = __ldg(reinterpret_cast<const double2 *>(p) + 1);
    v[0] = t0.x; v[1] = t0.y; v[2] = t1.x; v[3] = t1.y;
  }
};

// floor(v / voxel) exactly as np.floor(pts / voxel_size).astype(np.int32) on float64 data
// (voxel_grid_2d.py:67): IEEE division, floor, then a saturating conversion (NumPy's cast of an
// out-of-range value is undefined; extents keep real data far inside int32).
__device__ __forceinline__ int voxel_bin(double v, double voxel) {
  return __double2int_rd(__ddiv_rn(v, voxel));
}

struct BlockStage {
  int slice_pts[DODT_MAX_SLICES + 2];   // points per slice seen by this block
};

template <typename T, int VEC>
__global__ void __launch_bounds__(kBlock)
bev_accumulate(const T *__restrict__ px, const T *__restrict__ py, const T *__restrict__ pz,
               long long n_max, const int *__restrict__ n_dev, int vec_ok,
               const __grid_constant__ BevDev P,
               unsigned *__restrict__ maps, unsigned char *__restrict__ occ,
               int *__restrict__ stats) {
  __shared__ BlockStage st;

  const int lane = threadIdx.x & 31;
  if (threadIdx.x < DODT_MAX_SLICES + 2) st.slice_pts[threadIdx.x] = 0;
  __syncthreads();

  const long long n = n_dev ? min(n_max, static_cast<long long>(__ldg(n_dev))) : n_max;
  const long long i0 = (static_cast<long long>(blockIdx.x) * kBlock + threadIdx.x) * VEC;
  T x[VEC], y[VEC], z[VEC];
  if (vec_ok && i0 + VEC <= n) {
    VecLoad<T, VEC>::load(px + i0, x);
    VecLoad<T, VEC>::load(py + i0, y);
    VecLoad<T, VEC>::load(pz + i0, z);
  } else {
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const bool ok = i0 + j < n;
      x[j] = ok ? __ldg(px + i0 + j) : T(0);
      y[j] = ok ? __ldg(py + i0 + j) : T(0);
      z[j] = ok ? __ldg(pz + i0 + j) : T(0);
    }
  }

  const int HW = P.nx * P.nz;
  const int n_maps = P.S + 1;          // height slices + density
  const int n_pred = P.S + 2;          // + occupancy slice

#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    const long long idx = i0 + j;
    const double xd = static_cast<double>(x[j]);
    const double yd = static_cast<double>(y[j]);
    const double zd = static_cast<double>(z[j]);
    bool live = idx < n;
    if (P.filter_mode) {
      // open-interval extents test, obj_utils.py:474-479
      live = live && xd > P.ext[0] && xd < P.ext[1] && yd > P.ext[2] && yd < P.ext[3] &&
             zd > P.ext[4] && zd < P.ext[5];
    }
    int cell = 0;
    unsigned inv_key = 0;
    if (live) {
      const int bx = voxel_bin(xd, P.voxel) - P.min_x;
      const int bz = voxel_bin(zd, P.voxel) - P.min_z;
      if (bx < 0 || bx >= P.nx || bz < 0 || bz >= P.nz) {
        // voxel_grid_2d.py:133-138 raises ValueError; counted and reported to the host
        atomicAdd(&stats[DODT_BEV_STAT_OOB], 1);
        live = false;
      } else {
        int yb = voxel_bin(yd, P.voxel) - P.min_y_biased;
        yb = yb < 0 ? 0 : (yb > static_cast<int>(P.yb_max) ? static_cast<int>(P.yb_max) : yb);
        // rotated output: out[r][c] = grid[ix = c][iz = nz-1-r]  (bev_slices.py:115-116)
        cell = (P.nz - 1 - bz) * P.nx + bx;
        const unsigned key = (static_cast<unsigned>(yb) << P.idx_bits) |
                             static_cast<unsigned>(idx);
        inv_key = 0xFFFFFFFFu - key;  // larger == lower y-bin, then lower index; never 0
        if (P.has_occ == 2) occ[static_cast<size_t>(bx) * P.nz + bz] = 1;  // unfiltered occupancy
      }
    }
    // plane side, get_point_filter: dot(plane - [0,0,0,off], [x,y,z,1]) < 0, obj_utils.py:488-492
    const double base =
        __dadd_rn(__dadd_rn(__dmul_rn(P.a, xd), __dmul_rn(P.b, yd)), __dmul_rn(P.c, zd));

    for (int m = 0; m < n_pred; ++m) {
      bool in_m = live;
      if (P.filter_mode) {
        const bool below_hi = __dadd_rn(base, P.dhi[m]) < 0.0;
        const bool below_lo = __dadd_rn(base, P.dlo[m]) < 0.0;
        in_m = in_m && (below_hi != below_lo);  // np.logical_xor, kitti_utils.py:108
      } else {
        in_m = in_m && (m == 0 || m == P.S);   // plain voxelize_2d: slice 0 + counts
      }
      if (m == n_maps && !P.has_occ) in_m = false;
      const unsigned active = __ballot_sync(0xffffffffu, in_m);
      if (active == 0) continue;
      if (lane == 0) atomicAdd(&st.slice_pts[m], __popc(active));
      if (m == n_maps) {  // occupancy slice of the anchor filter (leaf_layout_2d)
        if (in_m && P.has_occ == 1) {
          const int bz = P.nz - 1 - cell / P.nx;
          const int bx = cell % P.nx;
          occ[static_cast<size_t>(bx) * P.nz + bz] = 1;
        }
        continue;
      }
      if (in_m) {
        const unsigned e = static_cast<unsigned>(m) * HW + cell;
        const unsigned peers = __match_any_sync(active, e);
        const bool leader = lane == __ffs(peers) - 1;
        if (m < P.S) {
          const unsigned best = __reduce_max_sync(peers, inv_key);
          if (leader) atomicMax(&maps[e], best);                                   // RED.MAX
        } else {
          if (leader) atomicAdd(&maps[e], static_cast<unsigned>(__popc(peers)));   // RED.ADD
        }
      }
    }
  }

  __syncthreads();
  if (threadIdx.x < n_pred && st.slice_pts[threadIdx.x]) {
    const int m = threadIdx.x;
    const int slot = m < P.S ? m : (m == P.S ? DODT_BEV_STAT_DENSITY : DODT_BEV_STAT_OCC);
    atomicAdd(&stats[slot], st.slice_pts[m]);
  }
}

// zero `bytes` bytes at p (16-byte aligned) with 16-byte stores, then a byte tail
__device__ __forceinline__ void clear_span(unsigned char *p, size_t bytes) {
  const size_t n16 = bytes / 16;
  uint4 *v = reinterpret_cast<uint4 *>(p);
  const size_t stride = static_cast<size_t>(gridDim.x) * kBlock;
  for (size_t i = static_cast<size_t>(blockIdx.x) * kBlock + threadIdx.x; i < n16; i += stride)
    v[i] = make_uint4(0u, 0u, 0u, 0u);
  if (blockIdx.x == 0)
    for (size_t i = n16 * 16 + threadIdx.x; i < bytes; i += kBlock) p[i] = 0;
}

__global__ void __launch_bounds__(kBlock)
bev_clear(uint4 *__restrict__ maps, size_t map_bytes, unsigned char *__restrict__ occ, size_t occ_bytes,
          int *__restrict__ stats) {
  clear_span(reinterpret_cast<unsigned char *>(maps), map_bytes);
  if (occ_bytes) clear_span(occ, occ_bytes);
  if (blockIdx.x == 0 && threadIdx.x < DODT_BEV_STATS_LEN) stats[threadIdx.x] = 0;
}

// one map entry: key / count -> final float (see the file header)
template <typename T>
__device__ __forceinline__ unsigned resolve_entry(const T *__restrict__ px, const T *__restrict__ py,
                                                  const T *__restrict__ pz, const BevDev &P,
                                                  const int *__restrict__ stats, unsigned raw, int m,
                                                  int cell, unsigned e, int *__restrict__ winner_idx,
                                                  int *__restrict__ counts) {
  if (m < P.S) {
    if (P.filter_mode && stats[m] <= 1) {
      // bev_slices.py:76-99: a slice with 0 or 1 points is replaced by one origin point
      return cell == P.origin_cell ? __float_as_uint(P.origin_val[m]) : 0u;
    }
    const unsigned key = 0xFFFFFFFFu - raw;
    const unsigned idx = key & P.idx_mask;
    const double xd = static_cast<double>(__ldg(px + idx));
    const double yd = static_cast<double>(__ldg(py + idx));
    const double zd = static_cast<double>(__ldg(pz + idx));
    double h;
    if (P.height_from_y) {
      h = yd;  // voxel_grid_2d.py:105-106 (no ground plane)
    } else {
      // dist_to_plane, geometry_utils.py:40: ((a*x + b*y) + c*z + d) / sqrt(a^2+b^2+c^2)
      h = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(P.a, xd), __dmul_rn(P.b, yd)),
                              __dmul_rn(P.c, zd)), P.d);
      h = __ddiv_rn(h, P.norm);
    }
    // bev_slices.py:107-109: (height - slice_lo) / height_per_division
    const double v = __ddiv_rn(__dsub_rn(h, P.lo[m]), P.hpd);
    if (winner_idx) winner_idx[e] = static_cast<int>(idx);
    return __float_as_uint(__double2float_rn(v));
  }
  // bev_generator.py:34-35: min(1, log(n + 1) / norm) through the host LUT
  if (counts) counts[cell] = static_cast<int>(raw);
  return __float_as_uint(raw < static_cast<unsigned>(P.lut_len) ? P.lut[raw] : 1.0f);
}

// VEC = 4 (nx*nz a multiple of 4; uint4 loads): the maps are ~93 % zeros, so each warp first
// sweeps its entries and queues the non-empty ones in shared memory, then resolves the queue with
// all 32 lanes busy — the fp64 evaluation is not executed once per sparse hit with 31 lanes idle.
// VEC = 1: plain entry-per-thread scan for odd grid sizes.
template <typename T, int VEC>
__global__ void __launch_bounds__(kBlock)
bev_resolve_scan(const T *__restrict__ px, const T *__restrict__ py, const T *__restrict__ pz,
                 const __grid_constant__ BevDev P, unsigned *__restrict__ maps,
                 const int *__restrict__ stats, int *__restrict__ winner_idx,
                 int *__restrict__ counts) {
  const int HW = P.nx * P.nz;
  const unsigned total = static_cast<unsigned>(HW) * (P.S + 1);
  if (VEC == 4) {
    constexpr int kQueue = 32 + 4 * 32;
    __shared__ uint2 queue[kBlock / 32][kQueue];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint2 *q = queue[warp];
    int qn = 0;                                              // warp-uniform
    auto drain = [&](int first, int count) {                 // resolve q[first .. first+count)
      if (lane < count) {
        const uint2 it = q[first + lane];
        const int m = it.x / HW;
        maps[it.x] = resolve_entry<T>(px, py, pz, P, stats, it.y, m, it.x - m * HW, it.x, winner_idx, counts);
      }
    };
    // kScanLoads independent 16-byte loads per lane are in flight before the first is inspected:
    // with one load per iteration the scan was latency-bound (ncu, dense config: long-scoreboard
    // 11.4 stalls per issue, 2.2 TB/s)
    constexpr int kScanLoads = 4;
    const unsigned warp_span = 32 * 4 * kScanLoads;
    const unsigned stride = gridDim.x * (kBlock / 32) * warp_span;
    for (unsigned base = (blockIdx.x * (kBlock / 32) + warp) * warp_span; base < total; base += stride) {
      uint4 v[kScanLoads];
#pragma unroll
      for (int u = 0; u < kScanLoads; ++u) {
        const unsigned e0 = base + u * 128 + lane * 4;
        v[u] = make_uint4(0u, 0u, 0u, 0u);
        if (e0 < total) v[u] = *reinterpret_cast<const uint4 *>(maps + e0);
      }
#pragma unroll
      for (int u = 0; u < kScanLoads; ++u) {
        const unsigned e0 = base + u * 128 + lane * 4;
        const unsigned raw[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const bool nz = raw[k] != 0u;
          const unsigned mask = __ballot_sync(0xffffffffu, nz);
          if (nz) q[qn + __popc(mask & ((1u << lane) - 1u))] = make_uint2(e0 + k, raw[k]);
          qn += __popc(mask);
        }
        __syncwarp();
        while (qn >= 32) {
          qn -= 32;
          drain(qn, 32);
          __syncwarp();
        }
      }
    }
    drain(0, qn);
  } else {
    const unsigned stride = gridDim.x * kBlock;
    for (unsigned e = blockIdx.x * kBlock + threadIdx.x; e < total; e += stride) {
      const unsigned raw = maps[e];
      if (raw == 0u) continue;
      const int m = e / HW;
      maps[e] = resolve_entry<T>(px, py, pz, P, stats, raw, m, e - m * HW, e, winner_idx, counts);
    }
  }
  if (blockIdx.x == 0 && P.filter_mode && threadIdx.x < P.S && P.origin_cell >= 0) {
    const int s = threadIdx.x;
    if (stats[s] <= 1)
      maps[static_cast<size_t>(s) * HW + P.origin_cell] = __float_as_uint(P.origin_val[s]);
  }
}

int host_floor_div(double v, double voxel) { return static_cast<int>(floor(v / voxel)); }

}  // namespace
}  // namespace dodt

extern "C" {

int dodt_bev_grid(const double extents[6], double voxel_size, int32_t grid[6]) {
  if (!extents || !grid || !(voxel_size > 0.0)) return DODT_EINVAL;
  for (int ax = 0; ax < 3; ++ax) {
    // voxel_grid_2d.py:125-127: floor(min / voxel), ceil(max / voxel - 1)
    const double lo = floor(extents[2 * ax] / voxel_size);
    const double hi = ceil(extents[2 * ax + 1] / voxel_size - 1.0);
    if (!(hi >= lo) || hi - lo > 1.0e8 || fabs(lo) > 1.0e9) return DODT_ESHAPE;
    grid[ax] = static_cast<int32_t>(hi - lo + 1.0);
    grid[3 + ax] = static_cast<int32_t>(lo);
  }
  return DODT_OK;
}

size_t dodt_bev_workspace_bytes(int64_t n_points, int32_t num_slices, int32_t nx, int32_t nz) {
  if (n_points < 0 || num_slices < 0 || nx <= 0 || nz <= 0) return 0;
  return 256;   // no scratch is needed any more (keys and counts live in the maps); kept in the ABI
}

int dodt_bev_slices(const void *pts, int32_t pts_dtype, int64_t n, const int32_t *n_dev,
                    int64_t row_stride,
                    const dodt_bev_params *p, float *maps, uint8_t *occ, int32_t *stats,
                    int32_t *winner_idx, int32_t *counts, void *workspace,
                    size_t workspace_bytes, dodt_stream_t stream_) {
  using namespace dodt;
  if (!p || !maps || !stats || n < 0 || (n > 0 && !pts) || row_stride < n) return DODT_EINVAL;
  if (pts_dtype != DODT_F32 && pts_dtype != DODT_F64) return DODT_EINVAL;
  if (p->num_slices < 0 || p->num_slices > DODT_MAX_SLICES) return DODT_EINVAL;
  if (p->filter_mode && p->num_slices < 0) return DODT_EINVAL;
  if (!p->filter_mode && p->num_slices != 1) return DODT_EINVAL;
  if (p->density_lut_len < 0 || p->density_lut_len > DODT_MAX_DENSITY_LUT) return DODT_EINVAL;
  int32_t grid[6];
  int rc = dodt_bev_grid(p->extents, p->voxel_size, grid);
  if (rc != DODT_OK) return rc;
  cudaStream_t stream = as_stream(stream_);

  BevDev P;
  memset(&P, 0, sizeof(P));
  const int S = p->num_slices;
  P.a = p->plane[0]; P.b = p->plane[1]; P.c = p->plane[2]; P.d = p->plane[3];
  P.height_from_y = (P.a == 0.0 && P.b == 0.0 && P.c == 0.0) ? 1 : 0;
  if (P.height_from_y && p->filter_mode) return DODT_EINVAL;
  // volatile: keep each product/sum individually rounded like NumPy does
  volatile double aa = P.a * P.a, bb = P.b * P.b, cc = P.c * P.c;
  volatile double s2 = aa + bb;
  s2 = s2 + cc;
  P.norm = P.height_from_y ? 1.0 : sqrt(s2);
  for (int i = 0; i < 6; ++i) P.ext[i] = p->extents[i];
  P.voxel = p->voxel_size;
  P.S = S;
  P.filter_mode = p->filter_mode ? 1 : 0;
  if (P.filter_mode) {
    // bev_slices.py:30-31,63-64
    volatile double hpd = (p->height_hi - p->height_lo) / static_cast<double>(S > 0 ? S : 1);
    P.hpd = hpd;
    for (int s = 0; s < S; ++s) {
      volatile double lo = p->height_lo + static_cast<double>(s) * hpd;
      volatile double hi = lo + hpd;
      P.lo[s] = lo;
      volatile double dhi = P.d + (-hi), dlo = P.d + (-lo);  // obj_utils.py:488
      P.dhi[s] = dhi;
      P.dlo[s] = dlo;
    }
    volatile double t;
    t = P.d + (-p->height_hi); P.dhi[S] = t;
    t = P.d + (-p->height_lo); P.dlo[S] = t;
    t = P.d + (-p->occ_hi); P.dhi[S + 1] = t;
    t = P.d + (-p->occ_lo); P.dlo[S + 1] = t;
  } else {
    P.hpd = 1.0;  // raw heights: (h - 0) / 1
  }
  P.has_occ = occ ? (P.filter_mode ? 1 : 2) : 0;
  P.nx = grid[0]; P.nz = grid[2];
  P.min_x = grid[3]; P.min_z = grid[5];
  // y bins only order the points inside a cell; one spare bin on each side absorbs the
  // rounding of a quotient that lands exactly on the extent
  const int ny = grid[1] + 2;
  P.min_y_biased = grid[4] - 1;
  int ybits = 1;
  while ((1 << ybits) < ny + 1) ++ybits;  // yb <= ny-1 < 2^ybits - 1, so key != 0xFFFFFFFF
  if (ybits > 20) return DODT_ECAPACITY;
  P.idx_bits = 32 - ybits;
  P.idx_mask = (P.idx_bits >= 32) ? 0xFFFFFFFFu : ((1u << P.idx_bits) - 1u);
  P.yb_max = static_cast<unsigned>(ny - 1);
  if (static_cast<uint64_t>(n) > (1ull << P.idx_bits)) return DODT_ECAPACITY;
  const int64_t HW = static_cast<int64_t>(P.nx) * P.nz;
  if (HW * (S + 1) > 0x7FFFFFFFll) return DODT_ECAPACITY;
  P.lut_len = p->density_lut_len;
  for (int i = 0; i < P.lut_len; ++i) P.lut[i] = static_cast<float>(p->density_lut[i]);
  // origin fallback cell and values (bev_slices.py:86-99)
  P.origin_cell = -1;
  if (P.filter_mode) {
    const int ox = host_floor_div(0.0, P.voxel) - P.min_x;
    const int oz = host_floor_div(0.0, P.voxel) - P.min_z;
    if (ox >= 0 && ox < P.nx && oz >= 0 && oz < P.nz) P.origin_cell = (P.nz - 1 - oz) * P.nx + ox;
    for (int s = 0; s < S; ++s) {
      volatile double h = P.d / P.norm;
      volatile double v = (h - P.lo[s]) / P.hpd;
      P.origin_val[s] = static_cast<float>(v);
    }
  }
  (void)workspace;
  (void)workspace_bytes;

  // pass 0: maps, occupancy grid and counters cleared by ONE launch of a few long-lived CTAs (three
  // driver memsets — 1640 short CTAs for the maps alone — cost the co-running kernels of the frame
  // pipeline more)
  static int clear_blocks = -1, resolve_blocks_env = -1;
  if (clear_blocks < 0) {
    clear_blocks = DODT_KNOB("DODT_BEV_CLEAR_BLOCKS", 4 * kNumSMs);   // frame pipeline: 0 (driver memsets) 10.80, 296 10.98,
    resolve_blocks_env = DODT_KNOB("DODT_BEV_RESOLVE_BLOCKS", 4 * kNumSMs);   // 592 11.03 k frames/s; resolve 1184 -> 592 CTAs: 11.07
  }
  if (clear_blocks > 0 && reinterpret_cast<uintptr_t>(maps) % 16 == 0 && (!occ || reinterpret_cast<uintptr_t>(occ) % 16 == 0)) {
    bev_clear<<<clear_blocks, kBlock, 0, stream>>>(reinterpret_cast<uint4 *>(maps),
                                                   static_cast<size_t>(HW) * (S + 1) * sizeof(float), occ,
                                                   occ ? static_cast<size_t>(HW) : 0, stats);
    DODT_AFTER_LAUNCH();
  } else {
    DODT_CUDA_TRY(cudaMemsetAsync(maps, 0, sizeof(float) * HW * (S + 1), stream));
    DODT_CUDA_TRY(cudaMemsetAsync(stats, 0, sizeof(int32_t) * DODT_BEV_STATS_LEN, stream));
    if (occ) DODT_CUDA_TRY(cudaMemsetAsync(occ, 0, HW, stream));
  }
  if (winner_idx) DODT_CUDA_TRY(cudaMemsetAsync(winner_idx, 0xFF, sizeof(int32_t) * HW * S, stream));
  if (counts) DODT_CUDA_TRY(cudaMemsetAsync(counts, 0, sizeof(int32_t) * HW, stream));
  if (n == 0) {
    // no point at all: every height slice is degenerate; resolve still applies the origin rule
  }

  unsigned *umaps = reinterpret_cast<unsigned *>(maps);
  const int resolve_blocks = resolve_blocks_env;
  if (pts_dtype == DODT_F32) {
    const float *px = static_cast<const float *>(pts);
    const float *py = px + row_stride, *pz = px + 2 * row_stride;
    if (n > 0) {
      // one point per thread keeps a 120k-point frame inside a single wave of 148 SMs; dense
      // clouds take 4 points per thread through float4 loads
      const bool dense = n > 4ll * kNumSMs * 2048;
      const int vec = dense ? 4 : 1;
      const int vec_ok = (reinterpret_cast<uintptr_t>(px) % 16 == 0 && row_stride % 4 == 0) ? 1 : 0;
      const int blocks = ceil_div(n, static_cast<int64_t>(kBlock) * vec);
      if (dense)
        bev_accumulate<float, 4><<<blocks, kBlock, 0, stream>>>(px, py, pz, n, n_dev, vec_ok, P, umaps, occ, stats);
      else
        bev_accumulate<float, 1><<<blocks, kBlock, 0, stream>>>(px, py, pz, n, n_dev, 1, P, umaps, occ, stats);
      DODT_AFTER_LAUNCH();
    }
    if (HW % 4 == 0)
      bev_resolve_scan<float, 4><<<resolve_blocks, kBlock, 0, stream>>>(px, py, pz, P, umaps, stats, winner_idx, counts);
    else
      bev_resolve_scan<float, 1><<<resolve_blocks, kBlock, 0, stream>>>(px, py, pz, P, umaps, stats, winner_idx, counts);
    DODT_AFTER_LAUNCH();
  } else {
    const double *px = static_cast<const double *>(pts);
    const double *py = px + row_stride, *pz = px + 2 * row_stride;
    if (n > 0) {
      const bool dense = n > 2ll * kNumSMs * 2048;
      const int vec = dense ? 2 : 1;
      const int vec_ok = (reinterpret_cast<uintptr_t>(px) % 16 == 0 && row_stride % 2 == 0) ? 1 : 0;
      const int blocks = ceil_div(n, static_cast<int64_t>(kBlock) * vec);
      if (dense)
        bev_accumulate<double, 2><<<blocks, kBlock, 0, stream>>>(px, py, pz, n, n_dev, vec_ok, P, umaps, occ, stats);
      else
        bev_accumulate<double, 1><<<blocks, kBlock, 0, stream>>>(px, py, pz, n, n_dev, 1, P, umaps, occ, stats);
      DODT_AFTER_LAUNCH();
    }
    if (HW % 4 == 0)
      bev_resolve_scan<double, 4><<<resolve_blocks, kBlock, 0, stream>>>(px, py, pz, P, umaps, stats, winner_idx, counts);
    else
      bev_resolve_scan<double, 1><<<resolve_blocks, kBlock, 0, stream>>>(px, py, pz, P, umaps, stats, winner_idx, counts);
    DODT_AFTER_LAUNCH();
  }
  return DODT_OK;
}

}  // extern "C"
