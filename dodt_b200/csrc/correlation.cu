// S4 — FlowNet-style correlation of two NHWC float32 feature maps (forward), sm_100a.
//
// Replaces the reference's TensorFlow custom op (paths relative to the Guoxs/DODT checkout):
//   avod/core/ops/correlation/correlation_op.cc:9-62        op registration + shape function
//   avod/core/ops/correlation/correlation_kernel.cc:26-124  output shape, two padded temporaries
//   avod/core/ops/correlation/pad.cu.cc:14-74               PadData (memset + uncoalesced copy)
//   avod/core/ops/correlation/correlation_kernel.cu.cc:21-119 CorrelateData (one warp per output
//                                                            pixel, B re-read D^2 times from HBM)
//
//   out[n,y,x,k] = 1/(ks^2*C) * sum_{j,i<ks} sum_c  Apad[n, y*s1+md+j,       x*s1+md+i,       c]
//                                                 * Bpad[n, y*s1+md+j+s2p,   x*s1+md+i+s2o,   c]
//   s2o = (k mod Wn - r)*s2, s2p = (k div Wn - r)*s2, r = md div s2, Wn = 2r+1,
//   Apad/Bpad = inputs zero-padded by `pad` on H and W.
//
// Padding is never materialised: out-of-image taps are predicated to zero.
//
// Kernels
//   corr_generic   any (kernel_size, max_displacement, stride_1, stride_2, pad, C): one thread per
//                  (output pixel, displacement); float4 over channels when C % 4 == 0.
//   corr_tile_k1   the DODT configuration family (kernel_size 1, stride_1 1, C % 4 == 0,
//                  (2r+1) <= 7): a CTA computes a TH x TW tile of output pixels; the A tile and the
//                  B tile with its r*s2 halo are staged once in shared memory (zero-filled outside
//                  the image, which is the reference's padding), each thread owns PX pixels spaced
//                  s2 apart along x so that one B column feeds up to PX of its (pixel, displacement)
//                  accumulators, and all (2r+1)^2 results of a pixel are produced from registers.
//                  B is read from HBM once instead of D^2 times and nothing is padded in memory.
#include "common.cuh"

namespace dodt {

// correlation_tma.cu: TMA-pipelined persistent kernel; returns 1 when it does not apply
int correlation_tma(const float *a, const float *b, int N, int H, int W, int C, int r, int out_h,
                    int out_w, int shift, float *out, int max_ctas, cudaStream_t stream);

int correlation_stream_tma(const float *const *maps, int n_pairs, float *const *outs, int H, int W,
                           int C, int r, int out_h, int out_w, int shift, int max_ctas,
                           cudaStream_t stream);

// correlation_feed.cu: TMA-fed persistent kernel (16-channel chunks); returns 1 when it does not apply
int correlation_feed(const float *a, const float *b, int N, int H, int W, int C, int r, int out_h,
                     int out_w, int shift, float *out, int max_ctas, cudaStream_t stream);

int correlation_stream_feed(const float *const *maps, int n_pairs, float *const *outs, int H, int W,
                            int C, int r, int out_h, int out_w, int shift, int max_ctas,
                            cudaStream_t stream);

namespace {

// diagnostic build only: DODT_CORR_FEED=0 routes around the TMA-fed kernel (A/B timing)
inline bool use_feed() {
  static int v = -1;
  if (v < 0) v = DODT_KNOB("DODT_CORR_FEED", 1);
  return v != 0;
}

struct CorrGeom {
  int batch, H, W, C;
  int ks, md, s1, s2, pad;
  int out_h, out_w, out_c;
  int r, wn;
};

template <int VEC>
__global__ void __launch_bounds__(256)
corr_generic(const float *__restrict__ a, const float *__restrict__ b, CorrGeom g,
             long long total, float *__restrict__ out) {
  const long long t = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (t >= total) return;
  const int k = static_cast<int>(t % g.out_c);
  long long rr = t / g.out_c;
  const int x = static_cast<int>(rr % g.out_w);
  rr /= g.out_w;
  const int y = static_cast<int>(rr % g.out_h);
  const int n = static_cast<int>(rr / g.out_h);

  const int s2o = (k % g.wn - g.r) * g.s2;
  const int s2p = (k / g.wn - g.r) * g.s2;
  // top-left corner of the kernel patch in UNPADDED coordinates
  const int ay0 = y * g.s1 + g.md - g.pad;
  const int ax0 = x * g.s1 + g.md - g.pad;
  const float *an = a + static_cast<size_t>(n) * g.H * g.W * g.C;
  const float *bn = b + static_cast<size_t>(n) * g.H * g.W * g.C;

  float acc = 0.0f;
  for (int j = 0; j < g.ks; ++j) {
    const int ay = ay0 + j, by = ay + s2p;
    if (ay < 0 || ay >= g.H || by < 0 || by >= g.H) continue;  // a zero-padded operand
    for (int i = 0; i < g.ks; ++i) {
      const int ax = ax0 + i, bx = ax + s2o;
      if (ax < 0 || ax >= g.W || bx < 0 || bx >= g.W) continue;
      const float *pa = an + (static_cast<size_t>(ay) * g.W + ax) * g.C;
      const float *pb = bn + (static_cast<size_t>(by) * g.W + bx) * g.C;
      if (VEC == 4) {
        for (int c = 0; c < g.C; c += 4) {
          const float4 va = __ldg(reinterpret_cast<const float4 *>(pa + c));
          const float4 vb = __ldg(reinterpret_cast<const float4 *>(pb + c));
          acc = fmaf(va.x, vb.x, acc);
          acc = fmaf(va.y, vb.y, acc);
          acc = fmaf(va.z, vb.z, acc);
          acc = fmaf(va.w, vb.w, acc);
        }
      } else {
        for (int c = 0; c < g.C; ++c) acc = fmaf(__ldg(pa + c), __ldg(pb + c), acc);
      }
    }
  }
  const float sumelems = static_cast<float>(g.ks * g.ks * g.C);
  out[t] = __fdiv_rn(acc, sumelems);
}

// ---------------------------------------------------------------------------------------------
// Tiled kernel for kernel_size == 1, stride_1 == 1.
// ---------------------------------------------------------------------------------------------
constexpr int kTW = 64;      // output pixels per tile along x
constexpr int kTH = 8;       // output rows per tile
constexpr int kPX = 4;       // pixels per thread, spaced s2 apart
constexpr int kCC = 8;       // channels staged per pass (2 float4)
constexpr int kTileThreads = kTW / kPX * kTH;  // 128

// Shared-memory pixel vectors are kCC floats = 32 bytes. The two 16-byte halves of pixel `p` are
// swapped when bit 2 of p is set, so that eight lanes reading eight different pixels (p mod 8 all
// distinct) hit eight different 16-byte bank groups.
__device__ __forceinline__ int smem_off(int pixel, int half) {
  return pixel * kCC + ((half ^ ((pixel >> 2) & 1)) << 2);
}

template <int R>
__global__ void __launch_bounds__(kTileThreads)
corr_tile_k1(const float *__restrict__ a, const float *__restrict__ b, CorrGeom g,
             float *__restrict__ out) {
  constexpr int WN = 2 * R + 1;
  constexpr int D2 = WN * WN;
  constexpr int NB = kPX + 2 * R;  // B columns a thread touches per displacement row
  extern __shared__ __align__(16) float smem[];
  const int halo = R * g.s2;
  const int bw = kTW + 2 * halo;          // B tile width in pixels
  // pitch = 2 (mod 8): lanes of a quarter-warp sit on four consecutive rows x two parities
  const int bpitch = ((bw + 7) / 8) * 8 + 2;
  const int bh = kTH + 2 * halo;
  const int apitch = kTW + 2;
  float *sa = smem;                              // [kTH][apitch][kCC]
  float *sb = smem + kTH * apitch * kCC;         // [bh][bpitch][kCC]

  const int n = blockIdx.z;
  const int ty0 = blockIdx.y * kTH, tx0 = blockIdx.x * kTW;
  const int shift = g.md - g.pad;  // output pixel (y,x) reads A at (y+shift, x+shift)
  const float *an = a + static_cast<size_t>(n) * g.H * g.W * g.C;
  const float *bn = b + static_cast<size_t>(n) * g.H * g.W * g.C;

  // thread -> (row, parity, group): lane bits [0]=parity, [1..2]=row&3, [3..4]=group&3
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int kGroupsX = kTW / (kPX * 2);     // groups of 2*kPX pixels along x (8)
  // a warp covers 4 rows x 4 groups; warps tile (kTH/4) x (kGroupsX/4)
  const int wy = warp / (kGroupsX / 4), wx = warp % (kGroupsX / 4);
  const int row = wy * 4 + ((lane >> 1) & 3);
  const int grp = wx * 4 + (lane >> 3);
  const int par = lane & 1;
  // with s2 == 2 the thread's pixels are x0, x0+2, x0+4, x0+6; for other s2 they are x0 + j*s2
  // inside a group of kPX*s2 pixels, parity selecting the residue (only s2 == 2 uses this kernel)
  const int x0 = grp * (kPX * 2) + par;

  float acc[kPX][D2];
#pragma unroll
  for (int j = 0; j < kPX; ++j)
#pragma unroll
    for (int k = 0; k < D2; ++k) acc[j][k] = 0.0f;

  for (int c0 = 0; c0 < g.C; c0 += kCC) {
    __syncthreads();
    // ---- stage A tile (rows ty0+shift .., cols tx0+shift ..) ----
    for (int e = threadIdx.x; e < kTH * kTW * 2; e += kTileThreads) {
      const int half = e & 1, p = e >> 1;
      const int py = p / kTW, px = p % kTW;
      const int gy = ty0 + py + shift, gx = tx0 + px + shift;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gy >= 0 && gy < g.H && gx >= 0 && gx < g.W)
        v = __ldg(reinterpret_cast<const float4 *>(an + (static_cast<size_t>(gy) * g.W + gx) * g.C + c0) + half);
      *reinterpret_cast<float4 *>(sa + smem_off(py * apitch + px, half)) = v;
    }
    // ---- stage B tile with halo ----
    for (int e = threadIdx.x; e < bh * bw * 2; e += kTileThreads) {
      const int half = e & 1, p = e >> 1;
      const int py = p / bw, px = p % bw;
      const int gy = ty0 + py + shift - halo, gx = tx0 + px + shift - halo;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gy >= 0 && gy < g.H && gx >= 0 && gx < g.W)
        v = __ldg(reinterpret_cast<const float4 *>(bn + (static_cast<size_t>(gy) * g.W + gx) * g.C + c0) + half);
      *reinterpret_cast<float4 *>(sb + smem_off(py * bpitch + px, half)) = v;
    }
    __syncthreads();

#pragma unroll
    for (int half = 0; half < 2; ++half) {
      float4 va[kPX];
#pragma unroll
      for (int j = 0; j < kPX; ++j)
        va[j] = *reinterpret_cast<const float4 *>(sa + smem_off(row * apitch + x0 + j * 2, half));
#pragma unroll
      for (int p = 0; p < WN; ++p) {
        float4 vb[NB];
        const int brow = row + p * g.s2;  // (row + halo) + (p - R) * s2
#pragma unroll
        for (int q = 0; q < NB; ++q)
          vb[q] = *reinterpret_cast<const float4 *>(sb + smem_off(brow * bpitch + x0 + q * 2, half));
#pragma unroll
        for (int j = 0; j < kPX; ++j)
#pragma unroll
          for (int o = 0; o < WN; ++o) {
            float s = acc[j][p * WN + o];
            s = fmaf(va[j].x, vb[j + o].x, s);
            s = fmaf(va[j].y, vb[j + o].y, s);
            s = fmaf(va[j].z, vb[j + o].z, s);
            s = fmaf(va[j].w, vb[j + o].w, s);
            acc[j][p * WN + o] = s;
          }
      }
    }
  }

  const float sumelems = static_cast<float>(g.C);
  const int oy = ty0 + row;
  if (oy < g.out_h) {
#pragma unroll
    for (int j = 0; j < kPX; ++j) {
      const int ox = tx0 + x0 + j * 2;
      if (ox < g.out_w) {
        float *dst = out + ((static_cast<size_t>(n) * g.out_h + oy) * g.out_w + ox) * D2;
#pragma unroll
        for (int k = 0; k < D2; ++k) dst[k] = __fdiv_rn(acc[j][k], sumelems);
      }
    }
  }
}

int fill_geom(int32_t batch, int32_t H, int32_t W, int32_t C, int32_t ks, int32_t md, int32_t s1,
              int32_t s2, int32_t pad, CorrGeom *g) {
  if (batch <= 0 || H <= 0 || W <= 0 || C <= 0 || ks <= 0 || md < 0 || s1 <= 0 || s2 <= 0 || pad < 0)
    return DODT_EINVAL;
  if (ks % 2 == 0) return DODT_EINVAL;  // "kernel_size must be odd", correlation_kernel.cc:23
  int32_t hwc[3];
  const int rc = dodt_correlation_out_shape(H, W, ks, md, s1, s2, pad, hwc);
  if (rc != DODT_OK) return rc;
  g->batch = batch; g->H = H; g->W = W; g->C = C;
  g->ks = ks; g->md = md; g->s1 = s1; g->s2 = s2; g->pad = pad;
  g->out_h = hwc[0]; g->out_w = hwc[1]; g->out_c = hwc[2];
  g->r = md / s2; g->wn = 2 * g->r + 1;
  return DODT_OK;
}

template <int R>
int launch_tile(const float *a, const float *b, const CorrGeom &g, float *out, cudaStream_t stream) {
  const int halo = R * g.s2;
  const int bw = kTW + 2 * halo, bpitch = ((bw + 7) / 8) * 8 + 2, bh = kTH + 2 * halo;
  const size_t smem = (static_cast<size_t>(kTH) * (kTW + 2) + static_cast<size_t>(bh) * bpitch) * kCC * sizeof(float);
  if (smem > 200 * 1024) return 1;  // not applicable: fall back to corr_generic
  if (smem > 48 * 1024)
    DODT_CUDA_TRY(cudaFuncSetAttribute(corr_tile_k1<R>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem)));
  dim3 grid(ceil_div(g.out_w, kTW), ceil_div(g.out_h, kTH), g.batch);
  corr_tile_k1<R><<<grid, kTileThreads, smem, stream>>>(a, b, g, out);
  DODT_AFTER_LAUNCH();
  return DODT_OK;
}

}  // namespace
}  // namespace dodt

extern "C" {

int dodt_correlation_out_shape(int32_t height, int32_t width, int32_t kernel_size,
                               int32_t max_displacement, int32_t stride_1, int32_t stride_2,
                               int32_t pad, int32_t out_hwc[3]) {
  if (!out_hwc || height <= 0 || width <= 0 || kernel_size <= 0 || max_displacement < 0 ||
      stride_1 <= 0 || stride_2 <= 0 || pad < 0)
    return DODT_EINVAL;
  if (kernel_size % 2 == 0) return DODT_EINVAL;  // correlation_kernel.cc:23 "kernel_size must be odd"
  // correlation_kernel.cc:39-57 (float ceil, as the reference computes it)
  const int kernel_radius = (kernel_size - 1) / 2;
  const int border = max_displacement + kernel_radius;
  const int ph = height + 2 * pad, pw = width + 2 * pad;
  const int oh = static_cast<int>(ceilf(static_cast<float>(ph - border * 2) / static_cast<float>(stride_1)));
  const int ow = static_cast<int>(ceilf(static_cast<float>(pw - border * 2) / static_cast<float>(stride_1)));
  if (oh < 1 || ow < 1) return DODT_ESHAPE;  // "Neighborhood and kernel don't fit in input"
  const int r = max_displacement / stride_2;
  out_hwc[0] = oh;
  out_hwc[1] = ow;
  out_hwc[2] = (2 * r + 1) * (2 * r + 1);
  return DODT_OK;
}

int dodt_correlation(const float *a, const float *b, int32_t batch, int32_t height, int32_t width,
                     int32_t channels, int32_t kernel_size, int32_t max_displacement,
                     int32_t stride_1, int32_t stride_2, int32_t pad, float *out,
                     dodt_stream_t stream_) {
  return dodt_correlation_shared(a, b, batch, height, width, channels, kernel_size, max_displacement,
                                 stride_1, stride_2, pad, out, 0, stream_);
}

int dodt_correlation_shared(const float *a, const float *b, int32_t batch, int32_t height,
                            int32_t width, int32_t channels, int32_t kernel_size,
                            int32_t max_displacement, int32_t stride_1, int32_t stride_2,
                            int32_t pad, float *out, int32_t max_ctas, dodt_stream_t stream_) {
  using namespace dodt;
  if (max_ctas < 0) return DODT_EINVAL;
  if (!a || !b || !out) return DODT_EINVAL;
  CorrGeom g;
  const int rc = fill_geom(batch, height, width, channels, kernel_size, max_displacement, stride_1,
                           stride_2, pad, &g);
  if (rc != DODT_OK) return rc;
  cudaStream_t stream = as_stream(stream_);
  const bool aligned = reinterpret_cast<uintptr_t>(a) % 16 == 0 && reinterpret_cast<uintptr_t>(b) % 16 == 0;
  // The reference reads the padded temporaries without bounds checks, so parameters that make a
  // displaced patch leave the padded image (pad < max_displacement with large displacements) are
  // undefined there; here such taps are zero.
  if (g.ks == 1 && g.s1 == 1 && g.s2 == 2 && aligned && use_feed()) {
    const int done = correlation_feed(a, b, g.batch, g.H, g.W, g.C, g.r, g.out_h, g.out_w,
                                      g.md - g.pad, out, max_ctas, stream);
    if (done <= 0) return done;
  }
  if (g.ks == 1 && g.s1 == 1 && g.s2 == 2 && aligned) {
    const int done = correlation_tma(a, b, g.batch, g.H, g.W, g.C, g.r, g.out_h, g.out_w,
                                     g.md - g.pad, out, max_ctas, stream);
    if (done <= 0) return done;
  }
  if (g.ks == 1 && g.s1 == 1 && g.s2 == 2 && channels % kCC == 0 && aligned && g.batch <= 65535) {
    int done = 1;
    switch (g.r) {
      case 1: done = launch_tile<1>(a, b, g, out, stream); break;
      case 2: done = launch_tile<2>(a, b, g, out, stream); break;
      default: break;
    }
    if (done <= 0) return done;  // launched (DODT_OK) or failed (DODT_E*)
  }
  const long long total = static_cast<long long>(g.batch) * g.out_h * g.out_w * g.out_c;
  const long long blocks = (total + 255) / 256;
  if (blocks > 0x7FFFFFFFll) return DODT_ECAPACITY;
  if (channels % 4 == 0 && aligned)
    corr_generic<4><<<static_cast<unsigned>(blocks), 256, 0, stream>>>(a, b, g, total, out);
  else
    corr_generic<1><<<static_cast<unsigned>(blocks), 256, 0, stream>>>(a, b, g, total, out);
  DODT_AFTER_LAUNCH();
  return DODT_OK;
}

int dodt_correlation_stream(const float *const *maps, int32_t n_maps, float *const *outs,
                            int32_t height, int32_t width, int32_t channels, int32_t kernel_size,
                            int32_t max_displacement, int32_t stride_1, int32_t stride_2,
                            int32_t pad, int32_t max_ctas, dodt_stream_t stream_) {
  using namespace dodt;
  if (!maps || !outs || n_maps < 2 || max_ctas < 0) return DODT_EINVAL;
  for (int k = 0; k < n_maps; ++k)
    if (!maps[k] || (k + 1 < n_maps && !outs[k])) return DODT_EINVAL;
  CorrGeom g;
  const int rc = fill_geom(1, height, width, channels, kernel_size, max_displacement, stride_1,
                           stride_2, pad, &g);
  if (rc != DODT_OK) return rc;
  cudaStream_t stream = as_stream(stream_);
  int j = 0;
  if (g.ks == 1 && g.s1 == 1 && g.s2 == 2) {
    // groups of up to DODT_CORR_STREAM_MAX_PAIRS pairs per launch; neighbouring groups share a map
    while (j < n_maps - 1) {
      const int n = n_maps - 1 - j < DODT_CORR_STREAM_MAX_PAIRS ? n_maps - 1 - j : DODT_CORR_STREAM_MAX_PAIRS;
      int done = 1;
      if (use_feed())
        done = correlation_stream_feed(maps + j, n, outs + j, g.H, g.W, g.C, g.r, g.out_h, g.out_w,
                                       g.md - g.pad, max_ctas, stream);
      if (done > 0)
        done = correlation_stream_tma(maps + j, n, outs + j, g.H, g.W, g.C, g.r, g.out_h,
                                      g.out_w, g.md - g.pad, max_ctas, stream);
      if (done < 0) return done;
      if (done > 0) break;   // not applicable: pair by pair below
      j += n;
    }
  }
  for (; j < n_maps - 1; ++j) {
    const int rc2 = dodt_correlation_shared(maps[j], maps[j + 1], 1, height, width, channels, kernel_size,
                                            max_displacement, stride_1, stride_2, pad, outs[j], max_ctas,
                                            stream_);
    if (rc2 != DODT_OK) return rc2;
  }
  return DODT_OK;
}

}  // extern "C"
