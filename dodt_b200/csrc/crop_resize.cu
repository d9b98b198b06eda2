// S3 — tf.image.crop_and_resize (bilinear) on NHWC float32 feature maps, sm_100a.
//
// Semantics follow TensorFlow 1.3.0 core/kernels/crop_and_resize_op.cc (CropAndResize functor),
// the op the reference calls at avod/core/models/dt_rpn_model.py:418-428 (3x3 RPN crops of the
// 1-channel bottlenecks) and avod/core/models/dt_avod_model.py:253-273 (7x7 crops of the 32-ch
// BEV / image maps and of the 25-ch correlation map):
//   scale   = (y2 - y1) * (H - 1) / (crop_h - 1)            (0 when crop_h == 1)
//   in_y    = y1 * (H - 1) + y * scale                      (0.5 * (y1 + y2) * (H - 1) if crop_h == 1)
//   sample  = extrapolation_value  if in_y < 0 or in_y > H - 1   (same test on x)
//   top     = tl + (tr - tl) * x_lerp ; bottom = bl + (br - bl) * x_lerp
//   out     = top + (bottom - top) * y_lerp                 (floorf / ceilf tap indices)
// every operation individually rounded in fp32 (no FMA contraction), so tap selection and the
// extrapolation decision are identical to the CPU op.
//
// Layout/mapping (crop_tables_kernel). A CTA owns R consecutive ROIs of one feature map:
//   phase 1  the crop_h + crop_w axis coordinates of each ROI are evaluated ONCE (one thread per
//            (ROI, axis, i)): tap offsets i0*stride / i1*stride, the lerp weight and the
//            extrapolation flag go to a shared-memory table — the two fp32 divisions and the
//            floor/ceil of the TF formula are not repeated per sample and per channel;
//   phase 2  one thread per output float4 (C % 4 == 0) or float: two table reads, four adds, four
//            128-bit tap loads, the individually rounded bilinear blend, one store. Consecutive
//            lanes are consecutive channel vectors of one sample, so every tap and every output is
//            a full 128-byte line for the 32-channel maps. Index decoding uses multiply-high by
//            precomputed reciprocals (no integer division in the loop).
// Feature maps are far smaller than the 126 MB L2, so taps that neighbouring ROIs share are served
// from L2; no shared-memory staging of pixels is used because a 7x7 crop touches at most 196 of
// the several hundred pixels under its box (staging the box would read more lines than the gather).
// The element-per-thread kernel (crop_resize_kernel) remains for tensors beyond 2^31 elements.
#include <stdlib.h>

#include "common.cuh"

namespace dodt {
namespace {

struct CropGeom {
  int batch, H, W, C;
  int crop_h, crop_w;
  float extrap;
};

// returns false if the sample is extrapolated
__device__ __forceinline__ bool sample_coord(float lo, float hi, int size, int crop, int i,
                                             int *i0, int *i1, float *lerp) {
  const float sm1 = static_cast<float>(size - 1);
  float in;
  if (crop > 1) {
    const float scale = __fdiv_rn(__fmul_rn(__fsub_rn(hi, lo), sm1), static_cast<float>(crop - 1));
    in = __fadd_rn(__fmul_rn(lo, sm1), __fmul_rn(static_cast<float>(i), scale));
  } else {
    // 0.5 * (y1 + y2) * (H - 1): the float sum is promoted to double by the 0.5 literal
    const double mid = __dmul_rn(__dmul_rn(0.5, static_cast<double>(__fadd_rn(lo, hi))),
                                 static_cast<double>(size - 1));
    in = __double2float_rn(mid);
  }
  if (in < 0.0f || in > sm1) return false;
  const float f = floorf(in);
  *i0 = static_cast<int>(f);
  *i1 = static_cast<int>(ceilf(in));
  *lerp = __fsub_rn(in, f);
  return true;
}

__device__ __forceinline__ float bilerp(float tl, float tr, float bl, float br, float xl,
                                        float yl) {
  const float top = __fadd_rn(tl, __fmul_rn(__fsub_rn(tr, tl), xl));
  const float bot = __fadd_rn(bl, __fmul_rn(__fsub_rn(br, bl), xl));
  return __fadd_rn(top, __fmul_rn(__fsub_rn(bot, top), yl));
}

template <int VEC>
__global__ void __launch_bounds__(256)
crop_resize_kernel(const float *__restrict__ image, const float *__restrict__ boxes,
                   const int *__restrict__ box_ind, const int *__restrict__ n_dev, long long total,
                   CropGeom g, float *__restrict__ crops) {
  const long long t = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (t >= total) return;
  const int cv = g.C / VEC;
  const int c = static_cast<int>(t % cv) * VEC;
  long long r = t / cv;
  const int x = static_cast<int>(r % g.crop_w);
  r /= g.crop_w;
  const int y = static_cast<int>(r % g.crop_h);
  const long long b = r / g.crop_h;

  if (n_dev && b >= __ldg(n_dev)) return;   // box count produced on the device
  const int b_in = box_ind ? __ldg(box_ind + b) : 0;
  if (b_in < 0 || b_in >= g.batch) return;  // TF leaves such rows untouched
  const float4 box = __ldg(reinterpret_cast<const float4 *>(boxes) + b);  // y1, x1, y2, x2

  float *dst = crops + t * VEC;
  int y0, y1i, x0, x1i;
  float yl, xl;
  const bool in_y = sample_coord(box.x, box.z, g.H, g.crop_h, y, &y0, &y1i, &yl);
  const bool in_x = sample_coord(box.y, box.w, g.W, g.crop_w, x, &x0, &x1i, &xl);
  if (!(in_y && in_x)) {
    if (VEC == 4) {
      *reinterpret_cast<float4 *>(dst) = make_float4(g.extrap, g.extrap, g.extrap, g.extrap);
    } else {
      *dst = g.extrap;
    }
    return;
  }
  const float *img = image + static_cast<size_t>(b_in) * g.H * g.W * g.C + c;
  const float *ptl = img + (static_cast<size_t>(y0) * g.W + x0) * g.C;
  const float *ptr = img + (static_cast<size_t>(y0) * g.W + x1i) * g.C;
  const float *pbl = img + (static_cast<size_t>(y1i) * g.W + x0) * g.C;
  const float *pbr = img + (static_cast<size_t>(y1i) * g.W + x1i) * g.C;
  if (VEC == 4) {
    const float4 tl = __ldg(reinterpret_cast<const float4 *>(ptl));
    const float4 tr = __ldg(reinterpret_cast<const float4 *>(ptr));
    const float4 bl = __ldg(reinterpret_cast<const float4 *>(pbl));
    const float4 br = __ldg(reinterpret_cast<const float4 *>(pbr));
    float4 o;
    o.x = bilerp(tl.x, tr.x, bl.x, br.x, xl, yl);
    o.y = bilerp(tl.y, tr.y, bl.y, br.y, xl, yl);
    o.z = bilerp(tl.z, tr.z, bl.z, br.z, xl, yl);
    o.w = bilerp(tl.w, tr.w, bl.w, br.w, xl, yl);
    *reinterpret_cast<float4 *>(dst) = o;
  } else {
    *dst = bilerp(__ldg(ptl), __ldg(ptr), __ldg(pbl), __ldg(pbr), xl, yl);
  }
}

constexpr int kCropThreads = 256;
constexpr int kCropTableMax = 1024;   // axis-table entries per CTA (16 KB; 2048 entries measured slower: 10.3 k frames/s)

struct CropMulti {
  const float *image[DODT_MAX_CROP_MAPS];
  const float *boxes[DODT_MAX_CROP_MAPS];
  float *crops[DODT_MAX_CROP_MAPS];
  int H[DODT_MAX_CROP_MAPS], W[DODT_MAX_CROP_MAPS], C[DODT_MAX_CROP_MAPS], vec[DODT_MAX_CROP_MAPS];
  int rois_per_cta[DODT_MAX_CROP_MAPS];
  unsigned mul_cv[DODT_MAX_CROP_MAPS];   // ceil(2^32 / (C / vec))
  unsigned mul_s, mul_cw;                // ceil(2^32 / (crop_h*crop_w)), ceil(2^32 / crop_w)
  int batch, crop_h, crop_w;
  float extrap;
};

// n / d for n < 65536 through the precomputed reciprocal (exact: n * (mul*d - 2^32) < 2^32)
// mul == 0 encodes d == 1 (2^32 does not fit)
__device__ __forceinline__ int fast_div(int n, unsigned mul) {
  return mul ? static_cast<int>(__umulhi(static_cast<unsigned>(n), mul)) : n;
}

template <int VEC>
__device__ __forceinline__ void crop_tables_body(const CropMulti &m, int s,
                                                 const int *__restrict__ box_ind, int n_eff,
                                                 int4 *tab) {
  const int C = m.C[s], H = m.H[s], W = m.W[s];
  const int ch = m.crop_h, cw = m.crop_w;
  const int cv = C / VEC, S = ch * cw, axes = ch + cw;
  const int R = m.rois_per_cta[s];
  const int roi0 = blockIdx.x * R;
  if (roi0 >= n_eff) return;
  const int nr = min(R, n_eff - roi0);
  const float4 *boxes = reinterpret_cast<const float4 *>(m.boxes[s]);

  // ---- phase 1: axis tables, one thread per (ROI, axis): the scale (one fp32 division) is
  // computed once and the crop positions of the axis are walked in a loop
  for (int t = threadIdx.x; t < nr * 2; t += kCropThreads) {
    const int rl = t >> 1;
    const bool is_y = (t & 1) == 0;
    const float4 box = __ldg(boxes + roi0 + rl);  // y1, x1, y2, x2
    const float lo = is_y ? box.x : box.y, hi = is_y ? box.z : box.w;
    const int size = is_y ? H : W, crop = is_y ? ch : cw;
    const int stride = is_y ? W * C : C;
    int base = 0, bad = 0;
    if (is_y) {
      const int b_in = box_ind ? __ldg(box_ind + roi0 + rl) : 0;
      if (b_in < 0 || b_in >= m.batch) bad = 1;   // TF leaves such rows untouched
      else base = b_in * H * W * C;
    }
    int4 *row = tab + rl * axes + (is_y ? 0 : ch);
    const float sm1 = static_cast<float>(size - 1);
    if (crop > 1) {
      const float scale = __fdiv_rn(__fmul_rn(__fsub_rn(hi, lo), sm1), static_cast<float>(crop - 1));
      const float start = __fmul_rn(lo, sm1);
      for (int i = 0; i < crop; ++i) {
        const float in = __fadd_rn(start, __fmul_rn(static_cast<float>(i), scale));
        const bool ok = !(in < 0.0f || in > sm1);
        const float f = floorf(in);
        const int i0 = ok ? static_cast<int>(f) : 0, i1 = ok ? static_cast<int>(ceilf(in)) : 0;
        row[i] = make_int4(base + i0 * stride, base + i1 * stride, __float_as_int(__fsub_rn(in, f)),
                           bad ? -1 : (ok ? 1 : 0));
      }
    } else {
      int i0 = 0, i1 = 0;
      float lerp = 0.0f;
      const bool ok = sample_coord(lo, hi, size, 1, 0, &i0, &i1, &lerp);
      row[0] = make_int4(base + i0 * stride, base + i1 * stride, __float_as_int(lerp), bad ? -1 : (ok ? 1 : 0));
    }
  }
  __syncthreads();

  // ---- phase 2: one thread per output vector
  const float *img = m.image[s];
  float *out = m.crops[s] + static_cast<size_t>(roi0) * S * C;
  const int items = nr * S * cv;
  const unsigned mul_cv = m.mul_cv[s];
  for (int item = threadIdx.x; item < items; item += kCropThreads) {
    const int slot = cv == 1 ? item : fast_div(item, mul_cv);
    const int c = (item - slot * cv) * VEC;
    const int rl = fast_div(slot, m.mul_s);
    const int sidx = slot - rl * S;
    const int y = fast_div(sidx, m.mul_cw);
    const int x = sidx - y * cw;
    const int4 ey = tab[rl * axes + y];
    const int4 ex = tab[rl * axes + ch + x];
    if (ey.w < 0) continue;
    float *dst = out + slot * C + c;
    const bool inside = ey.w > 0 && ex.w > 0;
    const float yl = __int_as_float(ey.z), xl = __int_as_float(ex.z);
    if (VEC == 4) {
      float4 o = make_float4(m.extrap, m.extrap, m.extrap, m.extrap);
      if (inside) {
        const float *p = img + c;
        const float4 tl = __ldg(reinterpret_cast<const float4 *>(p + ey.x + ex.x));
        const float4 tr = __ldg(reinterpret_cast<const float4 *>(p + ey.x + ex.y));
        const float4 bl = __ldg(reinterpret_cast<const float4 *>(p + ey.y + ex.x));
        const float4 br = __ldg(reinterpret_cast<const float4 *>(p + ey.y + ex.y));
        o.x = bilerp(tl.x, tr.x, bl.x, br.x, xl, yl);
        o.y = bilerp(tl.y, tr.y, bl.y, br.y, xl, yl);
        o.z = bilerp(tl.z, tr.z, bl.z, br.z, xl, yl);
        o.w = bilerp(tl.w, tr.w, bl.w, br.w, xl, yl);
      }
      *reinterpret_cast<float4 *>(dst) = o;
    } else {
      float o = m.extrap;
      if (inside) {
        const float *p = img + c;
        o = bilerp(__ldg(p + ey.x + ex.x), __ldg(p + ey.x + ex.y), __ldg(p + ey.y + ex.x),
                   __ldg(p + ey.y + ex.y), xl, yl);
      }
      *dst = o;
    }
  }
}

// blockIdx.y selects the map; every map uses its own channel vector width and ROIs per CTA
__global__ void __launch_bounds__(kCropThreads)
crop_tables_kernel(const CropMulti m, const int *__restrict__ box_ind,
                   const int *__restrict__ n_dev, int n) {
  __shared__ int4 tab[kCropTableMax];
  const int s = blockIdx.y;
  const int n_eff = n_dev ? min(n, __ldg(n_dev)) : n;
  if (m.vec[s] == 4) crop_tables_body<4>(m, s, box_ind, n_eff, tab);
  else crop_tables_body<1>(m, s, box_ind, n_eff, tab);
}

unsigned recip32(int d) {
  return d <= 1 ? 0u : static_cast<unsigned>(((1ull << 32) + d - 1) / static_cast<unsigned long long>(d));
}

// Fills the launch geometry of crop_tables_kernel; returns false if a map is outside its limits
// (more than 2^31 elements, or one ROI needs more than 65535 items / kCropTableMax table entries).
bool plan_tables(CropMulti *m, int n_specs, int64_t n, unsigned *grid_x) {
  const int S = m->crop_h * m->crop_w, axes = m->crop_h + m->crop_w;
  if (axes > kCropTableMax || n > 0x7FFFFFFF) return false;
  m->mul_s = recip32(S);
  m->mul_cw = recip32(m->crop_w);
  long long gx = 0;
  for (int k = 0; k < n_specs; ++k) {
    const long long elems = static_cast<long long>(m->batch) * m->H[k] * m->W[k] * m->C[k];
    const int cv = m->C[k] / m->vec[k];
    const long long per_roi = static_cast<long long>(S) * cv;
    if (elems >= 0x7FFFFFFFll || per_roi > 65535) return false;
    // ~4096 items (16 per thread) per CTA: measured in the frame pipeline, 512 / 1024 / 2048 / 4096 /
    // 8192 items give 10.50 / 10.62 / 10.70 / 10.88 / 10.80 k frames/s — fewer, longer CTAs cost the
    // co-running kernels less than many short ones, until too few CTAs are left to balance the SMs
    static int target = -1;   // DODT_CROP_ITEMS overrides it in the diagnostic build only
    if (target < 0) target = DODT_KNOB("DODT_CROP_ITEMS", 16 * kCropThreads);
    int R = static_cast<int>((target + per_roi - 1) / per_roi);
    if (R < 1) R = 1;
    if (R > kCropTableMax / axes) R = kCropTableMax / axes;
    while (R > 1 && R * per_roi > 65535) --R;
    m->rois_per_cta[k] = R;
    m->mul_cv[k] = recip32(cv);
    const long long g = (n + R - 1) / R;
    if (g > gx) gx = g;
  }
  for (int k = n_specs; k < DODT_MAX_CROP_MAPS; ++k) { m->rois_per_cta[k] = 1; m->mul_cv[k] = 0; }
  if (gx > 0x7FFFFFFFll) return false;
  *grid_x = static_cast<unsigned>(gx);
  return true;
}

}  // namespace
}  // namespace dodt

extern "C" int dodt_crop_and_resize_multi(const dodt_crop_spec *specs, int32_t n_specs,
                                          int32_t batch, const int32_t *box_ind, int64_t n,
                                          const int32_t *n_dev, int32_t crop_h, int32_t crop_w,
                                          float extrapolation_value, dodt_stream_t stream_) {
  using namespace dodt;
  if (!specs || n_specs <= 0 || n_specs > DODT_MAX_CROP_MAPS || n < 0 || batch <= 0 || crop_h <= 0 ||
      crop_w <= 0)
    return DODT_EINVAL;
  if (n == 0) return DODT_OK;
  CropMulti m;
  m.batch = batch; m.crop_h = crop_h; m.crop_w = crop_w; m.extrap = extrapolation_value;
  for (int k = 0; k < DODT_MAX_CROP_MAPS; ++k) {
    const int j = k < n_specs ? k : 0;
    const dodt_crop_spec &sp = specs[j];
    if (!sp.image || !sp.boxes || !sp.crops || sp.height <= 0 || sp.width <= 0 || sp.channels <= 0)
      return DODT_EINVAL;
    if (reinterpret_cast<uintptr_t>(sp.boxes) % 16 != 0) return DODT_EALIGN;
    m.image[k] = sp.image; m.boxes[k] = sp.boxes; m.crops[k] = sp.crops;
    m.H[k] = sp.height; m.W[k] = sp.width; m.C[k] = sp.channels;
    m.vec[k] = (sp.channels % 4 == 0 && reinterpret_cast<uintptr_t>(sp.image) % 16 == 0 &&
                reinterpret_cast<uintptr_t>(sp.crops) % 16 == 0) ? 4 : 1;
  }
  cudaStream_t stream = as_stream(stream_);
  unsigned grid_x = 0;
  if (plan_tables(&m, n_specs, n, &grid_x)) {
    dim3 grid(grid_x, n_specs);
    crop_tables_kernel<<<grid, kCropThreads, 0, stream>>>(m, box_ind, n_dev, static_cast<int>(n));
    DODT_AFTER_LAUNCH();
    return DODT_OK;
  }
  // outside the table kernel's limits: one element-per-thread launch per map
  for (int k = 0; k < n_specs; ++k) {
    CropGeom g{batch, m.H[k], m.W[k], m.C[k], crop_h, crop_w, extrapolation_value};
    const long long total = static_cast<long long>(n) * crop_h * crop_w * (m.C[k] / m.vec[k]);
    const long long blocks = (total + 255) / 256;
    if (blocks > 0x7FFFFFFFll) return DODT_ECAPACITY;
    if (m.vec[k] == 4)
      crop_resize_kernel<4><<<static_cast<unsigned>(blocks), 256, 0, stream>>>(m.image[k], m.boxes[k], box_ind, n_dev, total, g, m.crops[k]);
    else
      crop_resize_kernel<1><<<static_cast<unsigned>(blocks), 256, 0, stream>>>(m.image[k], m.boxes[k], box_ind, n_dev, total, g, m.crops[k]);
    DODT_AFTER_LAUNCH();
  }
  return DODT_OK;
}

extern "C" int dodt_crop_and_resize(const float *image, int32_t batch, int32_t height,
                                    int32_t width, int32_t channels, const float *boxes,
                                    const int32_t *box_ind, int64_t n, const int32_t *n_dev,
                                    int32_t crop_h, int32_t crop_w, float extrapolation_value,
                                    float *crops, dodt_stream_t stream_) {
  if (n < 0 || batch <= 0 || height <= 0 || width <= 0 || channels <= 0 || crop_h <= 0 ||
      crop_w <= 0)
    return DODT_EINVAL;
  if (n == 0) return DODT_OK;
  if (!image || !boxes || !crops) return DODT_EINVAL;
  dodt_crop_spec sp;
  sp.image = image; sp.boxes = boxes; sp.crops = crops;
  sp.height = height; sp.width = width; sp.channels = channels; sp.reserved = 0;
  return dodt_crop_and_resize_multi(&sp, 1, batch, box_ind, n, n_dev, crop_h, crop_w,
                                    extrapolation_value, stream_);
}
