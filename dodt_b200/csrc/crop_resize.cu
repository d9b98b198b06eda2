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
// Layout/mapping: one thread per output float4 (4 consecutive channels of one crop sample), so a
// 32-channel sample is written by 8 adjacent lanes as one 128-byte line and each of the four taps
// is read as one 128-byte line. Feature maps are far smaller than the 126 MB L2, so taps that
// neighbouring ROIs share are served from L2; no shared-memory staging is used because a 7x7 crop
// touches at most 196 of the several hundred pixels under its box (staging the box would read
// more lines than the gather does).
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

struct CropMulti {
  const float *image[DODT_MAX_CROP_MAPS];
  const float *boxes[DODT_MAX_CROP_MAPS];
  float *crops[DODT_MAX_CROP_MAPS];
  int H[DODT_MAX_CROP_MAPS], W[DODT_MAX_CROP_MAPS], C[DODT_MAX_CROP_MAPS], vec[DODT_MAX_CROP_MAPS];
  int batch, crop_h, crop_w;
  float extrap;
};

// blockIdx.y selects the map; every map uses its own channel vector width
__global__ void __launch_bounds__(256)
crop_resize_multi_kernel(const CropMulti m, const int *__restrict__ box_ind,
                         const int *__restrict__ n_dev, long long n) {
  const int s = blockIdx.y;
  const int vec = m.vec[s];
  const int C = m.C[s], H = m.H[s], W = m.W[s];
  const int cv = C / vec;
  const long long total = n * m.crop_h * m.crop_w * cv;
  const long long t = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (t >= total) return;
  const int c = static_cast<int>(t % cv) * vec;
  long long r = t / cv;
  const int x = static_cast<int>(r % m.crop_w);
  r /= m.crop_w;
  const int y = static_cast<int>(r % m.crop_h);
  const long long b = r / m.crop_h;
  if (n_dev && b >= __ldg(n_dev)) return;
  const int b_in = box_ind ? __ldg(box_ind + b) : 0;
  if (b_in < 0 || b_in >= m.batch) return;
  const float4 box = __ldg(reinterpret_cast<const float4 *>(m.boxes[s]) + b);
  float *dst = m.crops[s] + t * vec;
  int y0, y1i, x0, x1i;
  float yl, xl;
  const bool in_y = sample_coord(box.x, box.z, H, m.crop_h, y, &y0, &y1i, &yl);
  const bool in_x = sample_coord(box.y, box.w, W, m.crop_w, x, &x0, &x1i, &xl);
  const float *img = m.image[s] + static_cast<size_t>(b_in) * H * W * C + c;
  if (vec == 4) {
    float4 o = make_float4(m.extrap, m.extrap, m.extrap, m.extrap);
    if (in_y && in_x) {
      const float4 tl = __ldg(reinterpret_cast<const float4 *>(img + (static_cast<size_t>(y0) * W + x0) * C));
      const float4 tr = __ldg(reinterpret_cast<const float4 *>(img + (static_cast<size_t>(y0) * W + x1i) * C));
      const float4 bl = __ldg(reinterpret_cast<const float4 *>(img + (static_cast<size_t>(y1i) * W + x0) * C));
      const float4 br = __ldg(reinterpret_cast<const float4 *>(img + (static_cast<size_t>(y1i) * W + x1i) * C));
      o.x = bilerp(tl.x, tr.x, bl.x, br.x, xl, yl);
      o.y = bilerp(tl.y, tr.y, bl.y, br.y, xl, yl);
      o.z = bilerp(tl.z, tr.z, bl.z, br.z, xl, yl);
      o.w = bilerp(tl.w, tr.w, bl.w, br.w, xl, yl);
    }
    *reinterpret_cast<float4 *>(dst) = o;
  } else {
    float o = m.extrap;
    if (in_y && in_x)
      o = bilerp(__ldg(img + (static_cast<size_t>(y0) * W + x0) * C),
                 __ldg(img + (static_cast<size_t>(y0) * W + x1i) * C),
                 __ldg(img + (static_cast<size_t>(y1i) * W + x0) * C),
                 __ldg(img + (static_cast<size_t>(y1i) * W + x1i) * C), xl, yl);
    *dst = o;
  }
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
  long long max_threads = 0;
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
    const long long th = static_cast<long long>(n) * crop_h * crop_w * (sp.channels / m.vec[k]);
    if (k < n_specs && th > max_threads) max_threads = th;
  }
  const long long blocks = (max_threads + 255) / 256;
  if (blocks > 0x7FFFFFFFll) return DODT_ECAPACITY;
  dim3 grid(static_cast<unsigned>(blocks), n_specs);
  crop_resize_multi_kernel<<<grid, 256, 0, as_stream(stream_)>>>(m, box_ind, n_dev, n);
  DODT_AFTER_LAUNCH();
  return DODT_OK;
}

extern "C" int dodt_crop_and_resize(const float *image, int32_t batch, int32_t height,
                                    int32_t width, int32_t channels, const float *boxes,
                                    const int32_t *box_ind, int64_t n, const int32_t *n_dev,
                                    int32_t crop_h, int32_t crop_w, float extrapolation_value,
                                    float *crops, dodt_stream_t stream_) {
  using namespace dodt;
  if (n < 0 || batch <= 0 || height <= 0 || width <= 0 || channels <= 0 || crop_h <= 0 ||
      crop_w <= 0)
    return DODT_EINVAL;
  if (n == 0) return DODT_OK;
  if (!image || !boxes || !crops) return DODT_EINVAL;
  if (reinterpret_cast<uintptr_t>(boxes) % 16 != 0) return DODT_EALIGN;
  cudaStream_t stream = as_stream(stream_);
  CropGeom g{batch, height, width, channels, crop_h, crop_w, extrapolation_value};
  const bool vec4 = channels % 4 == 0 && reinterpret_cast<uintptr_t>(image) % 16 == 0 &&
                    reinterpret_cast<uintptr_t>(crops) % 16 == 0;
  const long long total = static_cast<long long>(n) * crop_h * crop_w * (vec4 ? channels / 4 : channels);
  const long long blocks = (total + 255) / 256;
  if (blocks > 0x7FFFFFFFll) return DODT_ECAPACITY;
  if (vec4)
    crop_resize_kernel<4><<<static_cast<unsigned>(blocks), 256, 0, stream>>>(image, boxes, box_ind, n_dev, total, g, crops);
  else
    crop_resize_kernel<1><<<static_cast<unsigned>(blocks), 256, 0, stream>>>(image, boxes, box_ind, n_dev, total, g, crops);
  DODT_AFTER_LAUNCH();
  return DODT_OK;
}
