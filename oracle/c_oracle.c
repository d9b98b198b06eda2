/*
 * CPU oracle in plain C — TEST INFRASTRUCTURE, NOT PRODUCT CODE (see oracle/np_oracle.py header).
 *
 * Restates, loop for loop, the three stages whose reference arithmetic is compiled code:
 *   oracle_correlation      avod/core/ops/correlation/correlation_kernel.cu.cc:21-119 on inputs
 *                           padded as pad.cu.cc:14-74 (GPU-only TF op in the reference)
 *   oracle_correlation_grad avod/core/ops/correlation/correlation_grad_kernel.cu.cc:20-189
 *                           (CorrelateDataBackward0 / CorrelateDataBackward1), loop for loop,
 *                           with the multiply-add fused as nvcc's default -fmad=true compiles it
 *   oracle_crop_and_resize  TensorFlow 1.3.0 core/kernels/crop_and_resize_op.cc (CPU functor)
 *   oracle_nms              TensorFlow 1.3.0 core/kernels/non_max_suppression_op.cc
 * Built by oracle/build_oracle.py with gcc -O2 -ffp-contract=off (every fp32 operation rounded
 * on its own, like the scalar C++ of those ops). It is checked against oracle/np_oracle.py in
 * tests/ and serves as the CPU baseline of bench.py (single thread per call; bench.py spreads
 * frames over processes).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

int oracle_correlation(const float *a, const float *b, int N, int H, int W, int C, int ks, int md,
                       int s1, int s2, int pad, float *out) {
  if (ks % 2 == 0) return -1;
  const int kr = (ks - 1) / 2, border = md + kr;
  const int oh = (int)ceilf((float)(H + 2 * pad - 2 * border) / (float)s1);
  const int ow = (int)ceilf((float)(W + 2 * pad - 2 * border) / (float)s1);
  if (oh < 1 || ow < 1) return -2;
  const int r = md / s2, wn = 2 * r + 1, oc = wn * wn;
  const float sumelems = (float)(ks * ks * C);
  for (int n = 0; n < N; ++n)
    for (int y = 0; y < oh; ++y)
      for (int x = 0; x < ow; ++x) {
        /* patch corner in padded coordinates (correlation_kernel.cu.cc:45-46) -> unpadded */
        const int y1 = y * s1 + md - pad, x1 = x * s1 + md - pad;
        for (int k = 0; k < oc; ++k) {
          const int s2o = (k % wn - r) * s2, s2p = (k / wn - r) * s2;
          float lane[32];
          for (int t = 0; t < 32; ++t) lane[t] = 0.0f;
          for (int j = 0; j < ks; ++j)
            for (int i = 0; i < ks; ++i) {
              const int ay = y1 + j, ax = x1 + i, by = ay + s2p, bx = ax + s2o;
              if (ay < 0 || ay >= H || ax < 0 || ax >= W) continue; /* zero padding of A */
              if (by < 0 || by >= H || bx < 0 || bx >= W) continue; /* zero padding of B */
              const float *pa = a + (((size_t)n * H + ay) * W + ax) * C;
              const float *pb = b + (((size_t)n * H + by) * W + bx) * C;
              /* sum[ch_off] += patch * b (correlation_kernel.cu.cc:93) is one FFMA under nvcc's default -fmad=true */
              for (int ch = 0; ch < C; ++ch) lane[ch & 31] = fmaf(pa[ch], pb[ch], lane[ch & 31]);
            }
          float total = 0.0f;
          for (int t = 0; t < 32; ++t) total += lane[t];
          out[(((size_t)n * oh + y) * ow + x) * oc + k] = total / sumelems;
        }
      }
  return 0;
}

/* correlation_grad_kernel.cu.cc:47-62: the ROUND_OFF trick is ceil / floor of a signed quotient */
static int grad_lo(int v, int s1) { return (v + 50000 * s1 - 1) / s1 + 1 - 50000; }
static int grad_hi(int v, int s1) { return (v + 50000 * s1) / s1 - 50000; }

/* which = 0: d/d input_a (other = input_b, CorrelateDataBackward0, :20-107);
 * which = 1: d/d input_b (other = input_a, CorrelateDataBackward1, :109-189).
 * Padded temporaries are not built: a tap outside the image is the zero PadData wrote (or, outside
 * the padded image, the value this restatement DEFINES as zero; the reference reads out of bounds). */
static void grad_one(int which, const float *grad, const float *other, int N, int H, int W, int C,
                     int ks, int md, int s1, int s2, int pad, int oh, int ow, float *dst) {
  const int kr = (ks - 1) / 2, r = md / s2, wn = 2 * r + 1, oc = wn * wn;
  const float sumelems = (float)(ks * ks * C);
  for (int n = 0; n < N; ++n)
    for (int yy = 0; yy < H; ++yy)
      for (int xx = 0; xx < W; ++xx)
        for (int k = 0; k < C; ++k) {
          const int x = xx + pad, y = yy + pad;
          float sum = 0.0f;
          for (int p = -r; p <= r; ++p)
            for (int o = -r; o <= r; ++o) {
              const int s2o = s2 * o, s2p = s2 * p;
              const int wx = which ? x - s2o : x, wy = which ? y - s2p : y;
              int xmin = grad_lo(wx - 2 * kr - md, s1), ymin = grad_lo(wy - 2 * kr - md, s1);
              int xmax = grad_hi(wx - md, s1), ymax = grad_hi(wy - md, s1);
              if (!(xmax >= 0 && ymax >= 0 && xmin <= ow - 1 && ymin <= oh - 1)) continue;
              if (xmin < 0) xmin = 0;
              if (xmax > ow - 1) xmax = ow - 1;
              if (ymin < 0) ymin = 0;
              if (ymax > oh - 1) ymax = oh - 1;
              const int ty = (which ? y - s2p : y + s2p) - pad, tx = (which ? x - s2o : x + s2o) - pad;
              float v = 0.0f;
              if (ty >= 0 && ty < H && tx >= 0 && tx < W) v = other[(((size_t)n * H + ty) * W + tx) * C + k];
              const int op = (p + r) * wn + (o + r);
              for (int gy = ymin; gy <= ymax; ++gy)
                for (int gx = xmin; gx <= xmax; ++gx)
                  sum = fmaf(grad[(((size_t)n * oh + gy) * ow + gx) * oc + op], v, sum);
            }
          dst[(((size_t)n * H + yy) * W + xx) * C + k] = sum / sumelems;
        }
}

int oracle_correlation_grad(const float *grad, const float *a, const float *b, int N, int H, int W,
                            int C, int ks, int md, int s1, int s2, int pad, float *ga, float *gb) {
  if (ks % 2 == 0) return -1;
  const int kr = (ks - 1) / 2, border = md + kr;
  const int oh = (int)ceilf((float)(H + 2 * pad - 2 * border) / (float)s1);
  const int ow = (int)ceilf((float)(W + 2 * pad - 2 * border) / (float)s1);
  if (oh < 1 || ow < 1) return -2;
  if (ga) grad_one(0, grad, b, N, H, W, C, ks, md, s1, s2, pad, oh, ow, ga);
  if (gb) grad_one(1, grad, a, N, H, W, C, ks, md, s1, s2, pad, oh, ow, gb);
  return 0;
}

int oracle_crop_and_resize(const float *image, int B, int H, int W, int C, const float *boxes,
                           const int32_t *box_ind, int n, int ch, int cw, float extrap,
                           float *crops) {
  for (int b = 0; b < n; ++b) {
    const float y1 = boxes[b * 4 + 0], x1 = boxes[b * 4 + 1];
    const float y2 = boxes[b * 4 + 2], x2 = boxes[b * 4 + 3];
    const int b_in = box_ind ? box_ind[b] : 0;
    if (b_in < 0 || b_in >= B) continue;
    const float hs = (ch > 1) ? (y2 - y1) * (H - 1) / (ch - 1) : 0;
    const float ws = (cw > 1) ? (x2 - x1) * (W - 1) / (cw - 1) : 0;
    for (int y = 0; y < ch; ++y) {
      const float in_y = (ch > 1) ? y1 * (H - 1) + y * hs : 0.5 * (y1 + y2) * (H - 1);
      float *row = crops + (((size_t)b * ch + y) * cw) * C;
      if (in_y < 0 || in_y > H - 1) {
        for (int e = 0; e < cw * C; ++e) row[e] = extrap;
        continue;
      }
      const int top = (int)floorf(in_y), bot = (int)ceilf(in_y);
      const float yl = in_y - top;
      for (int x = 0; x < cw; ++x) {
        const float in_x = (cw > 1) ? x1 * (W - 1) + x * ws : 0.5 * (x1 + x2) * (W - 1);
        float *px = row + (size_t)x * C;
        if (in_x < 0 || in_x > W - 1) {
          for (int d = 0; d < C; ++d) px[d] = extrap;
          continue;
        }
        const int left = (int)floorf(in_x), right = (int)ceilf(in_x);
        const float xl = in_x - left;
        const float *base = image + (size_t)b_in * H * W * C;
        const float *ptl = base + ((size_t)top * W + left) * C, *ptr = base + ((size_t)top * W + right) * C;
        const float *pbl = base + ((size_t)bot * W + left) * C, *pbr = base + ((size_t)bot * W + right) * C;
        for (int d = 0; d < C; ++d) {
          const float t = ptl[d] + (ptr[d] - ptl[d]) * xl;
          const float bo = pbl[d] + (pbr[d] - pbl[d]) * xl;
          px[d] = t + (bo - t) * yl;
        }
      }
    }
  }
  return 0;
}

static const float *g_scores;
static int cmp_desc_stable(const void *pa, const void *pb) {
  const int ia = *(const int *)pa, ib = *(const int *)pb;
  if (g_scores[ia] > g_scores[ib]) return -1;
  if (g_scores[ia] < g_scores[ib]) return 1;
  return ia < ib ? -1 : (ia > ib ? 1 : 0);
}

static float iou(const float *bx, int i, int j) {
  const float ymin_i = fminf(bx[i * 4], bx[i * 4 + 2]), xmin_i = fminf(bx[i * 4 + 1], bx[i * 4 + 3]);
  const float ymax_i = fmaxf(bx[i * 4], bx[i * 4 + 2]), xmax_i = fmaxf(bx[i * 4 + 1], bx[i * 4 + 3]);
  const float ymin_j = fminf(bx[j * 4], bx[j * 4 + 2]), xmin_j = fminf(bx[j * 4 + 1], bx[j * 4 + 3]);
  const float ymax_j = fmaxf(bx[j * 4], bx[j * 4 + 2]), xmax_j = fmaxf(bx[j * 4 + 1], bx[j * 4 + 3]);
  const float area_i = (ymax_i - ymin_i) * (xmax_i - xmin_i);
  const float area_j = (ymax_j - ymin_j) * (xmax_j - xmin_j);
  if (area_i <= 0 || area_j <= 0) return 0.0f;
  const float iy0 = fmaxf(ymin_i, ymin_j), ix0 = fmaxf(xmin_i, xmin_j);
  const float iy1 = fminf(ymax_i, ymax_j), ix1 = fminf(xmax_i, xmax_j);
  const float inter = fmaxf(iy1 - iy0, 0.0f) * fmaxf(ix1 - ix0, 0.0f);
  return inter / (area_i + area_j - inter);
}

/* returns the number of selected indices written to `selected` (capacity max_out) */
int oracle_nms(const float *boxes, const float *scores, int n, int max_out, float thr,
               int32_t *selected) {
  const int out_size = max_out < n ? max_out : n;
  if (out_size <= 0) return 0;
  int *order = (int *)malloc(sizeof(int) * (size_t)n);
  if (!order) return -1;
  for (int i = 0; i < n; ++i) order[i] = i;
  g_scores = scores;
  qsort(order, (size_t)n, sizeof(int), cmp_desc_stable);
  int k = 0;
  for (int i = 0; i < n && k < out_size; ++i) {
    int keep = 1;
    for (int j = k - 1; j >= 0; --j) /* most recent first, as TF does */
      if (iou(boxes, order[i], selected[j]) > thr) { keep = 0; break; }
    if (keep) selected[k++] = order[i];
  }
  free(order);
  return k;
}
