// S4 backward — gradients of the FlowNet-style correlation with respect to
// both NHWC float32 inputs, sm_100a.
//
// Replaces the reference's TensorFlow custom op "CorrelationGrad" (paths relative to the
// Guoxs/DODT checkout):
//   avod/core/corr_layers/correlation.py:30-48                 @tf.RegisterGradient("Correlation")
//   avod/core/ops/correlation/correlation_grad_kernel.cc:28-151 shape math, two padded temporaries
//   avod/core/ops/correlation/correlation_grad_kernel.cu.cc:20-107  CorrelateDataBackward0 (d/dA)
//   avod/core/ops/correlation/correlation_grad_kernel.cu.cc:109-189 CorrelateDataBackward1 (d/dB)
//
// With kr = (ks-1)/2, Y = y + pad, X = x + pad (padded coordinates of input pixel (y, x)),
// s2p = p*s2, s2o = o*s2, k = (p+r)*Wn + (o+r):
//   gA[n,y,x,c] = 1/(ks^2 C) * sum_{p,o} Bpad[n, Y+s2p, X+s2o, c] * sum_{(oy,ox) in win(Y,X)}     G[n,oy,ox,k]
//   gB[n,y,x,c] = 1/(ks^2 C) * sum_{p,o} Apad[n, Y-s2p, X-s2o, c] * sum_{(oy,ox) in win(Y-s2p,X-s2o)} G[n,oy,ox,k]
//   win(Y,X) = output pixels whose ks x ks patch covers (Y,X):
//              ceil((Y - 2kr - md)/s1) <= oy <= floor((Y - md)/s1), clamped to the output.
// The padded temporaries are never materialised: taps outside the image are zero. (The reference
// reads its padded copies without bounds checks, so parameters that push a displaced tap outside
// the PADDED image are undefined there; here those taps are zero as well.)
//
// Kernels
//   corr_grad_generic<WHICH>  any parameters: one thread per (pixel, channel), the reference's
//                             loops in the reference's order.
//   corr_grad_k1<R, REV, GB>  the DODT family (kernel_size 1, stride_1 1, stride_2 2, C % 8 == 0,
//                             r in {1,2}). The reduction runs over displacements, not channels, so
//                             the structure mirrors the forward kernel with the roles swapped: a
//                             thread owns 4 pixels spaced 2 apart on one row and keeps their
//                             4 x (2r+1)^2 gradient coefficients in REGISTERS for the whole tile
//                             (staged once through shared memory so that the global reads are
//                             coalesced), streams 8-channel chunks of the other input (tile + halo)
//                             through shared memory, and writes 4 x 8 finished gradients per chunk.
//                             Every value of the other input is read from HBM once per tile
//                             (halo re-reads hit L2) instead of (2r+1)^2 times.
//   corr_grad_k1<.., FLIP>    gB as the same contraction over the displacement-flipped gradient
//                             Gf[n,y,x,k'] = G[n, y+2p'-shift, x+2o'-shift, D2-1-k'], walked backwards
//                             (REV), which is the reference's summation order for gB, so both
//                             gradients are bit-identical to a scalar restatement that uses fmaf.
//                             Gf is never materialised (round 1 and 2 wrote it to a workspace with a
//                             separate pass, corr_grad_flip: 28.5 us and 196 MB of HBM traffic per call):
//                             a tile gathers its 4 x D2 coefficients per thread from a shared-memory
//                             copy of the gradient's (8 + 4r) x (64 + 4r) pixel neighbourhood, staged
//                             eight rows at a time over the two chunk stages.
//
// Algorithmic HBM bytes (both gradients, 700x800x32, 25 displacements): read G 56 MB + A, B 2 x 71.68
// MB, write gA, gB 2 x 71.68 MB = 342.7 MB (the neighbourhood rows of gB's gather are re-read from L2).
#include "common.cuh"
#include "tma_util.cuh"

namespace dodt {
namespace {

struct GradGeom {
  int batch, H, W, C;
  int ks, kr, md, s1, s2, pad;
  int out_h, out_w, out_c;
  int r, wn;
};

// floor / ceil of a / b for b > 0 and any sign of a (the reference's ROUND_OFF trick,
// correlation_grad_kernel.cu.cc:47-62, computes exactly these)
__device__ __forceinline__ int floor_div(int a, int b) {
  const int q = a / b;
  return (a % b != 0 && (a < 0)) ? q - 1 : q;
}
__device__ __forceinline__ int ceil_div_s(int a, int b) { return -floor_div(-a, b); }

// WHICH = 0: gradient w.r.t. input_a (other = input_b); 1: w.r.t. input_b (other = input_a)
template <int WHICH>
__global__ void __launch_bounds__(256)
corr_grad_generic(const float *__restrict__ grad, const float *__restrict__ other, GradGeom g,
                  long long total, float *__restrict__ dst) {
  const long long t = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (t >= total) return;
  const int c = static_cast<int>(t % g.C);
  long long rr = t / g.C;
  const int x = static_cast<int>(rr % g.W);
  rr /= g.W;
  const int y = static_cast<int>(rr % g.H);
  const int n = static_cast<int>(rr / g.H);
  const int X = x + g.pad, Y = y + g.pad;
  const float *gn = grad + static_cast<size_t>(n) * g.out_h * g.out_w * g.out_c;
  const float *on = other + static_cast<size_t>(n) * g.H * g.W * g.C;

  float sum = 0.0f;
  for (int p = -g.r; p <= g.r; ++p)
    for (int o = -g.r; o <= g.r; ++o) {
      const int s2o = g.s2 * o, s2p = g.s2 * p;
      // WHICH 0: window of (Y, X), tap Bpad[Y+s2p, X+s2o]; WHICH 1: window and tap at (Y-s2p, X-s2o)
      const int wy = WHICH == 0 ? Y : Y - s2p, wx = WHICH == 0 ? X : X - s2o;
      const int ty = WHICH == 0 ? Y + s2p : Y - s2p, tx = WHICH == 0 ? X + s2o : X - s2o;
      int xmin = ceil_div_s(wx - 2 * g.kr - g.md, g.s1), xmax = floor_div(wx - g.md, g.s1);
      int ymin = ceil_div_s(wy - 2 * g.kr - g.md, g.s1), ymax = floor_div(wy - g.md, g.s1);
      if (!(xmax >= 0 && ymax >= 0 && xmin <= g.out_w - 1 && ymin <= g.out_h - 1)) continue;
      xmin = max(0, xmin); xmax = min(g.out_w - 1, xmax);
      ymin = max(0, ymin); ymax = min(g.out_h - 1, ymax);
      const int uy = ty - g.pad, ux = tx - g.pad;      // unpadded tap
      float v = 0.0f;
      if (uy >= 0 && uy < g.H && ux >= 0 && ux < g.W)
        v = __ldg(on + (static_cast<size_t>(uy) * g.W + ux) * g.C + c);
      const int op = (p + g.r) * g.wn + (o + g.r);
      for (int oy = ymin; oy <= ymax; ++oy)
        for (int ox = xmin; ox <= xmax; ++ox)
          sum = fmaf(__ldg(gn + (static_cast<size_t>(oy) * g.out_w + ox) * g.out_c + op), v, sum);
    }
  const float sumelems = static_cast<float>(g.ks * g.ks * g.C);
  dst[t] = __fdiv_rn(sum, sumelems);
}

// ---------------------------------------------------------------------------------------------
// DODT family: kernel_size 1, stride_1 1, stride_2 2.
// ---------------------------------------------------------------------------------------------
constexpr int kTW = 64, kTH = 8, kPX = 4, kCC = 8;
constexpr int kThreads = kTW / (2 * kPX) * 2 * kTH;   // 128

// 32-byte pixel vectors; the two 16-byte halves of pixel p swap when bit 2 of p is set (as in the
// forward kernel: eight lanes reading eight different pixels hit eight different bank groups)
__device__ __forceinline__ int smem_off(int pixel, int half) {
  return pixel * kCC + ((half ^ ((pixel >> 2) & 1)) << 2);
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
// LDGSTS with zero fill: `bytes` = BYTES copies, 0 writes zeros (src is not dereferenced)
template <int BYTES>
__device__ __forceinline__ void cp_async(uint32_t dst, const void *src, bool valid) {
  const int bytes = valid ? BYTES : 0;
  if constexpr (BYTES == 16)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
  else
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2, %3;" ::"r"(dst), "l"(src), "n"(BYTES), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int R>
struct GradCfg {
  static constexpr int WN = 2 * R + 1, D2 = WN * WN, HALO = 2 * R;
  static constexpr int BW = kTW + 2 * HALO;
  static constexpr int BPITCH = ((BW + 7) / 8) * 8 + 2;       // 2 (mod 8) pixels
  static constexpr int BH = kTH + 2 * HALO;
  static constexpr int G_PITCH = kTW * D2 + ((2 - kTW * D2 % 32) + 32) % 32;  // 2 (mod 32) floats
  static constexpr int G_FLOATS = kTH * G_PITCH;
  static constexpr int B_FLOATS = BH * BPITCH * kCC;          // one stage
  // the coefficient tile is only needed until it sits in registers: it aliases the two stages
  static constexpr int FLOATS = 2 * B_FLOATS > G_FLOATS ? 2 * B_FLOATS : G_FLOATS;
  // FLIP (gB straight from the gradient map): the coefficients of a tile come from the BH x BW pixel
  // neighbourhood of the gradient, staged eight rows at a time (two phases) over the two stages
  static constexpr int FW = BW * D2;                                    // floats per neighbourhood row
  static constexpr int F_PITCH = FW + ((2 - FW % 32) + 32) % 32;        // 2 (mod 32) floats
  static_assert(8 * F_PITCH <= FLOATS, "a flip phase fits the stage memory");
  static constexpr size_t BAR_OFF = ((static_cast<size_t>(FLOATS) * sizeof(float) + 127) / 128) * 128;
  static constexpr size_t SMEM = BAR_OFF + 64;                // + the two full barriers of the TMA feed
};

// coef [batch, ch, cw, D2]: ch x cw is the extent of the coefficient map; input pixel (y, x) uses
// coef[y - cshift, x - cshift] (zero outside). other/dst [batch, H, W, C].
//   dst[n,y,x,c] = 1/C * sum_{p,o} coef[n, y-cshift, x-cshift, k] * other[n, y+2p, x+2o, c]
// REV: walk k = D2-1 .. 0 instead of 0 .. D2-1.
// All staging is LDGSTS (cp.async with zero fill = the reference's padding): the copies of a phase
// are in flight together, and chunk c+1 of the other input lands while chunk c is consumed.
// GB = bytes per coefficient copy (8 when every tile row starts 8-byte aligned, else 4).
// TMA: the other input's tiles arrive by cp.async.bulk.tensor (UTMALDG) issued by one thread and
// completing on an mbarrier, zero fill by the TMA unit (round 2: the forward kernel's feed);
// otherwise by LDGSTS from every thread (round 1; kept for hosts without tensor maps).
// FLIP (with REV): coef is the gradient map itself and the coefficient of tap k = (p, o) of input pixel
// (y, x) is coef[y + 2p - cshift, x + 2o - cshift, D2-1-k] — what corr_grad_flip used to materialise
// for the whole map (a 56 MB read + 70 MB write + 70 MB re-read per call) is gathered per tile from a
// shared-memory copy of the tile's neighbourhood, eight rows per phase.
template <int R, bool REV, int GB, bool TMA, bool FLIP = false>
__global__ void __launch_bounds__(kThreads, 2)
corr_grad_k1(const float *__restrict__ coef, const float *__restrict__ other, const __grid_constant__ CUtensorMap map_other,
             int batch, int H, int W, int C, int ch, int cw, int cshift, int tiles_x, int tiles_y,
             float *__restrict__ dst) {
  using Cfg = GradCfg<R>;
  constexpr int WN = Cfg::WN, D2 = Cfg::D2, NB = kPX + 2 * R;
  constexpr int GE = GB / 4;            // floats per coefficient copy
  extern __shared__ __align__(1024) float smem[];
  const uint32_t bar_full = tma::smem_u32(reinterpret_cast<unsigned char *>(smem) + Cfg::BAR_OFF);
  if (TMA) {
    if (threadIdx.x == 0) {
      tma::mbar_init(bar_full, 1);
      tma::mbar_init(bar_full + 8, 1);
      tma::mbar_fence_init();
    }
    __syncthreads();
  }
  unsigned it = 0;                      // chunks consumed so far by this CTA (TMA: stage / phase)
  float *sg = smem;                     // [kTH][G_PITCH], dead once gk[] is loaded
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // lane bits: [0] parity, [1..2] row & 3, [3..4] group & 3; warps tile 2 (rows) x 2 (x halves)
  const int row = (warp >> 1) * 4 + ((lane >> 1) & 3);
  const int x0 = ((warp & 1) * 4 + (lane >> 3)) * (2 * kPX) + (lane & 1);
  const float sumelems = static_cast<float>(C);
  const int n_tiles = tiles_x * tiles_y * batch;
  const int n_chunks = C / kCC;

  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int tx0 = (tile % tiles_x) * kTW;
    const int ty0 = ((tile / tiles_x) % tiles_y) * kTH;
    const int n = tile / (tiles_x * tiles_y);
    const float *cn = coef + static_cast<size_t>(n) * ch * cw * D2;
    const float *on = other + static_cast<size_t>(n) * H * W * C;
    float *dn = dst + static_cast<size_t>(n) * H * W * C;

    __syncthreads();   // the previous tile's last chunk has been consumed
    float gk[kPX][D2];
    if constexpr (FLIP) {
      static_assert(REV, "the flipped coefficients belong to the reversed walk");
      constexpr int kPerRow = Cfg::FW / GE;                         // copies per neighbourhood row
      constexpr int kIter = (kPerRow + kThreads - 1) / kThreads;
      const int gx0 = tx0 - Cfg::HALO - cshift;
      const int f_lo = max(0, -gx0) * D2, f_hi = min(Cfg::BW, cw - gx0) * D2;
      const int f0 = static_cast<int>(threadIdx.x) * GE;
#pragma unroll
      for (int ph = 0; ph < 2; ++ph) {
#pragma unroll
        for (int r_ = 0; r_ < 8; ++r_) {
          const int nr = ph * 8 + r_;
          if (nr < Cfg::BH) {
            const int gy = ty0 + nr - Cfg::HALO - cshift;
            const bool row_ok = gy >= 0 && gy < ch;
            const long long off = (static_cast<long long>(gy) * cw + gx0) * D2 + f0;
            const float *src0 = row_ok ? cn + off : cn;             // only dereferenced when ok
            const uint32_t dst0 = smem_u32(sg + r_ * Cfg::F_PITCH + f0);
#pragma unroll
            for (int i = 0; i < kIter; ++i) {
              const int f = f0 + i * kThreads * GE;
              if (kPerRow % kThreads == 0 || f < Cfg::FW) {
                const bool ok = row_ok && f >= f_lo && f < f_hi;
                cp_async<GB>(dst0 + i * kThreads * GE * 4, ok ? src0 + i * kThreads * GE : cn, ok);
              }
            }
          }
        }
        cp_async_commit();
        cp_async_wait<0>();
        __syncthreads();
        // tap row pi of pixel row `row` lives in neighbourhood row row + 2 pi
#pragma unroll
        for (int pi = 0; pi < WN; ++pi) {
          const int nr = row + 2 * pi;
          if ((nr >> 3) == ph) {
            const float *srow = sg + (nr & 7) * Cfg::F_PITCH;
#pragma unroll
            for (int j = 0; j < kPX; ++j)
#pragma unroll
              for (int oi = 0; oi < WN; ++oi)
                gk[j][pi * WN + oi] = srow[(x0 + 2 * j + 2 * oi) * D2 + (D2 - 1 - (pi * WN + oi))];
          }
        }
        __syncthreads();   // the phase has been read: the next one (or the stages) may be filled
      }
    } else {
    // ---- coefficients of the tile: each row is a contiguous run of kTW * D2 floats
    // (a row's copies differ by compile-time offsets from one global and one shared base; the
    // columns inside the coefficient map are the float range [f_lo, f_hi) of the row)
    {
      constexpr int kPerRow = kTW * D2 / GE;                       // copies per tile row
      constexpr int kIter = (kPerRow + kThreads - 1) / kThreads;
      const int f_lo = max(0, cshift - tx0) * D2, f_hi = min(kTW, cw + cshift - tx0) * D2;
      const int f0 = static_cast<int>(threadIdx.x) * GE;
#pragma unroll
      for (int r_ = 0; r_ < kTH; ++r_) {
        const int gy = ty0 + r_ - cshift;
        const bool row_ok = gy >= 0 && gy < ch;
        const long long off = (static_cast<long long>(gy) * cw + (tx0 - cshift)) * D2 + f0;
        const float *src0 = row_ok ? cn + off : cn;                // only dereferenced when ok
        const uint32_t dst0 = smem_u32(sg + r_ * Cfg::G_PITCH + f0);
#pragma unroll
        for (int i = 0; i < kIter; ++i) {
          const int f = f0 + i * kThreads * GE;
          if (kPerRow % kThreads == 0 || f < kTW * D2) {
            const bool ok = row_ok && f >= f_lo && f < f_hi;
            cp_async<GB>(dst0 + i * kThreads * GE * 4, ok ? src0 + i * kThreads * GE : cn, ok);
          }
        }
      }
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kPX; ++j)
#pragma unroll
      for (int k = 0; k < D2; ++k) gk[j][k] = sg[row * Cfg::G_PITCH + (x0 + 2 * j) * D2 + k];
    __syncthreads();   // sg is dead: the stages may be filled
    }

    // stage the other input's tile with its halo (zero outside the image = padding)
    // Loader thread (ly, lx) copies the 16-byte pieces lx, lx + 16, ... of tile rows ly, ly + 8, ...:
    // piece lx + 16 j is half (lx & 1) of pixel (lx >> 1) + 8 j, so the copies of a row differ by
    // compile-time offsets (8 pixels) from one global and one shared base (adding 8 to a pixel
    // index leaves the bank swizzle of smem_off unchanged).
    const int ly = threadIdx.x >> 4, lx = threadIdx.x & 15;
    const int lhalf = lx & 1, lpx = lx >> 1;
    auto issue_tma = [&](int chunk, unsigned u) {   // thread 0: chunk `chunk` of this tile is unit u
      const uint32_t bar = bar_full + 8 * (u & 1);
      tma::fence_proxy_async();
      tma::mbar_expect_tx(bar, Cfg::B_FLOATS * 4);
      tma::load_4d(tma::smem_u32(smem + (u & 1) * Cfg::B_FLOATS), &map_other, chunk * kCC, tx0 - Cfg::HALO,
                   ty0 - Cfg::HALO, n, bar);
    };
    auto issue = [&](int chunk) {
      float *sb = smem + (chunk & 1) * Cfg::B_FLOATS;
      const int c0 = chunk * kCC + 4 * lhalf;
      constexpr int kRowsPer = (Cfg::BH + kThreads / 16 - 1) / (kThreads / 16);
      constexpr int kPieces = (Cfg::BW * 2 + 15) / 16;
      const int gx0 = tx0 - Cfg::HALO + lpx;
#pragma unroll
      for (int rr = 0; rr < kRowsPer; ++rr) {
        const int py = ly + rr * (kThreads / 16);
        if (Cfg::BH % (kThreads / 16) == 0 || py < Cfg::BH) {
          const int gy = ty0 + py - Cfg::HALO;
          const bool row_ok = gy >= 0 && gy < H;
          const float *src0 = row_ok ? on + (static_cast<long long>(gy) * W + gx0) * C + c0 : on;
          const uint32_t dst0 = smem_u32(sb + smem_off(py * Cfg::BPITCH + lpx, lhalf));
#pragma unroll
          for (int j = 0; j < kPieces; ++j) {
            if (Cfg::BW * 2 % 16 == 0 || lx + 16 * j < Cfg::BW * 2) {
              const bool ok = row_ok && static_cast<unsigned>(gx0 + 8 * j) < static_cast<unsigned>(W);
              cp_async<16>(dst0 + j * 8 * kCC * 4, ok ? src0 + static_cast<long long>(j) * 8 * C : on, ok);
            }
          }
        }
      }
      cp_async_commit();
    };
    if (TMA) {
      if (threadIdx.x == 0) {
        issue_tma(0, it);
        if (n_chunks > 1) issue_tma(1, it + 1);
      }
    } else {
      issue(0);
    }
    for (int chunk = 0; chunk < n_chunks; ++chunk, ++it) {
      if (TMA) {
        tma::mbar_wait(bar_full + 8 * (it & 1), (it >> 1) & 1);
      } else {
        if (chunk + 1 < n_chunks) {
          issue(chunk + 1);       // its stage was released by the barrier that ended chunk - 1
          cp_async_wait<1>();
        } else {
          cp_async_wait<0>();
        }
        __syncthreads();
      }
      const float *sb = smem + ((TMA ? it : static_cast<unsigned>(chunk)) & 1) * Cfg::B_FLOATS;
      const int c0 = chunk * kCC;

      float4 acc[kPX][2];
#pragma unroll
      for (int j = 0; j < kPX; ++j) acc[j][0] = acc[j][1] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int pi = 0; pi < WN; ++pi) {
        const int p = REV ? WN - 1 - pi : pi;
        float4 vb[NB][2];
#pragma unroll
        for (int q = 0; q < NB; ++q)
#pragma unroll
          for (int half = 0; half < 2; ++half)
            vb[q][half] = *reinterpret_cast<const float4 *>(
                sb + smem_off((row + 2 * p) * Cfg::BPITCH + x0 + 2 * q, half));
#pragma unroll
        for (int oi = 0; oi < WN; ++oi) {
          const int o = REV ? WN - 1 - oi : oi;
#pragma unroll
          for (int j = 0; j < kPX; ++j) {
            const float w = gk[j][p * WN + o];
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              acc[j][half].x = fmaf(w, vb[j + o][half].x, acc[j][half].x);
              acc[j][half].y = fmaf(w, vb[j + o][half].y, acc[j][half].y);
              acc[j][half].z = fmaf(w, vb[j + o][half].z, acc[j][half].z);
              acc[j][half].w = fmaf(w, vb[j + o][half].w, acc[j][half].w);
            }
          }
        }
      }
      const int oy = ty0 + row;
      if (oy < H) {
#pragma unroll
        for (int j = 0; j < kPX; ++j) {
          const int ox = tx0 + x0 + 2 * j;
          if (ox < W) {
            float *d = dn + (static_cast<size_t>(oy) * W + ox) * C + c0;
            float4 v[2];
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              v[half] = acc[j][half];
              v[half].x = __fdiv_rn(v[half].x, sumelems); v[half].y = __fdiv_rn(v[half].y, sumelems);
              v[half].z = __fdiv_rn(v[half].z, sumelems); v[half].w = __fdiv_rn(v[half].w, sumelems);
            }
            // one 256-bit store per pixel chunk: the 32-byte sector is written whole (sm_100)
            asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(d), "f"(v[0].x),
                         "f"(v[0].y), "f"(v[0].z), "f"(v[0].w), "f"(v[1].x), "f"(v[1].y), "f"(v[1].z),
                         "f"(v[1].w)
                         : "memory");
          }
        }
      }
      __syncthreads();   // the stage may be refilled (cp.async: by issue(chunk + 2) of the next round)
      if (TMA && threadIdx.x == 0 && chunk + 2 < n_chunks) issue_tma(chunk + 2, it + 2);
    }
  }
}

template <int R>
int launch_k1(const float *grad, const float *a, const float *b, const GradGeom &g, float *ga,
              float *gb, cudaStream_t stream) {
  using Cfg = GradCfg<R>;
  const int shift = g.md - g.pad;
  const int tiles_x = ceil_div(g.W, kTW), tiles_y = ceil_div(g.H, kTH);
  const long long n_tiles = static_cast<long long>(tiles_x) * tiles_y * g.batch;
  if (n_tiles > 0x7FFFFFFFll) return 1;
  const int grid = static_cast<int>(n_tiles < 2 * kNumSMs ? n_tiles : 2 * kNumSMs);
  auto run = [&](auto kernel_tma, auto kernel_ldgsts, const float *coef, const float *other, int ch, int cw,
                 int cshift, float *dst) -> int {
    alignas(64) CUtensorMap map;
    const bool use_tma = tma::make_nhwc_map<kCC>(&map, other, g.batch, g.H, g.W, g.C, Cfg::BPITCH, Cfg::BH);
    if (use_tma) {
      DODT_CUDA_TRY(cudaFuncSetAttribute(kernel_tma, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(Cfg::SMEM)));
      kernel_tma<<<grid, kThreads, Cfg::SMEM, stream>>>(coef, other, map, g.batch, g.H, g.W, g.C, ch, cw, cshift,
                                                        tiles_x, tiles_y, dst);
    } else {
      DODT_CUDA_TRY(cudaFuncSetAttribute(kernel_ldgsts, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(Cfg::SMEM)));
      kernel_ldgsts<<<grid, kThreads, Cfg::SMEM, stream>>>(coef, other, map, g.batch, g.H, g.W, g.C, ch, cw,
                                                           cshift, tiles_x, tiles_y, dst);
    }
    DODT_AFTER_LAUNCH();
    return DODT_OK;
  };
  // 8-byte coefficient copies need every tile row to start 8-byte aligned (kTW is even)
  auto wide = [](const float *p, int cw, int cshift) {
    return reinterpret_cast<uintptr_t>(p) % 8 == 0 && cw % 2 == 0 && cshift % 2 == 0;
  };
  if (ga) {
    const int rc = wide(grad, g.out_w, shift)
                       ? run(corr_grad_k1<R, false, 8, true>, corr_grad_k1<R, false, 8, false>, grad, b, g.out_h,
                             g.out_w, shift, ga)
                       : run(corr_grad_k1<R, false, 4, true>, corr_grad_k1<R, false, 4, false>, grad, b, g.out_h,
                             g.out_w, shift, ga);
    if (rc != DODT_OK) return rc;
  }
  if (gb) {
    // gB: the same contraction walked backwards over the displacement-flipped gradient, gathered per
    // tile from the gradient map itself (FLIP)
    const int rc = wide(grad, g.out_w, shift)
                       ? run(corr_grad_k1<R, true, 8, true, true>, corr_grad_k1<R, true, 8, false, true>, grad, a,
                             g.out_h, g.out_w, shift, gb)
                       : run(corr_grad_k1<R, true, 4, true, true>, corr_grad_k1<R, true, 4, false, true>, grad, a,
                             g.out_h, g.out_w, shift, gb);
    if (rc != DODT_OK) return rc;
  }
  return DODT_OK;
}

bool k1_family(const GradGeom &g) {
  return g.ks == 1 && g.s1 == 1 && g.s2 == 2 && (g.r == 1 || g.r == 2) && g.C % kCC == 0 &&
         g.batch <= 65535;
}

int fill(int32_t batch, int32_t H, int32_t W, int32_t C, int32_t ks, int32_t md, int32_t s1,
         int32_t s2, int32_t pad, GradGeom *g) {
  if (batch <= 0 || H <= 0 || W <= 0 || C <= 0 || ks <= 0 || md < 0 || s1 <= 0 || s2 <= 0 || pad < 0)
    return DODT_EINVAL;
  int32_t hwc[3];
  const int rc = dodt_correlation_out_shape(H, W, ks, md, s1, s2, pad, hwc);  // odd ks, fits
  if (rc != DODT_OK) return rc;
  g->batch = batch; g->H = H; g->W = W; g->C = C;
  g->ks = ks; g->kr = (ks - 1) / 2; g->md = md; g->s1 = s1; g->s2 = s2; g->pad = pad;
  g->out_h = hwc[0]; g->out_w = hwc[1]; g->out_c = hwc[2];
  g->r = md / s2; g->wn = 2 * g->r + 1;
  return DODT_OK;
}

}  // namespace
}  // namespace dodt

extern "C" {

size_t dodt_correlation_grad_workspace_bytes(int32_t batch, int32_t height, int32_t width,
                                             int32_t channels, int32_t kernel_size,
                                             int32_t max_displacement, int32_t stride_1,
                                             int32_t stride_2, int32_t pad) {
  dodt::GradGeom g;
  if (dodt::fill(batch, height, width, channels, kernel_size, max_displacement, stride_1, stride_2,
                 pad, &g) != DODT_OK)
    return 0;
  return 0;   // no scratch since the displacement flip is gathered per tile (kept in the ABI)
}

int dodt_correlation_grad(const float *grad, const float *a, const float *b, int32_t batch,
                          int32_t height, int32_t width, int32_t channels, int32_t kernel_size,
                          int32_t max_displacement, int32_t stride_1, int32_t stride_2, int32_t pad,
                          float *grad_a, float *grad_b, void *workspace, size_t workspace_bytes,
                          dodt_stream_t stream_) {
  using namespace dodt;
  if (!grad || !a || !b || (!grad_a && !grad_b)) return DODT_EINVAL;
  GradGeom g;
  const int rc = fill(batch, height, width, channels, kernel_size, max_displacement, stride_1,
                      stride_2, pad, &g);
  if (rc != DODT_OK) return rc;
  cudaStream_t stream = as_stream(stream_);
  auto al16 = [](const void *p) { return reinterpret_cast<uintptr_t>(p) % 16 == 0; };
  (void)workspace;
  (void)workspace_bytes;
  if (k1_family(g) && al16(a) && al16(b) && (!grad_a || al16(grad_a)) && (!grad_b || al16(grad_b))) {
    int done = 1;
    if (g.r == 1) done = launch_k1<1>(grad, a, b, g, grad_a, grad_b, stream);
    if (g.r == 2) done = launch_k1<2>(grad, a, b, g, grad_a, grad_b, stream);
    if (done <= 0) return done;
  }
  const long long total = static_cast<long long>(g.batch) * g.H * g.W * g.C;
  const long long blocks = (total + 255) / 256;
  if (blocks > 0x7FFFFFFFll) return DODT_ECAPACITY;
  if (grad_a) {
    corr_grad_generic<0><<<static_cast<unsigned>(blocks), 256, 0, stream>>>(grad, b, g, total, grad_a);
    DODT_AFTER_LAUNCH();
  }
  if (grad_b) {
    corr_grad_generic<1><<<static_cast<unsigned>(blocks), 256, 0, stream>>>(grad, a, g, total, grad_b);
    DODT_AFTER_LAUNCH();
  }
  return DODT_OK;
}

}  // extern "C"
