// S4 — correlation, TMA-fed persistent kernel (round 2) for the DODT configuration family
// (kernel_size 1, stride_1 1, stride_2 2, neighbourhood radius R = max_displacement/2 in {1,2},
// C % 16 == 0). See correlation.cu for the reference citations and the generic kernels.
//
// Why this kernel exists (tools/micro/feed_bench.cu, profiles/r02_feed_bench.txt): feeding an SM
// with the tiles of an NHWC map costs 37 us per pair with 16-byte cp.async on 32-byte pieces (the
// round-1 loader: an 8-channel chunk of a pixel is one sector of its 128-byte line, 16 lines per
// warp instruction, every thread spending issue slots and L1TEX wavefronts on copies), 30.8 us with
// tensor-TMA boxes whose inner row is 32 bytes, 18.5 / 16.5 us with 64- / 128-byte inner rows. A
// thread's register tile (100 accumulators) caps the SM at 256 threads, and a 16-channel unit of an
// 8 x 64 tile is 107 KB: only one stage per CTA fits and a CTA cannot refill the stage it reads
// (measured 43.4 us per pair, a third of the warp samples in the mbarrier wait). Prefetch depth
// beats feed rate: the product configuration is
//
//   feed      8-channel units (an 8 x 64-pixel output tile x 8 channels) in a TWO-stage ring: one
//             elected thread issues two cp.async.bulk.tensor.4d loads (UTMALDG) per unit — the A
//             tile [8 x 66 px] and the B tile with its 4-pixel halo [16 x 74 px], 32-byte pixel
//             vectors under SWIZZLE_32B — two units ahead of their use. Out-of-image elements are
//             zero-filled by the TMA unit, which IS the reference's zero padding (PadData and the
//             two padded temporaries of pad.cu.cc / correlation_kernel.cc:69-107 never exist).
//             Completion arrives on an mbarrier (complete_tx). No thread issues a copy.
//   overlap   two CTAs per SM; one __syncthreads per unit hands the consumed stage back (warps
//             that free-run and return stages through a counter measured slower: 38.0 vs 36.5 us).
//   math      a thread owns 2 x 2 pixels spaced 2 apart in both directions and all (2R+1)^2
//             displacements of each (100 fp32 accumulators at R = 2) across the channel chunks;
//             per B row it reads 6 B-pixel float4s from shared memory (36 + 4 LDS.128 per 400
//             FFMA; the 4 x 1 tile of round 1, kept as an A/B instantiation, needs 40 + 4). Tile
//             pitches are 2 (mod 8) pixels, so under the swizzle the eight lanes of a quarter-warp
//             hit eight different 16-byte bank groups.
//   epilogue  the finished tile (8 x 64 px x 25 floats) is staged through the stage just consumed
//             (the next tile's first unit is already in flight in the other stage) and written
//             to HBM as coalesced rows.
//
// Summation order per output element: channels in ascending order, one fused multiply-add each,
// the same as corr_async_k1 (correlation_tma.cu) — results are bit-identical to it.
//
// Algorithmic HBM bytes per launch: 2*H*W*C*4 read once + H*W*D^2*4 written once (199.36 MB at
// 700x800x32, D^2 = 25); halo re-reads are served by the 126 MB L2. 35.9 us per pair in the 8-pair
// frame-stream launch (profiles/r02_corr_feed_ncu.txt: shared-memory pipe 77 %, DRAM 54 %).
#include <cuda.h>

#include "common.cuh"
#include "tma_util.cuh"

namespace dodt {
namespace {

constexpr int kTW = 64, kTH = 8, kPX = 4;
constexpr int kThr = kTW / (2 * kPX) * 2 * kTH;   // 128
constexpr int kMaxPairs = DODT_CORR_STREAM_MAX_PAIRS;
constexpr int kExclusiveSmem = 116 * 1024;   // > (228 KB - 2 x 1 KB reserved) / 2: two such CTAs never share an SM

// CH: channels per work unit (16: 64-byte pixel vectors, SWIZZLE_64B; 8: 32-byte vectors, SWIZZLE_32B)
// NST: stages per CTA (unit u lives in stage u % NST and is requested NST units ahead)
template <int R, int CH, int NST>
struct FeedCfg {
  static constexpr int kCH = CH;
  static constexpr int WN = 2 * R + 1;
  static constexpr int D2 = WN * WN;
  static constexpr int HALO = 2 * R;                   // stride_2 == 2
  static constexpr int AW = kTW + 2;                   // pitch = 2 (mod 8)
  static constexpr int BW = ((kTW + 2 * HALO + 7) / 8) * 8 + 2;
  static constexpr int BH = kTH + 2 * HALO;
  static constexpr int A_BYTES = kTH * AW * kCH * 4;
  static constexpr int B_BYTES = BH * BW * kCH * 4;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int OUT_PITCH = kTW * D2 + 2;       // floats; +2 keeps staging stores conflict-free
  static constexpr int OUT_BYTES = kTH * OUT_PITCH * 4;
  static constexpr int STAGE_PITCH = ((STAGE_BYTES + 1023) / 1024) * 1024;
  static constexpr int BAR_OFF = NST * STAGE_PITCH;
  static constexpr int SMEM_BYTES = BAR_OFF + 64;
  static_assert(A_BYTES % (CH * 32) == 0, "the B tile must start on a swizzle repeat");
  static_assert(OUT_BYTES <= STAGE_BYTES, "output staging fits the stage");
  static_assert(AW % 8 == 2 && BW % 8 == 2, "pitches 2 (mod 8) keep float4 reads conflict-free");
};

// 64-byte pixel vectors under CU_TENSOR_MAP_SWIZZLE_64B: bits [4,5] of the byte offset are XORed
// with bits [7,8], i.e. the 16-byte piece `c` of pixel p sits at piece c ^ ((p >> 1) & 3).
// 32-byte vectors under SWIZZLE_32B: bit 4 is XORed with bit 7, piece c sits at c ^ ((p >> 2) & 1).
// Offset (in floats) of piece 0; piece c is at (offset ^ (c << 2)).
template <int CH>
__device__ __forceinline__ int swz(int pixel) {
  return CH == 16 ? pixel * 16 + (((pixel >> 1) & 3) << 2) : pixel * 8 + (((pixel >> 2) & 1) << 2);
}
// B rows advance by 2*BW = 148 pixels per displacement row p: the swizzle term of a pixel changes by
// 74p mod 4 = 2(p & 1) under SWIZZLE_64B and by 37p mod 2 = p & 1 under SWIZZLE_32B
template <int CH>
__device__ __forceinline__ constexpr int row_flip(int p) { return CH == 16 ? ((p & 1) << 1) : (p & 1); }

struct FeedGeom {
  int batch, C, out_h, out_w, shift;   // shift = max_displacement - pad
  int tiles_x, tiles_y, n_tiles;
  int pow2;                            // C is a power of two: divide by multiplying with 1/C (exact)
  float inv_c;
  int n_stream;                        // > 0: item n correlates image n with image n + 1 into outs[n]
  float *outs[kMaxPairs];
};

// Tensor maps of one launch. Pair (or batch item group) n reads its A tile through a[n] and its B
// tile through b[n]; without the frame-stream form only a[0] / b[0] are used (4-D over the batch).
struct FeedMaps {
  CUtensorMap a[kMaxPairs];
  CUtensorMap b[kMaxPairs];
};

// SQ: a thread owns 2 x 2 pixels (spaced 2 apart in both directions) instead of 4 in a row: the
// B positions its (pixel, displacement) pairs need shrink from 8 x WN to (WN+1) x (WN+1) per
// 4-channel step (36 instead of 40 LDS.128 at R = 2), same 4 x WN^2 accumulators, same bits.
template <int R, int CH, int NST, bool SQ>
__global__ void __launch_bounds__(kThr, 2)
corr_feed_k1(const __grid_constant__ FeedMaps maps, const __grid_constant__ FeedGeom g, float *__restrict__ out) {
  using Cfg = FeedCfg<R, CH, NST>;
  constexpr int kCH = CH;
  static_assert(Cfg::BW == 74, "row_flip() is written for a 74-pixel B pitch");
  constexpr int WN = Cfg::WN, D2 = Cfg::D2, NB = kPX + 2 * R;
  constexpr int kWarps = kThr / 32;
  constexpr int kWX = kTW / (8 * kPX);                  // warps along x
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t smem_base = tma::smem_u32(smem);
  const uint32_t bar_full = smem_base + Cfg::BAR_OFF;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_chunks = g.C / kCH;

  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < NST; ++s) tma::mbar_init(bar_full + 8 * s, 1);
    tma::mbar_fence_init();
  }
  __syncthreads();

  // unit u = (tile, channel chunk) of this CTA, in order, lives in stage u % NST; issued by thread 0
  const int my_tiles = (g.n_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) /
                       static_cast<int>(gridDim.x);
  const int n_units = my_tiles * n_chunks;
  auto issue = [&](int u) {
    const int tile = blockIdx.x + (u / n_chunks) * gridDim.x;
    const int ch = u % n_chunks;
    const uint32_t stage_base = smem_base + static_cast<uint32_t>(u % NST) * Cfg::STAGE_PITCH;
    const uint32_t bar = bar_full + 8 * (u % NST);
    // batch-interleaved tile order: the same spatial tile of consecutive items is worked on at the
    // same time (by neighbouring CTAs), so a map that two items share is fetched from HBM once
    const int n = tile % g.batch, sp = tile / g.batch;
    const int tx = sp % g.tiles_x, ty = sp / g.tiles_x;
    const int ax = tx * kTW + g.shift, ay = ty * kTH + g.shift;
    const CUtensorMap *ma = g.n_stream ? &maps.a[n] : &maps.a[0];
    const CUtensorMap *mb = g.n_stream ? &maps.b[n] : &maps.b[0];
    const int item = g.n_stream ? 0 : n;
    // the stage was read (math) or written (output staging) through the generic proxy
    tma::fence_proxy_async();
    tma::mbar_expect_tx(bar, Cfg::STAGE_BYTES);
    tma::load_4d(stage_base, ma, ch * kCH, ax, ay, item, bar);
    tma::load_4d(stage_base + Cfg::A_BYTES, mb, ch * kCH, ax - Cfg::HALO, ay - Cfg::HALO, item, bar);
  };

  // ---- compute role
  // row tile:    lane bits [0] parity, [1..2] row & 3, [3..4] group & 3; pixels (row, x0 + 2j)
  // square tile: lane bits [0] x parity, [1] g & 1, [2] y0 & 1, [3..4] (g >> 1) & 3, warp bits [0] g >> 3,
  //              [1] y0 >> 2; pixels (row + 2 iy, x0 + 2 ix), x0 = 4 g + parity, row in {0, 1, 4, 5}.
  // Either way the eight lanes of a quarter-warp read eight pixels whose indices differ mod 8
  // (pitches are 2 mod 8), i.e. eight different 16-byte bank groups under the swizzle.
  constexpr int NBQ = SQ ? WN + 1 : NB;            // B columns a thread reads per B row
  constexpr int NBR = SQ ? WN + 1 : WN;            // B rows a thread reads per 4-channel step
  int row, x0;
  if constexpr (SQ) {
    const int g4 = ((lane >> 1) & 1) | (((lane >> 3) & 3) << 1) | ((warp & 1) << 3);
    row = ((warp >> 1) << 2) | ((lane >> 2) & 1);
    x0 = 4 * g4 + (lane & 1);
  } else {
    row = (warp / kWX) * 4 + ((lane >> 1) & 3);
    x0 = ((warp % kWX) * 4 + (lane >> 3)) * (2 * kPX) + (lane & 1);
  }
  // pixel j of the thread: row tile (row, x0 + 2j); square tile (row + 2 (j >> 1), x0 + 2 (j & 1))
  auto px_row = [&](int j) { return SQ ? row + 2 * (j >> 1) : row; };
  auto px_col = [&](int j) { return SQ ? x0 + 2 * (j & 1) : x0 + 2 * j; };
  int aoff[kPX], boff[NBQ];
#pragma unroll
  for (int j = 0; j < kPX; ++j) aoff[j] = swz<CH>(px_row(j) * Cfg::AW + px_col(j));
#pragma unroll
  for (int q = 0; q < NBQ; ++q) boff[q] = swz<CH>(row * Cfg::BW + x0 + 2 * q);

  if (threadIdx.x == 0) {
#pragma unroll
    for (int u = 0; u < NST; ++u)
      if (u < n_units) issue(u);
  }

  int it = 0;
  for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x) {
    float acc[kPX][D2];
#pragma unroll
    for (int j = 0; j < kPX; ++j)
#pragma unroll
      for (int k = 0; k < D2; ++k) acc[j][k] = 0.0f;

    int stage = 0;
    for (int ch = 0; ch < n_chunks; ++ch, ++it) {
      stage = it % NST;
      const float *sa = reinterpret_cast<const float *>(smem + stage * Cfg::STAGE_PITCH);
      const float *sb = sa + Cfg::A_BYTES / 4;
      tma::mbar_wait(bar_full + 8 * stage, (it / NST) & 1);
#pragma unroll
      for (int cv = 0; cv < kCH / 4; ++cv) {
        float4 va[kPX];
#pragma unroll
        for (int j = 0; j < kPX; ++j)
          va[j] = *reinterpret_cast<const float4 *>(sa + (aoff[j] ^ (cv << 2)));
#pragma unroll
        for (int br = 0; br < NBR; ++br) {
          float4 vb[NBQ];
          // rows advance by 2*BW = 148 pixels: the swizzle term of a pixel flips with every odd br;
          // the rest of the offset is a compile-time constant
#pragma unroll
          for (int q = 0; q < NBQ; ++q)
            vb[q] = *reinterpret_cast<const float4 *>(
                sb + (boff[q] ^ ((row_flip<CH>(br) ^ cv) << 2)) + br * 2 * Cfg::BW * kCH);
#pragma unroll
          for (int j = 0; j < kPX; ++j) {
            // square tile: B row br serves displacement row p = br - iy of pixel row iy
            const int p = SQ ? br - (j >> 1) : br;
            const int jc = SQ ? (j & 1) : j;
            if (p < 0 || p >= WN) continue;
#pragma unroll
            for (int o = 0; o < WN; ++o) {
              float s = acc[j][p * WN + o];
              s = fmaf(va[j].x, vb[jc + o].x, s);
              s = fmaf(va[j].y, vb[jc + o].y, s);
              s = fmaf(va[j].z, vb[jc + o].z, s);
              s = fmaf(va[j].w, vb[jc + o].w, s);
              acc[j][p * WN + o] = s;
            }
          }
        }
      }
      __syncthreads();   // everyone is done reading the stage
      // the last unit of a tile hands its stage to the epilogue first
      if (threadIdx.x == 0 && ch + 1 < n_chunks && it + NST < n_units) issue(it + NST);
    }

    // ---- epilogue: stage the tile through the (free) stage, coalesced stores
    const int n = tile % g.batch, sp = tile / g.batch;
    const int tx = sp % g.tiles_x, ty = sp / g.tiles_x;
    float *stg = reinterpret_cast<float *>(smem + stage * Cfg::STAGE_PITCH);   // the stage just consumed
    if (g.pow2) {
#pragma unroll
      for (int j = 0; j < kPX; ++j) {
        float *dst = stg + px_row(j) * Cfg::OUT_PITCH + px_col(j) * D2;
#pragma unroll
        for (int k = 0; k < D2; ++k) dst[k] = __fmul_rn(acc[j][k], g.inv_c);   // exact: 1/2^k
      }
    } else {
      const float sumelems = static_cast<float>(g.C);
#pragma unroll
      for (int j = 0; j < kPX; ++j) {
        float *dst = stg + px_row(j) * Cfg::OUT_PITCH + px_col(j) * D2;
#pragma unroll
        for (int k = 0; k < D2; ++k) dst[k] = __fdiv_rn(acc[j][k], sumelems);
      }
    }
    __syncthreads();
    const int valid_rows = min(kTH, g.out_h - ty * kTH);
    const int valid_floats = min(kTW, g.out_w - tx * kTW) * D2;
    float *out_img = g.n_stream ? g.outs[n] : out + static_cast<size_t>(n) * g.out_h * g.out_w * D2;
    float *gout = out_img + (static_cast<size_t>(ty * kTH) * g.out_w + tx * kTW) * D2;
    const bool vec_ok = (reinterpret_cast<uintptr_t>(out_img) % 8 == 0) && ((g.out_w * D2) % 2 == 0) &&
                        (valid_floats % 2 == 0);
    for (int r = warp; r < valid_rows; r += kWarps) {
      const float *src = stg + r * Cfg::OUT_PITCH;
      float *dstrow = gout + static_cast<size_t>(r) * g.out_w * D2;
      if (vec_ok) {
        for (int e = lane; e < valid_floats / 2; e += 32)
          reinterpret_cast<float2 *>(dstrow)[e] = reinterpret_cast<const float2 *>(src)[e];
      } else {
        for (int e = lane; e < valid_floats; e += 32) dstrow[e] = src[e];
      }
    }
    __syncthreads();   // staging rows are read out: the stage may be refilled
    if (threadIdx.x == 0 && it - 1 + NST < n_units) issue(it - 1 + NST);
  }
}

template <int R, int CH, int NST, bool SQ = true>
int launch_feed(const float *a, const float *b, int N, int H, int W, int C, int out_h, int out_w, int shift,
                float *out, int max_ctas, cudaStream_t stream, const float *const *maps, float *const *outs) {
  using Cfg = FeedCfg<R, CH, NST>;
  FeedMaps tm;
  FeedGeom g;
  g.batch = N; g.C = C; g.out_h = out_h; g.out_w = out_w; g.shift = shift;
  g.tiles_x = ceil_div(out_w, kTW);
  g.tiles_y = ceil_div(out_h, kTH);
  g.n_tiles = g.tiles_x * g.tiles_y * N;
  g.pow2 = (C & (C - 1)) == 0 ? 1 : 0;
  g.inv_c = 1.0f / static_cast<float>(C);
  g.n_stream = 0;
  for (int k = 0; k < kMaxPairs; ++k) g.outs[k] = nullptr;
  if (maps) {   // frame-stream form: N pairs over N + 1 images
    g.n_stream = N;
    for (int k = 0; k < N; ++k) {
      if (!tma::make_nhwc_map<CH>(&tm.a[k], maps[k], 1, H, W, C, Cfg::AW, kTH) ||
          !tma::make_nhwc_map<CH>(&tm.b[k], maps[k + 1], 1, H, W, C, Cfg::BW, Cfg::BH))
        return 1;
      g.outs[k] = outs[k];
    }
    for (int k = N; k < kMaxPairs; ++k) { tm.a[k] = tm.a[0]; tm.b[k] = tm.b[0]; }
  } else {
    if (!tma::make_nhwc_map<CH>(&tm.a[0], a, N, H, W, C, Cfg::AW, kTH) || !tma::make_nhwc_map<CH>(&tm.b[0], b, N, H, W, C, Cfg::BW, Cfg::BH))
      return 1;  // driver without tensor maps: the caller falls back to the cp.async kernel
    for (int k = 1; k < kMaxPairs; ++k) { tm.a[k] = tm.a[0]; tm.b[k] = tm.b[0]; }
  }
  int grid = g.n_tiles < 2 * kNumSMs ? g.n_tiles : 2 * kNumSMs;
  if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;   // persistent CTAs: any count works
  // A cap at or below the SM count asks for ONE CTA per SM, the other half of every SM (registers,
  // shared memory) left to the co-running kernels of the frame stream: the launch then requests
  // more than half of an SM's shared memory, so the block scheduler cannot put two of its CTAs
  // (or one of a second correlation launch) on the same SM.
  int smem_bytes = Cfg::SMEM_BYTES;
  if (max_ctas > 0 && max_ctas <= kNumSMs && smem_bytes < kExclusiveSmem) smem_bytes = kExclusiveSmem;
  // the attribute belongs to the (function, device) pair: set it on every launch (cheap)
  DODT_CUDA_TRY(cudaFuncSetAttribute(corr_feed_k1<R, CH, NST, SQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  corr_feed_k1<R, CH, NST, SQ><<<grid, kThr, smem_bytes, stream>>>(tm, g, out);
  DODT_AFTER_LAUNCH();
  return DODT_OK;
}

constexpr int kFeedCH = 8, kFeedNST = 2;

bool feed_applies(int C, int out_h, int out_w, int H, int W) {
  // boxes are at most 256 elements per dimension and strides multiples of 16 bytes by construction;
  // tiny maps are mostly padding and stay with the tile kernel
  return C % 16 == 0 && static_cast<long long>(out_h) * out_w >= 1024 && H >= 1 && W >= 1;
}

// diagnostic build: DODT_CORR_FEED = 1 -> 8-channel units, two stages; 2 -> 16-channel units, one stage;
// 3 / 4 -> 8-channel units, four / three stages (one CTA per SM); 5 -> the product configuration with the
// 4 x 1 pixel tile of round 1 instead of 2 x 2
template <int R>
int launch_variant(const float *a, const float *b, int N, int H, int W, int C, int out_h, int out_w, int shift,
                   float *out, int max_ctas, cudaStream_t stream, const float *const *maps, float *const *outs) {
#ifdef DODT_DIAG
  static int v = -1;
  if (v < 0) v = DODT_KNOB("DODT_CORR_FEED", 1);
  if (v == 2) return launch_feed<R, 16, 1>(a, b, N, H, W, C, out_h, out_w, shift, out, max_ctas, stream, maps, outs);
  if (v == 3) return launch_feed<R, 8, 4>(a, b, N, H, W, C, out_h, out_w, shift, out, max_ctas, stream, maps, outs);
  if (v == 4) return launch_feed<R, 8, 3>(a, b, N, H, W, C, out_h, out_w, shift, out, max_ctas, stream, maps, outs);
  if (v == 5) return launch_feed<R, kFeedCH, kFeedNST, false>(a, b, N, H, W, C, out_h, out_w, shift, out, max_ctas, stream, maps, outs);
#endif
  return launch_feed<R, kFeedCH, kFeedNST>(a, b, N, H, W, C, out_h, out_w, shift, out, max_ctas, stream, maps, outs);
}

}  // namespace

// returns DODT_OK if launched, 1 if this path does not apply, DODT_E* on failure
int correlation_feed(const float *a, const float *b, int N, int H, int W, int C, int r, int out_h,
                     int out_w, int shift, float *out, int max_ctas, cudaStream_t stream) {
  if (!feed_applies(C, out_h, out_w, H, W) || reinterpret_cast<uintptr_t>(a) % 16 ||
      reinterpret_cast<uintptr_t>(b) % 16)
    return 1;
  switch (r) {
    case 1: return launch_variant<1>(a, b, N, H, W, C, out_h, out_w, shift, out, max_ctas, stream, nullptr, nullptr);
    case 2: return launch_variant<2>(a, b, N, H, W, C, out_h, out_w, shift, out, max_ctas, stream, nullptr, nullptr);
    default: return 1;
  }
}

// Frame-stream form: pair j correlates maps[j] (frame t) with maps[j + 1] (frame t + 1) into
// outs[j], all pairs in ONE launch with batch-interleaved tiles, so that a map shared by two pairs
// (B of pair j, A of pair j + 1) is read from HBM once and served from L2 the second time.
int correlation_stream_feed(const float *const *maps, int n_pairs, float *const *outs, int H, int W,
                            int C, int r, int out_h, int out_w, int shift, int max_ctas,
                            cudaStream_t stream) {
  if (n_pairs < 1 || n_pairs > kMaxPairs || !feed_applies(C, out_h, out_w, H, W)) return 1;
  for (int k = 0; k <= n_pairs; ++k)
    if (reinterpret_cast<uintptr_t>(maps[k]) % 16) return 1;
  switch (r) {
    case 1: return launch_variant<1>(nullptr, nullptr, n_pairs, H, W, C, out_h, out_w, shift, nullptr, max_ctas,
                                  stream, maps, outs);
    case 2: return launch_variant<2>(nullptr, nullptr, n_pairs, H, W, C, out_h, out_w, shift, nullptr, max_ctas,
                                  stream, maps, outs);
    default: return 1;
  }
}

}  // namespace dodt
