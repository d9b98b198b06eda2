// S4 — correlation, TMA-pipelined persistent kernel for the DODT configuration family
// (kernel_size 1, stride_1 1, stride_2 2, neighbourhood radius R = max_displacement/2 in {1,2},
// C % 8 == 0). See correlation.cu for the reference citations and the generic kernels.
//
// corr_tma_k1: one persistent CTA per SM (grid = 148), 8 warps; thread 0 doubles as the TMA
// producer. Work unit = a 16 x 64 tile of output pixels x one chunk of 8 channels.
// corr_async_k1 (default, further down): same math, fed by cp.async, 8 x 64 tiles, 2 CTAs per SM.
//
//   producer  cp.async.bulk.tensor.4d (TMA) loads the A tile [16 x 66 px x 8 ch] and the B tile
//             with its 4-pixel halo [24 x 74 px x 8 ch] into one of two 90 KB stages; tile
//             coordinates may be negative or beyond the image — TMA zero-fills out-of-bounds
//             elements, which IS the reference's zero padding (PadData + two padded temporaries in
//             pad.cu.cc / correlation_kernel.cc:69-107 are never materialised). Completion is
//             signalled on an mbarrier (complete_tx), buffers are handed back through a second
//             mbarrier, so loads of chunk i+1 overlap the math of chunk i without any CTA-wide
//             barrier.
//   consumers each thread owns 4 pixels spaced 2 apart on one row (x0, x0+2, x0+4, x0+6) and all
//             25 displacements of each: 100 fp32 accumulators live in registers across the four
//             channel chunks. Per displacement row it reads 8 B-pixel float4s from shared memory
//             and feeds 20 (pixel, displacement) pairs from them — 10 FMAs per shared-memory
//             float, which is what keeps the kernel off the shared-memory roofline. Pixel vectors
//             are 32 bytes with TMA's 32-byte swizzle and the tile pitches are 2 (mod 8) pixels, so
//             the eight lanes of a quarter-warp always hit eight different 16-byte bank groups.
//   epilogue  the finished tile (16 x 64 px x 25 floats) is staged through the shared memory of
//             the stage that was just consumed (plus a 12 KB spare region) and written to HBM
//             as fully coalesced 128-byte rows; the 100-byte pixel stride of the NHWC(25) output
//             would otherwise turn every store into 32 partial sectors.
//
// Algorithmic HBM bytes per launch: 2*H*W*C*4 read once + H*W*25*4 written once (199.36 MB at
// 700x800x32); halo re-reads (1.74x of B) are served by the 126 MB L2.
#include <cuda.h>

#include "common.cuh"

namespace dodt {
namespace {

constexpr int kTW = 64, kTH = 16, kPX = 4, kCC = 8;
constexpr int kConsumers = kTW / (2 * kPX) * 2 * kTH;  // 256
// No dedicated producer warp: ptxas budgets registers for the thread count rounded up to 128, so
// 288 threads would cap the 100-accumulator consumers at 168 registers (spills). Thread 0 issues
// the two TMA loads of the NEXT work unit before it starts computing the current one.
constexpr int kThreads = kConsumers;

template <int R, int TH = kTH>
struct TmaCfg {
  static constexpr int WN = 2 * R + 1;
  static constexpr int D2 = WN * WN;
  static constexpr int HALO = 2 * R;                   // stride_2 == 2
  static constexpr int AW = kTW + 2;                   // pitch = 2 (mod 8)
  static constexpr int BW = ((kTW + 2 * HALO + 7) / 8) * 8 + 2;
  static constexpr int BH = TH + 2 * HALO;
  static constexpr int A_BYTES = TH * AW * kCC * 4;
  static constexpr int B_BYTES = BH * BW * kCC * 4;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int OUT_PITCH = kTW * D2 + 2;       // floats; +2 keeps staging stores conflict-free
  static constexpr int OUT_BYTES = TH * OUT_PITCH * 4;
  static constexpr int SPARE = OUT_BYTES > STAGE_BYTES ? OUT_BYTES - STAGE_BYTES : 0;
  static constexpr int STAGE1_OFF = ((STAGE_BYTES + SPARE + 1023) / 1024) * 1024;
  static constexpr int BAR_OFF = STAGE1_OFF + STAGE_BYTES;
  static constexpr int SMEM_BYTES = BAR_OFF + 64;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap *map, int c0, int c1,
                                            int c2, int c3, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%2, %3, %4, %5}], [%6];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3),
      "r"(bar)
      : "memory");
}
__device__ __forceinline__ void consumer_sync() {   // the 256 compute threads only
  asm volatile("bar.sync 1, %0;" ::"n"(kConsumers) : "memory");
}

// 32-byte pixel vectors under CU_TENSOR_MAP_SWIZZLE_32B: bit 4 of the byte offset is XORed with
// bit 7, i.e. the two 16-byte halves of pixel p swap when bit 2 of p is set.
__device__ __forceinline__ int swz(int pixel, int half) {
  return pixel * kCC + ((half ^ ((pixel >> 2) & 1)) << 2);
}

struct CorrTmaGeom {
  int batch, C, out_h, out_w, shift;  // shift = max_displacement - pad
  int tiles_x, tiles_y, n_tiles;
  float inv_unused;
};

template <int R>
__global__ void __launch_bounds__(kThreads, 1)
corr_tma_k1(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
            const CorrTmaGeom g, float *__restrict__ out) {
  using Cfg = TmaCfg<R>;
  static_assert(Cfg::A_BYTES % 1024 == 0, "B tile must stay 1024-byte aligned for the swizzle");
  constexpr int WN = Cfg::WN, D2 = Cfg::D2, NB = kPX + 2 * R;
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bar_full = smem_base + Cfg::BAR_OFF;        // [2]
  const uint32_t bar_empty = smem_base + Cfg::BAR_OFF + 16;  // [2]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_chunks = g.C / kCC;

  if (threadIdx.x == 0) {
    mbar_init(bar_full, 1);
    mbar_init(bar_full + 8, 1);
    mbar_init(bar_empty, kConsumers / 32);
    mbar_init(bar_empty + 8, kConsumers / 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // ------------------------------------------------------------------ TMA issue (thread 0)
  // work unit `u` = (tile, channel chunk); unit u lives in stage u & 1
  const int my_tiles = (g.n_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) /
                       static_cast<int>(gridDim.x);
  const int n_units = my_tiles * n_chunks;
  auto issue = [&](int u) {
    const int tile = blockIdx.x + (u / n_chunks) * gridDim.x;
    const int ch = u % n_chunks;
    const int tx = tile % g.tiles_x;
    const int ty = (tile / g.tiles_x) % g.tiles_y;
    const int n = tile / (g.tiles_x * g.tiles_y);
    const int ax = tx * kTW + g.shift, ay = ty * kTH + g.shift;
    const int stage = u & 1;
    const uint32_t base = smem_base + (stage ? Cfg::STAGE1_OFF : 0);
    mbar_wait(bar_empty + 8 * stage, ((u >> 1) & 1) ^ 1);   // consumers released the stage
    mbar_expect_tx(bar_full + 8 * stage, Cfg::STAGE_BYTES);
    tma_load_4d(base, &map_a, ch * kCC, ax, ay, n, bar_full + 8 * stage);
    tma_load_4d(base + Cfg::A_BYTES, &map_b, ch * kCC, ax - Cfg::HALO, ay - Cfg::HALO, n,
                bar_full + 8 * stage);
  };
  if (threadIdx.x == 0 && n_units > 0) issue(0);

  // -------------------------------------------------------------------- consumers
  // lane bits: [0] parity, [1..2] row & 3, [3..4] group & 3; warps tile 4 (rows) x 2 (x halves)
  const int row = (warp >> 1) * 4 + ((lane >> 1) & 3);
  const int x0 = ((warp & 1) * 4 + (lane >> 3)) * (2 * kPX) + (lane & 1);
  const float sumelems = static_cast<float>(g.C);

  int it = 0;
  for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x) {
    float acc[kPX][D2];
#pragma unroll
    for (int j = 0; j < kPX; ++j)
#pragma unroll
      for (int k = 0; k < D2; ++k) acc[j][k] = 0.0f;

    int stage = 0;
    for (int ch = 0; ch < n_chunks; ++ch, ++it) {
      stage = it & 1;
      if (threadIdx.x == 0 && it + 1 < n_units) issue(it + 1);   // prefetch one unit ahead
      __syncwarp();
      const float *sa = reinterpret_cast<const float *>(smem + (stage ? Cfg::STAGE1_OFF : 0));
      const float *sb = sa + Cfg::A_BYTES / 4;
      mbar_wait(bar_full + 8 * stage, (it >> 1) & 1);
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        float4 va[kPX];
#pragma unroll
        for (int j = 0; j < kPX; ++j)
          va[j] = *reinterpret_cast<const float4 *>(sa + swz(row * Cfg::AW + x0 + 2 * j, half));
#pragma unroll
        for (int p = 0; p < WN; ++p) {
          float4 vb[NB];
#pragma unroll
          for (int q = 0; q < NB; ++q)
            vb[q] = *reinterpret_cast<const float4 *>(
                sb + swz((row + 2 * p) * Cfg::BW + x0 + 2 * q, half));
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
      if (ch + 1 < n_chunks) {
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_empty + 8 * stage);
      }
    }

    // ---- epilogue: stage the tile through the just-consumed stage (+ spare), coalesced stores
    const int tx = tile % g.tiles_x;
    const int ty = (tile / g.tiles_x) % g.tiles_y;
    const int n = tile / (g.tiles_x * g.tiles_y);
    float *stg = reinterpret_cast<float *>(
        smem + (stage ? Cfg::STAGE1_OFF + Cfg::STAGE_BYTES - Cfg::OUT_BYTES : 0));
    consumer_sync();  // every consumer is done reading this stage
#pragma unroll
    for (int j = 0; j < kPX; ++j) {
      float *dst = stg + row * Cfg::OUT_PITCH + (x0 + 2 * j) * D2;
#pragma unroll
      for (int k = 0; k < D2; ++k) dst[k] = __fdiv_rn(acc[j][k], sumelems);
    }
    consumer_sync();
    const int valid_rows = min(kTH, g.out_h - ty * kTH);
    const int valid_floats = min(kTW, g.out_w - tx * kTW) * D2;
    float *gout = out + ((static_cast<size_t>(n) * g.out_h + ty * kTH) * g.out_w + tx * kTW) * D2;
    for (int r = warp; r < valid_rows; r += kConsumers / 32) {
      const float *src = stg + r * Cfg::OUT_PITCH;
      float *dstrow = gout + static_cast<size_t>(r) * g.out_w * D2;
      for (int e = lane; e < valid_floats; e += 32) dstrow[e] = src[e];
    }
    // generic-proxy accesses to this stage are finished before TMA (async proxy) refills it
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_empty + 8 * stage);
  }
}


// ---------------------------------------------------------------------------------------------
// cp.async variant (default). ncu on corr_tma_k1 (profiles/r01_corr_tma_ncu.json) shows the
// tensor-TMA engine delivering only ~8 B/clk/SM for this access pattern: an 8-channel chunk of
// an NHWC pixel is a 32-byte box row and the engine issues roughly one row request every four
// cycles, so the consumers sit on the full-barrier. Here the same two-stage pipeline is fed by
// the 256 compute threads themselves with 16-byte cp.async.cg (LDGSTS, zero-fill for the
// reference's padding): a thread's 26 copies per work unit differ only by compile-time offsets
// from one shared-memory and one global base, so the issue cost is ~150 instructions per 900 of
// math, and one __syncthreads per unit both publishes stage u and frees stage u-1.
// ---------------------------------------------------------------------------------------------
constexpr int kMaxStreamPairs = DODT_CORR_STREAM_MAX_PAIRS;
struct CorrAsyncGeom {
  int batch, H, W, C, out_h, out_w, shift;
  int tiles_x, tiles_y, n_tiles;
  int pow2;       // C is a power of two: divide by multiplying with the exact reciprocal
  float inv_c;
  // frame-stream mode (n_stream > 0): item n correlates maps[n] with maps[n + 1] into outs[n];
  // the a / b / out kernel arguments are unused
  int n_stream;
  const float *maps[kMaxStreamPairs + 1];
  float *outs[kMaxStreamPairs];
};

// PX pixels per thread (spaced 2 apart). PX = 4: 256 threads, 100 accumulators, 10 FMAs per
// shared-memory float; PX = 2: 512 threads (16 warps), 50 accumulators, 6.7 FMAs per float.
// NST: stages of the cp.async ring (unit u lives in stage u % NST and is requested NST-1 units
// ahead of its use).
template <int R, int PX, int FRONT, int TH, int CTAS, int NST, int DBG = 0>
__global__ void __launch_bounds__(kTW / (2 * PX) * 2 * TH, CTAS)
corr_async_k1(const float *__restrict__ a, const float *__restrict__ b, const CorrAsyncGeom g,
              float *__restrict__ out) {
  using Cfg = TmaCfg<R, TH>;
  constexpr int WN = Cfg::WN, D2 = Cfg::D2, NB = PX + 2 * R;
  constexpr int kThr = kTW / (2 * PX) * 2 * TH;
  constexpr int kWarps = kThr / 32;
  constexpr int kWX = kTW / (8 * PX);                   // warps along x
  constexpr int kBPieces = (kTW + 2 * Cfg::HALO) * 2;   // 16-byte pieces per B tile row
  constexpr int kAPieces = kTW * 2;
  constexpr int kRowThreads = kThr / 16;                // loader rows covered per pass
  constexpr int kLoadRows = Cfg::BH + TH;               // B rows then A rows
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t smem_base = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_chunks = g.C / kCC;
  const int my_tiles = (g.n_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) /
                       static_cast<int>(gridDim.x);
  const int n_units = my_tiles * n_chunks;

  // ---- loader role: thread (ly, lx) copies the 16-byte pieces lx, lx+16, ... of load rows
  // ly, ly+kRowThreads, ...; a load row is a B tile row (first BH rows) or an A tile row. All
  // copies of a row share one global and one shared base and differ by compile-time offsets;
  // out-of-image pieces use src-size 0 (zero fill = the reference's padding).
  const int ly = threadIdx.x >> 4, lx = threadIdx.x & 15;
  const int lhalf = lx & 1, lpx = lx >> 1;

  // Per-unit loader state: one (global base, shared base, row-valid, first column) per load row.
  constexpr int kRowsPerThread = (kLoadRows + kRowThreads - 1) / kRowThreads;
  constexpr int kPiecesPerRow = (kBPieces + 15) / 16;
  constexpr int kSlices = 2 * WN;                         // one slice per (half, p) math step
  constexpr int kPiecesPerSlice = (kRowsPerThread * kPiecesPerRow + kSlices - 1) / kSlices;
  const float *l_src[kRowsPerThread];
  uint32_t l_dst[kRowsPerThread];
  int l_gx0[kRowsPerThread];      // first column, or a value that fails every bounds test
  int l_pieces[kRowsPerThread];

  auto prepare = [&](int u) {
    const int tile = blockIdx.x + (u / n_chunks) * gridDim.x;
    const int c0 = (u % n_chunks) * kCC + lhalf * 4;
    // batch-interleaved tile order: the same spatial tile of consecutive batch items is worked on
    // at the same time (by neighbouring CTAs), so a map that two items share is fetched once
    const int n = tile % g.batch, sp = tile / g.batch;
    const int tx = sp % g.tiles_x;
    const int ty = sp / g.tiles_x;
    const uint32_t sbase = smem_base + static_cast<uint32_t>(u % NST) * Cfg::STAGE1_OFF;
    const size_t img = static_cast<size_t>(n) * g.H * g.W * g.C;
    const float *a_img = g.n_stream ? g.maps[n] : a + img;
    const float *b_img = g.n_stream ? g.maps[n + 1] : b + img;
#pragma unroll
    for (int rr = 0; rr < kRowsPerThread; ++rr) {
      const int lr = ly + rr * kRowThreads;
      const bool is_b = lr < Cfg::BH;
      const int trow = is_b ? lr : lr - Cfg::BH;
      const int halo = is_b ? Cfg::HALO : 0;
      const int gy = ty * TH + g.shift - halo + trow;
      const int gx0 = tx * kTW + g.shift - halo + lpx;
      const bool row_ok = static_cast<unsigned>(gy) < static_cast<unsigned>(g.H);
      l_src[rr] = (is_b ? b_img : a_img) + (static_cast<long long>(gy) * g.W + gx0) * g.C + c0;
      l_dst[rr] = sbase + (is_b ? Cfg::A_BYTES : 0) +
                  static_cast<uint32_t>(swz(trow * (is_b ? Cfg::BW : Cfg::AW) + lpx, lhalf)) * 4u;
      l_gx0[rr] = row_ok ? gx0 : (1 << 30);
      l_pieces[rr] = lr >= kLoadRows ? 0 : (is_b ? kBPieces : kAPieces);
    }
  };
  // copies number slice*kPiecesPerSlice ... of this thread's (row, piece) list
  auto issue_slice = [&](int slice) {
#pragma unroll
    for (int k = 0; k < kPiecesPerSlice; ++k) {
      const int idx = slice * kPiecesPerSlice + k;
      const int rr = idx / kPiecesPerRow, j = idx % kPiecesPerRow;
      if (rr < kRowsPerThread && !(DBG & 2)) {
        if (lx + 16 * j < l_pieces[rr]) {
          const bool ok = static_cast<unsigned>(l_gx0[rr] + 8 * j) < static_cast<unsigned>(g.W);
          const int bytes = ok ? 16 : 0;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(l_dst[rr] + j * 8 * kCC * 4),
                       "l"(l_src[rr] + j * 8 * g.C), "r"(bytes)
                       : "memory");
        }
      }
    }
  };
  auto issue = [&](int u) {
    prepare(u);
#pragma unroll
    for (int sl = 0; sl < kSlices; ++sl) issue_slice(sl);
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  // ---- compute role: lane bits [0] parity, [1..2] row & 3, [3..4] group & 3
  const int row = (warp / kWX) * 4 + ((lane >> 1) & 3);
  const int x0 = ((warp % kWX) * 4 + (lane >> 3)) * (2 * PX) + (lane & 1);
  int aoff[PX], boff[NB];
#pragma unroll
  for (int j = 0; j < PX; ++j) aoff[j] = swz(row * Cfg::AW + x0 + 2 * j, 0);
#pragma unroll
  for (int q = 0; q < NB; ++q) boff[q] = swz(row * Cfg::BW + x0 + 2 * q, 0);

#pragma unroll
  for (int u0 = 0; u0 < NST - 1; ++u0) {
    if (u0 < n_units) issue(u0);
    else asm volatile("cp.async.commit_group;" ::: "memory");   // keep the group count uniform
  }
  int it = 0;
  for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x) {
    float acc[PX][D2];
#pragma unroll
    for (int j = 0; j < PX; ++j)
#pragma unroll
      for (int k = 0; k < D2; ++k) acc[j][k] = 0.0f;

    int stage = 0;
    for (int ch = 0; ch < n_chunks; ++ch, ++it) {
      stage = it % NST;
      asm volatile("cp.async.wait_group %0;" ::"n"(NST - 2) : "memory");
      __syncthreads();   // unit `it` has landed for everyone; everyone is done with unit it-1
      const bool more = it + NST - 1 < n_units;
      if (more) prepare(it + NST - 1);
      const float *sa = reinterpret_cast<const float *>(smem + stage * Cfg::STAGE1_OFF);
      const float *sb = sa + Cfg::A_BYTES / 4;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        if (DBG & 1) {   // diagnostic: no shared-memory reads, no FMAs; the copies still ride along
#pragma unroll
          for (int p = 0; p < WN; ++p)
            if (more) issue_slice(half * WN + p);
          if (half == 0) acc[0][0] += sa[threadIdx.x];
          continue;
        }
        float4 va[PX];
#pragma unroll
        for (int j = 0; j < PX; ++j)
          va[j] = *reinterpret_cast<const float4 *>(sa + (aoff[j] ^ (half << 2)));
#pragma unroll
        for (int p = 0; p < WN; ++p) {
          float4 vb[NB];
          // rows advance by 2*BW = 148 pixels: bit 2 of the pixel index (the swizzle bit) flips
          // with every p, the rest of the offset is a compile-time constant
#pragma unroll
          for (int q = 0; q < NB; ++q)
            vb[q] = *reinterpret_cast<const float4 *>(
                sb + (boff[q] ^ (((p + half) & 1) << 2)) + p * 2 * Cfg::BW * kCC);
          // a few of the next unit's cp.async copies ride along with every math step, so the
          // LSU queue never sees a burst and the copies' latency hides behind the FMAs
          if (FRONT) {
            // all copies of the next unit go out during the first half of this unit's math, so
            // the last of them has half a unit of FMAs to land behind
            if (more && half == 0) { issue_slice(2 * p); issue_slice(2 * p + 1); }
          } else {
            if (more) issue_slice(half * WN + p);
          }
#pragma unroll
          for (int j = 0; j < PX; ++j)
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
      asm volatile("cp.async.commit_group;" ::: "memory");
    }

    // ---- epilogue: stage the tile through the just-consumed stage (+ spare), coalesced stores
    const int n = tile % g.batch, sp = tile / g.batch;
    const int tx = sp % g.tiles_x;
    const int ty = sp / g.tiles_x;
    // DBG & 16: bulk-store epilogue (rows 16-byte aligned in shared memory: pitch +4 instead of +2)
    constexpr int OUT_PITCH = (DBG & 16) ? kTW * D2 + 4 : Cfg::OUT_PITCH;
    constexpr int OUT_BYTES = (DBG & 16) ? TH * OUT_PITCH * 4 : Cfg::OUT_BYTES;
    static_assert(OUT_BYTES <= Cfg::STAGE_BYTES + Cfg::SPARE && OUT_BYTES % 16 == 0, "staging fits a stage");
    float *stg = reinterpret_cast<float *>(
        smem + (stage == NST - 1 ? stage * Cfg::STAGE1_OFF + Cfg::STAGE_BYTES - OUT_BYTES
                                 : stage * Cfg::STAGE1_OFF));
    __syncthreads();  // everyone is done reading this stage
    if (g.pow2) {
#pragma unroll
      for (int j = 0; j < PX; ++j) {
        float *dst = stg + row * OUT_PITCH + (x0 + 2 * j) * D2;
#pragma unroll
        for (int k = 0; k < D2; ++k) dst[k] = __fmul_rn(acc[j][k], g.inv_c);   // exact: 1/2^k
      }
    } else {
      const float sumelems = static_cast<float>(g.C);
#pragma unroll
      for (int j = 0; j < PX; ++j) {
        float *dst = stg + row * OUT_PITCH + (x0 + 2 * j) * D2;
#pragma unroll
        for (int k = 0; k < D2; ++k) dst[k] = __fdiv_rn(acc[j][k], sumelems);
      }
    }
    __syncthreads();
    const int valid_rows = min(TH, g.out_h - ty * TH);
    const int valid_floats = min(kTW, g.out_w - tx * kTW) * D2;
    float *out_img = g.n_stream ? g.outs[n] : out + static_cast<size_t>(n) * g.out_h * g.out_w * D2;
    float *gout = out_img + (static_cast<size_t>(ty * TH) * g.out_w + tx * kTW) * D2;
    const bool vec_ok = (reinterpret_cast<uintptr_t>(out_img) % 8 == 0) && ((g.out_w * D2) % 2 == 0) &&
                        (valid_floats % 2 == 0);
    const bool bulk_ok = (DBG & 16) && reinterpret_cast<uintptr_t>(out_img) % 16 == 0 &&
                         (g.out_w * D2) % 4 == 0 && valid_floats % 4 == 0;
    if (bulk_ok) {
      // one bulk copy (TMA, shared -> global) per tile row, issued by the first valid_rows threads;
      // the issuing threads wait until the engine has READ the staging rows, the next unit's
      // __syncthreads then keeps the refill of this stage behind that
      if (static_cast<int>(threadIdx.x) < valid_rows) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        const uint32_t src = smem_u32(stg + threadIdx.x * OUT_PITCH);
        float *dstrow = gout + static_cast<size_t>(threadIdx.x) * g.out_w * D2;
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dstrow), "r"(src),
                     "r"(valid_floats * 4)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      }
    }
    for (int r = warp; r < (((DBG & 4) || bulk_ok) ? 0 : valid_rows); r += kWarps) {
      const float *src = stg + r * OUT_PITCH;
      float *dstrow = gout + static_cast<size_t>(r) * g.out_w * D2;
      if (vec_ok) {
        for (int e = lane; e < valid_floats / 2; e += 32)
          reinterpret_cast<float2 *>(dstrow)[e] = reinterpret_cast<const float2 *>(src)[e];
      } else {
        for (int e = lane; e < valid_floats; e += 32) dstrow[e] = src[e];
      }
    }
    // the next iteration's __syncthreads orders these shared-memory reads before the cp.async
    // writes that refill this stage
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                  const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
    else
      (void)cudaGetLastError();
  }
  return fn;
}

bool make_map(CUtensorMap *map, const float *ptr, int N, int H, int W, int C, int box_w, int box_h) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  const cuuint64_t dims[4] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(W),
                              static_cast<cuuint64_t>(H), static_cast<cuuint64_t>(N)};
  const cuuint64_t strides[3] = {static_cast<cuuint64_t>(C) * 4, static_cast<cuuint64_t>(W) * C * 4,
                                 static_cast<cuuint64_t>(H) * W * C * 4};
  const cuuint32_t box[4] = {kCC, static_cast<cuuint32_t>(box_w), static_cast<cuuint32_t>(box_h), 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float *>(ptr), dims, strides, box,
             estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B,
             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int R>
int launch(const float *a, const float *b, int N, int H, int W, int C, int out_h, int out_w,
           int shift, float *out, cudaStream_t stream) {
  using Cfg = TmaCfg<R>;
  alignas(64) CUtensorMap map_a, map_b;
  if (!make_map(&map_a, a, N, H, W, C, Cfg::AW, kTH) || !make_map(&map_b, b, N, H, W, C, Cfg::BW, Cfg::BH))
    return 1;  // not applicable (driver without tensor maps): caller falls back
  // the attribute belongs to the (function, device) pair: set on every launch (cheap)
  DODT_CUDA_TRY(cudaFuncSetAttribute(corr_tma_k1<R>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     Cfg::SMEM_BYTES));
  CorrTmaGeom g;
  g.batch = N; g.C = C; g.out_h = out_h; g.out_w = out_w; g.shift = shift;
  g.tiles_x = ceil_div(out_w, kTW);
  g.tiles_y = ceil_div(out_h, kTH);
  g.n_tiles = g.tiles_x * g.tiles_y * N;
  g.inv_unused = 0.f;
  const int grid = g.n_tiles < kNumSMs ? g.n_tiles : kNumSMs;
  corr_tma_k1<R><<<grid, kThreads, Cfg::SMEM_BYTES, stream>>>(map_a, map_b, g, out);
  DODT_AFTER_LAUNCH();
  return DODT_OK;
}

template <int R, int PX, int FRONT, int TH = kTH, int CTAS = 1, int NST = 2, int DBG = 0>
int launch_async(const float *a, const float *b, int N, int H, int W, int C, int out_h, int out_w,
                 int shift, float *out, int max_ctas, cudaStream_t stream,
                 const float *const *maps = nullptr, float *const *outs = nullptr) {
  using Cfg = TmaCfg<R, TH>;
  constexpr int kThr = kTW / (2 * PX) * 2 * TH;
  constexpr int smem_bytes = (NST - 1) * Cfg::STAGE1_OFF + Cfg::STAGE_BYTES + 64;
  static_assert(NST == 2 || Cfg::SPARE == 0, "the staging spare region is laid out for two stages");
  // the attribute belongs to the (function, device) pair: set on every launch (cheap)
  DODT_CUDA_TRY(cudaFuncSetAttribute(corr_async_k1<R, PX, FRONT, TH, CTAS, NST, DBG>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  CorrAsyncGeom g;
  g.batch = N; g.H = H; g.W = W; g.C = C; g.out_h = out_h; g.out_w = out_w; g.shift = shift;
  g.tiles_x = ceil_div(out_w, kTW);
  g.tiles_y = ceil_div(out_h, TH);
  g.n_tiles = g.tiles_x * g.tiles_y * N;
  g.pow2 = (C & (C - 1)) == 0 ? 1 : 0;
  g.inv_c = 1.0f / static_cast<float>(C);
  g.n_stream = 0;
  for (int k = 0; k <= kMaxStreamPairs; ++k) g.maps[k] = nullptr;
  for (int k = 0; k < kMaxStreamPairs; ++k) g.outs[k] = nullptr;
  if (maps) {   // frame-stream mode: N pairs over N + 1 maps
    g.n_stream = N;
    for (int k = 0; k <= N; ++k) g.maps[k] = maps[k];
    for (int k = 0; k < N; ++k) g.outs[k] = outs[k];
  }
  int grid = g.n_tiles < CTAS * kNumSMs ? g.n_tiles : CTAS * kNumSMs;
  if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;   // persistent CTAs: any count works
  corr_async_k1<R, PX, FRONT, TH, CTAS, NST, DBG><<<grid, kThr, smem_bytes, stream>>>(a, b, g, out);
  DODT_AFTER_LAUNCH();
  return DODT_OK;
}

}  // namespace

// returns DODT_OK if launched, 1 if this path does not apply, DODT_E* on failure.
// impl: 0 = cp.async pipeline (default), 1 = tensor-TMA pipeline (kept for A/B measurements)
int correlation_tma(const float *a, const float *b, int N, int H, int W, int C, int r, int out_h,
                    int out_w, int shift, float *out, int max_ctas, cudaStream_t stream) {
  if (C % kCC != 0 || reinterpret_cast<uintptr_t>(a) % 16 || reinterpret_cast<uintptr_t>(b) % 16)
    return 1;
  if (static_cast<long long>(out_h) * out_w < 1024) return 1;  // tiny maps: tiles mostly padding
#ifdef DODT_DIAG
  // A/B instantiations and diagnostic builds (parts of the kernel compiled out: results INVALID for
  // DBG 1/2/4/5/6) exist in libdodt_fe_diag.so only; the product library has neither the
  // instantiations nor any environment lookup.
  static int impl = -1;
  if (impl < 0) impl = DODT_KNOB("DODT_CORR_IMPL", 0);
  if (impl == 1) {   // tensor-TMA pipeline with 32-byte box rows
    switch (r) {
      case 1: return launch<1>(a, b, N, H, W, C, out_h, out_w, shift, out, stream);
      case 2: return launch<2>(a, b, N, H, W, C, out_h, out_w, shift, out, stream);
      default: return 1;
    }
  }
  if (impl == 2 && r == 2) return launch_async<2, 2, 1>(a, b, N, H, W, C, out_h, out_w, shift, out, max_ctas, stream);
  if (impl == 3 && r == 2) return launch_async<2, 4, 1>(a, b, N, H, W, C, out_h, out_w, shift, out, max_ctas, stream);
  if (impl == 6 && r == 2) return launch_async<2, 4, 0, 8, 1, 3>(a, b, N, H, W, C, out_h, out_w, shift, out, max_ctas, stream);
  if (impl == 4 && r == 2) return launch_async<2, 4, 0>(a, b, N, H, W, C, out_h, out_w, shift, out, max_ctas, stream);
  {
    static int dbg = -1;
    if (dbg < 0) dbg = DODT_KNOB("DODT_CORR_DBG", 0);
    if (r == 2 && dbg == 1) return launch_async<2, 4, 0, 8, 2, 2, 1>(a, b, N, H, W, C, out_h, out_w, shift, out, max_ctas, stream);
    if (r == 2 && dbg == 2) return launch_async<2, 4, 0, 8, 2, 2, 2>(a, b, N, H, W, C, out_h, out_w, shift, out, max_ctas, stream);
    if (r == 2 && dbg == 4) return launch_async<2, 4, 0, 8, 2, 2, 4>(a, b, N, H, W, C, out_h, out_w, shift, out, max_ctas, stream);
    if (r == 2 && dbg == 5) return launch_async<2, 4, 0, 8, 2, 2, 5>(a, b, N, H, W, C, out_h, out_w, shift, out, max_ctas, stream);
    if (r == 2 && dbg == 6) return launch_async<2, 4, 0, 8, 2, 2, 6>(a, b, N, H, W, C, out_h, out_w, shift, out, max_ctas, stream);
    if (r == 2 && dbg == 16) return launch_async<2, 4, 0, 8, 2, 2, 16>(a, b, N, H, W, C, out_h, out_w, shift, out, max_ctas, stream);
  }
#endif
  // default: 8-row tiles, 4 warps per CTA, TWO CTAs per SM — the per-unit barrier, the exposed
  // tail of the copies and the epilogue of one CTA hide behind the other CTA's math
  switch (r) {
    case 1: return launch_async<1, 4, 0, 8, 2>(a, b, N, H, W, C, out_h, out_w, shift, out, max_ctas, stream);
    case 2: return launch_async<2, 4, 0, 8, 2>(a, b, N, H, W, C, out_h, out_w, shift, out, max_ctas, stream);
    default: return 1;
  }
}

// Frame-stream form: pair j correlates maps[j] (frame t) with maps[j + 1] (frame t + 1) into
// outs[j], all pairs in ONE launch with batch-interleaved tiles, so that a map shared by two pairs
// (B of pair j, A of pair j + 1) is read from HBM once and served from L2 the second time.
// maps / outs: host arrays of device pointers. Returns like correlation_tma.
int correlation_stream_tma(const float *const *maps, int n_pairs, float *const *outs, int H, int W,
                           int C, int r, int out_h, int out_w, int shift, int max_ctas,
                           cudaStream_t stream) {
  if (n_pairs < 1 || n_pairs > kMaxStreamPairs || C % kCC != 0) return 1;
  for (int k = 0; k <= n_pairs; ++k)
    if (reinterpret_cast<uintptr_t>(maps[k]) % 16) return 1;
  if (static_cast<long long>(out_h) * out_w < 1024) return 1;
#ifdef DODT_DIAG
  static int dbg = -1;
  if (dbg < 0) dbg = DODT_KNOB("DODT_CORR_DBG", 0);
  if (r == 2 && dbg == 16)
    return launch_async<2, 4, 0, 8, 2, 2, 16>(nullptr, nullptr, n_pairs, H, W, C, out_h, out_w, shift, nullptr,
                                              max_ctas, stream, maps, outs);
#endif
  switch (r) {
    case 1: return launch_async<1, 4, 0, 8, 2>(nullptr, nullptr, n_pairs, H, W, C, out_h, out_w, shift,
                                               nullptr, max_ctas, stream, maps, outs);
    case 2: return launch_async<2, 4, 0, 8, 2>(nullptr, nullptr, n_pairs, H, W, C, out_h, out_w, shift,
                                               nullptr, max_ctas, stream, maps, outs);
    default: return 1;
  }
}

}  // namespace dodt
