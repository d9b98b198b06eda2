// S5 — greedy axis-aligned NMS with tf.image.non_max_suppression semantics, sm_100a.
//
// Semantics follow TensorFlow 1.3.0 core/kernels/non_max_suppression_op.cc, the op the reference
// calls at avod/core/models/dt_rpn_model.py:587-591 (RPN: IoU 0.8, 1024/300 outputs) and
// avod/core/models/dt_avod_model.py:609-613 (final: IoU 0.01, 100 outputs):
//   - candidates are visited in order of decreasing score;
//   - corners are normalised with min/max, area = (ymax-ymin)*(xmax-xmin), IoU = 0 if either area
//     is <= 0, else inter / (area_i + area_j - inter), all in fp32 with IEEE division;
//   - a candidate is dropped iff its IoU with an already selected box is > iou_threshold;
//   - selection stops at min(max_output_size, n).
// Equal scores are visited in ascending index order (a stable sort; TF's std::sort leaves that
// unspecified).
//
// Greedy NMS is a sequential recurrence. It is evaluated here in windows of kWin score-sorted
// candidates, one kernel launch per window:
//   phase 1 (all CTAs)   64x64 tiles of pairwise IoU tests. For window candidate i the lanes of a
//                        warp test 32 earlier candidates at a time and __ballot_sync packs the
//                        results into the candidate's "suppressor" bitmask (bit j set iff j < i and
//                        IoU(j, i) > thr); tiles against the boxes kept by earlier windows reduce
//                        to one "dead on arrival" bit per candidate.
//   phase 2 (last CTA to finish, classic threadfence hand-off — no CTA ever waits on another)
//                        pulls the triangular bitmask into shared memory and solves the recurrence
//                        by monotone relaxation: a candidate is KEPT once every suppressor is
//                        known-removed, REMOVED once any suppressor is known-kept. Each sweep
//                        decides at least the first undecided candidate, typical data needs a
//                        handful of sweeps, and the fixed point is exactly the greedy result.
//                        The kept prefix is cut at max_out, appended to the output and to the
//                        kept-box list that later windows test against.
// Windows after the one that completes the selection exit immediately.
//
// Score order without a full sort. NMS consumes candidates in score order but (with max_out of
// 1024 / 100) almost never looks past the first few thousand of the 10^4..10^5 candidates, so the
// candidates are ordered lazily in chunks of kChunk = 2 windows:
//   nms_select       one thread-block CLUSTER of 8 CTAs finds, by MSB-first radix selection on the
//                    64-bit key (descending score, ascending index), the threshold below which
//                    exactly the next kChunk candidates lie, and compacts them. Score keys stay
//                    in registers; the 256-bin digit histograms live in each CTA's shared memory
//                    and are combined through distributed shared memory, one cluster barrier per
//                    digit; with tie-free scores the four score digits settle the threshold.
//   nms_rank_gather  orders the chunk by counting: 16 lanes per candidate count how many of the
//                    chunk's keys are smaller (keys staged in shared memory), which is the
//                    candidate's rank; the lane group then writes the candidate's index, its
//                    corner-normalised box and area at that rank. No sort network, no passes.
// Sets of up to kChunk candidates (the final NMS) skip the selection.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace dodt {
namespace {

constexpr int kWin = DODT_NMS_WINDOW;  // candidates per window (1536)
constexpr int kWords = kWin / 64;      // 24 bitmask words per window
constexpr int kTriWords = 64 * (kWords * (kWords + 1) / 2);  // triangular suppressor store
constexpr int kRoundThreads = 1024;
constexpr int kResolveThreads = kRoundThreads;
constexpr int kPerThread = (kWin + kResolveThreads - 1) / kResolveThreads;  // candidates per thread
constexpr int kChunkWins = 2;                // windows per lazily ordered chunk
constexpr int kChunk = kChunkWins * kWin;    // 3072 candidates
constexpr int kSelCluster = 8;               // CTAs per selection cluster
constexpr int kSelThreads = 1024;
constexpr int kSelBins = 256;                // 8-bit digits
constexpr int kSelCache = 12;                // score keys a thread keeps in registers
constexpr int kRankThreads = 256;
constexpr int kRankLanes = 16;               // lanes that share one candidate's count
constexpr int kRankEach = 2;                 // candidates counted by one lane group (shared key loads)
constexpr int kRankPerCta = kRankEach * kRankThreads / kRankLanes;

struct NmsState {  // lives in the workspace, zeroed per call
  int n_kept;      // boxes selected so far
  int done;        // selection complete
  unsigned tiles_done;  // CTA completion ticket of the current round
  int sweeps;           // relaxation sweeps of the last solved window (diagnostic)
  unsigned long long t_ns[6];  // globaltimer at phase boundaries of the last solved window (diagnostic)
  unsigned long long bound;    // 64-bit keys below this are already ordered (end of the last chunk)
};

__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

struct NmsBox {
  float ymin, xmin, ymax, xmax;
};

__device__ __forceinline__ bool iou_exceeds(const NmsBox &a, float area_a, const NmsBox &b,
                                            float area_b, float thr) {
  float iou = 0.0f;
  if (area_a > 0.0f && area_b > 0.0f) {
    const float iy0 = fmaxf(a.ymin, b.ymin), ix0 = fmaxf(a.xmin, b.xmin);
    const float iy1 = fminf(a.ymax, b.ymax), ix1 = fminf(a.xmax, b.xmax);
    const float inter = __fmul_rn(fmaxf(__fsub_rn(iy1, iy0), 0.0f), fmaxf(__fsub_rn(ix1, ix0), 0.0f));
    if (inter > 0.0f)
      iou = __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_a, area_b), inter));
  }
  return iou > thr;
}

__device__ __forceinline__ unsigned desc_key(float f) {
  const unsigned u = __float_as_uint(f);
  const unsigned asc = (u & 0x80000000u) ? ~u : (u | 0x80000000u);  // ascending float order
  return ~asc;                                                       // smaller = higher score
}
// total order of the candidates: descending score, then ascending index (all keys distinct)
__device__ __forceinline__ unsigned long long cand_key(float score, int i) {
  return (static_cast<unsigned long long>(desc_key(score)) << 32) | static_cast<unsigned>(i);
}

// Chunk `chunk` = the candidates of rank [chunk*kChunk, (chunk+1)*kChunk) in key order. Finds the
// exclusive upper key bound T of the chunk (the lower bound is where the previous chunk ended) and
// writes the chunk's keys, unordered, to cand[0 .. min(kChunk, remaining)).
// CACHED: every thread keeps its (at most kSelCache) score keys in registers across the digits.
template <bool CACHED>
__global__ void __cluster_dims__(kSelCluster, 1, 1) __launch_bounds__(kSelThreads)
nms_select(const float *__restrict__ scores, int n_max, const int *__restrict__ n_dev, int chunk,
           NmsState *__restrict__ st, unsigned long long *__restrict__ cand) {
  __shared__ int hist[2][kSelBins];
  __shared__ int warp_tot[kSelBins / 32];
  __shared__ int s_bin, s_excl;
  __shared__ int s_count, s_pos;
  cg::cluster_group cluster = cg::this_cluster();
  const int crank = static_cast<int>(cluster.block_rank());
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (st->done) return;                                  // uniform over the cluster
  const int n = n_dev ? min(n_max, __ldg(n_dev)) : n_max;
  const int lo_rank = chunk * kChunk;
  if (lo_rank >= n) return;                              // uniform
  const unsigned long long lo_bound = st->bound;
  const int remaining = n - lo_rank;                     // candidates with key >= lo_bound
  constexpr int stride = kSelCluster * kSelThreads;
  const int first = crank * kSelThreads + tid;
  const int iters = (n + stride - 1) / stride;           // uniform trip count

  unsigned key[kSelCache];
  if (CACHED) {
#pragma unroll
    for (int e = 0; e < kSelCache; ++e) {
      const int i = first + e * stride;
      key[e] = (e < iters && i < n) ? desc_key(__ldg(scores + i)) : 0u;
    }
  }
  // 64-bit key of this thread's e-th candidate; false when there is none or it is below lo_bound
  auto fetch = [&](int e, unsigned long long *c) {
    const int i = first + e * stride;
    if (i >= n) return false;
    const unsigned k = CACHED ? key[e] : desc_key(__ldg(scores + i));
    *c = (static_cast<unsigned long long>(k) << 32) | static_cast<unsigned>(i);
    return *c >= lo_bound;
  };

  unsigned long long T = ~0ull;                          // exclusive upper bound of the chunk
  if (remaining > kChunk) {
    // MSB-first radix selection (8-bit digits) of the key of rank kChunk among the keys >=
    // lo_bound: after each digit the search continues inside one bin; it ends as soon as the
    // wanted key is the smallest of its bin (k_rem == 0), because then "prefix with zero low
    // bits" separates. Tie-free scores settle within the four score digits.
    int k_rem = kChunk;
    unsigned long long prefix = 0;
    int bits = 0;
    for (int p = 0; p < 8; ++p) {
      int *h = hist[p & 1];
      if (tid < kSelBins) h[tid] = 0;
      __syncthreads();
      const int shift = 56 - bits;
      // Scores of real candidates crowd into a few high-digit bins (floats of one octave share
      // sign, exponent and leading mantissa bits), so the lanes of a warp first combine equal
      // bins (match.any) and one lane per distinct bin adds the group's count.
#pragma unroll
      for (int e = 0; e < (CACHED ? kSelCache : 1); ++e) {
        for (int ee = e; ee < iters; ee += (CACHED ? iters : 1)) {   // CACHED: exactly once
          unsigned long long c = 0;
          bool part = fetch(ee, &c);
          part = part && (bits == 0 || (c >> (64 - bits)) == prefix);
          const int bin = static_cast<int>(c >> shift) & (kSelBins - 1);
          const unsigned active = __ballot_sync(0xffffffffu, part);
          if (part) {
            const unsigned peers = __match_any_sync(active, bin);
            if (lane == __ffs(peers) - 1) atomicAdd(&h[bin], __popc(peers));
          }
        }
      }
      cluster.sync();                                    // every CTA's histogram is complete
      // cluster-wide total of bin `tid` through distributed shared memory, then a block scan
      int t = 0;
      if (tid < kSelBins) {
#pragma unroll
        for (int r = 0; r < kSelCluster; ++r) t += cluster.map_shared_rank(h, r)[tid];
      }
      int incl = t;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int up = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += up;
      }
      if (tid < kSelBins && lane == 31) warp_tot[warp] = incl;
      __syncthreads();
      if (tid < kSelBins) {
        int before = 0;
        for (int w2 = 0; w2 < warp; ++w2) before += warp_tot[w2];
        const int excl = before + incl - t;
        if (k_rem >= excl && k_rem < excl + t) { s_bin = tid; s_excl = excl; }
      }
      __syncthreads();
      prefix = (prefix << 8) | static_cast<unsigned>(s_bin);
      bits += 8;
      k_rem -= s_excl;
      __syncthreads();                                   // s_bin / warp_tot are reused next digit
      if (k_rem == 0) break;                             // every CTA computes the same values
    }
    T = bits >= 64 ? prefix : prefix << (64 - bits);
  }

  // compaction of lo_bound <= key < T: count per CTA, offsets across the cluster, scatter
  if (tid == 0) { s_count = 0; s_pos = 0; }
  __syncthreads();
  int mine = 0;
#pragma unroll
  for (int e = 0; e < (CACHED ? kSelCache : 1); ++e)
    for (int ee = e; ee < iters; ee += (CACHED ? iters : 1)) {
      unsigned long long c = 0;
      mine += (fetch(ee, &c) && c < T) ? 1 : 0;
    }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) mine += __shfl_down_sync(0xffffffffu, mine, d);
  if (lane == 0 && mine) atomicAdd(&s_count, mine);
  cluster.sync();
  int offset = 0;
  for (int r = 0; r < crank; ++r) offset += *cluster.map_shared_rank(&s_count, r);
#pragma unroll
  for (int e = 0; e < (CACHED ? kSelCache : 1); ++e)
    for (int ee = e; ee < iters; ee += (CACHED ? iters : 1)) {
      unsigned long long c = 0;
      if (fetch(ee, &c) && c < T) cand[offset + atomicAdd(&s_pos, 1)] = c;
    }
  cluster.sync();                                        // nobody still reads this CTA's shared memory
  if (crank == 0 && tid == 0) st->bound = T;
}

// Orders one chunk by counting and writes index / normalised box / area at each candidate's rank.
// cand == nullptr: the chunk is the whole input (n <= kChunk), keys are built from the scores.
__global__ void __launch_bounds__(kRankThreads)
nms_rank_gather(const float *__restrict__ boxes, const float *__restrict__ scores,
                const unsigned long long *__restrict__ cand, int n_max,
                const int *__restrict__ n_dev, int chunk, const NmsState *__restrict__ st,
                int *__restrict__ order, NmsBox *__restrict__ sbox, float *__restrict__ sarea) {
  __shared__ unsigned long long s_key[kChunk];   // 24 KB
  if (st->done) return;
  const int n = n_dev ? min(n_max, __ldg(n_dev)) : n_max;
  const int lo_rank = chunk * kChunk;
  const int m = min(kChunk, n - lo_rank);
  if (static_cast<int>(blockIdx.x) * kRankPerCta >= m) return;
#pragma unroll 4
  for (int j = threadIdx.x; j < m; j += kRankThreads)
    s_key[j] = cand ? __ldcg(cand + j) : cand_key(__ldg(scores + j), j);
  __syncthreads();
  const int sub = threadIdx.x % kRankLanes;
  const int c0 = blockIdx.x * kRankPerCta + (threadIdx.x / kRankLanes) * kRankEach;
  unsigned long long me[kRankEach];
  int cnt[kRankEach];
#pragma unroll
  for (int e = 0; e < kRankEach; ++e) {
    me[e] = c0 + e < m ? s_key[c0 + e] : 0ull;   // key 0 never occurs: nothing is smaller
    cnt[e] = 0;
  }
#pragma unroll 8
  for (int j = sub; j < m; j += kRankLanes) {
    const unsigned long long k = s_key[j];
#pragma unroll
    for (int e = 0; e < kRankEach; ++e) cnt[e] += k < me[e] ? 1 : 0;
  }
#pragma unroll
  for (int e = 0; e < kRankEach; ++e)
#pragma unroll
    for (int d = kRankLanes / 2; d > 0; d >>= 1) cnt[e] += __shfl_xor_sync(0xffffffffu, cnt[e], d);
#pragma unroll
  for (int e = 0; e < kRankEach; ++e) {
    if (sub != e || c0 + e >= m) continue;      // lane e of the group writes candidate e
    const int pos = lo_rank + cnt[e];
    const int src = static_cast<int>(me[e] & 0xFFFFFFFFull);
    order[pos] = src;
    const float4 b = __ldg(reinterpret_cast<const float4 *>(boxes) + src);
    NmsBox o;
    o.ymin = fminf(b.x, b.z); o.xmin = fminf(b.y, b.w);
    o.ymax = fmaxf(b.x, b.z); o.xmax = fmaxf(b.y, b.w);
    sbox[pos] = o;
    sarea[pos] = __fmul_rn(__fsub_rn(o.ymax, o.ymin), __fsub_rn(o.xmax, o.xmin));
  }
}

// Triangular suppressor store: block bi (64 candidates) owns words 0..bi, laid out word-major
// ([w][64 candidates]) so that the 32 lanes of a warp, which solve 32 consecutive candidates, read
// 32 consecutive 8-byte words (a candidate-major layout strides by (bi+1)*8 bytes: 32-way bank
// conflicts at bi = 15).
__device__ __forceinline__ int tri_base(int i) {
  const int bi = i >> 6;
  return 64 * (bi * (bi + 1) / 2) + (i & 63);
}
__device__ __forceinline__ int tri_word(int i, int w) { return tri_base(i) + w * 64; }

constexpr int kTileThreads = 256;

// Phase 1 of a window: one CTA per 64x64 tile of pairwise IoU tests (small CTAs, no shared-memory
// footprint to speak of, so they share SMs with whatever else is running).
__global__ void __launch_bounds__(kTileThreads)
nms_tiles(const NmsBox *__restrict__ sbox, const float *__restrict__ sarea, int n_max,
          const int *__restrict__ n_dev, int base, float thr,
          unsigned long long *__restrict__ sup,      // [kTriWords] suppressor bitmasks
          unsigned *__restrict__ dead,               // [2*kWin/32]: killed by earlier windows,
                                                     // then "has a suppressor in this window";
                                                     // then [kWin]: which suppressor words of
                                                     // each candidate are non-zero
          const NmsBox *__restrict__ kbox, const float *__restrict__ karea,  // kept boxes so far
          const NmsState *__restrict__ st) {
  __shared__ NmsBox jb[64];
  __shared__ float ja[64];
  if (st->done) return;
  const int n = n_dev ? min(n_max, __ldg(n_dev)) : n_max;
  const int wcount = max(0, min(kWin, n - base));  // candidates in this window
  const int nb = (wcount + 63) >> 6;               // 64-blocks in this window
  const int n_prev = st->n_kept;                   // boxes kept by earlier windows
  const int pb = (n_prev + 63) >> 6;
  const int tri_tiles = nb * (nb + 1) / 2;
  const int n_tiles = tri_tiles + nb * pb;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kWarps = kTileThreads / 32;

  // ---------------- phase 1: IoU tiles ----------------
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    int bi, bj;          // candidate block, suppressor block
    bool vs_prev;
    if (tile < tri_tiles) {
      // tile -> (bi, bj) with bj <= bi
      bi = static_cast<int>((sqrtf(8.0f * tile + 1.0f) - 1.0f) * 0.5f);
      while (bi * (bi + 1) / 2 > tile) --bi;
      while ((bi + 1) * (bi + 2) / 2 <= tile) ++bi;
      bj = tile - bi * (bi + 1) / 2;
      vs_prev = false;
    } else {
      const int t2 = tile - tri_tiles;
      bi = t2 / pb;
      bj = t2 % pb;
      vs_prev = true;
    }
    __syncthreads();
    if (threadIdx.x < 64) {
      const int j = bj * 64 + threadIdx.x;
      if (vs_prev) {
        const bool ok = j < n_prev;
        jb[threadIdx.x] = ok ? kbox[j] : NmsBox{0.f, 0.f, 0.f, 0.f};
        ja[threadIdx.x] = ok ? karea[j] : 0.0f;
      } else {
        const bool ok = j < wcount;
        jb[threadIdx.x] = ok ? sbox[base + j] : NmsBox{0.f, 0.f, 0.f, 0.f};
        ja[threadIdx.x] = ok ? sarea[base + j] : 0.0f;   // area 0 never suppresses
      }
    }
    __syncthreads();
    // each warp takes candidates i = bi*64 + warp, + kWarps, ...; lanes take suppressors j
    for (int ii = warp; ii < 64; ii += kWarps) {
      const int i = bi * 64 + ii;
      if (i >= wcount) break;  // warp-uniform
      const NmsBox b = sbox[base + i];
      const float area = sarea[base + i];
      const int jn = (vs_prev ? n_prev : wcount) - bj * 64;  // valid suppressors in this tile
      bool h0 = lane < jn && iou_exceeds(jb[lane], ja[lane], b, area, thr);
      bool h1 = lane + 32 < jn && iou_exceeds(jb[lane + 32], ja[lane + 32], b, area, thr);
      if (!vs_prev && bj == bi) {  // diagonal tile: only earlier candidates suppress
        h0 = h0 && lane < ii;
        h1 = h1 && lane + 32 < ii;
      }
      const unsigned m0 = __ballot_sync(0xffffffffu, h0);
      const unsigned m1 = __ballot_sync(0xffffffffu, h1);
      if (lane == 0) {
        if (vs_prev) {
          if (m0 | m1) atomicOr(&dead[i >> 5], 1u << (i & 31));
        } else {
          if (m0 | m1) {
            sup[tri_word(i, bj)] = (static_cast<unsigned long long>(m1) << 32) | m0;
            atomicOr(&dead[kWin / 32 + (i >> 5)], 1u << (i & 31));
            atomicOr(&dead[2 * (kWin / 32) + i], 1u << bj);
          }
        }
      }
    }
  }

}

// Phase 2 of a window: one CTA solves the recurrence from the suppressor bitmasks and emits.
__global__ void __launch_bounds__(kRoundThreads)
nms_solve(const NmsBox *__restrict__ sbox, const float *__restrict__ sarea,
          const int *__restrict__ order, int n_max, const int *__restrict__ n_dev, int base,
          int max_out, const unsigned long long *__restrict__ sup, unsigned *__restrict__ dead,
          NmsBox *__restrict__ kbox, float *__restrict__ karea, NmsState *__restrict__ st,
          int *__restrict__ keep, int *__restrict__ n_keep) {
  extern __shared__ unsigned long long smem_sup[];   // [64 * nb*(nb+1)/2] suppressor words
  __shared__ unsigned long long s_kept[kWords], s_removed[kWords], s_hassup[kWords];
  __shared__ int s_prefix[kWords + 1];
  if (st->done) return;
  const unsigned long long t_start = global_ns();
  const int n = n_dev ? min(n_max, __ldg(n_dev)) : n_max;
  const int wcount = max(0, min(kWin, n - base));
  const int n_prev = st->n_kept;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { st->t_ns[0] = t_start; st->t_ns[1] = t_start; }

  // ---------------- phase 2: solve the window ----------------
  // only the suppressor words that hold a bit are fetched (at IoU 0.8 that is a few per cent of
  // the 225 KB triangle); a thread reads back only the words it fetched itself
  unsigned nz[kPerThread];
#pragma unroll
  for (int q = 0; q < kPerThread; ++q) {
    const int i = q * kResolveThreads + threadIdx.x;
    nz[q] = i < wcount ? __ldcg(dead + 2 * (kWin / 32) + i) : 0u;
    const int row = tri_base(i < kWin ? i : 0);
    for (unsigned m = nz[q]; m; m &= m - 1) {
      const int w = __ffs(m) - 1;
      smem_sup[row + w * 64] = __ldcg(sup + row + w * 64);
    }
  }
  if (threadIdx.x < kWords) {
    const int w = threadIdx.x;
    // bits beyond wcount and candidates killed by earlier windows start as removed
    unsigned long long rem = (static_cast<unsigned long long>(__ldcg(dead + 2 * w + 1)) << 32) |
                             __ldcg(dead + 2 * w);
    const int valid = wcount - w * 64;
    if (valid <= 0) rem = ~0ull;
    else if (valid < 64) rem |= ~0ull << valid;
    s_removed[w] = rem;
    // candidates nobody in this window can suppress are decided at once
    const unsigned long long hs = (static_cast<unsigned long long>(__ldcg(dead + kWin / 32 + 2 * w + 1)) << 32) |
                                  __ldcg(dead + kWin / 32 + 2 * w);
    s_hassup[w] = hs;
    s_kept[w] = ~rem & ~hs;
  }
  __syncthreads();

  if (threadIdx.x == 0) st->t_ns[2] = global_ns();
  bool undecided[kPerThread];
#pragma unroll
  for (int q = 0; q < kPerThread; ++q) undecided[q] = true;
  int sweeps = 0;
  while (true) {
    ++sweeps;
    int pending = 0;
    bool now_kept[kPerThread], now_removed[kPerThread];
#pragma unroll
    for (int q = 0; q < kPerThread; ++q) {
      const int i = q * kResolveThreads + threadIdx.x;
      now_kept[q] = false;
      now_removed[q] = false;
      if (!undecided[q]) continue;
      if (i >= wcount || (((s_removed[i >> 6] | ~s_hassup[i >> 6]) >> (i & 63)) & 1ull)) {
        undecided[q] = false;   // out of range, dead on arrival, or kept because unsuppressable
        continue;
      }
      const unsigned long long *row = smem_sup + tri_base(i);
      bool hit_kept = false, all_removed = true;
      for (unsigned m = nz[q]; m; m &= m - 1) {
        const int w = __ffs(m) - 1;
        const unsigned long long s = row[w * 64];
        if (s & s_kept[w]) { hit_kept = true; break; }
        if (s & ~s_removed[w]) all_removed = false;
      }
      now_removed[q] = hit_kept;
      now_kept[q] = !hit_kept && all_removed;
      if (hit_kept || all_removed) undecided[q] = false; else pending = 1;
    }
    // publish: a warp's 32 candidates of slot q are exactly one 32-bit half of a state word, so
    // one lane ORs the ballot in — no shared-memory atomics (64-bit ones are CAS loops)
#pragma unroll
    for (int q = 0; q < kPerThread; ++q) {
      const unsigned mk = __ballot_sync(0xffffffffu, now_kept[q]);
      const unsigned mr = __ballot_sync(0xffffffffu, now_removed[q]);
      const int i0 = q * kResolveThreads + (threadIdx.x & ~31);
      if (lane == 0 && i0 < kWin) {
        unsigned *k32 = reinterpret_cast<unsigned *>(s_kept) + (i0 >> 5);
        unsigned *r32 = reinterpret_cast<unsigned *>(s_removed) + (i0 >> 5);
        if (mk) *k32 |= mk;
        if (mr) *r32 |= mr;
      }
    }
    if (!__syncthreads_or(pending)) break;
  }

  // ---------------- emit: kept candidates in score order, cut at max_out ----------------
  if (threadIdx.x == 0) { st->t_ns[3] = global_ns(); st->sweeps = sweeps; }
  if (threadIdx.x == 0) {
    int run = 0;
    for (int w = 0; w < kWords; ++w) { s_prefix[w] = run; run += __popcll(s_kept[w]); }
    s_prefix[kWords] = run;
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < kPerThread; ++q) {
    const int i = q * kResolveThreads + threadIdx.x;
    if (i >= wcount) continue;
    const unsigned long long kw = s_kept[i >> 6];
    if (!((kw >> (i & 63)) & 1ull)) continue;
    const int pos = n_prev + s_prefix[i >> 6] + __popcll(kw & ((1ull << (i & 63)) - 1ull));
    if (pos < max_out) {
      keep[pos] = order[base + i];
      kbox[pos] = sbox[base + i];
      karea[pos] = sarea[base + i];
    }
  }
  // reset the per-round scratch for the next window
  for (int w = threadIdx.x; w < 2 * (kWin / 32) + kWin; w += kResolveThreads) dead[w] = 0u;
  __syncthreads();
  const int total = min(max_out, n_prev + s_prefix[kWords]);
  // the first window owns the initialisation of the outputs: unused entries of keep are -1
  if (base == 0)
    for (int k = total + threadIdx.x; k < max_out; k += kResolveThreads) keep[k] = -1;
  if (threadIdx.x == 0) {
    const bool complete = total >= max_out || base + wcount >= n;
    st->n_kept = total;
    st->t_ns[4] = global_ns();
    n_keep[0] = total;
    if (complete) st->done = 1;
    if (complete || base == 0) n_keep[1] = complete ? 1 : 0;   // selection complete?
  }
}

struct NmsLayout {
  size_t order, sbox, sarea, kbox, karea, cand, sup, dead, state, total;
};

size_t align_up(size_t v) { return (v + 255) & ~static_cast<size_t>(255); }

int nms_layout(int64_t n, NmsLayout *L) {
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes); return o; };
  const size_t nn = static_cast<size_t>(n > 0 ? n : 1);
  L->order = take(nn * 4);
  L->sbox = take(nn * sizeof(NmsBox));
  L->sarea = take(nn * 4);
  L->kbox = take(nn * sizeof(NmsBox));
  L->karea = take(nn * 4);
  L->cand = take(static_cast<size_t>(kChunk) * 8);
  L->sup = take(static_cast<size_t>(kTriWords) * 8);
  L->dead = take((2 * (kWin / 32) + kWin) * 4);
  L->state = take(sizeof(NmsState));
  L->total = off;
  return DODT_OK;
}

}  // namespace
}  // namespace dodt

extern "C" {

size_t dodt_nms_state_offset(int64_t n) {
  if (n < 0 || n > 0x7FFFFFFF) return 0;
  dodt::NmsLayout L;
  dodt::nms_layout(n, &L);
  return L.state;
}

size_t dodt_nms_workspace_bytes(int64_t n) {
  if (n < 0 || n > 0x7FFFFFFF) return 0;
  dodt::NmsLayout L;
  dodt::nms_layout(n, &L);
  return L.total;
}

int dodt_nms(const float *boxes, const float *scores, int64_t n, const int32_t *n_dev,
             int32_t max_out, float iou_threshold, int32_t first_window, int32_t max_windows,
             int32_t *keep,
             int32_t *n_keep, void *workspace, size_t workspace_bytes, dodt_stream_t stream_) {
  using namespace dodt;
  if (n < 0 || n > 0x7FFFFFFF || max_out < 0 || max_windows < 0 || !n_keep || (max_out > 0 && !keep))
    return DODT_EINVAL;
  if (first_window < 0 || first_window % dodt::kChunkWins != 0) return DODT_EINVAL;
  const bool resume = first_window > 0;   // continue on the state a previous call left in `workspace`
  if (n > 0 && (!boxes || !scores)) return DODT_EINVAL;
  if (reinterpret_cast<uintptr_t>(boxes) % 16 != 0) return DODT_EALIGN;
  cudaStream_t stream = as_stream(stream_);
  if (n == 0 || max_out == 0) {
    // nothing to select: complete. One byte of 0x01 on the zeroed little-endian int32 is 1.
    DODT_CUDA_TRY(cudaMemsetAsync(n_keep, 0, 2 * sizeof(int32_t), stream));
    if (max_out > 0) DODT_CUDA_TRY(cudaMemsetAsync(keep, 0xFF, sizeof(int32_t) * max_out, stream));
    DODT_CUDA_TRY(cudaMemsetAsync(n_keep + 1, 1, 1, stream));
    return DODT_OK;
  }
  // (keep and n_keep are initialised by the first window's solve kernel)
  NmsLayout L;
  nms_layout(n, &L);
  if (!workspace || workspace_bytes < L.total) return DODT_ECAPACITY;
  if (reinterpret_cast<uintptr_t>(workspace) % 256 != 0) return DODT_EALIGN;
  char *ws = static_cast<char *>(workspace);
  int *order = reinterpret_cast<int *>(ws + L.order);
  NmsBox *sbox = reinterpret_cast<NmsBox *>(ws + L.sbox);
  float *sarea = reinterpret_cast<float *>(ws + L.sarea);
  NmsBox *kbox = reinterpret_cast<NmsBox *>(ws + L.kbox);
  float *karea = reinterpret_cast<float *>(ws + L.karea);
  unsigned long long *cand = reinterpret_cast<unsigned long long *>(ws + L.cand);
  unsigned long long *sup = reinterpret_cast<unsigned long long *>(ws + L.sup);
  unsigned *dead = reinterpret_cast<unsigned *>(ws + L.dead);
  NmsState *st = reinterpret_cast<NmsState *>(ws + L.state);

  const int ni = static_cast<int>(n);
  // dead bits and state start at zero (one memset: they are adjacent up to alignment padding)
  if (!resume)
    DODT_CUDA_TRY(cudaMemsetAsync(ws + L.dead, 0, (L.state - L.dead) + sizeof(NmsState), stream));

  const size_t smem = static_cast<size_t>(kTriWords) * sizeof(unsigned long long);
  // the attribute belongs to the (function, device) pair: set on every launch (cheap)
  DODT_CUDA_TRY(cudaFuncSetAttribute(nms_solve, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     static_cast<int>(smem)));
  const bool lazy = ni > kChunk;   // more than one chunk: order the candidates chunk by chunk
  int windows = 0;
  for (int chunk = first_window / kChunkWins; chunk * kChunk < ni; ++chunk) {
    if (max_windows > 0 && windows >= max_windows) break;
    const int cbase = chunk * kChunk;
    const int m = ni - cbase < kChunk ? ni - cbase : kChunk;
    if (lazy) {
      if (ni <= kSelCache * kSelCluster * kSelThreads)
        nms_select<true><<<kSelCluster, kSelThreads, 0, stream>>>(scores, ni, n_dev, chunk, st, cand);
      else
        nms_select<false><<<kSelCluster, kSelThreads, 0, stream>>>(scores, ni, n_dev, chunk, st, cand);
      DODT_AFTER_LAUNCH();
    }
    nms_rank_gather<<<ceil_div(m, kRankPerCta), kRankThreads, 0, stream>>>(
        boxes, scores, lazy ? cand : nullptr, ni, n_dev, chunk, st, order, sbox, sarea);
    DODT_AFTER_LAUNCH();
    for (int base = cbase; base < cbase + m; base += kWin) {
      if (max_windows > 0 && windows >= max_windows) break;
      ++windows;
      const int wcount = ni - base < kWin ? ni - base : kWin;
      const int nb = (wcount + 63) / 64;
      const int pb = (max_out + 63) / 64;  // upper bound of kept blocks from earlier windows
      const int tiles = nb * (nb + 1) / 2 + (base > 0 ? nb * pb : 0);
      // later windows usually find the selection complete and exit at once: launch them narrower
      const int cap = base == 0 ? tiles : 2 * kNumSMs;
      nms_tiles<<<tiles < cap ? tiles : cap, kTileThreads, 0, stream>>>(
          sbox, sarea, ni, n_dev, base, iou_threshold, sup, dead, kbox, karea, st);
      DODT_AFTER_LAUNCH();
      const size_t need = static_cast<size_t>(64) * (nb * (nb + 1) / 2) * sizeof(unsigned long long);
      nms_solve<<<1, kRoundThreads, need, stream>>>(sbox, sarea, order, ni, n_dev, base, max_out, sup,
                                                    dead, kbox, karea, st, keep, n_keep);
      DODT_AFTER_LAUNCH();
    }
  }
  return DODT_OK;
}

}  // extern "C"
