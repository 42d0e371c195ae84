// grid_build.cu -- 2-D NDT grid construction on sm_100a: a scatter (count + bucket) followed by an
// in-input-order per-cell reduction and the regularised inverse covariance.
//
// Replaces pcl::VoxelGridCovariance::applyFilter as run by ndt.setInputTarget(target_cloud)
// [REF src/PoseEstimator.cpp:19; resolution set at include/ndt_slam/PoseEstimator.h:81].
// Specification: SURVEY.md App. A.2.
//
// This translation unit is compiled with --fmad=false: every fp64 expression below is evaluated
// as separate IEEE multiply / add / divide / sqrt, in the same order as the generic x86-64 code path
// of the reference stack (and as the CPU oracle used by the tests), so counts, means, centroids and inverse
// covariances come out bit-identical rather than merely within tolerance.
//
// Pipeline (all on the handle's stream; host round trips: the bounds read-back when the cloud is large or already on the
// device -- small host clouds get their bounds during staging -- and one read-back of the leaf / slot counters):
//   k_bounds       float min/max + finite count, one set of global atomics per CTA
//   memsets        per-cell counts = 0, probe table = NaN (all-ones), occupancy bitmap = 0; the slot table is never cleared
//   k_count        cell id per point (float32, bit-exact), warp-aggregated int atomics -> count + arbitrary rank
//   k_alloc        occupied cells only: leaf ids and contiguous bucket ranges (two sweeps, one pair of atomics per CTA),
//                  count -> leaf id + 1 in place, list of dense leaves
//   k_fill, k_rank scatter point indices into their bucket, then order every bucket by point index (rank by counting)
//   k_tile_hist / k_tile_scan / k_tile_place   instead of count / fill / rank for small grids with dense buckets: stable
//                  counting sort by cell over tiles of the input (linear, not quadratic in the bucket size)
//   k_finalize     walk every bucket in input order (fp32 centroid, fp64 sums): one thread per sparse leaf, one warp per
//                  dense leaf; mean, single-pass covariance, 2x2 eigen clamp, inverse, 64-byte record, probe / slot /
//                  dilated-occupancy entries
//   k_nn_*         finer lattice for the exact 1-NN of the fitness score when the NDT buckets are dense
// Batched scan pairs: k_pair_bounds + k_pair_scan compute every pair's geometry on the device, then one pass of the
// same kernels builds all grids in one shared padded table.
#include "ndt_host.h"

#include <cooperative_groups.h>
#include <climits>
#include <cmath>
#include <cstring>
#include <limits>
#include <vector>

namespace cg = cooperative_groups;

namespace ndt {

namespace {

__device__ __forceinline__ int float_ord(float f) {
  const int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ord_to_float(int o) { return __int_as_float(o >= 0 ? o : o ^ 0x7fffffff); }
inline float ord_float(int o) {
  int i = o >= 0 ? o : o ^ 0x7fffffff;
  float f;
  std::memcpy(&f, &i, 4);
  return f;
}

__global__ void k_init_counters(int32_t *ctr, int32_t *bounds) {
  const int t = threadIdx.x;
  if (t < CTR_COUNT) ctr[t] = 0;
  if (t == 0) { bounds[0] = INT_MAX; bounds[1] = INT_MAX; bounds[2] = INT_MIN; bounds[3] = INT_MIN; }
}

// getMinMax3D over finite points
__global__ void __launch_bounds__(256) k_bounds(const float4 *__restrict__ pts, int64_t n, int32_t *bounds,
                                               int32_t *ctr) {
  int mnx = INT_MAX, mny = INT_MAX, mxx = INT_MIN, mxy = INT_MIN, nf = 0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += 4 * stride) {
    float4 p[4];                                    // four independent 16-byte loads in flight per thread
#pragma unroll
    for (int u = 0; u < 4; ++u) p[u] = (i + u * stride < n) ? __ldg(pts + i + u * stride) : make_float4(NAN, NAN, NAN, 0.f);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (isfinite(p[u].x) && isfinite(p[u].y) && isfinite(p[u].z)) {
        const int ox = float_ord(p[u].x), oy = float_ord(p[u].y);
        mnx = min(mnx, ox); mny = min(mny, oy); mxx = max(mxx, ox); mxy = max(mxy, oy);
        ++nf;
      }
    }
  }
  mnx = __reduce_min_sync(0xffffffffu, mnx); mny = __reduce_min_sync(0xffffffffu, mny);
  mxx = __reduce_max_sync(0xffffffffu, mxx); mxy = __reduce_max_sync(0xffffffffu, mxy);
  nf = __reduce_add_sync(0xffffffffu, nf);
  // one set of global atomics per CTA, not per warp: they all hit the same five addresses
  __shared__ int s_b[5];
  if (threadIdx.x == 0) { s_b[0] = INT_MAX; s_b[1] = INT_MAX; s_b[2] = INT_MIN; s_b[3] = INT_MIN; s_b[4] = 0; }
  __syncthreads();
  if ((threadIdx.x & 31) == 0 && nf > 0) {
    atomicMin(s_b + 0, mnx); atomicMin(s_b + 1, mny); atomicMax(s_b + 2, mxx); atomicMax(s_b + 3, mxy); atomicAdd(s_b + 4, nf);
  }
  __syncthreads();
  mnx = s_b[0]; mny = s_b[1]; mxx = s_b[2]; mxy = s_b[3]; nf = s_b[4];
  if (threadIdx.x == 0 && nf > 0) {
    atomicMin(bounds + 0, mnx); atomicMin(bounds + 1, mny);
    atomicMax(bounds + 2, mxx); atomicMax(bounds + 3, mxy);
    atomicAdd(ctr + CTR_NFIN, nf);
  }
}

struct Dims { int32_t min_bx, min_by, div_x, div_y; float inv_leaf; };
constexpr int FINALIZE_BIG_LEAF = 96;     // leaves with more points are reduced by a warp instead of a thread

// Which grid does point i belong to? (batched scan pairs: point ranges per pair; one grid: always 0)
__device__ __forceinline__ int pair_of(const int64_t *__restrict__ off, int n_pairs, int64_t i) {
  if (n_pairs <= 1) return 0;
  int lo = 0, hi = n_pairs;               // off[lo] <= i < off[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(off + mid) <= i) lo = mid; else hi = mid;
  }
  return lo;
}

// pass 1a: cell of each point + a unique rank inside the cell (warp-aggregated int atomics).
// All grids share one padded table; grid g owns entries [dims[g].base, dims[g].base + W*H).
__global__ void __launch_bounds__(256) k_count(const float4 *__restrict__ pts, int64_t n,
                                              const int64_t *__restrict__ off, int n_pairs,
                                              const PairDims *__restrict__ dims, float inv_leaf,
                                              int32_t *__restrict__ count, int32_t *__restrict__ cell_of,
                                              int32_t *__restrict__ rank_of) {
  const int lane = threadIdx.x & 31;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t n_round = ((n + 31) / 32) * 32;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += stride) {
    int pcell = -1;
    if (i < n) {
      const float4 p = __ldg(pts + i);
      if (isfinite(p.x) && isfinite(p.y) && isfinite(p.z)) {
        const PairDims d = dims[pair_of(off, n_pairs, i)];
        if (d.div_x > 0) {
          const int i0 = cell_coord(p.x, inv_leaf, d.min_bx);
          const int i1 = cell_coord(p.y, inv_leaf, d.min_by);
          pcell = d.base + (i1 + 2) * d.W + i0 + 2;     // position in the shared padded count / slot table
        }
      }
    }
    // lanes of this warp that hit the same cell share one atomic
    const unsigned grp = __match_any_sync(0xffffffffu, pcell);
    const int leader = __ffs(grp) - 1;
    int base = 0;
    if (pcell >= 0 && lane == leader) base = atomicAdd(count + pcell, __popc(grp));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (i < n) {
      cell_of[i] = pcell;
      rank_of[i] = base + __popc(grp & ((1u << lane) - 1u));
    }
  }
}

// pass 1b: walk the padded count tables row by row (blockIdx.x = grid * chunks + chunk); every occupied cell gets a leaf
// id and a bucket range and its count turns into leaf id + 1 in place (0 stays "empty"). Only occupied cells are written:
// the probe table was NaN-filled and the count table zeroed by memsets (a C3 grid has 16.8 M cells for 0.6 M leaves), and
// the slot table needs no initialisation at all -- it is only ever read for cells whose centroid is not NaN.
__global__ void __launch_bounds__(256) k_alloc(int32_t *__restrict__ count_leaf, const PairDims *__restrict__ dims, int chunks,
                                              int32_t *__restrict__ leaf_cell, int32_t *__restrict__ leaf_pair,
                                              int32_t *__restrict__ leaf_n, int32_t *__restrict__ leaf_start,
                                              int32_t *__restrict__ big_list, int32_t *__restrict__ ctr) {
  __shared__ int s_leaves[8], s_pts[8], s_base[2];
  const int lane = threadIdx.x & 31;
  const int pair = blockIdx.x / chunks, chunk = blockIdx.x - pair * chunks;
  const PairDims d = dims[pair];
  const int W = d.W;
  const int warps_per_block = blockDim.x >> 5, warp_in_block = threadIdx.x >> 5;
  // sweep 1: how many leaves / points does this warp own? One pair of global atomics per CTA (577 k leaves at C3 would
  // otherwise serialise on two addresses).
  int my_pts = 0, wl = 0;
  for (int r = 2 + chunk; r < d.div_y + 2; r += chunks) {           // interior rows only: the padding holds no points
    const size_t row0 = (size_t)d.base + (size_t)r * W;
    for (int c0 = warp_in_block * 32; c0 < W; c0 += warps_per_block * 32) {
      const int c = c0 + lane;
      const int n = (c < W) ? count_leaf[row0 + c] : 0;
      my_pts += n;
      wl += __popc(__ballot_sync(0xffffffffu, n > 0));
    }
  }
  const int wp = __reduce_add_sync(0xffffffffu, my_pts);
  if (lane == 0) { s_leaves[warp_in_block] = wl; s_pts[warp_in_block] = wp; }
  __syncthreads();
  if (threadIdx.x == 0) {
    int tl = 0, tp = 0;
    for (int w = 0; w < warps_per_block; ++w) { tl += s_leaves[w]; tp += s_pts[w]; }
    s_base[0] = tl ? atomicAdd(ctr + CTR_LEAVES, tl) : 0;
    s_base[1] = tp ? atomicAdd(ctr + CTR_PTS, tp) : 0;
  }
  __syncthreads();
  int run_leaf = s_base[0], run_pts = s_base[1];
  for (int w = 0; w < warp_in_block; ++w) { run_leaf += s_leaves[w]; run_pts += s_pts[w]; }
  if (wl == 0) return;
  // sweep 2: same walk, ids and bucket ranges from the warp's running offsets
  for (int r = 2 + chunk; r < d.div_y + 2; r += chunks) {
    const size_t row0 = (size_t)d.base + (size_t)r * W;
    for (int c0 = warp_in_block * 32; c0 < W; c0 += warps_per_block * 32) {
      const int c = c0 + lane;
      const int n = (c < W) ? count_leaf[row0 + c] : 0;
      const bool has = n > 0;
      const unsigned bal = __ballot_sync(0xffffffffu, has);
      if (bal == 0u) continue;
      int incl = n;
#pragma unroll
      for (int dlt = 1; dlt < 32; dlt <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, dlt);
        if (lane >= dlt) incl += t;
      }
      const int total = __shfl_sync(0xffffffffu, incl, 31);
      if (has) {
        const int leaf = run_leaf + __popc(bal & ((1u << lane) - 1u));
        leaf_cell[leaf] = (int32_t)(row0 + c);
        leaf_pair[leaf] = pair;
        leaf_n[leaf] = n;
        leaf_start[leaf] = run_pts + incl - n;
        count_leaf[row0 + c] = leaf + 1;
        if (n > FINALIZE_BIG_LEAF) big_list[atomicAdd(ctr + CTR_BIG, 1)] = leaf;   // dense leaves get a warp in k_finalize
      }
      run_leaf += __popc(bal);
      run_pts += total;
    }
  }
}

// pass 1c: point index -> its leaf bucket
__global__ void __launch_bounds__(256) k_fill(int64_t n, const int32_t *__restrict__ cell_of,
                                             const int32_t *__restrict__ rank_of,
                                             const int32_t *__restrict__ leaf_id,
                                             const int32_t *__restrict__ leaf_start, int32_t *__restrict__ list) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int cell = cell_of[i];
    if (cell < 0) continue;
    const int leaf = __ldg(leaf_id + cell) - 1;
    list[__ldg(leaf_start + leaf) + rank_of[i]] = (int32_t)i;
  }
}

// pass 1d: order every bucket by point index. One thread per bucket entry: its rank is the number of
// smaller indices in the same bucket (a thread does O(n) work for a bucket of n; the work of a dense
// cell is spread over all of its points instead of one warp). Threads of a warp mostly share a bucket,
// so the inner loads are broadcasts.
__global__ void __launch_bounds__(256) k_rank(int64_t n, const int32_t *__restrict__ cell_of,
                                             const int32_t *__restrict__ leaf_id,
                                             const int32_t *__restrict__ leaf_start,
                                             const int32_t *__restrict__ leaf_n, const int32_t *__restrict__ list,
                                             const float4 *__restrict__ pts, int32_t *__restrict__ sorted_idx,
                                             float2 *__restrict__ tgt_sorted, int32_t *__restrict__ ctr) {
  const int64_t n_bucketed = ctr[CTR_PTS];
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n_bucketed; p += (int64_t)gridDim.x * blockDim.x) {
    const int v = list[p];
    const int leaf = __ldg(leaf_id + cell_of[v]) - 1;
    const int st = __ldg(leaf_start + leaf), m = __ldg(leaf_n + leaf);
    int r = 0;
    // head up to a 16-byte boundary, then four indices per load (the loads of a warp are broadcasts: its threads mostly
    // share a bucket), tail
    const int32_t *__restrict__ bk = list + st;
    int j = 0;
    const int head = min(m, (int)((4 - (st & 3)) & 3));
    for (; j < head; ++j) r += (__ldg(bk + j) < v) ? 1 : 0;
    const int4 *__restrict__ bk4 = reinterpret_cast<const int4 *>(bk + j);
    const int n4 = (m - j) >> 2;
#pragma unroll 4
    for (int q = 0; q < n4; ++q) {
      const int4 t = __ldg(bk4 + q);
      r += (t.x < v) + (t.y < v) + (t.z < v) + (t.w < v);
    }
    for (j += 4 * n4; j < m; ++j) r += (__ldg(bk + j) < v) ? 1 : 0;
    sorted_idx[st + r] = v;
    const float4 pt = __ldg(pts + v);
    tgt_sorted[st + r] = make_float2(pt.x, pt.y);     // points in bucket order: finalize and the 1-NN read them contiguously
  }
}

// ---- stable counting sort by cell for small grids with dense buckets (a C2 local map: 300 k points in ~5 k cells) ----
// Rank-by-counting (k_rank) is quadratic in the bucket size; here the points are walked in input order in tiles of
// TILE_PTS: (1) per-tile histogram over the cells (shared-memory atomics) -> tile_hist[tile][cell], also the global
// per-cell counts and cell_of; (2) exclusive scan over the tiles, per cell; (3) every tile places its points: position =
// bucket start + points of the cell in earlier tiles + points of the cell earlier in this tile (rounds of 256 points, warps
// take turns in order). The result is exactly the input-order bucket layout k_fill + k_rank produce.
constexpr int TILE_PTS = 2048, TILE_CELLS_CAP = 12288;

__global__ void __launch_bounds__(256) k_tile_hist(const float4 *__restrict__ pts, int64_t n, const PairDims *__restrict__ dims,
                                                  float inv_leaf, int npad, int32_t *__restrict__ count,
                                                  int32_t *__restrict__ cell_of, int32_t *__restrict__ tile_hist) {
  extern __shared__ int32_t s_hist[];
  for (int c = threadIdx.x; c < npad; c += blockDim.x) s_hist[c] = 0;
  __syncthreads();
  const PairDims d = dims[0];
  const int64_t t0 = (int64_t)blockIdx.x * TILE_PTS;
  for (int k = threadIdx.x; k < TILE_PTS; k += blockDim.x) {
    const int64_t i = t0 + k;
    if (i >= n) break;
    const float4 p = __ldg(pts + i);
    int pcell = -1;
    if (isfinite(p.x) && isfinite(p.y) && isfinite(p.z) && d.div_x > 0) {
      const int i0 = cell_coord(p.x, inv_leaf, d.min_bx), i1 = cell_coord(p.y, inv_leaf, d.min_by);
      pcell = d.base + (i1 + 2) * d.W + i0 + 2;
      atomicAdd(&s_hist[pcell], 1);
    }
    cell_of[i] = pcell;
  }
  __syncthreads();
  int32_t *out = tile_hist + (size_t)blockIdx.x * npad;
  for (int c = threadIdx.x; c < npad; c += blockDim.x) {
    const int v = s_hist[c];
    out[c] = v;
    if (v) atomicAdd(count + c, v);
  }
}

__global__ void __launch_bounds__(64) k_tile_scan(int32_t *__restrict__ tile_hist, int n_tiles, int npad) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= npad) return;
  int run = 0;
  for (int t0 = 0; t0 < n_tiles; t0 += 8) {          // eight independent loads in flight, then the running sum
    int v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = (t0 + u < n_tiles) ? tile_hist[(size_t)(t0 + u) * npad + c] : 0;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (t0 + u < n_tiles) tile_hist[(size_t)(t0 + u) * npad + c] = run;
      run += v[u];
    }
  }
}

__global__ void __launch_bounds__(256) k_tile_place(const float4 *__restrict__ pts, int64_t n, int npad,
                                                   const int32_t *__restrict__ cell_of, const int32_t *__restrict__ leaf_id,
                                                   const int32_t *__restrict__ leaf_start, const int32_t *__restrict__ tile_off,
                                                   int32_t *__restrict__ sorted_idx, float2 *__restrict__ tgt_sorted) {
  extern __shared__ int32_t s_cnt[];
  for (int c = threadIdx.x; c < npad; c += blockDim.x) s_cnt[c] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
  const unsigned lt = (1u << lane) - 1u;
  const int32_t *__restrict__ off = tile_off + (size_t)blockIdx.x * npad;
  const int64_t t0 = (int64_t)blockIdx.x * TILE_PTS;
  for (int r = 0; r < TILE_PTS; r += blockDim.x) {
    const int64_t i = t0 + r + threadIdx.x;
    const int cell = (i < n) ? cell_of[i] : -1;
    const unsigned grp = __match_any_sync(0xffffffffu, cell);
    int local = 0;
    for (int w = 0; w < n_warps; ++w) {              // warps take turns in index order
      if (warp == w && cell >= 0) {
        local = s_cnt[cell] + __popc(grp & lt);
        __syncwarp(grp);
        if (lane == __ffs(grp) - 1) s_cnt[cell] += __popc(grp);
      }
      __syncthreads();
    }
    if (cell >= 0) {
      const int leaf = __ldg(leaf_id + cell) - 1;
      const int pos = __ldg(leaf_start + leaf) + __ldg(off + cell) + local;
      sorted_idx[pos] = (int32_t)i;
      const float4 p = __ldg(pts + i);
      tgt_sorted[pos] = make_float2(p.x, p.y);
    }
  }
}

// symmetric 2x2 eigen-decomposition (a b; b d): ascending eigenvalues, orthonormal columns.
// Expression order is fixed (IEEE ops only) so results are reproducible bit for bit.
__device__ inline void eig2(double a, double b, double d, double *lam, double *v0, double *v1) {
  const double tr = a + d, df = a - d;
  const double rt = sqrt(df * df + 4.0 * b * b);
  double l1, l0;
  if (tr >= 0) { l1 = 0.5 * (tr + rt); l0 = (l1 != 0.0) ? (a * d - b * b) / l1 : 0.5 * (tr - rt); }
  else { l0 = 0.5 * (tr - rt); l1 = (l0 != 0.0) ? (a * d - b * b) / l0 : 0.5 * (tr + rt); }
  if (l0 > l1) { const double t = l0; l0 = l1; l1 = t; }
  lam[0] = l0; lam[1] = l1;
  double ex, ey;
  if (fabs(b) > 0) {
    if (fabs(l1 - a) > fabs(l1 - d)) { ex = b; ey = l1 - a; }
    else { ex = l1 - d; ey = b; }
    double nn = sqrt(ex * ex + ey * ey);
    if (nn == 0) { ex = 1; ey = 0; nn = 1; }
    ex /= nn; ey /= nn;
  } else {
    if (a >= d) { ex = 1; ey = 0; } else { ex = 0; ey = 1; }
  }
  v1[0] = ex; v1[1] = ey;
  v0[0] = -ey; v0[1] = ex;
}

struct FinalizeParams { int32_t min_points; double eig_mult; int32_t quirks; };

// Neighbour masks for the batch kernels, derived from the probe table: entry t gets bit (dj + 1) * 4 + (di + 1) for every
// tree cell (non-NaN centroid) at offset (di, dj). The tables carry two empty cells of padding on every side of every
// grid, so the eight neighbours of any *interior or first-ring* entry exist; entries on the outermost ring are never
// addressed by a point (a candidate's own cell is at most one cell outside the grid) and get 0.
__global__ void __launch_bounds__(256) k_nbr_from_cen(const float2 *__restrict__ cen, const PairDims *__restrict__ dims, int n_grids,
                                                      int64_t n_entries, uint16_t *__restrict__ nbr) {
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n_entries; t += (int64_t)gridDim.x * blockDim.x) {
    int lo = 0, hi = n_grids;                 // which grid owns entry t: dims[lo].base <= t < dims[hi].base
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if ((int64_t)__ldg(&dims[mid].base) <= t) lo = mid; else hi = mid;
    }
    const int W = __ldg(&dims[lo].W), H = __ldg(&dims[lo].H);
    const int r = (int)(t - __ldg(&dims[lo].base));
    const int row = r / W, col = r - row * W;
    unsigned m = 0u;
    if (row >= 1 && row < H - 1 && col >= 1 && col < W - 1) {
#pragma unroll
      for (int dj = -1; dj <= 1; ++dj)
#pragma unroll
        for (int di = -1; di <= 1; ++di) {
          const float2 c = __ldg(cen + t + dj * W + di);
          if (c.x == c.x) m |= 1u << ((dj + 1) * 4 + (di + 1));
        }
    }
    nbr[t] = (uint16_t)m;
  }
}

// pass 2. The bucket of a leaf is contiguous and already in input order (k_rank); its sums must be taken in that
// order (fp32 centroid and fp64 sums accumulate exactly like the reference's pass 1: bit-identical results), so a
// leaf is one sequential chain of adds. Sparse leaves (C3: ~7 points) take one thread each. A dense leaf (C2 local
// map: thousands of points in a 0.5 m cell) takes a warp: coalesced 32-point loads, then every lane replays the
// same chain from register shuffles -- the chain then runs at add latency instead of one memory round trip per point.
struct LeafSums { double sx, sy, sxx, syx, syy; float cx, cy; };

struct FinalizeOut {
  int2 *__restrict__ leaf_range; int32_t *__restrict__ leaf_nr; double2 *__restrict__ leaf_mean; double *__restrict__ leaf_icov;
  float2 *__restrict__ leaf_cen; int32_t *__restrict__ slot; float2 *__restrict__ cen_tab; uint32_t *__restrict__ occ;
  CellRec *__restrict__ recs; int32_t *__restrict__ ctr;
  const int32_t *__restrict__ leaf_cell; const int32_t *__restrict__ leaf_pair; const PairDims *__restrict__ dims;
};

// Mean, single-pass covariance, eigenvalue clamp and inverse of one cell from its in-order sums (VoxelGridCovariance pass 2,
// SURVEY App. A.2). Shared by the full build (finish_leaf) and the incremental update (k_inc_update): same IEEE operations
// in the same order, so a cell comes out bit-identical whichever path touched it last.
struct LeafStats { int nr; bool in_tree; float cx, cy; double m0, m1, ic0, ic1, ic2, ic3; };
__device__ __forceinline__ LeafStats leaf_stats(const int n, const LeafSums sums, const FinalizeParams fp) {
  double sx = sums.sx, sy = sums.sy;
    const double sxx = sums.sxx, syx = sums.syx, syy = sums.syy;
    float cx = sums.cx, cy = sums.cy;
    const double nn = (double)n;
    cx = cx / (float)n; cy = cy / (float)n;
    const double psx = sx, psy = sy;
    const double m0 = sx / nn, m1 = sy / nn;
    int nr = n;
    double ic0 = 0, ic1 = 0, ic2 = 0, ic3 = 0;
    const bool in_tree = n >= fp.min_points;
    if (in_tree) {
      const double id = (fp.quirks & NDT_QUIRK_COV_INIT_IDENTITY) ? 1.0 : 0.0;
      const double Cxx = id + sxx, Cyy = id + syy, Cyx = syx, Cxy = syx, Czz = id;
      double cxx, cxy, cyx, cyy, czz;
      if (fp.quirks & NDT_QUIRK_COV_SCALE_NM1_N) {
        cxx = (Cxx - 2.0 * (psx * m0)) / nn + m0 * m0;
        cxy = (Cxy - 2.0 * (psx * m1)) / nn + m0 * m1;
        cyx = (Cyx - 2.0 * (psy * m0)) / nn + m1 * m0;
        cyy = (Cyy - 2.0 * (psy * m1)) / nn + m1 * m1;
        czz = Czz / nn;
        const double sc = (nn - 1.0) / nn;
        cxx *= sc; cxy *= sc; cyx *= sc; cyy *= sc; czz *= sc;
      } else {
        const double dn = nn - 1.0;
        cxx = (Cxx - psx * m0) / dn; cxy = (Cxy - psx * m1) / dn;
        cyx = (Cyx - psy * m0) / dn; cyy = (Cyy - psy * m1) / dn;
        czz = Czz / dn;
      }
      double lam2[2], v0[2], v1[2];
      eig2(cxx, cyx, cyy, lam2, v0, v1);
      double ev[3] = {czz, lam2[0], lam2[1]};
      int which[3] = {2, 0, 1};
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = a + 1; b < 3; ++b)
          if (ev[b] < ev[a]) { double t = ev[a]; ev[a] = ev[b]; ev[b] = t; int w = which[a]; which[a] = which[b]; which[b] = w; }
      if (ev[0] < 0 || ev[1] < 0 || ev[2] <= 0) {
        nr = -1;
      } else {
        const double mcv = fp.eig_mult * ev[2];
        if (ev[0] < mcv) {
          ev[0] = mcv;
          if (ev[1] < mcv) ev[1] = mcv;
          double l0 = lam2[0], l1 = lam2[1];
#pragma unroll
          for (int a = 0; a < 3; ++a) {
            if (which[a] == 0) l0 = ev[a];
            else if (which[a] == 1) l1 = ev[a];
            else czz = ev[a];
          }
          const double det = v0[0] * v1[1] - v1[0] * v0[1];
          const double i00 = v1[1] / det, i01 = -v1[0] / det, i10 = -v0[1] / det, i11 = v0[0] / det;
          const double a00 = v0[0] * l0, a01 = v1[0] * l1, a10 = v0[1] * l0, a11 = v1[1] * l1;
          cxx = a00 * i00 + a01 * i10; cxy = a00 * i01 + a01 * i11;
          cyx = a10 * i00 + a11 * i10; cyy = a10 * i01 + a11 * i11;
        }
        const double det = cxx * cyy - cxy * cyx;
        ic0 = cyy / det; ic1 = -cxy / det; ic2 = -cyx / det; ic3 = cxx / det;
        const double izz = 1.0 / czz;
        const double mxc = fmax(fmax(fmax(ic0, ic1), fmax(ic2, ic3)), fmax(izz, 0.0));
        const double mnc = fmin(fmin(fmin(ic0, ic1), fmin(ic2, ic3)), fmin(izz, 0.0));
        if (mxc == (double)INFINITY || mnc == -(double)INFINITY) nr = -1;
      }
    }
    LeafStats z;
    z.nr = nr; z.in_tree = in_tree; z.cx = cx; z.cy = cy; z.m0 = m0; z.m1 = m1; z.ic0 = ic0; z.ic1 = ic1; z.ic2 = ic2; z.ic3 = ic3;
    return z;
}

__device__ __forceinline__ void finish_leaf(const int leaf, const int n, const int st, const LeafSums sums, const FinalizeParams fp,
                                            const FinalizeOut &o) {
    int2 *leaf_range = o.leaf_range; int32_t *leaf_nr = o.leaf_nr; double2 *leaf_mean = o.leaf_mean; double *leaf_icov = o.leaf_icov;
    float2 *leaf_cen = o.leaf_cen; int32_t *slot = o.slot; float2 *cen_tab = o.cen_tab; uint32_t *occ = o.occ; CellRec *recs = o.recs;
    int32_t *ctr = o.ctr; const int32_t *leaf_cell = o.leaf_cell; const int32_t *leaf_pair = o.leaf_pair; const PairDims *dims = o.dims;
    leaf_range[leaf] = make_int2(st, n);
    const LeafStats z = leaf_stats(n, sums, fp);
    const float cx = z.cx, cy = z.cy;
    const double m0 = z.m0, m1 = z.m1, ic0 = z.ic0, ic1 = z.ic1, ic2 = z.ic2, ic3 = z.ic3;
    const int nr = z.nr;
    const bool in_tree = z.in_tree;
    leaf_nr[leaf] = nr;
    leaf_mean[leaf] = make_double2(m0, m1);
    leaf_icov[4 * (size_t)leaf + 0] = ic0; leaf_icov[4 * (size_t)leaf + 1] = ic1;
    leaf_icov[4 * (size_t)leaf + 2] = ic2; leaf_icov[4 * (size_t)leaf + 3] = ic3;
    leaf_cen[leaf] = make_float2(cx, cy);
    if (in_tree) {
      // warp-aggregated slot allocation (one atomic per group of lanes that got here together)
      cg::coalesced_group act = cg::coalesced_threads();
      int s = 0;
      if (act.thread_rank() == 0) s = atomicAdd(ctr + CTR_SLOTS, (int)act.size());
      s = act.shfl(s, 0) + (int)act.thread_rank();
      const unsigned okm = act.ballot(nr > 0);
      if (act.thread_rank() == 0 && okm) atomicAdd(ctr + CTR_VALID, __popc(okm));
      CellRec r;
      r.cx = cx; r.cy = cy; r.nr_points = nr; r.cell = leaf_cell[leaf];
      r.mx = m0; r.my = m1; r.c00 = ic0; r.c01 = ic1; r.c10 = ic2; r.c11 = ic3;
      recs[s] = r;
      const int q = r.cell;                               // position in the shared padded tables
      const int W = dims[leaf_pair[leaf]].W;
      slot[q] = s;
      cen_tab[q] = make_float2(cx, cy);
      // dilated occupancy: every cell whose 3x3 block contains this tree cell
#pragma unroll
      for (int dj = -1; dj <= 1; ++dj)
#pragma unroll
        for (int di = -1; di <= 1; ++di) {
          const int t = q + dj * W + di;
          atomicOr(occ + (t >> 5), 1u << (t & 31));
        }
    }
}

__global__ void __launch_bounds__(128) k_finalize(const float2 *__restrict__ tgt_sorted, const int32_t *__restrict__ leaf_n,
                                                 const int32_t *__restrict__ leaf_start, const int32_t *__restrict__ big_list,
                                                 FinalizeParams fp, FinalizeOut o) {
  const int n_leaves = o.ctr[CTR_LEAVES];
  // sparse leaves: one thread each
  for (int leaf = blockIdx.x * blockDim.x + threadIdx.x; leaf < n_leaves; leaf += gridDim.x * blockDim.x) {
    const int n = leaf_n[leaf], st = leaf_start[leaf];
    if (n > FINALIZE_BIG_LEAF) continue;
    LeafSums a{0, 0, 0, 0, 0, 0.f, 0.f};
    const float2 *__restrict__ bucket = tgt_sorted + st;
    for (int k = 0; k < n; ++k) {
      const float2 p = __ldg(bucket + k);
      const double xd = (double)p.x, yd = (double)p.y;
      a.sx += xd; a.sy += yd;
      a.sxx += xd * xd; a.syx += yd * xd; a.syy += yd * yd;
      a.cx = __fadd_rn(a.cx, p.x); a.cy = __fadd_rn(a.cy, p.y);
    }
    finish_leaf(leaf, n, st, a, fp, o);
  }
  // dense leaves (listed by k_alloc): one warp each. Per 32-point chunk every lane converts its own point and forms its
  // three products (the same IEEE operations the sequential code performs per point), parks the seven terms in shared
  // memory, and the add chain then runs over broadcast shared-memory reads: nothing but the adds is left in the serial
  // part, and the next chunk's global loads are already in flight.
  __shared__ double s_d[4][32][5];
  __shared__ float s_f[4][32][2];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
  const int n_big = o.ctr[CTR_BIG];
  for (int b = warp; b < n_big; b += n_warps) {
    const int leaf = big_list[b];
    const int n = leaf_n[leaf], st = leaf_start[leaf];
    LeafSums a{0, 0, 0, 0, 0, 0.f, 0.f};
    const float2 *__restrict__ bucket = tgt_sorted + st;
    float2 nxt = (lane < n) ? __ldg(bucket + lane) : make_float2(0.f, 0.f);
    for (int c = 0; c < n; c += 32) {
      const float2 mp = nxt;
      if (c + 32 + lane < n) nxt = __ldg(bucket + c + 32 + lane);
      const double xd = (double)mp.x, yd = (double)mp.y;
      s_d[wib][lane][0] = xd; s_d[wib][lane][1] = yd;
      s_d[wib][lane][2] = xd * xd; s_d[wib][lane][3] = yd * xd; s_d[wib][lane][4] = yd * yd;
      s_f[wib][lane][0] = mp.x; s_f[wib][lane][1] = mp.y;
      __syncwarp();
      const int m = min(32, n - c);
#pragma unroll 8
      for (int k = 0; k < m; ++k) {
        a.sx += s_d[wib][k][0]; a.sy += s_d[wib][k][1];
        a.sxx += s_d[wib][k][2]; a.syx += s_d[wib][k][3]; a.syy += s_d[wib][k][4];
        a.cx = __fadd_rn(a.cx, s_f[wib][k][0]); a.cy = __fadd_rn(a.cy, s_f[wib][k][1]);
      }
      __syncwarp();
    }
    if (lane == 0) finish_leaf(leaf, n, st, a, fp, o);
  }
}

// ---- fine nearest-neighbour lattice (unordered buckets: an exact minimum does not care about order) ----
__global__ void __launch_bounds__(256) k_nn_count(const float4 *__restrict__ pts, int64_t n, Dims d,
                                                 int32_t *__restrict__ cnt, int32_t *__restrict__ cell_of,
                                                 int32_t *__restrict__ rank_of) {
  const int lane = threadIdx.x & 31;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t n_round = ((n + 31) / 32) * 32;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += stride) {
    int cell = -1;
    if (i < n) {
      const float4 p = __ldg(pts + i);
      if (isfinite(p.x) && isfinite(p.y) && isfinite(p.z)) {
        int i0 = cell_coord(p.x, d.inv_leaf, d.min_bx), i1 = cell_coord(p.y, d.inv_leaf, d.min_by);
        i0 = min(max(i0, 0), d.div_x - 1); i1 = min(max(i1, 0), d.div_y - 1);
        cell = i0 + i1 * d.div_x;
      }
    }
    const unsigned grp = __match_any_sync(0xffffffffu, cell);
    const int leader = __ffs(grp) - 1;
    int base = 0;
    if (cell >= 0 && lane == leader) base = atomicAdd(cnt + cell, __popc(grp));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (i < n) { cell_of[i] = cell; rank_of[i] = base + __popc(grp & ((1u << lane) - 1u)); }
  }
}

__global__ void __launch_bounds__(256) k_nn_alloc(const int32_t *__restrict__ cnt, int64_t n_cells,
                                                 int2 *__restrict__ range, int32_t *__restrict__ ctr) {
  const int lane = threadIdx.x & 31;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t n_round = ((n_cells + 31) / 32) * 32;
  for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < n_round; c += stride) {
    const int n = (c < n_cells) ? cnt[c] : 0;
    int incl = n;
#pragma unroll
    for (int dlt = 1; dlt < 32; dlt <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, dlt);
      if (lane >= dlt) incl += t;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    int base = 0;
    if (lane == 0 && total > 0) base = atomicAdd(ctr + CTR_JOB, total);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (c < n_cells) range[c] = make_int2(base + incl - n, n);
  }
}

__global__ void __launch_bounds__(256) k_nn_fill(const float4 *__restrict__ pts, int64_t n,
                                                const int32_t *__restrict__ cell_of,
                                                const int32_t *__restrict__ rank_of, const int2 *__restrict__ range,
                                                float2 *__restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int cell = cell_of[i];
    if (cell < 0) continue;
    const float4 p = __ldg(pts + i);
    out[__ldg(range + cell).x + rank_of[i]] = make_float2(p.x, p.y);
  }
}

__global__ void __launch_bounds__(256) k_cell_index(const float4 *__restrict__ pts, int64_t n, Dims d,
                                                   int32_t *__restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 p = __ldg(pts + i);
    out[i] = cell_coord(p.x, d.inv_leaf, d.min_bx) + cell_coord(p.y, d.inv_leaf, d.min_by) * d.div_x;
  }
}

// ---- batched scan pairs: geometry of every pair's grid, computed on the device ----
// One warp per pair: getMinMax3D over the pair's finite target points, then the same float32 bounding-box
// arithmetic as the single-grid host code in grid_build() (VoxelGridCovariance::applyFilter).
// An empty / degenerate target gets a 4 x 4 all-empty table so the matcher needs no special case.
__global__ void __launch_bounds__(256) k_pair_bounds(const float4 *__restrict__ tgt, const int64_t *__restrict__ tgt_off,
                                                    const int64_t *__restrict__ src_off, int n_pairs, float inv_leaf,
                                                    PairDims *__restrict__ dims, int64_t *__restrict__ pad) {
  const int lane = threadIdx.x & 31;
  const int pair = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (pair >= n_pairs) return;
  const int64_t t0 = tgt_off[pair], t1 = tgt_off[pair + 1];
  int mnx = INT_MAX, mny = INT_MAX, mxx = INT_MIN, mxy = INT_MIN, nf = 0;
  for (int64_t i = t0 + lane; i < t1; i += 32) {
    const float4 p = __ldg(tgt + i);
    if (isfinite(p.x) && isfinite(p.y) && isfinite(p.z)) {
      const int ox = float_ord(p.x), oy = float_ord(p.y);
      mnx = min(mnx, ox); mny = min(mny, oy); mxx = max(mxx, ox); mxy = max(mxy, oy);
      ++nf;
    }
  }
  mnx = __reduce_min_sync(0xffffffffu, mnx); mny = __reduce_min_sync(0xffffffffu, mny);
  mxx = __reduce_max_sync(0xffffffffu, mxx); mxy = __reduce_max_sync(0xffffffffu, mxy);
  nf = __reduce_add_sync(0xffffffffu, nf);
  if (lane != 0) return;
  PairDims d{};
  d.src_off = src_off[pair]; d.tgt_off = t0; d.nt = t1 - t0;
  d.ns = (int32_t)(src_off[pair + 1] - src_off[pair]);
  bool empty = (nf == 0);
  if (!empty) {
    const float fx0 = ord_to_float(mnx), fy0 = ord_to_float(mny), fx1 = ord_to_float(mxx), fy1 = ord_to_float(mxy);
    const int64_t dx = (int64_t)((fx1 - fx0) * inv_leaf) + 1;
    const int64_t dy = (int64_t)((fy1 - fy0) * inv_leaf) + 1;
    if (dx * dy > (int64_t)INT_MAX) empty = true;          // PCL: "leaf size too small", empty grid
    if (!empty) {
      d.min_bx = (int)floorf(fx0 * inv_leaf); d.min_by = (int)floorf(fy0 * inv_leaf);
      d.div_x = (int)floorf(fx1 * inv_leaf) - d.min_bx + 1;
      d.div_y = (int)floorf(fy1 * inv_leaf) - d.min_by + 1;
    }
  }
  if (empty) { d.min_bx = d.min_by = 0; d.div_x = d.div_y = 0; }
  d.W = d.div_x + 4; d.H = d.div_y + 4;
  dims[pair] = d;
  pad[pair] = (int64_t)d.W * d.H;
}

// exclusive scan of the padded table sizes -> PairDims::base; out[0] = total entries, out[1] = tallest table
__global__ void __launch_bounds__(1024) k_pair_scan(PairDims *__restrict__ dims, const int64_t *__restrict__ pad, int n_pairs,
                                                   int64_t *__restrict__ out) {
  __shared__ int64_t s_sum[32];
  __shared__ int s_max[32];
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  const int chunk = (n_pairs + 1023) / 1024;
  const int lo = min(n_pairs, t * chunk), hi = min(n_pairs, lo + chunk);
  int64_t mine = 0;
  int mh = 0;
  for (int i = lo; i < hi; ++i) { mine += pad[i]; mh = max(mh, dims[i].H); }
  int64_t incl = mine;
#pragma unroll
  for (int dlt = 1; dlt < 32; dlt <<= 1) {
    const int64_t v = __shfl_up_sync(0xffffffffu, incl, dlt);
    if (lane >= dlt) incl += v;
  }
  mh = __reduce_max_sync(0xffffffffu, mh);
  if (lane == 31) s_sum[w] = incl;
  if (lane == 0) s_max[w] = mh;
  __syncthreads();
  int64_t before = 0;
  for (int k = 0; k < w; ++k) before += s_sum[k];
  int64_t run = before + incl - mine;
  for (int i = lo; i < hi; ++i) {
    dims[i].base = (int32_t)min(run, (int64_t)INT_MAX);   // the host rejects totals beyond int32 before any table is touched
    run += pad[i];
  }
  if (t == 1023) {
    out[0] = before + incl;
    int m = 0;
    for (int k = 0; k < 32; ++k) m = max(m, s_max[k]);
    out[1] = m;
  }
}


// ------------------------------------------------------------------------------------------------------------------
// Incremental target (ndt_set_target_incremental): a map whose settled part only grows. Per padded cell the running sums
// of the settled prefix stay on the device (CellAcc); a call folds the newly settled points into them, adds the provisional
// tail on top and re-derives the cells touched by either (and by the previous call's tail). Sums are taken in input order --
// all settled points precede the tail in the cloud, and the settled prefix only grows at its end -- so every cell comes out
// bit-identical to a full build over the whole cloud (tests/test_gpu_incremental.py compares every table).
// ------------------------------------------------------------------------------------------------------------------
struct __align__(16) CellAcc { double sx, sy, sxx, syx, syy; float cx, cy; int32_t n; int32_t pad; };
static_assert(sizeof(CellAcc) == 64, "CellAcc must be 64 bytes");
enum { ST_OCC = 1, ST_TREE = 2, ST_VALID = 4, ST_SLOT = 8, ST_CHANGED = 128 };
enum { INC_A = 0, INC_B = 1, INC_P = 2, INC_U = 3 };     // inc_cnt entries: fold cells, tail cells, previous tail cells, union

// The batch of a call = the points that became settled since the last call followed by the provisional tail: cloud[lo .. hi),
// at most INC_BATCH_CAP points, touching a few hundred cells. The whole update is ONE kernel on ONE CTA (the phases are
// far too small to fill a GPU, and as eight separate launches they cost the host more time to issue than the device
// needs to run them); phases are separated by __syncthreads(), buffers written and read inside the kernel are accessed
// through plain pointers (no read-only path):
//   rank      cell of every batch point; its rank among the batch points of the same cell (all-pairs compare over a
//             shared-memory copy of the batch's cells: B^2 / threads steps, B ~ 1,000); the first point of a cell
//             registers the cell (list, local id, point count)
//   starts    exclusive scan of the per-cell counts
//   scatter   order[start(cell) + rank] = batch index: the batch ordered by (cell, input index)
//   union     U = batch cells + the previous call's tail cells, without duplicates
//   finalize  every cell of U: add its batch points in input order to the settled sums -- the settled ones (cloud index
//             < n_stable) for good (stored back), the tail ones on top for this call only -- then statistics -> probe
//             tables, record, counters
//   occ       dilated occupancy around every cell whose tree status changed: each of its nine neighbours gets the OR over
//             ITS 3x3 block
constexpr int INC_BATCH_CAP = 4096;
constexpr int INC_THREADS = 512;

struct IncTables {
  uint8_t *status; int32_t *slot; float2 *cen;
  CellRec *recs; int32_t *ctr;
};
struct IncBatch {
  const float4 *__restrict__ pts;      // the whole target cloud (device copy), read-only here
  int64_t lo, n_stable;                // batch = pts[lo .. lo + B); indices < n_stable are settled for good
  int B;
  const PairDims *__restrict__ dims;
  float inv_leaf;
  int32_t epoch;
  int W;                               // padded row length
  int32_t *cell_of, *rank_of, *order, *cell_cnt, *cell_start;    // [INC_BATCH_CAP] each
  int2 *lid_tab;                       // per padded cell: (epoch, local id) of a cell that has batch points
  int32_t *lb, *lp, *lu;               // this call's batch cells, the previous call's, their union
  int32_t *cnt, *mark;
  CellAcc *acc;
  uint32_t *occ;
};

__global__ void __launch_bounds__(INC_THREADS) k_inc_update(IncBatch b, FinalizeParams fp, IncTables T) {
  __shared__ __align__(16) int s_cell[INC_BATCH_CAP];
  __shared__ int s_warp[INC_THREADS / 32];
  const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, w = tid >> 5, nw = nthr >> 5;
  const int B = b.B;
  if (tid == 0) { b.cnt[INC_B] = 0; b.cnt[INC_U] = 0; }
  // ---- rank ----
  const PairDims d = b.dims[0];
  const int B4 = (B + 3) & ~3;
  for (int k = tid; k < B4; k += nthr) {
    int c = -2;                                    // padding up to a multiple of four: matches nothing
    if (k < B) {
      const float4 p = __ldg(b.pts + b.lo + k);
      c = -1;
      if (isfinite(p.x) && isfinite(p.y) && isfinite(p.z))
        c = d.base + (cell_coord(p.y, b.inv_leaf, d.min_by) + 2) * d.W + cell_coord(p.x, b.inv_leaf, d.min_bx) + 2;
    }
    s_cell[k] = c;
  }
  __syncthreads();
  for (int k = tid; k < B; k += nthr) {
    const int c = s_cell[k];
    int before = 0, total = 0;
    if (c >= 0) {
      // four cells per 16-byte shared load; the threads of a warp read the same words (broadcast). Matches in the group
      // that contains k itself are split by position, groups before k count fully towards the rank.
      const int4 *s4 = reinterpret_cast<const int4 *>(s_cell);
      const int kg = k >> 2;
#pragma unroll 4
      for (int g = 0; g < B4 / 4; ++g) {
        const int4 v = s4[g];
        const int m = (v.x == c) + (v.y == c) + (v.z == c) + (v.w == c);
        total += m;
        before += g < kg ? m : 0;
      }
      const int4 v = s4[kg];
      const int r = k & 3;
      before += (r > 0 && v.x == c) + (r > 1 && v.y == c) + (r > 2 && v.z == c);
    }
    b.cell_of[k] = c;
    b.rank_of[k] = before;
    if (c >= 0 && before == 0) {
      const int lid = atomicAdd(b.cnt + INC_B, 1);
      b.lb[lid] = c;
      b.lid_tab[c] = make_int2(b.epoch, lid);
      b.cell_cnt[lid] = total;
    }
  }
  __syncthreads();
  // ---- starts: exclusive scan of cell_cnt[0 .. nb) ----
  const int nb = b.cnt[INC_B];
  {
    int carry = 0;
    for (int base = 0; base < nb; base += nthr) {
      const int i = base + tid;
      const int v = i < nb ? b.cell_cnt[i] : 0;
      int incl = v;
#pragma unroll
      for (int dlt = 1; dlt < 32; dlt <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, dlt); if (lane >= dlt) incl += t; }
      if (lane == 31) s_warp[w] = incl;
      __syncthreads();
      int before = carry;
      for (int q = 0; q < w; ++q) before += s_warp[q];
      if (i < nb) b.cell_start[i] = before + incl - v;
      int tot = 0;
      for (int q = 0; q < nw; ++q) tot += s_warp[q];
      carry += tot;
      __syncthreads();
    }
  }
  __syncthreads();
  // ---- scatter ----
  for (int k = tid; k < B; k += nthr) {
    const int c = b.cell_of[k];
    if (c >= 0) b.order[b.cell_start[b.lid_tab[c].y] + b.rank_of[k]] = k;
  }
  // ---- union (independent of the scatter) ----
  const int np = b.cnt[INC_P];
  for (int t = tid; t < nb + np; t += nthr) {
    const int c = t < nb ? b.lb[t] : b.lp[t - nb];
    if (atomicExch(b.mark + c, b.epoch) != b.epoch) b.lu[atomicAdd(b.cnt + INC_U, 1)] = c;
  }
  __syncthreads();
  // ---- finalize ----
  const int m = b.cnt[INC_U];
  for (int t = tid; t < m; t += nthr) {
    const int c = b.lu[t];
    CellAcc v = b.acc[c];
    LeafSums a{v.sx, v.sy, v.sxx, v.syx, v.syy, v.cx, v.cy};
    int n = v.n;
    const int2 e = b.lid_tab[c];
    if (e.x == b.epoch) {                            // the cell has points in this call's batch
      const int st = b.cell_start[e.y], num = b.cell_cnt[e.y];
      bool settled_part = true, grew = false;
      for (int q = 0; q < num; ++q) {
        const int64_t i = b.lo + b.order[st + q];
        if (settled_part && i >= b.n_stable) {       // everything from here on is tail: park the settled sums first
          if (grew) { v.sx = a.sx; v.sy = a.sy; v.sxx = a.sxx; v.syx = a.syx; v.syy = a.syy; v.cx = a.cx; v.cy = a.cy; v.n = n; b.acc[c] = v; }
          settled_part = false;
        }
        const float4 p = __ldg(b.pts + i);
        const double xd = (double)p.x, yd = (double)p.y;
        a.sx += xd; a.sy += yd;
        a.sxx += xd * xd; a.syx += yd * xd; a.syy += yd * yd;
        a.cx = __fadd_rn(a.cx, p.x); a.cy = __fadd_rn(a.cy, p.y);
        ++n;
        grew = true;
      }
      if (settled_part && grew) { v.sx = a.sx; v.sy = a.sy; v.sxx = a.sxx; v.syx = a.syx; v.syy = a.syy; v.cx = a.cx; v.cy = a.cy; v.n = n; b.acc[c] = v; }
    }
    const int old = T.status[c];
    int now = old & ST_SLOT;
    if (n > 0) {
      const LeafStats z = leaf_stats(n, a, fp);
      now |= ST_OCC;
      if (z.in_tree) {
        int sl;
        if (old & ST_SLOT) sl = T.slot[c]; else { sl = atomicAdd(T.ctr + CTR_SLOTS, 1); now |= ST_SLOT; }
        CellRec r;
        r.cx = z.cx; r.cy = z.cy; r.nr_points = z.nr; r.cell = c;
        r.mx = z.m0; r.my = z.m1; r.c00 = z.ic0; r.c01 = z.ic1; r.c10 = z.ic2; r.c11 = z.ic3;
        T.recs[sl] = r;
        T.slot[c] = sl;
        T.cen[c] = make_float2(z.cx, z.cy);
        now |= ST_TREE | (z.nr > 0 ? ST_VALID : 0);
      }
    }
    if (!(now & ST_TREE)) T.cen[c] = make_float2(__int_as_float(-1), __int_as_float(-1));      // all-ones NaN, like the memset of a full build
    if ((now ^ old) & ST_TREE) now |= ST_CHANGED;
    const int d_occ = ((now & ST_OCC) ? 1 : 0) - ((old & ST_OCC) ? 1 : 0), d_val = ((now & ST_VALID) ? 1 : 0) - ((old & ST_VALID) ? 1 : 0);
    if (d_occ) atomicAdd(T.ctr + CTR_LEAVES, d_occ);
    if (d_val) atomicAdd(T.ctr + CTR_VALID, d_val);
    T.status[c] = (uint8_t)now;
  }
  __syncthreads();
  // ---- occ ----
  for (int t = tid; t < m; t += nthr) {
    const int c = b.lu[t];
    if (!(T.status[c] & ST_CHANGED)) continue;
    for (int dj = -1; dj <= 1; ++dj)
      for (int di = -1; di <= 1; ++di) {
        const int q = c + dj * b.W + di;
        bool any = false;
        for (int ej = -1; ej <= 1; ++ej)
          for (int ei = -1; ei <= 1; ++ei) any = any || (T.status[q + ej * b.W + ei] & ST_TREE);
        if (any) atomicOr(b.occ + (q >> 5), 1u << (q & 31)); else atomicAnd(b.occ + (q >> 5), ~(1u << (q & 31)));
      }
  }
  __syncthreads();
  for (int t = tid; t < m; t += nthr) T.status[b.lu[t]] &= (uint8_t)~ST_CHANGED;
  if (tid == 0) b.cnt[INC_P] = nb;                   // this call's batch cells are the next call's "previous" cells
}

// After a full build: the settled sums of every leaf (its bucket is in input order; sorted_idx carries the original indices, the
// first one >= n_stable ends the settled part), the state byte of its cell, and the list of cells the tail touches.
// One warp per leaf: 32 bucket entries per load, every lane replays the adds from shuffles.
__global__ void __launch_bounds__(128) k_inc_init(const float2 *__restrict__ tgt_sorted, const int32_t *__restrict__ sorted_idx,
                                                 const int32_t *__restrict__ leaf_n, const int32_t *__restrict__ leaf_start,
                                                 const int32_t *__restrict__ leaf_cell, const int32_t *__restrict__ leaf_nr, int32_t n_stable,
                                                 int32_t min_points, const int32_t *__restrict__ ctr, CellAcc *__restrict__ acc,
                                                 uint8_t *__restrict__ status, int32_t *__restrict__ tail_list, int32_t *__restrict__ tail_n) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
  const int n_leaves = ctr[CTR_LEAVES];
  for (int leaf = warp; leaf < n_leaves; leaf += n_warps) {
    const int n = leaf_n[leaf], st = leaf_start[leaf], c = leaf_cell[leaf];
    LeafSums a{0, 0, 0, 0, 0, 0.f, 0.f};
    int settled = 0;
    bool more = true;
    for (int k0 = 0; k0 < n && more; k0 += 32) {
      const int k = k0 + lane;
      float2 p = make_float2(0.f, 0.f);
      int idx = INT_MAX;
      if (k < n) { p = __ldg(tgt_sorted + st + k); idx = __ldg(sorted_idx + st + k); }
      const int m = min(32, n - k0);
      for (int j = 0; j < m; ++j) {
        const int ij = __shfl_sync(0xffffffffu, idx, j);
        if (ij >= n_stable) { more = false; break; }          // ascending indices: everything after is tail
        const float px = __shfl_sync(0xffffffffu, p.x, j), py = __shfl_sync(0xffffffffu, p.y, j);
        const double xd = (double)px, yd = (double)py;
        a.sx += xd; a.sy += yd;
        a.sxx += xd * xd; a.syx += yd * xd; a.syy += yd * yd;
        a.cx = __fadd_rn(a.cx, px); a.cy = __fadd_rn(a.cy, py);
        ++settled;
      }
    }
    if (lane == 0) {
      CellAcc v;
      v.sx = a.sx; v.sy = a.sy; v.sxx = a.sxx; v.syx = a.syx; v.syy = a.syy; v.cx = a.cx; v.cy = a.cy; v.n = settled; v.pad = 0;
      acc[c] = v;
      const int nr = leaf_nr[leaf];
      const bool tree = n >= min_points;
      status[c] = (uint8_t)(ST_OCC | (tree ? (ST_TREE | ST_SLOT) : 0) | ((tree && nr > 0) ? ST_VALID : 0));
      if (settled < n) tail_list[atomicAdd(tail_n, 1)] = c;
    }
  }
}

inline int grid_for(int64_t work, int threads, int sm_count, int per_sm = 8) {
  int64_t b = (work + threads - 1) / threads;
  const int64_t cap = (int64_t)sm_count * per_sm;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace

// Build every table from the points in gb.tgt for n_grids grids laid out back to back.
int grid_build_tables(Handle *h, int64_t n, int n_grids, int64_t total_pad, int max_h) {
  GridBuffers &gb = h->gb;
  cudaStream_t st = h->stream;
  int32_t *ctr = gb.counters.as<int32_t>();
  const size_t npts = (size_t)(n > 0 ? n : 1);
  const size_t npad = (size_t)total_pad;
  NDT_CUDA(h, gb.cell_of.reserve(npts * 4));
  NDT_CUDA(h, gb.rank_of.reserve(npts * 4));
  NDT_CUDA(h, gb.list.reserve(npts * 4));
  NDT_CUDA(h, gb.sorted_idx.reserve(npts * 4));
  NDT_CUDA(h, gb.tgt_sorted.reserve(npts * sizeof(float2)));
  NDT_CUDA(h, gb.slot.reserve(npad * 4));
  NDT_CUDA(h, gb.cen.reserve(npad * sizeof(float2)));
  NDT_CUDA(h, gb.leaf_id.reserve(npad * 4));
  const size_t occ_words = (npad + 31) / 32 + 1;
  NDT_CUDA(h, gb.occ.reserve(occ_words * 4));
  const size_t max_leaves = std::min(npts, npad);
  NDT_CUDA(h, gb.leaf_cell.reserve(max_leaves * 4));
  NDT_CUDA(h, gb.leaf_pair.reserve(max_leaves * 4));
  NDT_CUDA(h, gb.big_list.reserve((npts / FINALIZE_BIG_LEAF + 1) * 4));
  NDT_CUDA(h, gb.leaf_n.reserve(max_leaves * 4));
  NDT_CUDA(h, gb.leaf_start.reserve(max_leaves * 4));
  NDT_CUDA(h, gb.leaf_range.reserve(max_leaves * sizeof(int2)));
  NDT_CUDA(h, gb.leaf_nr.reserve(max_leaves * 4));
  NDT_CUDA(h, gb.leaf_mean.reserve(max_leaves * sizeof(double2)));
  NDT_CUDA(h, gb.leaf_icov.reserve(max_leaves * 4 * sizeof(double)));
  NDT_CUDA(h, gb.leaf_cen.reserve(max_leaves * sizeof(float2)));
  NDT_CUDA(h, gb.recs.reserve(max_leaves * sizeof(CellRec)));

  const PairDims *dims = gb.dims.as<PairDims>();
  const int64_t *off = gb.pair_off.as<int64_t>();
  NDT_CUDA(h, cudaMemsetAsync(gb.occ.p, 0, occ_words * 4, st));
  h->have_nbr = false; h->nbr_cells = (int64_t)npad; h->nbr_grids = n_grids;
  NDT_CUDA(h, cudaMemsetAsync(gb.leaf_id.p, 0, npad * 4, st));               // per-cell counts, then leaf id + 1
  NDT_CUDA(h, cudaMemsetAsync(gb.cen.p, 0xff, npad * sizeof(float2), st));   // all-ones is a NaN: "no tree cell here"
  // enough CTAs to fill the machine: several row chunks per grid when there are few grids
  int chunks = 1;
  if (n_grids < h->sm_count * 8) chunks = std::max(1, std::min(max_h, (h->sm_count * 8) / std::max(n_grids, 1)));
  const bool tiled = n_grids == 1 && total_pad <= TILE_CELLS_CAP && n >= 8192;     // small grid, dense buckets
  if (tiled) {
    const int n_tiles = (int)((n + TILE_PTS - 1) / TILE_PTS), np = (int)total_pad;
    NDT_CUDA(h, gb.tile_hist.reserve((size_t)n_tiles * np * 4));
    k_tile_hist<<<n_tiles, 256, np * 4, st>>>(gb.tgt.as<float4>(), n, dims, h->gd.inv_leaf, np, gb.leaf_id.as<int32_t>(),
                                               gb.cell_of.as<int32_t>(), gb.tile_hist.as<int32_t>());
    k_alloc<<<(unsigned)((int64_t)n_grids * chunks), 256, 0, st>>>(gb.leaf_id.as<int32_t>(), dims, chunks, gb.leaf_cell.as<int32_t>(),
                                                                  gb.leaf_pair.as<int32_t>(), gb.leaf_n.as<int32_t>(),
                                                                  gb.leaf_start.as<int32_t>(), gb.big_list.as<int32_t>(), ctr);
    k_tile_scan<<<(np + 63) / 64, 64, 0, st>>>(gb.tile_hist.as<int32_t>(), n_tiles, np);
    k_tile_place<<<n_tiles, 256, np * 4, st>>>(gb.tgt.as<float4>(), n, np, gb.cell_of.as<int32_t>(), gb.leaf_id.as<int32_t>(),
                                                gb.leaf_start.as<int32_t>(), gb.tile_hist.as<int32_t>(), gb.sorted_idx.as<int32_t>(),
                                                gb.tgt_sorted.as<float2>());
  } else {
    k_count<<<grid_for(n, 256, h->sm_count), 256, 0, st>>>(gb.tgt.as<float4>(), n, off, n_grids, dims, h->gd.inv_leaf,
                                                          gb.leaf_id.as<int32_t>(), gb.cell_of.as<int32_t>(),
                                                          gb.rank_of.as<int32_t>());
    k_alloc<<<(unsigned)((int64_t)n_grids * chunks), 256, 0, st>>>(gb.leaf_id.as<int32_t>(), dims, chunks, gb.leaf_cell.as<int32_t>(),
                                                                  gb.leaf_pair.as<int32_t>(), gb.leaf_n.as<int32_t>(),
                                                                  gb.leaf_start.as<int32_t>(), gb.big_list.as<int32_t>(), ctr);
    k_fill<<<grid_for(n, 256, h->sm_count), 256, 0, st>>>(n, gb.cell_of.as<int32_t>(), gb.rank_of.as<int32_t>(),
                                                         gb.leaf_id.as<int32_t>(), gb.leaf_start.as<int32_t>(),
                                                         gb.list.as<int32_t>());
    k_rank<<<grid_for(n, 256, h->sm_count), 256, 0, st>>>(n, gb.cell_of.as<int32_t>(), gb.leaf_id.as<int32_t>(),
                                                         gb.leaf_start.as<int32_t>(), gb.leaf_n.as<int32_t>(),
                                                         gb.list.as<int32_t>(), gb.tgt.as<float4>(), gb.sorted_idx.as<int32_t>(),
                                                         gb.tgt_sorted.as<float2>(), ctr);
  }
  FinalizeParams fp{h->prm.min_points, h->prm.eig_mult, h->prm.quirks};
  FinalizeOut fo{gb.leaf_range.as<int2>(), gb.leaf_nr.as<int32_t>(), gb.leaf_mean.as<double2>(), gb.leaf_icov.as<double>(),
                 gb.leaf_cen.as<float2>(), gb.slot.as<int32_t>(), gb.cen.as<float2>(), gb.occ.as<uint32_t>(), gb.recs.as<CellRec>(), ctr,
                 gb.leaf_cell.as<int32_t>(), gb.leaf_pair.as<int32_t>(), dims};
  k_finalize<<<std::max(grid_for((int64_t)max_leaves, 128, h->sm_count, 16), 2 * h->sm_count), 128, 0, st>>>(gb.tgt_sorted.as<float2>(), gb.leaf_n.as<int32_t>(),
                                                                                  gb.leaf_start.as<int32_t>(), gb.big_list.as<int32_t>(), fp, fo);
  h->launches += 5;
  NDT_CUDA(h, cudaGetLastError());
  return NDT_OK;
}

// Batched scan pairs: gb.tgt holds every pair's target points back to back, gb.pair_off the device copies of
// the target offsets [0, n_pairs] followed by the source offsets [n_pairs + 1, 2 n_pairs + 1]. Computes every
// pair's grid geometry (gb.dims) and its place in the shared tables; one 16-byte read-back sizes the tables.
int pairs_prepare(Handle *h, int64_t n_pairs, int64_t *total_pad, int *max_h) {
  GridBuffers &gb = h->gb;
  cudaStream_t st = h->stream;
  NDT_CUDA(h, gb.dims.reserve((size_t)n_pairs * sizeof(PairDims)));
  NDT_CUDA(h, h->scratch2.reserve(((size_t)n_pairs + 2) * sizeof(int64_t)));
  NDT_CUDA(h, gb.counters.reserve((CTR_COUNT + 4) * sizeof(int32_t)));
  if (ensure_pinned(h, 256)) return NDT_ERR_CUDA;
  int64_t *pad = h->scratch2.as<int64_t>(), *out = pad + n_pairs;
  const int64_t *off = gb.pair_off.as<int64_t>();
  k_init_counters<<<1, 32, 0, st>>>(gb.counters.as<int32_t>(), gb.counters.as<int32_t>() + CTR_COUNT);
  k_pair_bounds<<<(unsigned)((n_pairs + 7) / 8), 256, 0, st>>>(gb.tgt.as<float4>(), off, off + n_pairs + 1, (int)n_pairs,
                                                              h->gd.inv_leaf, gb.dims.as<PairDims>(), pad);
  k_pair_scan<<<1, 1024, 0, st>>>(gb.dims.as<PairDims>(), pad, (int)n_pairs, out);
  h->launches += 3;
  int64_t *hp = (int64_t *)h->pinned;
  NDT_CUDA(h, cudaMemcpyAsync(hp, out, 16, cudaMemcpyDeviceToHost, st));
  NDT_CUDA(h, cudaStreamSynchronize(st));
  *total_pad = hp[0];
  *max_h = (int)hp[1];
  if (*total_pad > (int64_t)INT_MAX) return set_err(h, NDT_ERR_CAPACITY, "ndt_match_pairs: the pairs' grids exceed 2^31-1 cells in total");
  return NDT_OK;
}


// Finer lattice for the exact 1-NN of the fitness score: built when the NDT buckets are dense (walls: hundreds of points per
// 0.5 m cell), or always (force: factor >= 1) for incrementally maintained targets, whose ordered buckets go stale.
// side = true: the kernels run on the handle's second stream, forked from the main stream here (everything queued on it so
// far -- the upload of the points -- comes first); the caller joins with join_side_stream() before anything reads the lattice.
static int side_stream(Handle *h) {
  if (h->copy_stream) return NDT_OK;
  NDT_CUDA(h, cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
  for (int k = 0; k < 2; ++k) {
    NDT_CUDA(h, cudaEventCreateWithFlags(&h->ev_up[k], cudaEventDisableTiming));
    NDT_CUDA(h, cudaEventCreateWithFlags(&h->ev_done[k], cudaEventDisableTiming));
  }
  return NDT_OK;
}
static int join_side_stream(Handle *h) {
  NDT_CUDA(h, cudaEventRecord(h->ev_done[0], h->copy_stream));
  NDT_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_done[0], 0));
  return NDT_OK;
}

static int build_nn_lattice(Handle *h, int64_t n, const float mn[2], const float mx[2], int64_t nfin, bool force, bool side = false) {
  GridBuffers &gb = h->gb;
  GridDims &gd = h->gd;
  cudaStream_t st = h->stream;
  int32_t *ctr = gb.counters.as<int32_t>();
  const size_t npts = (size_t)(n > 0 ? n : 1);
  gd.nn_f = 0;
  const int64_t n_leaves = h->h_counters[CTR_LEAVES];
  int f = 0;
  if (n_leaves > 0 && nfin / n_leaves > 24) {
    f = 2;
    while (f < 64 && nfin / (n_leaves * f) > 12) f *= 2;                   // walls: points per fine cell fall like 1 / f
  } else if (force && n_leaves > 0) {
    f = 1;
  }
  while (f > 1 && (int64_t)gd.div_x * f * ((int64_t)gd.div_y * f) > (int64_t)8 * 1024 * 1024) f /= 2;   // lattice <= 64 MB
  if (f > 1 || (force && f == 1)) {
    gd.nn_f = f;
    gd.nn_leaf = gd.leaf / (float)f;
    gd.nn_inv_leaf = 1.0f / gd.nn_leaf;
    gd.nn_min_bx = (int)std::floor(mn[0] * gd.nn_inv_leaf); gd.nn_min_by = (int)std::floor(mn[1] * gd.nn_inv_leaf);
    gd.nn_div_x = (int)std::floor(mx[0] * gd.nn_inv_leaf) - gd.nn_min_bx + 1;
    gd.nn_div_y = (int)std::floor(mx[1] * gd.nn_inv_leaf) - gd.nn_min_by + 1;
    const int64_t ncf = (int64_t)gd.nn_div_x * gd.nn_div_y;
    NDT_CUDA(h, gb.cell_of.reserve(npts * 4));
    NDT_CUDA(h, gb.rank_of.reserve(npts * 4));
    NDT_CUDA(h, gb.nn_cnt.reserve((size_t)ncf * 4));
    NDT_CUDA(h, gb.nn_range.reserve((size_t)ncf * sizeof(int2)));
    NDT_CUDA(h, gb.nn_pts.reserve(npts * sizeof(float2)));
    if (side) {
      if (int rc = side_stream(h)) return rc;
      NDT_CUDA(h, cudaEventRecord(h->ev_up[0], st));
      st = h->copy_stream;
      NDT_CUDA(h, cudaStreamWaitEvent(st, h->ev_up[0], 0));
    }
    NDT_CUDA(h, cudaMemsetAsync(gb.nn_cnt.p, 0, (size_t)ncf * 4, st));
    NDT_CUDA(h, cudaMemsetAsync(ctr + CTR_JOB, 0, 4, st));
    Dims df{gd.nn_min_bx, gd.nn_min_by, gd.nn_div_x, gd.nn_div_y, gd.nn_inv_leaf};
    k_nn_count<<<grid_for(n, 256, h->sm_count), 256, 0, st>>>(gb.tgt.as<float4>(), n, df, gb.nn_cnt.as<int32_t>(),
                                                             gb.cell_of.as<int32_t>(), gb.rank_of.as<int32_t>());
    k_nn_alloc<<<grid_for(ncf, 256, h->sm_count), 256, 0, st>>>(gb.nn_cnt.as<int32_t>(), ncf, gb.nn_range.as<int2>(), ctr);
    k_nn_fill<<<grid_for(n, 256, h->sm_count), 256, 0, st>>>(gb.tgt.as<float4>(), n, gb.cell_of.as<int32_t>(),
                                                            gb.rank_of.as<int32_t>(), gb.nn_range.as<int2>(),
                                                            gb.nn_pts.as<float2>());
    h->launches += 3;
    NDT_CUDA(h, cudaGetLastError());
  }
  return NDT_OK;
}

constexpr int64_t INC_MAX_CELLS = 1 << 20;

// exact bounds / count of the finite points among xyzw[lo .. hi)
static void host_bounds(const float *xyzw, int64_t lo, int64_t hi, float mn[2], float mx[2], int64_t &nfin) {
  for (int64_t i = lo; i < hi; ++i) {
    const float x = xyzw[4 * i], y = xyzw[4 * i + 1], z = xyzw[4 * i + 2];
    if (std::isfinite(x) && std::isfinite(y) && std::isfinite(z)) {
      mn[0] = std::min(mn[0], x); mn[1] = std::min(mn[1], y);
      mx[0] = std::max(mx[0], x); mx[1] = std::max(mx[1], y);
      ++nfin;
    }
  }
}

// After a full build of one grid from a host cloud: set up the state incremental updates continue from.
static int inc_init(Handle *h, const float *xyzw, int64_t n, int64_t n_stable, int64_t nfin_all) {
  GridBuffers &gb = h->gb;
  const GridDims &gd = h->gd;
  cudaStream_t st = h->stream;
  h->inc_ok = false; h->inc_active = false;
  const int64_t npad = gd.n_cells > 0 ? (int64_t)(gd.div_x + 4) * (gd.div_y + 4) : 0;
  if (npad == 0 || npad > INC_MAX_CELLS || nfin_all == 0) return NDT_OK;            // not worth it / not possible: stay with full builds
  NDT_CUDA(h, gb.inc_acc.reserve((size_t)npad * sizeof(CellAcc)));
  NDT_CUDA(h, gb.inc_status.reserve((size_t)npad));
  NDT_CUDA(h, gb.inc_mark.reserve((size_t)npad * 4));
  NDT_CUDA(h, gb.inc_lists.reserve((size_t)4 * npad * 4));
  NDT_CUDA(h, gb.inc_lid.reserve((size_t)npad * sizeof(int2)));
  NDT_CUDA(h, gb.inc_cellof.reserve((size_t)5 * INC_BATCH_CAP * 4));          // cell_of | rank_of | order | cell_cnt | cell_start
  NDT_CUDA(h, gb.inc_cnt.reserve(8 * 4));
  NDT_CUDA(h, gb.recs.reserve((size_t)npad * sizeof(CellRec), (size_t)h->h_counters[CTR_SLOTS] * sizeof(CellRec)));   // room for every cell's record
  NDT_CUDA(h, cudaMemsetAsync(gb.inc_acc.p, 0, (size_t)npad * sizeof(CellAcc), st));
  NDT_CUDA(h, cudaMemsetAsync(gb.inc_status.p, 0, (size_t)npad, st));
  NDT_CUDA(h, cudaMemsetAsync(gb.inc_mark.p, 0, (size_t)npad * 4, st));
  NDT_CUDA(h, cudaMemsetAsync(gb.inc_lid.p, 0, (size_t)npad * sizeof(int2), st));
  NDT_CUDA(h, cudaMemsetAsync(gb.inc_cnt.p, 0, 8 * 4, st));
  int32_t *lists = gb.inc_lists.as<int32_t>();
  const int64_t n_leaves = h->h_counters[CTR_LEAVES];
  k_inc_init<<<grid_for(n_leaves * 32, 128, h->sm_count, 16), 128, 0, st>>>(
      gb.tgt_sorted.as<float2>(), gb.sorted_idx.as<int32_t>(), gb.leaf_n.as<int32_t>(), gb.leaf_start.as<int32_t>(), gb.leaf_cell.as<int32_t>(),
      gb.leaf_nr.as<int32_t>(), (int32_t)n_stable, h->prm.min_points, gb.counters.as<int32_t>(), gb.inc_acc.as<CellAcc>(),
      gb.inc_status.as<uint8_t>(), lists + (size_t)INC_P * npad, gb.inc_cnt.as<int32_t>() + INC_P);
  ++h->launches;
  NDT_CUDA(h, cudaGetLastError());
  h->inc_mn[0] = h->inc_mn[1] = std::numeric_limits<float>::max();
  h->inc_mx[0] = h->inc_mx[1] = -std::numeric_limits<float>::max();
  h->inc_nfin = 0;
  host_bounds(xyzw, 0, n_stable, h->inc_mn, h->inc_mx, h->inc_nfin);
  h->inc_m = n_stable;
  h->inc_epoch = 0;
  h->inc_list_prev = INC_P;
  h->inc_ok = true;
  return NDT_OK;
}

int grid_build(Handle *h, const float *xyzw, int64_t n, int memspace, int64_t n_same, int64_t n_stable) {
  GridBuffers &gb = h->gb;
  GridDims &gd = h->gd;
  cudaStream_t st = h->stream;
  h->have_grid = false; h->have_readback = false; h->grid_has_points = false;
  h->inc_ok = false; h->inc_active = false;
  if (memspace != NDT_MEM_HOST) n_stable = -1;      // the settled prefix is tracked with host-side bounds: host clouds only
  // n_same: the caller promises that the first n_same points equal the first n_same points of the previous target of
  // this handle (a map that only changed at its end). They are still on the device: only the rest is staged and copied.
  if (n_same < 0 || n_same > n || n_same > h->tgt_on_device || memspace != NDT_MEM_HOST) n_same = 0;
  h->tgt_on_device = 0;
  if (n < 0 || (n > 0 && !xyzw)) return set_err(h, NDT_ERR_ARG, "ndt_set_target: bad points");
  if (n > (int64_t)INT_MAX) return set_err(h, NDT_ERR_CAPACITY, "ndt_set_target: more than 2^31-1 points");
  const size_t npts = (size_t)(n > 0 ? n : 1);
  NDT_CUDA(h, gb.tgt.reserve(npts * sizeof(float4), (size_t)n_same * sizeof(float4), st));
  NDT_CUDA(h, gb.counters.reserve((CTR_COUNT + 4) * sizeof(int32_t)));
  NDT_CUDA(h, gb.dims.reserve(sizeof(PairDims)));
  NDT_CUDA(h, gb.pair_off.reserve(2 * sizeof(int64_t)));
  int32_t *ctr = gb.counters.as<int32_t>();
  int32_t *bounds = ctr + CTR_COUNT;
  gd = GridDims();
  gd.leaf = h->prm.resolution;
  gd.inv_leaf = 1.0f / gd.leaf;
  gd.r2 = (float)((double)gd.leaf * (double)gd.leaf);
  gd.n_tgt = n;
  std::memset(h->h_counters, 0, sizeof(h->h_counters));

  float mn[2] = {std::numeric_limits<float>::max(), std::numeric_limits<float>::max()};
  float mx[2] = {-std::numeric_limits<float>::max(), -std::numeric_limits<float>::max()};
  int64_t nfin = 0;
  // Host clouds always go through the pinned stage (a copy from pageable memory is several times slower).
  // Small ones get their bounds in the same host pass; large ones are memcpy'd and take the bounds kernel.
  const bool host_in = (memspace == NDT_MEM_HOST);
  const bool host_bounds = host_in && (n < 32768) && n_same == 0;
  const void *copy_from = xyzw;
  if (host_in && n > 0) {
    if (ensure_pinned(h, (npts - (size_t)n_same) * sizeof(float4) + 256)) return NDT_ERR_CUDA;
    float *stage = (float *)h->pinned;
    if (n_same > 0) {
      std::memcpy(stage, xyzw + 4 * n_same, (size_t)(n - n_same) * sizeof(float4));      // the changed tail only
    } else if (host_bounds) {
      for (int64_t i = 0; i < n; ++i) {
        const float x = xyzw[4 * i], y = xyzw[4 * i + 1], z = xyzw[4 * i + 2];
        stage[4 * i] = x; stage[4 * i + 1] = y; stage[4 * i + 2] = z; stage[4 * i + 3] = xyzw[4 * i + 3];
        if (std::isfinite(x) && std::isfinite(y) && std::isfinite(z)) {
          mn[0] = std::min(mn[0], x); mn[1] = std::min(mn[1], y);
          mx[0] = std::max(mx[0], x); mx[1] = std::max(mx[1], y);
          ++nfin;
        }
      }
    } else {
      std::memcpy(stage, xyzw, (size_t)n * sizeof(float4));
    }
    copy_from = stage;
  }
  if (h->timing) cudaEventRecord(h->ev0, st);       // device time only: host staging is not in it
  k_init_counters<<<1, 32, 0, st>>>(ctr, bounds);
  ++h->launches;
  if (n > 0) {
    if (n > n_same)
      NDT_CUDA(h, cudaMemcpyAsync(gb.tgt.as<float4>() + n_same, copy_from, (size_t)(n - n_same) * sizeof(float4),
                                  host_in ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, st));
    if (!host_bounds) {
      k_bounds<<<grid_for(n, 256, h->sm_count), 256, 0, st>>>(gb.tgt.as<float4>(), n, bounds, ctr);
      ++h->launches;
      int32_t *hb = h->pinned_ctr;                  // pinned: a pageable destination makes this a staged, slower copy
      NDT_CUDA(h, cudaMemcpyAsync(hb, ctr, (CTR_COUNT + 4) * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
      NDT_CUDA(h, cudaStreamSynchronize(st));
      nfin = hb[CTR_NFIN];
      if (nfin > 0) {
        mn[0] = ord_float(hb[CTR_COUNT + 0]); mn[1] = ord_float(hb[CTR_COUNT + 1]);
        mx[0] = ord_float(hb[CTR_COUNT + 2]); mx[1] = ord_float(hb[CTR_COUNT + 3]);
      }
      NDT_CUDA(h, cudaMemsetAsync(ctr + CTR_NFIN, 0, 4, st));
    }
  }
  bool empty = (nfin == 0);
  if (!empty) {
    const int64_t dx = (int64_t)((mx[0] - mn[0]) * gd.inv_leaf) + 1;
    const int64_t dy = (int64_t)((mx[1] - mn[1]) * gd.inv_leaf) + 1;
    if (dx * dy > (int64_t)std::numeric_limits<int32_t>::max()) empty = true;  // PCL: "leaf size too small", empty grid
  }
  if (!empty) {
    const int min_bx = (int)std::floor(mn[0] * gd.inv_leaf), min_by = (int)std::floor(mn[1] * gd.inv_leaf);
    const int max_bx = (int)std::floor(mx[0] * gd.inv_leaf), max_by = (int)std::floor(mx[1] * gd.inv_leaf);
    gd.min_bx = min_bx; gd.min_by = min_by;
    gd.div_x = max_bx - min_bx + 1; gd.div_y = max_by - min_by + 1;
    gd.n_cells = (int64_t)gd.div_x * gd.div_y;
    if ((int64_t)(gd.div_x + 4) * (gd.div_y + 4) > (int64_t)INT_MAX)
      return set_err(h, NDT_ERR_CAPACITY, "ndt_set_target: grid exceeds 2^31-1 cells");
    // the matcher takes floor(x / leaf) with one float -> int conversion, exact while cell indices stay below 2^22 in
    // magnitude (at 0.1 m cells: +-419 km); beyond that PCL's own float index arithmetic starts to round as well
    if (std::abs((int64_t)min_bx) >= NDT_MAX_CELL_INDEX || std::abs((int64_t)min_by) >= NDT_MAX_CELL_INDEX ||
        std::abs((int64_t)max_bx) >= NDT_MAX_CELL_INDEX || std::abs((int64_t)max_by) >= NDT_MAX_CELL_INDEX)
      return set_err(h, NDT_ERR_CAPACITY, "ndt_set_target: cell indices beyond +-4,194,304 are not supported");
  }
  if (empty) {
    gd.div_x = gd.div_y = 0; gd.n_cells = 0;
    // a 4 x 4 all-empty padded table (clear occupancy bits): a match against an empty target probes nothing
    NDT_CUDA(h, gb.occ.reserve(64)); NDT_CUDA(h, gb.slot.reserve(64)); NDT_CUDA(h, gb.cen.reserve(128)); NDT_CUDA(h, gb.leaf_id.reserve(64));
    NDT_CUDA(h, gb.nbr.reserve(64)); NDT_CUDA(h, cudaMemsetAsync(gb.nbr.p, 0, 64, st));
    h->have_nbr = true; h->nbr_cells = 0; h->nbr_grids = 0;
    NDT_CUDA(h, cudaMemsetAsync(gb.occ.p, 0, 64, st)); NDT_CUDA(h, cudaMemsetAsync(gb.slot.p, 0xff, 64, st));
    NDT_CUDA(h, cudaMemsetAsync(gb.cen.p, 0xff, 128, st)); NDT_CUDA(h, cudaMemsetAsync(gb.leaf_id.p, 0, 64, st));
    h->have_grid = true; h->have_readback = true; h->grid_has_points = true;
    if (h->timing) { cudaEventRecord(h->ev1, st); }
    NDT_CUDA(h, cudaStreamSynchronize(st));
    if (h->timing) cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1);
    return NDT_OK;
  }
  // one grid: a single PairDims at base 0
  PairDims pd{};
  pd.min_bx = gd.min_bx; pd.min_by = gd.min_by; pd.div_x = gd.div_x; pd.div_y = gd.div_y;
  pd.W = gd.div_x + 4; pd.H = gd.div_y + 4; pd.base = 0; pd.ns = 0; pd.src_off = 0; pd.tgt_off = 0; pd.nt = n;
  NDT_CUDA(h, cudaMemcpyAsync(gb.dims.p, &pd, sizeof(pd), cudaMemcpyHostToDevice, st));
  const int64_t npad_total = (int64_t)pd.W * pd.H;
  if (int rc = grid_build_tables(h, n, 1, npad_total, pd.H)) return rc;
  NDT_CUDA(h, cudaMemcpyAsync(h->pinned_ctr, ctr, sizeof(h->h_counters), cudaMemcpyDeviceToHost, st));
  NDT_CUDA(h, cudaStreamSynchronize(st));
  std::memcpy(h->h_counters, h->pinned_ctr, sizeof(h->h_counters));
  NDT_CUDA(h, cudaGetLastError());
  h->h_counters[CTR_NFIN] = (int32_t)nfin;

  if (int rc = build_nn_lattice(h, n, mn, mx, nfin, /*force=*/n_stable >= 0)) return rc;
  if (n_stable >= 0) { if (int rc = inc_init(h, xyzw, n, std::min(n_stable, n), nfin)) return rc; }
  if (h->timing) {
    cudaEventRecord(h->ev1, st);
    NDT_CUDA(h, cudaEventSynchronize(h->ev1));
    cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1);
  }
  NDT_CUDA(h, cudaGetLastError());
  h->have_grid = true; h->have_readback = true; h->grid_has_points = true;
  h->tgt_on_device = n;
  return NDT_OK;
}


// End of an incremental update whose caller did not wait (ndt_set_target_incremental_async): wait for the stream, take the
// counters the kernels left in pinned memory. Every entry point of the ABI comes through here first.
int finish_pending(Handle *h) {
  if (!h->pending) return NDT_OK;
  h->pending = false;
  NDT_CUDA(h, cudaStreamSynchronize(h->stream));
  std::memcpy(h->h_counters, h->pinned_ctr, sizeof(h->h_counters));
  h->h_counters[CTR_NFIN] = (int32_t)h->pending_nfin;
  h->h_counters[CTR_PTS] = (int32_t)h->pending_nfin;
  if (h->timing) cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1);
  return NDT_OK;
}

int grid_build_incremental(Handle *h, const float *xyzw, int64_t n, int64_t n_same, int64_t n_stable, int memspace, bool defer) {
  if (n < 0 || (n > 0 && !xyzw)) return set_err(h, NDT_ERR_ARG, "ndt_set_target_incremental: bad points");
  if (n_stable < 0) n_stable = 0;
  if (n_stable > n) n_stable = n;
  if (n_same < 0 || n_same > n || n_same > h->tgt_on_device) n_same = 0;
  const int64_t m = h->inc_m;
  // anything the running sums cannot continue from is a full build (which re-initialises them)
  const bool can = memspace == NDT_MEM_HOST && h->inc_ok && h->have_grid && n_same >= m && n_stable >= m && n <= (int64_t)INT_MAX &&
                   n - m <= INC_BATCH_CAP && n > 0;
  if (!can) return grid_build(h, xyzw, n, memspace, n_same, n_stable);
  GridBuffers &gb = h->gb;
  GridDims &gd = h->gd;
  cudaStream_t st = h->stream;
  // bounds of the whole cloud = settled prefix (cached) + newly settled points + tail, exact floats like getMinMax3D
  float mnA[2] = {h->inc_mn[0], h->inc_mn[1]}, mxA[2] = {h->inc_mx[0], h->inc_mx[1]};
  int64_t nfinA = h->inc_nfin;
  host_bounds(xyzw, m, n_stable, mnA, mxA, nfinA);
  float mn[2] = {mnA[0], mnA[1]}, mx[2] = {mxA[0], mxA[1]};
  int64_t nfin = nfinA;
  host_bounds(xyzw, n_stable, n, mn, mx, nfin);
  bool same_geom = nfin > 0;
  if (same_geom) {
    const int64_t dx = (int64_t)((mx[0] - mn[0]) * gd.inv_leaf) + 1, dy = (int64_t)((mx[1] - mn[1]) * gd.inv_leaf) + 1;
    if (dx * dy > (int64_t)std::numeric_limits<int32_t>::max()) same_geom = false;
  }
  if (same_geom) {
    const int min_bx = (int)std::floor(mn[0] * gd.inv_leaf), min_by = (int)std::floor(mn[1] * gd.inv_leaf);
    const int max_bx = (int)std::floor(mx[0] * gd.inv_leaf), max_by = (int)std::floor(mx[1] * gd.inv_leaf);
    same_geom = min_bx == gd.min_bx && min_by == gd.min_by && max_bx - min_bx + 1 == gd.div_x && max_by - min_by + 1 == gd.div_y;
  }
  if (!same_geom) return grid_build(h, xyzw, n, memspace, n_same, n_stable);      // the map grew past the grid: every index shifts

  h->have_grid = false; h->have_readback = false; h->have_nbr = false;
  const int64_t npad = (int64_t)(gd.div_x + 4) * (gd.div_y + 4);
  const size_t npts = (size_t)n;
  NDT_CUDA(h, gb.tgt.reserve(npts * sizeof(float4), (size_t)n_same * sizeof(float4), st));
  const int B = (int)(n - m);                        // this call's batch: newly settled points, then the tail
  if (n > n_same) {
    if (ensure_pinned(h, (size_t)(n - n_same) * sizeof(float4) + 256)) return NDT_ERR_CUDA;
    std::memcpy(h->pinned, xyzw + 4 * n_same, (size_t)(n - n_same) * sizeof(float4));
  }
  if (h->timing) cudaEventRecord(h->ev0, st);
  if (n > n_same)
    NDT_CUDA(h, cudaMemcpyAsync(gb.tgt.as<float4>() + n_same, h->pinned, (size_t)(n - n_same) * sizeof(float4), cudaMemcpyHostToDevice, st));
  const int32_t epoch = ++h->inc_epoch;
  int32_t *lists = gb.inc_lists.as<int32_t>(), *cnt = gb.inc_cnt.as<int32_t>(), *mark = gb.inc_mark.as<int32_t>();
  int32_t *ctr = gb.counters.as<int32_t>();
  // list buffers: the previous call's cells sit in buffer inc_list_prev; this call's batch cells and the union take two others
  int buf[3], k = 0;
  for (int b = 0; b < 4; ++b) if (b != h->inc_list_prev) buf[k++] = b;
  int32_t *LB = lists + (size_t)buf[0] * npad, *LU = lists + (size_t)buf[1] * npad, *LP = lists + (size_t)h->inc_list_prev * npad;
  const PairDims *dims = gb.dims.as<PairDims>();
  const float4 *pts = gb.tgt.as<float4>();
  int32_t *cell_of = gb.inc_cellof.as<int32_t>(), *rank_of = cell_of + INC_BATCH_CAP, *order = rank_of + INC_BATCH_CAP,
          *cell_cnt = order + INC_BATCH_CAP, *cell_start = cell_cnt + INC_BATCH_CAP;
  FinalizeParams fp{h->prm.min_points, h->prm.eig_mult, h->prm.quirks};
  IncTables T{gb.inc_status.as<uint8_t>(), gb.slot.as<int32_t>(), gb.cen.as<float2>(), gb.recs.as<CellRec>(), ctr};
  IncBatch ib{pts, m, n_stable, B, dims, gd.inv_leaf, epoch, gd.div_x + 4, cell_of, rank_of, order, cell_cnt, cell_start,
              gb.inc_lid.as<int2>(), LB, LP, LU, cnt, mark, gb.inc_acc.as<CellAcc>(), gb.occ.as<uint32_t>()};
  gd.n_tgt = n;
  // The exact 1-NN lattice only depends on the points: it is rebuilt on the second stream while k_inc_update runs. It is
  // sized from the leaf count of the previous call (h_counters: a handful of leaves off at most, and only the lattice pitch
  // depends on it), so the whole update needs one stream synchronisation, at its end.
  if (int rc = build_nn_lattice(h, n, mn, mx, nfin, /*force=*/true, /*side=*/true)) return rc;
  const bool forked = gd.nn_f > 0;
  k_inc_update<<<1, INC_THREADS, 0, st>>>(ib, fp, T);
  ++h->launches;
  if (forked) { if (int rc = join_side_stream(h)) return rc; }
  h->inc_list_prev = buf[0];
  NDT_CUDA(h, cudaGetLastError());
  if (h->timing) cudaEventRecord(h->ev1, st);
  // counters -> host (also the stream fence that lets the pinned stage be reused)
  NDT_CUDA(h, cudaMemcpyAsync(h->pinned_ctr, ctr, sizeof(h->h_counters), cudaMemcpyDeviceToHost, st));
  h->pending_nfin = nfin;
  h->pending = true;                                 // the host copy of the counters is taken by finish_pending()
  if (!defer) { if (int rc = finish_pending(h)) return rc; }
  h->inc_mn[0] = mnA[0]; h->inc_mn[1] = mnA[1]; h->inc_mx[0] = mxA[0]; h->inc_mx[1] = mxA[1];
  h->inc_nfin = nfinA;
  h->inc_m = n_stable;
  h->inc_active = true;
  h->nbr_cells = npad; h->nbr_grids = 1;
  h->have_grid = true; h->grid_has_points = true;
  h->tgt_on_device = n;
  return NDT_OK;
}

int grid_cell_index(Handle *h, const float *xyzw, int64_t n, int memspace, int32_t *idx_out) {
  if (!h->have_grid) return set_err(h, NDT_ERR_STATE, "ndt_cell_index: no target set");
  if (n <= 0) return NDT_OK;
  cudaStream_t st = h->stream;
  NDT_CUDA(h, h->scratch.reserve((size_t)n * sizeof(float4)));
  NDT_CUDA(h, h->scratch2.reserve((size_t)n * 4));
  const float4 *d_in = reinterpret_cast<const float4 *>(xyzw);
  if (memspace == NDT_MEM_HOST) {
    NDT_CUDA(h, cudaMemcpyAsync(h->scratch.p, xyzw, (size_t)n * sizeof(float4), cudaMemcpyHostToDevice, st));
    d_in = h->scratch.as<float4>();
  }
  Dims d{h->gd.min_bx, h->gd.min_by, h->gd.div_x, h->gd.div_y, h->gd.inv_leaf};
  k_cell_index<<<grid_for(n, 256, h->sm_count), 256, 0, st>>>(d_in, n, d, h->scratch2.as<int32_t>());
  ++h->launches;
  // idx_out is always a host buffer (parity hook)
  NDT_CUDA(h, cudaMemcpyAsync(idx_out, h->scratch2.p, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
  NDT_CUDA(h, cudaStreamSynchronize(st));
  return NDT_OK;
}

int ensure_nbr(Handle *h) {
  if (h->have_nbr) return NDT_OK;
  GridBuffers &gb = h->gb;
  if (h->nbr_cells > 0) {
    NDT_CUDA(h, gb.nbr.reserve((size_t)h->nbr_cells * sizeof(uint16_t)));
    k_nbr_from_cen<<<grid_for(h->nbr_cells, 256, h->sm_count), 256, 0, h->stream>>>(gb.cen.as<float2>(), gb.dims.as<PairDims>(), h->nbr_grids,
                                                                                    h->nbr_cells, gb.nbr.as<uint16_t>());
    ++h->launches;
    NDT_CUDA(h, cudaGetLastError());
  } else {
    NDT_CUDA(h, gb.nbr.reserve(64));
    NDT_CUDA(h, cudaMemsetAsync(gb.nbr.p, 0, 64, h->stream));
  }
  h->have_nbr = true;
  return NDT_OK;
}

GridView grid_view(const Handle *h) {
  GridView G{};
  const GridBuffers &gb = h->gb;
  G.slot = gb.slot.as<int32_t>();
  G.slot_w = h->gd.div_x + 4;
  G.table_base = 0;
  G.cen = gb.cen.as<float2>();
  G.occ = gb.occ.as<uint32_t>();
  G.nbr = gb.nbr.as<uint16_t>();
  G.recs = gb.recs.as<CellRec>();
  G.min_bx = h->gd.min_bx; G.min_by = h->gd.min_by; G.div_x = h->gd.div_x; G.div_y = h->gd.div_y;
  G.inv_leaf = h->gd.inv_leaf; G.r2 = h->gd.r2; G.leaf = h->gd.leaf;
  // after an incremental update only the lattice is current: no ordered buckets (the 1-NN then stays on the lattice)
  G.leaf_id = h->inc_active ? nullptr : gb.leaf_id.as<int32_t>();
  G.leaf_range = h->inc_active ? nullptr : gb.leaf_range.as<int2>();
  G.tgt_sorted = h->inc_active ? nullptr : gb.tgt_sorted.as<float2>();
  G.tgt = gb.tgt.as<float4>();
  G.n_tgt = h->gd.n_tgt;
  G.nn_f = h->gd.nn_f;
  G.nn_min_bx = h->gd.nn_min_bx; G.nn_min_by = h->gd.nn_min_by; G.nn_div_x = h->gd.nn_div_x; G.nn_div_y = h->gd.nn_div_y;
  G.nn_inv_leaf = h->gd.nn_inv_leaf; G.nn_leaf = h->gd.nn_leaf;
  G.nn_range = gb.nn_range.as<int2>();
  G.nn_pts = gb.nn_pts.as<float2>();
  return G;
}

MatchParams match_params(const Handle *h, bool want_fitness) {
  MatchParams mp{};
  // PCL computeTransformation; resolution_ is a float member
  const double r = (double)h->prm.resolution;
  const double c1 = 10.0 * (1.0 - h->prm.outlier_ratio);
  const double c2 = h->prm.outlier_ratio / std::pow(r, 3);
  const double d3 = -std::log(c2);
  mp.d1 = -std::log(c1 + c2) - d3;
  mp.d2 = -2.0 * std::log((-std::log(c1 * std::exp(-0.5) + c2) - d3) / mp.d1);
  mp.step_size = h->prm.step_size;
  mp.step_min = h->prm.trans_eps / 2;
  mp.trans_eps = h->prm.trans_eps;
  mp.max_iter = h->prm.max_iter;
  mp.quirks = h->prm.quirks;
  mp.want_fitness = want_fitness ? 1 : 0;
  return mp;
}

}  // namespace ndt
