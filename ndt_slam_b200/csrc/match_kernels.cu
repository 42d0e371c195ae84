// match_kernels.cu -- NDT objective evaluation and the persistent Newton / More-Thuente matcher (sm_100a).
//
// Replaces, for z = 0 data, pcl::NormalDistributionsTransform::{computeDerivatives, computeHessian,
// computeTransformation, computeStepLengthMT} and Registration::getFitnessScore as called by
// PoseEstimator::estimatePose [REF src/PoseEstimator.cpp:28, 43, 56]. The whole optimisation runs
// inside one kernel launch: no host round trip per iteration.
//
// Kernels
//   k_eval_partial / k_eval_final  one objective pass for a few poses (parity hook)
//   k_eval_warp     one objective pass for thousands of poses, one warp per pose (relocalisation score sweep)
//   k_align_block   one CTA per match, optional shared-memory tile of the whole grid (single scans)
//   k_align_cluster one thread-block cluster per match, DSMEM reduction (single scans)
//   k_align_grid    one match on every SM, cooperative launch + grid-wide reduction (very large source clouds)
//   k_align_warp    persistent CTAs, one warp per match pulled from an atomic work counter (batches)
//   k_align_pairs / k_align_pairs_block  batched scan pairs, every job with its own grid: warp per pair / CTA per pair
//   k_best_of       arg-max of the batch results
//   k_voxel_filter(_pairs)  ApproximateVoxelGrid, one warp per cloud (slot-parallel replay of the hash history)
#include "ndt_host.h"

#include <cooperative_groups.h>
#include <algorithm>
#include <cstdio>
#include <type_traits>
#include <cstdlib>
#include <cstring>

namespace cg = cooperative_groups;

namespace ndt {

namespace {

struct GlobalSrc {
  const float4 *__restrict__ p;
  __device__ __forceinline__ float2 operator()(int i) const {
    const float4 v = __ldg(p + i);   // coalesced 16-byte loads
    return make_float2(v.x, v.y);
  }
};
struct SmemSrc {
  uint32_t a;       // shared address of float2[ns]
  __device__ __forceinline__ float2 operator()(int i) const { return lds_float2(a + 8u * i); }
};

// One cooperative objective pass. All threads of the group call pass<MODE>() with identical
// arguments and leave with identical totals. Everything the hot loop touches is copied into
// locals first (the object itself lives in local memory behind `this`).
template <class Coop, class OccL, class NbrL, class CenL, class SlotL, class RecL, class SrcL>
struct Objective {
  ProbeGeom geom;
  OccL occ;
  NbrL nbr;
  CenL cen;
  SlotL slot;
  RecL rec;
  SrcL src;
  int ns;
  double d1, d2;
  int sse;
  Coop coop;
  HitQueue Q;          // this warp's hit queue (shared memory)

#ifndef NDT_PHASE_CLOCK
#define NDT_PHASE_CLOCK 0           // 1: per-phase clock64() totals of block 0 / thread 0 printed at kernel end (diagnostic builds only)
#endif
#if NDT_PHASE_CLOCK
  long long t_enter = 0, t_leave = 0, c_pose = 0, c_acc = 0, c_red = 0, c_between = 0, n_pass = 0;
#endif
  __device__ __noinline__ void pass(const int mode_asked, const double *p, const AngleCache &ac, double *out) {
    // the optimiser only ever asks for mode 0 when its trial passes carry the Hessian (NDT_TRIAL_MODE 0, ndt_device.cuh):
    // the mode is then a compile-time constant and the Hessian-only / gradient-only predication drops out of the hot loop
    const int mode = NDT_TRIAL_MODE == 0 ? 0 : mode_asked;
#if NDT_PHASE_CLOCK
    t_enter = clock64();
    if (t_leave) c_between += t_enter - t_leave;
    ++n_pass;
#endif
    double acc[NACC];
#pragma unroll
    for (int k = 0; k < NACC; ++k) acc[k] = 0.0;
    const PoseF pf = pose_to_float(p);
#if NDT_PHASE_CLOCK
    const long long t1 = clock64(); c_pose += t1 - t_enter;
#endif
    const double cs = ac.cs, sn = ac.sn;
    const Coop co = coop;
    int pairs = 0;
    accumulate_points(mode, geom, occ, nbr, cen, slot, rec, src, co.rank(), co.size(), ns, pf, sse != 0, cs, sn, d1, d2,
                      Q, acc, pairs);
#if NDT_PHASE_CLOCK
    const long long t2 = clock64(); c_acc += t2 - t1;
#endif
    if (mode == 0) co.template allreduce<13>(acc);
    else if (mode == 1) co.template allreduce<4>(acc);
    else co.template allreduce<9>(acc + 4);
#pragma unroll
    for (int k = 0; k < NACC; ++k) out[k] = acc[k];     // identical in every cooperating thread (fixed-order reduction)
#if NDT_PHASE_CLOCK
    t_leave = clock64(); c_red += t_leave - t2;
#endif
  }
  __device__ void report(const char *who, long long t_fit) const {
#if NDT_PHASE_CLOCK
    if (blockIdx.x == 0 && threadIdx.x == 0)
      printf("PHASE %s passes %lld: pose_to_float %lld  accumulate %lld  allreduce %lld  optimiser(between passes) %lld  fitness %lld  [cycles per pass: %lld %lld %lld %lld]\n",
             who, n_pass, c_pose, c_acc, c_red, c_between, t_fit, c_pose / n_pass, c_acc / n_pass, c_red / n_pass, c_between / (n_pass > 1 ? n_pass - 1 : 1));
#endif
  }
};

template <class Coop, class OccL, class NbrL, class CenL, class SlotL, class RecL, class SrcL>
__device__ __forceinline__ Objective<Coop, OccL, NbrL, CenL, SlotL, RecL, SrcL> make_objective(
    const GridView &G, const MatchParams &mp, const Coop &coop, OccL occ, NbrL nbr, CenL cen, SlotL slot, RecL rec, SrcL src, int ns,
    HitQueue Q) {
  return Objective<Coop, OccL, NbrL, CenL, SlotL, RecL, SrcL>{probe_geom(G), occ, nbr, cen, slot, rec, src, ns, mp.d1, mp.d2,
                                                  (mp.quirks & NDT_QUIRK_TRANSFORM_SSE_ORDER) ? 1 : 0, coop, Q};
}

// per-warp rings live at the start of the CTA's dynamic shared memory: [nw][QCAP] int2 hits | [nw][CQCAP] int2 candidates
constexpr int QUEUE_BYTES = 8 * QUEUE_BYTES_PER_WARP;          // 8 warps per CTA: 28,672 B
__device__ __forceinline__ HitQueue my_queue(unsigned char *smem) {
  const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const uint32_t base = smem_addr(smem);
  return HitQueue{base + 8u * (w * QCAP), base + 8u * (nw * QCAP + w * CQCAP)};
}

template <class Coop, class SrcL>
__device__ inline double fitness_pass(const GridView &G, const SrcL &src, int ns, const MatchParams &mp,
                                      const double *p, const Coop &coop) {
  const PoseF pf = pose_to_float(p);
  const bool sse = (mp.quirks & NDT_QUIRK_TRANSFORM_SSE_ORDER) != 0;
  double sum[1] = {0.0};
  const int lane = threadIdx.x & 31;
  for (int i0 = coop.rank() - lane; i0 < ns; i0 += coop.size()) {        // warp-uniform trip count: the 1-NN is warp-collective
    const int i = i0 + lane;
    const bool valid = i < ns;
    const float2 xy = src(valid ? i : ns - 1);
    float xt, yt;
    xform(pf, sse, xy.x, xy.y, xt, yt);
    const float d = nn_dist2_warp(G, xt, yt, valid);
    if (valid) sum[0] += (double)d;
  }
  coop.template allreduce<1>(sum);
  return sum[0];
}

// k_align_warp rarely (relocalisation: never) runs the fitness pass: kept out of line there, so the 1-NN search does not
// sit in the instruction stream of the persistent hot loop (C4 11.5 -> 11.25 ms); the single-match kernels inline it (C1
// 0.065 vs 0.069 ms).
template <class Coop, class SrcL>
__device__ __noinline__ double fitness_pass_cold(const GridView &G, const SrcL &src, int ns, const MatchParams &mp, const double *p,
                                                 const Coop &coop) {
  return fitness_pass(G, src, ns, mp, p, coop);
}

__device__ inline void write_result(ndt_result *out, const MatchOut &mo, int ns, double fitness_sum,
                                    bool have_fitness, int64_t n_tgt) {
  ndt_result r;
  r.pose[0] = mo.p[0]; r.pose[1] = mo.p[1]; r.pose[2] = mo.p[2];
  const PoseF pf = pose_to_float(mo.p);
#pragma unroll
  for (int k = 0; k < 16; ++k) r.T[k] = 0.f;
  r.T[0] = pf.c; r.T[1] = pf.s; r.T[4] = -pf.s; r.T[5] = pf.c; r.T[10] = 1.f; r.T[15] = 1.f;
  r.T[12] = pf.tx; r.T[13] = pf.ty;
  r.score = mo.score;
  r.trans_prob = ns ? mo.score / (double)ns : 0.0;
  if (have_fitness) r.fitness = (ns > 0 && n_tgt > 0) ? fitness_sum / (double)ns : DBL_MAX;
  else r.fitness = nan("");
#pragma unroll
  for (int k = 0; k < 9; ++k) r.hess[k] = mo.H[k];
  r.converged = mo.converged; r.iters = mo.iters; r.evals = mo.evals; r.passes_run = mo.passes;
  r.point_evals = (int64_t)mo.evals * (int64_t)ns;
  *out = r;
}

// ---------------------------------------------------------------------------------------------
// objective only
// ---------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(256) k_eval_partial(GridView G, MatchParams mp, const float4 *__restrict__ src,
                                                     int ns, const double *__restrict__ poses, int slices,
                                                     double *__restrict__ partial, int *__restrict__ pairs_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double scratch[9 * NACC];
  __shared__ int s_pairs[8];
  const int pose_i = blockIdx.x / slices, slice = blockIdx.x % slices;
  const double p[3] = {poses[3 * pose_i], poses[3 * pose_i + 1], poses[3 * pose_i + 2]};
  AngleCache ac;
  angle_terms(mp, p[2], ac);
  const PoseF pf = pose_to_float(p);
  const bool sse = (mp.quirks & NDT_QUIRK_TRANSFORM_SSE_ORDER) != 0;
  double acc[NACC];
#pragma unroll
  for (int k = 0; k < NACC; ++k) acc[k] = 0.0;
  int pairs = 0;
  // slice s owns points [s * chunk, (s + 1) * chunk)
  const int chunk = (ns + slices - 1) / slices;
  const int lo = slice * chunk, hi = min(ns, lo + chunk);
  accumulate_points(MODE, probe_geom(G), GlobalOcc{G.occ}, NoNbr{}, GlobalCen{G.cen}, GlobalSlot{G.slot}, GlobalRec{G.recs}, GlobalSrc{src},
                          lo + (int)threadIdx.x, (int)blockDim.x, hi, pf, sse, ac.cs, ac.sn, mp.d1, mp.d2, my_queue(smem_raw),
                          acc, pairs);
  BlockCoop coop{scratch};
  coop.allreduce<NACC>(acc);
  if ((threadIdx.x & 31) == 0) s_pairs[threadIdx.x >> 5] = pairs;   // warp-uniform count
  __syncthreads();
  if (threadIdx.x == 0) {
    int tp = 0;
    for (int w = 0; w < (blockDim.x >> 5); ++w) tp += s_pairs[w];
    double *o = partial + (size_t)blockIdx.x * NACC;
#pragma unroll
    for (int k = 0; k < NACC; ++k) o[k] = acc[k];
    pairs_out[blockIdx.x] = tp;
  }
}

__global__ void k_eval_final(const double *__restrict__ partial, const int *__restrict__ pairs_in, int slices,
                             int64_t n_poses, double *__restrict__ out14, int64_t *__restrict__ pairs_out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_poses) return;
  double acc[NACC];
#pragma unroll
  for (int k = 0; k < NACC; ++k) acc[k] = 0.0;
  int64_t tp = 0;
  for (int s = 0; s < slices; ++s) {
    const double *q = partial + ((size_t)i * slices + s) * NACC;
#pragma unroll
    for (int k = 0; k < NACC; ++k) acc[k] += q[k];
    tp += pairs_in[(size_t)i * slices + s];
  }
#pragma unroll
  for (int k = 0; k < NACC; ++k) out14[(size_t)i * (NACC + 1) + k] = acc[k];
  out14[(size_t)i * (NACC + 1) + NACC] = (double)tp;
  if (pairs_out) pairs_out[i] = tp;
}

// ---------------------------------------------------------------------------------------------
// one CTA per match
// ---------------------------------------------------------------------------------------------
template <bool TILE>
__global__ void __launch_bounds__(256) k_align_block(GridView G, MatchParams mp, const float4 *__restrict__ src,
                                                    int ns, const double *__restrict__ guesses,
                                                    ndt_result *__restrict__ out, int n_slots, int n_cells) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double scratch[9 * NACC];
  BlockCoop coop{scratch};
  const int job = blockIdx.x;
  const double guess[3] = {guesses[3 * job], guesses[3 * job + 1], guesses[3 * job + 2]};
  const GlobalSrc gsrc{src};
  MatchOut mo;
  OptState opt;
  if (TILE) {
    // stage the local map tile (here: the whole grid) in shared memory: records first (64-B aligned), then slots
    CellRec *s_recs = reinterpret_cast<CellRec *>(smem_raw + QUEUE_BYTES);
    float2 *s_cen = reinterpret_cast<float2 *>(smem_raw + QUEUE_BYTES + (size_t)n_slots * sizeof(CellRec));
    int32_t *s_slot = reinterpret_cast<int32_t *>(s_cen + n_cells);
    uint32_t *s_occ = reinterpret_cast<uint32_t *>(s_slot + n_cells);
    for (int i = threadIdx.x; i < (n_cells + 31) / 32 + 1; i += blockDim.x) s_occ[i] = __ldg(G.occ + i);
    const int4 *gr = reinterpret_cast<const int4 *>(G.recs);
    int4 *sr = reinterpret_cast<int4 *>(s_recs);
    for (int i = threadIdx.x; i < n_slots * 4; i += blockDim.x) sr[i] = __ldg(gr + i);
    for (int i = threadIdx.x; i < n_cells; i += blockDim.x) { s_cen[i] = __ldg(G.cen + i); s_slot[i] = __ldg(G.slot + i); }
    __syncthreads();
    auto obj = make_objective(G, mp, coop, SmemOcc{smem_addr(s_occ)}, NoNbr{}, SmemCen{smem_addr(s_cen)}, SmemSlot{smem_addr(s_slot)}, SmemRec{smem_addr(s_recs)}, gsrc, ns, my_queue(smem_raw));
    match_device(obj, mp, guess, mo, opt);
    obj.report("block<tile>", 0);
  } else {
    auto obj = make_objective(G, mp, coop, GlobalOcc{G.occ}, NoNbr{}, GlobalCen{G.cen}, GlobalSlot{G.slot}, GlobalRec{G.recs}, gsrc, ns, my_queue(smem_raw));
    match_device(obj, mp, guess, mo, opt);
    obj.report("block", 0);
  }
  double fsum = 0.0;
  if (mp.want_fitness) fsum = fitness_pass(G, gsrc, ns, mp, mo.p, coop);
  if (threadIdx.x == 0) write_result(out + job, mo, ns, fsum, mp.want_fitness != 0, G.n_tgt);
}

// ---------------------------------------------------------------------------------------------
// one thread-block cluster per match: block reduction, then a DSMEM exchange of the 13 partials
// ---------------------------------------------------------------------------------------------
#ifndef NDT_CLUSTER_SIZE
#define NDT_CLUSTER_SIZE 8
#endif
struct ClusterCoop {
  double *scratch;       // [9 * NACC] block scratch (one row per warp + the block totals)
  double *xchg;          // [2][16] this CTA's partial (double-buffered), read by every CTA of the cluster through DSMEM
  int *epoch;            // reductions so far (selects the exchange buffer)
  int crank, csize;
  __device__ __forceinline__ int rank() const { return crank * blockDim.x + threadIdx.x; }
  __device__ __forceinline__ int size() const { return csize * blockDim.x; }
  // One cluster barrier per reduction: thread k forms the block partial of component k, publishes it, and after the
  // barrier adds the partials of all CTAs in rank order (the same order everywhere: bit-identical totals). The exchange
  // buffer alternates, so a CTA can run ahead into the next reduction while a neighbour still reads this one.
  template <int N> __device__ __forceinline__ void allreduce(double *v) const {
    cg::cluster_group cluster = cg::this_cluster();
    warp_allreduce<N>(v);
    const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    double *total = scratch + nw * NACC;
    double *xb = xchg + ((*epoch) & 1) * 16;
    ++(*epoch);
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
      for (int k = 0; k < N; ++k) scratch[w * NACC + k] = v[k];
    }
    __syncthreads();
    if (threadIdx.x < N) {
      double s = scratch[threadIdx.x];
      for (int i = 1; i < nw; ++i) s += scratch[i * NACC + threadIdx.x];
      xb[threadIdx.x] = s;
    }
    cluster.sync();
    if (threadIdx.x < N) {
      // all remote partials are fetched before the first add: one DSMEM round trip instead of csize dependent ones
      double part[NDT_CLUSTER_SIZE];
#pragma unroll
      for (int r = 0; r < NDT_CLUSTER_SIZE; ++r) part[r] = r < csize ? cluster.map_shared_rank(xb, r)[threadIdx.x] : 0.0;
      double t = 0.0;
#pragma unroll
      for (int r = 0; r < NDT_CLUSTER_SIZE; ++r) if (r < csize) t += part[r];      // rank order, like before: bit-identical totals
      total[threadIdx.x] = t;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < N; ++k) v[k] = total[k];
  }
};

__global__ void __launch_bounds__(256) k_align_cluster(GridView G, MatchParams mp, const float4 *__restrict__ src,
                                                      int ns, const double *__restrict__ guesses,
                                                      ndt_result *__restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double scratch[9 * NACC];
  __shared__ double xchg[2 * 16];
  int epoch = 0;
  cg::cluster_group cluster = cg::this_cluster();
  ClusterCoop coop{scratch, xchg, &epoch, (int)cluster.block_rank(), (int)cluster.num_blocks()};
  const int job = blockIdx.x / coop.csize;
  const double guess[3] = {guesses[3 * job], guesses[3 * job + 1], guesses[3 * job + 2]};
  const GlobalSrc gsrc{src};
  MatchOut mo;
  OptState opt;
  auto obj = make_objective(G, mp, coop, GlobalOcc{G.occ}, NoNbr{}, GlobalCen{G.cen}, GlobalSlot{G.slot}, GlobalRec{G.recs}, gsrc, ns, my_queue(smem_raw));
  match_device(obj, mp, guess, mo, opt);
  double fsum = 0.0;
#if NDT_PHASE_CLOCK
  const long long tf0 = clock64();
#endif
  if (mp.want_fitness) fsum = fitness_pass(G, gsrc, ns, mp, mo.p, coop);
#if NDT_PHASE_CLOCK
  obj.report("cluster", clock64() - tf0);
#endif
  if (coop.crank == 0 && threadIdx.x == 0) write_result(out + job, mo, ns, fsum, mp.want_fitness != 0, G.n_tgt);
  cluster.sync();        // no CTA may retire while a neighbour can still read its exchange buffer through DSMEM
}

// ---------------------------------------------------------------------------------------------
// one match on the whole GPU (very large source clouds, C3): cooperative launch, one grid-wide barrier per objective
// pass. Every CTA publishes its 13 block partials, grid.sync(), then every CTA sums all partials in the same fixed order
// -- all threads of the grid hold bit-identical totals and run the optimiser redundantly, like the other matchers.
// ---------------------------------------------------------------------------------------------
struct GridCoop {
  double *scratch;       // [16 * 16] block scratch (shared memory)
  double *gpart;         // [2][gridDim.x][16] partials of every CTA (global), double-buffered
  int *epoch;            // this thread's count of reductions so far (selects the buffer)
  __device__ __forceinline__ int rank() const { return blockIdx.x * blockDim.x + threadIdx.x; }
  __device__ __forceinline__ int size() const { return gridDim.x * blockDim.x; }
  template <int N> __device__ __forceinline__ void allreduce(double *v) const {
    BlockCoop b{scratch};
    b.allreduce<N>(v);
    double *buf = gpart + (size_t)((*epoch) & 1) * gridDim.x * 16;
    ++(*epoch);
#pragma unroll
    for (int k = 0; k < N; ++k)
      if (threadIdx.x == k) buf[blockIdx.x * 16 + k] = v[k];
    __threadfence();
    cg::this_grid().sync();
    // thread (k, j): component k over CTAs j, j + 16, ... ; then the 16 strided sums are added in order
    const int k = threadIdx.x & 15, j = threadIdx.x >> 4;
    double s = 0.0;
    if (k < N) for (int c = j; c < (int)gridDim.x; c += 16) s += __ldcg(buf + c * 16 + k);
    scratch[j * 16 + k] = s;
    __syncthreads();
#pragma unroll
    for (int q = 0; q < N; ++q) {
      double t = scratch[q];
#pragma unroll
      for (int i = 1; i < 16; ++i) t += scratch[i * 16 + q];
      v[q] = t;
    }
    __syncthreads();
  }
};

__global__ void __launch_bounds__(256, 2) k_align_grid(GridView G, MatchParams mp, const float4 *__restrict__ src, int ns,
                                                    const double *__restrict__ guess3, ndt_result *__restrict__ out,
                                                    double *__restrict__ gpart) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double scratch[16 * 16];
  int epoch = 0;
  GridCoop coop{scratch, gpart, &epoch};
  const double guess[3] = {guess3[0], guess3[1], guess3[2]};
  const GlobalSrc gsrc{src};
  MatchOut mo;
  OptState opt;
  auto obj = make_objective(G, mp, coop, GlobalOcc{G.occ}, NoNbr{}, GlobalCen{G.cen}, GlobalSlot{G.slot}, GlobalRec{G.recs}, gsrc, ns, my_queue(smem_raw));
  match_device(obj, mp, guess, mo, opt);
  double fsum = 0.0;
  if (mp.want_fitness) fsum = fitness_pass(G, gsrc, ns, mp, mo.p, coop);
  if (blockIdx.x == 0 && threadIdx.x == 0) write_result(out, mo, ns, fsum, mp.want_fitness != 0, G.n_tgt);
}

// ---------------------------------------------------------------------------------------------
// persistent batch matcher: one warp per match, work pulled from an atomic counter
// ---------------------------------------------------------------------------------------------
#ifndef NDT_WARP_KERNEL_MIN_CTAS
#define NDT_WARP_KERNEL_MIN_CTAS 2
#endif
// ndt_align_batch picks the warps per match from the batch size relative to the resident warps R (2,368 on a B200).
// Measured on C4 (ms per call for 1 / 2 / 4 / 8 warps per match, profiles/ab_c4.py --team):
//   65,536 matches 7.84 / 9.61 / 12.6 / 18.2     8,192: 1.54 / 1.53 / 1.75 / 2.41     4,096: 1.26 / 1.05 / 1.03 / 1.30
//    2,048: 0.80 / 0.64 / 0.57 / 0.65              512: 0.65 / 0.44 / 0.34 / 0.34
// A team trades throughput (helpers idle while their leader runs the optimiser) for the length of the longest match.
// The eight 8,192-hypothesis shards of the 65,536-hypothesis job (profiles/shard_sweep.py): 1.55 - 1.79 ms with one warp per
// match, 1.40 - 1.59 ms with two; the four 16,384-hypothesis shards: 2.42 - 2.66 against 2.61 - 2.79.
#ifndef NDT_TEAM1_FROM_X
#define NDT_TEAM1_FROM_X 4        // n >= 4 R: one warp per match
#endif
#ifndef NDT_TEAM2_FROM_X
#define NDT_TEAM2_FROM_X 2        // n >= 2 R: two; below: four, and eight when not even every CTA gets a match
#endif
#ifndef NDT_WARP_OCC_SMEM_MAX
#define NDT_WARP_OCC_SMEM_MAX (64 * 1024)
#endif
#ifndef NDT_WARP_KERNEL_THREADS
#define NDT_WARP_KERNEL_THREADS 256
#endif
constexpr int WK_THREADS = NDT_WARP_KERNEL_THREADS, WK_WARPS = WK_THREADS / 32;
constexpr int WK_QUEUE_BYTES = WK_WARPS * QUEUE_BYTES_PER_WARP;
struct __align__(16) WarpState { MatchOut mo; };
#ifndef NDT_WARP_JOB_CHUNK
#define NDT_WARP_JOB_CHUNK 8
#endif
// Work distribution of the persistent batch kernels: a CTA claims NDT_WARP_JOB_CHUNK consecutive jobs at a time from the
// global counter and its warps take them one by one (shared-memory state word = chunk base << 32 | jobs handed out).
// Consecutive relocalisation hypotheses share their position (16 headings per lattice point), so the warps of a CTA
// probe the same neighbourhood of the map and share its centroid / record lines in L1; a single global counter would
// scatter neighbouring hypotheses over all SMs. Returns the job for this warp (>= n_jobs: nothing left).
template <int CHUNK = NDT_WARP_JOB_CHUNK>
__device__ __forceinline__ int next_job(unsigned long long *s_state, int32_t *job_counter, int lane) {
  int job = 0;
  if (lane == 0) {
    for (;;) {
      const unsigned long long st = atomicAdd(s_state, 1ull);
      const int cnt = (int)(st & 0xffffffffu), base = (int)(st >> 32);
      if (cnt < CHUNK) { job = base + cnt; break; }
      if (cnt == CHUNK) {                                    // first to find the chunk empty: fetch the next one
        const int nb = atomicAdd(job_counter, CHUNK);
        atomicExch(s_state, ((unsigned long long)(unsigned)nb << 32) | 1ull);
        job = nb;
        break;
      }
      while ((int)(atomicAdd(s_state, 0ull) & 0xffffffffu) > CHUNK) {}   // a neighbour is fetching
    }
  }
  return __shfl_sync(0xffffffffu, job, 0);
}

// One instantiation per staging combination, chosen at launch: each holds exactly one copy of the optimiser + objective.
template <bool SRC_SMEM, bool OCC_SMEM>
__global__ void __launch_bounds__(WK_THREADS, NDT_WARP_KERNEL_MIN_CTAS) k_align_warp(GridView G, MatchParams mp, const float4 *__restrict__ src,
                                                   int ns, const double *__restrict__ guesses,
                                                   ndt_result *__restrict__ out, int64_t n_jobs,
                                                   int32_t *__restrict__ job_counter, int occ_words) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // dynamic shared memory: hit queues | occupancy bitmap (OCC_SMEM) | source points (SRC_SMEM) | per-warp results
  uint32_t *s_occ = reinterpret_cast<uint32_t *>(smem_raw + WK_QUEUE_BYTES);
  const int occ_bytes = OCC_SMEM ? ((occ_words * 4 + 15) & ~15) : 0;
  if (OCC_SMEM) for (int i = threadIdx.x; i < occ_words; i += blockDim.x) s_occ[i] = __ldg(G.occ + i);
  float2 *s_src = reinterpret_cast<float2 *>(smem_raw + WK_QUEUE_BYTES + occ_bytes);
  if (SRC_SMEM) {
    for (int i = threadIdx.x; i < ns; i += blockDim.x) {
      const float4 v = __ldg(src + i);
      s_src[i] = make_float2(v.x, v.y);
    }
  }
  // the result of a match: one copy per warp in shared memory (written once per match, read by lane 0)
  WarpState *ws = reinterpret_cast<WarpState *>(smem_raw + WK_QUEUE_BYTES + occ_bytes + (SRC_SMEM ? ((ns * 8 + 15) & ~15) : 0)) + (threadIdx.x >> 5);
  __shared__ unsigned long long s_state;
  if (threadIdx.x == 0) s_state = NDT_WARP_JOB_CHUNK;          // "chunk exhausted": the first warp fetches one
  __syncthreads();
  const int lane = threadIdx.x & 31;
  WarpCoop coop{lane};
  using OccL = typename std::conditional<OCC_SMEM, SmemOcc, GlobalOcc>::type;
  using SrcL = typename std::conditional<SRC_SMEM, SmemSrc, GlobalSrc>::type;
  OccL occ_acc; SrcL src_acc;
  if constexpr (OCC_SMEM) occ_acc = SmemOcc{smem_addr(s_occ)}; else occ_acc = GlobalOcc{G.occ};
  if constexpr (SRC_SMEM) src_acc = SmemSrc{smem_addr(s_src)}; else src_acc = GlobalSrc{src};
  for (;;) {
    const int job = next_job(&s_state, job_counter, lane);
    if (job >= n_jobs) break;
    const double guess[3] = {guesses[3 * (size_t)job], guesses[3 * (size_t)job + 1], guesses[3 * (size_t)job + 2]};
    MatchOut &mo = ws->mo;
    OptState opt;
    auto obj = make_objective(G, mp, coop, occ_acc, GlobalNbr{G.nbr}, GlobalCen{G.cen}, GlobalSlot{G.slot}, GlobalRec{G.recs}, src_acc, ns, my_queue(smem_raw));
    match_device(obj, mp, guess, mo, opt);
    double fsum = 0.0;
    if (mp.want_fitness) fsum = fitness_pass_cold(G, src_acc, ns, mp, mo.p, coop);
    if (lane == 0) write_result(out + job, mo, ns, fsum, mp.want_fitness != 0, G.n_tgt);
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------
// The same persistent schedule with a TEAM of WPM warps per match (WPM = 2, 4, 8). A match is a serial chain of passes:
// when a call brings only a few matches per resident warp (a shard of a relocalisation on 8 GPUs, a small multi-start),
// the call lasts as long as its longest match, however many warps sit idle. In a team the first warp (the leader) runs
// the match -- optimiser, line search, job queue -- and posts the pose of every objective pass in a shared-memory mailbox;
// all WPM warps then take their slice of the points (rank, rank + 32 WPM, ...) through the three stages with their own
// rings, leave their warp totals in the mailbox, and the leader adds them in warp order. Two named barriers per pass
// (pose posted / totals written); between them the helpers are parked on the barrier and cost no issue slots -- the
// optimiser runs once per match, not once per warp. Fixed summation order: results are run-to-run deterministic (they
// differ from the one-warp kernel's in the last bits: another summation tree).
// ---------------------------------------------------------------------------------------------
enum { TEAM_CMD_PASS = 1, TEAM_CMD_FITNESS = 2, TEAM_CMD_EXIT = 3 };
template <int WPM>
struct __align__(16) TeamBox {
  PoseF pf;
  double cs, sn;
  int cmd, pad;
  double rows[WPM][NACC];
};

template <int WPM, class OccL, class NbrL, class CenL, class SlotL, class RecL, class SrcL>
struct TeamObjective {
  ProbeGeom geom;
  OccL occ; NbrL nbr; CenL cen; SlotL slot; RecL rec; SrcL src;
  int ns;
  double d1, d2;
  int sse;
  TeamBox<WPM> *box;
  int wit, lane, bar_id;
  HitQueue Q;
  __device__ __forceinline__ void barrier() const { asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(32 * WPM) : "memory"); }
  // what every warp of the team runs for a posted pass: its slice of the points, its warp total into its mailbox row
  __device__ __noinline__ void slice(const PoseF pf, const double cs, const double sn) {
    double acc[NACC];
#pragma unroll
    for (int k = 0; k < NACC; ++k) acc[k] = 0.0;
    int pairs = 0;
    accumulate_points(0, geom, occ, nbr, cen, slot, rec, src, wit * 32 + lane, 32 * WPM, ns, pf, sse != 0, cs, sn, d1, d2, Q, acc, pairs);
    warp_allreduce<NACC>(acc);
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < NACC; ++k) box->rows[wit][k] = acc[k];
    }
    barrier();
  }
  __device__ __noinline__ void fitness_slice(const GridView &G, const PoseF pf) {
    double sum[1] = {0.0};
    for (int i0 = wit * 32; i0 < ns; i0 += 32 * WPM) {         // warp-uniform trip count: the 1-NN is warp-collective
      const int i = i0 + lane;
      const bool valid = i < ns;
      const float2 xy = src(valid ? i : ns - 1);
      float xt, yt;
      xform(pf, sse != 0, xy.x, xy.y, xt, yt);
      const float d = nn_dist2_warp(G, xt, yt, valid);
      if (valid) sum[0] += (double)d;
    }
    warp_allreduce<1>(sum);
    if (lane == 0) box->rows[wit][0] = sum[0];
    barrier();
  }
  __device__ __forceinline__ void post(const PoseF pf, double cs, double sn, int cmd) const {
    if (lane == 0) { box->pf = pf; box->cs = cs; box->sn = sn; box->cmd = cmd; }
    barrier();
  }
  // leader only: the objective pass match_device asks for (always the full mode: score, gradient, Hessian)
  __device__ __forceinline__ void pass(const int, const double *p, const AngleCache &ac, double *out) {
    const PoseF pf = pose_to_float(p);
    post(pf, ac.cs, ac.sn, TEAM_CMD_PASS);
    slice(pf, ac.cs, ac.sn);
#pragma unroll
    for (int k = 0; k < NACC; ++k) {
      double t = box->rows[0][k];
#pragma unroll
      for (int i = 1; i < WPM; ++i) t += box->rows[i][k];
      out[k] = t;
    }
  }
  __device__ __forceinline__ double fitness(const GridView &G, const double *p) {
    const PoseF pf = pose_to_float(p);
    post(pf, 0.0, 0.0, TEAM_CMD_FITNESS);
    fitness_slice(G, pf);
    double t = box->rows[0][0];
#pragma unroll
    for (int i = 1; i < WPM; ++i) t += box->rows[i][0];
    return t;
  }
  // helpers: serve posted passes until the leader says the queue is empty
  __device__ __forceinline__ void serve(const GridView &G) {
    for (;;) {
      barrier();
      const int cmd = box->cmd;
      if (cmd == TEAM_CMD_EXIT) break;
      const PoseF pf = box->pf;
      if (cmd == TEAM_CMD_PASS) slice(pf, box->cs, box->sn);
      else fitness_slice(G, pf);
    }
  }
};

template <int WPM>
__global__ void __launch_bounds__(WK_THREADS, NDT_WARP_KERNEL_MIN_CTAS) k_align_team(GridView G, MatchParams mp, const float4 *__restrict__ src,
                                                   int ns, const double *__restrict__ guesses,
                                                   ndt_result *__restrict__ out, int64_t n_jobs,
                                                   int32_t *__restrict__ job_counter, int occ_words) {
  static_assert(WK_WARPS % WPM == 0, "teams tile the CTA");
  constexpr int TEAMS = WK_WARPS / WPM;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // dynamic shared memory: hit queues | occupancy bitmap | source points | per-warp results (one per team used)
  uint32_t *s_occ = reinterpret_cast<uint32_t *>(smem_raw + WK_QUEUE_BYTES);
  const int occ_bytes = (occ_words * 4 + 15) & ~15;
  for (int i = threadIdx.x; i < occ_words; i += blockDim.x) s_occ[i] = __ldg(G.occ + i);
  float2 *s_src = reinterpret_cast<float2 *>(smem_raw + WK_QUEUE_BYTES + occ_bytes);
  for (int i = threadIdx.x; i < ns; i += blockDim.x) {
    const float4 v = __ldg(src + i);
    s_src[i] = make_float2(v.x, v.y);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, team = warp / WPM, wit = warp % WPM;
  WarpState *ws = reinterpret_cast<WarpState *>(smem_raw + WK_QUEUE_BYTES + occ_bytes + ((ns * 8 + 15) & ~15)) + team * WPM;
  __shared__ unsigned long long s_state;
  __shared__ TeamBox<WPM> s_box[TEAMS];
  if (threadIdx.x == 0) s_state = TEAMS;                       // "chunk exhausted": the first leader fetches one (a chunk = one job per team)
  __syncthreads();
  TeamObjective<WPM, SmemOcc, GlobalNbr, GlobalCen, GlobalSlot, GlobalRec, SmemSrc> obj{
      probe_geom(G), SmemOcc{smem_addr(s_occ)}, GlobalNbr{G.nbr}, GlobalCen{G.cen}, GlobalSlot{G.slot}, GlobalRec{G.recs},
      SmemSrc{smem_addr(s_src)}, ns, mp.d1, mp.d2, (mp.quirks & NDT_QUIRK_TRANSFORM_SSE_ORDER) ? 1 : 0,
      &s_box[team], wit, lane, 1 + team, my_queue(smem_raw)};
  if (wit != 0) { obj.serve(G); return; }
  for (;;) {
    const int job = next_job<TEAMS>(&s_state, job_counter, lane);
    if (job >= n_jobs) break;
    const double guess[3] = {guesses[3 * (size_t)job], guesses[3 * (size_t)job + 1], guesses[3 * (size_t)job + 2]};
    MatchOut &mo = ws->mo;
    OptState opt;
    match_device(obj, mp, guess, mo, opt);
    double fsum = 0.0;
    if (mp.want_fitness) fsum = obj.fitness(G, mo.p);
    if (lane == 0) write_result(out + job, mo, ns, fsum, mp.want_fitness != 0, G.n_tgt);
    __syncwarp();
  }
  obj.post(PoseF{0.f, 0.f, 0.f, 0.f}, 0.0, 0.0, TEAM_CMD_EXIT);
}

// ---------------------------------------------------------------------------------------------
// objective only, many poses (relocalisation score sweep): persistent CTAs, one warp per pose, source scan and occupancy
// bitmap staged once per CTA -- the batch matcher's schedule without the optimiser
// ---------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(256, 2) k_eval_warp(GridView G, MatchParams mp, const float4 *__restrict__ src, int ns,
                                                      const double *__restrict__ poses, double *__restrict__ out14,
                                                      int64_t *__restrict__ pairs_out, int64_t n_jobs,
                                                      int32_t *__restrict__ job_counter, int occ_words) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint32_t *s_occ = reinterpret_cast<uint32_t *>(smem_raw + QUEUE_BYTES);
  for (int i = threadIdx.x; i < occ_words; i += blockDim.x) s_occ[i] = __ldg(G.occ + i);
  float2 *s_src = reinterpret_cast<float2 *>(smem_raw + QUEUE_BYTES + ((occ_words * 4 + 15) & ~15));
  for (int i = threadIdx.x; i < ns; i += blockDim.x) {
    const float4 v = __ldg(src + i);
    s_src[i] = make_float2(v.x, v.y);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const bool sse = (mp.quirks & NDT_QUIRK_TRANSFORM_SSE_ORDER) != 0;
  const HitQueue Q = my_queue(smem_raw);
  const ProbeGeom geom = probe_geom(G);
  __shared__ unsigned long long s_state;
  if (threadIdx.x == 0) s_state = NDT_WARP_JOB_CHUNK;
  __syncthreads();
  for (;;) {
    const int job = next_job(&s_state, job_counter, lane);
    if (job >= n_jobs) break;
    const double p[3] = {poses[3 * (size_t)job], poses[3 * (size_t)job + 1], poses[3 * (size_t)job + 2]};
    AngleCache ac;
    angle_terms(mp, p[2], ac);
    const PoseF pf = pose_to_float(p);
    double acc[NACC];
#pragma unroll
    for (int k = 0; k < NACC; ++k) acc[k] = 0.0;
    int pairs = 0;
    accumulate_points(MODE, geom, SmemOcc{smem_addr(s_occ)}, GlobalNbr{G.nbr}, GlobalCen{G.cen}, GlobalSlot{G.slot}, GlobalRec{G.recs}, SmemSrc{smem_addr(s_src)}, lane, 32, ns,
                      pf, sse, ac.cs, ac.sn, mp.d1, mp.d2, Q, acc, pairs);
    warp_allreduce<NACC>(acc);
    if (lane < NACC) {
      double v = 0.0;
#pragma unroll
      for (int k = 0; k < NACC; ++k) if (lane == k) v = acc[k];
      out14[(size_t)job * (NACC + 1) + lane] = v;
    }
    if (lane == 0) {
      out14[(size_t)job * (NACC + 1) + NACC] = (double)pairs;
      if (pairs_out) pairs_out[job] = pairs;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// arg-max over batch results: highest score among converged matches, lowest index on ties
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_best_of(const ndt_result *__restrict__ res, int64_t n,
                                                 int64_t *__restrict__ best_index, ndt_result *__restrict__ best) {
  __shared__ double s_score[32];
  __shared__ long long s_idx[32];
  double bs = -DBL_MAX;
  long long bi = -1;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const double s = res[i].score;
    if (res[i].converged && (s > bs)) { bs = s; bi = i; }
  }
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) {
    const double os = __shfl_xor_sync(0xffffffffu, bs, m);
    const long long oi = __shfl_xor_sync(0xffffffffu, bi, m);
    if (oi >= 0 && (bi < 0 || os > bs || (os == bs && oi < bi))) { bs = os; bi = oi; }
  }
  if ((threadIdx.x & 31) == 0) { s_score[threadIdx.x >> 5] = bs; s_idx[threadIdx.x >> 5] = bi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    bs = -DBL_MAX; bi = -1;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
      const double os = s_score[w]; const long long oi = s_idx[w];
      if (oi >= 0 && (bi < 0 || os > bs || (os == bs && oi < bi))) { bs = os; bi = oi; }
    }
    *best_index = bi;
    if (bi >= 0) *best = res[bi];
  }
}

// ---------------------------------------------------------------------------------------------
// pcl::ApproximateVoxelGrid<PointXYZ>::applyFilter: an order-dependent 512-entry hash history.
// A point only ever touches the history slot its voxel hashes to, so points with different hashes commute; only the
// order *within a slot* and the order of the emitted centroids matter. A warp takes 32 consecutive points: lanes with
// the same hash are serialised in lane order (__match_any_sync groups, one round per group member), lanes with
// different hashes update their slots in parallel, and the centroids flushed by the chunk are written in lane order
// (ballot + popc) -- exactly the sequential output, bit for bit.
// ---------------------------------------------------------------------------------------------
struct VfTable { int ix[512], iy[512], iz[512], n[512]; float sx[512], sy[512], sz[512]; };   // 14,336 B per warp

__device__ __forceinline__ int voxel_filter_warp(const float4 *__restrict__ in, const int n, const float leaf,
                                                 float4 *__restrict__ o, VfTable &T) {
  const int lane = threadIdx.x & 31;
  for (int k = lane; k < 512; k += 32) { T.n[k] = 0; T.sx[k] = T.sy[k] = T.sz[k] = 0.f; T.ix[k] = T.iy[k] = T.iz[k] = 0; }
  __syncwarp();
  const float inv = __fdiv_rn(1.0f, leaf);
  const unsigned lt = (1u << lane) - 1u;
  int op = 0;
  for (int c = 0; c < n; c += 32) {
    const int i = c + lane;
    const bool valid = i < n;
    float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
    int ix = 0, iy = 0, iz = 0, hash = -1 - lane;           // invalid lanes: unique negative keys (groups of one, never active)
    if (valid) {
      p = __ldg(in + i);
      ix = (int)floorf(__fmul_rn(p.x, inv)); iy = (int)floorf(__fmul_rn(p.y, inv)); iz = (int)floorf(__fmul_rn(p.z, inv));
      hash = (int)(((unsigned)ix * 7171u + (unsigned)iy * 3079u + (unsigned)iz * 4231u) & 511u);
    }
    const unsigned grp = __match_any_sync(0xffffffffu, hash);
    const int rank = __popc(grp & lt);
    const int rounds = __reduce_max_sync(0xffffffffu, valid ? __popc(grp) : 0);
    bool flushed = false;
    float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = 0; r < rounds; ++r) {
      if (valid && rank == r) {
        int cnt = T.n[hash];
        float sx = T.sx[hash], sy = T.sy[hash], sz = T.sz[hash];
        if (cnt && (ix != T.ix[hash] || iy != T.iy[hash] || iz != T.iz[hash])) {      // a different voxel claims the slot
          const float cf = (float)cnt;
          f = make_float4(__fdiv_rn(sx, cf), __fdiv_rn(sy, cf), __fdiv_rn(sz, cf), 0.f);
          flushed = true;
          cnt = 0; sx = sy = sz = 0.f;
        }
        T.ix[hash] = ix; T.iy[hash] = iy; T.iz[hash] = iz; T.n[hash] = cnt + 1;
        T.sx[hash] = __fadd_rn(sx, p.x); T.sy[hash] = __fadd_rn(sy, p.y); T.sz[hash] = __fadd_rn(sz, p.z);
      }
      __syncwarp();
    }
    const unsigned bal = __ballot_sync(0xffffffffu, flushed);
    if (flushed) o[op + __popc(bal & lt)] = f;
    op += __popc(bal);
  }
  // end-of-cloud flush in slot order
  for (int k0 = 0; k0 < 512; k0 += 32) {
    const int k = k0 + lane;
    const int cnt = T.n[k];
    const unsigned bal = __ballot_sync(0xffffffffu, cnt != 0);
    if (cnt) {
      const float cf = (float)cnt;
      o[op + __popc(bal & lt)] = make_float4(__fdiv_rn(T.sx[k], cf), __fdiv_rn(T.sy[k], cf), __fdiv_rn(T.sz[k], cf), 0.f);
    }
    op += __popc(bal);
  }
  return op;
}

// one cloud (ndt_approx_voxel_filter): one warp
__global__ void __launch_bounds__(32) k_voxel_filter(const float4 *__restrict__ in, int64_t n, float leaf, float4 *__restrict__ out,
                                                     int32_t *__restrict__ n_out) {
  __shared__ VfTable tab;
  const int op = voxel_filter_warp(in, (int)n, leaf, out, tab);
  if (threadIdx.x == 0) *n_out = op;
}

// a batch of scan pairs, one warp per pair: pair i reads in[src_off .. src_off + ns) and writes its centroids to out at
// the same offset (the filter never grows a cloud), then records the new count in PairDims::ns. leaf <= 0: plain copy.
constexpr int VF_WARPS = 2;
__global__ void __launch_bounds__(32 * VF_WARPS) k_voxel_filter_pairs(const float4 *__restrict__ in, float4 *__restrict__ out,
                                                                      PairDims *__restrict__ dims, int n_pairs, float leaf) {
  __shared__ VfTable tabs[VF_WARPS];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int pair = blockIdx.x * VF_WARPS + w;
  if (pair >= n_pairs) return;
  const int64_t off = dims[pair].src_off;
  const int n = dims[pair].ns;
  if (!(leaf > 0.f)) {
    for (int i = lane; i < n; i += 32) out[off + i] = __ldg(in + off + i);
    return;
  }
  const int op = voxel_filter_warp(in + off, n, leaf, out + off, tabs[w]);
  if (lane == 0) dims[pair].ns = op;
}

// ---------------------------------------------------------------------------------------------
// persistent scan-pair matcher (loop-closure verification): one warp per pair pulled from an atomic work
// counter. Every pair has its own grid inside the shared padded tables (PairDims) and its own source cloud.
// ---------------------------------------------------------------------------------------------
#ifndef NDT_PAIRS_BLOCK_BELOW
#define NDT_PAIRS_BLOCK_BELOW 7000          // pairs per call below which one CTA (not one warp) matches a pair (measured: 1024 pairs 1.70 -> 1.02 ms, 2048 2.09 -> 1.59, 8192 equal)
#endif
#ifndef NDT_PAIRS_KERNEL_MIN_CTAS
#define NDT_PAIRS_KERNEL_MIN_CTAS 2
#endif
__global__ void __launch_bounds__(256, NDT_PAIRS_KERNEL_MIN_CTAS) k_align_pairs(GridView G0, MatchParams mp,
                                                    const PairDims *__restrict__ dims, const float4 *__restrict__ src_all,
                                                    const double *__restrict__ guesses, ndt_result *__restrict__ out,
                                                    int64_t n_jobs, int32_t *__restrict__ job_counter) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  WarpCoop coop{lane};
  WarpState *ws = reinterpret_cast<WarpState *>(smem_raw + QUEUE_BYTES) + (threadIdx.x >> 5);
  for (;;) {
    int job = 0;
    if (lane == 0) job = atomicAdd(job_counter, 1);
    job = __shfl_sync(0xffffffffu, job, 0);
    if (job >= n_jobs) break;
    const PairDims d = dims[job];
    GridView G = G0;
    G.min_bx = d.min_bx; G.min_by = d.min_by; G.div_x = d.div_x; G.div_y = d.div_y;
    G.slot_w = d.W; G.table_base = d.base;
    G.tgt = G0.tgt + d.tgt_off; G.n_tgt = d.nt; G.nn_f = 0;
    const double guess[3] = {guesses[3 * (size_t)job], guesses[3 * (size_t)job + 1], guesses[3 * (size_t)job + 2]};
    const GlobalSrc gsrc{src_all + d.src_off};
    MatchOut &mo = ws->mo;
    OptState opt;
    auto obj = make_objective(G, mp, coop, GlobalOcc{G.occ}, NoNbr{}, GlobalCen{G.cen}, GlobalSlot{G.slot}, GlobalRec{G.recs}, gsrc, d.ns, my_queue(smem_raw));   // cold per-pair tables: a mask table would be one more stream
    match_device(obj, mp, guess, mo, opt);
    double fsum = 0.0;
    if (mp.want_fitness) fsum = fitness_pass(G, gsrc, d.ns, mp, mo.p, coop);      // always taken for pairs: inlined (out of line: 4.4 -> 5.0 ms)
    if (lane == 0) write_result(out + job, mo, d.ns, fsum, mp.want_fitness != 0, G.n_tgt);
  }
}

// The same workload with one CTA per pair (persistent CTAs, block reduction per objective pass). A pair matched by one
// warp takes ~1.3 ms however many warps the GPU has to spare; eight warps finish it in ~0.15 ms, so whenever pairs are not
// plentiful compared with the warp slots of the GPU (sharded batches!) this is the lower-latency schedule.
// The CTA first stages what stages A and B read -- the pair's source cloud (as float2), its slice of the occupancy bitmap
// and of the centroid table -- in shared memory with coalesced loads; a pair too large for the staging area is read in
// place. The accessors take generic pointers, so both cases run the same code.
constexpr int PB_SRC_CAP = 1280, PB_CEN_CAP = 2560, PB_OCC_CAP = PB_CEN_CAP / 32 + 2, PB_REC_CAP = 320;
constexpr int PB_STAGE_BYTES = PB_SRC_CAP * 8 + PB_CEN_CAP * 8 + ((PB_OCC_CAP * 4 + 15) & ~15) + PB_CEN_CAP * 4 + PB_REC_CAP * 64;
struct AnyOcc {
  const uint32_t *p; int w0;
  __device__ __forceinline__ uint32_t operator()(int w) const { return p[w - w0]; }
};
struct AnyCen {
  const float2 *p; int i0;
  __device__ __forceinline__ float2 operator()(int i) const { return p[i - i0]; }
};

struct AnySlot {
  const int32_t *p; int i0;
  __device__ __forceinline__ int operator()(int i) const { return p[i - i0]; }
};
struct AnyRec {
  const CellRec *p;
  __device__ __forceinline__ void body(int s, double2 &m, double2 &r0, double2 &r1) const {
    const double2 *q = reinterpret_cast<const double2 *>(p + s);
    m = q[1]; r0 = q[2]; r1 = q[3];
  }
};
struct AnySrc {
  const float *p; int stride;      // floats between consecutive points (2: staged float2, 4: the caller's float4)
  __device__ __forceinline__ float2 operator()(int i) const { return *reinterpret_cast<const float2 *>(p + (size_t)i * stride); }
};

__global__ void __launch_bounds__(256, NDT_PAIRS_KERNEL_MIN_CTAS) k_align_pairs_block(GridView G0, MatchParams mp,
                                                    const PairDims *__restrict__ dims, const float4 *__restrict__ src_all,
                                                    const double *__restrict__ guesses, ndt_result *__restrict__ out,
                                                    int64_t n_jobs, int32_t *__restrict__ job_counter) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double scratch[9 * NACC];
  __shared__ int s_job;
  float2 *s_src = reinterpret_cast<float2 *>(smem_raw + QUEUE_BYTES);
  float2 *s_cen = s_src + PB_SRC_CAP;
  uint32_t *s_occ = reinterpret_cast<uint32_t *>(s_cen + PB_CEN_CAP);
  int32_t *s_slot = reinterpret_cast<int32_t *>(reinterpret_cast<unsigned char *>(s_occ) + ((PB_OCC_CAP * 4 + 15) & ~15));
  CellRec *s_recs = reinterpret_cast<CellRec *>(s_slot + PB_CEN_CAP);
  __shared__ int s_wcnt[8];
  BlockCoop coop{scratch};
  for (;;) {
    if (threadIdx.x == 0) s_job = atomicAdd(job_counter, 1);
    __syncthreads();
    const int job = s_job;
    __syncthreads();
    if (job >= n_jobs) break;
    const PairDims d = dims[job];
    GridView G = G0;
    G.min_bx = d.min_bx; G.min_by = d.min_by; G.div_x = d.div_x; G.div_y = d.div_y;
    G.slot_w = d.W; G.table_base = d.base;
    G.tgt = G0.tgt + d.tgt_off; G.n_tgt = d.nt; G.nn_f = 0;
    const double guess[3] = {guesses[3 * (size_t)job], guesses[3 * (size_t)job + 1], guesses[3 * (size_t)job + 2]};
    const int ncell = d.W * d.H, w0 = d.base >> 5, nw = ((d.base + ncell + 31) >> 5) - w0 + 1;
    AnySrc asrc{reinterpret_cast<const float *>(src_all + d.src_off), 4};
    AnyCen acen{G.cen, 0};
    AnyOcc aocc{G.occ, 0};
    if (d.ns <= PB_SRC_CAP) {
      for (int i = threadIdx.x; i < d.ns; i += blockDim.x) { const float4 v = __ldg(src_all + d.src_off + i); s_src[i] = make_float2(v.x, v.y); }
      asrc = AnySrc{reinterpret_cast<const float *>(s_src), 2};
    }
    if (ncell <= PB_CEN_CAP) {
      for (int i = threadIdx.x; i < ncell; i += blockDim.x) s_cen[i] = __ldg(G.cen + d.base + i);
      for (int i = threadIdx.x; i < nw; i += blockDim.x) s_occ[i] = __ldg(G.occ + w0 + i);
      acen = AnyCen{s_cen, d.base};
      aocc = AnyOcc{s_occ, w0};
    }
    __syncthreads();
    // the records of the pair's tree cells, compacted into shared memory (they sit scattered in the global array: slots are
    // handed out in completion order across all pairs) with a local cell -> record table: after this the whole match
    // -- probes, hits, records -- runs out of shared memory
    AnySlot aslot{G.slot, 0};
    AnyRec arec{G.recs};
    if (ncell <= PB_CEN_CAP) {
      const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
      int n_tree = 0;
      for (int c0 = 0; c0 < ncell; c0 += blockDim.x) {
        const int c = c0 + threadIdx.x;
        const bool tree = c < ncell && (s_cen[c].x == s_cen[c].x);
        const unsigned bal = __ballot_sync(0xffffffffu, tree);
        if (lane == 0) s_wcnt[warp] = __popc(bal);
        __syncthreads();
        int off = n_tree, tot = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) { const int v = s_wcnt[w]; tot += v; if (w < warp) off += v; }
        const int j = off + __popc(bal & ((1u << lane) - 1u));
        if (c < ncell) s_slot[c] = tree ? j : -1;
        if (tree && j < PB_REC_CAP) {
          const int4 *gr = reinterpret_cast<const int4 *>(G.recs + __ldg(G.slot + d.base + c));
          int4 *sr = reinterpret_cast<int4 *>(s_recs + j);
#pragma unroll
          for (int q = 0; q < 4; ++q) sr[q] = __ldg(gr + q);
        }
        n_tree += tot;
        __syncthreads();
      }
      if (n_tree <= PB_REC_CAP) { aslot = AnySlot{s_slot, d.base}; arec = AnyRec{s_recs}; }      // block-uniform
    }
    __syncthreads();
    MatchOut mo;
    OptState opt;
    auto obj = make_objective(G, mp, coop, aocc, NoNbr{}, acen, aslot, arec, asrc, d.ns, my_queue(smem_raw));
    match_device(obj, mp, guess, mo, opt);
    double fsum = 0.0;
#if NDT_PHASE_CLOCK
    const long long tf0 = clock64();
#endif
    if (mp.want_fitness) fsum = fitness_pass(G, asrc, d.ns, mp, mo.p, coop);
#if NDT_PHASE_CLOCK
    if (job < 3) obj.report("pairs_block", clock64() - tf0);
#endif
    if (threadIdx.x == 0) write_result(out + job, mo, d.ns, fsum, mp.want_fitness != 0, G.n_tgt);
  }
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
int launch_eval(Handle *h, const double *d_poses, int64_t n, int want_hessian, double *d_out14, int64_t *d_pairs) {
  if (n <= 0) return NDT_OK;
  const int ns = (int)h->ns;
  cudaStream_t st = h->stream;
  int slices = 1;
  if (n < 4 * h->sm_count) {
    slices = (int)std::min<int64_t>((ns + 1023) / 1024, std::max<int64_t>(1, (4 * h->sm_count) / n));
    if (slices < 1) slices = 1;
  }
  const int64_t blocks = n * slices;
  if (blocks > 0x7fffffffLL) return set_err(h, NDT_ERR_CAPACITY, "ndt_eval_batch: too many poses");
  NDT_CUDA(h, h->scratch.reserve((size_t)blocks * NACC * sizeof(double)));
  NDT_CUDA(h, h->scratch2.reserve((size_t)blocks * sizeof(int)));
  const GridView G = grid_view(h);
  const MatchParams mp = match_params(h, false);
  const float4 *src = h->src.as<float4>();
  {
    // many poses: the batch matcher's schedule (one warp per pose, staged source / occupancy)
    const int64_t npad = h->gd.n_cells > 0 ? (int64_t)(h->gd.div_x + 4) * (h->gd.div_y + 4) : 0;
    const int occ_words = (int)((npad + 31) / 32 + 1);
    const size_t smem = QUEUE_BYTES + (((size_t)occ_words * 4 + 15) & ~size_t(15)) + (size_t)ns * sizeof(float2);
    if (n >= 16 * h->sm_count && h->gd.n_cells > 0 && smem <= 100 * 1024) {
      if (int rc = ensure_nbr(h)) return rc;
      const GridView G = grid_view(h);                 // with the neighbour masks
      int32_t *ctr = h->gb.counters.as<int32_t>();
      if (h->timing) cudaEventRecord(h->ev0, st);
      NDT_CUDA(h, cudaMemsetAsync(ctr + CTR_JOB, 0, sizeof(int32_t), st));
      const int64_t grid = std::min<int64_t>((int64_t)h->sm_count * 2, (n + 7) / 8);
      if (want_hessian) {
        NDT_CUDA(h, cudaFuncSetAttribute(k_eval_warp<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_eval_warp<0><<<(unsigned)grid, 256, smem, st>>>(G, mp, src, ns, d_poses, d_out14, d_pairs, n, ctr + CTR_JOB, occ_words);
      } else {
        NDT_CUDA(h, cudaFuncSetAttribute(k_eval_warp<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_eval_warp<1><<<(unsigned)grid, 256, smem, st>>>(G, mp, src, ns, d_poses, d_out14, d_pairs, n, ctr + CTR_JOB, occ_words);
      }
      ++h->launches;
      if (h->timing) cudaEventRecord(h->ev1, st);
      NDT_CUDA(h, cudaGetLastError());
      return NDT_OK;
    }
  }
  if (h->timing) cudaEventRecord(h->ev0, st);
  NDT_CUDA(h, cudaFuncSetAttribute(k_eval_partial<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, QUEUE_BYTES));
  NDT_CUDA(h, cudaFuncSetAttribute(k_eval_partial<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, QUEUE_BYTES));
  if (want_hessian)
    k_eval_partial<0><<<(unsigned)blocks, 256, QUEUE_BYTES, st>>>(G, mp, src, ns, d_poses, slices, h->scratch.as<double>(), h->scratch2.as<int>());
  else
    k_eval_partial<1><<<(unsigned)blocks, 256, QUEUE_BYTES, st>>>(G, mp, src, ns, d_poses, slices, h->scratch.as<double>(), h->scratch2.as<int>());
  k_eval_final<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(h->scratch.as<double>(), h->scratch2.as<int>(), slices, n, d_out14, d_pairs);
  h->launches += 2;
  if (h->timing) cudaEventRecord(h->ev1, st);
  NDT_CUDA(h, cudaGetLastError());
  return NDT_OK;
}

#ifndef NDT_GRID_MIN_NS
#define NDT_GRID_MIN_NS 16384     // above this one match takes the whole GPU (cooperative launch)
#endif
#ifndef NDT_CLUSTER_MIN_NS
#define NDT_CLUSTER_MIN_NS 600      // measured on C1 (855 points): 8-CTA cluster 0.090 ms vs one CTA 0.113 ms
#endif
int launch_align(Handle *h, const double *d_guesses, int64_t n, ndt_result *d_results, bool want_fitness) {
  if (n <= 0) return NDT_OK;
  const int ns = (int)h->ns;
  cudaStream_t st = h->stream;
  const GridView G = grid_view(h);
  const MatchParams mp = match_params(h, want_fitness && h->grid_has_points);   // a replica imported without points reports fitness = NaN
  const float4 *src = h->src.as<float4>();
  int32_t *ctr = h->gb.counters.as<int32_t>();
  if (h->timing) cudaEventRecord(h->ev0, st);
  if (n >= 64) {
    // batch: persistent CTAs, one warp per match
    if (int rc = ensure_nbr(h)) return rc;
    const GridView G = grid_view(h);                   // with the neighbour masks
    NDT_CUDA(h, cudaMemsetAsync(ctr + CTR_JOB, 0, sizeof(int32_t), st));
    const bool src_smem = (size_t)ns * sizeof(float2) <= 64 * 1024;
    const int64_t npad = h->gd.n_cells > 0 ? (int64_t)(h->gd.div_x + 4) * (h->gd.div_y + 4) : 0;
    int occ_words = (int)((npad + 31) / 32 + 1);
    const bool occ_smem = src_smem && (size_t)occ_words * 4 <= NDT_WARP_OCC_SMEM_MAX;          // large grids: bitmap stays in global memory / L1
    const size_t smem = WK_QUEUE_BYTES + (occ_smem ? (((size_t)occ_words * 4 + 15) & ~size_t(15)) : 0) +
                        (src_smem ? (((size_t)ns * sizeof(float2) + 15) & ~size_t(15)) : 0) + WK_WARPS * sizeof(WarpState);
    const int ctas_per_sm = NDT_WARP_KERNEL_MIN_CTAS;
    int64_t grid = (int64_t)h->sm_count * ctas_per_sm;
    grid = std::min<int64_t>(grid, (n + WK_WARPS - 1) / WK_WARPS);
    auto go = [&](auto kern) -> cudaError_t {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
      kern<<<(unsigned)grid, WK_THREADS, smem, st>>>(G, mp, src, ns, d_guesses, d_results, n, ctr + CTR_JOB, occ_words);
      return cudaSuccess;
    };
    // warps per match: one when every resident warp has a queue of matches to hide the tail behind; teams otherwise
    // (measured on C4, profiles/ab_c4.py --team: see DESIGN.md 4.3). Teams need the staged source + bitmap.
    int wpm = h->prm.align_team;
    if (wpm != 1 && wpm != 2 && wpm != 4 && wpm != 8) {
      const int64_t R = (int64_t)h->sm_count * ctas_per_sm * WK_WARPS;     // resident warps
      wpm = n >= NDT_TEAM1_FROM_X * R ? 1 : n >= NDT_TEAM2_FROM_X * R ? 2 : n >= R / 8 ? 4 : 8;
    }
    if (!occ_smem) wpm = 1;
    if (wpm > 1) {
      grid = std::min<int64_t>((int64_t)h->sm_count * ctas_per_sm, (n * wpm + WK_WARPS - 1) / WK_WARPS);
      if (wpm == 2) NDT_CUDA(h, go(k_align_team<2>));
      else if (wpm == 4) NDT_CUDA(h, go(k_align_team<4>));
      else NDT_CUDA(h, go(k_align_team<8>));
    }
    else if (occ_smem) NDT_CUDA(h, go(k_align_warp<true, true>));
    else if (src_smem) NDT_CUDA(h, go(k_align_warp<true, false>));
    else NDT_CUDA(h, go(k_align_warp<false, false>));
  } else if (ns > NDT_GRID_MIN_NS && h->coop_launch) {
    // very large source cloud: one match at a time on every SM (cooperative launch, grid-wide reduction)
    int per_sm = 0;
    NDT_CUDA(h, cudaFuncSetAttribute(k_align_grid, cudaFuncAttributeMaxDynamicSharedMemorySize, QUEUE_BYTES));
    NDT_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_align_grid, 256, QUEUE_BYTES));
    int grid = h->sm_count * std::max(1, std::min(per_sm, 2));
    grid = (int)std::min<int64_t>(grid, (ns + 255) / 256);
    NDT_CUDA(h, h->scratch.reserve((size_t)2 * grid * 16 * sizeof(double)));
    double *gpart = h->scratch.as<double>();
    for (int64_t i = 0; i < n; ++i) {
      const double *gi = d_guesses + 3 * i;
      ndt_result *ri = d_results + i;
      int ns_arg = ns;
      void *args[] = {(void *)&G, (void *)&mp, (void *)&src, (void *)&ns_arg, (void *)&gi, (void *)&ri, (void *)&gpart};
      NDT_CUDA(h, cudaLaunchCooperativeKernel((const void *)k_align_grid, dim3(grid), dim3(256), args, QUEUE_BYTES, st));
      if (i > 0) ++h->launches;
    }
  } else if (ns > NDT_CLUSTER_MIN_NS) {
    // large source cloud: spread one match over a thread-block cluster (DSMEM reduction)
    int csize = NDT_CLUSTER_SIZE;
    if (csize > 8) {
      cudaError_t e = cudaFuncSetAttribute(k_align_cluster, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
      if (e != cudaSuccess) { csize = 8; (void)cudaGetLastError(); }
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(n * csize));
    cfg.blockDim = dim3(256);
    NDT_CUDA(h, cudaFuncSetAttribute(k_align_cluster, cudaFuncAttributeMaxDynamicSharedMemorySize, QUEUE_BYTES));
    cfg.dynamicSmemBytes = QUEUE_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = csize; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    NDT_CUDA(h, cudaLaunchKernelEx(&cfg, k_align_cluster, G, mp, src, ns, d_guesses, d_results));
  } else {
    const int n_slots = h->h_counters[CTR_SLOTS];
    const int n_cells = h->gd.n_cells > 0 ? (h->gd.div_x + 4) * (h->gd.div_y + 4) : 0;   // padded table
    const size_t tile = (size_t)n_slots * sizeof(CellRec) + (size_t)n_cells * (sizeof(float2) + sizeof(int32_t)) + ((size_t)(n_cells + 31) / 32 + 1) * 4;
    const size_t budget = (size_t)h->max_smem_optin > 4096 + QUEUE_BYTES ? (size_t)h->max_smem_optin - 4096 - QUEUE_BYTES : 0;
    if (tile > 0 && tile <= budget) {
      NDT_CUDA(h, cudaFuncSetAttribute(k_align_block<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(tile + QUEUE_BYTES)));
      k_align_block<true><<<(unsigned)n, 256, tile + QUEUE_BYTES, st>>>(G, mp, src, ns, d_guesses, d_results, n_slots, n_cells);
    } else {
      NDT_CUDA(h, cudaFuncSetAttribute(k_align_block<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, QUEUE_BYTES));
      k_align_block<false><<<(unsigned)n, 256, QUEUE_BYTES, st>>>(G, mp, src, ns, d_guesses, d_results, n_slots, n_cells);
    }
  }
  ++h->launches;
  if (h->timing) cudaEventRecord(h->ev1, st);
  NDT_CUDA(h, cudaGetLastError());
  return NDT_OK;
}

int launch_pairs_filter(Handle *h, const float4 *d_in, float4 *d_out, int64_t n_pairs, float leaf) {
  if (n_pairs <= 0) return NDT_OK;
  k_voxel_filter_pairs<<<(unsigned)((n_pairs + VF_WARPS - 1) / VF_WARPS), 32 * VF_WARPS, 0, h->stream>>>(d_in, d_out, h->gb.dims.as<PairDims>(),
                                                                                                       (int)n_pairs, leaf);
  ++h->launches;
  NDT_CUDA(h, cudaGetLastError());
  return NDT_OK;
}

int launch_align_pairs(Handle *h, const float4 *d_src, const double *d_guesses, int64_t n_pairs, ndt_result *d_results,
                       bool want_fitness) {
  if (n_pairs <= 0) return NDT_OK;
  cudaStream_t st = h->stream;
  GridView G = grid_view(h);        // shared tables; the per-pair geometry comes from PairDims inside the kernel
  G.nn_f = 0;
  const MatchParams mp = match_params(h, want_fitness);
  int32_t *ctr = h->gb.counters.as<int32_t>();
  NDT_CUDA(h, cudaMemsetAsync(ctr + CTR_JOB, 0, sizeof(int32_t), st));
  int64_t grid = (int64_t)h->sm_count * NDT_PAIRS_KERNEL_MIN_CTAS;
  grid = std::min<int64_t>(grid, (n_pairs + 7) / 8);
  const bool cta_per_pair = h->prm.pairs_schedule == NDT_PAIRS_CTA || (h->prm.pairs_schedule == NDT_PAIRS_AUTO && n_pairs < NDT_PAIRS_BLOCK_BELOW);
  if (cta_per_pair) {
    grid = std::min<int64_t>((int64_t)h->sm_count * NDT_PAIRS_KERNEL_MIN_CTAS, n_pairs);
    NDT_CUDA(h, cudaFuncSetAttribute(k_align_pairs_block, cudaFuncAttributeMaxDynamicSharedMemorySize, QUEUE_BYTES + PB_STAGE_BYTES));
    k_align_pairs_block<<<(unsigned)grid, 256, QUEUE_BYTES + PB_STAGE_BYTES, st>>>(G, mp, h->gb.dims.as<PairDims>(), d_src, d_guesses, d_results,
                                                                 n_pairs, ctr + CTR_JOB);
  } else {
    NDT_CUDA(h, cudaFuncSetAttribute(k_align_pairs, cudaFuncAttributeMaxDynamicSharedMemorySize, QUEUE_BYTES + 8 * (int)sizeof(WarpState)));
    k_align_pairs<<<(unsigned)grid, 256, QUEUE_BYTES + 8 * sizeof(WarpState), st>>>(G, mp, h->gb.dims.as<PairDims>(), d_src, d_guesses, d_results,
                                                           n_pairs, ctr + CTR_JOB);
  }
  ++h->launches;
  NDT_CUDA(h, cudaGetLastError());
  return NDT_OK;
}

int launch_best_of(Handle *h, const ndt_result *d_results, int64_t n, int64_t *d_best_index, ndt_result *d_best) {
  k_best_of<<<1, 1024, 0, h->stream>>>(d_results, n, d_best_index, d_best);
  ++h->launches;
  NDT_CUDA(h, cudaGetLastError());
  return NDT_OK;
}

int launch_voxel_filter(Handle *h, const float4 *d_in, int64_t n, float leaf, float4 *d_out, int32_t *d_nout) {
  k_voxel_filter<<<1, 32, 0, h->stream>>>(d_in, n, leaf, d_out, d_nout);
  ++h->launches;
  NDT_CUDA(h, cudaGetLastError());
  return NDT_OK;
}

}  // namespace ndt
