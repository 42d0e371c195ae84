// ndt_device.cuh -- device-side data layout and math of the 2-D NDT hot path (sm_100a).
//
// What the reference computes here lives in PCL (reached from src/PoseEstimator.cpp:19, 28, 43, 56):
// VoxelGridCovariance (grid), NormalDistributionsTransform::computeDerivatives / computeHessian /
// computeTransformation / computeStepLengthMT, Registration::getFitnessScore. SURVEY.md App. A/B is
// the specification; this file is a from-scratch device formulation of it:
//   * the grid is a dense int32 cell->slot table plus compact 64-byte cell records (one 64-B aligned
//     record = float centroid for the radius test + fp64 mean + 2x2 inverse covariance),
//   * neighbour search is the exact equivalent of PCL's centroid radius search: the 3x3 cell block
//     around the transformed point, float32 distance test d^2 < r^2,
//   * the float32 point transform uses explicit round-to-nearest mul/add (no FMA contraction),
//   * accumulation is fp64 with fixed-order reductions so every cooperating thread ends up with
//     bit-identical totals and takes identical branches in the optimiser (no broadcast needed).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <float.h>

#include "ndt_b200.h"

namespace ndt {

// ------------------------------------------------------------------------------------------------
// layout
// ------------------------------------------------------------------------------------------------
struct __align__(16) CellRec {   // 64 bytes
  float cx, cy;                  // float32 centroid, accumulated in input order (radius test)
  int32_t nr_points;             // point count, or -1 when the eigenvalue / inf check failed (icov = 0)
  int32_t cell;                  // dense cell index (ijk0 + ijk1 * div_x)
  double mx, my;                 // fp64 mean
  double c00, c01, c10, c11;     // inverse covariance (xx, xy, yx, yy)
};
static_assert(sizeof(CellRec) == 64, "CellRec must be 64 bytes");

struct GridView {
  const int32_t *__restrict__ slot;     // [(div_x + 4) * (div_y + 4)] padded by 2 cells on every side (defined only where cen is not NaN):
                                        // cell (i, j) lives at (j + 2) * slot_w + i + 2; value = record index or -1
  int32_t slot_w;                       // div_x + 4
  int32_t table_base;                   // offset of this grid inside the shared tables (0 for a single grid;
                                        // batched scan-pair matching packs one padded table per pair back to back)
  const uint32_t *__restrict__ occ;     // 1 bit per padded cell: some tree cell lies in the 3x3 block around it
                                        // (dilated occupancy: a clear bit means the point cannot hit anything)
  const float2 *__restrict__ cen;       // same padded indexing: float32 centroid of tree cells, NaN elsewhere
                                        // (the probe is one load + a float compare: NaN < r2 is false)
  const uint16_t *__restrict__ nbr;     // same padded indexing: which cells of the 3x3 block around a cell are tree cells,
                                        // bit (dj + 1) * 4 + (di + 1) for the neighbour at (di, dj); nbr != 0 <=> occ bit.
                                        // Derived from cen on demand, for the batch kernels only (ensure_nbr)
  const CellRec *__restrict__ recs;     // compact records of cells with n >= min_points
  int32_t min_bx, min_by, div_x, div_y;
  float inv_leaf;                       // 1.0f / leaf
  float r2;                             // (float)((double)leaf * leaf)
  float leaf;
  // 1-NN buckets (fitness): every occupied cell
  const int32_t *__restrict__ leaf_id;  // same padded indexing -> leaf index + 1, 0 for an empty cell
  const int2 *__restrict__ leaf_range;  // per leaf: (start, n) into tgt_sorted
  const float2 *__restrict__ tgt_sorted;// target (x, y) in bucket order (cell by cell, input order inside a cell)
  const float4 *__restrict__ tgt;       // target points, input order
  int64_t n_tgt;
  // optional finer lattice used only for nearest-neighbour queries when the NDT buckets are dense
  // (nn_f = 0: search the NDT buckets above; otherwise cells of leaf / nn_f with a dense (start, n) table)
  int32_t nn_f;
  int32_t nn_min_bx, nn_min_by, nn_div_x, nn_div_y;
  float nn_inv_leaf, nn_leaf;
  const int2 *__restrict__ nn_range;    // [nn_div_x * nn_div_y]
  const float2 *__restrict__ nn_pts;    // target (x, y) in fine-bucket order
};

// geometry of one grid inside the shared padded tables (one entry for ndt_set_target, one per pair for
// ndt_match_pairs); for the batched matcher it also names the pair's point ranges
struct PairDims {
  int32_t min_bx, min_by, div_x, div_y;
  int32_t W, H;            // padded extents: div + 4
  int32_t base;            // first entry of this grid in the shared tables
  int32_t ns;              // pairs: number of (filtered) source points
  int64_t src_off, tgt_off;// pairs: first source / target point
  int64_t nt;              // pairs: number of target points
};

struct MatchParams {
  double d1, d2;          // gauss_d1_, gauss_d2_
  double step_size;       // step_max
  double step_min;        // transformation_epsilon_ / 2
  double trans_eps;
  int32_t max_iter;
  int32_t quirks;
  int32_t want_fitness;
  int32_t pad;
};

struct PoseF { float c, s, tx, ty; };

// ------------------------------------------------------------------------------------------------
// float32 pieces that must be bit-exact with the oracle
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int cell_coord(float v, float inv, int min_b) {
  // VoxelGridCovariance pass 1: float multiply, float floor, float subtract, truncate
  return (int)__fsub_rn(floorf(__fmul_rn(v, inv)), (float)min_b);
}

// fp64 sincos carries a long slow path; one out-of-line copy serves every call site of a kernel
static __device__ __noinline__ void sincos_once(double a, double *s, double *c) { sincos(a, s, c); }

__device__ __forceinline__ PoseF pose_to_float(const double p[3]) {
  PoseF f;
  const float yaw = (float)p[2];
  double sd, cd;
  sincos_once((double)yaw, &sd, &cd);
  f.c = (float)cd; f.s = (float)sd;
  f.tx = (float)p[0]; f.ty = (float)p[1];
  return f;
}

// Two float32 additions in one instruction (sm_100 FADD2): (rx, ry) = (ax + bx, ay + by), each rounded to nearest like
// __fadd_rn. Only ADDITIONS are packed: ptxas contracts a packed multiply that feeds a packed add into FFMA2 even when both
// carry .rn (measured: profiles/microbench.cu, 29,204 of 65,531 transform chains differ from the scalar result), which
// would break the bit-exact transform; scalar __fmul_rn products feeding a packed add stay unfused (0 mismatches).
#ifndef NDT_PACKED_ADD
#define NDT_PACKED_ADD 1
#endif
__device__ __forceinline__ void add2_rn(float &rx, float &ry, float ax, float ay, float bx, float by) {
#if NDT_PACKED_ADD
  asm("{ .reg .b64 ra, rb, rc; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; add.rn.f32x2 rc, ra, rb; mov.b64 {%0, %1}, rc; }"
      : "=f"(rx), "=f"(ry) : "f"(ax), "f"(ay), "f"(bx), "f"(by));
#else
  rx = __fadd_rn(ax, bx); ry = __fadd_rn(ay, by);
#endif
}

__device__ __forceinline__ void xform(const PoseF &f, bool sse_order, float x, float y, float &ox, float &oy) {
  const float ns = -f.s;
  const float cx = __fmul_rn(f.c, x), sx = __fmul_rn(f.s, x), nsy = __fmul_rn(ns, y), cy = __fmul_rn(f.c, y);
  float ux, uy;
  if (sse_order) {               // x' = c x + ((-s) y + tx)
    add2_rn(ux, uy, nsy, cy, f.tx, f.ty);
    add2_rn(ox, oy, cx, sx, ux, uy);
  } else {                       // x' = (c x + (-s) y) + tx
    add2_rn(ux, uy, cx, sx, nsy, cy);
    add2_rn(ox, oy, ux, uy, f.tx, f.ty);
  }
}

__device__ __forceinline__ float dist2f(float ax, float ay, float bx, float by) {
  float dx, dy;
  add2_rn(dx, dy, ax, ay, -bx, -by);          // a - b == a + (-b) exactly
  return __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
}

// ------------------------------------------------------------------------------------------------
// objective
// MODE 0: score + gradient + Hessian; 1: score + gradient; 2: Hessian only (computeHessian). The mode is a run-time,
// warp-uniform value: the three stages exist once in the instruction stream of the persistent matchers (they are
// instruction-cache bound when every warp of an SM sits in a different pass of a different match).
// acc layout: [0] score, [1..3] gradient, [4..12] Hessian row-major
//
// accumulate_points() below runs it as three compacted stages (transform -> probe -> fp64 hit path) linked by
// two per-warp rings in shared memory.
// ------------------------------------------------------------------------------------------------
constexpr int NACC = 13;
constexpr int NDT_MAX_CELL_INDEX = 1 << 22;    // |floor(x / leaf)| of every target point stays below this (host-checked)
#ifndef NDT_HITS_PER_LANE
#define NDT_HITS_PER_LANE 2
#endif
constexpr int HITS_PER_LANE = NDT_HITS_PER_LANE;   // hits a lane evaluates side by side in one drain
constexpr int QCAP = 32 * HITS_PER_LANE + 288;   // hit ring per warp: < 32 * HITS_PER_LANE queued before a probe, a probe adds <= 9 * 32
#ifndef NDT_A_STEPS
#define NDT_A_STEPS 3      // measured on C4: 1 -> 13.9 ms, 2 -> 13.4, 3 -> 13.1, 4 -> 13.3
#endif
constexpr int A_STEPS = NDT_A_STEPS;      // warp steps (32 points each) stage A transforms per iteration
constexpr int CQCAP = A_STEPS <= 3 ? 128 : 256;   // candidate ring per warp: < 32 queued before an iteration, which adds <= 32 * A_STEPS
constexpr int QUEUE_BYTES_PER_WARP = (QCAP + CQCAP) * 8;

// Ring entries are 8 bytes: the source point index plus a table position. The float32 transform of a point is
// eight flops, so stages B and C redo it from the (shared-memory) source point instead of carrying 16 more bytes
// per entry -- the rings of a CTA then take 24 KB and three CTAs fit on an SM.
// Shared memory is addressed through 32-bit shared-window addresses and explicit ld.shared / st.shared: the rings and the
// staged tables reach the hot loop through a struct and a non-inlined call, where the compiler no longer knows that a
// plain pointer is a shared-memory pointer and falls back to generic 64-bit loads and stores.
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ int2 lds_int2(uint32_t a) {
  int2 v;
  asm volatile("ld.shared.v2.s32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts_int2(uint32_t a, int2 v) {
  asm volatile("st.shared.v2.s32 [%0], {%1, %2};" ::"r"(a), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ float2 lds_float2(uint32_t a) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ int lds_s32(uint32_t a) { return (int)lds_u32(a); }
__device__ __forceinline__ double2 lds_double2(uint32_t a) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ float4 lds_float4(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}

struct HitQueue {
  uint32_t hit;   // shared address of int2[QCAP]:  (source point index, padded-table index of the hit cell)
  uint32_t cand;  // shared address of int2[CQCAP]: (source point index, padded-table index of the point's own cell)
};

// NH hits per lane, evaluated side by side: the code is straight-line (the reference's guard on e turns into
// selects), so the independent dependency chains of the hits interleave and hide each other's fp64 latency.
// valid[u] = the lane really holds hit u; a lane without a hit computes on a clamped copy and adds exact zeros.
template <int NH, class RecL>
__device__ __forceinline__ void hit_path(const int MODE, const RecL &rec_at, const float4 *e, const int *s, const bool *valid,
                                         const double cs, const double sn, const double d1, const double d2, double *acc) {
  double dx[NH], dy[NH], x[NH], y[NH], c00[NH], c01[NH], c10[NH], c11[NH], ex[NH], sc[NH];
#pragma unroll
  for (int u = 0; u < NH; ++u) {
    double2 m, r0, r1;
    rec_at.body(s[u], m, r0, r1);
    x[u] = (double)e[u].z; y[u] = (double)e[u].w;
    dx[u] = (double)e[u].x - m.x; dy[u] = (double)e[u].y - m.y;
    c00[u] = r0.x; c01[u] = r0.y; c10[u] = r1.x; c11[u] = r1.y;
  }
#pragma unroll
  for (int u = 0; u < NH; ++u) {
    const double Cdx = c00[u] * dx[u] + c01[u] * dy[u], Cdy = c10[u] * dx[u] + c11[u] * dy[u];
    const double q = dx[u] * Cdx + dy[u] * Cdy;
    ex[u] = exp(-d2 * q / 2.0);
  }
#pragma unroll
  for (int u = 0; u < NH; ++u) {
    const double t = d2 * ex[u];
    const bool ok = valid[u] && !(t > 1.0 || t < 0.0 || t != t);   // updateDerivatives returns 0 for a rejected hit: no score term either
    sc[u] = ok ? -d1 * ex[u] : 0.0;
    ex[u] = ok ? t * d1 : 0.0;
    dx[u] = ok ? dx[u] : 0.0; dy[u] = ok ? dy[u] : 0.0;      // keep a rejected term finite: its contributions are exact zeros
  }
#pragma unroll
  for (int u = 0; u < NH; ++u) {
    const double Jx = -sn * x[u] - cs * y[u], Jy = cs * x[u] - sn * y[u];
    // cov_dxd_pi = C * J_i for i = x, y, yaw ; a_i = d . (C J_i)
    const double CJ2x = c00[u] * Jx + c01[u] * Jy, CJ2y = c10[u] * Jx + c11[u] * Jy;
    const double a0 = dx[u] * c00[u] + dy[u] * c10[u];
    const double a1 = dx[u] * c01[u] + dy[u] * c11[u];
    const double a2 = dx[u] * CJ2x + dy[u] * CJ2y;
    if (MODE != 2) {
      acc[0] += sc[u];
      acc[1] += a0 * ex[u]; acc[2] += a1 * ex[u]; acc[3] += a2 * ex[u];
    }
    if (MODE != 1) {
      const double Hx = -cs * x[u] + sn * y[u], Hy = -sn * x[u] - cs * y[u];
      const double dCH = dx[u] * (c00[u] * Hx + c01[u] * Hy) + dy[u] * (c10[u] * Hx + c11[u] * Hy);
      const double k0 = -d2 * a0, k1 = -d2 * a1, k2 = -d2 * a2;
      // H(i,j) += e * (-d2 a_i a_j + J_j . (C J_i) [+ d.(C H_yawyaw) for i = j = yaw])
      acc[4]  += ex[u] * (k0 * a0 + c00[u]);
      acc[5]  += ex[u] * (k0 * a1 + c10[u]);
      acc[6]  += ex[u] * (k0 * a2 + (Jx * c00[u] + Jy * c10[u]));
      acc[7]  += ex[u] * (k1 * a0 + c01[u]);
      acc[8]  += ex[u] * (k1 * a1 + c11[u]);
      acc[9]  += ex[u] * (k1 * a2 + (Jx * c01[u] + Jy * c11[u]));
      acc[10] += ex[u] * (k2 * a0 + CJ2x);
      acc[11] += ex[u] * (k2 * a1 + CJ2y);
      acc[12] += ex[u] * (k2 * a2 + (Jx * CJ2x + Jy * CJ2y) + dCH);
    }
  }
}

// the scalars of the grid the probe loop needs, held in registers (not re-read through a struct pointer)
struct ProbeGeom {
  int W, base;
  int lo_x, lo_y;         // min_b - 1: the floor index of the first column / row whose 3x3 block touches the grid
  unsigned span_x, span_y;// div + 1: floor indices lo .. lo + span (inclusive) can hit something
  int cell0;              // padded-table index of floor index (0, 0): a point with floor indices (fi, fj) lives at cell0 + fj * W + fi
  float inv_leaf, r2;
};
__device__ __forceinline__ ProbeGeom probe_geom(const GridView &G) {
  return ProbeGeom{G.slot_w, G.table_base, G.min_bx - 1, G.min_by - 1, (unsigned)(G.div_x + 1), (unsigned)(G.div_y + 1),
                   G.table_base + (2 - G.min_by) * G.slot_w + 2 - G.min_bx, G.inv_leaf, G.r2};
}
// floor(v * inv) as an integer in one conversion (cvt.rmi). Equal to the reference's float expression
// (int)(floorf(v * inv) - (float)min_b) + min_b whenever every quantity stays below 2^23 in magnitude, which the host
// guarantees for the grid (NDT_MAX_CELL_INDEX) -- a point farther out lands outside the grid either way.
__device__ __forceinline__ int floor_index(float v, float inv) { return __float2int_rd(__fmul_rn(v, inv)); }

// Accumulate the objective over points i = first + k * stride (k = 0, 1, ...), i < hi, where `first`
// is lane-contiguous inside a warp (first = warp_first + lane): every lane of a warp iterates the same
// number of times. acc must be a register array of the caller. pairs: warp-uniform hit count.
//
// Three stages, each run by a full warp on compacted work (a warp-uniform state machine; every stage
// exists once in the instruction stream -- the matcher is instruction-cache sensitive):
//   A  transform  2 x 32 source points: float32 transform (bit-exact), own cell, dilated-occupancy bit. Points
//                 that can hit something (typically a third of a scan against a building-sized map) are
//                 pushed into the per-warp candidate ring (ballot + popc offsets).
//   B  probe      whenever 32 candidates are queued: nine centroid loads each, issued back to back (no
//                 bounds checks: padded table; no emptiness branch: NaN centroids fail the compare), the
//                 float32 radius test gives a 9-bit hit mask per lane, one warp prefix scan turns the
//                 per-lane hit counts into offsets in the hit ring.
//   C  drain      whenever 32 hits are queued: every lane pops one (point, cell) hit and runs the fp64
//                 hit path fully converged.
// Without the rings the probe would run for every point (all 32 lanes pay when one is a candidate) and the
// hit path once per neighbour position with a handful of active lanes.
// stage A for one point: padded-table index of its own cell if the point can hit anything, else -1
template <class OccL, class SrcL>
__device__ __forceinline__ int candidate_cell(const ProbeGeom &g, const OccL &occ_at, const SrcL &src, const int i,
                                              const bool valid, const PoseF &pf, const bool sse_order) {
  const float2 xy = src(i);
  float xt, yt;
  xform(pf, sse_order, xy.x, xy.y, xt, yt);
  const int fi = floor_index(xt, g.inv_leaf), fj = floor_index(yt, g.inv_leaf);
  // cell -1 .. div in both directions: the cells whose 3x3 block can touch the grid
  const bool in = valid && (unsigned)(fi - g.lo_x) <= g.span_x && (unsigned)(fj - g.lo_y) <= g.span_y;
  const int base = in ? (g.cell0 + fj * g.W + fi) : g.base;      // out of range: any valid entry of this grid (unused)
  const unsigned word = occ_at(base >> 5);
  const bool cand = in && ((word >> (base & 31)) & 1u);
  return cand ? base : -1;
}

template <class OccL, class NbrL, class CenL, class SlotL, class RecL, class SrcL>
__device__ __forceinline__ void accumulate_points(const int MODE, const ProbeGeom g, const OccL occ_at, const NbrL nbr_at, const CenL cen_at, const SlotL slot_at,
                                                  const RecL rec_at, const SrcL src, const int first,
                                                  const int stride, const int hi, const PoseF pf,
                                                  const bool sse_order, const double cs, const double sn,
                                                  const double d1, const double d2, const HitQueue Q, double *acc,
                                                  int &pairs) {
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  int qhead = 0, qn = 0;           // hit ring
  int chead = 0, cn = 0;           // candidate ring
  int i0 = first - lane;
  for (;;) {
    // ---- A: transform source points A_STEPS warp steps at a time until 32 candidates are queued ----
    while (cn < 32 && i0 < hi) {
      int b[A_STEPS];
#pragma unroll
      for (int u = 0; u < A_STEPS; ++u) {
        const int i = i0 + lane + u * stride;
        b[u] = candidate_cell(g, occ_at, src, min(i, hi - 1), i < hi, pf, sse_order);
      }
      int at = chead + cn;
#pragma unroll
      for (int u = 0; u < A_STEPS; ++u) {
        const unsigned bal = __ballot_sync(0xffffffffu, b[u] >= 0);
        if (b[u] >= 0) sts_int2(Q.cand + 8u * ((at + __popc(bal & lt)) & (CQCAP - 1)), make_int2(i0 + lane + u * stride, b[u]));
        at += __popc(bal);
      }
      cn = at - chead;
      i0 += A_STEPS * stride;
    }
    const bool done = (i0 >= hi);
    if (cn >= 32 || (done && cn > 0)) {
      // ---- B: probe up to 32 queued candidates ----
      const int n = min(cn, 32);
      __syncwarp();
      unsigned mask = 0u;
      int2 cd = make_int2(0, 0);
      if (lane < n) {
        cd = lds_int2(Q.cand + 8u * ((chead + lane) & (CQCAP - 1)));
        const float2 xy = src(cd.x);
        float xt, yt;
        xform(pf, sse_order, xy.x, xy.y, xt, yt);
        const int org = cd.y - (g.W + 1);
        if constexpr (NbrL::enabled) {
          // throughput-bound batch kernels: the neighbour mask names the tree cells of the 3x3 block (a wall crosses ~3 of
          // the 9); only those are loaded and tested, four at a time so the loads overlap (half the L1 wavefronts of
          // nine blind loads). bit b of the mask <-> cell offset (b >> 2) * W + (b & 3) - (W + 1)
          unsigned todo = nbr_at(cd.y);
          while (todo) {
            int bit[4];
            float2 c[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              bit[u] = __ffs(todo) - 1;                 // -1 once the mask is exhausted
              todo &= todo - 1u;
              c[u] = bit[u] >= 0 ? cen_at(org + (bit[u] >> 2) * g.W + (bit[u] & 3)) : make_float2(NAN, NAN);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) mask |= (dist2f(xt, yt, c[u].x, c[u].y) < g.r2) ? (1u << (bit[u] & 31)) : 0u;
          }
        } else {
          // latency-bound single matches: nine independent centroid loads issued back to back (no dependent mask load in
          // front of them; NaN centroids of non-tree cells fail the compare)
          float2 c[9];
#pragma unroll
          for (int k = 0; k < 9; ++k) c[k] = cen_at(org + (k / 3) * g.W + (k % 3));
#pragma unroll
          for (int k = 0; k < 9; ++k) mask |= (dist2f(xt, yt, c[k].x, c[k].y) < g.r2) ? (1u << ((k / 3) * 4 + (k % 3))) : 0u;
        }
      }
      chead = (chead + n) & (CQCAP - 1);
      cn -= n;
      if (__any_sync(0xffffffffu, mask != 0u)) {
        const int cnt = __popc(mask);
        int incl = cnt;
#pragma unroll
        for (int dlt = 1; dlt < 32; dlt <<= 1) {
          const int t = __shfl_up_sync(0xffffffffu, incl, dlt);
          if (lane >= dlt) incl += t;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        int pos = qhead + qn + incl - cnt;
        if (pos >= QCAP) pos -= QCAP;
        const int org = cd.y - (g.W + 1);
        while (mask) {
          const int k = __ffs(mask) - 1;
          mask &= mask - 1u;
          sts_int2(Q.hit + 8u * pos, make_int2(cd.x, org + (k >> 2) * g.W + (k & 3)));
          if (++pos == QCAP) pos = 0;
        }
        qn += total;
      }
    }
    // ---- C: every lane pops HITS_PER_LANE queued (point, cell) hits and runs the fp64 hit path converged ----
    while (qn >= 32 * HITS_PER_LANE || (done && cn == 0 && qn > 0)) {
      const int n = min(qn, 32 * HITS_PER_LANE);
      __syncwarp();
      float4 e[HITS_PER_LANE];
      int sl[HITS_PER_LANE];
      bool valid[HITS_PER_LANE];
#pragma unroll
      for (int u = 0; u < HITS_PER_LANE; ++u) {
        const int k = lane + 32 * u;
        valid[u] = k < n;
        int pos = qhead + (valid[u] ? k : 0);      // a lane without a hit recomputes the first one with weight zero
        if (pos >= QCAP) pos -= QCAP;
        const int2 h = lds_int2(Q.hit + 8u * pos);
        const float2 xy = src(h.x);
        e[u].z = xy.x; e[u].w = xy.y;
        xform(pf, sse_order, xy.x, xy.y, e[u].x, e[u].y);
        sl[u] = slot_at(h.y);
      }
      hit_path<HITS_PER_LANE>(MODE, rec_at, e, sl, valid, cs, sn, d1, d2, acc);
      __syncwarp();
      qhead += n;
      if (qhead >= QCAP) qhead -= QCAP;
      qn -= n;
      pairs += n;
    }
    if (done && cn == 0) break;       // the drain loop above has emptied the hit ring
  }
  __syncwarp();
}

// global-memory accessors (read-only path, L1/L2 cached)
struct GlobalOcc {
  const uint32_t *__restrict__ p;
  __device__ __forceinline__ uint32_t operator()(int w) const { return __ldg(p + w); }
};
struct SmemOcc {            // bitmap staged in shared memory
  uint32_t a;
  __device__ __forceinline__ uint32_t operator()(int w) const { return lds_u32(a + 4u * w); }
};
struct MixedOcc {           // staged when it fits, else read in place (warp-uniform choice)
  uint32_t a;
  const uint32_t *__restrict__ g;
  __device__ __forceinline__ uint32_t operator()(int w) const { return g ? __ldg(g + w) : lds_u32(a + 4u * w); }
};
struct GlobalNbr {
  static constexpr bool enabled = true;
  const uint16_t *__restrict__ p;
  __device__ __forceinline__ unsigned operator()(int i) const { return (unsigned)__ldg(p + i); }
};
struct NoNbr {                // probe all nine cells (single-match kernels)
  static constexpr bool enabled = false;
  __device__ __forceinline__ unsigned operator()(int) const { return 0u; }
};
struct GlobalCen {
  const float2 *__restrict__ p;
  __device__ __forceinline__ float2 operator()(int i) const { return __ldg(p + i); }
};
struct SmemCen {
  uint32_t a;
  __device__ __forceinline__ float2 operator()(int i) const { return lds_float2(a + 8u * i); }
};
struct GlobalSlot {
  const int32_t *__restrict__ p;
  __device__ __forceinline__ int operator()(int i) const { return __ldg(p + i); }
};
struct GlobalRec {
  const CellRec *__restrict__ p;
  __device__ __forceinline__ float4 head(int s) const {
    return __ldg(reinterpret_cast<const float4 *>(p + s));
  }
  __device__ __forceinline__ void body(int s, double2 &m, double2 &r0, double2 &r1) const {
    const double2 *q = reinterpret_cast<const double2 *>(p + s);
    m = __ldg(q + 1); r0 = __ldg(q + 2); r1 = __ldg(q + 3);
  }
};
// shared-memory tile accessors (local map tile staged once per match)
struct SmemSlot {
  uint32_t a;
  __device__ __forceinline__ int operator()(int i) const { return lds_s32(a + 4u * i); }
};
struct SmemRec {
  uint32_t a;
  __device__ __forceinline__ float4 head(int s) const { return lds_float4(a + 64u * s); }
  __device__ __forceinline__ void body(int s, double2 &m, double2 &r0, double2 &r1) const {
    m = lds_double2(a + 64u * s + 16u); r0 = lds_double2(a + 64u * s + 32u); r1 = lds_double2(a + 64u * s + 48u);
  }
};

// ------------------------------------------------------------------------------------------------
// cooperative reductions: every participating thread returns with bit-identical totals
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double shfl_xor_d(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }

template <int N>
__device__ __forceinline__ void warp_allreduce(double *v) {
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) {
#pragma unroll
    for (int k = 0; k < N; ++k) v[k] += shfl_xor_d(v[k], m);   // a + b == b + a: both partners agree
  }
}

struct WarpCoop {             // one warp per match
  int lane;
  __device__ __forceinline__ int rank() const { return lane; }
  __device__ __forceinline__ int size() const { return 32; }
  template <int N> __device__ __forceinline__ void allreduce(double *v) const { warp_allreduce<N>(v); }
  __device__ __forceinline__ void sync() const { __syncwarp(); }
};

struct BlockCoop {            // one CTA per match; scratch = [nwarps + 1][NACC] doubles in shared memory
  double *scratch;
  __device__ __forceinline__ int rank() const { return threadIdx.x; }
  __device__ __forceinline__ int size() const { return blockDim.x; }
  // warp butterflies -> one row per warp -> thread k adds column k over the warps in fixed order -> everyone reads the
  // N totals (broadcast loads). Two barriers, 8 + N shared loads per thread.
  template <int N> __device__ __forceinline__ void allreduce(double *v) const {
    warp_allreduce<N>(v);
    const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    double *total = scratch + nw * NACC;
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
      for (int k = 0; k < N; ++k) scratch[w * NACC + k] = v[k];
    }
    __syncthreads();
    if (threadIdx.x < N) {
      double s = scratch[threadIdx.x];
      for (int i = 1; i < nw; ++i) s += scratch[i * NACC + threadIdx.x];
      total[threadIdx.x] = s;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < N; ++k) v[k] = total[k];
  }
  __device__ __forceinline__ void sync() const { __syncthreads(); }
};

// ------------------------------------------------------------------------------------------------
// 3x3 solve like Eigen::JacobiSVD(H).solve(b): one-sided Jacobi SVD + pseudo-inverse with Eigen's
// default rank threshold (diagSize * epsilon * sigma_max, diagSize = 6 in PCL's 6x6 solve)
// ------------------------------------------------------------------------------------------------
static __device__ __noinline__ void svd_solve3_jacobi(const double *Hin, const double *b, double *x);

__device__ __forceinline__ void svd_solve3(const double *Hin, const double *b, double *x) {
  {
    // Fast path: a well-conditioned H has full rank under Eigen's threshold, and the pseudo-inverse
    // solution is H^-1 b. Adjugate solve (one division); the condition estimate
    // ||H||_F ||adj H||_F / |det| decides whether the Jacobi SVD below is needed at all.
    const double h00 = Hin[0], h01 = Hin[1], h02 = Hin[2], h10 = Hin[3], h11 = Hin[4], h12 = Hin[5],
                 h20 = Hin[6], h21 = Hin[7], h22 = Hin[8];
    const double c00 = h11 * h22 - h12 * h21, c01 = h12 * h20 - h10 * h22, c02 = h10 * h21 - h11 * h20;
    const double c10 = h02 * h21 - h01 * h22, c11 = h00 * h22 - h02 * h20, c12 = h01 * h20 - h00 * h21;
    const double c20 = h01 * h12 - h02 * h11, c21 = h02 * h10 - h00 * h12, c22 = h00 * h11 - h01 * h10;
    const double det = h00 * c00 + h01 * c01 + h02 * c02;
    const double nh = h00 * h00 + h01 * h01 + h02 * h02 + h10 * h10 + h11 * h11 + h12 * h12 + h20 * h20 + h21 * h21 + h22 * h22;
    const double na = c00 * c00 + c01 * c01 + c02 * c02 + c10 * c10 + c11 * c11 + c12 * c12 + c20 * c20 + c21 * c21 + c22 * c22;
    // cond^2 ~ nh * na / det^2 ; accept cond < 1e7 (1e14 squared). NaN / inf / zero fall through.
    if (nh * na < 1e14 * det * det) {
      const double id = 1.0 / det;
      // inverse = adj / det, adj = cofactor^T
      x[0] = (c00 * b[0] + c10 * b[1] + c20 * b[2]) * id;
      x[1] = (c01 * b[0] + c11 * b[1] + c21 * b[2]) * id;
      x[2] = (c02 * b[0] + c12 * b[1] + c22 * b[2]) * id;
      return;
    }
  }
  svd_solve3_jacobi(Hin, b, x);
}

static __device__ __noinline__ void svd_solve3_jacobi(const double *Hin, const double *b, double *x) {
  double A[3][3], V[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) { A[i][j] = Hin[i * 3 + j]; V[i][j] = (i == j) ? 1.0 : 0.0; }
  for (int sweep = 0; sweep < 30; ++sweep) {
    double off = 0.0;
#pragma unroll
    for (int p = 0; p < 2; ++p) {
#pragma unroll
      for (int q = p + 1; q < 3; ++q) {
        double alpha = 0, beta = 0, gamma = 0;
#pragma unroll
        for (int k = 0; k < 3; ++k) { alpha += A[k][p] * A[k][p]; beta += A[k][q] * A[k][q]; gamma += A[k][p] * A[k][q]; }
        if (gamma != 0.0) {
          off = fmax(off, fabs(gamma) / sqrt(alpha * beta));
          const double zeta = (beta - alpha) / (2.0 * gamma);
          const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
          const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            const double ap = A[k][p], aq = A[k][q];
            A[k][p] = c * ap - s * aq; A[k][q] = s * ap + c * aq;
            const double vp = V[k][p], vq = V[k][q];
            V[k][p] = c * vp - s * vq; V[k][q] = s * vp + c * vq;
          }
        }
      }
    }
    if (off < 1e-15) break;   // columns orthogonal to working precision (quadratic convergence: ~4-6 sweeps)
  }
  double sig[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) sig[j] = sqrt(A[0][j] * A[0][j] + A[1][j] * A[1][j] + A[2][j] * A[2][j]);
  const double smax = fmax(sig[0], fmax(sig[1], sig[2]));
  const double thr = fmax(smax * 6.0 * DBL_EPSILON, DBL_MIN);
  x[0] = x[1] = x[2] = 0.0;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    if (sig[j] > thr) {
      const double ub = (A[0][j] * b[0] + A[1][j] * b[1] + A[2][j] * b[2]) / sig[j];
      const double cf = ub / sig[j];
#pragma unroll
      for (int k = 0; k < 3; ++k) x[k] += V[k][j] * cf;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// More-Thuente helpers
// ------------------------------------------------------------------------------------------------
static __device__ __noinline__ double mt_trial_value(double a_l, double f_l, double g_l, double a_u, double f_u, double g_u,
                                        double a_t, double f_t, double g_t) {
  if (f_t > f_l) {
    const double z = 3 * (f_t - f_l) / (a_t - a_l) - g_t - g_l;
    const double w = sqrt(z * z - g_t * g_l);
    const double a_c = a_l + (a_t - a_l) * (w - g_l - z) / (g_t - g_l + 2 * w);
    const double a_q = a_l - 0.5 * (a_l - a_t) * g_l / (g_l - (f_l - f_t) / (a_l - a_t));
    if (fabs(a_c - a_l) < fabs(a_q - a_l)) return a_c;
    return 0.5 * (a_q + a_c);
  } else if (g_t * g_l < 0) {
    const double z = 3 * (f_t - f_l) / (a_t - a_l) - g_t - g_l;
    const double w = sqrt(z * z - g_t * g_l);
    const double a_c = a_l + (a_t - a_l) * (w - g_l - z) / (g_t - g_l + 2 * w);
    const double a_s = a_l - (a_l - a_t) / (g_l - g_t) * g_l;
    if (fabs(a_c - a_t) >= fabs(a_s - a_t)) return a_c;
    return a_s;
  } else if (fabs(g_t) <= fabs(g_l)) {
    const double z = 3 * (f_t - f_l) / (a_t - a_l) - g_t - g_l;
    const double w = sqrt(z * z - g_t * g_l);
    const double a_c = a_l + (a_t - a_l) * (w - g_l - z) / (g_t - g_l + 2 * w);
    const double a_s = a_l - (a_l - a_t) / (g_l - g_t) * g_l;
    const double a_n = (fabs(a_c - a_t) < fabs(a_s - a_t)) ? a_c : a_s;
    if (a_t > a_l) return fmin(a_t + 0.66 * (a_u - a_t), a_n);
    return fmax(a_t + 0.66 * (a_u - a_t), a_n);
  } else {
    const double z = 3 * (f_t - f_u) / (a_t - a_u) - g_t - g_u;
    const double w = sqrt(z * z - g_t * g_u);
    return a_u + (a_t - a_u) * (w - g_u - z) / (g_t - g_u + 2 * w);
  }
}

__device__ inline bool mt_update_interval(double &a_l, double &f_l, double &g_l, double &a_u, double &f_u,
                                          double &g_u, double a_t, double f_t, double g_t) {
  if (f_t > f_l) { a_u = a_t; f_u = f_t; g_u = g_t; return false; }
  if (g_t * (a_l - a_t) > 0) { a_l = a_t; f_l = f_t; g_l = g_t; return false; }
  if (g_t * (a_l - a_t) < 0) { a_u = a_l; f_u = f_l; g_u = g_l; a_l = a_t; f_l = f_t; g_l = g_t; return false; }
  return true;
}

// std::min / std::max semantics of the reference code path (NaN handling differs from fmin/fmax)
__device__ __forceinline__ double std_min(double a, double b) { return (b < a) ? b : a; }
__device__ __forceinline__ double std_max(double a, double b) { return (a < b) ? b : a; }

// ------------------------------------------------------------------------------------------------
// the matcher: Newton direction + More-Thuente line search, entirely on device.
// `Obj` evaluates one objective pass cooperatively: obj.pass(mode, p, angle_cache, out13)
// (MODE 2 must not touch score/gradient). All threads of the cooperating group call this with
// identical arguments and receive identical results.
// ------------------------------------------------------------------------------------------------
struct AngleCache { double cs, sn; };

__device__ __forceinline__ void angle_terms(const MatchParams &mp, double yaw, AngleCache &ac) {
  if ((mp.quirks & NDT_QUIRK_ANGLE_SNAP) && fabs(yaw) < 10e-5) { ac.cs = 1.0; ac.sn = 0.0; }
  else { sincos_once(yaw, &ac.sn, &ac.cs); }
}

struct MatchOut {
  double p[3];
  double score;
  double H[9];
  int converged, iters, evals, passes;   // evals: passes the reference makes; passes: the ones run here (see step_length_mt)
};

// Optimiser state of one match, kept together in one struct that the Newton loop and the line search share (a per-thread
// local: registers where they fit, local memory around the non-inlined objective passes). Passing the pieces around as
// separate by-reference scalars made the compiler spill 880 bytes per thread around every pass; as one object it is 250
// (C4 12.6 -> 11.8 ms on the same GPU).
// Round 2 experiments, all measured on the GPU and all rejected (profiles/r2_summary.md): (a) a per-warp copy in shared
// memory runs in k_align_pairs but raises "illegal instruction" in k_align_warp (bisected with profiles/smoke_variants.py:
// MatchOut in shared memory is fine, OptState is not; compute-sanitizer is closed on this pool, root cause open);
// (b) the optimiser as a state machine around one inlined objective call site keeps the state in registers but makes the
// compiler spill inside the hot loop (C4 11.8 -> 14.4 ms); (c) the same with the call not inlined grows the frame to 760 B.
struct OptState {
  double p[3], dp[3], g[3], H[9], x_t[3], acc[NACC];
  double score, a_l, f_l, g_l, a_u, f_u, g_u, a_t, phi_0, d_phi_0;
  AngleCache ac;
};

#ifndef NDT_TRIAL_MODE
#define NDT_TRIAL_MODE 0     // mode of a line-search trial pass: 0 = with Hessian (no computeHessian pass afterwards), 1 = PCL's split
#endif
template <class Obj>
__device__ inline double step_length_mt(Obj &obj, const MatchParams &mp, OptState &S, double step_init, int &evals, int &repeats) {
  const double *x = S.p;
  double *dir = S.dp, *g = S.g, *H = S.H, *x_t = S.x_t, *acc = S.acc;
  double &score = S.score;
  AngleCache &ac = S.ac;
  const double step_max = mp.step_size, step_min = mp.step_min;
  double &phi_0 = S.phi_0, &d_phi_0 = S.d_phi_0;
  phi_0 = -score;
  d_phi_0 = -(g[0] * dir[0] + g[1] * dir[1] + g[2] * dir[2]);
  if (d_phi_0 >= 0) {
    if (d_phi_0 == 0) return 0.0;
    d_phi_0 *= -1;
    dir[0] *= -1; dir[1] *= -1; dir[2] *= -1;
  }
  const int max_step_iterations = 10;
  int step_iterations = 0;
  const double mu = 1.e-4, nu = 0.9;
  double &a_l = S.a_l, &a_u = S.a_u, &f_l = S.f_l, &g_l = S.g_l, &f_u = S.f_u, &g_u = S.g_u, &a_t = S.a_t;
  a_l = 0; a_u = 0;
  f_l = 0; g_l = d_phi_0 - mu * d_phi_0;
  f_u = 0; g_u = d_phi_0 - mu * d_phi_0;
  bool interval_converged = (mp.quirks & NDT_QUIRK_MT_INTERVAL_LT0) ? ((step_max - step_min) < 0)
                                                                    : ((step_max - step_min) > 0);
  bool open_interval = true;
  a_t = step_init;
  a_t = std_min(a_t, step_max);
  a_t = std_max(a_t, step_min);
#pragma unroll
  for (int k = 0; k < 3; ++k) x_t[k] = x[k] + dir[k] * a_t;
  angle_terms(mp, x_t[2], ac);
  obj.pass(0, x_t, ac, acc); ++evals;
  score = acc[0]; g[0] = acc[1]; g[1] = acc[2]; g[2] = acc[3];
#pragma unroll
  for (int k = 0; k < 9; ++k) H[k] = acc[4 + k];
  double phi_t = -score;
  double d_phi_t = -(g[0] * dir[0] + g[1] * dir[1] + g[2] * dir[2]);
  double psi_t = phi_t - phi_0 - mu * d_phi_0 * a_t;
  double d_psi_t = d_phi_t - mu * d_phi_0;
  // A search whose interval has collapsed onto a clamp (typically a_t == step_min) asks for the same trial step again and
  // again until the iteration cap ends it: on C4 31 % of all passes are such repeats. x_t = x + dir * a_t is then the same
  // three doubles, so the pass would transform the same points with the same matrix and return the same score and
  // gradient bit for bit; it is not run, its result (still in score / g) is reused. `evals` keeps counting what the
  // reference executes, `repeats` what was skipped.
  double a_run = -1.0;                               // step of the last trial pass that ran (a_t >= step_min > 0)
  while (!interval_converged && step_iterations < max_step_iterations &&
         !(psi_t <= 0 && d_phi_t <= -nu * d_phi_0)) {
    a_t = mt_trial_value(a_l, f_l, g_l, a_u, f_u, g_u, a_t, open_interval ? psi_t : phi_t, open_interval ? d_psi_t : d_phi_t);
    a_t = std_min(a_t, step_max);
    a_t = std_max(a_t, step_min);
    ++evals;
    if (a_t == a_run) {
      ++repeats;
    } else {
#pragma unroll
      for (int k = 0; k < 3; ++k) x_t[k] = x[k] + dir[k] * a_t;
      angle_terms(mp, x_t[2], ac);
      obj.pass(NDT_TRIAL_MODE, x_t, ac, acc);
      score = acc[0]; g[0] = acc[1]; g[1] = acc[2]; g[2] = acc[3];
      a_run = a_t;
    }
    phi_t = -score;
    d_phi_t = -(g[0] * dir[0] + g[1] * dir[1] + g[2] * dir[2]);
    psi_t = phi_t - phi_0 - mu * d_phi_0 * a_t;
    d_psi_t = d_phi_t - mu * d_phi_0;
    if (open_interval && (psi_t <= 0 && d_psi_t >= 0)) {
      open_interval = false;
      f_l = f_l + phi_0 - mu * d_phi_0 * a_l; g_l = g_l + mu * d_phi_0;
      f_u = f_u + phi_0 - mu * d_phi_0 * a_u; g_u = g_u + mu * d_phi_0;
    }
    interval_converged = mt_update_interval(a_l, f_l, g_l, a_u, f_u, g_u, a_t, open_interval ? psi_t : phi_t,
                                            open_interval ? d_psi_t : d_phi_t);
    step_iterations++;
  }
  if (step_iterations) {
    // computeHessian at the last trial's x_t. The trial passes already accumulated the Hessian next to score and gradient
    // (same instruction stream as a Hessian-only pass, so the same bits): 92 % of the searches that try at all run one
    // distinct trial, and a second pass over the same transformed points would cost more than the nine extra sums did.
    ++evals;
#if NDT_TRIAL_MODE == 0
    ++repeats;
#else
    obj.pass(2, x_t, ac, acc);
#endif
#pragma unroll
    for (int k = 0; k < 9; ++k) H[k] = acc[4 + k];
  }
  return a_t;
}

template <class Obj>
__device__ inline void match_device(Obj &obj, const MatchParams &mp, const double *guess, MatchOut &mo, OptState &S) {
  // guess -> float matrix -> p: every component passes through float32 (Registration::align takes a Matrix4f)
  double *p = S.p, *g = S.g, *H = S.H, *dp = S.dp, *acc = S.acc;
  double &score = S.score;
  __syncwarp();
  p[0] = (double)(float)guess[0]; p[1] = (double)(float)guess[1]; p[2] = (double)(float)guess[2];
  int evals = 0, repeats = 0, nr_iterations = 0;
  bool converged = false;
  angle_terms(mp, p[2], S.ac);
  obj.pass(0, p, S.ac, acc); ++evals;
  score = acc[0]; g[0] = acc[1]; g[1] = acc[2]; g[2] = acc[3];
#pragma unroll
  for (int k = 0; k < 9; ++k) H[k] = acc[4 + k];
  while (!converged) {
    const double mg[3] = {-g[0], -g[1], -g[2]};
    double sol[3];
    const double Hc[9] = {H[0], H[1], H[2], H[3], H[4], H[5], H[6], H[7], H[8]};
    svd_solve3(Hc, mg, sol);
    const double nrm = sqrt(sol[0] * sol[0] + sol[1] * sol[1] + sol[2] * sol[2]);
    if (nrm == 0.0 || nrm != nrm) { converged = (nrm == nrm); break; }
    dp[0] = sol[0] / nrm; dp[1] = sol[1] / nrm; dp[2] = sol[2] / nrm;
    const double a = step_length_mt(obj, mp, S, nrm, evals, repeats);
#pragma unroll
    for (int k = 0; k < 3; ++k) { const double d = dp[k] * a; p[k] = p[k] + d; }
    if (nr_iterations > mp.max_iter || (nr_iterations && fabs(a) < mp.trans_eps)) converged = true;
    nr_iterations++;
  }
  mo.p[0] = p[0]; mo.p[1] = p[1]; mo.p[2] = p[2];
  mo.score = score;
#pragma unroll
  for (int k = 0; k < 9; ++k) mo.H[k] = H[k];
  mo.converged = converged ? 1 : 0;
  mo.iters = nr_iterations;
  mo.evals = evals;
  mo.passes = evals - repeats;
  __syncwarp();
}

// ------------------------------------------------------------------------------------------------
// Registration::getFitnessScore: exact float32 1-NN squared distance from a transformed source point to the
// target cloud (PCL: a second kd-tree over all target points). Here: ring search over the cell buckets built with
// the grid (the NDT cells, or a finer lattice when those are dense), exact because a ring is only trusted once every
// point outside the rings visited so far is provably farther than the best found.
//   stage 1  one query per lane, a few rings; the bucket descriptors of a ring are fetched eight at a time so the
//            loads of a ring overlap instead of forming one dependent chain per cell;
//   stage 2  queries stage 1 could not certify (points far from the map) are finished by the whole warp: one cell of
//            the ring per lane (sparse buckets) or all lanes striding over each bucket's points (dense buckets), one
//            warp-min per ring; exhaustive scan (lanes stride over the cloud) as last resort.
// ------------------------------------------------------------------------------------------------
constexpr int NN_STAGE1_RINGS_FINE = 6, NN_STAGE1_RINGS_COARSE = 3, NN_MAX_RINGS = 48;

// t-th cell of ring r around (0, 0), t in [0, 8r) (r = 0: the centre)
__device__ __forceinline__ void ring_cell(int r, int t, int &di, int &dj) {
  const int w = 2 * r + 1;
  if (t < w) { di = t - r; dj = -r; }
  else if (t < 2 * w) { di = t - w - r; dj = r; }
  else { const int u = t - 2 * w; dj = (u >> 1) - r + 1; di = (u & 1) ? r : -r; }
}

template <bool FINE>
__device__ __forceinline__ int2 nn_bucket(const GridView &G, int a, int b) {
  if (FINE) {
    if (a < 0 || a >= G.nn_div_x || b < 0 || b >= G.nn_div_y) return make_int2(0, 0);
    return __ldg(G.nn_range + (size_t)b * G.nn_div_x + a);
  }
  if (a < 0 || a >= G.div_x || b < 0 || b >= G.div_y) return make_int2(0, 0);
  const int lf = __ldg(G.leaf_id + G.table_base + (b + 2) * G.slot_w + a + 2);      // leaf id + 1, 0 = empty cell
  return lf > 0 ? __ldg(G.leaf_range + lf - 1) : make_int2(0, 0);
}

// one ring, one query per lane
template <bool FINE>
__device__ __forceinline__ void nn_ring(const GridView &G, int ci, int cj, int r, float xt, float yt, float &best) {
  const float2 *__restrict__ pts = FINE ? G.nn_pts : G.tgt_sorted;
  const int ncell = r == 0 ? 1 : 8 * r;
  for (int t0 = 0; t0 < ncell; t0 += 8) {
    int2 rg[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      int di = 0, dj = 0;
      if (r > 0) ring_cell(r, min(t0 + k, ncell - 1), di, dj);
      rg[k] = (t0 + k < ncell) ? nn_bucket<FINE>(G, ci + di, cj + dj) : make_int2(0, 0);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float2 *__restrict__ q = pts + rg[k].x;
      // four points per trip, loads issued together (a bucket along a wall holds ~10 points; one dependent L2 round trip
      // per point made this loop half of a warp-per-pair match); a NaN stand-in fails the compare
      for (int j = 0; j < rg[k].y; j += 4) {
        float2 t[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) t[u] = (j + u < rg[k].y) ? __ldg(q + j + u) : make_float2(NAN, NAN);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float dd = dist2f(xt, yt, t[u].x, t[u].y);
          if (dd < best) best = dd;
        }
      }
    }
  }
}

// rings that lie entirely outside the lattice hold nothing: the first ring that touches it
__device__ __forceinline__ int first_ring(int ci, int cj, int dvx, int dvy) {
  int ox = 0, oy = 0;
  if (ci < 0) ox = -ci; else if (ci >= dvx) ox = ci - (dvx - 1);
  if (cj < 0) oy = -cj; else if (cj >= dvy) oy = cj - (dvy - 1);
  return max(ox, oy);
}

// Every point outside rings 0..ring is at least ring cells plus the query's distance to the wall of its own cell away;
// 0.01 cell of slack covers the float rounding of the cell assignment (x * inv_leaf). margin = that distance in cells
// (0 for the warp-cooperative stage, which does not track it).
__device__ __forceinline__ bool ring_certifies(int ring, double leaf, float best, float margin = 0.f) {
  const double lim = ((double)ring + (double)margin - 0.01) * leaf;
  return lim > 0.0 && (double)best < lim * lim;
}
// distance (in cells) from a coordinate to the nearer wall of its own cell, by the same float arithmetic as cell_coord
__device__ __forceinline__ float cell_margin(float v, float inv) {
  const float t = __fmul_rn(v, inv);
  const float f = __fsub_rn(t, floorf(t));
  return fminf(f, 1.0f - f);
}

// stage 1: returns true when `best` is the exact answer
__device__ __forceinline__ bool nn_stage1(const GridView &G, float xt, float yt, float &best) {
  best = FLT_MAX;
  if (G.div_x <= 0) return false;
  if (G.nn_f > 0) {
    const int ci = cell_coord(xt, G.nn_inv_leaf, G.nn_min_bx), cj = cell_coord(yt, G.nn_inv_leaf, G.nn_min_by);
    const int r0 = first_ring(ci, cj, G.nn_div_x, G.nn_div_y);
    if (r0 > 2 * NN_STAGE1_RINGS_FINE) return false;
    const float margin = r0 == 0 ? fminf(cell_margin(xt, G.nn_inv_leaf), cell_margin(yt, G.nn_inv_leaf)) : 0.f;
    for (int ring = r0; ring <= r0 + NN_STAGE1_RINGS_FINE; ++ring) {
      nn_ring<true>(G, ci, cj, ring, xt, yt, best);
      if (ring_certifies(ring, (double)G.nn_leaf, best, margin)) return true;
    }
    return false;
  }
  const int ci = cell_coord(xt, G.inv_leaf, G.min_bx), cj = cell_coord(yt, G.inv_leaf, G.min_by);
  const int r0 = first_ring(ci, cj, G.div_x, G.div_y);
  if (r0 > 2 * NN_STAGE1_RINGS_COARSE) return false;
  const float margin = r0 == 0 ? fminf(cell_margin(xt, G.inv_leaf), cell_margin(yt, G.inv_leaf)) : 0.f;
  for (int ring = r0; ring <= r0 + NN_STAGE1_RINGS_COARSE; ++ring) {
    nn_ring<false>(G, ci, cj, ring, xt, yt, best);
    if (ring_certifies(ring, (double)G.leaf, best, margin)) return true;
  }
  return false;
}

// stage 2: the whole warp finishes one query (all lanes pass the same xt, yt, best); returns the exact answer
__device__ inline float nn_stage2_warp(const GridView &G, float xt, float yt, float best) {
  const int lane = threadIdx.x & 31;
  if (G.div_x > 0 && G.leaf_id == nullptr && G.nn_f > 0) {
    // incrementally maintained target: only the lattice is current. Lanes take the cells of a lattice ring, 32 at a time, and
    // every lane walks its own bucket; rings until the certificate holds.
    const int ci = cell_coord(xt, G.nn_inv_leaf, G.nn_min_bx), cj = cell_coord(yt, G.nn_inv_leaf, G.nn_min_by);
    const int r0 = first_ring(ci, cj, G.nn_div_x, G.nn_div_y);
    for (int ring = r0; ring <= r0 + NN_MAX_RINGS * max(G.nn_f, 1); ++ring) {
      const int ncell = ring == 0 ? 1 : 8 * ring;
      for (int t = lane; t < ncell; t += 32) {
        int di = 0, dj = 0;
        if (ring > 0) ring_cell(ring, t, di, dj);
        const int2 rg = nn_bucket<true>(G, ci + di, cj + dj);
        const float2 *__restrict__ q = G.nn_pts + rg.x;
        for (int j = 0; j < rg.y; ++j) {
          const float2 p = __ldg(q + j);
          const float dd = dist2f(xt, yt, p.x, p.y);
          if (dd < best) best = dd;
        }
      }
      best = __int_as_float(__reduce_min_sync(0xffffffffu, __float_as_int(best)));
      if (ring_certifies(ring, (double)G.nn_leaf, best)) return best;
    }
  } else if (G.div_x > 0) {
    const int ci = cell_coord(xt, G.inv_leaf, G.min_bx), cj = cell_coord(yt, G.inv_leaf, G.min_by);
    const int r0 = first_ring(ci, cj, G.div_x, G.div_y);
    const bool dense = G.nn_f > 0;          // dense NDT buckets (hundreds of points): lanes share a bucket instead of a ring
    for (int ring = r0; ring <= r0 + NN_MAX_RINGS; ++ring) {
      const int ncell = ring == 0 ? 1 : 8 * ring;
      if (dense) {
        // lanes fetch the bucket descriptors of 32 ring cells at once, then all lanes stride over the points of every
        // non-empty one (a ring of a building-sized map is mostly empty cells)
        for (int t0 = 0; t0 < ncell; t0 += 32) {
          int di = 0, dj = 0;
          const int t = t0 + lane;
          if (ring > 0 && t < ncell) ring_cell(ring, t, di, dj);
          const int2 mine = (t < ncell) ? nn_bucket<false>(G, ci + di, cj + dj) : make_int2(0, 0);
          unsigned full = __ballot_sync(0xffffffffu, mine.y > 0);
          while (full) {
            const int k = __ffs(full) - 1;
            full &= full - 1u;
            const int start = __shfl_sync(0xffffffffu, mine.x, k), cnt = __shfl_sync(0xffffffffu, mine.y, k);
            const float2 *__restrict__ q = G.tgt_sorted + start;
            for (int j = lane; j < cnt; j += 32) {
              const float2 p = __ldg(q + j);
              const float dd = dist2f(xt, yt, p.x, p.y);
              if (dd < best) best = dd;
            }
          }
        }
      } else {
        for (int t = lane; t < ncell; t += 32) {                       // one cell of the ring per lane
          int di = 0, dj = 0;
          if (ring > 0) ring_cell(ring, t, di, dj);
          const int2 rg = nn_bucket<false>(G, ci + di, cj + dj);
          const float2 *__restrict__ q = G.tgt_sorted + rg.x;
          for (int j = 0; j < rg.y; j += 4) {
            float2 p[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) p[u] = (j + u < rg.y) ? __ldg(q + j + u) : make_float2(NAN, NAN);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const float dd = dist2f(xt, yt, p[u].x, p[u].y);
              if (dd < best) best = dd;
            }
          }
        }
      }
      best = __int_as_float(__reduce_min_sync(0xffffffffu, __float_as_int(best)));   // best >= 0: int order == float order
      if (ring_certifies(ring, (double)G.leaf, best)) return best;
    }
  }
  // exhaustive fallback (query far from every occupied cell): still exact
  for (int64_t j = lane; j < G.n_tgt; j += 32) {
    const float4 t = __ldg(G.tgt + j);
    if (!isfinite(t.x) || !isfinite(t.y) || !isfinite(t.z)) continue;
    const float dd = dist2f(xt, yt, t.x, t.y);
    if (dd < best) best = dd;
  }
  return __int_as_float(__reduce_min_sync(0xffffffffu, __float_as_int(best)));
}

// Warp-collective: every lane of the warp must call this (valid = false for lanes without a query). Returns the
// squared distance for this lane's query.
__device__ inline float nn_dist2_warp(const GridView &G, float xt, float yt, bool valid) {
  float best = FLT_MAX;
  const bool sure = valid ? nn_stage1(G, xt, yt, best) : true;
  unsigned todo = __ballot_sync(0xffffffffu, !sure);
  while (todo) {
    const int l = __ffs(todo) - 1;
    todo &= todo - 1u;
    const float qx = __shfl_sync(0xffffffffu, xt, l), qy = __shfl_sync(0xffffffffu, yt, l), qb = __shfl_sync(0xffffffffu, best, l);
    const float r = nn_stage2_warp(G, qx, qy, qb);
    if ((int)(threadIdx.x & 31) == l) best = r;
  }
  return best;
}

}  // namespace ndt
