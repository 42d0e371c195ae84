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
  const int32_t *__restrict__ slot;     // [div_x * div_y] -> record index, -1 = not in the centroid tree
  const CellRec *__restrict__ recs;     // compact records of cells with n >= min_points
  int32_t min_bx, min_by, div_x, div_y;
  float inv_leaf;                       // 1.0f / leaf
  float r2;                             // (float)((double)leaf * leaf)
  float leaf;
  // 1-NN buckets (fitness): every occupied cell
  const int32_t *__restrict__ leaf_id;  // [div_x * div_y] -> leaf index or -1
  const int32_t *__restrict__ leaf_start;
  const int32_t *__restrict__ leaf_n;
  const int32_t *__restrict__ sorted_idx;
  const float4 *__restrict__ tgt;       // target points
  int64_t n_tgt;
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

__device__ __forceinline__ PoseF pose_to_float(const double p[3]) {
  PoseF f;
  const float yaw = (float)p[2];
  double sd, cd;
  sincos((double)yaw, &sd, &cd);
  f.c = (float)cd; f.s = (float)sd;
  f.tx = (float)p[0]; f.ty = (float)p[1];
  return f;
}

__device__ __forceinline__ void xform(const PoseF &f, bool sse_order, float x, float y, float &ox, float &oy) {
  const float ns = -f.s;
  if (sse_order) {
    ox = __fadd_rn(__fmul_rn(f.c, x), __fadd_rn(__fmul_rn(ns, y), f.tx));
    oy = __fadd_rn(__fmul_rn(f.s, x), __fadd_rn(__fmul_rn(f.c, y), f.ty));
  } else {
    ox = __fadd_rn(__fadd_rn(__fmul_rn(f.c, x), __fmul_rn(ns, y)), f.tx);
    oy = __fadd_rn(__fadd_rn(__fmul_rn(f.s, x), __fmul_rn(f.c, y)), f.ty);
  }
}

__device__ __forceinline__ float dist2f(float ax, float ay, float bx, float by) {
  const float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by);
  return __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
}

// ------------------------------------------------------------------------------------------------
// objective: one source point against its 3x3 neighbourhood
// MODE 0: score + gradient + Hessian; 1: score + gradient; 2: Hessian only (computeHessian)
// acc layout: [0] score, [1..3] gradient, [4..12] Hessian row-major
// ------------------------------------------------------------------------------------------------
constexpr int NACC = 13;

template <int MODE, class SlotLoad, class RecLoad>
__device__ __forceinline__ void eval_point(const GridView &G, const SlotLoad &slot_at, const RecLoad &rec_at,
                                           float xf, float yf, const PoseF &pf, bool sse_order, double cs,
                                           double sn, double d1, double d2, double *acc, int &pairs) {
  float xt, yt;
  xform(pf, sse_order, xf, yf, xt, yt);
  const int ci = cell_coord(xt, G.inv_leaf, G.min_bx);
  const int cj = cell_coord(yt, G.inv_leaf, G.min_by);
  if (ci < -1 || cj < -1 || ci > G.div_x || cj > G.div_y) return;
  const double x = (double)xf, y = (double)yf;
  const double Jx = -sn * x - cs * y, Jy = cs * x - sn * y;
  const double Hx = -cs * x + sn * y, Hy = -sn * x - cs * y;
  const double xtd = (double)xt, ytd = (double)yt;
#pragma unroll
  for (int dj = -1; dj <= 1; ++dj) {
    const int b = cj + dj;
    if (b < 0 || b >= G.div_y) continue;
#pragma unroll
    for (int di = -1; di <= 1; ++di) {
      const int a = ci + di;
      if (a < 0 || a >= G.div_x) continue;
      const int s = slot_at(b * G.div_x + a);
      if (s < 0) continue;
      const float4 head = rec_at.head(s);     // cx, cy, nr_points, cell
      if (!(dist2f(xt, yt, head.x, head.y) < G.r2)) continue;
      ++pairs;
      double2 m, r0, r1;
      rec_at.body(s, m, r0, r1);
      const double dx = xtd - m.x, dy = ytd - m.y;
      const double c00 = r0.x, c01 = r0.y, c10 = r1.x, c11 = r1.y;
      const double Cdx = c00 * dx + c01 * dy, Cdy = c10 * dx + c11 * dy;
      const double q = dx * Cdx + dy * Cdy;
      double e = exp(-d2 * q / 2.0);
      const double score_inc = -d1 * e;
      e = d2 * e;
      if (e > 1.0 || e < 0.0 || e != e) continue;
      e *= d1;
      // cov_dxd_pi = C * J_i for i = x, y, yaw
      const double CJ2x = c00 * Jx + c01 * Jy, CJ2y = c10 * Jx + c11 * Jy;
      const double a0 = dx * c00 + dy * c10;
      const double a1 = dx * c01 + dy * c11;
      const double a2 = dx * CJ2x + dy * CJ2y;
      if (MODE != 2) {
        acc[0] += score_inc;
        acc[1] += a0 * e; acc[2] += a1 * e; acc[3] += a2 * e;
      }
      if (MODE != 1) {
        const double CHx = c00 * Hx + c01 * Hy, CHy = c10 * Hx + c11 * Hy;
        const double dCH = dx * CHx + dy * CHy;
        // H(i,j) += e * (-d2 a_i a_j + J_j . (C J_i) [+ d.(C H_yawyaw) for i=j=yaw])
        acc[4]  += e * (-d2 * a0 * a0 + c00);
        acc[5]  += e * (-d2 * a0 * a1 + c10);
        acc[6]  += e * (-d2 * a0 * a2 + (Jx * c00 + Jy * c10));
        acc[7]  += e * (-d2 * a1 * a0 + c01);
        acc[8]  += e * (-d2 * a1 * a1 + c11);
        acc[9]  += e * (-d2 * a1 * a2 + (Jx * c01 + Jy * c11));
        acc[10] += e * (-d2 * a2 * a0 + CJ2x);
        acc[11] += e * (-d2 * a2 * a1 + CJ2y);
        acc[12] += e * (-d2 * a2 * a2 + (Jx * CJ2x + Jy * CJ2y) + dCH);
      }
    }
  }
}

// global-memory accessors (read-only path, L1/L2 cached)
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
  const int32_t *p;
  __device__ __forceinline__ int operator()(int i) const { return p[i]; }
};
struct SmemRec {
  const CellRec *p;
  __device__ __forceinline__ float4 head(int s) const { return *reinterpret_cast<const float4 *>(p + s); }
  __device__ __forceinline__ void body(int s, double2 &m, double2 &r0, double2 &r1) const {
    const double2 *q = reinterpret_cast<const double2 *>(p + s);
    m = q[1]; r0 = q[2]; r1 = q[3];
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

struct BlockCoop {            // one CTA per match; scratch = [nwarps][NACC] doubles in shared memory
  double *scratch;
  __device__ __forceinline__ int rank() const { return threadIdx.x; }
  __device__ __forceinline__ int size() const { return blockDim.x; }
  template <int N> __device__ __forceinline__ void allreduce(double *v) const {
    warp_allreduce<N>(v);
    const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
      for (int k = 0; k < N; ++k) scratch[w * NACC + k] = v[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < N; ++k) {
      double s = scratch[k];
      for (int i = 1; i < nw; ++i) s += scratch[i * NACC + k];
      v[k] = s;
    }
    __syncthreads();
  }
  __device__ __forceinline__ void sync() const { __syncthreads(); }
};

// ------------------------------------------------------------------------------------------------
// 3x3 solve like Eigen::JacobiSVD(H).solve(b): one-sided Jacobi SVD + pseudo-inverse with Eigen's
// default rank threshold (diagSize * epsilon * sigma_max, diagSize = 6 in PCL's 6x6 solve)
// ------------------------------------------------------------------------------------------------
__device__ inline void svd_solve3(const double *Hin, const double *b, double *x) {
  double A[3][3], V[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) { A[i][j] = Hin[i * 3 + j]; V[i][j] = (i == j) ? 1.0 : 0.0; }
  for (int sweep = 0; sweep < 60; ++sweep) {
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
    if (off < 1e-17) break;
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
__device__ inline double mt_trial_value(double a_l, double f_l, double g_l, double a_u, double f_u, double g_u,
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
// `Obj` evaluates one objective pass cooperatively: obj.template pass<MODE>(p, cs, sn, out13)
// (MODE 2 must not touch score/gradient). All threads of the cooperating group call this with
// identical arguments and receive identical results.
// ------------------------------------------------------------------------------------------------
struct AngleCache { double cs, sn; };

__device__ __forceinline__ void angle_terms(const MatchParams &mp, double yaw, AngleCache &ac) {
  if ((mp.quirks & NDT_QUIRK_ANGLE_SNAP) && fabs(yaw) < 10e-5) { ac.cs = 1.0; ac.sn = 0.0; }
  else { sincos(yaw, &ac.sn, &ac.cs); }
}

struct MatchOut {
  double p[3];
  double score;
  double H[9];
  int converged, iters, evals;
};

template <class Obj>
__device__ inline double step_length_mt(Obj &obj, const MatchParams &mp, const double *x, double *dir,
                                        double step_init, double &score, double *g, double *H, double *x_t,
                                        AngleCache &ac, int &evals) {
  const double step_max = mp.step_size, step_min = mp.step_min;
  const double phi_0 = -score;
  double d_phi_0 = -(g[0] * dir[0] + g[1] * dir[1] + g[2] * dir[2]);
  if (d_phi_0 >= 0) {
    if (d_phi_0 == 0) return 0.0;
    d_phi_0 *= -1;
    dir[0] *= -1; dir[1] *= -1; dir[2] *= -1;
  }
  const int max_step_iterations = 10;
  int step_iterations = 0;
  const double mu = 1.e-4, nu = 0.9;
  double a_l = 0, a_u = 0;
  double f_l = 0, g_l = d_phi_0 - mu * d_phi_0;
  double f_u = 0, g_u = d_phi_0 - mu * d_phi_0;
  bool interval_converged = (mp.quirks & NDT_QUIRK_MT_INTERVAL_LT0) ? ((step_max - step_min) < 0)
                                                                    : ((step_max - step_min) > 0);
  bool open_interval = true;
  double a_t = step_init;
  a_t = std_min(a_t, step_max);
  a_t = std_max(a_t, step_min);
  double acc[NACC];
#pragma unroll
  for (int k = 0; k < 3; ++k) x_t[k] = x[k] + dir[k] * a_t;
  angle_terms(mp, x_t[2], ac);
  obj.template pass<0>(x_t, ac, acc); ++evals;
  score = acc[0]; g[0] = acc[1]; g[1] = acc[2]; g[2] = acc[3];
#pragma unroll
  for (int k = 0; k < 9; ++k) H[k] = acc[4 + k];
  double phi_t = -score;
  double d_phi_t = -(g[0] * dir[0] + g[1] * dir[1] + g[2] * dir[2]);
  double psi_t = phi_t - phi_0 - mu * d_phi_0 * a_t;
  double d_psi_t = d_phi_t - mu * d_phi_0;
  while (!interval_converged && step_iterations < max_step_iterations &&
         !(psi_t <= 0 && d_phi_t <= -nu * d_phi_0)) {
    if (open_interval) a_t = mt_trial_value(a_l, f_l, g_l, a_u, f_u, g_u, a_t, psi_t, d_psi_t);
    else a_t = mt_trial_value(a_l, f_l, g_l, a_u, f_u, g_u, a_t, phi_t, d_phi_t);
    a_t = std_min(a_t, step_max);
    a_t = std_max(a_t, step_min);
#pragma unroll
    for (int k = 0; k < 3; ++k) x_t[k] = x[k] + dir[k] * a_t;
    angle_terms(mp, x_t[2], ac);
    obj.template pass<1>(x_t, ac, acc); ++evals;
    score = acc[0]; g[0] = acc[1]; g[1] = acc[2]; g[2] = acc[3];
    phi_t = -score;
    d_phi_t = -(g[0] * dir[0] + g[1] * dir[1] + g[2] * dir[2]);
    psi_t = phi_t - phi_0 - mu * d_phi_0 * a_t;
    d_psi_t = d_phi_t - mu * d_phi_0;
    if (open_interval && (psi_t <= 0 && d_psi_t >= 0)) {
      open_interval = false;
      f_l = f_l + phi_0 - mu * d_phi_0 * a_l; g_l = g_l + mu * d_phi_0;
      f_u = f_u + phi_0 - mu * d_phi_0 * a_u; g_u = g_u + mu * d_phi_0;
    }
    if (open_interval) interval_converged = mt_update_interval(a_l, f_l, g_l, a_u, f_u, g_u, a_t, psi_t, d_psi_t);
    else interval_converged = mt_update_interval(a_l, f_l, g_l, a_u, f_u, g_u, a_t, phi_t, d_phi_t);
    step_iterations++;
  }
  if (step_iterations) {
    // computeHessian: Hessian only, angle terms as cached by the last computeDerivatives (same x_t)
    obj.template pass<2>(x_t, ac, acc); ++evals;
#pragma unroll
    for (int k = 0; k < 9; ++k) H[k] = acc[4 + k];
  }
  return a_t;
}

template <class Obj>
__device__ inline void match_device(Obj &obj, const MatchParams &mp, const double *guess, MatchOut &mo) {
  // guess -> float matrix -> p: every component passes through float32 (Registration::align takes a Matrix4f)
  double p[3] = {(double)(float)guess[0], (double)(float)guess[1], (double)(float)guess[2]};
  double score, g[3], H[9], dp[3], acc[NACC];
  AngleCache ac;
  int evals = 0, nr_iterations = 0;
  bool converged = false;
  angle_terms(mp, p[2], ac);
  obj.template pass<0>(p, ac, acc); ++evals;
  score = acc[0]; g[0] = acc[1]; g[1] = acc[2]; g[2] = acc[3];
#pragma unroll
  for (int k = 0; k < 9; ++k) H[k] = acc[4 + k];
  while (!converged) {
    const double mg[3] = {-g[0], -g[1], -g[2]};
    svd_solve3(H, mg, dp);
    const double nrm = sqrt(dp[0] * dp[0] + dp[1] * dp[1] + dp[2] * dp[2]);
    if (nrm == 0.0 || nrm != nrm) { converged = (nrm == nrm); break; }
    dp[0] /= nrm; dp[1] /= nrm; dp[2] /= nrm;
    double x_t[3];
    const double a = step_length_mt(obj, mp, p, dp, nrm, score, g, H, x_t, ac, evals);
#pragma unroll
    for (int k = 0; k < 3; ++k) { dp[k] *= a; p[k] = p[k] + dp[k]; }
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
}

// ------------------------------------------------------------------------------------------------
// Registration::getFitnessScore: exact float 1-NN squared distance via ring search over the cell
// buckets built with the grid; exhaustive scan if nothing is found within `max_rings`.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void nn_visit(const GridView &G, int a, int b, float xt, float yt, float &best) {
  if (a < 0 || a >= G.div_x || b < 0 || b >= G.div_y) return;
  const int lf = __ldg(G.leaf_id + (size_t)b * G.div_x + a);
  if (lf < 0) return;
  const int st = __ldg(G.leaf_start + lf), n = __ldg(G.leaf_n + lf);
  for (int k = 0; k < n; ++k) {
    const float4 t = __ldg(G.tgt + __ldg(G.sorted_idx + st + k));
    const float dd = dist2f(xt, yt, t.x, t.y);
    if (dd < best) best = dd;
  }
}

__device__ inline float nn_dist2(const GridView &G, float xt, float yt, int max_rings) {
  float best = FLT_MAX;
  if (G.div_x > 0) {
    const int ci = cell_coord(xt, G.inv_leaf, G.min_bx);
    const int cj = cell_coord(yt, G.inv_leaf, G.min_by);
    // rings that lie entirely outside the grid hold nothing: start at the first ring that touches it
    int ox = 0, oy = 0;
    if (ci < 0) ox = -ci; else if (ci >= G.div_x) ox = ci - (G.div_x - 1);
    if (cj < 0) oy = -cj; else if (cj >= G.div_y) oy = cj - (G.div_y - 1);
    const int r_start = max(ox, oy);
    for (int ring = r_start; ring <= r_start + max_rings; ++ring) {
      if (ring == 0) {
        nn_visit(G, ci, cj, xt, yt, best);
      } else {
        for (int di = -ring; di <= ring; ++di) {
          nn_visit(G, ci + di, cj - ring, xt, yt, best);
          nn_visit(G, ci + di, cj + ring, xt, yt, best);
        }
        for (int dj = -ring + 1; dj <= ring - 1; ++dj) {
          nn_visit(G, ci - ring, cj + dj, xt, yt, best);
          nn_visit(G, ci + ring, cj + dj, xt, yt, best);
        }
        // every point outside rings 0..ring is at least (ring - 0.01) cells away
        const double lim = ((double)ring - 0.01) * (double)G.leaf;
        if ((double)best < lim * lim) return best;
      }
    }
  }
  // exhaustive fallback (query far from every occupied cell): still exact
  for (int64_t j = 0; j < G.n_tgt; ++j) {
    const float4 t = __ldg(G.tgt + j);
    if (!isfinite(t.x) || !isfinite(t.y) || !isfinite(t.z)) continue;
    const float dd = dist2f(xt, yt, t.x, t.y);
    if (dd < best) best = dd;
  }
  return best;
}

}  // namespace ndt
