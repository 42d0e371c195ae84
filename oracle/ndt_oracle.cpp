/*
 * ndt_oracle.cpp -- CPU ORACLE. TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library; the product (ndt_slam_b200/) never links, imports or calls it.
 *
 * PARITY STATUS: "parity unpinned" by the reference itself -- hibikid39/ndt_slam ships no tests,
 * golden vectors or fixtures, and the arithmetic lives in an un-vendored,
 * unpinned PCL (CMakeLists.txt:21, inferred 1.10.0). This file is a plain C++ restatement
 * (no Eigen, no PCL, no ROS; builds anywhere with g++) of the published PCL 1.10.0 algorithms
 * the reference calls, reduced exactly to the z = 0 case the reference feeds them
 * (SURVEY.md App. A, B). What pins it instead: the gauss_d1/d2 known answers, finite-difference
 * self-consistency, ground-truth recovery, and agreement with oracle/minipcl (a 6-DoF,
 * Eigen-based restatement compiled together with the reference's own sources into oracle/_ref).
 *
 * Reference call sites restated here:
 *   grid build      pcl::VoxelGridCovariance::applyFilter   <- ndt.setInputTarget   src/PoseEstimator.cpp:19
 *   source filter   pcl::ApproximateVoxelGrid::applyFilter  <- src/PoseEstimator.cpp:6-10, src/PointCloudMap.cpp:4-13
 *   objective       pcl::NDT::computeDerivatives/updateDerivatives/computeHessian <- ndt.align src/PoseEstimator.cpp:28, getHessian :56
 *   optimiser       pcl::NDT::computeTransformation/computeStepLengthMT/trialValueSelectionMT/updateIntervalMT <- :28
 *   fitness         pcl::Registration::getFitnessScore      <- src/PoseEstimator.cpp:43
 *   resampler       ScanPointResampler::resamplePoints      <- src/ScanPointResampler.cpp:4-62
 *   fusion          PoseFuser::fusePose/calOdometryCovariance <- src/PoseFuser.cpp:3-61
 *   pose algebra    Pose2D::calMotion/calPredPose, MyUtil::add_angle/sub_angle <- src/Pose2D.cpp:5-37, src/MyUtil.cpp:4-24
 */
#include "../include/ndt_b200.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <map>
#include <vector>

namespace {

// oracle-only switch (never set by the product): fp64 point transform, for finite-difference self-tests
constexpr int ORACLE_DEBUG_DOUBLE_TRANSFORM = 1 << 16;

struct Leaf {
  int n = 0;             // points accumulated (pass 1)
  int nr_points = 0;     // PCL's nr_points after pass 2 (-1 = failed eigen / inf check)
  double sx = 0, sy = 0; // mean_ accumulator (fp64)
  double sxx = 0, syx = 0, syy = 0; // cov_ accumulator (fp64)
  float cx = 0, cy = 0;  // centroid accumulator (fp32, input order)
  double mean[2] = {0, 0};
  double icov[4] = {0, 0, 0, 0}; // xx, xy, yx, yy (cov_.inverse() of a slightly asymmetric cov_)
  bool in_tree = false;  // n >= min_points: member of the centroid kd-tree
};

struct Oracle {
  ndt_params prm;
  // grid
  float leaf = 1.f, inv_leaf = 1.f;
  int min_b[2] = {0, 0}, div_b[2] = {0, 0};
  std::map<int64_t, Leaf> leaves;  // key = ijk0 + ijk1*div_b[0], like PCL's leaves_
  int64_t n_target = 0;
  std::vector<float> target;       // xyzw copy (for fitness)
  // 1-NN helper: target points bucketed by cell
  std::map<int64_t, std::vector<int>> tgt_bucket;
  // source
  std::vector<float> source;       // xyzw
  // gauss constants
  double d1 = 0, d2 = 0;
  // angle terms cached by the last computeDerivatives (computeHessian reuses them)
  double cs = 1, sn = 0;
  // stats
  int64_t outside3x3 = 0;          // hits found outside the 3x3 block around the point's own cell
  double min_boundary_gap = 1e300; // min |d^2 - r^2| over all radius tests of the last eval
  std::vector<double> trace;       // per objective pass: x, y, yaw, score, a_t, kind
  bool tracing = false;
  bool want_fitness = true;       // batched GPU calls (n >= 64) skip the fitness score; the CPU arm mirrors that
};

inline float fmul(float a, float b) { volatile float r = a * b; return r; }
inline float fadd(float a, float b) { volatile float r = a + b; return r; }

void gauss_constants(Oracle &o) {
  // PCL computeTransformation, SURVEY App. A.3. resolution_ is a float member.
  double r = (double)o.prm.resolution;
  double c1 = 10.0 * (1.0 - o.prm.outlier_ratio);
  double c2 = o.prm.outlier_ratio / std::pow(r, 3);
  double d3 = -std::log(c2);
  o.d1 = -std::log(c1 + c2) - d3;
  o.d2 = -2.0 * std::log((-std::log(c1 * std::exp(-0.5) + c2) - d3) / o.d1);
}

// VoxelGridCovariance pass-1 cell coordinate: float multiply, float floor, float subtract, truncate.
inline int cell_coord(float v, float inv, int min_b) {
  float t = std::floor(fmul(v, inv));
  return (int)(t - (float)min_b);
}
inline int abs_coord(float v, float inv) { return (int)std::floor(fmul(v, inv)); }

// symmetric 2x2 eigen-decomposition (a b; b d): ascending eigenvalues, orthonormal columns
void eig2(double a, double b, double d, double lam[2], double v0[2], double v1[2]) {
  double tr = a + d, df = a - d;
  double rt = std::sqrt(df * df + 4.0 * b * b);
  double l1 = 0.5 * (tr + rt), l0;
  // l0 via the stable product form when possible
  if (tr >= 0) { l1 = 0.5 * (tr + rt); l0 = (l1 != 0.0) ? (a * d - b * b) / l1 : 0.5 * (tr - rt); }
  else { l0 = 0.5 * (tr - rt); l1 = (l0 != 0.0) ? (a * d - b * b) / l0 : 0.5 * (tr + rt); }
  if (l0 > l1) std::swap(l0, l1);
  lam[0] = l0; lam[1] = l1;
  // eigenvector of the larger eigenvalue
  double ex, ey;
  if (std::fabs(b) > 0) {
    // (b, l1 - a) and (l1 - d, b) are both eigenvectors; take the better conditioned one
    if (std::fabs(l1 - a) > std::fabs(l1 - d)) { ex = b; ey = l1 - a; }
    else { ex = l1 - d; ey = b; }
    double nn = std::sqrt(ex * ex + ey * ey);
    if (nn == 0) { ex = 1; ey = 0; nn = 1; }
    ex /= nn; ey /= nn;
  } else {
    if (a >= d) { ex = 1; ey = 0; } else { ex = 0; ey = 1; }
  }
  v1[0] = ex; v1[1] = ey;
  v0[0] = -ey; v0[1] = ex;
}

void build_grid(Oracle &o, const float *p, int64_t n) {
  o.leaves.clear(); o.tgt_bucket.clear();
  o.target.assign(p, p + 4 * n);
  o.n_target = 0;
  o.leaf = o.prm.resolution;
  o.inv_leaf = 1.0f / o.leaf;
  const bool q_ident = o.prm.quirks & NDT_QUIRK_COV_INIT_IDENTITY;
  const bool q_nm1 = o.prm.quirks & NDT_QUIRK_COV_SCALE_NM1_N;
  // getMinMax3D over finite points (input cloud is_dense = false, PoseEstimator.h:94)
  float mn[2] = {std::numeric_limits<float>::max(), std::numeric_limits<float>::max()};
  float mx[2] = {-std::numeric_limits<float>::max(), -std::numeric_limits<float>::max()};
  int64_t nfin = 0;
  for (int64_t i = 0; i < n; ++i) {
    float x = p[4 * i], y = p[4 * i + 1], z = p[4 * i + 2];
    if (!std::isfinite(x) || !std::isfinite(y) || !std::isfinite(z)) continue;
    mn[0] = std::min(mn[0], x); mn[1] = std::min(mn[1], y);
    mx[0] = std::max(mx[0], x); mx[1] = std::max(mx[1], y);
    ++nfin;
  }
  o.div_b[0] = o.div_b[1] = 0; o.min_b[0] = o.min_b[1] = 0;
  if (nfin == 0) return;
  int64_t dx = (int64_t)((mx[0] - mn[0]) * o.inv_leaf) + 1;
  int64_t dy = (int64_t)((mx[1] - mn[1]) * o.inv_leaf) + 1;
  if (dx * dy > (int64_t)std::numeric_limits<int32_t>::max()) return;  // PCL warns, empty grid
  int max_b[2];
  for (int a = 0; a < 2; ++a) {
    o.min_b[a] = (int)std::floor(fmul(mn[a], o.inv_leaf));
    max_b[a] = (int)std::floor(fmul(mx[a], o.inv_leaf));
    o.div_b[a] = max_b[a] - o.min_b[a] + 1;
  }
  // pass 1
  for (int64_t i = 0; i < n; ++i) {
    float x = p[4 * i], y = p[4 * i + 1], z = p[4 * i + 2];
    if (!std::isfinite(x) || !std::isfinite(y) || !std::isfinite(z)) continue;
    int i0 = cell_coord(x, o.inv_leaf, o.min_b[0]);
    int i1 = cell_coord(y, o.inv_leaf, o.min_b[1]);
    int64_t idx = (int64_t)i0 + (int64_t)i1 * o.div_b[0];
    Leaf &l = o.leaves[idx];
    double xd = x, yd = y;
    l.sx += xd; l.sy += yd;
    l.sxx += xd * xd; l.syx += yd * xd; l.syy += yd * yd;
    l.cx = fadd(l.cx, x); l.cy = fadd(l.cy, y);
    ++l.n;
    o.tgt_bucket[idx].push_back((int)i);
    ++o.n_target;
  }
  // pass 2
  for (auto &kv : o.leaves) {
    Leaf &l = kv.second;
    const double nn = (double)l.n;
    l.cx = l.cx / (float)l.n; l.cy = l.cy / (float)l.n;
    const double psx = l.sx, psy = l.sy;  // pt_sum
    l.mean[0] = l.sx / nn; l.mean[1] = l.sy / nn;
    l.nr_points = l.n;
    if (l.n < o.prm.min_points) continue;
    l.in_tree = true;
    const double id = q_ident ? 1.0 : 0.0;
    double Cxx = id + l.sxx, Cyy = id + l.syy, Cyx = l.syx, Cxy = l.syx, Czz = id;
    double cxx, cxy, cyx, cyy, czz;
    const double m0 = l.mean[0], m1 = l.mean[1];
    if (q_nm1) {
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
    // SelfAdjointEigenSolver reads the lower triangle: (cxx, cyx, cyy) + zz, block diagonal for z = 0.
    double lam2[2], v0[2], v1[2];
    eig2(cxx, cyx, cyy, lam2, v0, v1);
    // three eigenvalues ascending: the z one and the two in-plane ones
    double ev[3] = {czz, lam2[0], lam2[1]};
    int which[3] = {2, 0, 1};  // 2 = z, 0/1 = in-plane index
    for (int a = 0; a < 3; ++a)
      for (int b = a + 1; b < 3; ++b)
        if (ev[b] < ev[a]) { std::swap(ev[a], ev[b]); std::swap(which[a], which[b]); }
    if (ev[0] < 0 || ev[1] < 0 || ev[2] <= 0) { l.nr_points = -1; continue; }
    const double mcv = o.prm.eig_mult * ev[2];
    if (ev[0] < mcv) {
      ev[0] = mcv;
      if (ev[1] < mcv) ev[1] = mcv;
      // cov = V * diag * V^-1, restricted to the (decoupled) in-plane block
      double l0 = lam2[0], l1 = lam2[1];
      for (int a = 0; a < 3; ++a) {
        if (which[a] == 0) l0 = ev[a];
        else if (which[a] == 1) l1 = ev[a];
        else czz = ev[a];
      }
      // V = [v0 v1]; V^-1 by cofactors
      double det = v0[0] * v1[1] - v1[0] * v0[1];
      double i00 = v1[1] / det, i01 = -v1[0] / det, i10 = -v0[1] / det, i11 = v0[0] / det;
      // (V * L) * V^-1
      double a00 = v0[0] * l0, a01 = v1[0] * l1, a10 = v0[1] * l0, a11 = v1[1] * l1;
      cxx = a00 * i00 + a01 * i10; cxy = a00 * i01 + a01 * i11;
      cyx = a10 * i00 + a11 * i10; cyy = a10 * i01 + a11 * i11;
    }
    // icov = cov.inverse()
    double det = cxx * cyy - cxy * cyx;
    l.icov[0] = cyy / det; l.icov[1] = -cxy / det; l.icov[2] = -cyx / det; l.icov[3] = cxx / det;
    double izz = 1.0 / czz;
    double mxc = std::max(std::max(std::max(l.icov[0], l.icov[1]), std::max(l.icov[2], l.icov[3])), std::max(izz, 0.0));
    double mnc = std::min(std::min(std::min(l.icov[0], l.icov[1]), std::min(l.icov[2], l.icov[3])), std::min(izz, 0.0));
    const double finf = (double)std::numeric_limits<float>::infinity();
    if (mxc == finf || mnc == -finf) l.nr_points = -1;
    if (l.nr_points == -1) { /* PCL keeps the (inf) icov; such leaves cannot occur for finite data */ }
  }
}

// A.6 float transform, no FMA. Default order (c*x + (-s)*y) + tx.
inline void xform(const Oracle &o, float c, float s, float tx, float ty, float x, float y, float &ox, float &oy) {
  float ns = -s;
  if (o.prm.quirks & NDT_QUIRK_TRANSFORM_SSE_ORDER) {
    ox = fadd(fmul(c, x), fadd(fmul(ns, y), tx));
    oy = fadd(fmul(s, x), fadd(fmul(c, y), ty));
  } else {
    ox = fadd(fadd(fmul(c, x), fmul(ns, y)), tx);
    oy = fadd(fadd(fmul(s, x), fmul(c, y)), ty);
  }
}

struct PoseF { float c, s, tx, ty; };
inline PoseF pose_to_float(const double p[3]) {
  PoseF f;
  float yaw = (float)p[2];
  f.c = (float)std::cos((double)yaw);
  f.s = (float)std::sin((double)yaw);
  f.tx = (float)p[0]; f.ty = (float)p[1];
  return f;
}

struct Hit { float d2; const Leaf *leaf; int order; };

// One objective pass. mode: 0 = computeDerivatives(hessian=true), 1 = computeDerivatives(false),
// 2 = computeHessian (angle terms NOT recomputed, score/gradient untouched).
void objective(Oracle &o, const double p[3], int mode, double &score, double g[3], double H[9], int64_t *n_pairs) {
  const int64_t ns = (int64_t)o.source.size() / 4;
  if (mode != 2) {
    // computeAngleDerivatives
    if ((o.prm.quirks & NDT_QUIRK_ANGLE_SNAP) && std::fabs(p[2]) < 10e-5) { o.cs = 1.0; o.sn = 0.0; }
    else { o.cs = std::cos(p[2]); o.sn = std::sin(p[2]); }
    score = 0; g[0] = g[1] = g[2] = 0;
  }
  if (mode != 1 || true) for (int k = 0; k < 9; ++k) H[k] = 0;  // computeDerivatives zeroes H even without hessian
  const bool want_h = (mode != 1);
  const double cs = o.cs, sn = o.sn;
  const PoseF pf = pose_to_float(p);
  const float r = o.prm.resolution;
  const float r2 = (float)((double)r * (double)r);
  int64_t pairs = 0;
  o.min_boundary_gap = 1e300;
  std::vector<Hit> hits;
  for (int64_t i = 0; i < ns; ++i) {
    const float xf = o.source[4 * i], yf = o.source[4 * i + 1];
    float xt, yt;
    xform(o, pf.c, pf.s, pf.tx, pf.ty, xf, yf, xt, yt);
    double xtd = (double)xt, ytd = (double)yt;
    if (o.prm.quirks & ORACLE_DEBUG_DOUBLE_TRANSFORM) {  // formula self-tests only (finite differences)
      const double cd = std::cos(p[2]), sd = std::sin(p[2]);
      xtd = cd * (double)xf - sd * (double)yf + p[0];
      ytd = sd * (double)xf + cd * (double)yf + p[1];
      xt = (float)xtd; yt = (float)ytd;
    }
    // radius search over the centroids of in-tree leaves: scan a 5x5 block (a centroid lies in its
    // own cell, so hits can only be in the 3x3 block; the outer ring is checked to prove it).
    if (o.div_b[0] == 0) continue;
    const int ci = cell_coord(xt, o.inv_leaf, o.min_b[0]);
    const int cj = cell_coord(yt, o.inv_leaf, o.min_b[1]);
    hits.clear();
    int order = 0;
    for (int dj = -2; dj <= 2; ++dj)
      for (int di = -2; di <= 2; ++di) {
        int a = ci + di, b = cj + dj;
        if (a < 0 || b < 0 || a >= o.div_b[0] || b >= o.div_b[1]) continue;
        auto it = o.leaves.find((int64_t)a + (int64_t)b * o.div_b[0]);
        if (it == o.leaves.end() || !it->second.in_tree) continue;
        const Leaf &l = it->second;
        float ddx = xt - l.cx, ddy = yt - l.cy;
        float dd = fadd(fmul(ddx, ddx), fmul(ddy, ddy));
        double gap = std::fabs((double)dd - (double)r2);
        if (gap < o.min_boundary_gap) o.min_boundary_gap = gap;
        if (dd < r2) {
          hits.push_back({dd, &l, order++});
          if (di < -1 || di > 1 || dj < -1 || dj > 1) ++o.outside3x3;
        }
      }
    // FLANN returns radius hits sorted by distance
    std::stable_sort(hits.begin(), hits.end(), [](const Hit &a, const Hit &b) { return a.d2 < b.d2; });
    const double x = (double)xf, y = (double)yf;
    const double Jx = -sn * x - cs * y, Jy = cs * x - sn * y;     // d x'/d yaw
    const double Hx = -cs * x + sn * y, Hy = -sn * x - cs * y;    // d2 x'/d yaw2
    for (const Hit &h : hits) {
      const Leaf &l = *h.leaf;
      ++pairs;
      const double dx = xtd - l.mean[0], dy = ytd - l.mean[1];
      const double c00 = l.icov[0], c01 = l.icov[1], c10 = l.icov[2], c11 = l.icov[3];
      const double Cdx = c00 * dx + c01 * dy, Cdy = c10 * dx + c11 * dy;
      const double q = dx * Cdx + dy * Cdy;
      double e = std::exp(-o.d2 * q / 2.0);
      const double score_inc = -o.d1 * e;
      e = o.d2 * e;
      if (e > 1 || e < 0 || e != e) continue;
      e *= o.d1;
      // cov_dxd_pi = C * J_i
      const double CJ[3][2] = {{c00, c10}, {c01, c11}, {c00 * Jx + c01 * Jy, c10 * Jx + c11 * Jy}};
      const double J[3][2] = {{1, 0}, {0, 1}, {Jx, Jy}};
      double a[3];
      for (int k = 0; k < 3; ++k) a[k] = dx * CJ[k][0] + dy * CJ[k][1];
      if (mode != 2) {
        score += score_inc;
        for (int k = 0; k < 3; ++k) g[k] += a[k] * e;
      }
      if (want_h) {
        const double CHx = c00 * Hx + c01 * Hy, CHy = c10 * Hx + c11 * Hy;
        const double dCH = dx * CHx + dy * CHy;
        for (int ii = 0; ii < 3; ++ii)
          for (int jj = 0; jj < 3; ++jj) {
            double t = -o.d2 * a[ii] * a[jj] + J[jj][0] * CJ[ii][0] + J[jj][1] * CJ[ii][1];
            if (ii == 2 && jj == 2) t += dCH;
            H[ii * 3 + jj] += e * t;
          }
      }
    }
  }
  if (n_pairs) *n_pairs = pairs;
}

// Solve H x = b like Eigen::JacobiSVD(H).solve(b): pseudo-inverse with the default rank threshold.
// One-sided Jacobi SVD of the 3x3.
void svd_solve3(const double Hin[9], const double b[3], double x[3]) {
  double A[3][3], V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) A[i][j] = Hin[i * 3 + j];
  for (int sweep = 0; sweep < 30; ++sweep) {
    double off = 0;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        double alpha = 0, beta = 0, gamma = 0;
        for (int k = 0; k < 3; ++k) { alpha += A[k][p] * A[k][p]; beta += A[k][q] * A[k][q]; gamma += A[k][p] * A[k][q]; }
        if (gamma == 0) continue;
        off = std::max(off, std::fabs(gamma) / std::sqrt(alpha * beta));
        double zeta = (beta - alpha) / (2.0 * gamma);
        double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
        double c = 1.0 / std::sqrt(1.0 + t * t), s = c * t;
        for (int k = 0; k < 3; ++k) {
          double ap = A[k][p], aq = A[k][q];
          A[k][p] = c * ap - s * aq; A[k][q] = s * ap + c * aq;
          double vp = V[k][p], vq = V[k][q];
          V[k][p] = c * vp - s * vq; V[k][q] = s * vp + c * vq;
        }
      }
    if (off < 1e-15) break;
  }
  double sig[3];
  for (int j = 0; j < 3; ++j) sig[j] = std::sqrt(A[0][j] * A[0][j] + A[1][j] * A[1][j] + A[2][j] * A[2][j]);
  double smax = std::max(sig[0], std::max(sig[1], sig[2]));
  // Eigen: threshold = diagSize * epsilon (diagSize = 6 for the 6x6 PCL solves), premultiplied by smax
  double thr = std::max(smax * 6.0 * std::numeric_limits<double>::epsilon(), std::numeric_limits<double>::min());
  x[0] = x[1] = x[2] = 0;
  for (int j = 0; j < 3; ++j) {
    if (!(sig[j] > thr)) continue;
    // u_j = A[:,j]/sig_j ; coefficient = (u_j . b)/sig_j
    double ub = (A[0][j] * b[0] + A[1][j] * b[1] + A[2][j] * b[2]) / sig[j];
    double cf = ub / sig[j];
    for (int k = 0; k < 3; ++k) x[k] += V[k][j] * cf;
  }
}

// ---- More-Thuente (PCL ndt.hpp) -------------------------------------------
double trial_value(double a_l, double f_l, double g_l, double a_u, double f_u, double g_u,
                   double a_t, double f_t, double g_t) {
  if (f_t > f_l) {  // case 1
    double z = 3 * (f_t - f_l) / (a_t - a_l) - g_t - g_l;
    double w = std::sqrt(z * z - g_t * g_l);
    double a_c = a_l + (a_t - a_l) * (w - g_l - z) / (g_t - g_l + 2 * w);
    double a_q = a_l - 0.5 * (a_l - a_t) * g_l / (g_l - (f_l - f_t) / (a_l - a_t));
    if (std::fabs(a_c - a_l) < std::fabs(a_q - a_l)) return a_c;
    return 0.5 * (a_q + a_c);
  } else if (g_t * g_l < 0) {  // case 2
    double z = 3 * (f_t - f_l) / (a_t - a_l) - g_t - g_l;
    double w = std::sqrt(z * z - g_t * g_l);
    double a_c = a_l + (a_t - a_l) * (w - g_l - z) / (g_t - g_l + 2 * w);
    double a_s = a_l - (a_l - a_t) / (g_l - g_t) * g_l;
    if (std::fabs(a_c - a_t) >= std::fabs(a_s - a_t)) return a_c;
    return a_s;
  } else if (std::fabs(g_t) <= std::fabs(g_l)) {  // case 3
    double z = 3 * (f_t - f_l) / (a_t - a_l) - g_t - g_l;
    double w = std::sqrt(z * z - g_t * g_l);
    double a_c = a_l + (a_t - a_l) * (w - g_l - z) / (g_t - g_l + 2 * w);
    double a_s = a_l - (a_l - a_t) / (g_l - g_t) * g_l;
    double a_n = (std::fabs(a_c - a_t) < std::fabs(a_s - a_t)) ? a_c : a_s;
    if (a_t > a_l) return std::min(a_t + 0.66 * (a_u - a_t), a_n);
    return std::max(a_t + 0.66 * (a_u - a_t), a_n);
  } else {  // case 4
    double z = 3 * (f_t - f_u) / (a_t - a_u) - g_t - g_u;
    double w = std::sqrt(z * z - g_t * g_u);
    return a_u + (a_t - a_u) * (w - g_u - z) / (g_t - g_u + 2 * w);
  }
}

bool update_interval(double &a_l, double &f_l, double &g_l, double &a_u, double &f_u, double &g_u,
                     double a_t, double f_t, double g_t) {
  if (f_t > f_l) { a_u = a_t; f_u = f_t; g_u = g_t; return false; }
  if (g_t * (a_l - a_t) > 0) { a_l = a_t; f_l = f_t; g_l = g_t; return false; }
  if (g_t * (a_l - a_t) < 0) { a_u = a_l; f_u = f_l; g_u = g_l; a_l = a_t; f_l = f_t; g_l = g_t; return false; }
  return true;
}

struct MatchState { int evals = 0; };

void trace_push(Oracle &o, const double p[3], double score, double a_t, int kind) {
  if (!o.tracing) return;
  o.trace.push_back(p[0]); o.trace.push_back(p[1]); o.trace.push_back(p[2]);
  o.trace.push_back(score); o.trace.push_back(a_t); o.trace.push_back((double)kind);
}

double step_length_mt(Oracle &o, const double x[3], double dir[3], double step_init, double step_max,
                      double step_min, double &score, double g[3], double H[9], double x_t[3], MatchState &ms) {
  double phi_0 = -score;
  double d_phi_0 = -(g[0] * dir[0] + g[1] * dir[1] + g[2] * dir[2]);
  if (d_phi_0 >= 0) {
    if (d_phi_0 == 0) return 0;
    d_phi_0 *= -1;
    for (int k = 0; k < 3; ++k) dir[k] *= -1;
  }
  const int max_step_iterations = 10;
  int step_iterations = 0;
  const double mu = 1.e-4, nu = 0.9;
  double a_l = 0, a_u = 0;
  double f_l = 0 /* psi(0) */, g_l = d_phi_0 - mu * d_phi_0;
  double f_u = 0, g_u = d_phi_0 - mu * d_phi_0;
  bool interval_converged = (o.prm.quirks & NDT_QUIRK_MT_INTERVAL_LT0) ? ((step_max - step_min) < 0)
                                                                        : ((step_max - step_min) > 0);
  bool open_interval = true;
  double a_t = step_init;
  a_t = std::min(a_t, step_max);
  a_t = std::max(a_t, step_min);
  for (int k = 0; k < 3; ++k) x_t[k] = x[k] + dir[k] * a_t;
  objective(o, x_t, 0, score, g, H, nullptr); ++ms.evals;
  trace_push(o, x_t, score, a_t, 0);
  double phi_t = -score;
  double d_phi_t = -(g[0] * dir[0] + g[1] * dir[1] + g[2] * dir[2]);
  double psi_t = phi_t - phi_0 - mu * d_phi_0 * a_t;
  double d_psi_t = d_phi_t - mu * d_phi_0;
  while (!interval_converged && step_iterations < max_step_iterations &&
         !(psi_t <= 0 && d_phi_t <= -nu * d_phi_0)) {
    if (open_interval) a_t = trial_value(a_l, f_l, g_l, a_u, f_u, g_u, a_t, psi_t, d_psi_t);
    else a_t = trial_value(a_l, f_l, g_l, a_u, f_u, g_u, a_t, phi_t, d_phi_t);
    a_t = std::min(a_t, step_max);
    a_t = std::max(a_t, step_min);
    for (int k = 0; k < 3; ++k) x_t[k] = x[k] + dir[k] * a_t;
    objective(o, x_t, 1, score, g, H, nullptr); ++ms.evals;
    trace_push(o, x_t, score, a_t, 1);
    phi_t = -score;
    d_phi_t = -(g[0] * dir[0] + g[1] * dir[1] + g[2] * dir[2]);
    psi_t = phi_t - phi_0 - mu * d_phi_0 * a_t;
    d_psi_t = d_phi_t - mu * d_phi_0;
    if (open_interval && (psi_t <= 0 && d_psi_t >= 0)) {
      open_interval = false;
      f_l = f_l + phi_0 - mu * d_phi_0 * a_l; g_l = g_l + mu * d_phi_0;
      f_u = f_u + phi_0 - mu * d_phi_0 * a_u; g_u = g_u + mu * d_phi_0;
    }
    if (open_interval) interval_converged = update_interval(a_l, f_l, g_l, a_u, f_u, g_u, a_t, psi_t, d_psi_t);
    else interval_converged = update_interval(a_l, f_l, g_l, a_u, f_u, g_u, a_t, phi_t, d_phi_t);
    step_iterations++;
  }
  if (step_iterations) {
    double dummy_s = 0, dummy_g[3];
    objective(o, x_t, 2, dummy_s, dummy_g, H, nullptr); ++ms.evals;
    trace_push(o, x_t, score, a_t, 2);
  }
  return a_t;
}

double fitness_score(Oracle &o, const double p[3]);

void align(Oracle &o, const double guess[3], ndt_result *res) {
  gauss_constants(o);
  MatchState ms;
  const int64_t ns = (int64_t)o.source.size() / 4;
  // guess -> float matrix -> p (project definition: each component rounded through float)
  double p[3] = {(double)(float)guess[0], (double)(float)guess[1], (double)(float)guess[2]};
  double score = 0, g[3], H[9], dp[3];
  int nr_iterations = 0;
  bool converged = false;
  objective(o, p, 0, score, g, H, nullptr); ++ms.evals;
  trace_push(o, p, score, 0.0, 0);
  while (!converged) {
    double mg[3] = {-g[0], -g[1], -g[2]};
    svd_solve3(H, mg, dp);
    double nrm = std::sqrt(dp[0] * dp[0] + dp[1] * dp[1] + dp[2] * dp[2]);
    if (nrm == 0 || nrm != nrm) { converged = (nrm == nrm); break; }
    for (int k = 0; k < 3; ++k) dp[k] /= nrm;
    double x_t[3];
    double a = step_length_mt(o, p, dp, nrm, o.prm.step_size, o.prm.trans_eps / 2, score, g, H, x_t, ms);
    for (int k = 0; k < 3; ++k) { dp[k] *= a; p[k] = p[k] + dp[k]; }
    if (nr_iterations > o.prm.max_iter || (nr_iterations && std::fabs(a) < o.prm.trans_eps)) converged = true;
    nr_iterations++;
  }
  std::memset(res, 0, sizeof(*res));
  res->pose[0] = p[0]; res->pose[1] = p[1]; res->pose[2] = p[2];
  PoseF pf = pose_to_float(p);
  float *T = res->T;
  for (int k = 0; k < 16; ++k) T[k] = 0.f;
  T[0] = pf.c; T[1] = pf.s; T[4] = -pf.s; T[5] = pf.c; T[10] = 1.f; T[15] = 1.f; T[12] = pf.tx; T[13] = pf.ty;
  res->score = score;
  res->trans_prob = ns ? score / (double)ns : 0.0;
  for (int k = 0; k < 9; ++k) res->hess[k] = H[k];
  res->converged = converged ? 1 : 0;
  res->iters = nr_iterations;
  res->evals = ms.evals;
  res->point_evals = (int64_t)ms.evals * ns;
  res->fitness = o.want_fitness ? fitness_score(o, p) : std::nan("");
}

// Registration::getFitnessScore: mean float squared distance to the nearest target point (all points).
double fitness_score(Oracle &o, const double p[3]) {
  const int64_t ns = (int64_t)o.source.size() / 4;
  const int64_t nt = (int64_t)o.target.size() / 4;
  if (nt == 0 || ns == 0) return std::numeric_limits<double>::max();
  PoseF pf = pose_to_float(p);
  double sum = 0; int64_t nr = 0;
  for (int64_t i = 0; i < ns; ++i) {
    float xt, yt;
    xform(o, pf.c, pf.s, pf.tx, pf.ty, o.source[4 * i], o.source[4 * i + 1], xt, yt);
    float best = std::numeric_limits<float>::max();
    // ring search over the cell buckets, then exhaustive fallback
    bool done = false;
    if (o.div_b[0] > 0) {
      const int ci = cell_coord(xt, o.inv_leaf, o.min_b[0]);
      const int cj = cell_coord(yt, o.inv_leaf, o.min_b[1]);
      for (int ring = 0; ring <= 6 && !done; ++ring) {
        for (int dj = -ring; dj <= ring; ++dj)
          for (int di = -ring; di <= ring; ++di) {
            if (std::max(std::abs(di), std::abs(dj)) != ring) continue;
            int a = ci + di, b = cj + dj;
            if (a < 0 || b < 0 || a >= o.div_b[0] || b >= o.div_b[1]) continue;
            auto it = o.tgt_bucket.find((int64_t)a + (int64_t)b * o.div_b[0]);
            if (it == o.tgt_bucket.end()) continue;
            for (int j : it->second) {
              float ddx = xt - o.target[4 * j], ddy = yt - o.target[4 * j + 1];
              float dd = fadd(fmul(ddx, ddx), fmul(ddy, ddy));
              if (dd < best) best = dd;
            }
          }
        // everything outside rings 0..ring is at least (ring - 0.01) cells away
        double lim = ((double)ring - 0.01) * (double)o.leaf;
        if (ring >= 1 && (double)best < lim * lim) done = true;
      }
    }
    if (!done) {
      for (int64_t j = 0; j < nt; ++j) {
        float tx = o.target[4 * j], ty = o.target[4 * j + 1];
        if (!std::isfinite(tx) || !std::isfinite(ty)) continue;
        float ddx = xt - tx, ddy = yt - ty;
        float dd = fadd(fmul(ddx, ddx), fmul(ddy, ddy));
        if (dd < best) best = dd;
      }
    }
    sum += (double)best; ++nr;
  }
  return nr ? sum / (double)nr : std::numeric_limits<double>::max();
}

}  // namespace

extern "C" {

void *oracle_create(const ndt_params *p) {
  Oracle *o = new Oracle();
  o->prm = *p;
  gauss_constants(*o);
  return o;
}
void oracle_destroy(void *h) { delete (Oracle *)h; }
void oracle_want_fitness(void *h, int on) { ((Oracle *)h)->want_fitness = on != 0; }

void oracle_gauss(void *h, double out[2]) { Oracle *o = (Oracle *)h; gauss_constants(*o); out[0] = o->d1; out[1] = o->d2; }

void oracle_set_target(void *h, const float *xyzw, int64_t n) { build_grid(*(Oracle *)h, xyzw, n); }
void oracle_set_source(void *h, const float *xyzw, int64_t n) { ((Oracle *)h)->source.assign(xyzw, xyzw + 4 * n); }

void oracle_grid_info(void *h, ndt_grid_info *gi) {
  Oracle *o = (Oracle *)h;
  std::memset(gi, 0, sizeof(*gi));
  gi->min_b[0] = o->min_b[0]; gi->min_b[1] = o->min_b[1];
  gi->div_b[0] = o->div_b[0]; gi->div_b[1] = o->div_b[1];
  gi->n_points = o->n_target;
  gi->n_leaves = (int32_t)o->leaves.size();
  for (auto &kv : o->leaves) { if (kv.second.in_tree) { gi->n_slots++; if (kv.second.nr_points > 0) gi->n_valid++; } }
}

int64_t oracle_grid_readback(void *h, int64_t cap, int32_t *cell_idx, int32_t *nr_points, double *mean2,
                             double *icov4, float *centroid2) {
  Oracle *o = (Oracle *)h;
  int64_t k = 0;
  for (auto &kv : o->leaves) {
    if (k >= cap) break;
    const Leaf &l = kv.second;
    if (cell_idx) cell_idx[k] = (int32_t)kv.first;
    if (nr_points) nr_points[k] = l.nr_points;
    if (mean2) { mean2[2 * k] = l.mean[0]; mean2[2 * k + 1] = l.mean[1]; }
    if (icov4) for (int a = 0; a < 4; ++a) icov4[4 * k + a] = l.icov[a];
    if (centroid2) { centroid2[2 * k] = l.cx; centroid2[2 * k + 1] = l.cy; }
    ++k;
  }
  return (int64_t)o->leaves.size();
}

void oracle_cell_index(void *h, const float *xyzw, int64_t n, int32_t *idx) {
  Oracle *o = (Oracle *)h;
  for (int64_t i = 0; i < n; ++i) {
    int i0 = cell_coord(xyzw[4 * i], o->inv_leaf, o->min_b[0]);
    int i1 = cell_coord(xyzw[4 * i + 1], o->inv_leaf, o->min_b[1]);
    idx[i] = i0 + i1 * o->div_b[0];
  }
}

void oracle_eval(void *h, const double pose[3], int want_hessian, ndt_eval_out *out) {
  Oracle *o = (Oracle *)h;
  gauss_constants(*o);
  double p[3] = {pose[0], pose[1], pose[2]};
  objective(*o, p, want_hessian ? 0 : 1, out->score, out->grad, out->hess, &out->n_pairs);
}

void oracle_eval_stats(void *h, int64_t *outside3x3, double *min_gap) {
  Oracle *o = (Oracle *)h;
  if (outside3x3) *outside3x3 = o->outside3x3;
  if (min_gap) *min_gap = o->min_boundary_gap;
}

void oracle_align(void *h, const double guess[3], ndt_result *res) {
  Oracle *o = (Oracle *)h;
  o->tracing = false;
  align(*o, guess, res);
}

// trace rows: x, y, yaw, score, a_t, kind(0 = derivatives+H, 1 = derivatives, 2 = hessian only)
int64_t oracle_align_trace(void *h, const double guess[3], ndt_result *res, double *trace6, int64_t cap_rows) {
  Oracle *o = (Oracle *)h;
  o->tracing = true; o->trace.clear();
  align(*o, guess, res);
  o->tracing = false;
  int64_t rows = (int64_t)o->trace.size() / 6;
  int64_t m = std::min(rows, cap_rows);
  if (trace6) std::memcpy(trace6, o->trace.data(), sizeof(double) * 6 * m);
  return rows;
}

double oracle_fitness(void *h, const double pose[3]) { return fitness_score(*(Oracle *)h, pose); }

// pcl::ApproximateVoxelGrid<PointXYZ>::applyFilter. out must hold n points.
int64_t oracle_approx_voxel_filter(const float *in, int64_t n, float leaf, float *out) {
  struct He { int ix, iy, iz, count; float cx, cy, cz; };
  const int hist = 512;
  std::vector<He> he(hist);
  for (auto &e : he) { e.ix = e.iy = e.iz = 0; e.count = 0; e.cx = e.cy = e.cz = 0.f; }
  const float inv = 1.0f / leaf;
  int64_t op = 0;
  auto flush = [&](He &e) {
    float c = (float)e.count;
    out[4 * op] = e.cx / c; out[4 * op + 1] = e.cy / c; out[4 * op + 2] = e.cz / c; out[4 * op + 3] = 0.f;
    ++op;
  };
  for (int64_t i = 0; i < n; ++i) {
    float x = in[4 * i], y = in[4 * i + 1], z = in[4 * i + 2];
    int ix = (int)std::floor(fmul(x, inv)), iy = (int)std::floor(fmul(y, inv)), iz = (int)std::floor(fmul(z, inv));
    unsigned hash = (unsigned)((uint32_t)ix * 7171u + (uint32_t)iy * 3079u + (uint32_t)iz * 4231u) & (hist - 1);
    He &e = he[hash];
    if (e.count && (ix != e.ix || iy != e.iy || iz != e.iz)) { flush(e); e.count = 0; e.cx = e.cy = e.cz = 0.f; }
    e.ix = ix; e.iy = iy; e.iz = iz; e.count++;
    e.cx = fadd(e.cx, x); e.cy = fadd(e.cy, y); e.cz = fadd(e.cz, z);
  }
  for (int k = 0; k < hist; ++k) if (he[k].count) flush(he[k]);
  return op;
}

// ScanPointResampler::resamplePoints (src/ScanPointResampler.cpp:4-62). xy: n x 2 doubles, out holds
// up to cap points; returns the resampled count (or -needed if cap is too small).
int64_t oracle_resample(const double *xy, int64_t n, double space, double space_thre, double *out, int64_t cap) {
  if (n == 0) return 0;
  std::vector<double> o;
  double dis = 0;
  double px = xy[0], py = xy[1];
  o.push_back(px); o.push_back(py);
  for (int64_t i = 1; i < n; ++i) {
    double cx = xy[2 * i], cy = xy[2 * i + 1];
    double dx = cx - px, dy = cy - py;
    double L = std::sqrt(dx * dx + dy * dy);
    if (dis + L < space) { dis += L; px = cx; py = cy; continue; }
    double nx, ny; bool inserted = false;
    if (dis + L >= space_thre) { nx = cx; ny = cy; }
    else { double ratio = (space - dis) / L; nx = dx * ratio + px; ny = dy * ratio + py; inserted = true; }
    o.push_back(nx); o.push_back(ny);
    px = nx; py = ny; dis = 0;
    if (inserted) --i;
  }
  int64_t m = (int64_t)o.size() / 2;
  if (m > cap) return -m;
  std::memcpy(out, o.data(), sizeof(double) * o.size());
  return m;
}

// MyUtil::add_angle / sub_angle (src/MyUtil.cpp:4-24), degrees.
double oracle_add_angle(double a1, double a2) { double s = a1 + a2; if (s < -180) s += 360; else if (s >= 180) s -= 360; return s; }
double oracle_sub_angle(double a1, double a2) { double d = a1 - a2; if (d < -180) d += 360; else if (d >= 180) d -= 360; return d; }

static const double kPi = 3.14159265358979323846;
static inline double deg2rad(double x) { return x * kPi / 180; }
static inline double rad2deg(double x) { return x * 180 / kPi; }

// Pose2D::calMotion (src/Pose2D.cpp:5-14): poses are (tx, ty, th_deg)
void oracle_cal_motion(const double cur[3], const double prev[3], double motion[3]) {
  double a = deg2rad(prev[2]);
  double c = std::cos(a), s = std::sin(a);
  double dx = cur[0] - prev[0], dy = cur[1] - prev[1];
  motion[0] = c * dx + s * dy;
  motion[1] = -s * dx + c * dy;
  motion[2] = oracle_sub_angle(cur[2], prev[2]);
}
// Pose2D::calPredPose (src/Pose2D.cpp:28-37)
void oracle_cal_pred_pose(const double motion[3], const double last[3], double pred[3]) {
  double a = deg2rad(last[2]);
  double c = std::cos(a), s = std::sin(a);
  pred[0] = c * motion[0] + (-s) * motion[1] + last[0];
  pred[1] = s * motion[0] + c * motion[1] + last[1];
  pred[2] = oracle_add_angle(last[2], motion[2]);
}

static void mat3_mul(const double A[9], const double B[9], double C[9]) {
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { double s = 0; for (int k = 0; k < 3; ++k) s += A[i * 3 + k] * B[k * 3 + j]; C[i * 3 + j] = s; }
}
static void mat3_inv(const double m[9], double o[9]) {
  double c00 = m[4] * m[8] - m[5] * m[7], c01 = m[5] * m[6] - m[3] * m[8], c02 = m[3] * m[7] - m[4] * m[6];
  double det = m[0] * c00 + m[1] * c01 + m[2] * c02;
  double id = 1.0 / det;
  o[0] = c00 * id; o[1] = (m[2] * m[7] - m[1] * m[8]) * id; o[2] = (m[1] * m[5] - m[2] * m[4]) * id;
  o[3] = c01 * id; o[4] = (m[0] * m[8] - m[2] * m[6]) * id; o[5] = (m[2] * m[3] - m[0] * m[5]) * id;
  o[6] = c02 * id; o[7] = (m[1] * m[6] - m[0] * m[7]) * id; o[8] = (m[0] * m[4] - m[1] * m[3]) * id;
}

// PoseFuser::calOdometryCovariance (src/PoseFuser.cpp:38-61)
void oracle_odometry_cov(const double motion[3], const double last[3], const double lastCov[9], double delTime,
                         double coeVel, double coeOmega, double cov[9]) {
  double v = std::sqrt(motion[0] * motion[0] + motion[1] * motion[1]) / delTime;
  double omega = deg2rad(motion[2] / delTime);
  double M0 = coeVel * v * v, M1 = coeOmega * omega * omega;
  double th = deg2rad(last[2]);
  double A[3][2] = {{delTime * std::cos(th), 0}, {delTime * std::sin(th), 0}, {0, delTime}};
  double F[9] = {1, 0, -v * delTime * std::sin(th), 0, 1, v * delTime * std::cos(th), 0, 0, 1};
  double Ft[9]; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Ft[i * 3 + j] = F[j * 3 + i];
  double T1[9], T2[9];
  mat3_mul(F, lastCov, T1); mat3_mul(T1, Ft, T2);
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j)
    cov[i * 3 + j] = T2[i * 3 + j] + (A[i][0] * M0 * A[j][0] + A[i][1] * M1 * A[j][1]);
}
// PoseFuser::fusePose (src/PoseFuser.cpp:3-36)
void oracle_fuse_pose(const double pred[3], const double est[3], const double motion[3], const double last[3],
                      const double lastCov[9], const double Q[9], double delTime, double coeVel, double coeOmega,
                      double fused[3], double cov[9]) {
  double ch[9]; oracle_odometry_cov(motion, last, lastCov, delTime, coeVel, coeOmega, ch);
  double S[9], Si[9], K[9];
  for (int k = 0; k < 9; ++k) S[k] = Q[k] + ch[k];
  mat3_inv(S, Si); mat3_mul(ch, Si, K);
  double IK[9]; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) IK[i * 3 + j] = (i == j ? 1.0 : 0.0) - K[i * 3 + j];
  mat3_mul(IK, ch, cov);
  double zh[3] = {est[0] - pred[0], est[1] - pred[1], deg2rad(oracle_sub_angle(est[2], pred[2]))};
  double mh[3] = {pred[0], pred[1], deg2rad(pred[2])};
  double mu[3];
  for (int i = 0; i < 3; ++i) mu[i] = K[i * 3] * zh[0] + K[i * 3 + 1] * zh[1] + K[i * 3 + 2] * zh[2] + mh[i];
  fused[0] = mu[0]; fused[1] = mu[1]; fused[2] = rad2deg(mu[2]);
}

}  // extern "C"
