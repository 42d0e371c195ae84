// pcl::NormalDistributionsTransform<PointSource, PointTarget> restated in its full 6-DoF form with
// Eigen (algorithm: Magnusson 2009 eq. 6.9 - 6.21, More & Thuente 1994).
// TEST INFRASTRUCTURE: together with the reference's own sources this is oracle/_ref. It is an
// independent second statement of the algorithm (generic 6 x 6 loops, JacobiSVD, SelfAdjointEigenSolver)
// against which the plain z = 0 oracle and the CUDA path are checked.
#pragma once
#include <cmath>
#include <limits>
#include <vector>
#include <Eigen/Dense>
#include <Eigen/Geometry>
#include <Eigen/SVD>
#include <pcl/point_cloud.h>
#include <pcl/filters/voxel_grid_covariance.h>

#ifndef MINIPCL_PROJECT_TRANSFORM
// 1: rotation entries are (float)cos((double)(float)yaw), the definition shared with the CUDA path
//    (SURVEY.md App. A.6). 0: Eigen's float AngleAxis (cosf / sinf), as a stock PCL build would do.
#define MINIPCL_PROJECT_TRANSFORM 1
#endif
#ifndef MINIPCL_MT_INTERVAL_LT0
#define MINIPCL_MT_INTERVAL_LT0 1
#endif

namespace pcl {

template <class PointT>
inline void transformPointCloud(const PointCloud<PointT> &in, PointCloud<PointT> &out, const Eigen::Matrix4f &T) {
  if (&in != &out) { out.header = in.header; out.is_dense = in.is_dense; out.width = in.width; out.height = in.height; out.points.resize(in.points.size()); }
  for (std::size_t i = 0; i < in.points.size(); ++i) {
    const float x = in.points[i].x, y = in.points[i].y, z = in.points[i].z;
    PointT p = in.points[i];
    // ((m0 x + m1 y) + m2 z) + t, float, no fused multiply-add
    p.x = ((T(0, 0) * x + T(0, 1) * y) + T(0, 2) * z) + T(0, 3);
    p.y = ((T(1, 0) * x + T(1, 1) * y) + T(1, 2) * z) + T(1, 3);
    p.z = ((T(2, 0) * x + T(2, 1) * y) + T(2, 2) * z) + T(2, 3);
    out.points[i] = p;
  }
}

template <class PointSource, class PointTarget> class NormalDistributionsTransform {
 public:
  typedef PointCloud<PointSource> PointCloudSource;
  typedef typename PointCloudSource::Ptr PointCloudSourcePtr;
  typedef typename PointCloudSource::ConstPtr PointCloudSourceConstPtr;
  typedef PointCloud<PointTarget> PointCloudTarget;
  typedef typename PointCloudTarget::ConstPtr PointCloudTargetConstPtr;
  typedef VoxelGridCovariance<PointTarget> TargetGrid;
  typedef typename TargetGrid::LeafConstPtr TargetGridLeafConstPtr;
  typedef Eigen::Matrix<double, 6, 1> Vector6d;
  typedef Eigen::Matrix<double, 6, 6> Matrix6d;

  NormalDistributionsTransform() {
    // PCL defaults
    resolution_ = 1.0f; step_size_ = 0.1; outlier_ratio_ = 0.55; transformation_epsilon_ = 0.1; max_iterations_ = 35;
    final_transformation_.setIdentity(); transformation_.setIdentity(); previous_transformation_.setIdentity();
    point_gradient_.setZero(); point_gradient_.block<3, 3>(0, 0).setIdentity(); point_hessian_.setZero();
  }
  virtual ~NormalDistributionsTransform() {}

  void setTransformationEpsilon(double e) { transformation_epsilon_ = e; }
  void setStepSize(double s) { step_size_ = s; }
  void setResolution(float r) { if (resolution_ != r) { resolution_ = r; if (target_) init(); } }
  void setMaximumIterations(int n) { max_iterations_ = n; }
  void setOulierRatio(double o) { outlier_ratio_ = o; }
  void setInputSource(const PointCloudSourceConstPtr &c) { input_ = c; }
  void setInputTarget(const PointCloudTargetConstPtr &c) { target_ = c; init(); }
  Eigen::Matrix4f getFinalTransformation() const { return final_transformation_; }
  bool hasConverged() const { return converged_; }
  double getTransformationProbability() const { return trans_probability_; }
  int getFinalNumIteration() const { return nr_iterations_; }
  const TargetGrid &getTargetCells() const { return target_cells_; }
  int objectivePasses() const { return passes_; }

  void align(PointCloudSource &output, const Eigen::Matrix4f &guess = Eigen::Matrix4f::Identity()) {
    output.points.resize(input_->points.size());
    output.header = input_->header; output.width = (uint32_t)input_->points.size(); output.height = 1; output.is_dense = input_->is_dense;
    for (std::size_t i = 0; i < input_->points.size(); ++i) output.points[i] = input_->points[i];
    converged_ = false;
    final_transformation_ = transformation_ = previous_transformation_ = Eigen::Matrix4f::Identity();
    computeTransformation(output, guess);
  }

  // Registration::getFitnessScore(max_range = DBL_MAX): mean float squared distance to the nearest target point
  double getFitnessScore(double max_range = std::numeric_limits<double>::max()) {
    double fitness = 0.0;
    PointCloudSource tr;
    transformPointCloud(*input_, tr, final_transformation_);
    int nr = 0;
    for (const auto &p : tr.points) {
      float best = std::numeric_limits<float>::max();
      for (const auto &t : target_->points) {
        if (!std::isfinite(t.x) || !std::isfinite(t.y) || !std::isfinite(t.z)) continue;
        const float dx = p.x - t.x, dy = p.y - t.y, dz = p.z - t.z;
        float d = dx * dx; d += dy * dy; d += dz * dz;
        if (d < best) best = d;
      }
      if (best <= max_range) { fitness += best; ++nr; }
    }
    return nr > 0 ? fitness / nr : std::numeric_limits<double>::max();
  }

  // parity hooks (not part of PCL's public API)
  double evaluate(const Vector6d &p, bool compute_hessian, Vector6d &g, Matrix6d &H) {
    gaussConstants();
    PointCloudSource tr;
    transformPointCloud(*input_, tr, matrixOf(p));
    return computeDerivatives(g, H, tr, const_cast<Vector6d &>(p), compute_hessian);
  }

 protected:
  void init() { target_cells_.setLeafSize(resolution_, resolution_, resolution_); target_cells_.setInputCloud(target_); target_cells_.filter(true); }

  void gaussConstants() {
    const double gauss_c1 = 10 * (1 - outlier_ratio_);
    const double gauss_c2 = outlier_ratio_ / std::pow(resolution_, 3);
    const double gauss_d3 = -std::log(gauss_c2);
    gauss_d1_ = -std::log(gauss_c1 + gauss_c2) - gauss_d3;
    gauss_d2_ = -2 * std::log((-std::log(gauss_c1 * std::exp(-0.5) + gauss_c2) - gauss_d3) / gauss_d1_);
  }

  static Eigen::Matrix4f matrixOf(const Vector6d &x) {
#if MINIPCL_PROJECT_TRANSFORM
    // z = 0 / pure-yaw use only: roll and pitch must be exactly zero on this path
    const float yaw = static_cast<float>(x(5));
    const float c = static_cast<float>(std::cos(static_cast<double>(yaw))), s = static_cast<float>(std::sin(static_cast<double>(yaw)));
    if (x(3) == 0.0 && x(4) == 0.0) {
      Eigen::Matrix4f T = Eigen::Matrix4f::Identity();
      T(0, 0) = c; T(0, 1) = -s; T(1, 0) = s; T(1, 1) = c;
      T(0, 3) = static_cast<float>(x(0)); T(1, 3) = static_cast<float>(x(1)); T(2, 3) = static_cast<float>(x(2));
      return T;
    }
#endif
    return (Eigen::Translation<float, 3>(static_cast<float>(x(0)), static_cast<float>(x(1)), static_cast<float>(x(2))) *
            Eigen::AngleAxis<float>(static_cast<float>(x(3)), Eigen::Vector3f::UnitX()) *
            Eigen::AngleAxis<float>(static_cast<float>(x(4)), Eigen::Vector3f::UnitY()) *
            Eigen::AngleAxis<float>(static_cast<float>(x(5)), Eigen::Vector3f::UnitZ())).matrix();
  }

  virtual void computeTransformation(PointCloudSource &output, const Eigen::Matrix4f &guess) {
    nr_iterations_ = 0; converged_ = false; passes_ = 0;
    gaussConstants();
    if (guess != Eigen::Matrix4f::Identity()) {
      final_transformation_ = guess;
      transformPointCloud(output, output, guess);
    }
    point_gradient_.setZero();
    point_gradient_.block<3, 3>(0, 0).setIdentity();
    point_hessian_.setZero();
    Eigen::Transform<float, 3, Eigen::Affine, Eigen::ColMajor> eig_transformation;
    eig_transformation.matrix() = final_transformation_;
    Vector6d p, delta_p, score_gradient;
    const Eigen::Vector3f init_translation = eig_transformation.translation();
    const Eigen::Vector3f init_rotation = eig_transformation.rotation().eulerAngles(0, 1, 2);
    p << init_translation(0), init_translation(1), init_translation(2), init_rotation(0), init_rotation(1), init_rotation(2);
    Matrix6d hessian;
    double score = computeDerivatives(score_gradient, hessian, output, p);
    while (!converged_) {
      previous_transformation_ = transformation_;
      Eigen::JacobiSVD<Matrix6d> sv(hessian, Eigen::ComputeFullU | Eigen::ComputeFullV);
      delta_p = sv.solve(-score_gradient);
      double delta_p_norm = delta_p.norm();
      if (delta_p_norm == 0 || delta_p_norm != delta_p_norm) {
        trans_probability_ = score / static_cast<double>(input_->points.size());
        converged_ = delta_p_norm == delta_p_norm;
        final_p_ = p;
        return;
      }
      delta_p.normalize();
      delta_p_norm = computeStepLengthMT(p, delta_p, delta_p_norm, step_size_, transformation_epsilon_ / 2, score, score_gradient, hessian, output);
      delta_p *= delta_p_norm;
      transformation_ = matrixOf(delta_p);
      p = p + delta_p;
      if (nr_iterations_ > max_iterations_ || (nr_iterations_ && (std::fabs(delta_p_norm) < transformation_epsilon_))) converged_ = true;
      nr_iterations_++;
    }
    trans_probability_ = score / static_cast<double>(input_->points.size());
    final_p_ = p; final_score_ = score; final_hessian_ = hessian;
  }

  void computeAngleDerivatives(Vector6d &p, bool compute_hessian = true) {
    double cx, cy, cz, sx, sy, sz;
    if (std::fabs(p(3)) < 10e-5) { cx = 1.0; sx = 0.0; } else { cx = std::cos(p(3)); sx = std::sin(p(3)); }
    if (std::fabs(p(4)) < 10e-5) { cy = 1.0; sy = 0.0; } else { cy = std::cos(p(4)); sy = std::sin(p(4)); }
    if (std::fabs(p(5)) < 10e-5) { cz = 1.0; sz = 0.0; } else { cz = std::cos(p(5)); sz = std::sin(p(5)); }
    // eq. 6.19 [Magnusson 2009]
    j_ang_[0] << (-sx * sz + cx * sy * cz), (-sx * cz - cx * sy * sz), (-cx * cy);
    j_ang_[1] << (cx * sz + sx * sy * cz), (cx * cz - sx * sy * sz), (-sx * cy);
    j_ang_[2] << (-sy * cz), sy * sz, cy;
    j_ang_[3] << sx * cy * cz, (-sx * cy * sz), sx * sy;
    j_ang_[4] << (-cx * cy * cz), cx * cy * sz, (-cx * sy);
    j_ang_[5] << (-cy * sz), (-cy * cz), 0;
    j_ang_[6] << (cx * cz - sx * sy * sz), (-cx * sz - sx * sy * cz), 0;
    j_ang_[7] << (sx * cz + cx * sy * sz), (cx * sy * cz - sx * sz), 0;
    if (compute_hessian) {
      // eq. 6.21 [Magnusson 2009]: a2 a3 b2 b3 c2 c3 d1 d2 d3 e1 e2 e3 f1 f2 f3
      h_ang_[0] << (-cx * sz - sx * sy * cz), (-cx * cz + sx * sy * sz), sx * cy;
      h_ang_[1] << (-sx * sz + cx * sy * cz), (-cx * sy * sz - sx * cz), (-cx * cy);
      h_ang_[2] << (cx * cy * cz), (-cx * cy * sz), (cx * sy);
      h_ang_[3] << (sx * cy * cz), (-sx * cy * sz), (sx * sy);
      h_ang_[4] << (-sx * cz - cx * sy * sz), (sx * sz - cx * sy * cz), 0;
      h_ang_[5] << (cx * cz - sx * sy * sz), (-sx * sy * cz - cx * sz), 0;
      h_ang_[6] << (-cy * cz), (cy * sz), (sy);
      h_ang_[7] << (-sx * sy * cz), (sx * sy * sz), (sx * cy);
      h_ang_[8] << (cx * sy * cz), (-cx * sy * sz), (-cx * cy);
      h_ang_[9] << (sy * sz), (sy * cz), 0;
      h_ang_[10] << (-sx * cy * sz), (-sx * cy * cz), 0;
      h_ang_[11] << (cx * cy * sz), (cx * cy * cz), 0;
      h_ang_[12] << (-cy * cz), (cy * sz), 0;
      h_ang_[13] << (-cx * sz - sx * sy * cz), (-cx * cz + sx * sy * sz), 0;
      h_ang_[14] << (-sx * sz + cx * sy * cz), (-cx * sy * sz - sx * cz), 0;
    }
  }

  void computePointDerivatives(Eigen::Vector3d &x, bool compute_hessian = true) {
    point_gradient_(1, 3) = x.dot(j_ang_[0]);
    point_gradient_(2, 3) = x.dot(j_ang_[1]);
    point_gradient_(0, 4) = x.dot(j_ang_[2]);
    point_gradient_(1, 4) = x.dot(j_ang_[3]);
    point_gradient_(2, 4) = x.dot(j_ang_[4]);
    point_gradient_(0, 5) = x.dot(j_ang_[5]);
    point_gradient_(1, 5) = x.dot(j_ang_[6]);
    point_gradient_(2, 5) = x.dot(j_ang_[7]);
    if (compute_hessian) {
      Eigen::Vector3d a, b, c, d, e, f;
      a << 0, x.dot(h_ang_[0]), x.dot(h_ang_[1]);
      b << 0, x.dot(h_ang_[2]), x.dot(h_ang_[3]);
      c << 0, x.dot(h_ang_[4]), x.dot(h_ang_[5]);
      d << x.dot(h_ang_[6]), x.dot(h_ang_[7]), x.dot(h_ang_[8]);
      e << x.dot(h_ang_[9]), x.dot(h_ang_[10]), x.dot(h_ang_[11]);
      f << x.dot(h_ang_[12]), x.dot(h_ang_[13]), x.dot(h_ang_[14]);
      point_hessian_.block<3, 1>(9, 3) = a;  point_hessian_.block<3, 1>(12, 3) = b; point_hessian_.block<3, 1>(15, 3) = c;
      point_hessian_.block<3, 1>(9, 4) = b;  point_hessian_.block<3, 1>(12, 4) = d; point_hessian_.block<3, 1>(15, 4) = e;
      point_hessian_.block<3, 1>(9, 5) = c;  point_hessian_.block<3, 1>(12, 5) = e; point_hessian_.block<3, 1>(15, 5) = f;
    }
  }

  double updateDerivatives(Vector6d &score_gradient, Matrix6d &hessian, Eigen::Vector3d &x_trans, Eigen::Matrix3d &c_inv, bool compute_hessian = true) {
    Eigen::Vector3d cov_dxd_pi;
    double e_x_cov_x = std::exp(-gauss_d2_ * x_trans.dot(c_inv * x_trans) / 2);
    const double score_inc = -gauss_d1_ * e_x_cov_x;
    e_x_cov_x = gauss_d2_ * e_x_cov_x;
    if (e_x_cov_x > 1 || e_x_cov_x < 0 || e_x_cov_x != e_x_cov_x) return 0;
    e_x_cov_x *= gauss_d1_;
    for (int i = 0; i < 6; i++) {
      cov_dxd_pi = c_inv * point_gradient_.col(i);
      score_gradient(i) += x_trans.dot(cov_dxd_pi) * e_x_cov_x;
      if (compute_hessian) {
        for (int j = 0; j < hessian.cols(); j++) {
          hessian(i, j) += e_x_cov_x * (-gauss_d2_ * x_trans.dot(cov_dxd_pi) * x_trans.dot(c_inv * point_gradient_.col(j)) +
                                        x_trans.dot(c_inv * point_hessian_.block<3, 1>(3 * i, j)) +
                                        point_gradient_.col(j).dot(cov_dxd_pi));
        }
      }
    }
    return score_inc;
  }

  double computeDerivatives(Vector6d &score_gradient, Matrix6d &hessian, PointCloudSource &trans_cloud, Vector6d &p, bool compute_hessian = true) {
    ++passes_;
    score_gradient.setZero();
    hessian.setZero();
    double score = 0;
    computeAngleDerivatives(p);
    for (std::size_t idx = 0; idx < input_->points.size(); idx++) {
      const PointSource x_trans_pt = trans_cloud.points[idx];
      std::vector<TargetGridLeafConstPtr> neighborhood;
      std::vector<float> distances;
      target_cells_.radiusSearch(x_trans_pt, resolution_, neighborhood, distances);
      for (TargetGridLeafConstPtr cell : neighborhood) {
        const PointSource x_pt = input_->points[idx];
        Eigen::Vector3d x(x_pt.x, x_pt.y, x_pt.z);
        Eigen::Vector3d x_trans(x_trans_pt.x, x_trans_pt.y, x_trans_pt.z);
        x_trans -= cell->getMean();
        Eigen::Matrix3d c_inv = cell->getInverseCov();
        computePointDerivatives(x, compute_hessian);
        score += updateDerivatives(score_gradient, hessian, x_trans, c_inv, compute_hessian);
      }
    }
    return score;
  }

  void updateHessian(Matrix6d &hessian, Eigen::Vector3d &x_trans, Eigen::Matrix3d &c_inv) {
    Eigen::Vector3d cov_dxd_pi;
    double e_x_cov_x = gauss_d2_ * std::exp(-gauss_d2_ * x_trans.dot(c_inv * x_trans) / 2);
    if (e_x_cov_x > 1 || e_x_cov_x < 0 || e_x_cov_x != e_x_cov_x) return;
    e_x_cov_x *= gauss_d1_;
    for (int i = 0; i < 6; i++) {
      cov_dxd_pi = c_inv * point_gradient_.col(i);
      for (int j = 0; j < hessian.cols(); j++) {
        hessian(i, j) += e_x_cov_x * (-gauss_d2_ * x_trans.dot(cov_dxd_pi) * x_trans.dot(c_inv * point_gradient_.col(j)) +
                                      x_trans.dot(c_inv * point_hessian_.block<3, 1>(3 * i, j)) +
                                      point_gradient_.col(j).dot(cov_dxd_pi));
      }
    }
  }

  // Hessian only; the angle terms are NOT recomputed and p is ignored (PCL behaviour)
  void computeHessian(Matrix6d &hessian, PointCloudSource &trans_cloud, Vector6d &) {
    ++passes_;
    hessian.setZero();
    for (std::size_t idx = 0; idx < input_->points.size(); idx++) {
      const PointSource x_trans_pt = trans_cloud.points[idx];
      std::vector<TargetGridLeafConstPtr> neighborhood;
      std::vector<float> distances;
      target_cells_.radiusSearch(x_trans_pt, resolution_, neighborhood, distances);
      for (TargetGridLeafConstPtr cell : neighborhood) {
        const PointSource x_pt = input_->points[idx];
        Eigen::Vector3d x(x_pt.x, x_pt.y, x_pt.z);
        Eigen::Vector3d x_trans(x_trans_pt.x, x_trans_pt.y, x_trans_pt.z);
        x_trans -= cell->getMean();
        Eigen::Matrix3d c_inv = cell->getInverseCov();
        computePointDerivatives(x);
        updateHessian(hessian, x_trans, c_inv);
      }
    }
  }

  static double psiMT(double a, double f_a, double f_0, double g_0, double mu) { return f_a - f_0 - mu * g_0 * a; }
  static double dPsiMT(double g_a, double g_0, double mu) { return g_a - mu * g_0; }

  static double trialValueSelectionMT(double a_l, double f_l, double g_l, double a_u, double f_u, double g_u, double a_t, double f_t, double g_t) {
    if (f_t > f_l) {
      const double z = 3 * (f_t - f_l) / (a_t - a_l) - g_t - g_l;
      const double w = std::sqrt(z * z - g_t * g_l);
      const double a_c = a_l + (a_t - a_l) * (w - g_l - z) / (g_t - g_l + 2 * w);
      const double a_q = a_l - 0.5 * (a_l - a_t) * g_l / (g_l - (f_l - f_t) / (a_l - a_t));
      return std::fabs(a_c - a_l) < std::fabs(a_q - a_l) ? a_c : 0.5 * (a_q + a_c);
    }
    if (g_t * g_l < 0) {
      const double z = 3 * (f_t - f_l) / (a_t - a_l) - g_t - g_l;
      const double w = std::sqrt(z * z - g_t * g_l);
      const double a_c = a_l + (a_t - a_l) * (w - g_l - z) / (g_t - g_l + 2 * w);
      const double a_s = a_l - (a_l - a_t) / (g_l - g_t) * g_l;
      return std::fabs(a_c - a_t) >= std::fabs(a_s - a_t) ? a_c : a_s;
    }
    if (std::fabs(g_t) <= std::fabs(g_l)) {
      const double z = 3 * (f_t - f_l) / (a_t - a_l) - g_t - g_l;
      const double w = std::sqrt(z * z - g_t * g_l);
      const double a_c = a_l + (a_t - a_l) * (w - g_l - z) / (g_t - g_l + 2 * w);
      const double a_s = a_l - (a_l - a_t) / (g_l - g_t) * g_l;
      const double a_t_next = std::fabs(a_c - a_t) < std::fabs(a_s - a_t) ? a_c : a_s;
      return a_t > a_l ? std::min(a_t + 0.66 * (a_u - a_t), a_t_next) : std::max(a_t + 0.66 * (a_u - a_t), a_t_next);
    }
    const double z = 3 * (f_t - f_u) / (a_t - a_u) - g_t - g_u;
    const double w = std::sqrt(z * z - g_t * g_u);
    return a_u + (a_t - a_u) * (w - g_u - z) / (g_t - g_u + 2 * w);
  }

  static bool updateIntervalMT(double &a_l, double &f_l, double &g_l, double &a_u, double &f_u, double &g_u, double a_t, double f_t, double g_t) {
    if (f_t > f_l) { a_u = a_t; f_u = f_t; g_u = g_t; return false; }
    if (g_t * (a_l - a_t) > 0) { a_l = a_t; f_l = f_t; g_l = g_t; return false; }
    if (g_t * (a_l - a_t) < 0) { a_u = a_l; f_u = f_l; g_u = g_l; a_l = a_t; f_l = f_t; g_l = g_t; return false; }
    return true;
  }

  double computeStepLengthMT(const Vector6d &x, Vector6d &step_dir, double step_init, double step_max, double step_min, double &score,
                             Vector6d &score_gradient, Matrix6d &hessian, PointCloudSource &trans_cloud) {
    const double phi_0 = -score;
    double d_phi_0 = -(score_gradient.dot(step_dir));
    Vector6d x_t;
    if (d_phi_0 >= 0) {
      if (d_phi_0 == 0) return 0;
      d_phi_0 *= -1;
      step_dir *= -1;
    }
    const int max_step_iterations = 10;
    int step_iterations = 0;
    const double mu = 1.e-4, nu = 0.9;
    double a_l = 0, a_u = 0;
    double f_l = psiMT(a_l, phi_0, phi_0, d_phi_0, mu), g_l = dPsiMT(d_phi_0, d_phi_0, mu);
    double f_u = psiMT(a_u, phi_0, phi_0, d_phi_0, mu), g_u = dPsiMT(d_phi_0, d_phi_0, mu);
#if MINIPCL_MT_INTERVAL_LT0
    bool interval_converged = (step_max - step_min) < 0, open_interval = true;
#else
    bool interval_converged = (step_max - step_min) > 0, open_interval = true;
#endif
    double a_t = step_init;
    a_t = std::min(a_t, step_max);
    a_t = std::max(a_t, step_min);
    x_t = x + step_dir * a_t;
    final_transformation_ = matrixOf(x_t);
    transformPointCloud(*input_, trans_cloud, final_transformation_);
    score = computeDerivatives(score_gradient, hessian, trans_cloud, x_t, true);
    double phi_t = -score;
    double d_phi_t = -(score_gradient.dot(step_dir));
    double psi_t = psiMT(a_t, phi_t, phi_0, d_phi_0, mu);
    double d_psi_t = dPsiMT(d_phi_t, d_phi_0, mu);
    while (!interval_converged && step_iterations < max_step_iterations && !(psi_t <= 0 && d_phi_t <= -nu * d_phi_0)) {
      if (open_interval) a_t = trialValueSelectionMT(a_l, f_l, g_l, a_u, f_u, g_u, a_t, psi_t, d_psi_t);
      else a_t = trialValueSelectionMT(a_l, f_l, g_l, a_u, f_u, g_u, a_t, phi_t, d_phi_t);
      a_t = std::min(a_t, step_max);
      a_t = std::max(a_t, step_min);
      x_t = x + step_dir * a_t;
      final_transformation_ = matrixOf(x_t);
      transformPointCloud(*input_, trans_cloud, final_transformation_);
      score = computeDerivatives(score_gradient, hessian, trans_cloud, x_t, false);
      phi_t = -score;
      d_phi_t = -(score_gradient.dot(step_dir));
      psi_t = psiMT(a_t, phi_t, phi_0, d_phi_0, mu);
      d_psi_t = dPsiMT(d_phi_t, d_phi_0, mu);
      if (open_interval && (psi_t <= 0 && d_psi_t >= 0)) {
        open_interval = false;
        f_l = f_l + phi_0 - mu * d_phi_0 * a_l; g_l = g_l + mu * d_phi_0;
        f_u = f_u + phi_0 - mu * d_phi_0 * a_u; g_u = g_u + mu * d_phi_0;
      }
      if (open_interval) interval_converged = updateIntervalMT(a_l, f_l, g_l, a_u, f_u, g_u, a_t, psi_t, d_psi_t);
      else interval_converged = updateIntervalMT(a_l, f_l, g_l, a_u, f_u, g_u, a_t, phi_t, d_phi_t);
      step_iterations++;
    }
    if (step_iterations) computeHessian(hessian, trans_cloud, x_t);
    return a_t;
  }

  PointCloudSourceConstPtr input_;
  PointCloudTargetConstPtr target_;
  TargetGrid target_cells_;
  float resolution_;
  double step_size_, outlier_ratio_, gauss_d1_ = 0, gauss_d2_ = 0, trans_probability_ = 0, transformation_epsilon_;
  int max_iterations_, nr_iterations_ = 0, passes_ = 0;
  bool converged_ = false;
  Eigen::Matrix4f final_transformation_, transformation_, previous_transformation_;
  Eigen::Vector3d j_ang_[8], h_ang_[15];
  Eigen::Matrix<double, 3, 6> point_gradient_;
  Eigen::Matrix<double, 18, 6> point_hessian_;

 public:
  Vector6d final_p_ = Vector6d::Zero();
  double final_score_ = 0;
  Matrix6d final_hessian_ = Matrix6d::Zero();
  EIGEN_MAKE_ALIGNED_OPERATOR_NEW
};
}  // namespace pcl
