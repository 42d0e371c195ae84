// pcl::VoxelGridCovariance<PointT> restated in its full 3-D form with Eigen.
// Independent of oracle/ndt_oracle.cpp (which carries the exact z = 0 reduction); tests compare the two.
// The centroid kd-tree (FLANN) is replaced by a bucket lattice with identical radius-search
// semantics: every centroid with float ||x - c||^2 < r^2, sorted by distance.
#pragma once
#include <algorithm>
#include <cmath>
#include <limits>
#include <map>
#include <unordered_map>
#include <vector>
#include <Eigen/Dense>
#include <Eigen/Eigenvalues>
#include <pcl/point_cloud.h>

#ifndef MINIPCL_COV_INIT_IDENTITY
#define MINIPCL_COV_INIT_IDENTITY 1   // PCL <= 1.11: Leaf::cov_ starts at Identity (App. A.7)
#endif
#ifndef MINIPCL_COV_SCALE_NM1_N
#define MINIPCL_COV_SCALE_NM1_N 1     // biased single-pass covariance, then *= (n - 1) / n
#endif

namespace pcl {
template <class PointT> class VoxelGridCovariance {
 public:
  struct Leaf {
    int nr_points = 0;
    Eigen::Vector3d mean_ = Eigen::Vector3d::Zero();
    Eigen::Vector3f centroid = Eigen::Vector3f::Zero();
    Eigen::Matrix3d cov_ = MINIPCL_COV_INIT_IDENTITY ? Eigen::Matrix3d(Eigen::Matrix3d::Identity()) : Eigen::Matrix3d(Eigen::Matrix3d::Zero());
    Eigen::Matrix3d icov_ = Eigen::Matrix3d::Zero();
    Eigen::Matrix3d evecs_ = Eigen::Matrix3d::Identity();
    Eigen::Vector3d evals_ = Eigen::Vector3d::Zero();
    const Eigen::Vector3d &getMean() const { return mean_; }
    const Eigen::Matrix3d &getInverseCov() const { return icov_; }
    int getPointCount() const { return nr_points; }
    EIGEN_MAKE_ALIGNED_OPERATOR_NEW
  };
  typedef const Leaf *LeafConstPtr;

 private:
  typename PointCloud<PointT>::ConstPtr input_;
  float leaf_ = 1.f, inv_ = 1.f;
  int min_points_per_voxel_ = 6;
  double min_covar_eigvalue_mult_ = 0.01;
  Eigen::Vector3i min_b_ = Eigen::Vector3i::Zero(), max_b_ = Eigen::Vector3i::Zero(), div_b_ = Eigen::Vector3i::Zero();
  std::map<std::size_t, Leaf> leaves_;
  std::vector<Eigen::Vector3f> voxel_centroids_;
  std::vector<int> voxel_centroids_leaf_indices_;
  std::unordered_map<long long, std::vector<int>> bucket_;     // stand-in for the FLANN kd-tree
  static long long bkey(long a, long b, long c) { return ((a & 0x1fffffLL) << 42) | ((b & 0x1fffffLL) << 21) | (c & 0x1fffffLL); }

 public:
  void setLeafSize(float lx, float, float) { leaf_ = lx; inv_ = 1.0f / lx; }
  void setInputCloud(const typename PointCloud<PointT>::ConstPtr &c) { input_ = c; }
  const std::map<std::size_t, Leaf> &getLeaves() const { return leaves_; }
  Eigen::Vector3i getMinBoxCoordinates() const { return min_b_; }
  Eigen::Vector3i getNrDivisions() const { return div_b_; }

  void filter(bool /*searchable*/ = true) {
    leaves_.clear(); voxel_centroids_.clear(); voxel_centroids_leaf_indices_.clear(); bucket_.clear();
    div_b_.setZero();
    if (!input_) return;
    Eigen::Vector3f min_p = Eigen::Vector3f::Constant(std::numeric_limits<float>::max());
    Eigen::Vector3f max_p = Eigen::Vector3f::Constant(-std::numeric_limits<float>::max());
    bool any = false;
    for (const auto &p : input_->points) {
      if (!input_->is_dense && (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z))) continue;
      Eigen::Vector3f v(p.x, p.y, p.z);
      min_p = min_p.cwiseMin(v); max_p = max_p.cwiseMax(v); any = true;
    }
    if (!any) return;
    const std::int64_t dx = static_cast<std::int64_t>((max_p[0] - min_p[0]) * inv_) + 1;
    const std::int64_t dy = static_cast<std::int64_t>((max_p[1] - min_p[1]) * inv_) + 1;
    const std::int64_t dz = static_cast<std::int64_t>((max_p[2] - min_p[2]) * inv_) + 1;
    if (dx * dy * dz > static_cast<std::int64_t>(std::numeric_limits<std::int32_t>::max())) return;
    for (int a = 0; a < 3; ++a) {
      min_b_[a] = static_cast<int>(std::floor(min_p[a] * inv_));
      max_b_[a] = static_cast<int>(std::floor(max_p[a] * inv_));
    }
    div_b_ = max_b_ - min_b_ + Eigen::Vector3i::Ones();
    const Eigen::Vector3i mul(1, div_b_[0], div_b_[0] * div_b_[1]);
    // first pass
    for (const auto &p : input_->points) {
      if (!input_->is_dense && (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z))) continue;
      const int ijk0 = static_cast<int>(std::floor(p.x * inv_) - static_cast<float>(min_b_[0]));
      const int ijk1 = static_cast<int>(std::floor(p.y * inv_) - static_cast<float>(min_b_[1]));
      const int ijk2 = static_cast<int>(std::floor(p.z * inv_) - static_cast<float>(min_b_[2]));
      const int idx = ijk0 * mul[0] + ijk1 * mul[1] + ijk2 * mul[2];
      Leaf &leaf = leaves_[idx];
      const Eigen::Vector3d pt3d(p.x, p.y, p.z);
      leaf.mean_ += pt3d;
      leaf.cov_ += pt3d * pt3d.transpose();
      leaf.centroid += Eigen::Vector3f(p.x, p.y, p.z);
      ++leaf.nr_points;
    }
    // second pass
    Eigen::SelfAdjointEigenSolver<Eigen::Matrix3d> eigensolver;
    for (auto &kv : leaves_) {
      Leaf &leaf = kv.second;
      leaf.centroid /= static_cast<float>(leaf.nr_points);
      const Eigen::Vector3d pt_sum = leaf.mean_;
      leaf.mean_ /= leaf.nr_points;
      if (leaf.nr_points < min_points_per_voxel_) continue;
      voxel_centroids_.push_back(leaf.centroid);
      voxel_centroids_leaf_indices_.push_back(static_cast<int>(kv.first));
#if MINIPCL_COV_SCALE_NM1_N
      leaf.cov_ = (leaf.cov_ - 2 * (pt_sum * leaf.mean_.transpose())) / leaf.nr_points + leaf.mean_ * leaf.mean_.transpose();
      leaf.cov_ *= (leaf.nr_points - 1.0) / leaf.nr_points;
#else
      leaf.cov_ = (leaf.cov_ - pt_sum * leaf.mean_.transpose()) / (leaf.nr_points - 1.0);
#endif
      eigensolver.compute(leaf.cov_);
      Eigen::Matrix3d eigen_val = eigensolver.eigenvalues().asDiagonal();
      leaf.evecs_ = eigensolver.eigenvectors();
      if (eigen_val(0, 0) < 0 || eigen_val(1, 1) < 0 || eigen_val(2, 2) <= 0) { leaf.nr_points = -1; continue; }
      const double min_covar_eigvalue = min_covar_eigvalue_mult_ * eigen_val(2, 2);
      if (eigen_val(0, 0) < min_covar_eigvalue) {
        eigen_val(0, 0) = min_covar_eigvalue;
        if (eigen_val(1, 1) < min_covar_eigvalue) eigen_val(1, 1) = min_covar_eigvalue;
        leaf.cov_ = leaf.evecs_ * eigen_val * leaf.evecs_.inverse();
      }
      leaf.evals_ = eigen_val.diagonal();
      leaf.icov_ = leaf.cov_.inverse();
      if (leaf.icov_.maxCoeff() == std::numeric_limits<float>::infinity() ||
          leaf.icov_.minCoeff() == -std::numeric_limits<float>::infinity())
        leaf.nr_points = -1;
    }
    for (std::size_t i = 0; i < voxel_centroids_.size(); ++i) {
      const Eigen::Vector3f &c = voxel_centroids_[i];
      bucket_[bkey((long)std::floor(c[0] * inv_), (long)std::floor(c[1] * inv_), (long)std::floor(c[2] * inv_))].push_back((int)i);
    }
  }

  // all tree leaves with float squared distance < radius^2, nearest first
  int radiusSearch(const PointT &point, double radius, std::vector<LeafConstPtr> &k_leaves, std::vector<float> &k_sqr_distances) const {
    k_leaves.clear(); k_sqr_distances.clear();
    const float r2 = static_cast<float>(radius * radius);
    const int reach = static_cast<int>(std::ceil(radius * inv_)) + 1;
    const long cx = (long)std::floor(point.x * inv_), cy = (long)std::floor(point.y * inv_), cz = (long)std::floor(point.z * inv_);
    std::vector<std::pair<float, int>> found;
    for (long dz = -reach; dz <= reach; ++dz)
      for (long dy = -reach; dy <= reach; ++dy)
        for (long dx = -reach; dx <= reach; ++dx) {
          auto it = bucket_.find(bkey(cx + dx, cy + dy, cz + dz));
          if (it == bucket_.end()) continue;
          for (int i : it->second) {
            const Eigen::Vector3f &c = voxel_centroids_[i];
            const float ddx = point.x - c[0], ddy = point.y - c[1], ddz = point.z - c[2];
            float d = ddx * ddx; d += ddy * ddy; d += ddz * ddz;
            if (d < r2) found.emplace_back(d, i);
          }
        }
    std::stable_sort(found.begin(), found.end(), [](const std::pair<float, int> &a, const std::pair<float, int> &b) { return a.first < b.first; });
    for (auto &f : found) {
      k_leaves.push_back(&leaves_.at(voxel_centroids_leaf_indices_[f.second]));
      k_sqr_distances.push_back(f.first);
    }
    return (int)found.size();
  }
};
}  // namespace pcl
