#pragma once
#include <limits>
#include <string>
#include <pcl/point_cloud.h>
namespace pcl {
template <class PointT> class PassThrough {
  typename PointCloud<PointT>::ConstPtr input_;
  std::string field_ = "x";
  float lo_ = -std::numeric_limits<float>::max(), hi_ = std::numeric_limits<float>::max();
 public:
  void setInputCloud(const typename PointCloud<PointT>::ConstPtr &c) { input_ = c; }
  void setFilterFieldName(const std::string &f) { field_ = f; }
  void setFilterLimits(float a, float b) { lo_ = a; hi_ = b; }
  void filter(PointCloud<PointT> &out) {
    out.clear();
    for (const auto &p : input_->points) {
      const float v = field_ == "x" ? p.x : (field_ == "y" ? p.y : p.z);
      if (v >= lo_ && v <= hi_) out.push_back(p);
    }
  }
};
}  // namespace pcl
