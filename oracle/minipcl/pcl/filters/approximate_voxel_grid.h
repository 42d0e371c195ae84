// pcl::ApproximateVoxelGrid<PointXYZ>::applyFilter restated: 512-entry hash
// history keyed on the float voxel index; a colliding different voxel flushes the entry.
#pragma once
#include <cmath>
#include <pcl/point_cloud.h>
namespace pcl {
template <class PointT> class ApproximateVoxelGrid {
  struct he { int ix, iy, iz, count; float c[3]; };
  typename PointCloud<PointT>::ConstPtr input_;
  float inv_[3] = {1.f, 1.f, 1.f};
  static const int histsize_ = 512;
 public:
  void setLeafSize(float lx, float ly, float lz) { inv_[0] = 1.0f / lx; inv_[1] = 1.0f / ly; inv_[2] = 1.0f / lz; }
  void setInputCloud(const typename PointCloud<PointT>::ConstPtr &c) { input_ = c; }
  void filter(PointCloud<PointT> &output) {
    std::vector<he> history(histsize_);
    for (auto &e : history) { e.ix = e.iy = e.iz = 0; e.count = 0; e.c[0] = e.c[1] = e.c[2] = 0.f; }
    output.points.resize(input_->points.size());
    std::size_t op = 0;
    auto flush = [&](he &e) {
      const float n = static_cast<float>(e.count);
      PointT p; p.x = e.c[0] / n; p.y = e.c[1] / n; p.z = e.c[2] / n;
      output.points[op++] = p;
    };
    for (std::size_t cp = 0; cp < input_->points.size(); ++cp) {
      const PointT &pt = input_->points[cp];
      const int ix = static_cast<int>(std::floor(pt.x * inv_[0]));
      const int iy = static_cast<int>(std::floor(pt.y * inv_[1]));
      const int iz = static_cast<int>(std::floor(pt.z * inv_[2]));
      const unsigned hash = static_cast<unsigned>((ix * 7171 + iy * 3079 + iz * 4231) & (histsize_ - 1));
      he &e = history[hash];
      if (e.count && (ix != e.ix || iy != e.iy || iz != e.iz)) { flush(e); e.count = 0; e.c[0] = e.c[1] = e.c[2] = 0.f; }
      e.ix = ix; e.iy = iy; e.iz = iz; e.count++;
      e.c[0] += pt.x; e.c[1] += pt.y; e.c[2] += pt.z;
    }
    for (int i = 0; i < histsize_; ++i) if (history[i].count) flush(history[i]);
    output.points.resize(op);
    output.width = static_cast<uint32_t>(op);
    output.height = 1;
    output.is_dense = false;
  }
};
}  // namespace pcl
