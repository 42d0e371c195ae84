// VoxelFilter.h -- host-side pcl::ApproximateVoxelGrid<PointXYZ>::filter.
// The reference runs it on the source scan inside estimatePose [REF src/PoseEstimator.cpp:6-10] and on
// sub-map clouds [REF src/PointCloudMap.cpp:4-13]. It is a sequential, order-dependent 512-entry hash
// history, so single clouds are filtered on the host; batches use the device
// kernel behind ndt_approx_voxel_filter.
#ifndef NDT_SLAM_B200_VOXELFILTER_H_
#define NDT_SLAM_B200_VOXELFILTER_H_

#include <cmath>
#include <pcl/point_cloud.h>

namespace ndt_host {

inline void approximate_voxel_grid(const pcl::PointCloud<pcl::PointXYZ> &in, float leaf, pcl::PointCloud<pcl::PointXYZ> &out) {
  struct Slot { int ix = 0, iy = 0, iz = 0, n = 0; float sx = 0.f, sy = 0.f, sz = 0.f; };
  constexpr unsigned kSlots = 512;
  Slot table[kSlots];
  const float inv = 1.0f / leaf;
  out.points.clear();
  out.points.reserve(in.points.size());
  auto emit = [&out](Slot &s) {
    const float cnt = static_cast<float>(s.n);
    out.points.emplace_back(s.sx / cnt, s.sy / cnt, s.sz / cnt);
    s.n = 0; s.sx = s.sy = s.sz = 0.f;
  };
  for (const pcl::PointXYZ &p : in.points) {
    const int ix = static_cast<int>(std::floor(p.x * inv));
    const int iy = static_cast<int>(std::floor(p.y * inv));
    const int iz = static_cast<int>(std::floor(p.z * inv));
    const unsigned h = (static_cast<unsigned>(ix) * 7171u + static_cast<unsigned>(iy) * 3079u + static_cast<unsigned>(iz) * 4231u) & (kSlots - 1);
    Slot &s = table[h];
    if (s.n != 0 && (s.ix != ix || s.iy != iy || s.iz != iz)) emit(s);   // a different voxel claims the slot
    s.ix = ix; s.iy = iy; s.iz = iz;
    ++s.n;
    s.sx += p.x; s.sy += p.y; s.sz += p.z;
  }
  for (Slot &s : table) if (s.n != 0) emit(s);
  out.width = static_cast<uint32_t>(out.points.size());
  out.height = 1;
  out.is_dense = false;
}

}  // namespace ndt_host
#endif
