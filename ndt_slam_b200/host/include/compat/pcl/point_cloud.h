// compat/pcl/point_cloud.h -- the two PCL value types that cross the reference's class interfaces
// (pcl::PointXYZ, pcl::PointCloud<pcl::PointXYZ>::Ptr: PointCloudMap.h:76-77, PoseEstimator.h:91).
// Memory layout of PointXYZ is the 16-byte {x, y, z, pad} PCL uses, so a cloud's points go to the
// C ABI (float4-strided) without conversion. With a real PCL on the include path this is not used.
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

namespace pcl {
struct alignas(16) PointXYZ {
  float x = 0.f, y = 0.f, z = 0.f, pad = 1.f;
  PointXYZ() {}
  PointXYZ(float a, float b, float c) : x(a), y(b), z(c) {}
};
static_assert(sizeof(PointXYZ) == 16, "PointXYZ must be float4-sized");
struct PCLHeader { uint32_t seq = 0; uint64_t stamp = 0; std::string frame_id; };

template <class P> class PointCloud {
 public:
  typedef std::shared_ptr<PointCloud<P>> Ptr;
  typedef std::shared_ptr<const PointCloud<P>> ConstPtr;
  PCLHeader header;
  std::vector<P> points;
  uint32_t width = 0, height = 0;
  bool is_dense = true;
  std::size_t size() const { return points.size(); }
  void clear() { points.clear(); width = height = 0; }
  void push_back(const P &p) { points.push_back(p); width = static_cast<uint32_t>(points.size()); height = 1; }
  PointCloud &operator+=(const PointCloud &o) {
    points.insert(points.end(), o.points.begin(), o.points.end());
    width = static_cast<uint32_t>(points.size());
    height = 1;
    is_dense = is_dense && o.is_dense;
    return *this;
  }
};
}  // namespace pcl
namespace boost { using std::make_shared; }   // reference-style call sites: boost::make_shared<pcl::PointCloud<...>>()
