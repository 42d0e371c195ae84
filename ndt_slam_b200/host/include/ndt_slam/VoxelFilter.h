// VoxelFilter.h -- host-side pcl::ApproximateVoxelGrid<PointXYZ>::filter.
// The reference runs it on the source scan inside estimatePose [REF src/PoseEstimator.cpp:6-10] and on
// sub-map clouds [REF src/PointCloudMap.cpp:4-13]. It is a sequential, order-dependent 512-entry hash
// history, so single clouds are filtered on the host; batches use the device
// kernel behind ndt_approx_voxel_filter.
#ifndef NDT_SLAM_B200_VOXELFILTER_H_
#define NDT_SLAM_B200_VOXELFILTER_H_

#include <cmath>
#include <vector>
#include <pcl/point_cloud.h>

namespace ndt_host {

// The filter consumes points strictly in order and a point only ever touches its own hash slot, so the state after
// a prefix of the cloud (the 512 slots + the centroids already emitted) is exactly the state a full run would have
// at that position. IncrementalVoxelGrid keeps that state: appending points to a cloud costs O(new points), and the
// complete output for the current cloud is `emitted` followed by a non-destructive flush of the live slots in slot
// order. The result is bit-identical to running the one-shot filter over the whole cloud.
class IncrementalVoxelGrid {
 public:
  struct Slot { int ix = 0, iy = 0, iz = 0, n = 0; float sx = 0.f, sy = 0.f, sz = 0.f; };
  static constexpr unsigned kSlots = 512;

  void reset(float leaf_size) {
    leaf = leaf_size;
    inv = 1.0f / leaf_size;
    for (Slot &s : table) s = Slot();
    emitted.clear();
    consumed = 0;
  }
  // consume points[consumed .. n) of a cloud that only ever grows at its end
  void feed(const pcl::PointXYZ *points, size_t n) {
    for (size_t i = consumed; i < n; ++i) step(table, points[i], emitted);
    if (n > consumed) consumed = n;
  }
  // centroids of the slots that are still live, in slot order (what the end-of-cloud flush would emit)
  template <class Vec>
  void append_live(Vec &out) const {
    for (const Slot &s : table) if (s.n != 0) out.push_back(centroid(s));
  }
  // What the filter would still emit if `tail[0 .. n)` followed the points consumed so far: the centroids flushed while
  // the tail is consumed, then the end-of-cloud flush -- computed on a copy of the 512 slots, so the state stays where it
  // is. Used when the end of the cloud is provisional (moving-object removal replaces the newest scan's raw points by a
  // filtered version one scan later).
  template <class Vec>
  void append_with_tail(const pcl::PointXYZ *tail, size_t n, Vec &out) const {
    if (n == 0) { append_live(out); return; }
    Slot scratch[kSlots];
    for (unsigned k = 0; k < kSlots; ++k) scratch[k] = table[k];
    for (size_t i = 0; i < n; ++i) step(scratch, tail[i], out);
    for (const Slot &s : scratch) if (s.n != 0) out.push_back(centroid(s));
  }
  float leaf = 0.f;
  size_t consumed = 0;
  std::vector<pcl::PointXYZ> emitted;

 private:
  static pcl::PointXYZ centroid(const Slot &s) {
    const float cnt = static_cast<float>(s.n);
    return pcl::PointXYZ(s.sx / cnt, s.sy / cnt, s.sz / cnt);
  }
  // one point of the hash history
  template <class Vec>
  void step(Slot *slots, const pcl::PointXYZ &p, Vec &flushed) const {
    const int ix = static_cast<int>(std::floor(p.x * inv));
    const int iy = static_cast<int>(std::floor(p.y * inv));
    const int iz = static_cast<int>(std::floor(p.z * inv));
    const unsigned h = (static_cast<unsigned>(ix) * 7171u + static_cast<unsigned>(iy) * 3079u + static_cast<unsigned>(iz) * 4231u) & (kSlots - 1);
    Slot &s = slots[h];
    if (s.n != 0 && (s.ix != ix || s.iy != iy || s.iz != iz)) {   // a different voxel claims the slot
      flushed.push_back(centroid(s));
      s.n = 0; s.sx = s.sy = s.sz = 0.f;
    }
    s.ix = ix; s.iy = iy; s.iz = iz;
    ++s.n;
    s.sx += p.x; s.sy += p.y; s.sz += p.z;
  }
  float inv = 0.f;
  Slot table[kSlots];
};

inline void approximate_voxel_grid(const pcl::PointCloud<pcl::PointXYZ> &in, float leaf, pcl::PointCloud<pcl::PointXYZ> &out) {
  IncrementalVoxelGrid f;
  f.reset(leaf);
  f.emitted.reserve(in.points.size());
  f.feed(in.points.data(), in.points.size());
  out.points.assign(f.emitted.begin(), f.emitted.end());
  f.append_live(out.points);
  out.width = static_cast<uint32_t>(out.points.size());
  out.height = 1;
  out.is_dense = false;
}

}  // namespace ndt_host
#endif
