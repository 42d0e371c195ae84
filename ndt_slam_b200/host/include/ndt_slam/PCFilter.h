// PCFilter.h -- moving-object removal used by Submap::makeMap when removeMoving is set
// [REF include/ndt_slam/PCFilter.h:17-107]. This is map-side pre-filtering, not NDT arithmetic, and
// sits outside the accelerated path. The voxel difference here uses an absolute
// floor(p / resol) lattice; PCL's OctreePointCloudChangeDetector anchors its voxels at a data-dependent
// bounding box, so individual points near voxel faces can be classified differently (documented
// deviation; the headline configuration runs with removeMoving = false).
#ifndef NDT_SLAM_B200_PCFILTER_H_
#define NDT_SLAM_B200_PCFILTER_H_

#include <cmath>
#include <set>
#include <string>
#include <tuple>
#include <pcl/point_cloud.h>
#include <ros/ros.h>

class PCFilter {
 public:
  double resol;
  double thre_neighbor;

  PCFilter() : resol(0.05), thre_neighbor(0.1) {
    ros::param::get("resol", resol);
    ros::param::get("thre_neighbor", thre_neighbor);
  }

  // cloud_base without the points closer than thre_neighbor to any point of point_list
  pcl::PointCloud<pcl::PointXYZ>::Ptr remove_neighborPoint(pcl::PointCloud<pcl::PointXYZ>::Ptr cloud_base,
                                                           pcl::PointCloud<pcl::PointXYZ>::Ptr point_list) {
    auto kept = std::make_shared<pcl::PointCloud<pcl::PointXYZ>>();
    for (const auto &p : cloud_base->points) {
      bool near_any = false;
      for (const auto &q : point_list->points) {
        const float dx = p.x - q.x, dy = p.y - q.y, dz = p.z - q.z;
        if (std::sqrt(dx * dx + dy * dy + dz * dz) < thre_neighbor) near_any = true;
      }
      if (!near_any) kept->points.push_back(p);
    }
    kept->width = static_cast<uint32_t>(kept->points.size());
    kept->height = 1;
    return kept;
  }

  // points of cloud_test that fall into voxels cloud_base does not occupy
  pcl::PointCloud<pcl::PointXYZ>::Ptr difference_extraction(pcl::PointCloud<pcl::PointXYZ>::Ptr cloud_base,
                                                            pcl::PointCloud<pcl::PointXYZ>::Ptr cloud_test) {
    typedef std::tuple<long, long, long> Key;
    auto key = [this](const pcl::PointXYZ &p) {
      return Key((long)std::floor(p.x / resol), (long)std::floor(p.y / resol), (long)std::floor(p.z / resol));
    };
    std::set<Key> occupied;
    for (const auto &p : cloud_base->points) occupied.insert(key(p));
    auto diff = std::make_shared<pcl::PointCloud<pcl::PointXYZ>>();
    for (const auto &p : cloud_test->points)
      if (!occupied.count(key(p))) diff->points.push_back(p);
    diff->width = static_cast<uint32_t>(diff->points.size());
    diff->height = 1;
    return diff;
  }

  pcl::PointCloud<pcl::PointXYZ>::Ptr pass_through(std::string axis, double lo, double hi,
                                                   pcl::PointCloud<pcl::PointXYZ>::Ptr cloud) {
    auto out = std::make_shared<pcl::PointCloud<pcl::PointXYZ>>();
    for (const auto &p : cloud->points) {
      const float v = axis == "x" ? p.x : (axis == "y" ? p.y : p.z);
      if (v >= lo && v <= hi) out->push_back(p);
    }
    return out;
  }
};

#endif
