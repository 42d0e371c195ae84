// PointCloudMap.h -- trajectory, sub-maps and the local / global map clouds.
// Same public data and methods as the reference [REF include/ndt_slam/PointCloudMap.h:22-145,
// src/PointCloudMap.cpp:4-134]. localMap_cloud is the only coupling into the matcher
// (ScanMatcher.cpp:40): it becomes the NDT target, i.e. the input of the device grid build.
#ifndef NDT_SLAM_B200_POINTCLOUDMAP_H_
#define NDT_SLAM_B200_POINTCLOUDMAP_H_

#include <string>
#include <vector>
#include <pcl/point_cloud.h>
#include <ros/ros.h>

#include "LPoint2D.h"
#include "PCFilter.h"
#include "Pose2D.h"
#include "Scan2D.h"
#include "Timer.h"
#include "VoxelFilter.h"

class Submap {
 public:
  PCFilter pcf;

  double atdS;        // accumulated travel distance where this sub-map starts
  size_t cntS;        // first scan number
  size_t cntE;        // last scan number
  bool newest;

  bool removeMoving;

  pcl::PointCloud<pcl::PointXYZ>::Ptr p_cloud;
  std::vector<pcl::PointCloud<pcl::PointXYZ>::Ptr> scans;

  double LeafSize;
  Timer timer;

  Submap() : atdS(0), cntS(0), cntE(static_cast<size_t>(-1)), newest(true), removeMoving(false), LeafSize(0.2) { readParams(); }
  Submap(double a, size_t s) : atdS(a), cntS(s), cntE(static_cast<size_t>(-1)), newest(true), removeMoving(false), LeafSize(0.2) { readParams(); }

  void addPoints(pcl::PointCloud<pcl::PointXYZ>::Ptr cloud_ptr) { scans.emplace_back(cloud_ptr); }
  pcl::PointCloud<pcl::PointXYZ>::Ptr filterPoints();
  void makeMap();

  // Thinned view of p_cloud as (prefix, tail): `prefix` = centroids the filter has already emitted for good (it only
  // grows while p_cloud grows at its end), `tail` = the end-of-cloud flush (<= 512 points, changes every scan).
  // filterPoints() == prefix + tail. Lets makeLocalMap update localMap_cloud in O(new points).
  const std::vector<pcl::PointXYZ> &thinnedPrefix();
  void thinnedTail(std::vector<pcl::PointXYZ> &tail) const {
    thin.append_with_tail(p_cloud->points.data() + thin.consumed, p_cloud->points.size() - thin.consumed, tail);
  }

 private:
  // Incremental state (not in the reference, which rebuilds p_cloud and re-filters it from scratch every scan --
  // quadratic over a sub-map's life, and with moving-object removal every rebuild re-runs the octree difference and the
  // brute-force neighbour removal of every scan triple). A sub-map only ever appends whole scans, and with removeMoving the
  // filtered version of scan i+1 depends on scans i, i+1, i+2 only, so: p_cloud = [stable part that never changes again]
  // + [the newest scan, raw, replaced one scan later by its filtered version]. Both the concatenation and the
  // order-dependent voxel filter continue from the end of the stable part; the clouds produced are bit-identical to the
  // from-scratch ones.
  size_t appended_scans = 0;                 // scans[first .. appended_scans) are already in p_cloud
  const void *appended_into = nullptr;       // the p_cloud object they were appended to
  size_t appended_points = 0;
  size_t rm_triples = 0;                     // removeMoving: triples (i, i+1, i+2), i < rm_triples, are folded into p_cloud
  size_t stable_points = 0;                  // p_cloud[0 .. stable_points) is final
  ndt_host::IncrementalVoxelGrid thin;       // has consumed a prefix of the stable part
  const void *thin_of = nullptr;             // the p_cloud object `thin` has consumed a prefix of
  void syncThin();

  void readParams() {
    ros::param::get("removeMoving", removeMoving);
    ros::param::get("LeafSize", LeafSize);
    p_cloud = std::make_shared<pcl::PointCloud<pcl::PointXYZ>>();
  }
};

class PointCloudMap {
 public:
  std::vector<Pose2D> poses;
  Pose2D lastPose;
  Scan2D lastScan;
  int startFrame;

  pcl::PointCloud<pcl::PointXYZ>::Ptr globalMap_cloud;
  pcl::PointCloud<pcl::PointXYZ>::Ptr localMap_cloud;

  double sepThre;     // travel distance after which a new sub-map is started [m]
  double atd;         // accumulated travel distance
  std::vector<Submap> submaps;

  Timer timer;

  std::vector<pcl::PointCloud<pcl::PointXYZ>::Ptr> maps;
  std::string map_name, separated_map_name;

  // not in the reference: makeLocalMap number `localMapEpoch` left the first `localMapStablePrefix` points of
  // localMap_cloud exactly as call localMapEpoch - 1 had them (lets the matcher upload only the changed tail)
  uint64_t localMapEpoch = 0;
  size_t localMapStablePrefix = 0;
  // ... and its first `localMapSettled` points (previous sub-map + what the current sub-map's voxel filter has emitted for
  // good) stay a prefix of every later local map until the layout changes (a new sub-map), when localMapStablePrefix drops to 0
  size_t localMapSettled = 0;

  PointCloudMap() : startFrame(0), sepThre(30), atd(0) {
    ros::param::get("start_frame", startFrame);
    ros::param::get("sepThre", sepThre);
    ros::param::get("map_name", map_name);
    ros::param::get("separated_map_name", separated_map_name);
    globalMap_cloud = std::make_shared<pcl::PointCloud<pcl::PointXYZ>>();
    localMap_cloud = std::make_shared<pcl::PointCloud<pcl::PointXYZ>>();
    submaps.emplace_back(Submap());
  }

  void setLastPose(const Pose2D &p) { lastPose = p; }
  Pose2D getLastPose() const { return lastPose; }
  void setLastScan(const Scan2D &s) { lastScan = s; }
  std::vector<Submap> &getSubmaps() { return submaps; }

  void saveGlobalMap();     // PCD ASCII: the global map and one file per sub-map

  void addPose(const Pose2D &p);
  void addPoints(const std::vector<LPoint2D> &lps);
  void makeGlobalMap();
  void makeLocalMap();

 private:
  // makeLocalMap keeps localMap_cloud = [previous sub-map][current sub-map's thinned prefix][tail] and only rewrites
  // what changed since the last call (same content as rebuilding it from scratch)
  const void *lm_prev = nullptr, *lm_cur = nullptr;
  size_t lm_prev_points = 0, lm_prefix_points = 0, lm_fixed = 0;
  // makeGlobalMap keeps the frozen sub-maps' part of globalMap_cloud / maps
  const void *gm_cloud = nullptr;
  size_t gm_frozen = 0, gm_frozen_points = 0;
};

// PCD v0.7 ASCII writer for x y z float clouds (what pcl::io::savePCDFileASCII emits)
int savePCDFileASCII(const std::string &path, const pcl::PointCloud<pcl::PointXYZ> &cloud);

#endif
