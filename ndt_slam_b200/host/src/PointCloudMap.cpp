// PointCloudMap.cpp -- sub-map bookkeeping [REF src/PointCloudMap.cpp:4-134; PointCloudMap.h:124-136].
#include "ndt_slam/PointCloudMap.h"

#include <algorithm>
#include <cmath>
#include <fstream>
#include <iomanip>
#include <sstream>

#include "ndt_slam/VoxelFilter.h"

typedef pcl::PointCloud<pcl::PointXYZ> Cloud;

// Bring the incremental filter up to date with the stable part of p_cloud. Anything other than "the same cloud object
// grew at the end of its stable part" (a new cloud object, a shorter cloud, a different leaf size) restarts it.
void Submap::syncThin() {
  const float leaf = static_cast<float>(LeafSize);
  const size_t stable = std::min(stable_points, p_cloud->points.size());
  if (thin_of != p_cloud.get() || stable < thin.consumed || thin.leaf != leaf) {
    thin.reset(leaf);
    thin_of = p_cloud.get();
  }
  thin.feed(p_cloud->points.data(), stable);
}

const std::vector<pcl::PointXYZ> &Submap::thinnedPrefix() {
  syncThin();
  return thin.emitted;
}

Cloud::Ptr Submap::filterPoints() {
  Cloud::Ptr thinned = std::make_shared<Cloud>();
  syncThin();
  thinned->points.assign(thin.emitted.begin(), thin.emitted.end());
  thinnedTail(thinned->points);
  thinned->width = static_cast<uint32_t>(thinned->points.size());
  thinned->height = 1;
  thinned->is_dense = false;
  return thinned;
}

// Bring p_cloud up to date with the stored scans [REF src/PointCloudMap.cpp:15-39].
//  removeMoving: for every consecutive triple (i, i+1, i+2) keep scan i+1 minus the points that lie near
//                voxels seen only by i+1 and not by {i, i+2}; the first sub-map also keeps its first scan
//                and the newest sub-map its latest scan unfiltered (so a single scan appears twice, like the reference).
//  otherwise:    plain concatenation; sub-maps after the first skip the two scans carried over from their
//                predecessor.
// Only what the scans added since the last call change is recomputed (same cloud as clearing and rebuilding everything).
void Submap::makeMap() {
  const size_t n = scans.size();
  const size_t first = removeMoving ? 0 : (cntS == 0 ? 0 : 2);
  if (appended_into != p_cloud.get() || p_cloud->points.size() != appended_points || stable_points > appended_points ||
      (removeMoving ? rm_triples + 2 > std::max<size_t>(n, 2) : (appended_scans < first || appended_scans > n))) {
    p_cloud->clear();                              // someone else touched the cloud: start over
    appended_into = p_cloud.get();
    appended_scans = first;
    rm_triples = 0;
    if (removeMoving && cntS == 0 && n > 0) *p_cloud += *scans[0];
    stable_points = p_cloud->points.size();
  }
  if (removeMoving) {
    p_cloud->points.resize(stable_points);         // the previous newest scan was there raw: its filtered version follows
    for (size_t i = rm_triples; i + 2 < n; ++i) {
      Cloud::Ptr outer = std::make_shared<Cloud>();
      *outer += *scans[i];
      *outer += *scans[i + 2];
      Cloud::Ptr transient = pcf.difference_extraction(outer, scans[i + 1]);
      Cloud::Ptr kept = pcf.remove_neighborPoint(scans[i + 1], transient);
      p_cloud->points.insert(p_cloud->points.end(), kept->points.begin(), kept->points.end());
    }
    rm_triples = n >= 2 ? n - 2 : 0;
    stable_points = p_cloud->points.size();
    if (newest && n > 0) p_cloud->points.insert(p_cloud->points.end(), scans[n - 1]->points.begin(), scans[n - 1]->points.end());
  } else {
    for (size_t i = appended_scans; i < n; ++i) p_cloud->points.insert(p_cloud->points.end(), scans[i]->points.begin(), scans[i]->points.end());
    appended_scans = std::max(appended_scans, n);
    stable_points = p_cloud->points.size();
  }
  p_cloud->width = static_cast<uint32_t>(p_cloud->points.size());
  p_cloud->height = 1;
  appended_points = p_cloud->points.size();
}

void PointCloudMap::addPose(const Pose2D &p) {
  if (poses.empty()) {
    atd = 0.0;
  } else {
    const Pose2D &before = poses.back();
    const double dx = p.tx - before.tx, dy = p.ty - before.ty;
    atd += std::sqrt(dx * dx + dy * dy);
  }
  poses.emplace_back(p);
}

void PointCloudMap::addPoints(const std::vector<LPoint2D> &lps) {
  Cloud::Ptr scan_cloud = std::make_shared<Cloud>();
  scan_cloud->points.resize(lps.size());
  for (size_t i = 0; i < lps.size(); ++i) {
    scan_cloud->points[i].x = static_cast<float>(lps[i].x);   // map frame, double -> float32
    scan_cloud->points[i].y = static_cast<float>(lps[i].y);
    scan_cloud->points[i].z = 0.f;
  }
  scan_cloud->width = static_cast<uint32_t>(lps.size());
  scan_cloud->height = 1;
  scan_cloud->is_dense = false;

  Submap &cur = submaps.back();
  if (atd - cur.atdS < sepThre) {
    cur.addPoints(scan_cloud);
    cur.makeMap();
    return;
  }
  // travelled far enough: freeze the current sub-map (thinned) and open the next one, seeded with the
  // two most recent scans so the moving-object filter has its triples
  const size_t n_poses = poses.size();             // the newest pose is already in
  cur.cntE = n_poses - 2;
  cur.p_cloud = cur.filterPoints();
  cur.newest = false;

  Submap next(atd, n_poses - 1);
  const size_t have = cur.scans.size();
  if (have >= 2) {
    next.addPoints(cur.scans[have - 2]);
    next.addPoints(cur.scans[have - 1]);
  }
  next.addPoints(scan_cloud);
  next.makeMap();
  submaps.emplace_back(next);
}

// [REF src/PointCloudMap.cpp:96-116] global map = every frozen sub-map's thinned cloud + the current sub-map thinned now.
// Frozen sub-maps never change, so their part of the cloud (and of `maps`) is kept from call to call and only the current
// sub-map's part is rebuilt: the same cloud the reference produces by clearing and re-concatenating everything every key frame.
void PointCloudMap::makeGlobalMap() {
  const size_t frozen = submaps.size() - 1;
  std::vector<pcl::PointXYZ> &gm = globalMap_cloud->points;
  bool keep = gm_cloud == globalMap_cloud.get() && gm_frozen <= frozen && gm.size() >= gm_frozen_points && maps.size() >= gm_frozen;
  for (size_t i = 0; keep && i < gm_frozen; ++i) keep = maps[i].get() == submaps[i].p_cloud.get();
  if (!keep) { gm.clear(); maps.clear(); gm_frozen = 0; gm_frozen_points = 0; gm_cloud = globalMap_cloud.get(); }
  gm.resize(gm_frozen_points);
  maps.resize(gm_frozen);
  for (size_t i = gm_frozen; i < frozen; ++i) {
    gm.insert(gm.end(), submaps[i].p_cloud->points.begin(), submaps[i].p_cloud->points.end());
    maps.emplace_back(submaps[i].p_cloud);
  }
  gm_frozen = frozen;
  gm_frozen_points = gm.size();
  Cloud::Ptr current = submaps.back().filterPoints();
  gm.insert(gm.end(), current->points.begin(), current->points.end());
  maps.emplace_back(current);
  globalMap_cloud->width = static_cast<uint32_t>(gm.size());
  globalMap_cloud->height = 1;
  globalMap_cloud->is_dense = false;
}

void PointCloudMap::makeLocalMap() {
  const Cloud *prev = submaps.size() >= 2 ? submaps[submaps.size() - 2].p_cloud.get() : nullptr;   // previous sub-map only
  Submap &cur = submaps.back();
  const std::vector<pcl::PointXYZ> &prefix = cur.thinnedPrefix();
  const size_t prev_points = prev ? prev->points.size() : 0;
  std::vector<pcl::PointXYZ> &lm = localMap_cloud->points;
  const bool same_layout = lm_cur == cur.p_cloud.get() && lm_prev == prev && lm_prev_points == prev_points &&
                           lm_prefix_points <= prefix.size() && lm_fixed == prev_points + lm_prefix_points && lm.size() >= lm_fixed;
  ++localMapEpoch;
  localMapStablePrefix = same_layout ? lm_fixed : 0;
  if (!same_layout) {
    lm.clear();
    if (prev) lm.insert(lm.end(), prev->points.begin(), prev->points.end());
    lm_prev = prev; lm_prev_points = prev_points; lm_cur = cur.p_cloud.get(); lm_prefix_points = 0;
  } else {
    lm.resize(lm_fixed);                           // drop the old tail
  }
  lm.insert(lm.end(), prefix.begin() + lm_prefix_points, prefix.end());
  lm_prefix_points = prefix.size();
  lm_fixed = lm.size();
  localMapSettled = lm_fixed;
  cur.thinnedTail(lm);
  localMap_cloud->width = static_cast<uint32_t>(lm.size());
  localMap_cloud->height = 1;
  localMap_cloud->is_dense = false;
}

int savePCDFileASCII(const std::string &path, const Cloud &cloud) {
  std::ofstream out(path.c_str());
  if (!out.is_open()) return -1;
  const size_t n = cloud.points.size();
  out << "# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\n"
      << "WIDTH " << n << "\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS " << n << "\nDATA ascii\n";
  out << std::setprecision(8);
  for (const auto &p : cloud.points) out << p.x << " " << p.y << " " << p.z << "\n";
  return 0;
}

void PointCloudMap::saveGlobalMap() {
  savePCDFileASCII(map_name, *globalMap_cloud);
  for (size_t i = 0; i < maps.size(); ++i) {
    std::ostringstream name;
    name << separated_map_name << i << ".pcd";
    savePCDFileASCII(name.str(), *maps[i]);
  }
}
