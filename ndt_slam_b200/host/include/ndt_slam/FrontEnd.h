// FrontEnd.h -- per-scan entry point: scan matching plus a global-map rebuild every keyframe_skip scans
// [REF include/ndt_slam/FrontEnd.h:16-103, src/FrontEnd.cpp:4-48]. The loop-closure / pose-graph parts are commented out
// in the reference [REF src/FrontEnd.cpp:20-44]; here they exist behind the parameter `loop_closure` (default false =
// the reference's behaviour): key frames become pose-graph nodes joined by odometry arcs, and LoopDetector verifies
// revisit candidates in one batched device call. No pose adjustment follows.
#ifndef NDT_SLAM_B200_FRONTEND_H_
#define NDT_SLAM_B200_FRONTEND_H_

#include <vector>
#include <ros/ros.h>
#include "LoopDetector.h"
#include "PointCloudMap.h"
#include "PoseGraph.h"
#include "Scan2D.h"
#include "ScanMatcher.h"
#include "Timer.h"

class FrontEnd {
 private:
  int cnt;
  int keyframeSkip;
  int startFrame;

  ScanMatcher smat;
  PointCloudMap *pcmap;
  Timer timer;
  bool loopClosure;

 public:
  PoseGraph pg;                  // key-frame nodes, odometry arcs and verified loop arcs (loop_closure only)
  LoopDetector lpd;
  int loopsDetected;

  FrontEnd() : cnt(0), keyframeSkip(0), startFrame(0), pcmap(nullptr), loopClosure(false), loopsDetected(0) {
    ros::param::get("keyframe_skip", keyframeSkip);
    ros::param::get("start_frame", startFrame);
    ros::param::get("loop_closure", loopClosure);
    lpd.setPoseGraph(&pg);
  }

  geometry_msgs::PoseArray get_poseArray() { return smat.get_poseArray(); }
  std::vector<Pose2D> get_poses() { return smat.poses; }
  const ScanMatcher &matcher() const { return smat; }      // instrumentation access (not in the reference)
  ScanMatcher &matcher() { return smat; }

  void saveMap() {
    pcmap->makeGlobalMap();
    pcmap->saveGlobalMap();
  }

  void setPoseEstimator(PoseEstimator *estim_) { smat.setPoseEstimator(estim_); }
  void setPointCloudMap(PointCloudMap *pcmap_) {
    pcmap = pcmap_;
    smat.setPointCloudMap(pcmap_);
  }

  void process(Scan2D &scan);
};

#endif
