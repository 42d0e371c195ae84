// FrontEnd.cpp [REF src/FrontEnd.cpp:4-48]
#include "ndt_slam/FrontEnd.h"

void FrontEnd::process(Scan2D &scan) {
  if (scan.sid < startFrame) return;
  smat.matchScan(scan);
  // keyframe_skip must be > 0 (the reference divides by it as well; its C++ default of 0 is unusable)
  const bool keyframe = keyframeSkip > 0 && cnt % keyframeSkip == 0;
  if (keyframe) pcmap->makeGlobalMap();
  if (keyframe && loopClosure) {
    // what the reference sketches in comments [REF src/FrontEnd.cpp:20-44]: a node per key frame, an odometry arc from the
    // previous one, then the loop detector (candidates + batched NDT verification + loop arcs)
    const Pose2D curPose = pcmap->getLastPose();
    PoseNode *last = pg.nodes.empty() ? nullptr : pg.nodes.back();
    PoseNode *node = pg.addNode(curPose);
    if (last) {
      Pose2D rel;
      Pose2D::calMotion(curPose, last->pose, rel);
      Eigen::Matrix3d cov = smat.lastCov;
      for (int i = 0; i < 3; ++i) cov(i, i) += 1e-12;          // the first fused covariance can be exactly singular
      pg.addArc(pg.makeArc(last->nid, node->nid, rel, cov));
    }
    if (lpd.detectLoop(&scan, curPose, node->nid, pcmap->atd)) ++loopsDetected;
  }
  ++cnt;
}
