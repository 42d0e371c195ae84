// ScanMatcher.cpp -- the per-scan pipeline [REF src/ScanMatcher.cpp:4-117].
#include "ndt_slam/ScanMatcher.h"

#include <chrono>
#include <cmath>

namespace {
inline double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
}  // namespace

geometry_msgs::Pose ScanMatcher::toMsg(const Pose2D &pose) {
  geometry_msgs::Pose m;
  m.position.x = pose.tx;
  m.position.y = pose.ty;
  const double half = 0.5 * DEG2RAD(pose.th);      // yaw-only quaternion
  m.orientation.z = std::sin(half);
  m.orientation.w = std::cos(half);
  return m;
}

void ScanMatcher::savePose(const std_msgs::Header &header, const Pose2D &pose, const Eigen::Matrix3d &cov) {
  poses.push_back(pose);
  Covs.push_back(cov);
  poseArray.poses.push_back(toMsg(pose));
  poseArray.header = header;
  poseArray.header.frame_id = "map";
}

void ScanMatcher::remakePoseArray(std::vector<Pose2D> &poses_) {
  poseArray.poses.clear();
  poses.clear();
  for (const Pose2D &p : poses_) {
    poseArray.poses.push_back(toMsg(p));
    poses.push_back(p);
  }
  poseArray.header.stamp = ros::Time::now();
  poseArray.header.frame_id = "map";
}

bool ScanMatcher::matchScan(Scan2D &curScan) {
  double t = now_ms();
  spres.resamplePoints(&curScan);
  msResample = now_ms() - t;

  if (cnt == 0) {
    // the first scan only seeds the map at its odometry pose
    initPose = curScan.pose;
    lastCov = Eigen::Matrix3d::Zero();
    t = now_ms();
    growMap(curScan, initPose);
    msGrowMap = now_ms() - t;
    msEstimate = msFuse = 0;
    savePose(curScan.header, initPose, Eigen::Matrix3d::Zero());
    tfb.publish_tf_map2odom(initPose);
    prevScan = curScan;
    ++cnt;
    return true;
  }

  // odometry increment since the previous scan, applied to the last estimated pose
  Pose2D odoMotion;
  Pose2D::calMotion(curScan.pose, prevScan.pose, odoMotion);
  const Pose2D lastPose = pcmap->getLastPose();
  Pose2D predPose;
  Pose2D::calPredPose(odoMotion, lastPose, predPose);

  // NDT against the local map, starting from the prediction
  t = now_ms();
  estim->setScanPair(&curScan, pcmap->localMap_cloud);
  estim->hintTargetPrefix(pcmap->localMapEpoch, pcmap->localMapStablePrefix, pcmap->localMapSettled);   // optional: upload only what makeLocalMap changed
  Pose2D estPose;
  Eigen::Matrix3d Qmat;
  const double cost = estim->estimatePose(predPose, estPose, Qmat);
  msEstimate = now_ms() - t;
  const bool successful = (cost <= scthre);

  t = now_ms();
  Eigen::Matrix3d cov;
  Pose2D fusedPose;
  if (successful) {
    pfu.fusePose(predPose, estPose, odoMotion, lastPose, lastCov, Qmat, fusedPose, cov);
  } else {
    pfu.calOdometryCovariance(odoMotion, lastPose, lastCov, cov);
    fusedPose = predPose;
  }
  lastCov = forceCov ? *forceCov : cov;
  msFuse = now_ms() - t;

  t = now_ms();
  growMap(curScan, forcePose ? *forcePose : fusedPose);
  msGrowMap = now_ms() - t;
  prevScan = curScan;

  savePose(curScan.header, fusedPose, cov);
  Pose2D drift;
  Pose2D::calGlobalMotion(fusedPose, curScan.pose, drift);
  tfb.publish_tf_map2odom(drift);

  ++cnt;
  return successful;
}

void ScanMatcher::growMap(const Scan2D &scan, const Pose2D &pose) {
  std::vector<LPoint2D> in_map;
  in_map.reserve(scan.lps.size());
  const double (*R)[2] = pose.Rmat;
  for (const LPoint2D &lp : scan.lps) {
    if (lp.type == ISOLATE) continue;
    LPoint2D g(scan.sid, R[0][0] * lp.x + R[0][1] * lp.y + pose.tx, R[1][0] * lp.x + R[1][1] * lp.y + pose.ty);
    g.setNormal(R[0][0] * lp.nx + R[0][1] * lp.ny, R[1][0] * lp.nx + R[1][1] * lp.ny);
    g.setType(lp.type);
    in_map.push_back(g);
  }
  pcmap->addPose(pose);
  pcmap->addPoints(in_map);
  pcmap->setLastPose(pose);
  pcmap->setLastScan(scan);
  pcmap->makeLocalMap();
  // the cloud the next estimatePose matches against exists now: let the device bring its grid up to date while the host
  // goes on (FrontEnd bookkeeping, the next scan's resampling and source filter)
  if (estim) estim->prefetchTarget(pcmap->localMap_cloud, pcmap->localMapEpoch, pcmap->localMapStablePrefix, pcmap->localMapSettled);
}
