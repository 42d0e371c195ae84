// ScanPointResampler.h -- arc-length resampling of a scan to a uniform point spacing.
// Interface of the reference class [REF include/ndt_slam/ScanPointResampler.h:12-37]; the algorithm is
// inherently sequential (each output point depends on the previous output), so it stays on the host.
#ifndef NDT_SLAM_B200_SCANPOINTRESAMPLER_H_
#define NDT_SLAM_B200_SCANPOINTRESAMPLER_H_

#include <ros/ros.h>
#include "LPoint2D.h"
#include "Scan2D.h"

class ScanPointResampler {
  double space;       // target spacing [m]
  double spaceThre;   // gaps at least this long are kept as they are (no interpolation) [m]
  double dis;         // arc length walked since the last emitted point

 public:
  ScanPointResampler() : space(0.0), spaceThre(0.0), dis(0.0) {
    ros::param::get("space", space);
    ros::param::get("space_thre", spaceThre);
  }
  // int parameters, as declared by the reference (h:29-32)
  void setDthre(int s, int l) { space = s; spaceThre = l; }

  void resamplePoints(Scan2D *scan);
  bool findInterpolatePoint(const LPoint2D &cp, const LPoint2D &pp, LPoint2D &np, bool &inserted);
};

#endif
