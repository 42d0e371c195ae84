// Scan2D.h -- one scan: id, odometry pose at capture time, points [REF include/ndt_slam/Scan2D.h:15-35].
#ifndef NDT_SLAM_B200_SCAN2D_H_
#define NDT_SLAM_B200_SCAN2D_H_

#include <string>
#include <vector>
#include <ros/ros.h>
#include "LPoint2D.h"
#include "Pose2D.h"

namespace std_msgs { struct Header { uint32_t seq = 0; ros::Time stamp; std::string frame_id; }; }

struct Scan2D {
  int sid;
  Pose2D pose;
  std::vector<LPoint2D> lps;
  std_msgs::Header header;

  Scan2D() : sid(0) {}
  void setLps(const std::vector<LPoint2D> &ps) { lps = ps; }
};

#endif
