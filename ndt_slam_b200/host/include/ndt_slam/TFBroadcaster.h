// TFBroadcaster.h -- the reference publishes a map->odom TF here [REF include/ndt_slam/TFBroadcaster.h:13-46].
// ROS visual plumbing, outside the hot path: kept as a no-op that remembers the last transform.
#ifndef NDT_SLAM_B200_TFBROADCASTER_H_
#define NDT_SLAM_B200_TFBROADCASTER_H_
#include "Pose2D.h"
class TFBroadcaster {
 public:
  Pose2D last_map2odom;
  TFBroadcaster() {}
  void publish_tf_map2odom(Pose2D &pose) { last_map2odom = pose; }
};
#endif
