// PoseFuser.h -- EKF-style fusion of the odometry prediction with the NDT pose.
// Interface of the reference class [REF include/ndt_slam/PoseFuser.h:10-38, src/PoseFuser.cpp:3-61].
#ifndef NDT_SLAM_B200_POSEFUSER_H_
#define NDT_SLAM_B200_POSEFUSER_H_

#include <Eigen/Core>
#include <ros/ros.h>
#include "Pose2D.h"

class PoseFuser {
  double delTime;    // scan interval [s]
  double coeVel;     // translational odometry noise coefficient
  double coeOmega;   // rotational odometry noise coefficient

 public:
  PoseFuser() : delTime(0.5), coeVel(0.1), coeOmega(0.1) {
    ros::param::get("delTime", delTime);
    ros::param::get("coeVel", coeVel);
    ros::param::get("coeOmega", coeOmega);
  }

  void fusePose(const Pose2D &predPose, const Pose2D &estPose, const Pose2D &odoMotion, const Pose2D &lastPose,
                const Eigen::Matrix3d &lastCov, const Eigen::Matrix3d &Qmat, Pose2D &fusedPose, Eigen::Matrix3d &cov);
  void calOdometryCovariance(const Pose2D &odoMotion, const Pose2D &lastPose, const Eigen::Matrix3d &lastCov,
                             Eigen::Matrix3d &cov);
};

#endif
