// PoseFuser.cpp -- EKF predict / update on (x, y, yaw) [REF src/PoseFuser.cpp:3-61].
// The state is metres and radians internally, Pose2D carries degrees.
#include "ndt_slam/PoseFuser.h"

#include <cmath>

void PoseFuser::calOdometryCovariance(const Pose2D &odoMotion, const Pose2D &lastPose, const Eigen::Matrix3d &lastCov,
                                      Eigen::Matrix3d &cov) {
  // velocity motion model: v, omega from the odometry increment over one scan interval
  const double v = odoMotion.calDistance() / delTime;
  const double omega = DEG2RAD(odoMotion.th / delTime);
  const double heading = DEG2RAD(lastPose.th);
  const double ch = std::cos(heading), sh = std::sin(heading);

  Eigen::Matrix2d M;                       // control noise
  M << coeVel * v * v, 0.0, 0.0, coeOmega * omega * omega;
  Eigen::Matrix<double, 3, 2> A;           // d state / d control
  A << delTime * ch, 0.0, delTime * sh, 0.0, 0.0, delTime;
  Eigen::Matrix3d F;                       // d state / d previous state
  F << 1.0, 0.0, -v * delTime * sh, 0.0, 1.0, v * delTime * ch, 0.0, 0.0, 1.0;

  cov = F * lastCov * F.transpose() + A * M * A.transpose();
}

void PoseFuser::fusePose(const Pose2D &predPose, const Pose2D &estPose, const Pose2D &odoMotion, const Pose2D &lastPose,
                         const Eigen::Matrix3d &lastCov, const Eigen::Matrix3d &Qmat, Pose2D &fusedPose, Eigen::Matrix3d &cov) {
  Eigen::Matrix3d predicted;
  calOdometryCovariance(odoMotion, lastPose, lastCov, predicted);

  const Eigen::Matrix3d gain = predicted * (Qmat + predicted).inverse();
  cov = (Eigen::Matrix3d::Identity() - gain) * predicted;

  Eigen::Vector3d innovation;
  innovation << estPose.tx - predPose.tx, estPose.ty - predPose.ty, DEG2RAD(MyUtil::sub_angle(estPose.th, predPose.th));
  Eigen::Vector3d prior;
  prior << predPose.tx, predPose.ty, DEG2RAD(predPose.th);
  const Eigen::Vector3d post = gain * innovation + prior;

  fusedPose.setPose(post(0), post(1), RAD2DEG(post(2)));
}
