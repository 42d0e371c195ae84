// MyUtil.h -- angle helpers used on the hot path's host side.
// Mirrors the part of the reference's MyUtil that the path uses [REF include/ndt_slam/MyUtil.h:22-23,
// src/MyUtil.cpp:4-24]: DEG2RAD / RAD2DEG and add_angle / sub_angle with wrap into [-180, 180).
// The reference's calEigen2D / svdInverse / convertCov / quaternion helpers are dead code or ROS
// plumbing and are not reproduced.
#ifndef NDT_SLAM_B200_MYUTIL_H_
#define NDT_SLAM_B200_MYUTIL_H_

#include <cmath>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

#define DEG2RAD(x) ((x)*M_PI/180)
#define RAD2DEG(x) ((x)*180/M_PI)

class MyUtil {
 public:
  // degrees in, degrees out, result wrapped once into [-180, 180)
  static double add_angle(double a1, double a2) { return wrap(a1 + a2); }
  static double sub_angle(double a1, double a2) { return wrap(a1 - a2); }

 private:
  static double wrap(double deg) {
    if (deg < -180) return deg + 360;
    if (deg >= 180) return deg - 360;
    return deg;
  }
};

#endif
