// Pose2D.h -- planar pose (tx, ty [m], th [DEGREES]) with its cached 2x2 rotation.
// Same members and static helpers as the reference [REF include/ndt_slam/Pose2D.h:11-70,
// src/Pose2D.cpp:5-59]; th is in degrees everywhere, Rmat must be refreshed with calRmat().
#ifndef NDT_SLAM_B200_POSE2D_H_
#define NDT_SLAM_B200_POSE2D_H_

#include <cmath>
#include "LPoint2D.h"
#include "MyUtil.h"

struct Pose2D {
  double tx, ty, th;
  double Rmat[2][2];

  Pose2D() : tx(0), ty(0), th(0) { Rmat[0][0] = Rmat[1][1] = 1.0; Rmat[0][1] = Rmat[1][0] = 0.0; }
  Pose2D(double x, double y, double a) : tx(x), ty(y), th(a) { calRmat(); }
  Pose2D(double mat[2][2], double x, double y, double a) : tx(x), ty(y), th(a) {
    Rmat[0][0] = mat[0][0]; Rmat[0][1] = mat[0][1]; Rmat[1][0] = mat[1][0]; Rmat[1][1] = mat[1][1];
  }

  void calRmat() {
    const double a = DEG2RAD(th);
    const double c = std::cos(a), s = std::sin(a);
    Rmat[0][0] = c; Rmat[1][1] = c; Rmat[1][0] = s; Rmat[0][1] = -s;
  }
  void setPose(double x, double y, double a) { tx = x; ty = y; th = a; calRmat(); }
  double calDistance() const { return std::sqrt(tx * tx + ty * ty); }

  // motion from prev to cur expressed in prev's frame
  static void calMotion(Pose2D cur, Pose2D prev, Pose2D &motion);
  // motion from prev to cur expressed in the world frame
  static void calGlobalMotion(const Pose2D cur, const Pose2D prev, Pose2D &motion);
  // last (+) motion
  static void calPredPose(Pose2D motion, Pose2D last, Pose2D &pred);
  void globalPoint(const LPoint2D &in, LPoint2D &out) const;
  LPoint2D relativePoint(const LPoint2D &p) const;
  LPoint2D globalPoint(const LPoint2D &p) const;
};

#endif
