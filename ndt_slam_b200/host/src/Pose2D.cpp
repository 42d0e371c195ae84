// Pose2D.cpp -- pose algebra [REF src/Pose2D.cpp:5-59]. Angles are degrees.
#include "ndt_slam/Pose2D.h"

namespace {
// R(p)^T * (dx, dy): world displacement expressed in p's frame
inline void to_frame(const Pose2D &p, double dx, double dy, double &fx, double &fy) {
  fx = p.Rmat[0][0] * dx + p.Rmat[1][0] * dy;
  fy = p.Rmat[0][1] * dx + p.Rmat[1][1] * dy;
}
// R(p) * (x, y) + t(p)
inline void to_world(const Pose2D &p, double x, double y, double &wx, double &wy) {
  wx = p.Rmat[0][0] * x + p.Rmat[0][1] * y + p.tx;
  wy = p.Rmat[1][0] * x + p.Rmat[1][1] * y + p.ty;
}
}  // namespace

void Pose2D::calMotion(Pose2D cur, Pose2D prev, Pose2D &motion) {
  to_frame(prev, cur.tx - prev.tx, cur.ty - prev.ty, motion.tx, motion.ty);
  motion.th = MyUtil::sub_angle(cur.th, prev.th);
  motion.calRmat();
}

void Pose2D::calGlobalMotion(const Pose2D cur, const Pose2D prev, Pose2D &motion) {
  motion.tx = cur.tx - prev.tx;
  motion.ty = cur.ty - prev.ty;
  motion.th = MyUtil::sub_angle(cur.th, prev.th);
  motion.calRmat();
}

void Pose2D::calPredPose(Pose2D motion, Pose2D last, Pose2D &pred) {
  to_world(last, motion.tx, motion.ty, pred.tx, pred.ty);
  pred.th = MyUtil::add_angle(last.th, motion.th);
  pred.calRmat();
}

void Pose2D::globalPoint(const LPoint2D &in, LPoint2D &out) const { to_world(*this, in.x, in.y, out.x, out.y); }

LPoint2D Pose2D::relativePoint(const LPoint2D &p) const {
  double x, y;
  to_frame(*this, p.x - tx, p.y - ty, x, y);
  return LPoint2D(p.sid, x, y);
}

LPoint2D Pose2D::globalPoint(const LPoint2D &p) const {
  double x, y;
  to_world(*this, p.x, p.y, x, y);
  return LPoint2D(p.sid, x, y);
}
