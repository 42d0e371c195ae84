// Stand-in for tf: only what MyUtil.h / TFBroadcaster.h touch (visual plumbing, no arithmetic on the hot path).
#pragma once
#include <cmath>
#include <string>
#include <ros/ros.h>
#include <geometry_msgs/Pose.h>
namespace tf {
struct Vector3 { double x, y, z; Vector3(double a = 0, double b = 0, double c = 0) : x(a), y(b), z(c) {} };
struct Quaternion {
  double x = 0, y = 0, z = 0, w = 1;
  Quaternion() {}
  Quaternion(double a, double b, double c, double d) : x(a), y(b), z(c), w(d) {}
  Quaternion(double yaw, double pitch, double roll) { setRPY(roll, pitch, yaw); }
  void setRPY(double roll, double pitch, double yaw) {
    double cy = std::cos(yaw / 2), sy = std::sin(yaw / 2), cp = std::cos(pitch / 2), sp = std::sin(pitch / 2);
    double cr = std::cos(roll / 2), sr = std::sin(roll / 2);
    x = sr * cp * cy - cr * sp * sy; y = cr * sp * cy + sr * cp * sy; z = cr * cp * sy - sr * sp * cy; w = cr * cp * cy + sr * sp * sy;
  }
};
struct Matrix3x3 {
  Quaternion q;
  explicit Matrix3x3(const Quaternion &qq) : q(qq) {}
  void getRPY(double &roll, double &pitch, double &yaw) const {
    roll = std::atan2(2 * (q.w * q.x + q.y * q.z), 1 - 2 * (q.x * q.x + q.y * q.y));
    pitch = std::asin(2 * (q.w * q.y - q.z * q.x));
    yaw = std::atan2(2 * (q.w * q.z + q.x * q.y), 1 - 2 * (q.y * q.y + q.z * q.z));
  }
};
inline Quaternion createQuaternionFromRPY(double r, double p, double y) { Quaternion q; q.setRPY(r, p, y); return q; }
inline void quaternionMsgToTF(const geometry_msgs::Quaternion &m, Quaternion &q) { q = Quaternion(m.x, m.y, m.z, m.w); }
inline void quaternionTFToMsg(const Quaternion &q, geometry_msgs::Quaternion &m) { m.x = q.x; m.y = q.y; m.z = q.z; m.w = q.w; }
struct StampedTransform {
  std::string frame_id_, child_frame_id_;
  ros::Time stamp_;
  Vector3 origin; Quaternion rot;
  void setOrigin(const Vector3 &v) { origin = v; }
  void setRotation(const Quaternion &q) { rot = q; }
};
struct TransformBroadcaster { void sendTransform(const StampedTransform &) {} };
}  // namespace tf
