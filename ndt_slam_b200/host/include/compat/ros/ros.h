// compat/ros/ros.h -- just enough of the ROS surface for the host classes to keep the reference's
// constructor behaviour (parameters are read with ros::param::get in constructors, e.g.
// PoseEstimator.h:65-70) without ROS. A catkin build drops this directory from the include path.
#pragma once
#include <cstdint>
#include <map>
#include <sstream>
#include <string>

namespace ros {
struct Time { uint32_t sec = 0, nsec = 0; static Time now() { return Time(); } };
namespace param {
inline std::map<std::string, std::string> &table() { static std::map<std::string, std::string> t; return t; }
inline void set(const std::string &name, const std::string &value) { table()[name] = value; }
inline void clear() { table().clear(); }
inline bool lookup(const std::string &name, std::string &raw) {
  auto it = table().find(name);
  if (it == table().end()) return false;
  raw = it->second;
  return true;
}
inline bool get(const std::string &name, std::string &out) { return lookup(name, out); }
inline bool get(const std::string &name, bool &out) {
  std::string raw;
  if (!lookup(name, raw)) return false;
  out = (raw == "true" || raw == "True" || raw == "1");
  return true;
}
template <class Num> inline bool get(const std::string &name, Num &out) {
  std::string raw;
  if (!lookup(name, raw)) return false;
  std::istringstream in(raw);
  Num v;
  if (!(in >> v)) return false;
  out = v;
  return true;
}
}  // namespace param
}  // namespace ros

#ifndef ROS_INFO
#define ROS_INFO(...) ((void)0)
#define ROS_INFO_STREAM(x) ((void)0)
#endif
