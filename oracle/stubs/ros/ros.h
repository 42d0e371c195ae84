// Minimal stand-in for <ros/ros.h> so the reference sources compile unmodified (TEST INFRASTRUCTURE).
// Parameters live in a process-wide string map; logging is off unless REF_ROS_LOG is set.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <map>
#include <sstream>
#include <string>

namespace ros {
struct Time {
  uint32_t sec = 0, nsec = 0;
  static Time now() { return Time(); }
};
inline bool &log_enabled() { static bool on = std::getenv("REF_ROS_LOG") != nullptr; return on; }
namespace param {
inline std::map<std::string, std::string> &store() { static std::map<std::string, std::string> m; return m; }
inline void set(const std::string &k, const std::string &v) { store()[k] = v; }
template <class T> inline bool get(const std::string &k, T &v) {
  auto it = store().find(k);
  if (it == store().end()) return false;
  std::istringstream is(it->second);
  T t; is >> t;
  if (is.fail()) return false;
  v = t; return true;
}
template <> inline bool get<bool>(const std::string &k, bool &v) {
  auto it = store().find(k);
  if (it == store().end()) return false;
  v = (it->second == "true" || it->second == "1" || it->second == "True");
  return true;
}
template <> inline bool get<std::string>(const std::string &k, std::string &v) {
  auto it = store().find(k);
  if (it == store().end()) return false;
  v = it->second; return true;
}
}  // namespace param
inline bool ok() { return true; }
inline void init(int &, char **, const std::string &) {}
struct Publisher { template <class M> void publish(const M &) const {} };
struct Subscriber {};
struct NodeHandle { template <class M> Publisher advertise(const std::string &, int) { return Publisher(); } };
}  // namespace ros

#define ROS_INFO(...) do { if (ros::log_enabled()) { std::printf(__VA_ARGS__); std::printf("\n"); } } while (0)
#define ROS_INFO_STREAM(x) do { if (ros::log_enabled()) { std::cout << x << std::endl; } } while (0)
