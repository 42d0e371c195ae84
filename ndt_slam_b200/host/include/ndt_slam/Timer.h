// Timer.h -- wall-clock stopwatch with the reference's method names [REF include/ndt_slam/Timer.h:9-37];
// steady_clock instead of system_clock, and the elapsed time is readable instead of only printed.
#ifndef NDT_SLAM_B200_TIMER_H_
#define NDT_SLAM_B200_TIMER_H_
#include <chrono>
class Timer {
  std::chrono::steady_clock::time_point t0_, t1_;
 public:
  void start_timer() { t0_ = std::chrono::steady_clock::now(); }
  void end_timer() { t1_ = std::chrono::steady_clock::now(); }
  double elapsed_ms() const { return std::chrono::duration<double, std::milli>(t1_ - t0_).count(); }
  void print_timer() const {}
};
#endif
