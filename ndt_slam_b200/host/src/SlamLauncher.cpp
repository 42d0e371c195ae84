// SlamLauncher.cpp -- text-log driven loop and output writers [REF src/SlamLauncher.cpp:7-141].
#include "ndt_slam/SlamLauncher.h"

#include <string>

SlamLauncher::SlamLauncher() : end_frame(100), rate(0), drawSkip(0), stamp(0), sidelidar(true), ok(true), scansProcessed(0) {
  ros::param::get("sidelidar", sidelidar);
  ros::param::get("end_frame", end_frame);
  ros::param::get("draw_skip", drawSkip);
  ros::param::get("filename_in", filename_in);
  ros::param::get("poses_name", poses_name);
  inputfile.open(filename_in, std::ios::in);
  if (!inputfile) ok = false;
  outputfile.open(poses_name, std::ios::out);
  if (!outputfile) ok = false;
}

void SlamLauncher::init() {
  frontEnd.setPoseEstimator(&estim);
  frontEnd.setPointCloudMap(&pcmap);
}

void SlamLauncher::readFormat() {
  std::string line;
  for (int i = 0; i < 4; ++i) std::getline(inputfile, line);
}

void SlamLauncher::output_file_poses(std::vector<Pose2D> poses) {
  outputfile << poses.size() << std::endl;
  for (size_t i = 0; i < poses.size(); i += 10) outputfile << poses[i].tx << " " << poses[i].ty << " " << poses[i].th << " " << std::endl;
}

// One record: "stamp x y theta_deg <rest of line: image name>" then three groups
// "count x y x y ..." for the front, left and right lidars; side lidars are used only if sidelidar.
bool SlamLauncher::input_file_line() {
  std::string tok;
  auto next = [&](char delim) -> std::string & { std::getline(inputfile, tok, delim); return tok; };
  stamp = std::stoi(next(' '));
  scan.sid = stamp;
  scan.pose.tx = std::stod(next(' '));
  scan.pose.ty = std::stod(next(' '));
  scan.pose.th = std::stod(next(' '));
  std::getline(inputfile, tok);                     // image name, unused

  std::vector<LPoint2D> pts;
  for (int group = 0; group < 3; ++group) {
    const int count = std::stoi(next(' '));
    const bool use = (group == 0) || sidelidar;
    for (int i = 0; i < count; ++i) {
      const double x = std::stod(next(' '));
      const double y = std::stod(next(' '));
      if (use) { LPoint2D lp; lp.setData(stamp, x, y); pts.push_back(lp); }
    }
  }
  scan.lps = pts;
  scan.pose.calRmat();
  if (inputfile.eof()) {
    inputfile.close();
    return true;
  }
  return false;
}

void SlamLauncher::loop_wait() {
  if (!ok) return;
  int cnt = 1;
  readFormat();
  for (;;) {
    if (cnt > end_frame || input_file_line()) {
      output_file_poses(frontEnd.get_poses());
      frontEnd.saveMap();
      return;
    }
    frontEnd.process(scan);
    ++scansProcessed;
    ++cnt;
  }
}
