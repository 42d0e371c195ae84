// SlamLauncher.h -- file-driven main loop: reads the reference's text scan log, feeds FrontEnd, writes
// the poses file and the PCD maps [REF include/ndt_slam/SlamLauncher.h:36-108, src/SlamLauncher.cpp:7-141;
// formats: SURVEY.md App. D]. The ROS publishers of the reference (pcmap / poses topics) are visual
// plumbing and are not reproduced; file formats are kept so outputs can be diffed with a reference run.
#ifndef NDT_SLAM_B200_SLAMLAUNCHER_H_
#define NDT_SLAM_B200_SLAMLAUNCHER_H_

#include <fstream>
#include <string>
#include <vector>
#include <ros/ros.h>

#include "FrontEnd.h"
#include "PointCloudMap.h"
#include "PoseEstimator.h"
#include "Timer.h"

class SlamLauncher {
 private:
  int end_frame;

 public:
  PointCloudMap pcmap;
  FrontEnd frontEnd;
  PoseEstimator estim;

  Timer timer;

  int rate;
  int drawSkip;
  int stamp;
  Scan2D scan;

  std::ifstream inputfile;
  std::ofstream outputfile;
  std::string filename_in, poses_name;

  bool sidelidar;
  bool ok;              // false when the input file could not be opened (the reference exits the process)
  int scansProcessed;

  SlamLauncher();

  void readFormat();                                  // skips the 4 header lines
  void output_file_poses(std::vector<Pose2D> poses);  // count, then every 10th pose "tx ty th "
  bool input_file_line();                             // parses one scan record; true = end of data
  void init();
  void loop_wait();
};

#endif
