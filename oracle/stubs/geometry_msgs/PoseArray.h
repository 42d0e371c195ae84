#pragma once
#include <vector>
#include <std_msgs/Header.h>
#include <geometry_msgs/Pose.h>
namespace geometry_msgs { struct PoseArray { std_msgs::Header header; std::vector<Pose> poses; }; }
