// stand-in for <pcl_conversions/pcl_conversions.h>: only toPCL(ros::Time, stamp) is used (SlamLauncher.cpp:132)
#pragma once
#include <cstdint>
#include <ros/ros.h>
namespace pcl_conversions {
template <class Stamp> inline void toPCL(const ros::Time &t, Stamp &pcl_stamp) { pcl_stamp = static_cast<Stamp>(t.sec) * 1000000u + t.nsec / 1000u; }
}
