// stand-in for <pcl_ros/point_cloud.h>: publishing a pcl::PointCloud is a no-op here (TEST INFRASTRUCTURE)
#pragma once
#include <pcl/point_cloud.h>
