#pragma once
namespace sensor_msgs { struct PointCloud {}; }
