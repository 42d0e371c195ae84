#pragma once
namespace sensor_msgs { struct ChannelFloat32 {}; }
