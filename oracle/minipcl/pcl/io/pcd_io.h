// pcl::io::savePCDFileASCII for PointXYZ (PCD v0.7 ASCII): the reference's map output format
// (PointCloudMap.h:124-136).
#pragma once
#include <fstream>
#include <iomanip>
#include <string>
#include <pcl/point_cloud.h>
namespace pcl { namespace io {
template <class PointT> inline int savePCDFileASCII(const std::string &file, const PointCloud<PointT> &cloud, int precision = 8) {
  std::ofstream fs(file.c_str());
  if (!fs.is_open()) return -1;
  const std::size_t n = cloud.points.size();
  const uint32_t w = cloud.width * cloud.height == n ? cloud.width : (uint32_t)n, h = cloud.width * cloud.height == n ? cloud.height : 1;
  fs << "# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\n"
     << "WIDTH " << w << "\nHEIGHT " << h << "\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS " << n << "\nDATA ascii\n";
  fs << std::setprecision(precision);
  for (const auto &p : cloud.points) fs << p.x << " " << p.y << " " << p.z << "\n";
  return 0;
}
}}  // namespace pcl::io
