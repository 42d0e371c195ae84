// minipcl -- a restated, header-only subset of PCL 1.10.0 (TEST INFRASTRUCTURE; see oracle/README.md).
// PCL itself is not vendored by the reference and is absent from this machine.
#pragma once
namespace pcl {
struct alignas(16) PointXYZ {
  union { float data[4]; struct { float x, y, z; }; };
  PointXYZ() : x(0.f), y(0.f), z(0.f) { data[3] = 1.0f; }
  PointXYZ(float a, float b, float c) : x(a), y(b), z(c) { data[3] = 1.0f; }
};
}  // namespace pcl
