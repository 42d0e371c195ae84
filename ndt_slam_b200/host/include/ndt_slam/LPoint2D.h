// LPoint2D.h -- one scan point (boundary value type; layout and semantics as the reference's
// [REF include/ndt_slam/LPoint2D.h:13-69]: everything double, sid = scan id, normals unused on this path).
#ifndef NDT_SLAM_B200_LPOINT2D_H_
#define NDT_SLAM_B200_LPOINT2D_H_

#include <cmath>
#include "MyUtil.h"

struct Vector2D { double x, y; };

enum ptype { UNKNOWN = 0, LINE = 1, CORNER = 2, ISOLATE = 3 };

struct LPoint2D {
  int sid;
  double x, y;
  double nx, ny;
  double atd;
  ptype type;

  LPoint2D() : x(0), y(0) { init(); }
  LPoint2D(int id, double px, double py) : x(px), y(py) { init(); sid = id; }
  LPoint2D(double px, double py) : x(px), y(py) { init(); }

  // resets everything except the position
  void init() { sid = -1; nx = ny = 0; atd = 0; type = UNKNOWN; }
  void setData(int id, double px, double py) { init(); sid = id; x = px; y = py; }
  void set_RangeAngle2XY(double range, double angle_deg) {
    const double a = DEG2RAD(angle_deg);
    x = range * std::cos(a);
    y = range * std::sin(a);
  }
  void setType(ptype t) { type = t; }
  void setNormal(double px, double py) { nx = px; ny = py; }
};

#endif
