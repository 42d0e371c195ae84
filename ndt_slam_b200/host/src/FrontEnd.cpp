// FrontEnd.cpp [REF src/FrontEnd.cpp:4-48]
#include "ndt_slam/FrontEnd.h"

void FrontEnd::process(Scan2D &scan) {
  if (scan.sid < startFrame) return;
  smat.matchScan(scan);
  // keyframe_skip must be > 0 (the reference divides by it as well; its C++ default of 0 is unusable)
  if (keyframeSkip > 0 && cnt % keyframeSkip == 0) pcmap->makeGlobalMap();
  ++cnt;
}
