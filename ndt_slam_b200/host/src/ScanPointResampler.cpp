// ScanPointResampler.cpp -- arc-length resampling [REF src/ScanPointResampler.cpp:4-62].
// Walk the scan in beam order accumulating the travelled arc length `dis`. A candidate point is
//   dropped            while dis + L <  space,
//   kept as it is      when  dis + L >= spaceThre (a real gap: no interpolation across it),
//   replaced by an interpolated point at exactly `space` otherwise; the same input point is then
//   examined again from the new anchor.
#include "ndt_slam/ScanPointResampler.h"

#include <cmath>
#include <vector>

void ScanPointResampler::resamplePoints(Scan2D *scan) {
  std::vector<LPoint2D> &in = scan->lps;
  if (in.empty()) return;

  std::vector<LPoint2D> kept;
  kept.reserve(in.size());
  dis = 0;
  LPoint2D anchor = in[0];                       // the point distances are measured from
  kept.push_back(LPoint2D(anchor.sid, anchor.x, anchor.y));

  size_t i = 1;
  while (i < in.size()) {
    const LPoint2D &cand = in[i];
    LPoint2D fresh;
    bool interpolated = false;
    if (findInterpolatePoint(cand, anchor, fresh, interpolated)) {
      kept.push_back(fresh);
      anchor = fresh;
      dis = 0;
      if (!interpolated) ++i;                    // an interpolated point sits before cand: look at cand again
    } else {
      anchor = cand;
      ++i;
    }
  }
  scan->setLps(kept);
}

bool ScanPointResampler::findInterpolatePoint(const LPoint2D &cp, const LPoint2D &pp, LPoint2D &np, bool &inserted) {
  const double dx = cp.x - pp.x, dy = cp.y - pp.y;
  const double L = std::sqrt(dx * dx + dy * dy);
  const double reach = dis + L;
  if (reach < space) {
    dis += L;
    return false;
  }
  if (reach >= spaceThre) {
    np.setData(cp.sid, cp.x, cp.y);
    return true;
  }
  const double ratio = (space - dis) / L;
  np.setData(cp.sid, dx * ratio + pp.x, dy * ratio + pp.y);
  inserted = true;
  return true;
}
