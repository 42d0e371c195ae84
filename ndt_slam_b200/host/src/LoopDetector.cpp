// LoopDetector.cpp -- see LoopDetector.h. [REF src/FrontEnd.cpp:32-44 (call site); src/PoseEstimator.cpp:4-69 (what one
// verification is)]
#include "ndt_slam/LoopDetector.h"

#include <algorithm>
#include <cmath>
#include <stdexcept>
#include <string>

#include "ndt_slam/MyUtil.h"

LoopDetector::LoopDetector()
    : radius(4.0), minTravel(15.0), maxCandidates(8), scoreThre(0.5), coeNDTCov(1.0), Resolution(1.0), StepSize(0.1),
      TransformationEpsilon(0.01), LeafSize(0.1), MaximumIterations(35) {
  ros::param::get("loop_radius", radius);
  ros::param::get("loop_min_travel", minTravel);
  ros::param::get("loop_max_candidates", maxCandidates);
  ros::param::get("loop_score_thre", scoreThre);
  ros::param::get("coeNDTCov", coeNDTCov);
  ros::param::get("Resolution", Resolution);            // the matcher's own parameters (PoseEstimator.h:63-84)
  ros::param::get("StepSize", StepSize);
  ros::param::get("TransformationEpsilon", TransformationEpsilon);
  ros::param::get("MaximumIterations", MaximumIterations);
  ros::param::get("LeafSize", LeafSize);
}

LoopDetector::~LoopDetector() {
  if (ndt) ndt_destroy(ndt);
}

void LoopDetector::ensureHandle() {
  if (ndt) return;
  ndt_params prm;
  ndt_params_default(&prm);
  prm.resolution = static_cast<float>(Resolution);
  prm.step_size = StepSize;
  prm.trans_eps = TransformationEpsilon;
  prm.max_iter = MaximumIterations;
  int dev = 0;
  ros::param::get("cuda_device", dev);
  prm.device = dev;
  if (ndt_create(&prm, &ndt) != NDT_OK) throw std::runtime_error(std::string("LoopDetector: ndt_create: ") + ndt_last_error(nullptr));
}

std::vector<int> LoopDetector::findCandidates(const Pose2D &curPose, double atd) const {
  std::vector<std::pair<double, int>> near;
  for (size_t k = 0; k < frames.size(); ++k) {
    const KeyFrame &f = frames[k];
    if (atd - f.atd < minTravel) continue;                          // still the stretch just driven
    const double d = std::hypot(curPose.tx - f.pose.tx, curPose.ty - f.pose.ty);
    if (d <= radius) near.emplace_back(d, static_cast<int>(k));
  }
  std::sort(near.begin(), near.end());
  std::vector<int> out;
  for (size_t k = 0; k < near.size() && static_cast<int>(k) < maxCandidates; ++k) out.push_back(near[k].second);
  return out;
}

bool LoopDetector::detectLoop(const Scan2D *curScan, const Pose2D &curPose, int nodeId, double atd) {
  lastMatches.clear();
  lastVerifyMs = 0.0;
  KeyFrame kf;
  kf.nodeId = nodeId; kf.pose = curPose; kf.atd = atd;
  kf.xyzw.resize(4 * curScan->lps.size());
  for (size_t i = 0; i < curScan->lps.size(); ++i) {               // double -> float32 exactly where setScanPair casts
    kf.xyzw[4 * i] = static_cast<float>(curScan->lps[i].x);
    kf.xyzw[4 * i + 1] = static_cast<float>(curScan->lps[i].y);
    kf.xyzw[4 * i + 2] = 0.f; kf.xyzw[4 * i + 3] = 0.f;
  }
  const std::vector<int> cand = findCandidates(curPose, atd);
  bool found = false;
  if (!cand.empty()) {
    ensureHandle();
    const int64_t n = static_cast<int64_t>(cand.size());
    const int64_t ns = static_cast<int64_t>(curScan->lps.size());
    std::vector<float> src, tgt;
    std::vector<int64_t> so(n + 1, 0), to(n + 1, 0);
    std::vector<double> guess(3 * n);
    for (int64_t k = 0; k < n; ++k) {
      const KeyFrame &f = frames[cand[k]];
      src.insert(src.end(), kf.xyzw.begin(), kf.xyzw.end());
      tgt.insert(tgt.end(), f.xyzw.begin(), f.xyzw.end());
      so[k + 1] = so[k] + ns;
      to[k + 1] = to[k] + static_cast<int64_t>(f.xyzw.size() / 4);
      Pose2D rel;
      Pose2D::calMotion(curPose, f.pose, rel);                     // the current pose seen from the candidate's
      guess[3 * k] = rel.tx; guess[3 * k + 1] = rel.ty; guess[3 * k + 2] = DEG2RAD(rel.th);
    }
    std::vector<ndt_result> res(n);
    if (ndt_match_pairs(ndt, src.data(), so.data(), tgt.data(), to.data(), guess.data(), n, static_cast<float>(LeafSize), NDT_MEM_HOST,
                        res.data()) != NDT_OK)
      throw std::runtime_error(std::string("LoopDetector: ndt_match_pairs: ") + ndt_last_error(ndt));
    float ms = 0.f;
    ndt_last_kernel_ms(ndt, &ms);
    lastVerifyMs = ms;
    pairsVerified += n;
    for (int64_t k = 0; k < n; ++k) {
      const KeyFrame &f = frames[cand[k]];
      const ndt_result &r = res[k];
      LoopMatch m;
      m.curId = nodeId; m.refId = f.nodeId; m.result = r;
      const float r00 = r.T[0], r10 = r.T[1];                     // yaw by quadrant, like estimatePose (PoseEstimator.cpp:31-35)
      double theta;
      if (r00 > 0 && r10 > 0) theta = std::asin(r10);
      else if (r00 > 0 && r10 < 0) theta = std::asin(r10);
      else if (r00 < 0 && r10 > 0) theta = std::acos(r00);
      else theta = std::acos(r00) * (-1.0);
      m.relPose.setPose(r.T[12], r.T[13], RAD2DEG(theta));
      m.cost = r.converged ? r.fitness : 10000000;
      Eigen::Matrix3d negH;
      for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) negH(a, b) = -r.hess[3 * a + b];
      m.cov = negH.inverse() * coeNDTCov;
      m.accepted = (m.cost <= scoreThre);
      if (m.accepted && pg && f.nodeId >= 0 && nodeId >= 0 && f.nodeId < static_cast<int>(pg->nodes.size()) &&
          nodeId < static_cast<int>(pg->nodes.size())) {
        PoseArc *arc = pg->makeArc(f.nodeId, nodeId, m.relPose, m.cov);
        arc->loop = true; arc->cost = m.cost;
        pg->addArc(arc);
        found = true;
      }
      lastMatches.push_back(m);
    }
  }
  frames.push_back(std::move(kf));
  return found;
}
