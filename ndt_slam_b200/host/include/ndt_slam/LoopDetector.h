// LoopDetector.h -- loop-closure candidates and their batched NDT verification.
//
// The reference calls lpd.detectLoop(&scan, curPose, cnt) at key frames and, on success, hands over to a back end
// [REF src/FrontEnd.cpp:32-44, all commented out; FrontEnd.h:60-96 for the wiring]; the class itself is not in its tree.
// What is built here is the front half of that pipeline around the batched verifier of the C ABI:
//   1. candidates: earlier key frames whose estimated position lies within `loop_radius` of the current one while the
//      accumulated travel distance in between exceeds `loop_min_travel` (a revisit, not the neighbourhood just driven),
//      at most `loop_max_candidates`, nearest first;
//   2. verification: ONE ndt_match_pairs call -- for every candidate a grid is built from the candidate's stored scan and
//      the current scan is matched against it from the relative pose the current estimates imply (every pair is
//      PoseEstimator::estimatePose [REF src/PoseEstimator.cpp:4-69], all pairs in one launch on the GPU);
//   3. consumption: candidates whose match converged with a fitness score <= `loop_score_thre` become loop arcs of the
//      pose graph (relative pose = the NDT result, covariance = (-H)^-1 * coeNDTCov like estimatePose).
// Scans are kept in their own sensor frame (resampled, as matchScan leaves them), so a pair is scan-to-scan like BASELINE
// config 5. No pose adjustment follows: the arcs are the product.
#ifndef NDT_SLAM_B200_LOOPDETECTOR_H_
#define NDT_SLAM_B200_LOOPDETECTOR_H_

#include <vector>
#include <Eigen/Core>
#include <ros/ros.h>

#include "Pose2D.h"
#include "PoseGraph.h"
#include "Scan2D.h"
#include "ndt_b200.h"

struct LoopMatch {
  int curId = -1, refId = -1;     // key-frame (pose-graph node) ids
  Pose2D relPose;                 // current pose seen from the candidate's pose
  Eigen::Matrix3d cov;
  double cost = 0.0;              // fitness score
  bool accepted = false;
  ndt_result result;              // everything the device reported for this pair
};

class LoopDetector {
 public:
  double radius;            // loop_radius [m]
  double minTravel;         // loop_min_travel [m]
  int maxCandidates;        // loop_max_candidates
  double scoreThre;         // loop_score_thre (fitness score, like ScanMatcher's score_thre)
  double coeNDTCov;

  std::vector<LoopMatch> lastMatches;      // every candidate of the last detectLoop call, accepted or not
  double lastVerifyMs = 0.0;               // device time of the batched verification
  long long pairsVerified = 0;

  LoopDetector();
  ~LoopDetector();
  LoopDetector(const LoopDetector &) = delete;
  LoopDetector &operator=(const LoopDetector &) = delete;

  void setPoseGraph(PoseGraph *pg_) { pg = pg_; }

  // A key frame: remembers the (resampled, sensor-frame) scan under pose-graph node `nodeId` at accumulated travel
  // distance `atd`, then looks for and verifies loop candidates. Returns true if at least one loop arc was added.
  bool detectLoop(const Scan2D *curScan, const Pose2D &curPose, int nodeId, double atd);

  // step 1 on its own (host only; unit-testable without a device)
  std::vector<int> findCandidates(const Pose2D &curPose, double atd) const;

 private:
  struct KeyFrame { int nodeId; Pose2D pose; double atd; std::vector<float> xyzw; };
  std::vector<KeyFrame> frames;
  PoseGraph *pg = nullptr;
  ndt_handle ndt = nullptr;
  double Resolution, StepSize, TransformationEpsilon, LeafSize;
  int MaximumIterations;
  void ensureHandle();
};

#endif
