// PoseGraph.h -- nodes (robot poses) and arcs (relative-pose constraints) of the SLAM pose graph.
// The reference only carries the call sites of this class, commented out [REF src/FrontEnd.cpp:20-30, 52-76:
// pg->addNode(curPose), pg->makeArc(lastNode->nid, curNode->nid, relPose, cov), pg->addArc(arc), pg->nodes];
// CMakeLists.txt:84-87 lists PoseGraph.cpp / LoopDetector.cpp / BackEnd.cpp as removed sources. This is row f3 of
// the structure the loop-closure verifier (LoopDetector.h) records its results in. Pose
// adjustment (the back end) is not built.
#ifndef NDT_SLAM_B200_POSEGRAPH_H_
#define NDT_SLAM_B200_POSEGRAPH_H_

#include <memory>
#include <vector>
#include <Eigen/Core>
#include "Pose2D.h"

struct PoseArc;

struct PoseNode {
  int nid = -1;
  Pose2D pose;
  std::vector<PoseArc *> arcs;
};

struct PoseArc {
  PoseNode *src = nullptr;      // start node
  PoseNode *dst = nullptr;      // end node
  Pose2D relPose;               // dst seen from src (Pose2D::calMotion convention: tx, ty in src's frame, th in degrees)
  Eigen::Matrix3d inf;          // information matrix = inverse covariance of relPose (x, y, yaw [rad])
  bool loop = false;            // odometry arc or loop arc
  double cost = 0.0;            // fitness score of the match that produced a loop arc
};

class PoseGraph {
 public:
  std::vector<PoseNode *> nodes;
  std::vector<PoseArc *> arcs;

  PoseGraph() {}
  PoseGraph(const PoseGraph &) = delete;
  PoseGraph &operator=(const PoseGraph &) = delete;

  PoseNode *addNode(const Pose2D &pose) {
    node_store.emplace_back(new PoseNode());
    PoseNode *n = node_store.back().get();
    n->nid = static_cast<int>(nodes.size());
    n->pose = pose;
    nodes.push_back(n);
    return n;
  }
  // an arc from node srcNid to node dstNid with relative pose relPose and covariance cov (inverted here)
  PoseArc *makeArc(int srcNid, int dstNid, const Pose2D &relPose, const Eigen::Matrix3d &cov) {
    arc_store.emplace_back(new PoseArc());
    PoseArc *a = arc_store.back().get();
    a->src = nodes[srcNid];
    a->dst = nodes[dstNid];
    a->relPose = relPose;
    a->inf = cov.inverse();
    return a;
  }
  void addArc(PoseArc *arc) {
    arc->src->arcs.push_back(arc);
    arc->dst->arcs.push_back(arc);
    arcs.push_back(arc);
  }

 private:
  std::vector<std::unique_ptr<PoseNode>> node_store;
  std::vector<std::unique_ptr<PoseArc>> arc_store;
};

#endif
