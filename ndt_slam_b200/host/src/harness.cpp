// harness.cpp -- extern "C" test / bench entry points over the host classes (libndt_slam_host.so).
// Mirrors oracle/ref_shim.cpp function for function so the same Python driver can run the reference
// build and this build side by side.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#include "ndt_slam/FrontEnd.h"
#include "ndt_slam/PointCloudMap.h"
#include "ndt_slam/PoseEstimator.h"
#include "ndt_slam/PoseFuser.h"
#include "ndt_slam/Relocalizer.h"
#include "ndt_slam/ScanMatcher.h"
#include "ndt_slam/ScanPointResampler.h"
#include "ndt_slam/SlamLauncher.h"
#include "ndt_slam/VoxelFilter.h"

namespace {
Scan2D make_scan(int sid, const double pose[3], const double *xy, int64_t n) {
  Scan2D s;
  s.sid = sid;
  s.pose.setPose(pose[0], pose[1], pose[2]);
  s.lps.resize(n);
  for (int64_t i = 0; i < n; ++i) s.lps[i].setData(sid, xy[2 * i], xy[2 * i + 1]);
  return s;
}
Eigen::Matrix3d m3(const double *a) { Eigen::Matrix3d m; for (int i = 0; i < 9; ++i) m(i / 3, i % 3) = a[i]; return m; }
void m3out(const Eigen::Matrix3d &m, double *a) { for (int i = 0; i < 9; ++i) a[i] = m(i / 3, i % 3); }
pcl::PointCloud<pcl::PointXYZ>::Ptr cloud_of(const float *xyzw, int64_t n) {
  auto c = std::make_shared<pcl::PointCloud<pcl::PointXYZ>>();
  c->points.resize(n);
  if (n) std::memcpy(c->points.data(), xyzw, sizeof(float) * 4 * n);
  c->width = (uint32_t)n; c->height = 1; c->is_dense = false;
  return c;
}
int64_t cloud_out(const pcl::PointCloud<pcl::PointXYZ> &c, float *xyzw, int64_t cap) {
  const int64_t m = std::min<int64_t>(cap, (int64_t)c.points.size());
  for (int64_t i = 0; i < m; ++i) { xyzw[4 * i] = c.points[i].x; xyzw[4 * i + 1] = c.points[i].y; xyzw[4 * i + 2] = c.points[i].z; xyzw[4 * i + 3] = 0.f; }
  return (int64_t)c.points.size();
}
struct Slam {
  PointCloudMap pcmap; FrontEnd fe; PoseEstimator estim;
  double ms[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};   // resample, estimate, fuse, growMap, device grid, device match, host filter, set_source wall, set_target wall, align wall
  int64_t evals = 0, point_evals = 0, matches = 0;
  Slam() { fe.setPoseEstimator(&estim); fe.setPointCloudMap(&pcmap); }
};
char g_err[512] = "";
}  // namespace

extern "C" {

const char *host_last_error() { return g_err; }
void host_param_set(const char *k, const char *v) { ros::param::set(k, v); }
void host_param_clear() { ros::param::clear(); }

int64_t host_resample(const double *xy, int64_t n, double *out, int64_t cap) {
  const double zero[3] = {0, 0, 0};
  Scan2D s = make_scan(0, zero, xy, n);
  ScanPointResampler r;
  r.resamplePoints(&s);
  const int64_t m = (int64_t)s.lps.size();
  if (m > cap) return -m;
  for (int64_t i = 0; i < m; ++i) { out[2 * i] = s.lps[i].x; out[2 * i + 1] = s.lps[i].y; }
  return m;
}
int64_t host_voxel_filter(const float *xyzw, int64_t n, float leaf, float *out) {
  pcl::PointCloud<pcl::PointXYZ> f;
  ndt_host::approximate_voxel_grid(*cloud_of(xyzw, n), leaf, f);
  return cloud_out(f, out, n);
}
double host_add_angle(double a, double b) { return MyUtil::add_angle(a, b); }
double host_sub_angle(double a, double b) { return MyUtil::sub_angle(a, b); }
void host_cal_motion(const double cur[3], const double prev[3], double m[3]) {
  Pose2D o; Pose2D::calMotion(Pose2D(cur[0], cur[1], cur[2]), Pose2D(prev[0], prev[1], prev[2]), o);
  m[0] = o.tx; m[1] = o.ty; m[2] = o.th;
}
void host_cal_pred_pose(const double motion[3], const double last[3], double pred[3]) {
  Pose2D o; Pose2D::calPredPose(Pose2D(motion[0], motion[1], motion[2]), Pose2D(last[0], last[1], last[2]), o);
  pred[0] = o.tx; pred[1] = o.ty; pred[2] = o.th;
}
void host_odometry_cov(const double motion[3], const double last[3], const double lastCov[9], double cov[9]) {
  PoseFuser f; Eigen::Matrix3d c;
  f.calOdometryCovariance(Pose2D(motion[0], motion[1], motion[2]), Pose2D(last[0], last[1], last[2]), m3(lastCov), c);
  m3out(c, cov);
}
void host_fuse_pose(const double pred[3], const double est[3], const double motion[3], const double last[3],
                    const double lastCov[9], const double Q[9], double fused[3], double cov[9]) {
  PoseFuser f; Eigen::Matrix3d c; Pose2D out;
  f.fusePose(Pose2D(pred[0], pred[1], pred[2]), Pose2D(est[0], est[1], est[2]), Pose2D(motion[0], motion[1], motion[2]),
             Pose2D(last[0], last[1], last[2]), m3(lastCov), m3(Q), out, c);
  fused[0] = out.tx; fused[1] = out.ty; fused[2] = out.th; m3out(c, cov);
}

// PoseEstimator::setScanPair + estimatePose on the GPU. Returns the cost; -1e300 on failure (see host_last_error).
double host_estimate_pose(const double *scan_xy, int64_t n, const float *tgt_xyzw, int64_t m, const double init[3],
                          double est[3], double cov[9], ndt_result *res_out) {
  try {
    Scan2D s = make_scan(0, init, scan_xy, n);
    PoseEstimator pe;
    pe.setScanPair(&s, cloud_of(tgt_xyzw, m));
    Pose2D ip(init[0], init[1], init[2]), ep; Eigen::Matrix3d c;
    const double cost = pe.estimatePose(ip, ep, c);
    est[0] = ep.tx; est[1] = ep.ty; est[2] = ep.th; m3out(c, cov);
    if (res_out) *res_out = pe.lastResult;
    return cost;
  } catch (const std::exception &e) {
    std::strncpy(g_err, e.what(), sizeof(g_err) - 1);
    return -1e300;
  }
}

void *host_slam_create() { return new Slam(); }
void host_slam_destroy(void *h) { delete (Slam *)h; }
int host_slam_process(void *h, int sid, const double odo[3], const double *xy, int64_t n) {
  Slam *s = (Slam *)h;
  try {
    Scan2D scan = make_scan(sid, odo, xy, n);
    s->fe.process(scan);
    const ScanMatcher &sm = s->fe.matcher();
    s->ms[0] += sm.msResample; s->ms[1] += sm.msEstimate; s->ms[2] += sm.msFuse; s->ms[3] += sm.msGrowMap;
    if (sm.msEstimate > 0) {
      s->ms[4] += s->estim.lastGridMs; s->ms[5] += s->estim.lastMatchMs; s->ms[6] += s->estim.lastFilterMs;
      s->ms[7] += s->estim.lastSetSourceWallMs; s->ms[8] += s->estim.lastSetTargetWallMs; s->ms[9] += s->estim.lastAlignWallMs;
      s->evals += s->estim.lastResult.evals; s->point_evals += s->estim.lastResult.point_evals; s->matches += 1;
    }
    return 0;
  } catch (const std::exception &e) {
    std::strncpy(g_err, e.what(), sizeof(g_err) - 1);
    return -1;
  }
}
// Teacher forcing (tests): process one scan, then let the map / the next fusion continue from the given pose (deg) and
// covariance -- the reference's own outputs for this scan -- instead of this scan's result (see ScanMatcher.h).
int host_slam_process_forced(void *h, int sid, const double odo[3], const double *xy, int64_t n, const double pose3[3], const double cov9[9]) {
  Slam *s = (Slam *)h;
  Pose2D fp(pose3[0], pose3[1], pose3[2]);
  Eigen::Matrix3d fc;
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) fc(i, j) = cov9[3 * i + j];
  ScanMatcher &sm = s->fe.matcher();
  sm.forcePose = &fp; sm.forceCov = &fc;
  const int rc = host_slam_process(h, sid, odo, xy, n);
  sm.forcePose = nullptr; sm.forceCov = nullptr;
  return rc;
}
int64_t host_slam_covs(void *h, double *out9, int64_t cap) {
  const ScanMatcher &sm = ((Slam *)h)->fe.matcher();
  const int64_t m = std::min<int64_t>(cap, (int64_t)sm.Covs.size());
  for (int64_t k = 0; k < m; ++k) for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) out9[9 * k + 3 * i + j] = sm.Covs[k](i, j);
  return (int64_t)sm.Covs.size();
}
int64_t host_slam_poses(void *h, double *out3, int64_t cap) {
  Slam *s = (Slam *)h;
  std::vector<Pose2D> p = s->fe.get_poses();
  const int64_t m = std::min<int64_t>(cap, (int64_t)p.size());
  for (int64_t i = 0; i < m; ++i) { out3[3 * i] = p[i].tx; out3[3 * i + 1] = p[i].ty; out3[3 * i + 2] = p[i].th; }
  return (int64_t)p.size();
}
int64_t host_slam_local_map(void *h, float *xyzw, int64_t cap) { return cloud_out(*((Slam *)h)->pcmap.localMap_cloud, xyzw, cap); }
int64_t host_slam_global_map(void *h, float *xyzw, int64_t cap) { return cloud_out(*((Slam *)h)->pcmap.globalMap_cloud, xyzw, cap); }
int host_slam_submaps(void *h) { return (int)((Slam *)h)->pcmap.submaps.size(); }
// ms: resample, estimate, fuse, growMap, device grid kernels, device match kernel ; counts: matches, evals, point_evals
void host_slam_stats(void *h, double ms10[10], int64_t counts3[3]) {
  Slam *s = (Slam *)h;
  for (int i = 0; i < 10; ++i) ms10[i] = s->ms[i];
  counts3[0] = s->matches; counts3[1] = s->evals; counts3[2] = s->point_evals;
}

// PointCloudMap on its own (no device): feed map-frame scans + poses like ScanMatcher::growMap does
// (addPose, addPoints, makeLocalMap per scan; makeGlobalMap at the end) and return both clouds.
// check_every > 0 additionally compares, every check_every scans, the incrementally maintained local map with one
// rebuilt from scratch by the one-shot filter (returns -(scan+1) at the first difference).
int64_t host_map_replay(const double *poses3, const double *xy, const int64_t *off, int n_scans, int check_every,
                        float *local_out, int64_t lcap, int64_t *n_local, float *global_out, int64_t gcap, int64_t *n_global) {
  PointCloudMap pcmap;
  for (int s = 0; s < n_scans; ++s) {
    pcmap.addPose(Pose2D(poses3[3 * s], poses3[3 * s + 1], poses3[3 * s + 2]));
    std::vector<LPoint2D> lps(off[s + 1] - off[s]);
    for (int64_t i = off[s]; i < off[s + 1]; ++i) lps[i - off[s]].setData(s, xy[2 * i], xy[2 * i + 1]);
    pcmap.addPoints(lps);
    pcmap.makeLocalMap();
    if (s % 5 == 0) pcmap.makeGlobalMap();        // like FrontEnd at key frames: the kept frozen part must not change the final cloud
    if (check_every > 0 && (s + 1) % check_every == 0) {
      pcl::PointCloud<pcl::PointXYZ> expect, thinned;
      if (pcmap.submaps.size() >= 2) expect += *pcmap.submaps[pcmap.submaps.size() - 2].p_cloud;
      const Submap &cur = pcmap.submaps.back();
      ndt_host::approximate_voxel_grid(*cur.p_cloud, static_cast<float>(cur.LeafSize), thinned);
      expect += thinned;
      const auto &got = pcmap.localMap_cloud->points;
      if (got.size() != expect.points.size()) return -(int64_t)(s + 1);
      for (size_t i = 0; i < got.size(); ++i)
        if (std::memcmp(&got[i], &expect.points[i], 12) != 0) return -(int64_t)(s + 1);
    }
  }
  pcmap.makeGlobalMap();
  *n_local = cloud_out(*pcmap.localMap_cloud, local_out, lcap);
  *n_global = cloud_out(*pcmap.globalMap_cloud, global_out, gcap);
  return (int64_t)pcmap.submaps.size();
}

// PCFilter on its own: diff = difference_extraction(base, test), kept = remove_neighborPoint(test, diff)
void host_pcfilter(const float *base_xyzw, int64_t n_base, const float *test_xyzw, int64_t n_test, float *diff_out, int64_t *n_diff,
                   float *kept_out, int64_t *n_kept) {
  PCFilter f;
  auto diff = f.difference_extraction(cloud_of(base_xyzw, n_base), cloud_of(test_xyzw, n_test));
  auto kept = f.remove_neighborPoint(cloud_of(test_xyzw, n_test), diff);
  *n_diff = cloud_out(*diff, diff_out, n_test);
  *n_kept = cloud_out(*kept, kept_out, n_test);
}

// Relocalizer (C++ host, C ABI only): n_handles handles on devices dev[0..n_handles), grid replicated, hypotheses sharded
int64_t host_relocalize(const int *dev, int n_handles, const float *map_xyzw, int64_t n_map, const float *scan_xyzw, int64_t n_scan,
                        const double *hyp, int64_t n_hyp, double resolution, ndt_result *best, ndt_result *results, double *device_ms) {
  try {
    Relocalizer rl(std::vector<int>(dev, dev + n_handles), resolution);
    rl.setMap(*cloud_of(map_xyzw, n_map));
    rl.setScan(*cloud_of(scan_xyzw, n_scan));
    const int64_t bi = rl.relocalize(hyp, n_hyp, best, results);
    if (device_ms) *device_ms = rl.lastDeviceMs;
    return bi;
  } catch (const std::exception &e) {
    std::strncpy(g_err, e.what(), sizeof(g_err) - 1);
    return -2;
  }
}

// loop closure: what the last processed scan's LoopDetector call did (cur node, ref node, relPose x y th_deg, cost, accepted,
// iters, evals, converged) x candidates; totals in counts3 = {pose-graph nodes, arcs, loop arcs}
int64_t host_slam_loops(void *h, double *rows10, int64_t cap, int64_t counts3[3]) {
  Slam *s = (Slam *)h;
  const std::vector<LoopMatch> &lm = s->fe.lpd.lastMatches;
  const int64_t m = std::min<int64_t>(cap, (int64_t)lm.size());
  for (int64_t k = 0; k < m; ++k) {
    double *r = rows10 + 10 * k;
    r[0] = lm[k].curId; r[1] = lm[k].refId; r[2] = lm[k].relPose.tx; r[3] = lm[k].relPose.ty; r[4] = lm[k].relPose.th; r[5] = lm[k].cost;
    r[6] = lm[k].accepted; r[7] = lm[k].result.iters; r[8] = lm[k].result.evals; r[9] = lm[k].result.converged;
  }
  int64_t loops = 0;
  for (const PoseArc *a : s->fe.pg.arcs) loops += a->loop ? 1 : 0;
  counts3[0] = (int64_t)s->fe.pg.nodes.size(); counts3[1] = (int64_t)s->fe.pg.arcs.size(); counts3[2] = loops;
  return (int64_t)lm.size();
}
// every loop arc of the pose graph: src node, dst node, relPose (x, y, th_deg), cost
int64_t host_slam_loop_arcs(void *h, double *rows6, int64_t cap) {
  Slam *s = (Slam *)h;
  int64_t n = 0;
  for (const PoseArc *a : s->fe.pg.arcs) {
    if (!a->loop) continue;
    if (n < cap) { double *r = rows6 + 6 * n; r[0] = a->src->nid; r[1] = a->dst->nid; r[2] = a->relPose.tx; r[3] = a->relPose.ty; r[4] = a->relPose.th; r[5] = a->cost; }
    ++n;
  }
  return n;
}
// key-frame poses (pose-graph nodes)
int64_t host_slam_nodes(void *h, double *rows3, int64_t cap) {
  Slam *s = (Slam *)h;
  const int64_t m = std::min<int64_t>(cap, (int64_t)s->fe.pg.nodes.size());
  for (int64_t k = 0; k < m; ++k) { rows3[3 * k] = s->fe.pg.nodes[k]->pose.tx; rows3[3 * k + 1] = s->fe.pg.nodes[k]->pose.ty; rows3[3 * k + 2] = s->fe.pg.nodes[k]->pose.th; }
  return (int64_t)s->fe.pg.nodes.size();
}
// LoopDetector::findCandidates on its own (host only): key frames at poses3 / atd, query at index n - 1
int64_t host_loop_candidates(const double *poses3, const double *atd, int64_t n, int *out, int64_t cap) {
  LoopDetector ld;
  Scan2D empty;
  const int keep = ld.maxCandidates;
  ld.maxCandidates = 0;                            // store the key frames only: nothing to verify, no device needed
  for (int64_t k = 0; k + 1 < n; ++k) ld.detectLoop(&empty, Pose2D(poses3[3 * k], poses3[3 * k + 1], poses3[3 * k + 2]), -1, atd[k]);
  ld.maxCandidates = keep;
  const std::vector<int> c = ld.findCandidates(Pose2D(poses3[3 * (n - 1)], poses3[3 * (n - 1) + 1], poses3[3 * (n - 1) + 2]), atd[n - 1]);
  for (size_t k = 0; k < c.size() && (int64_t)k < cap; ++k) out[k] = c[k];
  return (int64_t)c.size();
}

// SlamLauncher's reader and writers on their own (no device needed): same drivers as the ref_launcher_* / ref_save_maps
// entry points of oracle/ref_shim.cpp over the reference's own SlamLauncher.cpp
int64_t host_launcher_parse(double *meta5, int64_t meta_cap, double *xy, int64_t xy_cap, int64_t *n_points) {
  SlamLauncher sl;
  if (!sl.ok) return -1;
  sl.readFormat();
  int64_t n = 0, np = 0;
  while (!sl.input_file_line()) {
    if (n < meta_cap) { meta5[5 * n] = sl.scan.sid; meta5[5 * n + 1] = sl.scan.pose.tx; meta5[5 * n + 2] = sl.scan.pose.ty; meta5[5 * n + 3] = sl.scan.pose.th; meta5[5 * n + 4] = (double)sl.scan.lps.size(); }
    for (const LPoint2D &lp : sl.scan.lps) { if (np < xy_cap) { xy[2 * np] = lp.x; xy[2 * np + 1] = lp.y; } ++np; }
    ++n;
  }
  *n_points = np;
  return n;
}
void host_launcher_write_poses(const double *poses3, int64_t n) {
  std::vector<Pose2D> poses;
  for (int64_t i = 0; i < n; ++i) poses.push_back(Pose2D(poses3[3 * i], poses3[3 * i + 1], poses3[3 * i + 2]));
  SlamLauncher sl;
  sl.output_file_poses(poses);
  sl.outputfile.close();
}
void host_save_maps(const float *global_xyzw, int64_t n_global, const float *sub_xyzw, const int64_t *sub_off, int n_sub) {
  PointCloudMap pcmap;
  pcmap.globalMap_cloud = cloud_of(global_xyzw, n_global);
  for (int k = 0; k < n_sub; ++k) pcmap.maps.push_back(cloud_of(sub_xyzw + 4 * sub_off[k], sub_off[k + 1] - sub_off[k]));
  pcmap.saveGlobalMap();
}

// SlamLauncher: run a text scan log end to end (parameters filename_in / poses_name / map_name ... must be set)
int host_launcher_run() {
  try {
    SlamLauncher sl;
    if (!sl.ok) { std::strncpy(g_err, "cannot open input / output file", sizeof(g_err) - 1); return -1; }
    sl.init();
    sl.loop_wait();
    return sl.scansProcessed;
  } catch (const std::exception &e) {
    std::strncpy(g_err, e.what(), sizeof(g_err) - 1);
    return -1;
  }
}

}  // extern "C"
