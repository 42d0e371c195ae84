// ref_shim.cpp -- C entry points over the REFERENCE'S OWN classes (compiled unmodified from
// /root/reference/src by oracle/Makefile against oracle/stubs + oracle/minipcl). TEST INFRASTRUCTURE.
// Output: oracle/_ref/libndt_slam_ref.so. Nothing here is part of the product.
#include <cstring>
#include <iostream>
#include <vector>

#include "ndt_slam/PointCloudMap.h"
#include "ndt_slam/PoseEstimator.h"
#include "ndt_slam/PoseFuser.h"
#include "ndt_slam/ScanMatcher.h"
#include "ndt_slam/ScanPointResampler.h"
// FrontEnd keeps its ScanMatcher private (include/ndt_slam/FrontEnd.h:24). The shim -- and only the shim: the
// reference sources are compiled untouched -- opens the class so it can zero ScanMatcher::lastCov, which the
// reference declares (ScanMatcher.h:42) and reads (ScanMatcher.cpp:61 / 64) before anything writes it. Every header
// FrontEnd.h pulls in is already included above, so the macro touches nothing but the FrontEnd class itself.
#define private public
#include "ndt_slam/FrontEnd.h"
#undef private
#include "ndt_slam/SlamLauncher.h"

namespace {
struct Quiet {   // PoseFuser::fusePose prints matrices to std::cout unconditionally (PoseFuser.cpp:14-15, 27-28)
  std::streambuf *old = nullptr;
  Quiet() { if (!ros::log_enabled()) old = std::cout.rdbuf(nullptr); }
  ~Quiet() { if (old) std::cout.rdbuf(old); }
};
Scan2D make_scan(int sid, const double pose[3], const double *xy, int64_t n) {
  Scan2D s;
  s.sid = sid;
  s.pose.setPose(pose[0], pose[1], pose[2]);
  s.lps.reserve(n);
  for (int64_t i = 0; i < n; ++i) { LPoint2D lp; lp.setData(sid, xy[2 * i], xy[2 * i + 1]); s.lps.push_back(lp); }
  return s;
}
typedef pcl::NormalDistributionsTransform<pcl::PointXYZ, pcl::PointXYZ> NDT;
struct NdtBox { NDT ndt; pcl::PointCloud<pcl::PointXYZ>::Ptr src, tgt; };
pcl::PointCloud<pcl::PointXYZ>::Ptr cloud_of(const float *xyzw, int64_t n) {
  pcl::PointCloud<pcl::PointXYZ>::Ptr c(new pcl::PointCloud<pcl::PointXYZ>);
  c->points.resize(n); c->width = (uint32_t)n; c->height = 1; c->is_dense = false;
  for (int64_t i = 0; i < n; ++i) { c->points[i].x = xyzw[4 * i]; c->points[i].y = xyzw[4 * i + 1]; c->points[i].z = xyzw[4 * i + 2]; }
  return c;
}
struct Slam {
  PointCloudMap pcmap; FrontEnd fe; PoseEstimator estim;
  Slam() {
    fe.setPoseEstimator(&estim); fe.setPointCloudMap(&pcmap);
    fe.smat.lastCov.setZero();      // uninitialised in the reference: a run is only reproducible with a defined value
  }
};
}  // namespace

extern "C" {

void ref_param_set(const char *k, const char *v) { ros::param::set(k, v); }
void ref_param_clear() { ros::param::store().clear(); }

// ---- ScanPointResampler (src/ScanPointResampler.cpp) ------------------------------------------------
int64_t ref_resample(const double *xy, int64_t n, double *out, int64_t cap) {
  const double pose[3] = {0, 0, 0};
  Scan2D s = make_scan(0, pose, xy, n);
  ScanPointResampler r;
  r.resamplePoints(&s);
  const int64_t m = (int64_t)s.lps.size();
  if (m > cap) return -m;
  for (int64_t i = 0; i < m; ++i) { out[2 * i] = s.lps[i].x; out[2 * i + 1] = s.lps[i].y; }
  return m;
}

// ---- MyUtil / Pose2D (src/MyUtil.cpp, src/Pose2D.cpp) ------------------------------------------------
double ref_add_angle(double a, double b) { return MyUtil::add_angle(a, b); }
double ref_sub_angle(double a, double b) { return MyUtil::sub_angle(a, b); }
void ref_cal_motion(const double cur[3], const double prev[3], double m[3]) {
  Pose2D c(cur[0], cur[1], cur[2]), p(prev[0], prev[1], prev[2]), o;
  Pose2D::calMotion(c, p, o); m[0] = o.tx; m[1] = o.ty; m[2] = o.th;
}
void ref_cal_pred_pose(const double motion[3], const double last[3], double pred[3]) {
  Pose2D m(motion[0], motion[1], motion[2]), l(last[0], last[1], last[2]), o;
  Pose2D::calPredPose(m, l, o); pred[0] = o.tx; pred[1] = o.ty; pred[2] = o.th;
}

// ---- PoseFuser (src/PoseFuser.cpp) -------------------------------------------------------------------
static Eigen::Matrix3d m3(const double *a) { Eigen::Matrix3d m; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) m(i, j) = a[3 * i + j]; return m; }
static void m3out(const Eigen::Matrix3d &m, double *a) { for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) a[3 * i + j] = m(i, j); }
void ref_odometry_cov(const double motion[3], const double last[3], const double lastCov[9], double cov[9]) {
  PoseFuser f; Eigen::Matrix3d c;
  f.calOdometryCovariance(Pose2D(motion[0], motion[1], motion[2]), Pose2D(last[0], last[1], last[2]), m3(lastCov), c);
  m3out(c, cov);
}
void ref_fuse_pose(const double pred[3], const double est[3], const double motion[3], const double last[3],
                   const double lastCov[9], const double Q[9], double fused[3], double cov[9]) {
  Quiet q;
  PoseFuser f; Eigen::Matrix3d c; Pose2D out;
  f.fusePose(Pose2D(pred[0], pred[1], pred[2]), Pose2D(est[0], est[1], est[2]), Pose2D(motion[0], motion[1], motion[2]),
             Pose2D(last[0], last[1], last[2]), m3(lastCov), m3(Q), out, c);
  fused[0] = out.tx; fused[1] = out.ty; fused[2] = out.th; m3out(c, cov);
}

// ---- PoseEstimator::estimatePose (src/PoseEstimator.cpp) on a scan + target cloud --------------------
double ref_estimate_pose(const double *scan_xy, int64_t n, const float *tgt_xyzw, int64_t m, const double init[3],
                         double est[3], double cov[9]) {
  Quiet q;
  Scan2D s = make_scan(0, init, scan_xy, n);
  PoseEstimator pe;
  pe.setScanPair(&s, cloud_of(tgt_xyzw, m));
  Pose2D ip(init[0], init[1], init[2]), ep; Eigen::Matrix3d c;
  const double cost = pe.estimatePose(ip, ep, c);
  est[0] = ep.tx; est[1] = ep.ty; est[2] = ep.th; m3out(c, cov);
  return cost;
}

// ---- the restated PCL objects directly (parity of the two restatements) ------------------------------
void *ref_ndt_create(float resolution, double step, double eps, int max_iter) {
  NdtBox *b = new NdtBox();
  b->ndt.setTransformationEpsilon(eps); b->ndt.setStepSize(step); b->ndt.setResolution(resolution); b->ndt.setMaximumIterations(max_iter);
  return b;
}
void ref_ndt_destroy(void *h) { delete (NdtBox *)h; }
void ref_ndt_set_target(void *h, const float *xyzw, int64_t n) { NdtBox *b = (NdtBox *)h; b->tgt = cloud_of(xyzw, n); b->ndt.setInputTarget(b->tgt); }
void ref_ndt_set_source(void *h, const float *xyzw, int64_t n) { NdtBox *b = (NdtBox *)h; b->src = cloud_of(xyzw, n); b->ndt.setInputSource(b->src); }
// leaves in key order: cell index, nr_points, mean xy, icov (xx, xy, yx, yy), centroid xy
int64_t ref_ndt_grid(void *h, int64_t cap, int32_t *cell, int32_t *nr, double *mean2, double *icov4, float *cen2, int32_t dims[4]) {
  NdtBox *b = (NdtBox *)h;
  const auto &g = b->ndt.getTargetCells();
  const auto mb = g.getMinBoxCoordinates(), dv = g.getNrDivisions();
  if (dims) { dims[0] = mb[0]; dims[1] = mb[1]; dims[2] = dv[0]; dims[3] = dv[1]; }
  int64_t k = 0;
  for (const auto &kv : g.getLeaves()) {
    if (k >= cap) break;
    const auto &l = kv.second;
    if (cell) cell[k] = (int32_t)kv.first;
    if (nr) nr[k] = l.nr_points;
    if (mean2) { mean2[2 * k] = l.mean_[0]; mean2[2 * k + 1] = l.mean_[1]; }
    if (icov4) { icov4[4 * k] = l.icov_(0, 0); icov4[4 * k + 1] = l.icov_(0, 1); icov4[4 * k + 2] = l.icov_(1, 0); icov4[4 * k + 3] = l.icov_(1, 1); }
    if (cen2) { cen2[2 * k] = l.centroid[0]; cen2[2 * k + 1] = l.centroid[1]; }
    ++k;
  }
  return (int64_t)g.getLeaves().size();
}
// score, gradient (x, y, yaw), Hessian rows/cols {0, 1, 5}; returns max |entry| of everything outside that block
double ref_ndt_eval(void *h, const double pose[3], int want_hessian, double out13[13]) {
  NdtBox *b = (NdtBox *)h;
  NDT::Vector6d p, g; NDT::Matrix6d H;
  p << pose[0], pose[1], 0, 0, 0, pose[2];
  out13[0] = b->ndt.evaluate(p, want_hessian != 0, g, H);
  const int id[3] = {0, 1, 5};
  for (int i = 0; i < 3; ++i) out13[1 + i] = g(id[i]);
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) out13[4 + 3 * i + j] = H(id[i], id[j]);
  double off = 0;
  const int od[3] = {2, 3, 4};
  for (int i = 0; i < 3; ++i) { off = std::max(off, std::fabs(g(od[i]))); for (int j = 0; j < 3; ++j) { off = std::max(off, std::fabs(H(od[i], id[j]))); off = std::max(off, std::fabs(H(id[j], od[i]))); } }
  return off;
}
// align from guess (x, y, yaw rad): out = final p (x, y, yaw), score, iterations, converged, passes, fitness, H3x3
static void ndt_align_impl(void *h, const double guess[3], double out[17], bool want_fitness);
void ref_ndt_align(void *h, const double guess[3], double out[17]) { ndt_align_impl(h, guess, out, true); }
// the same without Registration::getFitnessScore (the batched device call ranks relocalisation hypotheses by score only)
void ref_ndt_align_nofit(void *h, const double guess[3], double out[17]) { ndt_align_impl(h, guess, out, false); }
static void ndt_align_impl(void *h, const double guess[3], double out[17], bool want_fitness) {
  NdtBox *b = (NdtBox *)h;
  const float yaw = (float)guess[2];
  Eigen::Matrix4f G = Eigen::Matrix4f::Identity();
  const float c = (float)std::cos((double)yaw), s = (float)std::sin((double)yaw);
  G(0, 0) = c; G(0, 1) = -s; G(1, 0) = s; G(1, 1) = c; G(0, 3) = (float)guess[0]; G(1, 3) = (float)guess[1];
  pcl::PointCloud<pcl::PointXYZ> outc;
  b->ndt.align(outc, G);
  out[0] = b->ndt.final_p_(0); out[1] = b->ndt.final_p_(1); out[2] = b->ndt.final_p_(5);
  out[3] = b->ndt.final_score_; out[4] = b->ndt.getFinalNumIteration(); out[5] = b->ndt.hasConverged() ? 1 : 0;
  out[6] = b->ndt.objectivePasses(); out[7] = want_fitness ? b->ndt.getFitnessScore() : std::nan("");
  const int id[3] = {0, 1, 5};
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) out[8 + 3 * i + j] = b->ndt.final_hessian_(id[i], id[j]);
}
int64_t ref_voxel_filter(const float *xyzw, int64_t n, float leaf, float *out) {
  pcl::PointCloud<pcl::PointXYZ>::Ptr in = cloud_of(xyzw, n);
  pcl::PointCloud<pcl::PointXYZ> f;
  pcl::ApproximateVoxelGrid<pcl::PointXYZ> vg;
  vg.setLeafSize(leaf, leaf, leaf); vg.setInputCloud(in); vg.filter(f);
  for (size_t i = 0; i < f.points.size(); ++i) { out[4 * i] = f.points[i].x; out[4 * i + 1] = f.points[i].y; out[4 * i + 2] = f.points[i].z; out[4 * i + 3] = 0.f; }
  return (int64_t)f.points.size();
}

// ---- the reference front end end to end: SlamLauncher::init wiring + FrontEnd::process per scan -------
void *ref_slam_create() { Quiet q; return new Slam(); }
void ref_slam_destroy(void *h) { delete (Slam *)h; }
void ref_slam_process(void *h, int sid, const double odo[3], const double *xy, int64_t n) {
  Quiet q;
  Slam *s = (Slam *)h;
  Scan2D scan = make_scan(sid, odo, xy, n);
  s->fe.process(scan);
}
int64_t ref_slam_poses(void *h, double *out3, int64_t cap) {
  Slam *s = (Slam *)h;
  std::vector<Pose2D> p = s->fe.get_poses();
  const int64_t m = std::min<int64_t>(cap, (int64_t)p.size());
  for (int64_t i = 0; i < m; ++i) { out3[3 * i] = p[i].tx; out3[3 * i + 1] = p[i].ty; out3[3 * i + 2] = p[i].th; }
  return (int64_t)p.size();
}
// per-scan pose covariances (ScanMatcher::Covs, public member; reached through the opened FrontEnd)
int64_t ref_slam_covs(void *h, double *out9, int64_t cap) {
  Slam *s = (Slam *)h;
  const std::vector<Eigen::Matrix3d> &c = s->fe.smat.Covs;
  const int64_t m = std::min<int64_t>(cap, (int64_t)c.size());
  for (int64_t k = 0; k < m; ++k) for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) out9[9 * k + 3 * i + j] = c[k](i, j);
  return (int64_t)c.size();
}
int64_t ref_slam_local_map(void *h, float *xyzw, int64_t cap) {
  Slam *s = (Slam *)h;
  const auto &c = *s->pcmap.localMap_cloud;
  const int64_t m = std::min<int64_t>(cap, (int64_t)c.points.size());
  for (int64_t i = 0; i < m; ++i) { xyzw[4 * i] = c.points[i].x; xyzw[4 * i + 1] = c.points[i].y; xyzw[4 * i + 2] = c.points[i].z; xyzw[4 * i + 3] = 0.f; }
  return (int64_t)c.points.size();
}
int64_t ref_slam_global_map(void *h, float *xyzw, int64_t cap) {
  Slam *s = (Slam *)h;
  const auto &c = *s->pcmap.globalMap_cloud;
  const int64_t m = std::min<int64_t>(cap, (int64_t)c.points.size());
  for (int64_t i = 0; i < m; ++i) { xyzw[4 * i] = c.points[i].x; xyzw[4 * i + 1] = c.points[i].y; xyzw[4 * i + 2] = c.points[i].z; xyzw[4 * i + 3] = 0.f; }
  return (int64_t)c.points.size();
}
int ref_slam_submaps(void *h) { return (int)((Slam *)h)->pcmap.submaps.size(); }

// ---- SlamLauncher (src/SlamLauncher.cpp, unmodified): the text scan-log reader and the poses writer -------------------
// Parses the log named by the filename_in parameter exactly like loop_wait does (readFormat, then input_file_line until it
// reports the end); returns the records: per scan (sid, pose x y th_deg, n points) in meta5 and all points in xy.
int64_t ref_launcher_parse(double *meta5, int64_t meta_cap, double *xy, int64_t xy_cap, int64_t *n_points) {
  Quiet q;
  SlamLauncher sl;
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
// output_file_poses on the given trajectory, into the file named by the poses_name parameter
void ref_launcher_write_poses(const double *poses3, int64_t n) {
  Quiet q;
  std::vector<Pose2D> poses;
  for (int64_t i = 0; i < n; ++i) poses.push_back(Pose2D(poses3[3 * i], poses3[3 * i + 1], poses3[3 * i + 2]));
  SlamLauncher sl;
  sl.output_file_poses(poses);
  sl.outputfile.close();
}
// PointCloudMap::saveGlobalMap (PointCloudMap.h:124-136) with the given clouds as global map / sub-maps
void ref_save_maps(const float *global_xyzw, int64_t n_global, const float *sub_xyzw, const int64_t *sub_off, int n_sub) {
  Quiet q;
  PointCloudMap pcmap;
  pcmap.globalMap_cloud = cloud_of(global_xyzw, n_global);
  for (int k = 0; k < n_sub; ++k) pcmap.maps.push_back(cloud_of(sub_xyzw + 4 * sub_off[k], sub_off[k + 1] - sub_off[k]));
  pcmap.saveGlobalMap();
}

// The reference's own PCFilter (include/ndt_slam/PCFilter.h, unmodified) on the restated change-detector octree:
// diff = difference_extraction(base, test), kept = remove_neighborPoint(test, diff)
void ref_pcfilter(const float *base_xyzw, int64_t n_base, const float *test_xyzw, int64_t n_test, float *diff_out, int64_t *n_diff,
                  float *kept_out, int64_t *n_kept) {
  PCFilter f;
  pcl::PointCloud<pcl::PointXYZ>::Ptr base = cloud_of(base_xyzw, n_base), test = cloud_of(test_xyzw, n_test);
  pcl::PointCloud<pcl::PointXYZ>::Ptr diff = f.difference_extraction(base, test);
  pcl::PointCloud<pcl::PointXYZ>::Ptr kept = f.remove_neighborPoint(test, diff);
  auto dump = [](const pcl::PointCloud<pcl::PointXYZ> &c, float *o) {
    for (size_t i = 0; i < c.points.size(); ++i) { o[4 * i] = c.points[i].x; o[4 * i + 1] = c.points[i].y; o[4 * i + 2] = c.points[i].z; o[4 * i + 3] = 0.f; }
    return (int64_t)c.points.size();
  };
  *n_diff = dump(*diff, diff_out);
  *n_kept = dump(*kept, kept_out);
}

// The reference's own PointCloudMap on its own: same driver as host_map_replay in the product's harness.
int64_t ref_map_replay(const double *poses3, const double *xy, const int64_t *off, int n_scans,
                       float *local_out, int64_t lcap, int64_t *n_local, float *global_out, int64_t gcap, int64_t *n_global) {
  Quiet q;
  PointCloudMap pcmap;
  for (int s = 0; s < n_scans; ++s) {
    pcmap.addPose(Pose2D(poses3[3 * s], poses3[3 * s + 1], poses3[3 * s + 2]));
    std::vector<LPoint2D> lps(off[s + 1] - off[s]);
    for (int64_t i = off[s]; i < off[s + 1]; ++i) lps[i - off[s]].setData(s, xy[2 * i], xy[2 * i + 1]);
    pcmap.addPoints(lps);
    pcmap.makeLocalMap();
  }
  pcmap.makeGlobalMap();
  auto dump = [](const pcl::PointCloud<pcl::PointXYZ> &c, float *o, int64_t cap) {
    const int64_t m = std::min<int64_t>(cap, (int64_t)c.points.size());
    for (int64_t i = 0; i < m; ++i) { o[4 * i] = c.points[i].x; o[4 * i + 1] = c.points[i].y; o[4 * i + 2] = c.points[i].z; o[4 * i + 3] = 0.f; }
    return (int64_t)c.points.size();
  };
  *n_local = dump(*pcmap.localMap_cloud, local_out, lcap);
  *n_global = dump(*pcmap.globalMap_cloud, global_out, gcap);
  return (int64_t)pcmap.submaps.size();
}

}  // extern "C"
