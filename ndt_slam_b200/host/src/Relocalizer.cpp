// Relocalizer.cpp -- see Relocalizer.h
#include "ndt_slam/Relocalizer.h"

#include <algorithm>
#include <stdexcept>
#include <string>

namespace {
void ck(int rc, ndt_handle h, const char *what) {
  if (rc != NDT_OK) throw std::runtime_error(std::string("Relocalizer: ") + what + ": " + ndt_last_error(h));
}
}  // namespace

Relocalizer::Relocalizer(const std::vector<int> &devices, double resolution, double stepSize, double transEps, int maxIter) {
  if (devices.empty()) throw std::runtime_error("Relocalizer: no device");
  for (int dev : devices) {
    ndt_params prm;
    ndt_params_default(&prm);
    prm.resolution = static_cast<float>(resolution);
    prm.step_size = stepSize; prm.trans_eps = transEps; prm.max_iter = maxIter;
    prm.device = dev;
    ndt_handle h = nullptr;
    if (ndt_create(&prm, &h) != NDT_OK) throw std::runtime_error(std::string("Relocalizer: ndt_create: ") + ndt_last_error(nullptr));
    handles.push_back(h);
    device_of.push_back(dev);
    d_guess.push_back(nullptr); d_res.push_back(nullptr); cap.push_back(0);
  }
}

Relocalizer::~Relocalizer() {
  for (size_t k = 0; k < handles.size(); ++k) {
    if (d_guess[k]) ndt_free(handles[k], d_guess[k]);
    if (d_res[k]) ndt_free(handles[k], d_res[k]);
    ndt_destroy(handles[k]);
  }
}

void Relocalizer::setMap(const pcl::PointCloud<pcl::PointXYZ> &map) {
  ck(ndt_set_target(handles[0], reinterpret_cast<const float *>(map.points.data()), (int64_t)map.points.size(), NDT_MEM_HOST), handles[0],
     "ndt_set_target");
  ck(ndt_replicate_grid(handles.data(), (int)handles.size(), /*flags=*/0), handles[0], "ndt_replicate_grid");
}

void Relocalizer::setScan(const pcl::PointCloud<pcl::PointXYZ> &scan) {
  for (ndt_handle h : handles)
    ck(ndt_set_source(h, reinterpret_cast<const float *>(scan.points.data()), (int64_t)scan.points.size(), NDT_MEM_HOST), h, "ndt_set_source");
}

void Relocalizer::reserve(size_t k, int64_t n) {
  if (n <= cap[k]) return;
  if (d_guess[k]) ndt_free(handles[k], d_guess[k]);
  if (d_res[k]) ndt_free(handles[k], d_res[k]);
  d_guess[k] = d_res[k] = nullptr;
  ck(ndt_alloc(handles[k], n * 3 * (int64_t)sizeof(double), &d_guess[k]), handles[k], "ndt_alloc");
  ck(ndt_alloc(handles[k], n * (int64_t)sizeof(ndt_result), &d_res[k]), handles[k], "ndt_alloc");
  cap[k] = n;
}

int64_t Relocalizer::relocalize(const double *hyp, int64_t n, ndt_result *best, ndt_result *results) {
  const int64_t G = (int64_t)handles.size();
  std::vector<int64_t> lo(G + 1);
  for (int64_t k = 0; k <= G; ++k) lo[k] = k * n / G;                     // block partition [k n / G, (k + 1) n / G)
  // launch every shard (device-space calls return without waiting), then wait for all of them
  for (int64_t k = 0; k < G; ++k) {
    const int64_t m = lo[k + 1] - lo[k];
    if (m == 0) continue;
    reserve((size_t)k, m);
    ck(ndt_upload(handles[k], d_guess[k], hyp + 3 * lo[k], m * 3 * (int64_t)sizeof(double)), handles[k], "ndt_upload");
    ck(ndt_align_batch(handles[k], static_cast<const double *>(d_guess[k]), m, NDT_MEM_DEVICE, /*want_fitness=*/0,
                       static_cast<ndt_result *>(d_res[k])), handles[k], "ndt_align_batch");
  }
  lastDeviceMs = 0.0;
  std::vector<const ndt_result *> ptrs(G);
  std::vector<int64_t> counts(G);
  for (int64_t k = 0; k < G; ++k) {
    ptrs[k] = static_cast<const ndt_result *>(d_res[k]);
    counts[k] = lo[k + 1] - lo[k];
    if (counts[k] == 0) continue;
    ck(ndt_synchronize(handles[k]), handles[k], "ndt_synchronize");
    float ms = 0.f;
    ndt_last_kernel_ms(handles[k], &ms);
    lastDeviceMs = std::max(lastDeviceMs, (double)ms);
    if (results) ck(ndt_download(handles[k], results + lo[k], d_res[k], counts[k] * (int64_t)sizeof(ndt_result)), handles[k], "ndt_download");
  }
  int bh = -1;
  int64_t bi = -1;
  ck(ndt_best_of_multi(handles.data(), ptrs.data(), counts.data(), (int)G, &bh, &bi, best), handles[0], "ndt_best_of_multi");
  return bh < 0 ? -1 : lo[bh] + bi;
}
