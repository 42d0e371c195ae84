// capi.cu -- the extern "C" boundary declared in include/ndt_b200.h.
// Each entry point names the reference call it replaces (see the header). No torch types, no
// exceptions, no CPU fallback: without a CUDA device ndt_create fails with NDT_ERR_NO_DEVICE.
#include "ndt_host.h"

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>

namespace ndt {

static thread_local std::string g_create_err;

int set_err(Handle *h, int code, const char *what, cudaError_t e) {
  std::string m = what ? what : "";
  if (e != cudaSuccess) { m += ": "; m += cudaGetErrorString(e); (void)cudaGetLastError(); }
  if (h) h->err = m; else g_create_err = m;
  return code;
}

int ensure_pinned(Handle *h, size_t bytes) {
  if (bytes <= h->pinned_cap) return 0;
  if (h->pinned) cudaFreeHost(h->pinned);
  h->pinned = nullptr; h->pinned_cap = 0;
  size_t want = h->pinned_cap == 0 ? ((bytes + 4095) & ~size_t(4095)) : (bytes + bytes / 2 + 4096);
  cudaError_t e = cudaMallocHost(&h->pinned, want);
  if (e != cudaSuccess) { set_err(h, NDT_ERR_CUDA, "cudaMallocHost", e); return 1; }
  h->pinned_cap = want;
  return 0;
}

// header of the flat grid blob used for replication
struct BlobHeader {
  uint64_t magic;
  int32_t flags, reserved;
  GridDims gd;
  int32_t counters[CTR_COUNT];
  int64_t off_slot, off_cen, off_occ, off_recs, off_leaf_id, off_leaf_range, off_sorted, off_tgt, off_nn_range, off_nn_pts, total;
};
static constexpr uint64_t kBlobMagic = 0x4e44544232303045ull;  // "NDTB200E"

static inline int64_t align256(int64_t v) { return (v + 255) & ~int64_t(255); }

// byte sizes of the blob's sections for a grid of the given geometry / counters
struct BlobSizes { int64_t slot, cen, occ, recs, leaf_id, leaf_range, sorted, tgt, nn_range, nn_pts; };
static BlobSizes blob_sizes(const GridDims &gd, const int32_t *counters, int flags) {
  const int64_t nc = gd.n_cells, nl = counters[CTR_LEAVES], nsl = counters[CTR_SLOTS], nt = gd.n_tgt;
  const int64_t npad = nc > 0 ? (int64_t)(gd.div_x + 4) * (gd.div_y + 4) : 0;
  const bool pts = (flags & NDT_BLOB_POINTS) != 0;
  const int64_t ncf = (pts && gd.nn_f > 0) ? (int64_t)gd.nn_div_x * gd.nn_div_y : 0;
  BlobSizes z{};
  z.slot = npad * 4; z.cen = npad * 8; z.occ = npad > 0 ? ((npad + 31) / 32 + 1) * 4 : 0;
  z.recs = nsl * (int64_t)sizeof(CellRec);
  z.leaf_id = pts ? npad * 4 : 0; z.leaf_range = pts ? nl * 8 : 0; z.sorted = pts ? nt * 8 : 0;
  z.tgt = pts ? nt * (int64_t)sizeof(float4) : 0;
  z.nn_range = ncf * 8; z.nn_pts = ncf > 0 ? nt * 8 : 0;
  return z;
}

static BlobHeader blob_layout(const Handle *h, int flags) {
  BlobHeader b{};
  b.magic = kBlobMagic;
  b.flags = flags & NDT_BLOB_POINTS;
  if (!h->grid_has_points || h->inc_active) b.flags = 0;   // a replica without points cannot hand any on; after an incremental
                                                          // update the ordered 1-NN buckets are stale (only the lattice is kept)
  b.gd = h->gd;
  std::memcpy(b.counters, h->h_counters, sizeof(b.counters));
  const BlobSizes z = blob_sizes(h->gd, h->h_counters, b.flags);
  int64_t o = align256(sizeof(BlobHeader));
  b.off_slot = o; o = align256(o + z.slot);
  b.off_cen = o; o = align256(o + z.cen);
  b.off_occ = o; o = align256(o + z.occ);
  b.off_recs = o; o = align256(o + z.recs);
  b.off_leaf_id = o; o = align256(o + z.leaf_id);
  b.off_leaf_range = o; o = align256(o + z.leaf_range);
  b.off_sorted = o; o = align256(o + z.sorted);
  b.off_tgt = o; o = align256(o + z.tgt);
  b.off_nn_range = o; o = align256(o + z.nn_range);
  b.off_nn_pts = o; o = align256(o + z.nn_pts);
  b.total = o;
  return b;
}

}  // namespace ndt

using namespace ndt;

#define H_OR_FAIL(hh)                                     \
  Handle *h = reinterpret_cast<Handle *>(hh);             \
  if (!h) return NDT_ERR_ARG;                             \
  if (cudaSetDevice(h->device) != cudaSuccess) return set_err(h, NDT_ERR_CUDA, "cudaSetDevice"); \
  if (h->pending) { if (int rc_pending = finish_pending(h)) return rc_pending; }

extern "C" {

const char *ndt_version(void) { return "ndt_b200 0.1.0 (sm_100a)"; }

int ndt_params_default(ndt_params *p) {
  if (!p) return NDT_ERR_ARG;
  // C++ defaults of PoseEstimator (include/ndt_slam/PoseEstimator.h:63-64) + PCL's internal constants
  p->resolution = 1.0f;
  p->step_size = 0.1;
  p->trans_eps = 0.01;
  p->max_iter = 35;
  p->outlier_ratio = 0.55;
  p->min_points = 6;
  p->eig_mult = 0.01;
  p->quirks = NDT_QUIRKS_PCL_1_10;
  p->device = 0;
  p->stream = nullptr;
  p->align_skip_fitness = 0;
  p->pairs_schedule = NDT_PAIRS_AUTO;
  p->pairs_batch_points = 0;
  p->align_team = 0;
  p->reserved0 = 0;
  return NDT_OK;
}

static void all_buffers(Handle *h, std::vector<DevBuf *> &v) {
  GridBuffers &g = h->gb;
  v = {&g.tgt, &g.cell_of, &g.rank_of, &g.list, &g.sorted_idx, &g.slot, &g.leaf_id, &g.leaf_cell,
       &g.leaf_n, &g.leaf_start, &g.leaf_nr, &g.leaf_mean, &g.leaf_icov, &g.leaf_cen, &g.recs,
       &g.counters, &g.leaf_pair, &g.big_list, &g.tile_hist, &g.dims, &g.pair_off, &g.cen, &g.occ, &g.nbr, &g.inc_acc, &g.inc_status, &g.inc_mark, &g.inc_lists, &g.inc_cellof, &g.inc_lid, &g.inc_cnt, &g.nn_cnt, &g.nn_range, &g.nn_pts,
       &g.tgt_sorted, &g.leaf_range, &h->src, &h->scratch, &h->scratch2, &h->pair_tgt_next, &h->pair_raw_next, &h->stage, &h->io};
}

int ndt_create(const ndt_params *p, ndt_handle *out) {
  if (!p || !out) return NDT_ERR_ARG;
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0) {
    set_err(nullptr, NDT_ERR_NO_DEVICE, "no CUDA device available (this library has no CPU fallback)", e);
    return NDT_ERR_NO_DEVICE;
  }
  if (p->device < 0 || p->device >= ndev) { set_err(nullptr, NDT_ERR_ARG, "bad device ordinal"); return NDT_ERR_ARG; }
  if (!(p->resolution > 0.f)) { set_err(nullptr, NDT_ERR_ARG, "resolution must be > 0"); return NDT_ERR_ARG; }
  if (cudaSetDevice(p->device) != cudaSuccess) { set_err(nullptr, NDT_ERR_CUDA, "cudaSetDevice"); return NDT_ERR_CUDA; }
  Handle *h = new Handle();
  h->prm = *p;
  h->device = p->device;
  if (p->stream) { h->stream = (cudaStream_t)p->stream; h->own_stream = false; }
  else {
    if ((e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess) {
      set_err(nullptr, NDT_ERR_CUDA, "cudaStreamCreate", e); delete h; return NDT_ERR_CUDA;
    }
    h->own_stream = true;
  }
  cudaEventCreate(&h->ev0); cudaEventCreate(&h->ev1);
  cudaMallocHost((void **)&h->pinned_ctr, 256);
  {
    // private stream-ordered pool that keeps what it is given back (no trimming at synchronisation points)
    int pools = 0;
    cudaDeviceGetAttribute(&pools, cudaDevAttrMemoryPoolsSupported, h->device);
    if (pools) {
      cudaMemPoolProps props{};
      props.allocType = cudaMemAllocationTypePinned;
      props.handleTypes = cudaMemHandleTypeNone;
      props.location.type = cudaMemLocationTypeDevice;
      props.location.id = h->device;
      if (cudaMemPoolCreate(&h->pool, &props) == cudaSuccess) {
        uint64_t keep_all = UINT64_MAX;
        cudaMemPoolSetAttribute(h->pool, cudaMemPoolAttrReleaseThreshold, &keep_all);
      } else { h->pool = nullptr; (void)cudaGetLastError(); }
    }
    std::vector<DevBuf *> bufs;
    all_buffers(h, bufs);
    for (DevBuf *b : bufs) b->bind(h->stream, h->pool);
  }
  cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, h->device);
  cudaDeviceGetAttribute(&h->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device);
  { int coop = 0; cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, h->device); h->coop_launch = coop != 0; }
  if ((e = h->gb.counters.reserve((CTR_COUNT + 4) * sizeof(int32_t))) != cudaSuccess ||
      (e = h->stage.reserve(4096)) != cudaSuccess) {
    set_err(nullptr, NDT_ERR_CUDA, "cudaMalloc", e); delete h; return NDT_ERR_CUDA;
  }
  *out = reinterpret_cast<ndt_handle>(h);
  return NDT_OK;
}

int ndt_destroy(ndt_handle hh) {
  Handle *h = reinterpret_cast<Handle *>(hh);
  if (!h) return NDT_OK;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  std::vector<DevBuf *> bufs;
  all_buffers(h, bufs);
  for (DevBuf *b : bufs) b->release();
  cudaStreamSynchronize(h->stream);
  if (h->pool) cudaMemPoolDestroy(h->pool);
  if (h->pinned) cudaFreeHost(h->pinned);
  if (h->pinned_ctr) cudaFreeHost(h->pinned_ctr);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  for (int k = 0; k < 2; ++k) {
    if (h->ev_up[k]) cudaEventDestroy(h->ev_up[k]);
    if (h->ev_done[k]) cudaEventDestroy(h->ev_done[k]);
  }
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return NDT_OK;
}

const char *ndt_last_error(ndt_handle hh) {
  Handle *h = reinterpret_cast<Handle *>(hh);
  return h ? h->err.c_str() : g_create_err.c_str();
}

int ndt_set_target(ndt_handle hh, const float *xyzw, int64_t n, int memspace) {
  H_OR_FAIL(hh);
  return grid_build(h, xyzw, n, memspace);
}

int ndt_set_target_prefix(ndt_handle hh, const float *xyzw, int64_t n, int64_t n_same, int memspace) {
  H_OR_FAIL(hh);
  return grid_build(h, xyzw, n, memspace, n_same);
}

int ndt_set_target_incremental(ndt_handle hh, const float *xyzw, int64_t n, int64_t n_same, int64_t n_stable, int memspace) {
  H_OR_FAIL(hh);
  return grid_build_incremental(h, xyzw, n, n_same, n_stable, memspace);
}

int ndt_set_target_incremental_async(ndt_handle hh, const float *xyzw, int64_t n, int64_t n_same, int64_t n_stable, int memspace) {
  H_OR_FAIL(hh);
  return grid_build_incremental(h, xyzw, n, n_same, n_stable, memspace, /*defer=*/true);
}

int ndt_get_grid_info(ndt_handle hh, ndt_grid_info *info) {
  H_OR_FAIL(hh);
  if (!info) return NDT_ERR_ARG;
  if (!h->have_grid) return set_err(h, NDT_ERR_STATE, "ndt_get_grid_info: no target set");
  std::memset(info, 0, sizeof(*info));
  info->min_b[0] = h->gd.min_bx; info->min_b[1] = h->gd.min_by;
  info->div_b[0] = h->gd.div_x; info->div_b[1] = h->gd.div_y;
  info->n_points = h->h_counters[CTR_PTS];
  info->n_leaves = h->h_counters[CTR_LEAVES];
  info->n_slots = h->h_counters[CTR_SLOTS];
  info->n_valid = h->h_counters[CTR_VALID];
  info->reserved = h->inc_active ? 1 : 0;          // 1: the last target call was an incremental update
  return NDT_OK;
}

int ndt_grid_readback(ndt_handle hh, int64_t cap, int32_t *cell_idx, int32_t *nr_points, double *mean2,
                      double *icov4, float *centroid2, int64_t *n_out) {
  H_OR_FAIL(hh);
  if (!h->have_grid) return set_err(h, NDT_ERR_STATE, "ndt_grid_readback: no target set");
  if (!h->have_readback) return set_err(h, NDT_ERR_STATE, "ndt_grid_readback: this grid was imported (ndt_grid_import / ndt_replicate_grid); the per-leaf read-back tables exist only on the handle that built it");
  const int64_t nl = h->h_counters[CTR_LEAVES];
  if (n_out) *n_out = nl;
  if (nl == 0 || cap <= 0) return NDT_OK;
  std::vector<int32_t> cell(nl), nr(nl);
  std::vector<double> mean(2 * nl), icov(4 * nl);
  std::vector<float> cen(2 * nl);
  cudaStream_t st = h->stream;
  NDT_CUDA(h, cudaMemcpyAsync(cell.data(), h->gb.leaf_cell.p, nl * 4, cudaMemcpyDeviceToHost, st));
  NDT_CUDA(h, cudaMemcpyAsync(nr.data(), h->gb.leaf_nr.p, nl * 4, cudaMemcpyDeviceToHost, st));
  NDT_CUDA(h, cudaMemcpyAsync(mean.data(), h->gb.leaf_mean.p, nl * 16, cudaMemcpyDeviceToHost, st));
  NDT_CUDA(h, cudaMemcpyAsync(icov.data(), h->gb.leaf_icov.p, nl * 32, cudaMemcpyDeviceToHost, st));
  NDT_CUDA(h, cudaMemcpyAsync(cen.data(), h->gb.leaf_cen.p, nl * 8, cudaMemcpyDeviceToHost, st));
  NDT_CUDA(h, cudaStreamSynchronize(st));
  // leaves are stored by their position in the padded table: convert to PCL's ijk0 + ijk1 * div_x
  {
    const int W = h->gd.div_x + 4, dx = h->gd.div_x;
    for (int64_t k = 0; k < nl; ++k) { const int q = cell[k], r = q / W, c = q - r * W; cell[k] = (r - 2) * dx + (c - 2); }
  }
  // PCL keeps leaves in a std::map keyed by cell index: report in that order
  std::vector<int64_t> order(nl);
  std::iota(order.begin(), order.end(), 0);
  std::sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return cell[a] < cell[b]; });
  const int64_t m = std::min(cap, nl);
  for (int64_t k = 0; k < m; ++k) {
    const int64_t s = order[k];
    if (cell_idx) cell_idx[k] = cell[s];
    if (nr_points) nr_points[k] = nr[s];
    if (mean2) { mean2[2 * k] = mean[2 * s]; mean2[2 * k + 1] = mean[2 * s + 1]; }
    if (icov4) for (int a = 0; a < 4; ++a) icov4[4 * k + a] = icov[4 * s + a];
    if (centroid2) { centroid2[2 * k] = cen[2 * s]; centroid2[2 * k + 1] = cen[2 * s + 1]; }
  }
  return NDT_OK;
}

int ndt_cell_index(ndt_handle hh, const float *xyzw, int64_t n, int memspace, int32_t *idx_out) {
  H_OR_FAIL(hh);
  if (n > 0 && (!xyzw || !idx_out)) return set_err(h, NDT_ERR_ARG, "ndt_cell_index: null buffer");
  return grid_cell_index(h, xyzw, n, memspace, idx_out);
}

int ndt_set_source(ndt_handle hh, const float *xyzw, int64_t n, int memspace) {
  H_OR_FAIL(hh);
  if (n < 0 || (n > 0 && !xyzw)) return set_err(h, NDT_ERR_ARG, "ndt_set_source: bad points");
  if (n > 0x7fffffffLL) return set_err(h, NDT_ERR_CAPACITY, "ndt_set_source: too many points");
  h->have_src = false;
  NDT_CUDA(h, h->src.reserve((size_t)std::max<int64_t>(n, 1) * sizeof(float4)));
  if (n > 0) {
    if (memspace == NDT_MEM_HOST) {
      if (ensure_pinned(h, (size_t)n * sizeof(float4))) return NDT_ERR_CUDA;
      std::memcpy(h->pinned, xyzw, (size_t)n * sizeof(float4));
      NDT_CUDA(h, cudaMemcpyAsync(h->src.p, h->pinned, (size_t)n * sizeof(float4), cudaMemcpyHostToDevice, h->stream));
      NDT_CUDA(h, cudaStreamSynchronize(h->stream));   // the pinned stage is reused by later calls
    } else {
      NDT_CUDA(h, cudaMemcpyAsync(h->src.p, xyzw, (size_t)n * sizeof(float4), cudaMemcpyDeviceToDevice, h->stream));
    }
  }
  h->ns = n;
  h->have_src = true;
  return NDT_OK;
}

int ndt_approx_voxel_filter(ndt_handle hh, const float *xyzw, int64_t n, float leaf, int memspace,
                            float *out_xyzw, int64_t *n_out) {
  H_OR_FAIL(hh);
  if (n < 0 || !n_out || (n > 0 && (!xyzw || !out_xyzw)) || !(leaf > 0.f))
    return set_err(h, NDT_ERR_ARG, "ndt_approx_voxel_filter: bad argument");
  *n_out = 0;
  if (n == 0) return NDT_OK;
  if (n > 0x7fffffffLL) return set_err(h, NDT_ERR_CAPACITY, "ndt_approx_voxel_filter: more than 2^31-1 points");
  cudaStream_t st = h->stream;
  const size_t bytes = (size_t)n * sizeof(float4);
  NDT_CUDA(h, h->scratch.reserve(2 * bytes + 256));
  float4 *d_in = h->scratch.as<float4>();
  float4 *d_out = d_in + n;
  int32_t *d_n = h->gb.counters.as<int32_t>() + CTR_JOB;
  const float4 *in = reinterpret_cast<const float4 *>(xyzw);
  if (memspace == NDT_MEM_HOST) {
    NDT_CUDA(h, cudaMemcpyAsync(d_in, xyzw, bytes, cudaMemcpyHostToDevice, st));
    in = d_in;
  }
  float4 *outp = (memspace == NDT_MEM_HOST) ? d_out : reinterpret_cast<float4 *>(out_xyzw);
  int rc = launch_voxel_filter(h, in, n, leaf, outp, d_n);
  if (rc) return rc;
  int32_t m = 0;
  NDT_CUDA(h, cudaMemcpyAsync(&m, d_n, 4, cudaMemcpyDeviceToHost, st));
  NDT_CUDA(h, cudaStreamSynchronize(st));
  if (memspace == NDT_MEM_HOST && m > 0) {
    NDT_CUDA(h, cudaMemcpyAsync(out_xyzw, d_out, (size_t)m * sizeof(float4), cudaMemcpyDeviceToHost, st));
    NDT_CUDA(h, cudaStreamSynchronize(st));
  }
  *n_out = m;
  return NDT_OK;
}

// the persistent batch kernels hand out jobs from an int32 counter in chunks: keep counter + chunk inside int32
static constexpr int64_t kMaxBatchJobs = (int64_t)INT32_MAX - 4096;

static int need_ready(Handle *h, const char *who) {
  if (!h->have_grid) return set_err(h, NDT_ERR_STATE, (std::string(who) + ": no target set").c_str());
  if (!h->have_src) return set_err(h, NDT_ERR_STATE, (std::string(who) + ": no source set").c_str());
  return NDT_OK;
}

int ndt_eval(ndt_handle hh, const double pose[3], int want_hessian, ndt_eval_out *out) {
  H_OR_FAIL(hh);
  if (!pose || !out) return set_err(h, NDT_ERR_ARG, "ndt_eval: null argument");
  if (int rc = need_ready(h, "ndt_eval")) return rc;
  cudaStream_t st = h->stream;
  if (ensure_pinned(h, 256)) return NDT_ERR_CUDA;
  static_assert(sizeof(double) * (3 + NACC + 1) + sizeof(int64_t) <= 256, "stage");
  double *hp = (double *)h->pinned;
  hp[0] = pose[0]; hp[1] = pose[1]; hp[2] = pose[2];
  double *d_pose = h->stage.as<double>(), *d_out = d_pose + 3;
  int64_t *d_pairs = (int64_t *)(d_out + NACC + 1);
  NDT_CUDA(h, cudaMemcpyAsync(d_pose, hp, 24, cudaMemcpyHostToDevice, st));
  int rc = launch_eval(h, d_pose, 1, want_hessian, d_out, d_pairs);
  if (rc) return rc;
  NDT_CUDA(h, cudaMemcpyAsync(hp + 3, d_out, sizeof(double) * (NACC + 1) + 8, cudaMemcpyDeviceToHost, st));
  NDT_CUDA(h, cudaStreamSynchronize(st));
  if (h->timing) cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1);
  h->ms_pending = false;
  out->score = hp[3];
  for (int k = 0; k < 3; ++k) out->grad[k] = hp[4 + k];
  for (int k = 0; k < 9; ++k) out->hess[k] = hp[7 + k];
  std::memcpy(&out->n_pairs, hp + 3 + NACC + 1, 8);
  return NDT_OK;
}

int ndt_eval_batch(ndt_handle hh, const double *poses, int64_t n, int want_hessian, int memspace, double *out14) {
  H_OR_FAIL(hh);
  if (n < 0 || (n > 0 && (!poses || !out14))) return set_err(h, NDT_ERR_ARG, "ndt_eval_batch: bad argument");
  if (int rc = need_ready(h, "ndt_eval_batch")) return rc;
  if (n == 0) return NDT_OK;
  if (n > kMaxBatchJobs) return set_err(h, NDT_ERR_CAPACITY, "ndt_eval_batch: too many poses for the 32-bit job counter");
  cudaStream_t st = h->stream;
  if (memspace == NDT_MEM_DEVICE) {
    h->ms_pending = true;
    return launch_eval(h, poses, n, want_hessian, out14, nullptr);
  }
  const size_t pb = (size_t)n * 3 * sizeof(double), ob = (size_t)n * (NACC + 1) * sizeof(double);
  NDT_CUDA(h, h->io.reserve(pb + ob));
  double *d_poses = h->io.as<double>(), *d_out = d_poses + 3 * n;
  NDT_CUDA(h, cudaMemcpyAsync(d_poses, poses, pb, cudaMemcpyHostToDevice, st));
  int rc = launch_eval(h, d_poses, n, want_hessian, d_out, nullptr);
  if (rc) return rc;
  NDT_CUDA(h, cudaMemcpyAsync(out14, d_out, ob, cudaMemcpyDeviceToHost, st));
  NDT_CUDA(h, cudaStreamSynchronize(st));
  if (h->timing) cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1);
  h->ms_pending = false;
  return NDT_OK;
}

int ndt_align(ndt_handle hh, const double guess[3], ndt_result *out) {
  H_OR_FAIL(hh);
  if (!guess || !out) return set_err(h, NDT_ERR_ARG, "ndt_align: null argument");
  if (int rc = need_ready(h, "ndt_align")) return rc;
  cudaStream_t st = h->stream;
  if (ensure_pinned(h, 1024)) return NDT_ERR_CUDA;
  double *hp = (double *)h->pinned;
  hp[0] = guess[0]; hp[1] = guess[1]; hp[2] = guess[2];
  ndt_result *hres = (ndt_result *)(hp + 4);
  double *d_guess = h->stage.as<double>();
  ndt_result *d_res = (ndt_result *)(d_guess + 4);
  NDT_CUDA(h, cudaMemcpyAsync(d_guess, hp, 24, cudaMemcpyHostToDevice, st));
  int rc = launch_align(h, d_guess, 1, d_res, !h->prm.align_skip_fitness);
  if (rc) return rc;
  NDT_CUDA(h, cudaMemcpyAsync(hres, d_res, sizeof(ndt_result), cudaMemcpyDeviceToHost, st));
  NDT_CUDA(h, cudaStreamSynchronize(st));
  if (h->timing) cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1);
  h->ms_pending = false;
  *out = *hres;
  return NDT_OK;
}

int ndt_align_batch(ndt_handle hh, const double *guesses, int64_t n, int memspace, int want_fitness_arg, ndt_result *results) {
  H_OR_FAIL(hh);
  if (n < 0 || (n > 0 && (!guesses || !results))) return set_err(h, NDT_ERR_ARG, "ndt_align_batch: bad argument");
  if (int rc = need_ready(h, "ndt_align_batch")) return rc;
  if (n == 0) return NDT_OK;
  if (n > kMaxBatchJobs) return set_err(h, NDT_ERR_CAPACITY, "ndt_align_batch: too many guesses for the 32-bit job counter");
  cudaStream_t st = h->stream;
  const bool want_fitness = want_fitness_arg != 0;
  if (memspace == NDT_MEM_DEVICE) { h->ms_pending = true; return launch_align(h, guesses, n, results, want_fitness); }
  const size_t gbytes = (size_t)n * 3 * sizeof(double), rbytes = (size_t)n * sizeof(ndt_result);
  NDT_CUDA(h, h->io.reserve(gbytes + rbytes + 256));
  double *d_g = h->io.as<double>();
  ndt_result *d_r = (ndt_result *)((char *)h->io.p + ((gbytes + 255) & ~size_t(255)));
  NDT_CUDA(h, cudaMemcpyAsync(d_g, guesses, gbytes, cudaMemcpyHostToDevice, st));
  int rc = launch_align(h, d_g, n, d_r, want_fitness);
  if (rc) return rc;
  NDT_CUDA(h, cudaMemcpyAsync(results, d_r, rbytes, cudaMemcpyDeviceToHost, st));
  NDT_CUDA(h, cudaStreamSynchronize(st));
  if (h->timing) cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1);
  h->ms_pending = false;
  return NDT_OK;
}

int ndt_best_of(ndt_handle hh, const ndt_result *results, int64_t n, int memspace, int64_t *best_index,
                ndt_result *best) {
  H_OR_FAIL(hh);
  if (n <= 0 || !results || !best_index || !best) return set_err(h, NDT_ERR_ARG, "ndt_best_of: bad argument");
  cudaStream_t st = h->stream;
  if (ensure_pinned(h, 1024)) return NDT_ERR_CUDA;
  const size_t rbytes = (size_t)n * sizeof(ndt_result);
  if (memspace == NDT_MEM_HOST) NDT_CUDA(h, h->io.reserve(rbytes));
  int64_t *d_bi = h->stage.as<int64_t>();
  ndt_result *d_best = (ndt_result *)(d_bi + 2);
  const ndt_result *d_res = results;
  if (memspace == NDT_MEM_HOST) {
    ndt_result *tmp = h->io.as<ndt_result>();
    NDT_CUDA(h, cudaMemcpyAsync(tmp, results, rbytes, cudaMemcpyHostToDevice, st));
    d_res = tmp;
  }
  int rc = launch_best_of(h, d_res, n, d_bi, d_best);
  if (rc) return rc;
  char *hp = (char *)h->pinned;
  NDT_CUDA(h, cudaMemcpyAsync(hp, d_bi, 8, cudaMemcpyDeviceToHost, st));
  NDT_CUDA(h, cudaMemcpyAsync(hp + 16, d_best, sizeof(ndt_result), cudaMemcpyDeviceToHost, st));
  NDT_CUDA(h, cudaStreamSynchronize(st));
  std::memcpy(best_index, hp, 8);
  std::memcpy(best, hp + 16, sizeof(ndt_result));
  return NDT_OK;
}

// ndt_grid_import: the matcher's arithmetic assumes finite records (a rejected hit is turned into exact zeros by selects on
// its inputs, and 0 * NaN is not 0). The grid build only ever writes finite records; a blob is checked.
__global__ void k_check_records(const CellRec *__restrict__ recs, int n, int32_t *__restrict__ bad) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double *d = reinterpret_cast<const double *>(recs + i);        // 64 B: {float cx, cy; int nr, cell; double mx, my, c00, c01, c10, c11}
    const float *f = reinterpret_cast<const float *>(recs + i);
    bool ok = isfinite(f[0]) && isfinite(f[1]);
#pragma unroll
    for (int k = 2; k < 8; ++k) ok = ok && isfinite(d[k]);
    if (!ok) atomicExch(bad, 1);
  }
}

// z = 0 planes of 2-D SLAM carry 8 useful bytes per point: the compact entry uploads (x, y) pairs and widens them on the device
__global__ void k_expand_xy(const float2 *__restrict__ in, float4 *__restrict__ out, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float2 p = __ldg(in + i);
    out[i] = make_float4(p.x, p.y, 0.f, 0.f);
  }
}

static int match_pairs_impl(ndt_handle hh, const float *src_xyzw, const int64_t *src_off, const float *tgt_xyzw,
                            const int64_t *tgt_off, const double *guesses, int64_t n_pairs, float source_leaf,
                            int memspace, ndt_result *results, const bool xy) {
  H_OR_FAIL(hh);
  if (n_pairs < 0 || (n_pairs > 0 && (!src_off || !tgt_off || !guesses || !results)))
    return set_err(h, NDT_ERR_ARG, "ndt_match_pairs: bad argument");
  if (n_pairs == 0) return NDT_OK;
  if (n_pairs > (int64_t)(1 << 30)) return set_err(h, NDT_ERR_CAPACITY, "ndt_match_pairs: too many pairs");
  // the offset arrays are host metadata (like the point counts of the single-match calls)
  const int64_t ns_total = src_off[n_pairs] - src_off[0], nt_total = tgt_off[n_pairs] - tgt_off[0];
  if (src_off[0] != 0 || tgt_off[0] != 0) return set_err(h, NDT_ERR_ARG, "ndt_match_pairs: offsets must start at 0");
  for (int64_t i = 0; i < n_pairs; ++i) {
    if (src_off[i + 1] < src_off[i] || tgt_off[i + 1] < tgt_off[i]) return set_err(h, NDT_ERR_ARG, "ndt_match_pairs: offsets must be non-decreasing");
    if (src_off[i + 1] - src_off[i] > 0x7fffffffLL) return set_err(h, NDT_ERR_CAPACITY, "ndt_match_pairs: source cloud too large");
  }
  if ((ns_total > 0 && !src_xyzw) || (nt_total > 0 && !tgt_xyzw)) return set_err(h, NDT_ERR_ARG, "ndt_match_pairs: null points");
  if (nt_total > (int64_t)INT32_MAX) return set_err(h, NDT_ERR_CAPACITY, "ndt_match_pairs: more than 2^31-1 target points");
  cudaStream_t st = h->stream;
  GridBuffers &gb = h->gb;
  GridDims &gd = h->gd;
  h->have_grid = false; h->have_src = false;          // the handle's single grid / source are overwritten
  h->have_readback = false; h->grid_has_points = false;
  h->inc_ok = false; h->inc_active = false;
  h->tgt_on_device = 0;
  gd = GridDims();
  gd.leaf = h->prm.resolution; gd.inv_leaf = 1.0f / gd.leaf; gd.r2 = (float)((double)gd.leaf * (double)gd.leaf);
  std::memset(h->h_counters, 0, sizeof(h->h_counters));
  const bool host = (memspace == NDT_MEM_HOST);
  const cudaMemcpyKind in_kind = host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
  const size_t gbytes = (size_t)n_pairs * 3 * sizeof(double), rbytes = (size_t)n_pairs * sizeof(ndt_result);
  const double *d_g = guesses;
  ndt_result *d_r = results;
  if (host) {
    NDT_CUDA(h, h->io.reserve(gbytes + rbytes + 256));
    d_g = h->io.as<double>();
    d_r = (ndt_result *)((char *)h->io.p + ((gbytes + 255) & ~size_t(255)));
    NDT_CUDA(h, cudaMemcpyAsync(h->io.p, guesses, gbytes, cudaMemcpyHostToDevice, st));
  }
  // Pairs go through the pipeline in batches of at most `batch_points` target points: bounds the size of the shared
  // tables for very large inputs. Device inputs: one batch is best (measured on C5: a pair is matched by one warp in
  // ~1 ms, so every batch ends with a latency tail). Host inputs: the upload of the clouds (16 B per point over PCIe)
  // takes as long as matching them, so the call is cut into a few batches and batch k+1 is uploaded on a second stream
  // into a second pair of buffers while batch k is filtered, gridded and matched.
  int64_t batch_points = h->prm.pairs_batch_points > 0 ? h->prm.pairs_batch_points : 32000000;
  if (host && !xy && h->prm.pairs_batch_points <= 0)
    batch_points = std::min<int64_t>(batch_points, std::max<int64_t>(nt_total / 4 + 1, (int64_t)1 << 20));
  std::vector<int64_t> cuts{0};
  int64_t max_nt = 1, max_ns = 1, max_nb = 1;
  for (int64_t p0 = 0; p0 < n_pairs;) {
    int64_t p1 = p0 + 1;
    while (p1 < n_pairs && tgt_off[p1 + 1] - tgt_off[p0] <= batch_points && src_off[p1 + 1] - src_off[p0] <= 2 * batch_points) ++p1;
    max_nt = std::max(max_nt, tgt_off[p1] - tgt_off[p0]);
    max_ns = std::max(max_ns, src_off[p1] - src_off[p0]);
    max_nb = std::max(max_nb, p1 - p0);
    cuts.push_back(p1);
    p0 = p1;
  }
  const int n_batches = (int)cuts.size() - 1;
  const bool pipelined = host && !xy && n_batches > 1;   // compact clouds halve the upload instead: one batch, best schedule
  // every buffer a batch needs is sized for the largest batch before the first upload: nothing is reallocated under a copy
  NDT_CUDA(h, gb.tgt.reserve((size_t)max_nt * sizeof(float4)));
  NDT_CUDA(h, h->src.reserve((size_t)max_ns * sizeof(float4)));
  NDT_CUDA(h, gb.pair_off.reserve(2 * ((size_t)max_nb + 1) * sizeof(int64_t)));
  if (host || xy) NDT_CUDA(h, h->scratch.reserve((size_t)max_ns * sizeof(float4)));
  if (xy && host) {                                        // staging of the (x, y) uploads
    NDT_CUDA(h, h->pair_tgt_next.reserve((size_t)max_nt * sizeof(float2)));
    NDT_CUDA(h, h->pair_raw_next.reserve((size_t)max_ns * sizeof(float2)));
  }
  cudaStream_t cs = nullptr;
  if (pipelined) {
    NDT_CUDA(h, h->pair_tgt_next.reserve((size_t)max_nt * sizeof(float4)));
    NDT_CUDA(h, h->pair_raw_next.reserve((size_t)max_ns * sizeof(float4)));
    if (!h->copy_stream) {
      NDT_CUDA(h, cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
      for (int k = 0; k < 2; ++k) {
        NDT_CUDA(h, cudaEventCreateWithFlags(&h->ev_up[k], cudaEventDisableTiming));
        NDT_CUDA(h, cudaEventCreateWithFlags(&h->ev_done[k], cudaEventDisableTiming));
      }
    }
    cs = h->copy_stream;
  }
  if (h->timing) cudaEventRecord(h->ev0, st);
  const int pf = xy ? 2 : 4;                               // floats per input point
  auto expand_blocks = [&](int64_t n) { return (unsigned)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, (int64_t)h->sm_count * 8)); };
  auto upload = [&](int k, DevBuf &tgt_buf, DevBuf &raw_buf, cudaStream_t s) -> cudaError_t {
    const int64_t q0 = cuts[k], q1 = cuts[k + 1];
    const int64_t nt = tgt_off[q1] - tgt_off[q0], ns = src_off[q1] - src_off[q0];
    cudaError_t e = cudaSuccess;
    if (xy) {
      const float2 *t2 = reinterpret_cast<const float2 *>(tgt_xyzw) + tgt_off[q0], *s2 = reinterpret_cast<const float2 *>(src_xyzw) + src_off[q0];
      if (host) {
        if (nt > 0) e = cudaMemcpyAsync(h->pair_tgt_next.p, t2, (size_t)nt * sizeof(float2), cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess && ns > 0) e = cudaMemcpyAsync(h->pair_raw_next.p, s2, (size_t)ns * sizeof(float2), cudaMemcpyHostToDevice, s);
        t2 = h->pair_tgt_next.as<float2>(); s2 = h->pair_raw_next.as<float2>();
      }
      if (e != cudaSuccess) return e;
      if (nt > 0) { k_expand_xy<<<expand_blocks(nt), 256, 0, s>>>(t2, tgt_buf.as<float4>(), nt); ++h->launches; }
      if (ns > 0) { k_expand_xy<<<expand_blocks(ns), 256, 0, s>>>(s2, raw_buf.as<float4>(), ns); ++h->launches; }
      return cudaGetLastError();
    }
    if (nt > 0) e = cudaMemcpyAsync(tgt_buf.p, tgt_xyzw + pf * tgt_off[q0], (size_t)nt * sizeof(float4), in_kind, s);
    if (e == cudaSuccess && host && ns > 0)
      e = cudaMemcpyAsync(raw_buf.p, src_xyzw + pf * src_off[q0], (size_t)ns * sizeof(float4), cudaMemcpyHostToDevice, s);
    return e;
  };
  auto fail = [&](int rc) { if (cs) cudaStreamSynchronize(cs); return rc; };     // no upload may outlive the call
  if (pipelined) {
    // the buffers were (re)allocated in order of `st`: the upload stream starts behind that point
    NDT_CUDA(h, cudaEventRecord(h->ev_done[1], st));
    NDT_CUDA(h, cudaStreamWaitEvent(cs, h->ev_done[1], 0));
    if (upload(0, h->pair_tgt_next, h->pair_raw_next, cs) != cudaSuccess || cudaEventRecord(h->ev_up[0], cs) != cudaSuccess)
      return fail(set_err(h, NDT_ERR_CUDA, "ndt_match_pairs: upload", cudaGetLastError()));
  }
  std::vector<int64_t> off_stage;
  for (int k = 0; k < n_batches; ++k) {
    const int64_t p0 = cuts[k], p1 = cuts[k + 1];
    const int64_t nb = p1 - p0, nt = tgt_off[p1] - tgt_off[p0], ns = src_off[p1] - src_off[p0];
    if (pipelined) {
      std::swap(gb.tgt, h->pair_tgt_next);           // batch k's clouds were uploaded into the `next` pair of buffers
      std::swap(h->scratch, h->pair_raw_next);
      cudaError_t e = cudaStreamWaitEvent(st, h->ev_up[k & 1], 0);
      if (e == cudaSuccess && k + 1 < n_batches) {
        // the `next` buffers now are the ones batch k-1 computed on: its kernels must be done before they are overwritten
        if (k >= 1) e = cudaStreamWaitEvent(cs, h->ev_done[(k - 1) & 1], 0);
        if (e == cudaSuccess) e = upload(k + 1, h->pair_tgt_next, h->pair_raw_next, cs);
        if (e == cudaSuccess) e = cudaEventRecord(h->ev_up[(k + 1) & 1], cs);
      }
      if (e != cudaSuccess) return fail(set_err(h, NDT_ERR_CUDA, "ndt_match_pairs: upload", e));
    } else if (upload(k, gb.tgt, h->scratch, st) != cudaSuccess) {
      return set_err(h, NDT_ERR_CUDA, "ndt_match_pairs: upload", cudaGetLastError());
    }
    off_stage.resize(2 * ((size_t)nb + 1));
    for (int64_t i = 0; i <= nb; ++i) {
      off_stage[i] = tgt_off[p0 + i] - tgt_off[p0];
      off_stage[nb + 1 + i] = src_off[p0 + i] - src_off[p0];
    }
    if (cudaMemcpyAsync(gb.pair_off.p, off_stage.data(), off_stage.size() * sizeof(int64_t), cudaMemcpyHostToDevice, st) != cudaSuccess)
      return fail(set_err(h, NDT_ERR_CUDA, "ndt_match_pairs: offsets", cudaGetLastError()));
    // raw source clouds: device inputs are filtered straight from the caller's buffer
    const float4 *d_raw = (host || xy) ? h->scratch.as<float4>() : reinterpret_cast<const float4 *>(src_xyzw) + src_off[p0];
    int64_t total_pad = 0;
    int max_h = 0;
    if (int rc = pairs_prepare(h, nb, &total_pad, &max_h)) return fail(rc);    // one 16-byte read-back (also fences off_stage)
    if (int rc = launch_pairs_filter(h, d_raw, h->src.as<float4>(), nb, source_leaf)) return fail(rc);
    gd.n_tgt = nt;
    if (int rc = grid_build_tables(h, nt, (int)nb, total_pad, max_h)) return fail(rc);
    if (int rc = launch_align_pairs(h, h->src.as<float4>(), d_g + 3 * p0, nb, d_r + p0, /*want_fitness=*/true)) return fail(rc);
    if (pipelined && cudaEventRecord(h->ev_done[k & 1], st) != cudaSuccess)
      return fail(set_err(h, NDT_ERR_CUDA, "ndt_match_pairs: event", cudaGetLastError()));
    (void)ns;
  }
  if (h->timing) cudaEventRecord(h->ev1, st);
  if (!host) { h->ms_pending = true; return NDT_OK; }
  NDT_CUDA(h, cudaMemcpyAsync(results, d_r, rbytes, cudaMemcpyDeviceToHost, st));
  NDT_CUDA(h, cudaStreamSynchronize(st));
  if (h->timing) cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1);
  h->ms_pending = false;
  return NDT_OK;
}

int ndt_match_pairs(ndt_handle hh, const float *src_xyzw, const int64_t *src_off, const float *tgt_xyzw,
                    const int64_t *tgt_off, const double *guesses, int64_t n_pairs, float source_leaf,
                    int memspace, ndt_result *results) {
  return match_pairs_impl(hh, src_xyzw, src_off, tgt_xyzw, tgt_off, guesses, n_pairs, source_leaf, memspace, results, false);
}

int ndt_match_pairs_xy(ndt_handle hh, const float *src_xy, const int64_t *src_off, const float *tgt_xy,
                       const int64_t *tgt_off, const double *guesses, int64_t n_pairs, float source_leaf,
                       int memspace, ndt_result *results) {
  return match_pairs_impl(hh, src_xy, src_off, tgt_xy, tgt_off, guesses, n_pairs, source_leaf, memspace, results, true);
}

int ndt_grid_blob_size(ndt_handle hh, int flags, int64_t *bytes) {
  H_OR_FAIL(hh);
  if (!bytes) return NDT_ERR_ARG;
  if (!h->have_grid) return set_err(h, NDT_ERR_STATE, "ndt_grid_blob_size: no target set");
  *bytes = blob_layout(h, flags).total;
  return NDT_OK;
}

int ndt_grid_export(ndt_handle hh, int flags, void *device_blob, int64_t bytes) {
  H_OR_FAIL(hh);
  if (!h->have_grid) return set_err(h, NDT_ERR_STATE, "ndt_grid_export: no target set");
  const BlobHeader b = blob_layout(h, flags);
  if (!device_blob || bytes < b.total) return set_err(h, NDT_ERR_ARG, "ndt_grid_export: blob too small");
  cudaStream_t st = h->stream;
  char *d = (char *)device_blob;
  if (ensure_pinned(h, sizeof(BlobHeader))) return NDT_ERR_CUDA;
  std::memcpy(h->pinned, &b, sizeof(b));
  NDT_CUDA(h, cudaMemcpyAsync(d, h->pinned, sizeof(b), cudaMemcpyHostToDevice, st));
  auto cp = [&](int64_t off, const DevBuf &src, int64_t nbytes) -> cudaError_t {
    if (nbytes <= 0) return cudaSuccess;
    return cudaMemcpyAsync(d + off, src.p, (size_t)nbytes, cudaMemcpyDeviceToDevice, st);
  };
  const BlobSizes z = blob_sizes(b.gd, b.counters, b.flags);
  NDT_CUDA(h, cp(b.off_slot, h->gb.slot, z.slot));
  NDT_CUDA(h, cp(b.off_cen, h->gb.cen, z.cen));
  NDT_CUDA(h, cp(b.off_occ, h->gb.occ, z.occ));
  NDT_CUDA(h, cp(b.off_recs, h->gb.recs, z.recs));
  NDT_CUDA(h, cp(b.off_leaf_id, h->gb.leaf_id, z.leaf_id));
  NDT_CUDA(h, cp(b.off_leaf_range, h->gb.leaf_range, z.leaf_range));
  NDT_CUDA(h, cp(b.off_sorted, h->gb.tgt_sorted, z.sorted));
  NDT_CUDA(h, cp(b.off_tgt, h->gb.tgt, z.tgt));
  NDT_CUDA(h, cp(b.off_nn_range, h->gb.nn_range, z.nn_range));
  NDT_CUDA(h, cp(b.off_nn_pts, h->gb.nn_pts, z.nn_pts));
  NDT_CUDA(h, cudaStreamSynchronize(st));
  return NDT_OK;
}

// A blob comes from another process / GPU: nothing in its header is trusted. Geometry and counters must be plausible and
// mutually consistent, every section must lie inside [header, total] in layout order, total must fit the buffer.
static const char *blob_check(const Handle *h, const BlobHeader &b, int64_t bytes) {
  if (b.magic != kBlobMagic) return "not a grid blob (magic)";
  if (b.flags & ~NDT_BLOB_POINTS) return "unknown flags";
  const GridDims &gd = b.gd;
  if (!(gd.leaf == h->prm.resolution)) return "the blob's cell size differs from this handle's resolution";
  if (gd.div_x < 0 || gd.div_y < 0 || gd.n_tgt < 0 || gd.n_tgt > (int64_t)INT32_MAX) return "bad geometry";
  if (gd.n_cells != (int64_t)gd.div_x * gd.div_y) return "bad cell count";
  if ((int64_t)(gd.div_x + 4) * (gd.div_y + 4) > (int64_t)INT32_MAX) return "grid too large";
  const int64_t npad = gd.n_cells > 0 ? (int64_t)(gd.div_x + 4) * (gd.div_y + 4) : 0;
  const int64_t nl = b.counters[CTR_LEAVES], nsl = b.counters[CTR_SLOTS];
  if (nl < 0 || nsl < 0 || nsl > nl || nl > npad || nl > gd.n_tgt) return "bad leaf / slot counters";
  if (gd.nn_f < 0 || gd.nn_f > 64 || (gd.nn_f > 0 && (gd.nn_div_x <= 0 || gd.nn_div_y <= 0 ||
      (int64_t)gd.nn_div_x * gd.nn_div_y > (int64_t)64 * 1024 * 1024))) return "bad 1-NN lattice";
  const BlobSizes z = blob_sizes(gd, b.counters, b.flags);
  const int64_t off[10] = {b.off_slot, b.off_cen, b.off_occ, b.off_recs, b.off_leaf_id, b.off_leaf_range, b.off_sorted,
                           b.off_tgt, b.off_nn_range, b.off_nn_pts};
  const int64_t len[10] = {z.slot, z.cen, z.occ, z.recs, z.leaf_id, z.leaf_range, z.sorted, z.tgt, z.nn_range, z.nn_pts};
  int64_t lo = (int64_t)sizeof(BlobHeader);
  for (int k = 0; k < 10; ++k) {
    if (off[k] < lo || (off[k] & 255) || off[k] > b.total || len[k] > b.total - off[k]) return "section offsets out of range";
    lo = off[k] + len[k];
  }
  if (b.total > bytes) return "blob truncated";
  return nullptr;
}

int ndt_grid_import(ndt_handle hh, const void *device_blob, int64_t bytes) {
  H_OR_FAIL(hh);
  if (!device_blob || bytes < (int64_t)sizeof(BlobHeader)) return set_err(h, NDT_ERR_ARG, "ndt_grid_import: bad blob");
  cudaStream_t st = h->stream;
  BlobHeader b{};
  NDT_CUDA(h, cudaMemcpyAsync(&b, device_blob, sizeof(b), cudaMemcpyDeviceToHost, st));
  NDT_CUDA(h, cudaStreamSynchronize(st));
  if (const char *why = blob_check(h, b, bytes)) return set_err(h, NDT_ERR_ARG, (std::string("ndt_grid_import: ") + why).c_str());
  h->have_grid = false; h->have_readback = false; h->grid_has_points = false;
  h->inc_ok = false; h->inc_active = false;
  h->tgt_on_device = 0;
  h->gd = b.gd;
  std::memcpy(h->h_counters, b.counters, sizeof(b.counters));
  const char *d = (const char *)device_blob;
  auto take = [&](DevBuf &dst, int64_t off, int64_t nbytes) -> cudaError_t {
    cudaError_t e = dst.reserve((size_t)std::max<int64_t>(nbytes, 16));
    if (e != cudaSuccess || nbytes <= 0) return e;
    return cudaMemcpyAsync(dst.p, d + off, (size_t)nbytes, cudaMemcpyDeviceToDevice, st);
  };
  const BlobSizes z = blob_sizes(b.gd, b.counters, b.flags);
  if (z.slot == 0) {
    // empty grid: the 4 x 4 all-empty padded table ndt_set_target builds for an empty target
    NDT_CUDA(h, h->gb.occ.reserve(64)); NDT_CUDA(h, h->gb.slot.reserve(64)); NDT_CUDA(h, h->gb.cen.reserve(128)); NDT_CUDA(h, h->gb.leaf_id.reserve(64));
    NDT_CUDA(h, cudaMemsetAsync(h->gb.occ.p, 0, 64, st)); NDT_CUDA(h, cudaMemsetAsync(h->gb.slot.p, 0xff, 64, st));
    NDT_CUDA(h, cudaMemsetAsync(h->gb.cen.p, 0xff, 128, st)); NDT_CUDA(h, cudaMemsetAsync(h->gb.leaf_id.p, 0, 64, st));
  }
  NDT_CUDA(h, take(h->gb.slot, b.off_slot, z.slot));
  NDT_CUDA(h, take(h->gb.cen, b.off_cen, z.cen));
  NDT_CUDA(h, take(h->gb.occ, b.off_occ, z.occ));
  NDT_CUDA(h, take(h->gb.recs, b.off_recs, z.recs));
  NDT_CUDA(h, take(h->gb.leaf_id, b.off_leaf_id, z.leaf_id));
  NDT_CUDA(h, take(h->gb.leaf_range, b.off_leaf_range, z.leaf_range));
  NDT_CUDA(h, take(h->gb.tgt_sorted, b.off_sorted, z.sorted));
  NDT_CUDA(h, take(h->gb.tgt, b.off_tgt, z.tgt));
  NDT_CUDA(h, take(h->gb.nn_range, b.off_nn_range, z.nn_range));
  NDT_CUDA(h, take(h->gb.nn_pts, b.off_nn_pts, z.nn_pts));
  // one grid at base 0 (the geometry the batch kernels' neighbour masks are derived from, on demand)
  {
    PairDims pd{};
    pd.min_bx = b.gd.min_bx; pd.min_by = b.gd.min_by; pd.div_x = b.gd.div_x; pd.div_y = b.gd.div_y;
    pd.W = b.gd.div_x + 4; pd.H = b.gd.div_y + 4; pd.base = 0; pd.nt = b.gd.n_tgt;
    NDT_CUDA(h, h->gb.dims.reserve(sizeof(PairDims)));
    if (ensure_pinned(h, sizeof(PairDims))) return NDT_ERR_CUDA;
    std::memcpy(h->pinned, &pd, sizeof(pd));
    NDT_CUDA(h, cudaMemcpyAsync(h->gb.dims.p, h->pinned, sizeof(pd), cudaMemcpyHostToDevice, st));
    h->have_nbr = false; h->nbr_grids = 1;
    h->nbr_cells = z.slot > 0 ? (int64_t)pd.W * pd.H : 0;
  }
  const int n_recs = (int)(z.recs / (int64_t)sizeof(CellRec));
  int32_t *bad = h->gb.counters.as<int32_t>() + CTR_BIG;
  NDT_CUDA(h, cudaMemsetAsync(bad, 0, sizeof(int32_t), st));
  if (n_recs > 0) {
    k_check_records<<<std::min(std::max(1, (n_recs + 255) / 256), h->sm_count * 8), 256, 0, st>>>(h->gb.recs.as<CellRec>(), n_recs, bad);
    ++h->launches;
  }
  NDT_CUDA(h, cudaMemcpyAsync(h->pinned_ctr, bad, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  NDT_CUDA(h, cudaStreamSynchronize(st));
  if (h->pinned_ctr[0] != 0) return set_err(h, NDT_ERR_ARG, "ndt_grid_import: non-finite cell record");
  h->have_grid = true;
  h->grid_has_points = (b.flags & NDT_BLOB_POINTS) != 0;
  return NDT_OK;
}

int ndt_replicate_grid(const ndt_handle *handles, int n_handles, int flags) {
  if (!handles || n_handles < 1 || !handles[0]) return NDT_ERR_ARG;
  Handle *h0 = reinterpret_cast<Handle *>(handles[0]);
  for (int k = 1; k < n_handles; ++k) {
    if (!handles[k] || handles[k] == handles[0]) return set_err(h0, NDT_ERR_ARG, "ndt_replicate_grid: null or repeated handle");
    for (int j = 1; j < k; ++j) if (handles[j] == handles[k]) return set_err(h0, NDT_ERR_ARG, "ndt_replicate_grid: repeated handle");
  }
  if (!h0->have_grid) return set_err(h0, NDT_ERR_STATE, "ndt_replicate_grid: handles[0] has no target set");
  if (n_handles == 1) return NDT_OK;
  if (cudaSetDevice(h0->device) != cudaSuccess) return set_err(h0, NDT_ERR_CUDA, "cudaSetDevice");
  const int64_t total = blob_layout(h0, flags).total;
  // the source blob lives in its own allocation: scratch buffers are reused by other calls on the handle
  void *src_blob = nullptr;
  NDT_CUDA(h0, cudaMalloc(&src_blob, (size_t)total));
  int rc = ndt_grid_export(handles[0], flags, src_blob, total);
  for (int k = 1; k < n_handles && rc == NDT_OK; ++k) {
    Handle *hk = reinterpret_cast<Handle *>(handles[k]);
    if (cudaSetDevice(hk->device) != cudaSuccess) { rc = set_err(hk, NDT_ERR_CUDA, "cudaSetDevice"); break; }
    void *dst_blob = src_blob;
    if (hk->device != h0->device) {
      int can = 0;
      cudaDeviceCanAccessPeer(&can, hk->device, h0->device);
      if (can) { cudaError_t e = cudaDeviceEnablePeerAccess(h0->device, 0); if (e != cudaSuccess) (void)cudaGetLastError(); }   // already enabled is fine
      cudaError_t e = cudaMalloc(&dst_blob, (size_t)total);
      if (e == cudaSuccess) e = cudaMemcpyPeerAsync(dst_blob, hk->device, src_blob, h0->device, (size_t)total, hk->stream);   // NVLink when peers, staged otherwise
      if (e == cudaSuccess) e = cudaStreamSynchronize(hk->stream);
      if (e != cudaSuccess) { rc = set_err(hk, NDT_ERR_CUDA, "ndt_replicate_grid: peer copy", e); if (dst_blob != src_blob) cudaFree(dst_blob); break; }
    }
    rc = ndt_grid_import(handles[k], dst_blob, total);
    if (rc != NDT_OK) set_err(h0, rc, (std::string("ndt_replicate_grid: import failed: ") + hk->err).c_str());
    if (dst_blob != src_blob) cudaFree(dst_blob);
  }
  cudaSetDevice(h0->device);
  cudaFree(src_blob);
  return rc;
}

int ndt_best_of_multi(const ndt_handle *handles, const ndt_result *const *device_results, const int64_t *counts,
                      int n_handles, int *best_handle, int64_t *best_index, ndt_result *best) {
  if (!handles || !device_results || !counts || n_handles < 1 || !best_handle || !best_index || !best) return NDT_ERR_ARG;
  *best_handle = -1; *best_index = -1;
  for (int k = 0; k < n_handles; ++k) {
    if (!handles[k]) return NDT_ERR_ARG;
    if (counts[k] <= 0) continue;
    int64_t bi = -1;
    ndt_result r;
    const int rc = ndt_best_of(handles[k], device_results[k], counts[k], NDT_MEM_DEVICE, &bi, &r);
    if (rc != NDT_OK) return rc;
    if (bi >= 0 && (*best_handle < 0 || r.score > best->score)) { *best_handle = k; *best_index = bi; *best = r; }
  }
  return NDT_OK;
}

int ndt_trim(ndt_handle hh) {
  H_OR_FAIL(hh);
  NDT_CUDA(h, cudaStreamSynchronize(h->stream));
  if (h->pool) NDT_CUDA(h, cudaMemPoolTrimTo(h->pool, 0));
  return NDT_OK;
}

int ndt_alloc(ndt_handle hh, int64_t bytes, void **device_ptr) {
  H_OR_FAIL(hh);
  if (!device_ptr || bytes < 0) return set_err(h, NDT_ERR_ARG, "ndt_alloc: bad argument");
  *device_ptr = nullptr;
  NDT_CUDA(h, cudaMalloc(device_ptr, (size_t)std::max<int64_t>(bytes, 16)));
  return NDT_OK;
}

int ndt_free(ndt_handle hh, void *device_ptr) {
  H_OR_FAIL(hh);
  if (device_ptr) { NDT_CUDA(h, cudaStreamSynchronize(h->stream)); NDT_CUDA(h, cudaFree(device_ptr)); }
  return NDT_OK;
}

int ndt_upload(ndt_handle hh, void *device_dst, const void *host_src, int64_t bytes) {
  H_OR_FAIL(hh);
  if (bytes < 0 || (bytes > 0 && (!device_dst || !host_src))) return set_err(h, NDT_ERR_ARG, "ndt_upload: bad argument");
  if (bytes == 0) return NDT_OK;
  NDT_CUDA(h, cudaMemcpyAsync(device_dst, host_src, (size_t)bytes, cudaMemcpyHostToDevice, h->stream));
  NDT_CUDA(h, cudaStreamSynchronize(h->stream));
  return NDT_OK;
}

int ndt_download(ndt_handle hh, void *host_dst, const void *device_src, int64_t bytes) {
  H_OR_FAIL(hh);
  if (bytes < 0 || (bytes > 0 && (!host_dst || !device_src))) return set_err(h, NDT_ERR_ARG, "ndt_download: bad argument");
  if (bytes == 0) return NDT_OK;
  NDT_CUDA(h, cudaMemcpyAsync(host_dst, device_src, (size_t)bytes, cudaMemcpyDeviceToHost, h->stream));
  NDT_CUDA(h, cudaStreamSynchronize(h->stream));
  return NDT_OK;
}

int ndt_launch_count(ndt_handle hh, int64_t *n) {
  Handle *h = reinterpret_cast<Handle *>(hh);
  if (!h || !n) return NDT_ERR_ARG;
  *n = h->launches;
  return NDT_OK;
}

int ndt_last_kernel_ms(ndt_handle hh, float *ms) {
  H_OR_FAIL(hh);
  if (!ms) return NDT_ERR_ARG;
  if (h->ms_pending) {   // device-space batch calls return before completion: resolve the event pair now
    NDT_CUDA(h, cudaEventSynchronize(h->ev1));
    NDT_CUDA(h, cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1));
    h->ms_pending = false;
  }
  *ms = h->last_ms;
  return NDT_OK;
}

int ndt_synchronize(ndt_handle hh) {
  H_OR_FAIL(hh);
  NDT_CUDA(h, cudaStreamSynchronize(h->stream));
  return NDT_OK;
}

}  // extern "C"
