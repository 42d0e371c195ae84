// ndt_host.h -- host-side state of one ndt_handle and the launcher entry points shared between the
// translation units (grid_build.cu is compiled with --fmad=false, match_kernels.cu with FMA on).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

#include "ndt_b200.h"
#include "ndt_device.cuh"

namespace ndt {

// growable device buffer (capacity only ever grows). Buffers of a handle are bound to the handle's stream and to a private
// stream-ordered memory pool (cudaMallocFromPoolAsync / cudaFreeAsync): growing a buffer neither synchronises the device
// nor unmaps memory -- freed blocks stay in the pool and are stitched into the next, larger request. A map that grows scan
// by scan reallocates ~25 buffers O(log n) times; with cudaMalloc / cudaFree those scans cost tens of milliseconds.
struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
  cudaStream_t st = nullptr;
  cudaMemPool_t pool = nullptr;      // null: plain cudaMalloc / cudaFree
  void bind(cudaStream_t s, cudaMemPool_t m) { st = s; pool = m; }
  // keep > 0: the first `keep` bytes survive a reallocation (stream-ordered copy)
  cudaError_t reserve(size_t bytes, size_t keep = 0, cudaStream_t /*unused*/ = nullptr) {
    if (bytes <= cap) return cudaSuccess;
    // first allocation: what was asked for (a handle has ~30 tables; doubling every one of them doubled the footprint of
    // a C3-size grid for nothing); regrowth: geometric, so a map that grows scan by scan reallocates O(log n) times
    const size_t want = cap == 0 ? ((bytes + 4095) & ~size_t(4095)) : (bytes + bytes / 2 + 4096);
    void *q = nullptr;
    cudaError_t e = pool ? cudaMallocFromPoolAsync(&q, want, pool, st) : cudaMalloc(&q, want);
    if (e != cudaSuccess) return e;
    if (p && keep > 0) {
      e = cudaMemcpyAsync(q, p, keep < cap ? keep : cap, cudaMemcpyDeviceToDevice, st);
      if (e == cudaSuccess && !pool) e = cudaStreamSynchronize(st);
    }
    if (p) { if (pool) cudaFreeAsync(p, st); else cudaFree(p); }
    p = q; cap = want;
    return e;
  }
  void release() {
    if (p) { if (pool) cudaFreeAsync(p, st); else cudaFree(p); }
    p = nullptr; cap = 0;
  }
  template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

// device-side counters written by the build kernels
enum { CTR_LEAVES = 0, CTR_PTS = 1, CTR_SLOTS = 2, CTR_VALID = 3, CTR_NFIN = 4, CTR_JOB = 5, CTR_BIG = 6, CTR_COUNT = 8 };

struct GridBuffers {
  DevBuf tgt;          // float4[n]       target points (device copy or alias source)
  DevBuf cell_of;      // int32[n]        cell of each point (-1 = non-finite)
  DevBuf rank_of;      // int32[n]        arbitrary unique rank of the point inside its cell
  DevBuf list;         // int32[n]        per-leaf point indices, arbitrary order
  DevBuf sorted_idx;   // int32[n]        per-leaf point indices, ascending (input order)
  DevBuf tgt_sorted;   // float2[n]       target (x, y) in bucket order (1-NN scans read this contiguously)
  DevBuf leaf_range;   // int2[n]         per leaf (start, n)
  DevBuf nn_cnt;       // int32[nn cells] fine nearest-neighbour lattice: counts (build only)
  DevBuf nn_range;     // int2[nn cells]  (start, n)
  DevBuf nn_pts;       // float2[n]       target (x, y) in fine-bucket order
  DevBuf slot;         // int32[padded]   cell -> record slot (written for tree cells only; never initialised elsewhere)
  DevBuf cen;          // float2[padded]  probe table: float32 centroid of tree cells, NaN elsewhere
  DevBuf occ;          // uint32[padded/32] dilated occupancy bitmap (3x3 block contains a tree cell)
  DevBuf nbr;          // uint16[padded]   which cells of the 3x3 block around a cell are tree cells (bit (dj+1)*4 + di+1); on demand
  DevBuf leaf_id;      // int32[padded]   per-cell point count during the build, then leaf id + 1 (0 = empty)
  DevBuf leaf_cell;    // int32[n]        per leaf: position in the shared padded tables
  DevBuf leaf_pair;    // int32[n]        per leaf: which grid it belongs to
  DevBuf tile_hist;    // int32[tiles][padded] per-tile cell histograms / offsets (stable counting sort of small dense grids)
  DevBuf big_list;     // int32[]         leaves with more than FINALIZE_BIG_LEAF points (reduced by a warp each)
  DevBuf dims;         // PairDims[n_grids]
  DevBuf pair_off;     // int64[n_grids+1] target point ranges (batched pairs)
  DevBuf leaf_n;       // int32[n]
  DevBuf leaf_start;   // int32[n]
  DevBuf leaf_nr;      // int32[n]        PCL nr_points after pass 2 (n or -1)
  DevBuf leaf_mean;    // double2[n]
  DevBuf leaf_icov;    // double4[n] as 4 doubles
  DevBuf leaf_cen;     // float2[n]
  DevBuf recs;         // CellRec[n]      compact records (n >= min_points)
  DevBuf counters;     // int32[CTR_COUNT] + bounds int32[4]
  // incremental target (ndt_set_target_incremental): per-cell running sums of the settled prefix of the cloud, the state
  // of every cell after the last call, work lists
  DevBuf inc_acc;      // CellAcc[padded]  in-order sums over the settled prefix
  DevBuf inc_status;   // uint8[padded]    bit0 occupied, bit1 tree cell, bit2 valid, bit3 has a record slot, bit7 tree status changed
  DevBuf inc_mark;     // int32[padded]    epoch marks (distinct-cell lists without clearing)
  DevBuf inc_lists;    // int32[4][padded] this call's batch cells | the previous call's | union (rotating)
  DevBuf inc_cellof;   // int32[5][cap]    per batch point: cell, rank in its cell, order; per batch cell: count, start
  DevBuf inc_lid;      // int2[padded]     (epoch, local id) of the cells of the current batch
  DevBuf inc_cnt;      // int32[8]         list lengths
};

struct GridDims {
  int32_t min_bx = 0, min_by = 0, div_x = 0, div_y = 0;
  float leaf = 1.f, inv_leaf = 1.f, r2 = 1.f;
  int64_t n_cells = 0;
  int64_t n_tgt = 0;       // points handed to set_target
  int32_t nn_f = 0, nn_min_bx = 0, nn_min_by = 0, nn_div_x = 0, nn_div_y = 0;   // fine 1-NN lattice (nn_f = 0: none)
  float nn_leaf = 1.f, nn_inv_leaf = 1.f;
};

struct Handle {
  ndt_params prm{};
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  std::string err;
  int64_t launches = 0;
  float last_ms = 0.f;
  int sm_count = 148;
  int max_smem_optin = 0;
  bool coop_launch = false;      // cudaDevAttrCooperativeLaunch
  cudaMemPool_t pool = nullptr;  // private stream-ordered pool behind every DevBuf of this handle

  GridBuffers gb;
  GridDims gd;
  bool have_grid = false;
  // incremental target state (valid while inc_ok)
  bool inc_ok = false;           // the tables were produced by / are ready for incremental updates
  bool inc_active = false;       // ... and the last call WAS an incremental update (ordered buckets are stale)
  int64_t inc_m = 0;             // points [0, inc_m) are folded into inc_acc
  int32_t inc_epoch = 0;
  float inc_mn[2] = {0.f, 0.f}, inc_mx[2] = {0.f, 0.f};   // exact bounds of the finite points among [0, inc_m)
  int64_t inc_nfin = 0;          // finite points among [0, inc_m)
  int inc_list_prev = 2;         // which of the list buffers holds the previous call's tail cells
  bool have_nbr = false;         // gb.nbr matches the current tables (derived from gb.cen on demand: ensure_nbr)
  int64_t nbr_cells = 0;         // padded entries of all grids in the shared tables (what ensure_nbr covers)
  int nbr_grids = 0;             // entries of gb.dims
  bool have_readback = false;    // the per-leaf read-back tables belong to this grid (false on an imported replica)
  bool grid_has_points = false;  // target points + 1-NN buckets present (false on a replica imported without NDT_BLOB_POINTS)
  int64_t tgt_on_device = 0;     // leading points of gb.tgt that still hold the last ndt_set_target cloud (0: unknown)
  int32_t h_counters[CTR_COUNT] = {0};   // host copy after the last build

  DevBuf src;          // float4[ns]
  int64_t ns = 0;
  bool have_src = false;

  DevBuf scratch;      // misc device scratch (eval partials, guesses, results staging)
  DevBuf scratch2;
  DevBuf pair_tgt_next, pair_raw_next;   // ndt_match_pairs, host clouds: the batch being uploaded while the current one is matched
  cudaStream_t copy_stream = nullptr;    // its upload stream (created on first use)
  cudaEvent_t ev_up[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
  DevBuf stage;        // 4 KB persistent staging (single pose / single result)
  DevBuf io;           // host<->device staging of batched poses / results
  bool ms_pending = false;  // last_ms not yet resolved (async device-space call)
  int32_t *pinned_ctr = nullptr; // 256 B of pinned host memory for the small counter read-backs of the grid build
  void *pinned = nullptr;       // pinned host staging
  size_t pinned_cap = 0;
  bool timing = true;           // record ev0/ev1 around kernels
  bool pending = false;         // an incremental target update was queued without waiting (finish_pending)
  int64_t pending_nfin = 0;
};

// grid_build.cu
// n_stable >= 0: the caller promises that the first n_stable points stay a prefix of future targets; the build then also
// prepares the state ndt_set_target_incremental continues from
int grid_build(Handle *h, const float *xyzw, int64_t n, int memspace, int64_t n_same = 0, int64_t n_stable = -1);
int grid_build_incremental(Handle *h, const float *xyzw, int64_t n, int64_t n_same, int64_t n_stable, int memspace, bool defer = false);
int finish_pending(Handle *h);
// shared by ndt_set_target (one grid) and ndt_match_pairs (one grid per pair): target points are in gb.tgt,
// geometry in gb.dims (device), point ranges in gb.pair_off (device, n_grids + 1 entries; unused for one grid)
int grid_build_tables(Handle *h, int64_t n, int n_grids, int64_t total_pad, int max_h);
int pairs_prepare(Handle *h, int64_t n_pairs, int64_t *total_pad, int *max_h);
int grid_cell_index(Handle *h, const float *xyzw, int64_t n, int memspace, int32_t *idx_out);
GridView grid_view(const Handle *h);
// neighbour masks for the batch kernels (k_align_warp, k_eval_warp, k_align_pairs), derived from the probe table once per grid
int ensure_nbr(Handle *h);
MatchParams match_params(const Handle *h, bool want_fitness);

// match_kernels.cu
int launch_eval(Handle *h, const double *d_poses, int64_t n, int want_hessian, double *d_out14, int64_t *d_pairs);
int launch_align(Handle *h, const double *d_guesses, int64_t n, ndt_result *d_results, bool want_fitness);
int launch_best_of(Handle *h, const ndt_result *d_results, int64_t n, int64_t *d_best_index, ndt_result *d_best);
int launch_voxel_filter(Handle *h, const float4 *d_in, int64_t n, float leaf, float4 *d_out, int32_t *d_nout);
// batched scan pairs: per-pair source filter (writes PairDims::ns) and the persistent warp-per-pair matcher
int launch_pairs_filter(Handle *h, const float4 *d_in, float4 *d_out, int64_t n_pairs, float leaf);
int launch_align_pairs(Handle *h, const float4 *d_src, const double *d_guesses, int64_t n_pairs, ndt_result *d_results,
                       bool want_fitness);

// helpers (capi.cu)
int set_err(Handle *h, int code, const char *what, cudaError_t e = cudaSuccess);
int ensure_pinned(Handle *h, size_t bytes);

#define NDT_CUDA(h, call)                                                   \
  do {                                                                      \
    cudaError_t _e = (call);                                                \
    if (_e != cudaSuccess) return ::ndt::set_err((h), NDT_ERR_CUDA, #call, _e); \
  } while (0)

}  // namespace ndt
