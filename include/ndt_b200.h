/*
 * ndt_b200.h -- C ABI of the B200-native 2-D NDT hot path (grid build + scan matching).
 *
 * This is the drop-in boundary for the one data-parallel path of hibikid39/ndt_slam:
 * everything `PoseEstimator::estimatePose` asks of PCL
 *   [REF src/PoseEstimator.cpp:6-29, 43-56; include/ndt_slam/PoseEstimator.h:19-31, 77-83]
 * i.e. pcl::NormalDistributionsTransform::{setResolution,setStepSize,setTransformationEpsilon,
 * setMaximumIterations,setInputTarget,setInputSource,align,getFinalTransformation,hasConverged,
 * getFitnessScore,getTransformationProbability,computeHessian} and pcl::ApproximateVoxelGrid::filter.
 * The reference has no FFI of its own (it links PCL directly, CMakeLists.txt:21,104-108), so the
 * entry points below are what a maintainer would bind from PoseEstimator.{h,cpp}; see
 * INTEGRATION.md for the binding.
 *
 * Conventions: extern "C", plain pointers and sizes, POD structs, no exceptions cross the ABI.
 * Every call returns NDT_OK (0) or a negative ndt_status; the message is ndt_last_error(h).
 * The caller owns every buffer it passes; the handle owns all device memory. A handle is bound
 * to one CUDA device and one stream and is NOT thread-safe (the reference is single-threaded,
 * SlamLauncher.cpp:112-138). There is no CPU fallback: without a CUDA device ndt_create fails.
 *
 * Points are float4-strided {x, y, z(=0), pad}: the in-memory layout of pcl::PointXYZ
 * (16 bytes), exactly what PoseEstimator::setScanPair fills (PoseEstimator.h:91-104).
 * Poses are (x [m], y [m], yaw [rad]) in fp64.
 */
#ifndef NDT_B200_H_
#define NDT_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ndt_handle_s *ndt_handle;

typedef enum {
  NDT_OK = 0,
  NDT_ERR_ARG = -1,      /* bad argument / call order            */
  NDT_ERR_CUDA = -2,     /* a CUDA runtime call failed           */
  NDT_ERR_NO_DEVICE = -3,/* no usable CUDA device (no fallback)  */
  NDT_ERR_CAPACITY = -4, /* grid or batch exceeds what fits      */
  NDT_ERR_STATE = -5     /* target/source not set                */
} ndt_status;

/* Where a caller buffer lives. */
typedef enum { NDT_MEM_HOST = 0, NDT_MEM_DEVICE = 1 } ndt_memspace;

/* Behaviour switches that pin the restated PCL 1.10.0 semantics.
 * The default is NDT_QUIRKS_PCL_1_10. */
enum {
  NDT_QUIRK_COV_INIT_IDENTITY  = 1 << 0, /* Leaf::cov_ starts at Identity (adds I/n)          */
  NDT_QUIRK_COV_SCALE_NM1_N    = 1 << 1, /* biased single-pass cov then *= (n-1)/n            */
  NDT_QUIRK_MT_INTERVAL_LT0    = 1 << 2, /* interval_converged = (step_max - step_min) < 0    */
  NDT_QUIRK_ANGLE_SNAP         = 1 << 3, /* |yaw| < 10e-5 -> cos=1, sin=0 in the derivatives  */
  NDT_QUIRK_TRANSFORM_SSE_ORDER= 1 << 4, /* x' = c*x + ((-s)*y + tx) instead of (c*x+(-s)*y)+tx */
  NDT_QUIRKS_PCL_1_10 = NDT_QUIRK_COV_INIT_IDENTITY | NDT_QUIRK_COV_SCALE_NM1_N |
                        NDT_QUIRK_MT_INTERVAL_LT0 | NDT_QUIRK_ANGLE_SNAP
};

/* Replaces the ndt.set*() calls in the PoseEstimator constructor (PoseEstimator.h:77-83) plus
 * the PCL-internal constants the reference never touches. */
typedef struct {
  float  resolution;      /* ndt.setResolution((float)Resolution)        h:81 */
  double step_size;       /* ndt.setStepSize                             h:79 */
  double trans_eps;       /* ndt.setTransformationEpsilon                h:77 */
  int32_t max_iter;       /* ndt.setMaximumIterations                    h:83 */
  double outlier_ratio;   /* PCL default 0.55                                 */
  int32_t min_points;     /* VoxelGridCovariance min_points_per_voxel_ = 6    */
  double eig_mult;        /* min_covar_eigvalue_mult_ = 0.01                  */
  int32_t quirks;         /* NDT_QUIRK_* bit set                              */
  int32_t device;         /* CUDA device ordinal                              */
  void  *stream;          /* cudaStream_t to launch on, or NULL = own stream  */
  /* Behaviour / tuning knobs of the batched entry points. 0 = library default everywhere (ndt_params_default). */
  int32_t align_skip_fitness; /* 1: ndt_align skips the getFitnessScore pass (result.fitness = NaN)        */
  int32_t pairs_schedule;     /* ndt_match_pairs: NDT_PAIRS_AUTO / NDT_PAIRS_WARP / NDT_PAIRS_CTA           */
  int64_t pairs_batch_points; /* ndt_match_pairs: max target points per internal batch; 0 = 32,000,000     */
  int32_t align_team;         /* ndt_align_batch (n >= 64): warps that share one match, 1 / 2 / 4 / 8;
                                 0 = chosen from the batch size (few matches per resident warp: larger teams,
                                 shorter latency; many: one warp per match, best throughput)                */
  int32_t reserved0;
} ndt_params;

/* ndt_params.pairs_schedule: how ndt_match_pairs spreads pairs over the GPU. AUTO picks by batch size. */
enum { NDT_PAIRS_AUTO = 0, NDT_PAIRS_WARP = 1, NDT_PAIRS_CTA = 2 };

/* One evaluation of the NDT objective: pcl::NDT::computeDerivatives (reached through
 * ndt.align, PoseEstimator.cpp:28). H is row-major 3x3 over (x, y, yaw). */
typedef struct {
  double score;
  double grad[3];
  double hess[9];
  int64_t n_pairs;        /* (point, cell) hits that passed the radius test */
} ndt_eval_out;

/* Everything estimatePose reads back after ndt.align (PoseEstimator.cpp:29, 43-56). */
typedef struct {
  double pose[3];         /* final parameter vector p = (x, y, yaw)                       */
  float  T[16];           /* getFinalTransformation(), column-major like Eigen::Matrix4f  */
  double score;           /* NDT score at the final pose                                  */
  double trans_prob;      /* getTransformationProbability() = score / N_source            */
  double fitness;         /* getFitnessScore(): mean squared 1-NN distance                */
  double hess[9];         /* getHessian() rows/cols {x, y, yaw}, row-major                */
  int32_t converged;      /* hasConverged()                                               */
  int32_t iters;          /* nr_iterations_ (outer Newton iterations)                     */
  int32_t evals;          /* objective passes the reference makes (computeDerivatives + computeHessian) */
  int32_t passes_run;     /* the ones run on the device: a line-search trial at the very pose of the
                             previous trial returns the same numbers and is served from them */
  int64_t point_evals;    /* source points x evals, counted on the device                 */
} ndt_result;

typedef struct {
  int32_t min_b[2];       /* VoxelGridCovariance min_b_ (x, y)             */
  int32_t div_b[2];       /* div_b_ (x, y)                                 */
  int64_t n_points;       /* target points accepted                        */
  int32_t n_leaves;       /* occupied cells (any count)                    */
  int32_t n_slots;        /* cells with n >= min_points (kd-tree members)  */
  int32_t n_valid;        /* of those, cells that passed the eigen checks  */
  int32_t reserved;       /* 1 if the last target call was an incremental update   */
} ndt_grid_info;

/* ---- lifetime ------------------------------------------------------------------------- */
int ndt_params_default(ndt_params *p);
int ndt_create(const ndt_params *p, ndt_handle *out);
int ndt_destroy(ndt_handle h);
const char *ndt_last_error(ndt_handle h);      /* h may be NULL: last create error */
const char *ndt_version(void);

/* ---- grid build: ndt.setInputTarget(target_cloud)  (PoseEstimator.cpp:19) --------------- */
int ndt_set_target(ndt_handle h, const float *xyzw, int64_t n, int memspace);
/* Same, for a target that only changed at its end since the previous ndt_set_target* call on this handle (a map that
 * grows scan by scan: PointCloudMap::makeLocalMap [REF src/PointCloudMap.cpp:119-134]): the caller promises that the
 * first n_same points are identical to the previous call's, so only points [n_same, n) are staged and copied to the
 * device; the grid itself is rebuilt in full and is identical to ndt_set_target's. n_same = 0 is ndt_set_target. */
int ndt_set_target_prefix(ndt_handle h, const float *xyzw, int64_t n, int64_t n_same, int memspace);
/* The same for a map with a settled part (PointCloudMap's local map = previous sub-map + the current sub-map's thinned
 * prefix + a provisional tail that is replaced every scan): n_same as above; n_stable <= n: the caller promises that the
 * first n_stable points stay a prefix of every later target handed to this handle. The library then keeps per-cell running
 * sums of the settled prefix ON THE DEVICE, folds the points that became settled since the last call into them, adds the
 * tail on top and re-derives only the cells any of this touched (plus the cells of the previous tail) -- in the reference's
 * input order, so every cell is bit-identical to what ndt_set_target builds from the whole cloud. Falls back to a full
 * rebuild (which also (re)initialises the running sums) whenever that is not possible: first call, grid bounds moved, the
 * settled prefix shrank, device input. After an incremental update ndt_grid_readback returns NDT_ERR_STATE (the per-leaf
 * read-back tables are not maintained); the matcher, the fitness score and ndt_get_grid_info see the same grid. */
int ndt_set_target_incremental(ndt_handle h, const float *xyzw, int64_t n, int64_t n_same, int64_t n_stable, int memspace);
/* The same update, queued: returns once the new points are copied to a staging buffer and the kernels are launched, so the
 * host can prepare the next scan (resampling, source filter) while the device brings the grid up to date. xyzw may be
 * reused at once; the next call on the handle, whatever it is, waits for the update first. What cannot be continued
 * incrementally is rebuilt in full before the call returns, as above. (ScanMatcher::growMap calls this right after
 * makeLocalMap, src/ScanMatcher.cpp:93-117, instead of leaving the whole cost to the next estimatePose.) */
int ndt_set_target_incremental_async(ndt_handle h, const float *xyzw, int64_t n, int64_t n_same, int64_t n_stable, int memspace);
int ndt_get_grid_info(ndt_handle h, ndt_grid_info *info);
/* Leaves in ascending cell-index order (the std::map order of PCL's leaves_), for parity
 * checks: cell index, nr_points (-1 = failed eigen check), mean[2], icov[4] (xx,xy,yx,yy),
 * centroid[2]. Arrays may be NULL. Returns the leaf count through *n_out. */
int ndt_grid_readback(ndt_handle h, int64_t cap, int32_t *cell_idx, int32_t *nr_points,
                      double *mean2, double *icov4, float *centroid2, int64_t *n_out);
/* Cell index of each point exactly as VoxelGridCovariance pass 1 computes it (device kernel). */
int ndt_cell_index(ndt_handle h, const float *xyzw, int64_t n, int memspace, int32_t *idx_out);

/* ---- source: ndt.setInputSource(filtered_cloud)  (PoseEstimator.cpp:17) ----------------- */
int ndt_set_source(ndt_handle h, const float *xyzw, int64_t n, int memspace);
/* pcl::ApproximateVoxelGrid::filter (PoseEstimator.cpp:6-10; PointCloudMap.cpp:4-13). The
 * algorithm is a sequential 512-entry hash history, so it runs as one device thread per cloud.
 * out must hold n points; *n_out receives the filtered count. */
int ndt_approx_voxel_filter(ndt_handle h, const float *xyzw, int64_t n, float leaf,
                            int memspace, float *out_xyzw, int64_t *n_out);

/* ---- objective: computeDerivatives / computeHessian ------------------------------------ */
int ndt_eval(ndt_handle h, const double pose[3], int want_hessian, ndt_eval_out *out);
/* Score / gradient / Hessian for n poses of the same source (relocalisation sweep).
 * out14 receives n x 14 doubles: score, g[3], H[9], n_pairs. poses/out live in `memspace`. */
int ndt_eval_batch(ndt_handle h, const double *poses, int64_t n, int want_hessian,
                   int memspace, double *out14);

/* ---- matching: ndt.align(output, init_guess) + the getters ------------------------------ */
int ndt_align(ndt_handle h, const double guess[3], ndt_result *out);
/* n independent matches of the current source against the current grid, one per guess.
 * guesses: n x 3 doubles, results: n ndt_result, both in `memspace`.
 * want_fitness != 0: every result carries getFitnessScore() (one exact 1-NN pass per match); 0: result.fitness = NaN
 * (relocalisation ranks tens of thousands of hypotheses by score and re-runs ndt_align on the winner). */
int ndt_align_batch(ndt_handle h, const double *guesses, int64_t n, int memspace, int want_fitness,
                    ndt_result *results);
/* Index of the best result of an ndt_align_batch (max score among converged; lowest index on ties), device-side. */
int ndt_best_of(ndt_handle h, const ndt_result *results, int64_t n, int memspace,
                int64_t *best_index, ndt_result *best);
/* The same over several handles (one per GPU, each holding its shard's results in ITS device memory): every GPU
 * reduces its own shard, the host compares the n_handles winners (ties: lowest handle, then lowest index).
 * *best_handle = -1 when no result converged anywhere. Only n_handles x sizeof(ndt_result) bytes leave the GPUs. */
int ndt_best_of_multi(const ndt_handle *handles, const ndt_result *const *device_results, const int64_t *counts,
                      int n_handles, int *best_handle, int64_t *best_index, ndt_result *best);

/* n independent scan-pair matches (loop-closure verification): pair i matches
 * src[src_off[i] .. src_off[i+1]) against a grid built from tgt[tgt_off[i] .. tgt_off[i+1]).
 * source_leaf > 0 applies the ApproximateVoxelGrid source filter first (as estimatePose does).
 * This is n_pairs x PoseEstimator::estimatePose [REF src/PoseEstimator.cpp:4-69] in one call: all grids are
 * built by one pass of the grid-build kernels (one shared padded table, one slice per pair) and one
 * persistent kernel matches every pair (one warp per pair); results carry the fitness score.
 * src_off / tgt_off (n_pairs + 1 entries, starting at 0) are always HOST arrays; the points, guesses and
 * results live in `memspace`. The handle's own target / source (ndt_set_target / ndt_set_source) are
 * overwritten by this call. */
int ndt_match_pairs(ndt_handle h, const float *src_xyzw, const int64_t *src_off,
                    const float *tgt_xyzw, const int64_t *tgt_off, const double *guesses,
                    int64_t n_pairs, float source_leaf, int memspace, ndt_result *results);

/* The same call for planar clouds given as (x, y) float pairs, 8 bytes per point instead of 16 (z = 0 like every
 * LPoint2D the reference turns into a pcl::PointXYZ, PointCloudMap.cpp:72-86): halves the host-to-device upload
 * that bounds ndt_match_pairs on host inputs. src_xy / tgt_xy: 2 floats per point; offsets count points. */
int ndt_match_pairs_xy(ndt_handle h, const float *src_xy, const int64_t *src_off,
                       const float *tgt_xy, const int64_t *tgt_off, const double *guesses,
                       int64_t n_pairs, float source_leaf, int memspace, ndt_result *results);

/* ---- replication of a finished grid to other GPUs (one NVLink broadcast, no collectives
 *      per iteration): export to / import from a flat device blob ------------------------- */
/* flags: NDT_BLOB_POINTS also carries the target points and their 1-NN buckets (needed only for getFitnessScore on the
 * replica: more than half of the bytes); without it a replica matches and evaluates exactly like the original but
 * reports fitness = NaN. The per-leaf read-back tables of ndt_grid_readback are never replicated: on a replica that
 * call returns NDT_ERR_STATE. ndt_grid_import validates the blob (magic, offsets, sizes, resolution == the handle's). */
enum { NDT_BLOB_POINTS = 1 };
int ndt_grid_blob_size(ndt_handle h, int flags, int64_t *bytes);
int ndt_grid_export(ndt_handle h, int flags, void *device_blob, int64_t bytes);
int ndt_grid_import(ndt_handle h, const void *device_blob, int64_t bytes);
/* handles[0] holds a finished grid; copy it to handles[1 .. n_handles) (other GPUs of the box, or the same GPU):
 * export on the source device, one peer copy per destination over NVLink (cudaMemcpyPeerAsync), import there. This is
 * the whole multi-GPU data path of the batched workloads: the grid is replicated once, hypotheses / pairs are sharded
 * by the caller, no collective per iteration. A C++ host needs nothing else (no NCCL, no torch). */
int ndt_replicate_grid(const ndt_handle *handles, int n_handles, int flags);
/* Return the device memory the handle's private pool caches beyond its live buffers to the driver. */
int ndt_trim(ndt_handle h);

/* ---- device memory for callers without a CUDA runtime of their own (a plain C / C++ host driving the
 *      NDT_MEM_DEVICE forms of the batched calls): buffers live on the handle's device, copies are ordered on
 *      its stream; upload / download return when the copy is done --------------------------------------- */
int ndt_alloc(ndt_handle h, int64_t bytes, void **device_ptr);
int ndt_free(ndt_handle h, void *device_ptr);
int ndt_upload(ndt_handle h, void *device_dst, const void *host_src, int64_t bytes);
int ndt_download(ndt_handle h, void *host_dst, const void *device_src, int64_t bytes);

/* ---- instrumentation ------------------------------------------------------------------- */
/* Number of kernels this handle has launched since creation. */
int ndt_launch_count(ndt_handle h, int64_t *n);
/* Device milliseconds of the last ndt_set_target / ndt_align* / ndt_eval* / ndt_match_pairs
 * call, measured with CUDA events on the handle's stream around its kernels only. */
int ndt_last_kernel_ms(ndt_handle h, float *ms);
int ndt_synchronize(ndt_handle h);

#ifdef __cplusplus
}
#endif
#endif /* NDT_B200_H_ */
