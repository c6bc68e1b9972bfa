/*
 * quatro_b200.h -- C-ABI of the B200-native global-registration hot path.
 *
 * This is the drop-in boundary.  Every entry point replaces one stage
 * boundary of the reference (url-kaist/Quatro, paths relative to the reference root):
 *
 *   qb200_patchwork             <- PatchWork<PointT>::estimate_ground   include/patchwork.hpp:329-455 (pre-processing, 8f-1)
 *   qb200_segment_cloud         <- ImageProjection::segmentCloud        include/imageProjection.hpp:273-294 (pre-processing, 8f-1)
 *   qb200_voxelize              <- voxelize<T>()                    include/quatro.hpp:49-57
 *   qb200_compute_fpfh          <- FPFHEstimation::computeFPFHFeatures  src/teaser_utils/fpfh.cc:44-75
 *   qb200_match                 <- Matcher::calculateCorrespondences    include/teaser_utils/feature_matcher.h:42-74
 *                                  + Matcher::advancedMatching          src/teaser_utils/feature_matcher.cc:77-265
 *   qb200_build_graph           <- Quatro::computeTIMs + solveForScale + inlier_graph_.addEdge loop
 *                                  include/quatro.hpp:307-386, 784-789
 *   qb200_max_clique            <- teaser::MaxCliqueSolver::findMaxClique   src/graph.cc:12-130
 *   qb200_solve_pose            <- chain TIMs + GNC-TLS yaw + COTE      include/quatro.hpp:817-936
 *   qb200_solve_correspondences <- Quatro::computeTransformation(Eigen::Matrix4d&)  include/quatro.hpp:769-936
 *   qb200_match_and_pack        <- FPFHManager::setFeaturePair          include/fpfh_manager.hpp:98-153
 *   qb200_register_pair/_batch  <- examples/run_global_registration.cpp:206-246 (voxelize .. computeTransformation)
 *   qb200_register_batch_sharded / _rank  <- the same loop over a list of pairs, sharded over the GPUs of one box
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no C++/torch types.  All pointers are HOST pointers
 *     unless a function takes a qb200_mem_kind (then QB200_MEM_DEVICE pointers are accepted).
 *   - Points are 16-byte records {x, y, z, w} (pcl::PointXYZ layout, include/utility.h:109;
 *     KITTI .bin records, examples/run_global_registration.cpp:384-399).
 *   - Every function returns a qb200_status (0 = ok, <0 = bad argument / CUDA failure,
 *     >0 = per-pair degenerate result).  Nothing throws across this boundary.
 *   - There is NO CPU fallback: if no CUDA device is usable qb200_create() fails with
 *     QB200_ERR_NO_DEVICE and no other entry point can be called.
 *   - One handle = one device + one stream; a handle is not thread-safe, several handles are.
 */
#ifndef QUATRO_B200_H_
#define QUATRO_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QB200_VERSION 100

typedef enum qb200_status {
  QB200_OK = 0,
  /* >0: the call worked, the pair is degenerate (mirrors solution_.valid=false, quatro.hpp:809-813) */
  QB200_DEGENERATE_CLIQUE = 1,   /* max clique size <= 1 */
  QB200_DEGENERATE_INPUT = 2,    /* fewer than 2 correspondences / empty cloud */
  QB200_CAPACITY_EXCEEDED = 3,   /* a per-pair buffer (voxels / correspondences) overflowed; result invalid */
  /* <0: errors */
  QB200_ERR_BAD_ARG = -1,
  QB200_ERR_NO_DEVICE = -2,
  QB200_ERR_CUDA = -3,
  QB200_ERR_UNSUPPORTED = -4,    /* e.g. use_crosscheck = 0, libnccl absent */
  QB200_ERR_VOXEL_OVERFLOW = -5  /* dx*dy*dz > INT_MAX: PCL warns and returns the input unfiltered */
} qb200_status;

/* qb200_result.flags */
enum { QB200_FLAG_CLIQUE_TRUNCATED = 1 /* PMC_EXACT stopped at max_clique_node_limit: the clique is the best found, not proven maximum */ };
#define QB200_DEFAULT_CLIQUE_NODE_LIMIT 262144

/* Quatro::INLIER_SELECTION_MODE, include/quatro.hpp:184-189 */
enum { QB200_PMC_EXACT = 0, QB200_PMC_HEU = 1, QB200_KCORE_HEU = 2, QB200_INLIER_NONE = 3 };
/* Params::cote_mode, include/quatro.hpp:209 */
enum { QB200_COTE_MEDIAN = 0, QB200_COTE_WEIGHTED_MEAN = 1 };
typedef enum qb200_mem_kind { QB200_MEM_HOST = 0, QB200_MEM_DEVICE = 1 } qb200_mem_kind;

/* One POD for every knob on the path.  Defaults (qb200_default_params) = config/params.yaml:22-44
 * + the hard-coded matcher flags of include/fpfh_manager.hpp:126-127. */
typedef struct qb200_params {
  /* front end */
  float voxel_size;              /* 0.3   voxelize leaf                      params.yaml:22 */
  float normal_radius;           /* 0.5                                      params.yaml:24 */
  float fpfh_radius;             /* 0.75                                     params.yaml:25 */
  float grid_cell;               /* 0 -> (1 + 2^-9) fpfh_radius. Cell of the neighbour-search lattice; the neighbour SETS do not
                                    depend on it, only the (cell,index) order in which they are accumulated. */
  float tuple_scale;             /* 0.95                                     fpfh_manager.hpp:127 */
  int32_t use_crosscheck;        /* 1 */
  int32_t use_tuple_test;        /* 1 */
  int32_t tuple_trials_per_corr; /* 100                                      feature_matcher.cc:195 */
  int32_t skip_flagged;          /* 1: voxelize drops points with w < 0 (stand-in for the reference's
                                    ground/sub-cluster removal, which is out of scope; PCL itself only
                                    drops non-finite points) */
  int32_t reserved0;
  uint64_t seed;                 /* tuple-test RNG seed (replaces srand(time(NULL)), feature_matcher.cc:189) */
  /* solver, include/quatro.hpp:202-268 */
  double noise_bound;            /* 0.3   */
  double cbar2;                  /* 1.0   */
  double rot_noise_bound;        /* 0 -> 2*noise_bound (the value the reference's function-local static
                                    latches on its first call, quatro.hpp:469-470 after :851) */
  double cote_noise_bound;       /* 0.3   Quatro::noise_bound_ ctor constant, quatro.hpp:115,601 */
  double rotation_gnc_factor;    /* 1.4   */
  double rotation_cost_threshold;/* 1.1e-4 */
  double kcore_heuristic_threshold; /* 0.5 */
  int32_t rotation_max_iterations;  /* 50 */
  int32_t inlier_selection_mode;    /* QB200_PMC_HEU */
  int32_t cote_mode;                /* QB200_COTE_MEDIAN */
  int32_t using_rot_inliers_when_estimating_cote; /* 0 */
  int32_t use_pre_estimated_RyRx;   /* 0 */
  int32_t max_clique_node_limit;    /* PMC_EXACT: branch-and-bound nodes per pair before the search returns its best clique so far
                                       (deterministic stand-in for pmc's wall-clock time_limit, src/graph.cc:44); 0 -> QB200_DEFAULT_CLIQUE_NODE_LIMIT */
  double RyRx[9];                   /* row-major 3x3, setPreEstaimatedRyRx, quatro.hpp:276-279 */
} qb200_params;

/* Handle configuration: device and per-pair capacities (device workspaces are sized once). */
typedef struct qb200_config {
  int32_t device;            /* CUDA ordinal */
  int32_t max_batch_slots;   /* pairs resident in one wave of the batch pipeline (default 64) */
  int32_t max_raw_points;    /* per cloud (default 131072) */
  int32_t max_voxel_points;  /* per cloud (default 16384; multiple of 128) */
  int32_t max_corr;          /* per pair  (default 4096; multiple of 32, <= 8192; the pose solver holds cliques of <= 4096) */
  int32_t reserved[3];
} qb200_config;

/* Fixed-size per-pair record (also the unit gathered across GPUs). */
typedef struct qb200_result {
  int32_t valid;             /* solution_.valid */
  int32_t status;            /* qb200_status for this pair */
  int32_t n_src_vox, n_tgt_vox;
  int32_t n_mutual;          /* mutual nearest neighbours before the tuple test */
  int32_t n_corr;            /* L: correspondences after tuple test + dedupe */
  int32_t max_core;          /* pmc max core number (graph.cc:60) */
  int32_t clique_size;
  int32_t gnc_iters;
  int32_t n_rot_inliers;     /* getNumRotaionInliers */
  int32_t n_final_inliers;
  int32_t flags;             /* QB200_FLAG_* bits */
  int64_t n_edges;
  double cost;               /* Quatro::cost_ */
  double T[16];              /* column-major 4x4 (Eigen::Matrix4d memory order); identity when !valid */
} qb200_result;

typedef struct qb200_pair {
  const float* src; /* n_src x 4 floats */
  const float* tgt;
  int32_t n_src, n_tgt;
} qb200_pair;

typedef struct qb200_handle qb200_handle;

void qb200_default_params(qb200_params* p);
void qb200_default_config(qb200_config* c);
int qb200_version(void);

int qb200_create(const qb200_config* cfg, qb200_handle** out);
void qb200_destroy(qb200_handle* h);
/* Run on a caller-owned CUDA stream (cudaStream_t as void*); NULL restores the handle's own stream. */
int qb200_set_stream(qb200_handle* h, void* cuda_stream);
const char* qb200_last_error(const qb200_handle* h);
/* Number of kernels this handle has launched so far (bench.py's gpu_launches). */
int64_t qb200_launch_count(const qb200_handle* h);

/* --- stage boundaries (host pointers) ------------------------------------------------------ */

/* pcl::VoxelGrid semantics: centroid per occupied leaf cube, output ordered by ascending
 * (k,j,i) cell; non-finite points (and w<0 points when skip_flagged) are dropped. */
int qb200_voxelize(qb200_handle* h, const float* pts4, int32_t n, float leaf, int32_t skip_flagged,
                   float* out4, int32_t cap, int32_t* n_out);

/* --- pre-processing before the path: ground removal ------------------------------
 * qb200_patchwork <- PatchWork<PointT>::estimate_ground, include/patchwork.hpp:329-455 (concentric zone model
 * :512-543, region-wise ground plane fit :278-324,548-590, ground likelihood estimation :386-440), de-ROS-ed: the
 * parameters the reference reads from the ROS parameter server (patchwork.hpp:50-95, config/patchwork_params.yaml)
 * travel in this POD. */
#define QB200_PW_MAX_ZONES 4
#define QB200_PW_MAX_THRESHOLDS 8
typedef struct qb200_patchwork_params {
  double sensor_height;                   /* 1.723   patchwork_params.yaml:1 */
  double th_seeds;                        /* 0.25 */
  double th_dist;                         /* 0.125 */
  double max_range;                       /* 80.0 */
  double min_range;                       /* 2.7 (= min_ranges_each_zone[0]) */
  double uprightness_thr;                 /* 0.707 */
  double adaptive_seed_selection_margin;  /* -1.1 */
  double global_elevation_threshold;      /* -0.5 */
  double min_ranges_each_zone[QB200_PW_MAX_ZONES];       /* 2.7, 12.3625, 22.025, 41.35 */
  double elevation_thresholds[QB200_PW_MAX_THRESHOLDS];  /* -1.2, -0.9984, -0.851, -0.605 */
  double flatness_thresholds[QB200_PW_MAX_THRESHOLDS];   /* 0.0001, 0.000125, 0.000185, 0.000185 */
  int32_t num_iter;                       /* 3 */
  int32_t num_lpr;                        /* 20 */
  int32_t num_min_pts;                    /* 80 */
  int32_t using_global_elevation;         /* 0 */
  int32_t num_zones;                      /* 4 (the reference's binning is written for exactly four zones, patchwork.hpp:520-539) */
  int32_t num_thresholds;                 /* 4 = rings of interest (elevation_thresholds.size()) */
  int32_t num_sectors_each_zone[QB200_PW_MAX_ZONES];     /* 16, 32, 54, 32 */
  int32_t num_rings_each_zone[QB200_PW_MAX_ZONES];       /* 2, 4, 4, 4 */
} qb200_patchwork_params;
void qb200_default_patchwork_params(qb200_patchwork_params* p);

/* ground4 / nonground4: room for n points each (either may be NULL: only the counts are returned).  Point order of both
 * outputs = the reference's: patches in (zone, ring, sector) order, inside a patch ascending (z, input index); a patch whose
 * plane is rejected hands its ground part, then its non-ground part, to the non-ground output (patchwork.hpp:399-436).
 * Points outside (min_range, max_range], below -1.8 sensor_height, non-finite, or in patches of <= num_min_pts points appear in
 * neither output (as in the reference).  QB200_CAPACITY_EXCEEDED: a patch holds more than 16384 points. */
int qb200_patchwork(qb200_handle* h, const float* pts4, int32_t n, const qb200_patchwork_params* p,
                    float* ground4, int32_t* n_ground, float* nonground4, int32_t* n_nonground);

/* qb200_segment_cloud <- ImageProjection::segmentCloud in "Patchwork" mode, include/imageProjection.hpp:273-294: range-image
 * projection (:308-352), connected components of the range image under the LeGO-LOAM angle criterion (labelComponents, :483-579)
 * and the extraction of the valid segments / outliers (:424-481, getValidSegments :214-216, getOutliers :230-232).
 * The constructor's per-sensor constants (:86-131) travel in this POD; qb200_default_segment_params = "Velodyne-64-HDE",
 * "4CrossNeighbor" (examples/run_global_registration.cpp:53-55). */
enum { QB200_NEIGHBORS_4 = 0, QB200_NEIGHBORS_8 = 1, QB200_NEIGHBORS_4_CROSS = 2 };
typedef struct qb200_segment_params {
  int32_t n_scan;                     /* 64   (<= 64) */
  int32_t horizon_scan;               /* 1800 */
  float ang_res_x;                    /* 360 / 1800 */
  float ang_res_y;                    /* 26.9 / 63 */
  float ang_bottom;                   /* 25.0 */
  float segment_theta;                /* 60 deg in radians */
  int32_t neighbor_mode;              /* QB200_NEIGHBORS_4_CROSS */
  int32_t min_pts_for_subclustering;  /* 30 */
  int32_t segment_valid_point_num;    /* 5 */
  int32_t segment_valid_line_num;     /* 3 */
} qb200_segment_params;
void qb200_default_segment_params(qb200_segment_params* p);

/* valid4 / outlier4: room for n_scan * horizon_scan points each (either may be NULL).  Both outputs are in row-major pixel
 * order; a pixel holds the LAST input point projected into it (:340-346); w = 1. */
int qb200_segment_cloud(qb200_handle* h, const float* pts4, int32_t n, const qb200_segment_params* p,
                        float* valid4, int32_t* n_valid, float* outlier4, int32_t* n_outlier);

/* normals4: n x {nx,ny,nz,curvature}; desc33: n x 33 floats (pcl::FPFHSignature33). Either may be NULL. */
int qb200_compute_fpfh(qb200_handle* h, const float* pts4, int32_t n, float normal_radius,
                       float fpfh_radius, float grid_cell, float* normals4, float* desc33);

/* corr: n_corr x {src_idx, tgt_idx}, sorted lexicographically, unique. */
int qb200_match(qb200_handle* h, const float* src4, int32_t n_src, const float* src_desc33,
                const float* tgt4, int32_t n_tgt, const float* tgt_desc33, const qb200_params* p,
                int32_t* corr, int32_t cap, int32_t* n_corr, int32_t* n_mutual);

/* adj: L rows x words_per_row uint32, bit j of row i set iff edge (i,j); full symmetric matrix.
 * words_per_row >= ceil(L/32).  degree (L) and n_edges may be NULL. */
int qb200_build_graph(qb200_handle* h, const float* a4, const float* b4, int32_t L,
                      double noise_bound, double cbar2, uint32_t* adj, int32_t words_per_row,
                      int32_t* degree, int64_t* n_edges);

/* clique: ascending vertex ids.  kcore (L, pmc's core number + 1) and kcore_order (L) may be NULL. */
int qb200_max_clique(qb200_handle* h, const uint32_t* adj, int32_t L, int32_t words_per_row,
                     int32_t mode, double kcore_heuristic_threshold, int32_t* clique, int32_t* n_clique,
                     int32_t* kcore, int32_t* kcore_order, int32_t* max_core);
/* The same with the PMC_EXACT knobs: node_limit (0 = default) and the QB200_FLAG_* bits of the search (flags may be NULL).
 * PMC_EXACT = the heuristic clique as the incumbent, then a bit-parallel branch and bound with greedy-colouring bounds
 * (src/graph.cc:106-127 -> [EXT] pmc::pmcx_maxclique::search_dense); the clique SIZE is the maximum, membership follows the
 * canonical sequential order of DESIGN.md 5.3 (pmc's own choice among equal maximum cliques depends on thread timing). */
int qb200_max_clique_ex(qb200_handle* h, const uint32_t* adj, int32_t L, int32_t words_per_row,
                        int32_t mode, double kcore_heuristic_threshold, int64_t node_limit, int32_t* clique, int32_t* n_clique,
                        int32_t* kcore, int32_t* kcore_order, int32_t* max_core, int32_t* flags);

/* rotation + translation given the (sorted) clique. inlier_mask (n_clique bytes) may be NULL. */
int qb200_solve_pose(qb200_handle* h, const float* a4, const float* b4, int32_t L,
                     const int32_t* clique, int32_t n_clique, const qb200_params* p,
                     qb200_result* res, uint8_t* rot_inlier_mask, uint8_t* trans_inlier_mask);

/* = Quatro::computeTransformation on matched point pairs. */
int qb200_solve_correspondences(qb200_handle* h, const float* a4, const float* b4, int32_t L,
                                const qb200_params* p, qb200_result* res);

/* voxelized clouds in -> correspondences + matched point copies (FPFHManager::setFeaturePair). */
int qb200_match_and_pack(qb200_handle* h, const float* src4, int32_t n_src, const float* tgt4,
                         int32_t n_tgt, const qb200_params* p, int32_t* corr, float* src_matched4,
                         float* tgt_matched4, int32_t cap, int32_t* n_corr);

/* Batch of precomputed correspondence sets (matched point pairs a[i] <-> b[i], xyzw records): graph -> max clique ->
 * GNC-TLS yaw + COTE for every set, = Quatro::computeTransformation (include/quatro.hpp:769-936) called once per set
 * with setInputSource / setInputTarget already given matched clouds.  Sets are processed in waves of
 * max_batch_slots; kind says where a / b live; results is a host array of n records. */
typedef struct qb200_corr_set {
  const float* a;   /* L x 4 floats (source side) */
  const float* b;   /* L x 4 floats (target side) */
  int32_t L;        /* <= max_corr */
  int32_t reserved;
} qb200_corr_set;
int qb200_solve_batch(qb200_handle* h, const qb200_corr_set* sets, int32_t n_sets, const qb200_params* p,
                      qb200_mem_kind kind, qb200_result* results);

/* Pipelined form of qb200_register_batch for a stream of batches: _enqueue queues the batch and returns (it collects a lane's
 * earlier wave only when it needs that lane again), so the single-warp tail of one batch runs under the PCIe copies and front-end
 * kernels of the next; _flush waits for everything queued and completes the record arrays.  The scans (host kind) and `results` of
 * every queued batch must stay valid until a flush (or qb200_register_batch, = enqueue + flush) returns.  Other entry points flush
 * implicitly. */
int qb200_register_batch_enqueue(qb200_handle* h, const qb200_pair* pairs, int32_t n_pairs, const qb200_params* p, qb200_mem_kind kind,
                                 qb200_result* results);
int qb200_register_batch_flush(qb200_handle* h);

/* raw scans in -> pose out. */
int qb200_register_pair(qb200_handle* h, const float* src4, int32_t n_src, const float* tgt4,
                        int32_t n_tgt, const qb200_params* p, qb200_result* res);

/* Batch of independent pairs.  kind says where pairs[i].src/tgt live; results is a host array.
 * The batch is processed in waves of max_batch_slots pairs.  Batches larger than one wave rotate over up to
 * 4 lanes (each further lane -- same buffers again, own stream -- is allocated on first use) so that one wave's
 * host->device copies and solver tail overlap the other waves' dense kernels; QB200_LANES=n (1..4) in the
 * environment caps the lane count (1 = strictly one wave at a time).  Results never depend on the wave size
 * or the lane. */
int qb200_register_batch(qb200_handle* h, const qb200_pair* pairs, int32_t n_pairs,
                         const qb200_params* p, qb200_mem_kind kind, qb200_result* results);

/* --- scan cache: one scan against many (loop-closure sweeps) and odometry chains -------------------------------------------------
 * FPFHManager keeps the previous target's cloud and descriptors and reuses them as the next source (swapTgt2Src / is_odometry_test_,
 * include/fpfh_manager.hpp:74-77, 111-118) and hands its descriptors out (getObjDescriptor / getSceneDescriptor / getTgtNormals,
 * :161-177).  The cache keeps voxel points, normals and FPFH-33 of a scan resident on the DEVICE, so the front end runs once per
 * scan instead of once per pair.  Slots are indexed 0 .. n_slots-1; results never depend on whether a scan came from the cache
 * (tests/test_gpu_parity.py::test_scan_cache_*). */
typedef struct qb200_slot_pair { int32_t src_slot, tgt_slot; } qb200_slot_pair;
int qb200_cache_reserve(qb200_handle* h, int32_t n_slots);   /* (re)allocates; 0 frees.  ~3.1 MB per slot at max_voxel_points = 16384 */
/* voxelize + normals + FPFH of n_scans raw scans (scans4[i]: n_points[i] x {x,y,z,w}) into slots slot_ids[i] */
int qb200_cache_scans(qb200_handle* h, const float* const* scans4, const int32_t* n_points, const int32_t* slot_ids, int32_t n_scans,
                      const qb200_params* p, qb200_mem_kind kind);
/* match + graph + clique + pose for pairs of cached scans (= qb200_register_batch without its front end); p's front-end parameters
 * must be the ones the slots were cached with */
int qb200_register_cached(qb200_handle* h, const qb200_slot_pair* pairs, int32_t n_pairs, const qb200_params* p, qb200_result* results);
int qb200_cache_copy(qb200_handle* h, int32_t from_slot, int32_t to_slot);   /* swapTgt2Src */
/* read a cached scan back: voxel points (n x 4), normals (n x {nx,ny,nz,curvature}), descriptors (n x 33); any may be NULL */
int qb200_cache_read(qb200_handle* h, int32_t slot, float* vox4, float* normals4, float* desc33, int32_t cap, int32_t* n);

/* --- multi-GPU: batches of independent pairs shard across the GPUs of one box; the only communication is ONE all-gather (NCCL over
 * NVLink) of the fixed-size result records per batch -- north_star / SURVEY.md 8(e).  The reference has no counterpart (it is a
 * single-process CPU program: examples/run_global_registration.cpp processes one pair); these entry points are what a loop-closure
 * sweep over the reference's `quatro.computeTransformation` would call instead.  libnccl.so.2 is opened at the first call
 * (QB200_ERR_UNSUPPORTED when it is absent).
 *
 * (A) one process, several devices: handles[i] was created on device i' (any distinct devices); pair g runs on handles[g mod n_dev]
 *     (QB200_MEM_DEVICE pointers of pair g must live on that device); results come back in the order of `pairs`. */
#define QB200_UNIQUE_ID_BYTES 128
int qb200_comm_init_all(qb200_handle** handles, int32_t n_dev);
int qb200_register_batch_sharded(qb200_handle** handles, int32_t n_dev, const qb200_pair* pairs, int32_t n_pairs,
                                 const qb200_params* p, qb200_mem_kind kind, qb200_result* results);
/* (B) one process per device (torchrun / mpirun): rank 0 calls qb200_comm_unique_id and hands the 128 bytes to every rank (any
 *     out-of-band channel), every rank calls qb200_comm_init_rank.  qb200_register_batch_rank registers this rank's n_local pairs
 *     (the same n_local on every rank) and gathers all records: all_results[i * world + r] = record of rank r's i-th pair
 *     (round-robin sharding of a global list).  defer != 0: the call returns as soon as the gather is enqueued on the handle's
 *     communication stream -- all_results is complete after qb200_comm_wait (or the next qb200_register_batch_rank), so the
 *     gather overlaps the next batch and no rank waits for the slowest one inside a step.  defer == 2 (a stream of batches): the
 *     local batch is only queued (qb200_register_batch_enqueue), the call then completes the PREVIOUS batch's records and starts
 *     their gather; local_pairs' scans and all_results of a batch must stay valid until its records were delivered (two calls
 *     later, or qb200_comm_wait, which ends the stream). */
int qb200_comm_unique_id(void* id128);
int qb200_comm_init_rank(qb200_handle* h, int32_t world, int32_t rank, const void* id128);
int qb200_register_batch_rank(qb200_handle* h, const qb200_pair* local_pairs, int32_t n_local, const qb200_params* p,
                              qb200_mem_kind kind, qb200_result* all_results, int32_t defer);
int qb200_comm_wait(qb200_handle* h);
/* Bind the calling host thread to the cores of the NUMA node of the handle's GPU (kernel launches and pinned copies from the far
 * socket of a 2-socket box are slower).  Returns the number of cores bound (0: topology unknown, nothing changed). */
int qb200_bind_numa(qb200_handle* h);

/* Introspection of the most recent single-pair solve on this handle (getMaxCliques,
 * getFinalInliersIndices, getCorrespondences; quatro.hpp:949-972, fpfh_manager.hpp:234-236). */
int qb200_get_last_clique(qb200_handle* h, int32_t* idx, int32_t cap, int32_t* n);
int qb200_get_last_final_inliers(qb200_handle* h, int32_t* idx, int32_t cap, int32_t* n);
int qb200_get_last_correspondences(qb200_handle* h, int32_t* corr, float* src_matched4,
                                   float* tgt_matched4, int32_t cap, int32_t* n);

/* Normals (n x {nx,ny,nz,curvature}) and FPFH-33 descriptors (n x 33) the most recent qb200_match_and_pack (which = 0 source,
 * 1 target) or qb200_compute_fpfh (which = 0) left on the device: FPFHManager::getObjDescriptor / getSceneDescriptor /
 * getTgtNormals (include/fpfh_manager.hpp:161-177).  Either pointer may be NULL. */
int qb200_get_last_features(qb200_handle* h, int32_t which, float* normals4, float* desc33, int32_t cap, int32_t* n);

/* Per-stage device time of the last qb200_register_batch call in milliseconds (CUDA events):
 * [0]=h2d [1]=voxel [2]=fpfh [3]=match [4]=graph [5]=clique [6]=pose [7]=d2h; n<=8. */
int qb200_get_stage_ms(qb200_handle* h, float* ms, int32_t n);

/* Device time (CUDA events on the handle's stream) and launch count of the two roofline kernels
 * during the last qb200_register_batch call: [0] = the tensor-core nearest-neighbour passes (tc_match_kernel x3),
 * [1] = tim_graph_kernel (TIM consistency graph); n <= 2. */
int qb200_get_kernel_ms(qb200_handle* h, float* ms, int32_t* launches, int32_t n);

/* Diagnostics: the tensor-core (tcgen05, 3xTF32) approximate squared distances that pre-filter the 33-D
 * nearest-neighbour search, for up to 128 x 128 descriptors (out[128*128], row = a).  The matcher's results
 * never depend on these values (exact fp32 re-rank); tests use this to measure the filter's error margin. */
int qb200_debug_tc_distances(qb200_handle* h, const float* a33, int32_t na, const float* b33, int32_t nb, float* out);


/* Diagnostics: cumulative counters of the tensor-core matcher since creation / the last reset:
 * out4[0] = descriptor pairs that went through the exact fp32 chain, [1] = 128 x 128 tiles drained,
 * [2] = warm-up passes, [3] = stripes handed to the exact kernel.  Synchronises the handle's stream. */
int qb200_debug_match_stats(qb200_handle* h, uint64_t* out4, int32_t reset);
/* QB200_TC_PROF=1 only: per-role clock64 accounting of tc_nn_kernel (24 counters, see tools/tc_profile.py) */
int qb200_debug_tc_profile(qb200_handle* h, uint64_t* out24, int32_t reset);

/* Diagnostics: with QB200_TC_VERIFY=1 in the environment every batch is matched by the tensor-core path AND by the exact CUDA-core
 * kernel; out2[0] = nearest-neighbour table entries compared so far, out2[1] = entries whose packed (distance, index) differ
 * (0 unless the filter's error bound is violated).  Synchronises the handle's stream. */
int qb200_debug_match_verify(qb200_handle* h, uint64_t* out2, int32_t reset);

#ifdef __cplusplus
}
#endif
#endif /* QUATRO_B200_H_ */
