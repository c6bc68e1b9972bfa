// handle.cuh -- the opaque qb200_handle: one device, one stream, device workspaces for one wave of
// max_batch_slots pairs.  Replaces the reference's process-global state (function-local statics at
// include/quatro.hpp:53,64,469-470,660 and include/fpfh_manager.hpp:110) with per-handle state.
#pragma once
#include <string.h>

#include "common.cuh"

struct qb200_handle {
  qb200_config cfg;
  int S, R, V, Lc, W;        // slots, raw cap / cloud, voxel cap / cloud, corr cap / pair, words per adjacency row
  int NS;                    // match stripes per pair = V / kMatchTile
  int device;
  cudaStream_t own_stream, stream;
  char err[512];
  int64_t launches;

  // ---- wave description (host-known inputs) ----
  const float4** d_cloud_ptr; // [2S]
  int* d_cloud_n;             // [2S]
  int* d_raw_off;             // [2S+1] offsets into the concatenated sort arrays
  const float4** h_cloud_ptr; int* h_cloud_n; int* h_raw_off;  // pinned mirrors
  float4* raw_stage;          // [2S*R] staging for host inputs
  // Multi-wave batches rotate over this handle and up to 7 more lanes (own stream and buffers, created on first use):
  // the H2D copies and the latency-bound solver tail of one wave overlap the dense kernels of the others.
  qb200_handle* lane[7];
  int max_lanes;              // 1..8 (QB200_LANES, default 4)
  unsigned func_attr_set;     // which kernels already got their dynamic shared-memory opt-in on this handle's device
  int pend_w0, pend_np;       // wave in flight on this lane (pend_np == 0: none)
  qb200_result* pend_dst;     // ... and the caller's record array of its batch
  int lanes_active, lane_cursor;  // public handle: lanes of the rotation in use (0 = nothing in flight), next lane = busy longest
  cudaEvent_t ev_fork;
  cudaStream_t copy_stream;   // host scans of a multi-wave batch cross PCIe on ONE stream, wave after wave (api.cu: wave_submit)
  cudaEvent_t ev_copied;      // this lane's scans have arrived (recorded on the copy stream)

  // ---- sort workspace (voxel sort, then lattice sort) ----
  uint64_t *key_a, *key_b;    // [2S*R]
  uint32_t *val_a, *val_b;    // [2S*R]
  void* cub_temp; size_t cub_bytes;
  float* aos_scratch;         // [2*V*33] AoS descriptors of the stage entry points (qb200_compute_fpfh / qb200_match)

  // ---- front end ----
  int* vox_start;             // [2S*(V+1)] position (in the sorted raw array) of each voxel's first point
  float4* vox_pts;            // [2S*V] centroids, ascending (k,j,i)
  uint64_t* cell_key;         // [2S*V] occupied lattice cells, ascending
  int* cell_start;            // [2S*(V+1)]
  float4* normals;            // [2S*V]
  float* spfh;                // [2S*V*36] rows padded to 36 floats
  unsigned short* nbr_list;   // [2S][kNbrGlobalCap][V] neighbour indices found by K4 (lattice order), reused by K5
  int* nbr_cnt;               // [2S*V] neighbour count (self included); > kNbrGlobalCap: K5 walks the lattice itself
  float* desc_t;              // [2S*40*V] FPFH, dimension-major per cloud (row d = bin d over all points; rows 33..39 zero)
  float* desc_tiles;          // [2S*(V/128)*3*5120] per 128-point block: centred TF32 hi | lo | exact fp32 images in the UMMA
                              // shared-memory operand layout (one bulk copy per tile)
  float* desc_norm;           // [2S*V] squared norms (fp32 fma chain)
  int* tc_fallback;           // [S] 1 = too many exact ties for the filter to pay off: pair re-done by the exact fp32 kernel
  unsigned long long* tc_stats; // [32] diagnostics, cumulative: [0..3] exact evaluations, tiles drained, warm-up passes, aborted stripes; [4..5] QB200_TC_VERIFY; [8..31] QB200_TC_PROF
  int force_exact_match;      // 0 (default): tcgen05 filter + exact evaluation; 1 (QB200_MATCH_EXACT=1): exact CUDA-core K6 only
  // ---- matching ----
  unsigned long long* rowbest;// [S*V] packed (dist bits << 32 | tgt idx) per source point
  unsigned long long* colpart;// [S*NS*V] per-stripe partial column minima
  unsigned long long* colbest;// [S*V]
  int *mut_i, *mut_j;         // [S*V] mutual NN list (larger-cloud idx, smaller-cloud idx)
  unsigned char* mark;        // [S*V] tuple-test survivors
  int* partner;               // [S*V] tgt partner per source index (-1)
  float* mean;                // [2S*4]
  int *corr_src, *corr_tgt;   // [S*Lc]
  float4 *ma, *mb;            // [S*Lc] matched points (src, tgt)
  // ---- graph / clique ----
  uint32_t *adj, *adjp;       // [S*Lc*W] adjacency bits, and the same in (core, id)-rank space
  int* deg;                   // [S*Lc]
  int *kcore, *korder, *rank_of, *by_rank, *kbin;  // [S*(Lc+2)]
  int* clique;                // [S*Lc] ascending ids
  uint32_t *ex_stack, *ex_pool;  // PMC_EXACT scratch, allocated on first use: [min(S,64)*1024*W] candidate sets per level, [min(S,64)*2^17] list entries
  int* ex_lvl;                // [2*min(S,64)*1024] list segment (begin | remaining) per level
  unsigned short* ex_cur;     // [min(S,64)*1024] clique under construction (ranks)
  int* final_inl;             // [S*Lc]
  unsigned char *rot_mask, *trans_mask;  // [S*Lc]
  // ---- pre-processing (preprocess.cu), allocated on first use ----
  int* pw_ints;               // patch id / rank per point, per-patch counters and offsets
  float4* pw_out;             // [2*R] ground | non-ground
  void* ip_buf; int ip_npix;  // range-image scratch (per pixel: winner, parent, size, range, row set, two outputs)
  // ---- results ----
  qb200_result* d_results;    // [S]
  qb200_result* h_results;    // pinned [S]
  qb::WaveCounters ctr;
  int* ctr_block; size_t ctr_ints;

  // ---- scan cache (qb200_cache_*): front-end results of whole scans, resident on the device ----
  int c_slots;
  float4 *c_vox, *c_nrm;      // [slots*V]
  float* c_desc;              // [slots*40*V] dimension-major like desc_t
  int *c_n, *c_status;        // [slots]
  int *d_slot_of_cloud, *h_slot_of_cloud;   // [2S] cache slot of every cloud of the current wave (device / pinned)
  float* c_sig;               // host [slots*4]: (voxel, normal_r, fpfh_r, cell) a slot was computed with

  // ---- multi-GPU gather of the result records (comm.cu) ----
  void* comm;                 // ncclComm_t
  int comm_world, comm_rank, comm_cap;   // cap: records per rank the staging buffers hold
  cudaStream_t comm_stream;
  cudaEvent_t comm_done;
  qb200_result *d_send, *d_recv, *h_send, *h_recv;   // device staging; pinned host staging (h_recv is rank-major)
  int pend_gather_n;          // > 0: a deferred gather is in flight (records per rank)
  qb200_result* pend_gather_dst;
  int pipe_n, pipe_buf;       // pipelined rank mode: local batch queued, gather not started yet (records per rank, half of h_send)
  qb200_result* pipe_dst;

  // ---- state mirrored from the reference's statics ----
  double rot_noise_bound_latched;  // quatro.hpp:469-470 (0 = not latched yet)
  int last_n_corr, last_n_clique, last_n_final;  // slot 0 of the most recent single-pair call

  cudaEvent_t ev[9];
  float stage_ms[8];
  cudaEvent_t kev[4];        // [0,1] around match_stripe_kernel, [2,3] around tim_graph_kernel (last wave)
  float kernel_ms[2];
  int kernel_calls[2];
  int kev_armed[2];

  void fail(const char* file, int line, const char* msg) {
    snprintf(err, sizeof(err), "%s:%d: %s", file, line, msg);
  }
};

namespace qb {

// Stage launchers (each enqueues kernels on h->stream for clouds/pairs [0, n) of the current wave).
int launch_voxel(qb200_handle* h, int n_clouds, int total_raw, float leaf, int skip_flagged);
int launch_fpfh(qb200_handle* h, int n_clouds, float normal_radius, float fpfh_radius, float cell);
int launch_match(qb200_handle* h, int n_pairs, const qb200_params& p);
int launch_graph(qb200_handle* h, int n_pairs, double noise_bound, double cbar2);
int launch_clique(qb200_handle* h, int n_pairs, int mode, double kcore_thr, long long node_limit);
int launch_pose(qb200_handle* h, int n_pairs, const qb200_params& p);
int launch_fill_counters(qb200_handle* h, int n_pairs, int have_frontend);
int launch_finalize_status(qb200_handle* h, int n_pairs);
int launch_iota_clique(qb200_handle* h, int n_pairs);
int launch_segment_cloud(qb200_handle* h, const float4* pts, int n, const qb200_segment_params& sp, int* n_valid, int* n_outlier,
                         const float4** valid_dev, const float4** outlier_dev);
int launch_patchwork(qb200_handle* h, const float4* pts, int n, const qb200_patchwork_params& pp, int* n_ground, int* n_nonground, int* status);
int launch_match_nn(qb200_handle* h, int n_pairs);
int launch_match_exact(qb200_handle* h, int n_pairs, const int* only);
int launch_tc_debug_tile(qb200_handle* h, float* d_out);
int launch_desc_to_aos(qb200_handle* h, int cloud, int n, float* d_out33);
int launch_desc_from_aos(qb200_handle* h, int cloud, int n, const float* d_in33);
int desc_to_aos_rows(qb200_handle* h, const float* desc_rows, int n, float* d_out33);
size_t sort_temp_bytes(int max_items);
void comm_release(qb200_handle* h);
int collect_batch(qb200_handle* h, const qb200_result* dst);  // api.cu: wait for every wave in flight that writes into dst[...]
// Raise a kernel's dynamic shared-memory opt-in to at least `bytes` on the handle's device.  The attribute is a property of the
// (function, device), not of a handle: handles of different capacities share it, so it is only ever raised (process-wide maximum).
int ensure_dyn_smem(qb200_handle* h, const void* kernel, size_t bytes);
int sort_pairs(qb200_handle* h, int n_items, int end_bit);
int sort_keys(qb200_handle* h, int n_items, int begin_bit, int end_bit);
int launch_voxel_sort(qb200_handle* h, int n_clouds, float inv_leaf, int skip_flagged, int idx_bits);
int launch_cloud_sort(qb200_handle* h, int n_clouds, const int* n_items, int f1, int f2);

}  // namespace qb
