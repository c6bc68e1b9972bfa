// frontend.cu -- K1..K5: voxel down-sampling, neighbour lattice, normals, SPFH, FPFH  (sm_100a)
//
// Replaces: voxelize<T>() (include/quatro.hpp:49-57 -> [EXT] pcl::VoxelGrid) and
// FPFHEstimation::computeFPFHFeatures (src/teaser_utils/fpfh.cc:44-75 -> [EXT] pcl::NormalEstimation,
// pcl::FPFHEstimationOMP, pcl::search::KdTree).
//
// Design: every cloud of a batch wave is processed by the same launches (grid.y = cloud); sizes
// stay on the device.  Voxelisation is a stable per-cloud radix sort of the (voxel | point index) items of the KEPT
// points (voxsort.cu; the device-wide library sort below is the fallback for handles too small for its scratch and the
// QB200_VOXEL_SORT=cub A/B switch) followed by a segmented, in-order centroid sum (bit-identical to the sequential CPU sum).  The kd-tree is
// replaced by a sorted-cell lattice: a point's neighbours are found by (2m+1)^2 binary searches for
// x-runs of cells, visited in ascending (cell, index) order -- the accumulation order the CPU oracle
// uses, so the single-pass float covariance matches bit for bit.
#include <cub/device/device_radix_sort.cuh>

#include <cstdlib>
#include <cstring>

#include "fpfh_math.cuh"
#include "handle.cuh"

namespace qb {

// ------------------------------------------------------------------------------------------------
// sort plumbing
// ------------------------------------------------------------------------------------------------
size_t sort_temp_bytes(int max_items) {
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint64_t*)nullptr, (uint64_t*)nullptr, (const uint32_t*)nullptr,
                                  (uint32_t*)nullptr, max_items, 0, 64, (cudaStream_t)0);
  return bytes;
}

int sort_pairs(qb200_handle* h, int n_items, int end_bit) {
  if (n_items <= 0) return QB200_OK;
  size_t bytes = h->cub_bytes;
  QB_CUDA_TRY(h, cub::DeviceRadixSort::SortPairs(h->cub_temp, bytes, h->key_a, h->key_b, h->val_a, h->val_b, n_items, 0, end_bit,
                                                 h->stream));
  h->launches += 1 + (end_bit + 7) / 8;  // onesweep: histogram + one pass per 8 key bits
  return QB200_OK;
}

// keys only (the payload rides in the low key bits below begin_bit: 8 instead of 12 bytes per item and pass)
int sort_keys(qb200_handle* h, int n_items, int begin_bit, int end_bit) {
  if (n_items <= 0) return QB200_OK;
  size_t bytes = h->cub_bytes;
  QB_CUDA_TRY(h, cub::DeviceRadixSort::SortKeys(h->cub_temp, bytes, h->key_a, h->key_b, n_items, begin_bit, end_bit, h->stream));
  h->launches += 1 + (end_bit - begin_bit + 7) / 8;
  return QB200_OK;
}

static int clog2(int n) {
  int b = 0;
  while ((1 << b) < n) ++b;
  return b;
}

// ------------------------------------------------------------------------------------------------
// K1a: bounding box of the kept raw points (one pass, 128-bit loads), then raw point -> (cloud | voxel) key.
// The voxel key is PCL's own linear index  (i - min_i) + (j - min_j) dx + (k - min_k) dx dy  ([EXT] pcl::VoxelGrid,
// called from include/quatro.hpp:49-57), which the library requires to fit an int: at most 31 key bits, and for a given cloud
// only the bits of dx dy dz (vox_digits()); the library-sort fallback adds the cloud id above them (5 passes over 38 bits).
// ------------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(256) voxel_bbox_kernel(const float4* const* __restrict__ cloud_ptr, const int* __restrict__ cloud_n,
                                                         float inv_leaf, int skip_flagged, int* __restrict__ bbox,
                                                         int* __restrict__ n_valid, int* __restrict__ cloud_status, int* __restrict__ chunk_cnt) {
  const int cloud = blockIdx.y;
  const int n = cloud_n[cloud];
  const float4* __restrict__ pts = cloud_ptr[cloud];
  int mn0 = INT_MAX, mn1 = INT_MAX, mn2 = INT_MAX, mx0 = INT_MIN, mx1 = INT_MIN, mx2 = INT_MIN, cnt = 0, bad = 0;
  // eight independent 16-byte loads in flight per thread and ~30 points per thread: this pass streams the raw scans (230 MB per
  // 64-pair wave) from HBM, and the reduction tail (shuffles, atomics) is paid once per 30 points instead of once per 7.
  // A CTA walks kVsChunks / gridDim.x CONTIGUOUS chunks of the scan and also reports how many points of each chunk are kept:
  // the pack pass (voxsort.cu) writes the kept points compacted, in scan order, from these counts.
  constexpr int kInFlight = 8;
  __shared__ int s_cc[kVsChunks];
  if (threadIdx.x < kVsChunks) s_cc[threadIdx.x] = 0;
  __syncthreads();
  const int cs = vox_chunk_size(n);
  const int per_cta = kVsChunks / gridDim.x;
  for (int cl = 0; cl < per_cta; ++cl) {
    const int ch = blockIdx.x * per_cta + cl;
    const int c0 = ch * cs, c1 = min(n, c0 + cs);
    int ccnt = 0;
    for (int i0 = c0 + threadIdx.x; i0 < c1; i0 += kInFlight * 256) {
      float4 pp[kInFlight];
#pragma unroll
      for (int j = 0; j < kInFlight; ++j) pp[j] = i0 + j * 256 < c1 ? __ldg(pts + i0 + j * 256) : make_float4(NAN, NAN, NAN, 0.f);  // NaN = not kept
#pragma unroll
      for (int j = 0; j < kInFlight; ++j) {
        const float4 p = pp[j];
        if (i0 + j * 256 >= c1 || !raw_point_kept(p, skip_flagged)) continue;
        const int ci = (int)floorf(p.x * inv_leaf), cj = (int)floorf(p.y * inv_leaf), ck = (int)floorf(p.z * inv_leaf);
        if (cell_ok(ci, cj, ck)) {
          const int ox = float_ordered(p.x), oy = float_ordered(p.y), oz = float_ordered(p.z);
          mn0 = min(mn0, ox); mn1 = min(mn1, oy); mn2 = min(mn2, oz);
          mx0 = max(mx0, ox); mx1 = max(mx1, oy); mx2 = max(mx2, oz);
          ++ccnt;
        } else {
          bad = 1;  // outside the representable lattice: PCL's index would overflow as well
        }
      }
    }
    cnt += ccnt;
    ccnt = __reduce_add_sync(0xffffffffu, ccnt);
    if (lane_id() == 0 && ccnt) atomicAdd(&s_cc[ch], ccnt);
  }
  __syncthreads();
  if ((int)threadIdx.x < per_cta) chunk_cnt[cloud * kVsChunks + blockIdx.x * per_cta + threadIdx.x] = s_cc[blockIdx.x * per_cta + threadIdx.x];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn0 = min(mn0, __shfl_xor_sync(0xffffffffu, mn0, o)); mn1 = min(mn1, __shfl_xor_sync(0xffffffffu, mn1, o));
    mn2 = min(mn2, __shfl_xor_sync(0xffffffffu, mn2, o)); mx0 = max(mx0, __shfl_xor_sync(0xffffffffu, mx0, o));
    mx1 = max(mx1, __shfl_xor_sync(0xffffffffu, mx1, o)); mx2 = max(mx2, __shfl_xor_sync(0xffffffffu, mx2, o));
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o); bad |= __shfl_xor_sync(0xffffffffu, bad, o);
  }
  // one set of global atomics per CTA
  __shared__ int s_red[8];
  if (threadIdx.x == 0) { s_red[0] = s_red[1] = s_red[2] = INT_MAX; s_red[3] = s_red[4] = s_red[5] = INT_MIN; s_red[6] = 0; s_red[7] = 0; }
  __syncthreads();
  if (lane_id() == 0) {
    if (cnt) {
      atomicMin(&s_red[0], mn0); atomicMin(&s_red[1], mn1); atomicMin(&s_red[2], mn2);
      atomicMax(&s_red[3], mx0); atomicMax(&s_red[4], mx1); atomicMax(&s_red[5], mx2);
      atomicAdd(&s_red[6], cnt);
    }
    if (bad) s_red[7] = 1;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (s_red[6]) {
      int* b = bbox + cloud * 6;
      atomicMin(b + 0, s_red[0]); atomicMin(b + 1, s_red[1]); atomicMin(b + 2, s_red[2]);
      atomicMax(b + 3, s_red[3]); atomicMax(b + 4, s_red[4]); atomicMax(b + 5, s_red[5]);
      atomicAdd(n_valid + cloud, s_red[6]);
    }
    if (s_red[7]) cloud_status[cloud] = QB200_ERR_VOXEL_OVERFLOW;
  }
}

__global__ void __launch_bounds__(256) voxel_keys_kernel(const float4* const* __restrict__ cloud_ptr, const int* __restrict__ cloud_n,
                                                         const int* __restrict__ raw_off, float inv_leaf, int skip_flagged,
                                                         const int* __restrict__ bbox, const int* __restrict__ n_valid, int idx_bits,
                                                         uint64_t* __restrict__ keys) {
  const int cloud = blockIdx.y;
  const int n = cloud_n[cloud], off = raw_off[cloud];
  const float4* __restrict__ pts = cloud_ptr[cloud];
  // min_b / div_b of pcl::VoxelGrid::applyFilter
  long long m0 = 0, m1 = 0, m2 = 0, d0 = 1, d1 = 1, d2 = 1;
  if (n_valid[cloud] > 0) {
    const int* b = bbox + cloud * 6;
    m0 = (long long)floorf(ordered_float(b[0]) * inv_leaf); m1 = (long long)floorf(ordered_float(b[1]) * inv_leaf);
    m2 = (long long)floorf(ordered_float(b[2]) * inv_leaf);
    d0 = (long long)floorf(ordered_float(b[3]) * inv_leaf) - m0 + 1; d1 = (long long)floorf(ordered_float(b[4]) * inv_leaf) - m1 + 1;
    d2 = (long long)floorf(ordered_float(b[5]) * inv_leaf) - m2 + 1;
  }
  (void)d2;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float4 p = __ldg(pts + i);
    uint64_t cell = kVoxInvalid;
    if (raw_point_kept(p, skip_flagged)) {
      const int ci = (int)floorf(p.x * inv_leaf), cj = (int)floorf(p.y * inv_leaf), ck = (int)floorf(p.z * inv_leaf);
      if (cell_ok(ci, cj, ck)) {
        const long long lin = ((long long)ci - m0) + ((long long)cj - m1) * d0 + ((long long)ck - m2) * d0 * d1;
        if (lin >= 0 && lin < (long long)kVoxInvalid) cell = (uint64_t)lin;  // otherwise the cloud is refused (overflow) anyway
      }
    }
    keys[off + i] = ((((uint64_t)cloud << kVoxShift) | cell) << idx_bits) | (uint64_t)i;  // stable sort on the bits above idx_bits
  }
}

// ------------------------------------------------------------------------------------------------
// K1b / K2b: run heads of the sorted keys of one cloud -> start position of every voxel / cell.
// One CTA per cloud walks its segment with a carried block scan (sizes never leave the device).
//   mode 0 (voxels): segment = [raw_off, raw_off + n_raw), writes starts[], n_out = #voxels
//   mode 1 (cells):  segment = [cloud*V, cloud*V + V),     writes starts[] and cell keys
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) run_heads_kernel(int mode, const uint64_t* keys, const uint64_t* keys_alt, const int* __restrict__ seg_off,
                                                         const int* __restrict__ seg_n, int V, int key_shift, float inv_leaf, const int* __restrict__ bbox,
                                                         const int* __restrict__ n_valid_in, int* __restrict__ starts,
                                                         uint64_t* __restrict__ cell_keys, int* __restrict__ n_out, int* __restrict__ n_valid_out,
                                                         int* __restrict__ cloud_status) {
  __shared__ int sm[33];
  __shared__ int s_overflow;
  const int cloud = blockIdx.x;
  const int off = mode == 0 ? seg_off[cloud] : cloud * V;
  int n = mode == 0 ? seg_n[cloud] : V;
  if (mode == 0 && keys_alt) {  // own voxel sort: the kept points only, in A or B by the parity of the cloud's digit count
    n = n_valid_in[cloud];
    if (vox_digits(bbox + cloud * 6, n, inv_leaf) & 1) keys = keys_alt;
  }
  if (threadIdx.x == 0) {
    int ov = 0;
    if (mode == 0 && n_valid_in[cloud] > 0) {
      // [EXT] pcl::VoxelGrid: dx*dy*dz > INT_MAX -> "leaf size too small", input returned unfiltered
      const int* b = bbox + cloud * 6;
      const long long dx = (long long)((ordered_float(b[3]) - ordered_float(b[0])) * inv_leaf) + 1;
      const long long dy = (long long)((ordered_float(b[4]) - ordered_float(b[1])) * inv_leaf) + 1;
      const long long dz = (long long)((ordered_float(b[5]) - ordered_float(b[2])) * inv_leaf) + 1;
      if (dx * dy * dz > (long long)INT_MAX) ov = 1;
    }
    if (mode == 0 && cloud_status[cloud] == QB200_ERR_VOXEL_OVERFLOW) ov = 1;
    s_overflow = ov;
  }
  __syncthreads();
  if (s_overflow) {
    if (threadIdx.x == 0) {
      n_out[cloud] = 0;
      cloud_status[cloud] = QB200_ERR_VOXEL_OVERFLOW;
      starts[(size_t)cloud * (V + 1)] = 0;
    }
    return;
  }
  int carry = 0, valid_total = 0;
  // four consecutive keys per thread and round: a quarter of the block scans (three barriers each) of a key-per-thread loop
  constexpr int kPer = 4;
  for (int base = 0; base < n; base += kPer * blockDim.x) {
    const int p0 = base + kPer * threadIdx.x;
    uint64_t k[kPer];
    bool valid[kPer];
    int head[kPer], nh = 0, nv = 0;
    uint64_t kprev = (p0 > 0 && p0 < n) ? keys[off + p0 - 1] >> key_shift : 0;
#pragma unroll
    for (int j = 0; j < kPer; ++j) {
      const int p = p0 + j;
      k[j] = 0; valid[j] = false; head[j] = 0;
      if (p < n) {
        k[j] = keys[off + p] >> key_shift;  // mode 0: the low bits carry the point index
        valid[j] = mode == 0 ? (k[j] & kVoxMask) != kVoxInvalid : (k[j] & kCellMask) != kCellInvalid;
        head[j] = (valid[j] && (p == 0 || k[j] != kprev)) ? 1 : 0;
        kprev = k[j];
      }
      nh += head[j]; nv += valid[j] ? 1 : 0;
    }
    // one scan for both counts: heads in the low half, valid points in the high half (<= 4096 each per round)
    int both;
    const int exb = block_excl_scan(nh | (nv << 16), sm, &both);
    int rank = carry + (exb & 0xFFFF);
    const int tot = both & 0xFFFF, vtot = both >> 16;
#pragma unroll
    for (int j = 0; j < kPer; ++j) {
      if (head[j]) {
        if (rank <= V) {
          starts[(size_t)cloud * (V + 1) + rank] = p0 + j;
          if (mode == 1 && rank < V) cell_keys[(size_t)cloud * V + rank] = k[j] & kCellMask;
        }
        ++rank;
      }
    }
    carry += tot;
    valid_total += vtot;
  }
  if (threadIdx.x == 0) {
    int nv = carry;
    if (nv > V) {
      nv = V;
      cloud_status[cloud] = QB200_CAPACITY_EXCEEDED;
    } else {
      starts[(size_t)cloud * (V + 1) + nv] = valid_total;  // end sentinel: valid points sort to the front
    }
    n_out[cloud] = nv;
    if (n_valid_out) n_valid_out[cloud] = valid_total;
  }
}

// K1c: centroid of each voxel, summed in original point order (stable sort) -> identical to the
// sequential CPU sum.  One thread per voxel; points are gathered through the sorted index.
__global__ void __launch_bounds__(128) voxel_centroid_kernel(const float4* const* __restrict__ cloud_ptr, const int* __restrict__ raw_off,
                                                             const uint64_t* sorted_keys, const uint64_t* keys_alt, const int* __restrict__ bbox,
                                                             const int* __restrict__ n_valid, float inv_leaf, uint64_t idx_mask,
                                                             const int* __restrict__ starts, const int* __restrict__ n_vox, int V,
                                                             float4* __restrict__ vox_pts) {
  const int cloud = blockIdx.y;
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_vox[cloud]) return;
  if (keys_alt && (vox_digits(bbox + cloud * 6, n_valid[cloud], inv_leaf) & 1)) sorted_keys = keys_alt;
  const float4* __restrict__ pts = cloud_ptr[cloud];
  const int off = raw_off[cloud];
  const int a = starts[(size_t)cloud * (V + 1) + r], b = starts[(size_t)cloud * (V + 1) + r + 1];
  float sx = 0.f, sy = 0.f, sz = 0.f;
  int t = a;
  for (; t + 4 <= b; t += 4) {  // four gathers in flight, summed in order
    const uint64_t k0 = sorted_keys[off + t], k1 = sorted_keys[off + t + 1], k2 = sorted_keys[off + t + 2], k3 = sorted_keys[off + t + 3];
    const float4 p0 = __ldg(pts + (k0 & idx_mask)), p1 = __ldg(pts + (k1 & idx_mask)), p2 = __ldg(pts + (k2 & idx_mask)),
                 p3 = __ldg(pts + (k3 & idx_mask));
    sx += p0.x; sy += p0.y; sz += p0.z;
    sx += p1.x; sy += p1.y; sz += p1.z;
    sx += p2.x; sy += p2.y; sz += p2.z;
    sx += p3.x; sy += p3.y; sz += p3.z;
  }
  for (; t < b; ++t) {
    const float4 p = __ldg(pts + (sorted_keys[off + t] & idx_mask));
    sx += p.x; sy += p.y; sz += p.z;
  }
  const float cnt = (float)(b - a);
  vox_pts[(size_t)cloud * V + r] = make_float4(sx / cnt, sy / cnt, sz / cnt, 1.0f);
}

// K2a: lattice keys of the (voxelised) clouds
__global__ void __launch_bounds__(256) lattice_keys_kernel(const float4* __restrict__ pts, const int* __restrict__ n_pts, int V, float inv_cell,
                                                           uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  const int cloud = blockIdx.y;
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= V) return;
  uint64_t cell = kCellInvalid;
  if (r < n_pts[cloud]) {
    const float4 p = pts[(size_t)cloud * V + r];
    if (isfinite(p.x) && isfinite(p.y) && isfinite(p.z)) {
      const int ci = (int)floorf(p.x * inv_cell), cj = (int)floorf(p.y * inv_cell), ck = (int)floorf(p.z * inv_cell);
      if (cell_ok(ci, cj, ck)) cell = cell_key(ci, cj, ck);
    }
  }
  keys[(size_t)cloud * V + r] = ((uint64_t)cloud << kCloudShift) | cell;
  vals[(size_t)cloud * V + r] = (uint32_t)r;
}

// ------------------------------------------------------------------------------------------------
// Neighbour walk (device): ascending (cell, index) order; set = {d2 < r2}, self included.
// ------------------------------------------------------------------------------------------------
struct LatticeView {
  const float4* pts;       // cloud's points
  const uint64_t* ckeys;   // occupied cells, ascending
  const int* cstart;       // n_cells + 1
  const uint32_t* order;   // point indices sorted by (cell, index)
  int n_cells;
  float inv;
};

// candidates of the occupied cells [c0, c1): loads are issued four points at a time (index -> point are dependent loads;
// independent candidates overlap their latency), tests and callbacks stay in (cell, index) order
template <class F>
__device__ __forceinline__ void walk_cells(const LatticeView& L, const float4 pq, float r2, int c0, int c1, F&& f) {
  if (c0 >= c1) return;
  const int t1 = L.cstart[c1];
  for (int t = L.cstart[c0]; t < t1; t += 4) {
    int p[4];
    float4 pp[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) p[e] = t + e < t1 ? (int)L.order[t + e] : -1;
#pragma unroll
    for (int e = 0; e < 4; ++e) pp[e] = p[e] >= 0 ? L.pts[p[e]] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (p[e] < 0) continue;
      const float dx = pq.x - pp[e].x, dy = pq.y - pp[e].y, dz = pq.z - pp[e].z;
      const float d2 = (dx * dx + dy * dy) + dz * dz;
      if (d2 < r2) f(p[e], d2, pp[e]);
    }
  }
}

template <class F>
__device__ __forceinline__ void for_each_neighbor(const LatticeView& L, const float4 pq, int m, float r2, F&& f) {
  if (!(isfinite(pq.x) && isfinite(pq.y) && isfinite(pq.z))) return;
  const int ci = (int)floorf(pq.x * L.inv), cj = (int)floorf(pq.y * L.inv), ck = (int)floorf(pq.z * L.inv);
  if (!cell_ok(ci, cj, ck)) return;
  const int ilo = max(ci - m, -kOffIJ), ihi = min(ci + m, kOffIJ - 2);
  if (m == 1) {
    // the usual case (cell >= radius): the 9 (k, j) rows are located by 9 binary searches that advance in lockstep, so
    // their loads are independent (one round trip per step instead of nine), and each row is one contiguous cell range
    uint64_t lo[9], hi[9];
    int a[9], b[9], e[9];
#pragma unroll
    for (int r = 0; r < 9; ++r) {
      const int k = ck + r / 3 - 1, j = cj + r % 3 - 1;
      const bool ok = !(k < -kOffK || k >= kOffK - 1 || j < -kOffIJ || j >= kOffIJ - 1);
      lo[r] = ok ? cell_key(ilo, j, k) : 1;
      hi[r] = ok ? cell_key(ihi, j, k) : 0;
      a[r] = 0; b[r] = ok ? L.n_cells : 0;
      e[r] = 0;
    }
    const int steps = 33 - __clz(L.n_cells | 1);
    for (int s = 0; s < steps; ++s) {
#pragma unroll
      for (int r = 0; r < 9; ++r) {
        if (a[r] < b[r]) {
          const int mid = (a[r] + b[r]) >> 1;
          if (L.ckeys[mid] < lo[r]) a[r] = mid + 1; else b[r] = mid;
        }
      }
    }
    // end of each row's range: at most 3 cells (i - 1, i, i + 1)
#pragma unroll
    for (int r = 0; r < 9; ++r) {
      int c = a[r];
#pragma unroll
      for (int q = 0; q < 3; ++q)
        if (c < L.n_cells && hi[r] >= lo[r] && L.ckeys[c] <= hi[r]) ++c;
      e[r] = c;
    }
#pragma unroll
    for (int r = 0; r < 9; ++r) walk_cells(L, pq, r2, a[r], e[r], f);
    return;
  }
  for (int dk = -m; dk <= m; ++dk) {
    const int k = ck + dk;
    if (k < -kOffK || k >= kOffK - 1) continue;
    for (int dj = -m; dj <= m; ++dj) {
      const int j = cj + dj;
      if (j < -kOffIJ || j >= kOffIJ - 1) continue;
      const uint64_t lo = cell_key(ilo, j, k), hi = cell_key(ihi, j, k);
      int a = 0, b = L.n_cells;
      while (a < b) {
        const int mid = (a + b) >> 1;
        if (L.ckeys[mid] < lo) a = mid + 1; else b = mid;
      }
      int c = a;
      while (c < L.n_cells && L.ckeys[c] <= hi) ++c;
      walk_cells(L, pq, r2, a, c, f);
    }
  }
}

__device__ __forceinline__ LatticeView make_view(int cloud, int V, const float4* pts, const uint64_t* cell_key, const int* cell_start,
                                                 const uint32_t* order, const int* n_cells, float inv) {
  LatticeView L;
  L.pts = pts + (size_t)cloud * V;
  L.ckeys = cell_key + (size_t)cloud * V;
  L.cstart = cell_start + (size_t)cloud * (V + 1);
  L.order = order + (size_t)cloud * V;
  L.n_cells = n_cells[cloud];
  L.inv = inv;
  return L;
}

// K4 / K5 walk the same neighbourhoods.  A thread first COLLECTS its neighbour indices (cheap distance tests, divergent)
// into a shared-memory list and then processes the list in a dense loop, so the expensive per-neighbour work (pair
// features, 33-bin gathers) runs with most lanes of the warp active instead of whenever any lane found a neighbour.
// Neighbourhoods larger than the list are handled window by window (ordinals [base, base + cap)), order preserved.
constexpr int kNbrThreads = 128;
constexpr int kNbrCap = 96;

template <class Process>
__device__ __forceinline__ int for_each_neighbor_listed(const LatticeView& L, const float4 pq, int m, float r2,
                                                        unsigned short (*nbr)[kNbrThreads], Process&& process) {
  int k_total = 0;
  for (int base = 0;; base += kNbrCap) {
    int k = 0;
    for_each_neighbor(L, pq, m, r2, [&](int p, float, const float4) {
      if (k >= base && k < base + kNbrCap) nbr[k - base][threadIdx.x] = (unsigned short)p;
      ++k;
    });
    k_total = k;
    const int kl = k_total - base < kNbrCap ? k_total - base : kNbrCap;
    for (int t = 0; t < kl; ++t) process((int)nbr[t][threadIdx.x]);
    if (base + kNbrCap >= k_total) break;
  }
  return k_total;
}

// K2c: the fpfh_radius neighbourhood of every point, found ONCE: indices in lattice (cell, index) order, written
// [t][point] so that the stores of a CTA's points coalesce.  K3 (smaller radius: a subsequence of the same order), K4 and K5
// consume the list; a point with more than kNbrGlobalCap neighbours makes its consumers walk the lattice themselves.
__global__ void __launch_bounds__(kNbrThreads) nbr_list_kernel(const float4* __restrict__ pts, const int* __restrict__ n_pts, int V,
                                                               const uint64_t* __restrict__ cell_key, const int* __restrict__ cell_start,
                                                               const uint32_t* __restrict__ order, const int* __restrict__ n_cells, float inv,
                                                               int m, float r2, unsigned short* __restrict__ nbr_list,
                                                               int* __restrict__ nbr_cnt) {
  const int cloud = blockIdx.y;
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n_pts[cloud]) return;
  const LatticeView L = make_view(cloud, V, pts, cell_key, cell_start, order, n_cells, inv);
  unsigned short* __restrict__ gl = nbr_list + (size_t)cloud * kNbrGlobalCap * V + q;
  int k = 0;
  for_each_neighbor(L, L.pts[q], m, r2, [&](int p, float, const float4) {
    if (k < kNbrGlobalCap) gl[(size_t)k * V] = (unsigned short)p;
    ++k;
  });
  nbr_cnt[(size_t)cloud * V + q] = k;
}

// K3: normals.  One thread per point, sequential float accumulation in lattice order over the points within
// normal_radius: the subsequence of the K2c list that passes the (bit-identical) distance test.
__global__ void __launch_bounds__(128) normals_kernel(const float4* __restrict__ pts, const int* __restrict__ n_pts, int V,
                                                      const uint64_t* __restrict__ cell_key, const int* __restrict__ cell_start,
                                                      const uint32_t* __restrict__ order, const int* __restrict__ n_cells, float inv, int m,
                                                      float r2, const unsigned short* __restrict__ nbr_list, const int* __restrict__ nbr_cnt,
                                                      int list_usable, float4* __restrict__ normals) {
  const int cloud = blockIdx.y;
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n_pts[cloud]) return;
  const LatticeView L = make_view(cloud, V, pts, cell_key, cell_start, order, n_cells, inv);
  const float4 pq = L.pts[q];
  float accu[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  int cnt = 0;
  auto add = [&](const float4 pp) {
    accu[0] += pp.x * pp.x; accu[1] += pp.x * pp.y; accu[2] += pp.x * pp.z;
    accu[3] += pp.y * pp.y; accu[4] += pp.y * pp.z; accu[5] += pp.z * pp.z;
    accu[6] += pp.x; accu[7] += pp.y; accu[8] += pp.z;
    ++cnt;
  };
  const int kq = nbr_cnt[(size_t)cloud * V + q];
  if (list_usable && kq <= kNbrGlobalCap) {
    const unsigned short* __restrict__ gl = nbr_list + (size_t)cloud * kNbrGlobalCap * V + q;
    for (int t = 0; t < kq; ++t) {
      const float4 pp = L.pts[gl[(size_t)t * V]];
      const float dx = pq.x - pp.x, dy = pq.y - pp.y, dz = pq.z - pp.z;
      const float d2 = (dx * dx + dy * dy) + dz * dz;  // same expression as the lattice walk
      if (d2 < r2) add(pp);
    }
  } else {
    for_each_neighbor(L, pq, m, r2, [&](int, float, const float4 pp) { add(pp); });
  }
  float out[4];
  qb_normal_from_accu(accu, cnt, pq.x, pq.y, pq.z, out);
  normals[(size_t)cloud * V + q] = make_float4(out[0], out[1], out[2], out[3]);
}

// K4: SPFH.  Bin COUNTS are order-free; the float histogram value is rebuilt by repeated addition
// of the same increment, which is what the sequential reference loop produces.
// kRare = false: the points whose neighbourhood is in the K2c list (all but a handful): no lattice walk compiled in, no list
// buffer in shared memory.  kRare = true: only the points with more than kNbrGlobalCap neighbours, which walk the lattice.
// SPFH rows are stored as three padded thirds [11 bins, 0][11 bins, 0][11 bins, 0] (36 floats): a third is three 16-byte loads.
constexpr int kSpfhThreads = kNbrThreads;
__device__ __forceinline__ int spfh_slot(int b) { return b + b / 11; }
template <bool kRare>
__global__ void __launch_bounds__(kSpfhThreads) spfh_kernel(const float4* __restrict__ pts, const float4* __restrict__ normals,
                                                            const int* __restrict__ n_pts, int V, const uint64_t* __restrict__ cell_key,
                                                            const int* __restrict__ cell_start, const uint32_t* __restrict__ order,
                                                            const int* __restrict__ n_cells, float inv, int m, float r2,
                                                            float* __restrict__ spfh, const unsigned short* __restrict__ nbr_list,
                                                            const int* __restrict__ nbr_cnt) {
  __shared__ unsigned short cnts[kDescDim][kSpfhThreads];
  __shared__ unsigned short nbr[kRare ? kNbrCap : 1][kNbrThreads];
  const int cloud = blockIdx.y;
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n_pts[cloud]) return;
  int k = nbr_cnt[(size_t)cloud * V + q];
  if ((k > kNbrGlobalCap) != kRare) return;
  const LatticeView L = make_view(cloud, V, pts, cell_key, cell_start, order, n_cells, inv);
  const float4* __restrict__ nrm = normals + (size_t)cloud * V;
  const float4 pq = L.pts[q];
  const float4 nq = nrm[q];
#pragma unroll
  for (int b = 0; b < kDescDim; ++b) cnts[b][threadIdx.x] = 0;
  auto feature = [&](int p) {
    if (p == q) return;
    const float4 pp = L.pts[p];
    const float4 np = nrm[p];
    float f1, f2, f3;
    if (!qb_pair_features(pq.x, pq.y, pq.z, nq.x, nq.y, nq.z, pp.x, pp.y, pp.z, np.x, np.y, np.z, &f1, &f2, &f3)) return;
    int b1, b2, b3;
    qb_feature_bins(f1, f2, f3, &b1, &b2, &b3);
    cnts[b1][threadIdx.x]++;
    cnts[11 + b2][threadIdx.x]++;
    cnts[22 + b3][threadIdx.x]++;
  };
  if (!kRare) {
    const unsigned short* __restrict__ gl = nbr_list + (size_t)cloud * kNbrGlobalCap * V + q;
    for (int t = 0; t < k; ++t) feature((int)gl[(size_t)t * V]);
  } else {
    k = for_each_neighbor_listed(L, pq, m, r2, nbr, feature);
  }
  float* __restrict__ out = spfh + ((size_t)cloud * V + q) * kDescPad;  // three padded thirds: 16-byte gathers in K5
  const float incr = k >= 2 ? 100.0f / (float)(k - 1) : 0.0f;
  for (int b = 0; b < kDescDim; ++b) {
    const int c = cnts[b][threadIdx.x];
    float v = 0.0f;
    for (int t = 0; t < c; ++t) v += incr;
    out[spfh_slot(b)] = v;
  }
  out[11] = 0.0f; out[23] = 0.0f; out[35] = 0.0f;
}

// K5: FPFH = per-third renormalised sum of neighbour SPFHs weighted by 1/d^2, neighbours in lattice
// order.  Output is written dimension-major (desc_t[d][q]) for the matching kernel's tile loads.
//
// fpfh_list_kernel (the points whose neighbourhood is in the K2c list): the three thirds of the signature are independent
// (own 11 accumulators, own fp64 normaliser), so a point is served by three threads in three different warps -- 11 accumulators
// and three 16-byte gathers per neighbour each instead of 33 and nine: 3x the warps at well under half the registers.
// Per-bin accumulation order (lattice order of the neighbours) is unchanged.
__global__ void __launch_bounds__(3 * kNbrThreads, 3) fpfh_list_kernel(const float4* __restrict__ pts, const int* __restrict__ n_pts, int V,
                                                                    const float* __restrict__ spfh, const unsigned short* __restrict__ nbr_list,
                                                                    const int* __restrict__ nbr_cnt, float* __restrict__ desc_t) {
  const int cloud = blockIdx.y;
  const int third = threadIdx.x / kNbrThreads;  // warp-uniform
  const int q = blockIdx.x * kNbrThreads + (threadIdx.x - third * kNbrThreads);
  if (q >= n_pts[cloud]) return;
  const int kq = nbr_cnt[(size_t)cloud * V + q];
  if (kq > kNbrGlobalCap) return;  // fpfh_rare_kernel
  const float4* __restrict__ P = pts + (size_t)cloud * V;
  const float4* __restrict__ sp = reinterpret_cast<const float4*>(spfh + (size_t)cloud * V * kDescPad) + 3 * third;
  const float4 pq = P[q];
  float o[11];
#pragma unroll
  for (int b = 0; b < 11; ++b) o[b] = 0.0f;
  double sum = 0.0;
  const unsigned short* __restrict__ gl = nbr_list + (size_t)cloud * kNbrGlobalCap * V + q;
  // Two-deep software pipeline: the index of neighbour t + 2 and the point / SPFH third of neighbour t + 1 are in flight while
  // neighbour t is accumulated (the chain index -> rows -> 11 dependent adds was one full L1/L2 latency per step: 66 % of the
  // scheduler cycles had no eligible warp).  The accumulation order is unchanged.
  int p_n = kq > 0 ? (int)gl[0] : 0;
  int p_nn = kq > 1 ? (int)gl[(size_t)V] : 0;
  float4 pp_n = P[p_n];
  float4 a0 = __ldg(sp + (size_t)p_n * (kDescPad / 4)), a1 = __ldg(sp + (size_t)p_n * (kDescPad / 4) + 1), a2 = __ldg(sp + (size_t)p_n * (kDescPad / 4) + 2);
  for (int t = 0; t < kq; ++t) {
    const float4 pp = pp_n, t0 = a0, t1 = a1, t2 = a2;
    if (t + 1 < kq) {
      const int pn = p_nn;
      if (t + 2 < kq) p_nn = (int)gl[(size_t)(t + 2) * V];
      pp_n = P[pn];
      a0 = __ldg(sp + (size_t)pn * (kDescPad / 4)); a1 = __ldg(sp + (size_t)pn * (kDescPad / 4) + 1); a2 = __ldg(sp + (size_t)pn * (kDescPad / 4) + 2);
    }
    const float dx = pq.x - pp.x, dy = pq.y - pp.y, dz = pq.z - pp.z;
    const float d2 = (dx * dx + dy * dy) + dz * dz;  // the same expression as the neighbour test: bit-identical
    if (d2 == 0.0f) continue;
    const float weight = 1.0f / d2;
    const float sv[11] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w, t2.x, t2.y, t2.z};
#pragma unroll
    for (int b = 0; b < 11; ++b) { const float v = sv[b] * weight; sum += v; o[b] += v; }
  }
  if (sum != 0.0) sum = 100.0 / sum;
  const float g = (float)sum;
  float* __restrict__ out = desc_t + (size_t)cloud * kDescK * V + (size_t)(11 * third) * V + q;
#pragma unroll
  for (int b = 0; b < 11; ++b) out[(size_t)b * V] = o[b] * g;
}

// the rare point with more than kNbrGlobalCap neighbours walks the lattice itself (one thread, all 33 bins)
__global__ void __launch_bounds__(kNbrThreads) fpfh_rare_kernel(const float4* __restrict__ pts, const int* __restrict__ n_pts, int V,
                                                                const uint64_t* __restrict__ cell_key, const int* __restrict__ cell_start,
                                                                const uint32_t* __restrict__ order, const int* __restrict__ n_cells, float inv,
                                                                int m, float r2, const float* __restrict__ spfh,
                                                                const int* __restrict__ nbr_cnt, float* __restrict__ desc_t) {
  __shared__ unsigned short nbr[kNbrCap][kNbrThreads];
  const int cloud = blockIdx.y;
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n_pts[cloud]) return;
  if (nbr_cnt[(size_t)cloud * V + q] <= kNbrGlobalCap) return;  // fpfh_list_kernel
  const LatticeView L = make_view(cloud, V, pts, cell_key, cell_start, order, n_cells, inv);
  const float4* __restrict__ sp = reinterpret_cast<const float4*>(spfh + (size_t)cloud * V * kDescPad);
  const float4 pq = L.pts[q];
  float o[kDescDim];
#pragma unroll
  for (int b = 0; b < kDescDim; ++b) o[b] = 0.0f;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0;
  auto accumulate = [&](int p) {
    const float4 pp = L.pts[p];
    const float dx = pq.x - pp.x, dy = pq.y - pp.y, dz = pq.z - pp.z;
    const float d2 = (dx * dx + dy * dy) + dz * dz;  // the same expression as the neighbour test: bit-identical
    if (d2 == 0.0f) return;
    const float weight = 1.0f / d2;
    float s[kDescPad];
#pragma unroll
    for (int v4 = 0; v4 < kDescPad / 4; ++v4) {
      const float4 t = __ldg(sp + (size_t)p * (kDescPad / 4) + v4);
      s[4 * v4] = t.x; s[4 * v4 + 1] = t.y; s[4 * v4 + 2] = t.z; s[4 * v4 + 3] = t.w;
    }
#pragma unroll
    for (int b = 0; b < 11; ++b) { const float v = s[b] * weight; s0 += v; o[b] += v; }
#pragma unroll
    for (int b = 0; b < 11; ++b) { const float v = s[12 + b] * weight; s1 += v; o[11 + b] += v; }
#pragma unroll
    for (int b = 0; b < 11; ++b) { const float v = s[24 + b] * weight; s2 += v; o[22 + b] += v; }
  };
  for_each_neighbor_listed(L, pq, m, r2, nbr, accumulate);
  if (s0 != 0.0) s0 = 100.0 / s0;
  if (s1 != 0.0) s1 = 100.0 / s1;
  if (s2 != 0.0) s2 = 100.0 / s2;
  const float g0 = (float)s0, g1 = (float)s1, g2 = (float)s2;
  float* __restrict__ out = desc_t + (size_t)cloud * kDescK * V + q;
#pragma unroll
  for (int b = 0; b < 11; ++b) out[(size_t)b * V] = o[b] * g0;
#pragma unroll
  for (int b = 11; b < 22; ++b) out[(size_t)b * V] = o[b] * g1;
#pragma unroll
  for (int b = 22; b < 33; ++b) out[(size_t)b * V] = o[b] * g2;
}

// descriptor layout converters for the stage API (pcl::FPFHSignature33 rows <-> dimension-major)
__global__ void desc_to_aos_kernel(const float* __restrict__ desc_t, int V, int n, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * kDescDim) return;
  const int q = i / kDescDim, d = i % kDescDim;
  out[i] = desc_t[(size_t)d * V + q];
}
__global__ void desc_from_aos_kernel(const float* __restrict__ in, int V, int n, float* __restrict__ desc_t) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * kDescDim) return;
  const int q = i / kDescDim, d = i % kDescDim;
  desc_t[(size_t)d * V + q] = in[i];
}

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
int launch_voxel(qb200_handle* h, int n_clouds, int total_raw, float leaf, int skip_flagged) {
  if (n_clouds <= 0) return QB200_OK;
  const float inv = 1.0f / leaf;
  const dim3 gk(64, n_clouds);
  const dim3 gb(n_clouds >= 16 ? 16 : 64, n_clouds);  // ~30 points per thread when the batch fills the device on its own
  int* chunk_cnt = reinterpret_cast<int*>(h->val_b);   // [clouds][kVsChunks]
  voxel_bbox_kernel<<<gb, 256, 0, h->stream>>>(h->d_cloud_ptr, h->d_cloud_n, inv, skip_flagged, h->ctr.bbox, h->ctr.n_valid, h->ctr.cloud_status,
                                               chunk_cnt);
  const int idx_bits = clog2(h->R > 2 ? h->R : 2);  // point index inside its scan
  // QB200_VOXEL_SORT=cub: round 1's device-wide library sort of every raw point (kept for A/B runs; results are identical)
  static const bool use_cub = getenv("QB200_VOXEL_SORT") && !strcmp(getenv("QB200_VOXEL_SORT"), "cub");
  const bool own = !use_cub && h->R >= 16384;   // the histogram scratch (val_a) holds 256 x R / 2048 words per cloud
  const uint64_t* keys_alt = nullptr;
  if (own) {
    h->launches += 1;
    if (int rc = launch_voxel_sort(h, n_clouds, inv, skip_flagged, idx_bits)) return rc;
    keys_alt = h->key_b;
  } else {
    voxel_keys_kernel<<<gk, 256, 0, h->stream>>>(h->d_cloud_ptr, h->d_cloud_n, h->d_raw_off, inv, skip_flagged, h->ctr.bbox, h->ctr.n_valid,
                                                 idx_bits, h->key_a);
    h->launches += 2;
    const int rc = sort_keys(h, total_raw, idx_bits, idx_bits + kVoxShift + clog2(n_clouds > 1 ? n_clouds : 2));
    if (rc) return rc;
  }
  const uint64_t* sorted = own ? h->key_a : h->key_b;
  run_heads_kernel<<<n_clouds, 1024, 0, h->stream>>>(0, sorted, keys_alt, h->d_raw_off, h->d_cloud_n, h->V, idx_bits, inv, h->ctr.bbox, h->ctr.n_valid,
                                                     h->vox_start, nullptr, h->ctr.n_vox, nullptr, h->ctr.cloud_status);
  const dim3 gc((h->V + 127) / 128, n_clouds);
  voxel_centroid_kernel<<<gc, 128, 0, h->stream>>>(h->d_cloud_ptr, h->d_raw_off, sorted, keys_alt, h->ctr.bbox, h->ctr.n_valid, inv,
                                                   (1ull << idx_bits) - 1, h->vox_start, h->ctr.n_vox, h->V, h->vox_pts);
  h->launches += 2;
  QB_CUDA_TRY(h, cudaGetLastError());
  return QB200_OK;
}

int launch_fpfh(qb200_handle* h, int n_clouds, float normal_radius, float fpfh_radius, float cell) {
  if (n_clouds <= 0) return QB200_OK;
  const int V = h->V;
  const float inv = 1.0f / cell;
  const int mn = (int)ceilf(normal_radius * inv + 1e-3f), mf = (int)ceilf(fpfh_radius * inv + 1e-3f);
  const float rn2 = (float)((double)normal_radius * (double)normal_radius), rf2 = (float)((double)fpfh_radius * (double)fpfh_radius);
  const dim3 gl((V + 255) / 256, n_clouds);
  lattice_keys_kernel<<<gl, 256, 0, h->stream>>>(h->vox_pts, h->ctr.n_vox, V, inv, h->key_a, h->val_a);
  h->launches++;
  // per-cloud shared-memory sort (sort.cu); clouds too large for it go through the device-wide radix sort
  int rc = launch_cloud_sort(h, n_clouds, h->ctr.n_vox, 18, 36);  // fields of cell_key(): i | j | k
  if (rc == QB200_ERR_UNSUPPORTED) rc = sort_pairs(h, n_clouds * V, kCloudShift + clog2(n_clouds > 1 ? n_clouds : 2));
  if (rc) return rc;
  run_heads_kernel<<<n_clouds, 1024, 0, h->stream>>>(1, h->key_b, nullptr, nullptr, nullptr, V, 0, inv, nullptr, nullptr, h->cell_start, h->cell_key,
                                                     h->ctr.n_cells, h->ctr.n_lat, h->ctr.cloud_status);
  const dim3 gp((V + 127) / 128, n_clouds);
  nbr_list_kernel<<<gp, kNbrThreads, 0, h->stream>>>(h->vox_pts, h->ctr.n_vox, V, h->cell_key, h->cell_start, h->val_b, h->ctr.n_cells, inv, mf,
                                                     rf2, h->nbr_list, h->nbr_cnt);
  // the list serves K3 when the normal neighbourhood is a subset visited in the same order: radius <= fpfh radius (checked
  // by the callers) and the same lattice reach for both walks
  const int list_usable = (mn <= mf && rn2 <= rf2) ? 1 : 0;
  normals_kernel<<<gp, 128, 0, h->stream>>>(h->vox_pts, h->ctr.n_vox, V, h->cell_key, h->cell_start, h->val_b, h->ctr.n_cells, inv, mn, rn2,
                                            h->nbr_list, h->nbr_cnt, list_usable, h->normals);
  // listed neighbourhoods (all but a handful of points) and the lattice-walking rest are separate launches: the common kernels
  // carry neither the walk's registers nor its list buffer
  spfh_kernel<false><<<gp, kSpfhThreads, 0, h->stream>>>(h->vox_pts, h->normals, h->ctr.n_vox, V, h->cell_key, h->cell_start, h->val_b,
                                                         h->ctr.n_cells, inv, mf, rf2, h->spfh, h->nbr_list, h->nbr_cnt);
  spfh_kernel<true><<<gp, kSpfhThreads, 0, h->stream>>>(h->vox_pts, h->normals, h->ctr.n_vox, V, h->cell_key, h->cell_start, h->val_b,
                                                        h->ctr.n_cells, inv, mf, rf2, h->spfh, h->nbr_list, h->nbr_cnt);
  fpfh_list_kernel<<<gp, 3 * kNbrThreads, 0, h->stream>>>(h->vox_pts, h->ctr.n_vox, V, h->spfh, h->nbr_list, h->nbr_cnt, h->desc_t);
  fpfh_rare_kernel<<<gp, kNbrThreads, 0, h->stream>>>(h->vox_pts, h->ctr.n_vox, V, h->cell_key, h->cell_start, h->val_b, h->ctr.n_cells, inv, mf,
                                                      rf2, h->spfh, h->nbr_cnt, h->desc_t);
  h->launches += 3;
  h->launches += 4;
  QB_CUDA_TRY(h, cudaGetLastError());
  return QB200_OK;
}

int launch_desc_to_aos(qb200_handle* h, int cloud, int n, float* d_out33) {
  if (n <= 0) return QB200_OK;
  desc_to_aos_kernel<<<(n * kDescDim + 255) / 256, 256, 0, h->stream>>>(h->desc_t + (size_t)cloud * kDescK * h->V, h->V, n, d_out33);
  h->launches++;
  QB_CUDA_TRY(h, cudaGetLastError());
  return QB200_OK;
}
// any dimension-major descriptor block [40][V] (e.g. a cache slot) -> AoS n x 33
int desc_to_aos_rows(qb200_handle* h, const float* desc_rows, int n, float* d_out33) {
  if (n <= 0) return QB200_OK;
  desc_to_aos_kernel<<<(n * kDescDim + 255) / 256, 256, 0, h->stream>>>(desc_rows, h->V, n, d_out33);
  h->launches++;
  QB_CUDA_TRY(h, cudaGetLastError());
  return QB200_OK;
}
int launch_desc_from_aos(qb200_handle* h, int cloud, int n, const float* d_in33) {
  if (n <= 0) return QB200_OK;
  desc_from_aos_kernel<<<(n * kDescDim + 255) / 256, 256, 0, h->stream>>>(d_in33, h->V, n, h->desc_t + (size_t)cloud * kDescK * h->V);
  h->launches++;
  QB_CUDA_TRY(h, cudaGetLastError());
  return QB200_OK;
}

}  // namespace qb
