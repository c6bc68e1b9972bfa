// tc_match.cu -- K6 on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.   (QB200_MATCH_EXACT=1 bypasses it.)
//
// The N_src x N_tgt x 33 descriptor-distance matrix is the one genuinely dense contraction of the path
// (north_star): d(i,j) = |a_i|^2 + |b_j|^2 - 2 a_i.b_j.  The reference does two exact 1-NN searches (FLANN kd-trees,
// src/teaser_utils/feature_matcher.cc:97-125); the exact answer here is DEFINED by the fp32 fma chain of match.cu
// (lowest-index ties).  Tensor cores cannot reproduce that rounding, so they act as a conservative FILTER and the
// survivors are evaluated exactly inside the same kernel -- results are bit-identical to the exact kernel.
//
//   norm_key_kernel   : |x - mu|^2 by the fp32 chain (mu = FPFH signature of a plane: 100 in bins 5, 16, 27 -- street scenes
//                       are dominated by planar points, centring makes THEIR dot products tiny and therefore the filter's
//                       absolute error tiny exactly where near-ties are dense) -> one radix sort orders every cloud by norm
//   dedup_kernel      : runs of bit-identical descriptors (adjacent after the sort) collapse to their first rank = lowest index
//   split_desc_kernel : hi = TF32(x'), lo = TF32(x' - hi), exact x, per block of 128 unique descriptors as ready-made
//                       shared-memory operand images (hi | lo | exact), so an image is ONE cp.async.bulk
//   tc_nn_kernel      : per 128-row stripe, column tiles in nearest-norm-first order; a tile whose lower bound
//                       (gap of the norm ranges)^2 exceeds every current best of the stripe's rows and of its columns is
//                       skipped unloaded; otherwise
//                         d~ = |a'|^2 + |b'|^2 - 2 (hi.hi + hi.lo + lo.hi)      3 x 5 tcgen05.mma.kind::tf32, fp32 in TMEM
//                         |d~ - d| <= e_ij = c/2 (|a'_i|^2 + |b'_j|^2),  c = 6e-5   (split + accumulation + chain rounding)
//                         (i,j) is evaluated EXACTLY (fp32 chain from the exact images) iff its lower bound d~ - e_ij
//                         does not exceed the best exact distance known so far for row i or for column j;
//                         row bests live in shared memory, column bests in global memory (atomicMin on packed
//                         distance|index words).
//                       A stripe that evaluates too many entries (massive near-ties) flags its pair, which is redone by the
//                       exact CUDA-core kernel, so results never depend on the filter.
//   broadcast_best_kernel : class results -> every member, in point order for the mutual-NN stage
//
// tc_nn_kernel, one CTA per SM, 19 warps, mbarrier hand-offs only: warp 18 chooses tiles and issues the exact-image copies
// (4 stages), warp 17 the operand copies (2 stages), warp 16 issues the 15 MMAs per tile into one of 4 TMEM accumulator stages,
// warps 0..15 drain them with tcgen05.ld.32x32b.x32 (lane = row) and run the filter / exact evaluation.
#include "handle.cuh"
#include <cstdlib>

namespace qb {

constexpr int kTcM = 128, kTcN = 128;
constexpr int kTcKB = kDescK / 8;                 // K blocks of 8 (TF32 MMA K)
constexpr int kTcTileBytes = kDescK * 128 * 4;    // one operand image (128 points x 40 dims) = 20480 B
constexpr int kTileFloats = kDescK * 128;         // 5120
constexpr int kTcImages = 3;                      // hi | lo | exact
constexpr int kTcEpiWarps = 16;                   // filter / evaluation warps: TMEM lane quadrant = warp & 3, column quarter = warp >> 2
constexpr int kTcThreads = (kTcEpiWarps + 3) * 32;  // + MMA warp + operand-copy warp + scheduler warp
constexpr int kTcAcc = 4;                          // TMEM accumulator stages (4 x 128 columns = all of TMEM)
constexpr int kTcStages = 4;                      // ring of exact B images (prefetch distance 3); operand images: 2 stages
constexpr float kTcC = 1.2e-4f;                   // |d~ - d| <= kTcC/2 * (|a'|^2 + |b'|^2): 3x the worst error measured (test_tc_filter_error_bound)
constexpr int kSpinLimit = 400000;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Operand images use the canonical K-major / no-swizzle UMMA layout (validated by tools/tc_probe.cu on B200; MN-major
// TF32 without swizzle yields zeros): core matrix = 8 points x 16 B (4 consecutive K values), byte offset
// kc*2048 + p*16 for K chunk kc (4 dims) and point p of the 128-point block.
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t addr) {
  // start address >> 4 | LBO = 2048 B (next 4-wide K chunk) | SBO = 128 B (next 8-point group) | version 1 (sm_100) | SWIZZLE_NONE
  return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)(2048 >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
}

__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}

__device__ __forceinline__ float tc_mu(int d) { return (d == 5 || d == 16 || d == 27) ? 100.0f : 0.0f; }

// centred squared norm of every descriptor (the fp32 chain split_desc_kernel repeats) as the sort key cloud | norm bits:
// K6 processes the points of a cloud in ascending-norm order, which turns the reverse triangle inequality
// d(a,b) >= (|a'| - |b'|)^2 into a tile-level lower bound (whole 128 x 128 blocks are skipped without being loaded).
__global__ void __launch_bounds__(256) norm_key_kernel(const float* __restrict__ desc_t, const int* __restrict__ n_vox, int V,
                                                       uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  const int cloud = blockIdx.y;
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= V) return;
  uint32_t bits = 0xFFFFFFFFu;  // padding sorts last
  if (q < n_vox[cloud]) {
    const size_t base = (size_t)cloud * kDescK * V + q;
    float acc = 0.0f;
#pragma unroll
    for (int d = 0; d < kDescDim; ++d) {
      const float xc = desc_t[base + (size_t)d * V] - tc_mu(d);
      acc = __fmaf_rn(xc, xc, acc);
    }
    bits = acc == acc ? __float_as_uint(acc) : 0xFFFFFFFEu;  // squared norms are >= 0: the bit pattern orders them
  }
  keys[(size_t)cloud * V + q] = ((uint64_t)cloud << 32) | bits;
  vals[(size_t)cloud * V + q] = (uint32_t)q;
}

// centred TF32 split + exact image + centred squared norm, written block-wise in the shared-memory operand layout;
// rank r of the cloud (ascending norm, ties by index) is point perm[r]
__global__ void __launch_bounds__(256) split_desc_kernel(const float* __restrict__ desc_t, const int* __restrict__ n_vox, int V,
                                                         const uint32_t* __restrict__ perm, float* __restrict__ tiles,
                                                         float* __restrict__ norm) {
  const int cloud = blockIdx.y;
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = n_vox[cloud];
  const int NB = V >> 7;
  if (q >= ((n + 127) & ~127)) return;  // the last block is padded with zeros so stale data never reaches the tensor core
  const size_t base = (size_t)cloud * kDescK * V + (q < n ? perm[(size_t)cloud * V + q] : 0u);
  const int blk = q >> 7, p = q & 127;
  float4* __restrict__ img = reinterpret_cast<float4*>(tiles + (size_t)(cloud * NB + blk) * kTcImages * kTileFloats) + p;
  // The tensor core delivers the filter's LOWER BOUND itself: rows (source clouds, even) carry x' in dims 0..32, kLow |x'|^2 in
  // dim 33 and 1 in dim 34; columns (target clouds, odd) carry -2 x', 1 and kLow |x'|^2 -- the contraction over the 35 dims is
  // kLow (|a'|^2 + |b'|^2) - 2 a'.b'.  (Scaling by -2 is exact; the norm term is split hi | lo like every other value.)
  const bool is_col = (cloud & 1) != 0;
  const float kLow = 1.0f - 0.5f * kTcC;
  float acc = 0.0f;
#pragma unroll
  for (int kc = 0; kc < kDescK / 4; ++kc) {
    float xv[4], hv[4], lv[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int d = 4 * kc + e;
      const bool live = d < kDescDim && q < n;
      const float x = live ? desc_t[base + (size_t)d * V] : 0.0f;
      const float xc = live ? x - tc_mu(d) : 0.0f;
      xv[e] = x;
      acc = __fmaf_rn(xc, xc, acc);
      float xs = is_col ? -2.0f * xc : xc;
      if (q < n && d == kDescDim) xs = is_col ? 1.0f : kLow * acc;       // acc is complete here: d runs upwards
      if (q < n && d == kDescDim + 1) xs = is_col ? kLow * acc : 1.0f;
      hv[e] = __uint_as_float(__float_as_uint(xs) & 0xFFFFE000u);
      lv[e] = __uint_as_float(__float_as_uint(xs - hv[e]) & 0xFFFFE000u);
    }
    img[0 * (kTileFloats / 4) + kc * 128] = make_float4(hv[0], hv[1], hv[2], hv[3]);
    img[1 * (kTileFloats / 4) + kc * 128] = make_float4(lv[0], lv[1], lv[2], lv[3]);
    if (kc < kDescK / 4 - 1) img[2 * (kTileFloats / 4) + kc * 128] = make_float4(xv[0], xv[1], xv[2], xv[3]);
  }
  // the exact image has no data in dims 36..39: slot 36 carries the column's filter term kLow |x'|^2 (+inf = padding)
  img[2 * (kTileFloats / 4) + (kDescK / 4 - 1) * 128] = make_float4(q < n ? kLow * acc : INFINITY, 0.0f, 0.0f, 0.0f);
  if (q < n) norm[(size_t)cloud * V + q] = acc;  // rank order
}

// Exact duplicates.  Street scenes contain hundreds of points with bit-identical descriptors (the histogram of a perfect
// plane, ...): every pair of them ties at distance 0 and would have to go through the exact chain.  Identical descriptors
// have identical norms, so after the norm sort they are neighbours: each run is collapsed to its first rank (the stable
// sort makes that the LOWEST point index, exactly the tie-break of the reference search), K6 runs on the unique
// descriptors only and broadcast_best_kernel hands the class result to every member.  One CTA per cloud.
__global__ void __launch_bounds__(1024) dedup_kernel(const float* __restrict__ desc_t, const int* __restrict__ n_vox, int V,
                                                     const uint64_t* __restrict__ sorted_keys, const uint32_t* __restrict__ perm, int enable,
                                                     uint32_t* __restrict__ uperm, uint32_t* __restrict__ class_of, int* __restrict__ n_unique) {
  __shared__ int scan_smem[33];
  const int cloud = blockIdx.x;
  const int n = n_vox[cloud];
  const float* __restrict__ D = desc_t + (size_t)cloud * kDescK * V;
  const uint32_t* __restrict__ pm = perm + (size_t)cloud * V;
  const uint64_t* __restrict__ sk = sorted_keys + (size_t)cloud * V;
  int base = 0;
  for (int start = 0; start < n; start += blockDim.x) {
    const int r = start + threadIdx.x;
    int flag = 0;
    if (r < n) {
      flag = 1;
      if (enable && r > 0 && sk[r] == sk[r - 1]) {  // same norm bits: compare the descriptors bit by bit
        const uint32_t a = pm[r], b = pm[r - 1];
        bool same = true;  // 11 dimensions per round trip (true duplicates need all 33: one dependent load pair each was 33 L2 latencies)
#pragma unroll
        for (int g = 0; g < 3; ++g) {
          if (!same) break;
          unsigned diff = 0;
#pragma unroll
          for (int j = 0; j < 11; ++j) {
            const int d = 11 * g + j;
            diff |= __float_as_uint(D[(size_t)d * V + a]) ^ __float_as_uint(D[(size_t)d * V + b]);
          }
          same = diff == 0;
        }
        flag = same ? 0 : 1;
      }
    }
    int total;
    const int ex = block_excl_scan(flag, scan_smem, &total);
    if (r < n) {
      const int u = base + ex + flag - 1;  // a duplicate belongs to the class opened by the closest earlier rank
      class_of[(size_t)cloud * V + r] = (uint32_t)u;
      if (flag) uperm[(size_t)cloud * V + u] = pm[r];
    }
    base += total;
  }
  if (threadIdx.x == 0) n_unique[cloud] = base;
}

// class results (unique-rank order) -> every point of the class, in point order for the mutual-NN stage
__global__ void __launch_bounds__(256) broadcast_best_kernel(const unsigned long long* __restrict__ rowbest_u,
                                                             const unsigned long long* __restrict__ colbest_u, const int* __restrict__ n_vox,
                                                             int V, const uint32_t* __restrict__ perm, const uint32_t* __restrict__ class_of,
                                                             unsigned long long* __restrict__ rowbest, unsigned long long* __restrict__ colbest) {
  const int cloud = blockIdx.y, pair = cloud >> 1;
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_vox[cloud]) return;
  const uint32_t u = class_of[(size_t)cloud * V + r], q = perm[(size_t)cloud * V + r];
  if (cloud & 1) colbest[(size_t)pair * V + q] = colbest_u[(size_t)pair * V + u];
  else rowbest[(size_t)pair * V + q] = rowbest_u[(size_t)pair * V + u];
}

// ---- mbarrier / bulk-copy helpers ---------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst), "l"(src), "r"(bytes),
               "r"(bar)
               : "memory");
}
// bounded wait: returns false if the barrier never completed (a descriptor bug must not hang the box)
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
  for (int spin = 0; spin < kSpinLimit; ++spin) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return true;
  }
  return false;
}

__device__ __forceinline__ unsigned long long tc_pack(float d, int idx) {
  return ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)idx;
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ unsigned tc_fkey(float f) {  // order-preserving float -> uint, NaN last
  unsigned key = __float_as_uint(f);
  key = (key & 0x80000000u) ? ~key : (key | 0x80000000u);
  return f == f ? key : 0xFFFFFFFFu;
}
__device__ __forceinline__ float tc_fkey_inv(unsigned key) { return __uint_as_float((key & 0x80000000u) ? (key & 0x7FFFFFFFu) : ~key); }

// rows = source cloud (2*pair), columns = target cloud (2*pair+1): their UNIQUE descriptors in ascending-norm (rank) order;
// n_vox / perm are the per-cloud unique counts and the point index of every unique rank.  rowbest / colbest_r are indexed by
// unique rank (broadcast_best_kernel maps them back).  kDbg: additionally
// dump d~ of the first tile of stripe 0 (validation hook).
//
// Warp roles (no CTA-wide barrier inside the tile loop, everything is handed over through mbarriers):
//   warp 17      : copy + schedule warp.  Picks the stripe's next column tile, nearest norm range first, and SKIPS a tile when
//                  its lower bound (gap between the norm ranges)^2 exceeds every current best of the stripe's rows and of
//                  the tile's columns; issues the bulk copies (2 operand stages, 4 exact-image stages)
//   warp 16      : one elected lane issues the 15 MMAs per tile into one of 4 TMEM accumulator stages
//   warps 0..15  : warp w owns TMEM lanes 32 (w & 3) .. +31 (rows) and columns 32 (w >> 2) .. +31 of every tile:
//                  tcgen05.ld -> release the TMEM stage -> branch-free filter -> the warp's survivors are compacted
//                  into batches of 32 and evaluated exactly, ONE CANDIDATE PER LANE -> release the exact-image stage.
// kProf: clock64 accounting of every role's waits into stats[8..31] (tools/tc_profile.py; QB200_TC_PROF=1)
template <bool kDbg, bool kProf = false>
__global__ void __launch_bounds__(kTcThreads, 1)
tc_nn_kernel(const float* __restrict__ tiles, const float* __restrict__ norm, const int* __restrict__ n_vox, int V,
             const uint32_t* __restrict__ perm, unsigned long long* __restrict__ rowbest, unsigned long long* __restrict__ colbest_r,
             unsigned* __restrict__ tile_cmax, int* __restrict__ fallback, unsigned long long* __restrict__ stats,
             float* __restrict__ dbg_tile, int no_prune) {
  extern __shared__ __align__(128) unsigned char smem[];  // 220 KB of operand images; static + dynamic must stay <= 227 KB
  __shared__ uint64_t s_fullx[kTcStages], s_sfree[kTcStages], s_fullhl[2], s_mma[kTcAcc], s_tfree[kTcAcc], s_afull;
  __shared__ uint32_t s_tmem;
  __shared__ unsigned long long s_rbest[kTcM];                 // best exact (distance | target index) per row of the stripe
  __shared__ __align__(16) float s_wnbm[kTcEpiWarps][32];      // per warp: kLow |b'_j|^2 of its 32 columns
  __shared__ __align__(16) float s_wcj[kTcEpiWarps][32];       //           kLow |b'_j|^2 - (best exact distance of column j)
  __shared__ unsigned short s_queue[kTcEpiWarps][32];          //           one batch of candidates (row lane << 5 | column)
  __shared__ int s_seq[8];                                     // column tile of sequence position n (ring), -1 = end of the stripe
  __shared__ float s_tlb[128];                                 // lower bound of every distance between the stripe and column tile t
  __shared__ int s_dead, s_abort, s_evals, s_warm, s_npos;
  __shared__ int s_ndec;                                       // sequence positions 0 .. s_ndec-1 have been decided (s_seq ring)

  const int pair = blockIdx.y, stripe = blockIdx.x;
  const int cloudA = 2 * pair, cloudB = 2 * pair + 1;
  const int nA = n_vox[cloudA], nB = n_vox[cloudB];
  const int r0 = stripe * kTcM;
  if (r0 >= nA || nB <= 0) return;  // uniform for the CTA, before any barrier / allocation

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int NB = V >> 7;
  auto tick = [&]() -> long long { return kProf ? clock64() : 0ll; };
  auto prof = [&](int slot, long long cyc) { if (kProf && lane == 0) atomicAdd(stats + slot, (unsigned long long)cyc); };
  const long long t_begin = tick();
  // shared-memory map: A block (hi | lo | exact) | 2 stages of B (hi | lo) | 4 stages of B exact images.  The operand images
  // are dead as soon as the MMAs of their tile completed, the exact image only when every warp evaluated the tile.
  constexpr uint32_t kABytes = kTcImages * kTcTileBytes;
  constexpr uint32_t kHLBytes = 2 * kTcTileBytes;
  constexpr uint32_t kXBytes = kTcTileBytes;
  const uint32_t sA = smem_u32(smem), sHL0 = sA + kABytes, sX0 = sHL0 + 2 * kHLBytes;
  const float* __restrict__ tA = tiles + (size_t)cloudA * NB * kTcImages * kTileFloats;
  const float* __restrict__ tB = tiles + (size_t)cloudB * NB * kTcImages * kTileFloats;
  const float* __restrict__ nrmA = norm + (size_t)cloudA * V;   // rank order, ascending
  const float* __restrict__ nrmB = norm + (size_t)cloudB * V;
  const uint32_t* __restrict__ permA = perm + (size_t)cloudA * V;
  const uint32_t* __restrict__ permB = perm + (size_t)cloudB * V;
  unsigned long long* __restrict__ cbg = colbest_r + (size_t)pair * V;
  const uint32_t bar_fullx0 = smem_u32(&s_fullx[0]), bar_sfree0 = smem_u32(&s_sfree[0]), bar_fullhl0 = smem_u32(&s_fullhl[0]),
                 bar_mma0 = smem_u32(&s_mma[0]), bar_tfree0 = smem_u32(&s_tfree[0]), bar_a = smem_u32(&s_afull);
  const int n_tiles = (nB + kTcN - 1) / kTcN;

  if (warp == 0) {  // TMEM: 4 accumulator stages x 128 fp32 columns
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&s_tmem)), "r"(kTcAcc * kTcN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < kTcStages; ++i) { mbar_init(bar_fullx0 + 8 * i, 1); mbar_init(bar_sfree0 + 8 * i, kTcEpiWarps); }
    for (int i = 0; i < 2; ++i) mbar_init(bar_fullhl0 + 8 * i, 1);
    for (int i = 0; i < kTcAcc; ++i) { mbar_init(bar_mma0 + 8 * i, 1); mbar_init(bar_tfree0 + 8 * i, kTcEpiWarps); }
    mbar_init(bar_a, 1);
    s_dead = 0; s_abort = 0; s_evals = 0; s_warm = 0; s_npos = 0; s_ndec = 0;
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (threadIdx.x < kTcM) s_rbest[threadIdx.x] = ~0ull;
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = s_tmem;
  const long long t_setup = tick();
  if (kProf && threadIdx.x == 0) { atomicAdd(stats + 8, 1ull); atomicAdd(stats + 10, (unsigned long long)(t_setup - t_begin)); }
  volatile int* v_dead = &s_dead;
  volatile int* v_abort = &s_abort;
  volatile int* v_seq = s_seq;
  volatile int* v_ndec = &s_ndec;
  // bounded wait until sequence position n has been decided by the scheduler warp
  auto wait_decided = [&](int n) -> bool {
    for (int spin = 0; spin < kSpinLimit; ++spin) {
      if (*v_ndec > n) { __threadfence_block(); return true; }
      __nanosleep(32);
    }
    return false;
  };
  volatile unsigned long long* v_rbest = s_rbest;

  if (warp == kTcEpiWarps) {
    // ================= MMA warp =================
    // instruction descriptor: D=F32 (bits 4-5), A=B=TF32 (bits 7-9, 10-12), both K-major (bits 15,16 = 0), N>>3 (17-22), M>>4 (24-28)
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kTcN >> 3) << 17) | ((uint32_t)(kTcM >> 4) << 24);
    bool ok = mbar_wait(bar_a, 0);
    long long p_hl = 0, p_tf = 0, p_is = 0;
    prof(15, tick() - t_setup);
    for (int k = 0; ok; ++k) {
      const int ts = k & (kTcAcc - 1), hs = k & 1;
      const long long t0 = tick();
      ok = mbar_wait(bar_fullhl0 + 8 * hs, (uint32_t)((k >> 1) & 1));
      const long long t1 = tick();
      p_hl += t1 - t0;
      if (!ok || v_seq[k & 7] < 0) break;
      if (k >= kTcAcc) ok = mbar_wait(bar_tfree0 + 8 * ts, (uint32_t)(((k >> 2) - 1) & 1));  // accumulator stage drained (position k-4)
      const long long t2 = tick();
      p_tf += t2 - t1;
      if (!ok) break;
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      if (lane == 0) {
        const uint32_t aH = sA, aL = sA + kTcTileBytes, bH = sHL0 + hs * kHLBytes, bL = bH + kTcTileBytes, d = tmem + (uint32_t)(ts * kTcN);
        uint32_t acc = 0;
#pragma unroll
        for (int kb = 0; kb < kTcKB; ++kb) {  // small cross terms first, then hi.hi
          tc_mma_tf32(d, tc_smem_desc(aH + kb * 4096), tc_smem_desc(bL + kb * 4096), idesc, acc);
          acc = 1;
          tc_mma_tf32(d, tc_smem_desc(aL + kb * 4096), tc_smem_desc(bH + kb * 4096), idesc, 1);
        }
#pragma unroll
        for (int kb = 0; kb < kTcKB; ++kb) tc_mma_tf32(d, tc_smem_desc(aH + kb * 4096), tc_smem_desc(bH + kb * 4096), idesc, 1);
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar_mma0 + 8 * ts) : "memory");
      }
      __syncwarp();
      p_is += tick() - t2;
    }
    prof(16, p_hl); prof(17, p_tf); prof(18, p_is);
    if (!ok) *v_dead = 1;
  } else if (warp == kTcEpiWarps + 1) {
    // ================= operand-copy warp: A images once, then the hi | lo images of every decided position =================
    auto issue_hl = [&](int n) {  // operand images (hi | lo, 40 KB) of position n -> stage n & 1; end marker: plain arrive
      const int t = v_seq[n & 7];
      if (lane != 0) return;
      const uint32_t bar = bar_fullhl0 + 8 * (n & 1);
      if (t < 0) { mbar_arrive(bar); return; }
      mbar_expect_tx(bar, kHLBytes);
      bulk_g2s(sHL0 + (n & 1) * kHLBytes, tB + (size_t)t * kTcImages * kTileFloats, kHLBytes, bar);
    };
    if (lane == 0) {
      mbar_expect_tx(bar_a, kABytes);
      bulk_g2s(sA, tA + (size_t)stripe * kTcImages * kTileFloats, kABytes, bar_a);
    }
    bool ok = wait_decided(0);
    if (ok) { issue_hl(0); ok = wait_decided(1); }
    if (ok) issue_hl(1);
    long long p_wm = 0, p_wd = 0;
    bool ended = !ok || v_seq[0] < 0 || v_seq[1] < 0;  // (an end marker has gone out already)
    for (int k = 0; ok && !ended; ++k) {
      // operand stage k & 1 is free once the MMAs of position k completed
      const long long t0 = tick();
      ok = mbar_wait(bar_mma0 + 8 * (k & (kTcAcc - 1)), (uint32_t)((k >> 2) & 1));
      const long long t1 = tick();
      p_wm += t1 - t0;
      if (ok) ok = wait_decided(k + 2);
      p_wd += tick() - t1;
      if (!ok) break;
      issue_hl(k + 2);
      ended = v_seq[(k + 2) & 7] < 0;  // that was the end marker
    }
    prof(12, p_wm); prof(28, p_wd);
    if (!ok) *v_dead = 1;
  } else if (warp == kTcEpiWarps + 2) {
    // ================= scheduler warp: chooses the tiles and issues the exact-image copies =================
    // norm range of the stripe's rows and of a column tile (ranks are sorted by norm: first / last valid entry)
    const float amin = sqrtf(nrmA[r0]), amax = sqrtf(nrmA[(r0 + kTcM < nA ? r0 + kTcM : nA) - 1]);
    auto tile_lb = [&](int t) -> float {  // lower bound of every exact distance between the stripe and column tile t
      const float bmin = sqrtf(nrmB[t * kTcN]), bmax = sqrtf(nrmB[(t * kTcN + kTcN < nB ? t * kTcN + kTcN : nB) - 1]);
      const float gap = fmaxf(amin - bmax, bmin - amax) - 2.0e-3f;  // slack: rounding of the norm chains and square roots
      if (!(amin <= amax && bmin <= bmax)) return 0.0f;               // NaN norms: never skip
      return gap > 0.0f ? gap * gap * 0.9999f : 0.0f;
    };
    // Shared per-tile column maxima: tcm4[t][g] = an upper bound of the best exact distances of columns 32 g .. 32 g + 31 of tile t
    // (float bits; every filter warp refreshes its group after evaluating the tile, bests only shrink).  The values of the first
    // 128 tiles (4 tiles per lane) are fetched at the END of a choice for the NEXT one, so no choice waits for L2.
    const uint4* __restrict__ tcm4 = reinterpret_cast<const uint4*>(tile_cmax) + (size_t)pair * (V >> 7);
    uint4 pq0, pq1, pq2, pq3;
    auto fetch_tcm = [&]() {
      pq0 = pq1 = pq2 = pq3 = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
      if (!no_prune) {
        if (lane < n_tiles) pq0 = __ldcg(tcm4 + lane);
        if (lane + 32 < n_tiles) pq1 = __ldcg(tcm4 + lane + 32);
        if (lane + 64 < n_tiles) pq2 = __ldcg(tcm4 + lane + 64);
        if (lane + 96 < n_tiles) pq3 = __ldcg(tcm4 + lane + 96);
      }
    };
    auto max4 = [](const uint4& q) -> unsigned { return max(max(q.x, q.y), max(q.z, q.w)); };
    fetch_tcm();
    // lower bounds of all column tiles (at most 128: V <= 16384 ... 65536 / 128 = 512 tiles are covered by the loop stride),
    // kept in shared memory; start at the tile whose norm range is closest to the stripe's, then walk outwards on both sides
    int t0 = 0;
    {
      float best = INFINITY;
      for (int t = lane; t < n_tiles; t += 32) {
        const float g = tile_lb(t);
        if (t < 128) s_tlb[t] = g;
        if (g < best) { best = g; t0 = t; }
      }
      const unsigned key = __reduce_min_sync(0xffffffffu, tc_fkey(best));
      const unsigned who = __ballot_sync(0xffffffffu, tc_fkey(best) == key);
      t0 = __shfl_sync(0xffffffffu, t0, __ffs(who) - 1);
      __syncwarp();
    }
    auto tlb = [&](int t) -> float { return t < 128 ? s_tlb[t] : tile_lb(t); };
    int lo = t0 - 1, hi = t0 + 1;
    bool first = true, done = false;
    auto next_tile = [&]() -> int {  // warp-uniform
      if (first) { first = false; return t0; }
      // Inputs of a choice: the current worst row best of the stripe (+inf while any row is unknown) and the prefetched per-tile
      // column maxima.  A tile whose lower bound exceeds both is skipped for good (bests only shrink).
      unsigned rmax = 0;
#pragma unroll
      for (int i = 0; i < kTcM / 32; ++i) {
        const int row = lane + 32 * i;
        if (r0 + row < nA) {
          const unsigned long long rb = v_rbest[row];
          rmax = max(rmax, rb == ~0ull ? 0x7F800000u : (unsigned)(rb >> 32));
        }
      }
      const float rmaxf = __uint_as_float(__reduce_max_sync(0xffffffffu, rmax));  // distances are >= 0: the bit patterns order them
      const unsigned tc0 = max4(pq0), tc1 = max4(pq1), tc2 = max4(pq2), tc3 = max4(pq3);
      if (*v_abort || *v_dead) return -1;
      // The walk, 32 candidates per round: lanes 0..15 look at the next 16 tiles on the left (lo, lo-1, ...), lanes 16..31 at the
      // next 16 on the right.  Lower bounds grow outwards on either side, so the next tile in nearest-first order is the first
      // one that cannot be skipped on the left or on the right, whichever has the smaller bound (left on ties); the tiles in
      // front of them are skipped for good.
      const bool is_left = lane < 16;
      while (true) {
        if (lo < 0 && hi >= n_tiles) return -1;
        const int t = is_left ? lo - lane : hi + (lane - 16);
        const bool valid = is_left ? t >= 0 : t < n_tiles;
        const float lb = valid ? tlb(t) : INFINITY;
        unsigned cm = 0xFFFFFFFFu;
        {
          const int src = t & 31, sl = (t >> 5) & 3;
          const unsigned a0 = __shfl_sync(0xffffffffu, tc0, src), a1 = __shfl_sync(0xffffffffu, tc1, src);
          const unsigned a2 = __shfl_sync(0xffffffffu, tc2, src), a3 = __shfl_sync(0xffffffffu, tc3, src);
          if (valid && t < 128) cm = sl == 0 ? a0 : sl == 1 ? a1 : sl == 2 ? a2 : a3;
          else if (valid && !no_prune) cm = max4(__ldcg(tcm4 + t));
        }
        const bool visit = valid && (no_prune || lb <= 0.0f || !(lb > rmaxf) || !(lb > __uint_as_float(cm)));
        const unsigned vb = __ballot_sync(0xffffffffu, visit);
        const int fl = (vb & 0xFFFFu) ? __ffs(vb & 0xFFFFu) - 1 : 16;   // first tile to visit on either side (16 = none in this window)
        const int fr = (vb >> 16) ? __ffs(vb >> 16) - 1 : 16;
        if (fl == 16 && fr == 16) { lo -= 16; hi += 16; continue; }
        const float gl = __shfl_sync(0xffffffffu, lb, fl & 15), gh = __shfl_sync(0xffffffffu, lb, 16 + (fr & 15));
        const bool left = fr == 16 || (fl < 16 && gl <= gh);
        const int tv = left ? lo - fl : hi + fr;
        lo -= left ? fl + 1 : fl;
        hi += left ? fr : fr + 1;
        return tv;
      }
    };
    int n_issued = 0;
    auto decide = [&](int n) {  // fix the tile of sequence position n
      int t = -1;
      if (!done) {
        t = next_tile();
        if (t < 0) done = true; else ++n_issued;
      }
      if (lane == 0) {  // (no loads are in flight here: the fence is cheap)
        v_seq[n & 7] = t;
        __threadfence_block();
        *v_ndec = n + 1;
      }
      __syncwarp();
      if (!done) fetch_tcm();  // for the next choice: lands while this warp waits for the MMAs / the free exact-image stage
      return t;
    };
    auto issue_x = [&](int n) {  // exact image (20 KB) of position n -> stage n & 3
      const int t = v_seq[n & 7];
      if (lane != 0) return;
      const uint32_t bar = bar_fullx0 + 8 * (n & (kTcStages - 1));
      if (t < 0) { mbar_arrive(bar); return; }
      mbar_expect_tx(bar, kXBytes);
      bulk_g2s(sX0 + (n & (kTcStages - 1)) * kXBytes, tB + ((size_t)t * kTcImages + 2) * kTileFloats, kXBytes, bar);
    };
    for (int n = 0; n < 3; ++n) { decide(n); issue_x(n); }
    prof(11, tick() - t_setup);
    long long p_wm = 0, p_dec = 0, p_sf = 0;
    auto mbar_test = [&](uint32_t bar, uint32_t parity) -> bool {
      uint32_t ready;
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t}\n"
          : "=r"(ready)
          : "r"(bar), "r"(parity)
          : "memory");
      return __all_sync(0xffffffffu, ready != 0);
    };
    // Two duties, neither blocks the other: the exact image of position nx goes out as soon as every warp evaluated position nx-4
    // (its stage), and positions are chosen ahead (at most 3 beyond nx: the s_seq ring, and not before the MMAs of position nd-3
    // completed, so that a choice sees the bests of the tiles before it) while that stage is still busy.
    int nd = 3, nx = 3;
    bool ok = true, end_decided = v_seq[0] < 0 || v_seq[1] < 0 || v_seq[2] < 0;
    if (end_decided) nx = nd;  // (all three end-marker arrivals were issued above; nothing else to do)
    int idle = 0;
    long long t_idle = 0;
    while (ok && !(end_decided && nx == nd)) {
      const uint32_t bar_s = bar_sfree0 + 8 * (nx & (kTcStages - 1)), par_s = (uint32_t)(((nx - 4) >> 2) & 1);
      const uint32_t bar_m = bar_mma0 + 8 * ((nd - 3) & (kTcAcc - 1)), par_m = (uint32_t)(((nd - 3) >> 2) & 1);
      const bool can_x = nx < nd, can_d = !end_decided && nd < nx + 4;
      if (can_x && (nx < 4 || mbar_test(bar_s, par_s))) {
        if (kProf && idle) { (can_d ? p_wm : p_sf) += tick() - t_idle; }
        idle = 0;
        issue_x(nx);
        ++nx;
        continue;
      }
      if (can_d && mbar_test(bar_m, par_m)) {
        const long long t1 = tick();
        if (kProf && idle) { (can_x ? p_wm : p_sf) += t1 - t_idle; }
        idle = 0;
        if (decide(nd) < 0) end_decided = true;
        ++nd;
        p_dec += tick() - t1;
        continue;
      }
      // nothing is ready: poll both events (bounded)
      if (idle == 0) t_idle = tick();
      if (++idle > kSpinLimit) { ok = false; break; }
      __nanosleep(32);
    }
    prof(29, p_wm); prof(13, p_dec); prof(14, p_sf);
    if (!ok) *v_dead = 1;
    if (lane == 0) s_npos = n_issued;
  } else {
    // ================= filter / evaluation warps =================
    const int quad = warp & 3, cq = warp >> 2, cb = cq * 32;
    const int row = quad * 32 + lane, gi = r0 + row;
    const bool row_ok = gi < nA;
    const float4* __restrict__ aex = reinterpret_cast<const float4*>(smem + 2 * kTcTileBytes) + quad * 32;  // exact image of the A block
    // TMEM holds LB_ij = d~_ij - e_ij = kLow (na' + nb') - 2 dot (split_desc_kernel); an entry is a candidate iff LB_ij <= the best
    // exact distance known for row i or for column j
    const float kLow = 1.0f - 0.5f * kTcC;
    const float kW = kTcC / kLow;                            // d~ + e = LB + kW (kLow na' + kLow nb')
    const float nam = row_ok ? kLow * nrmA[gi] : INFINITY;   // +inf: a padded row yields no upper bounds
    const unsigned oa = row_ok ? permA[gi] : 0u;             // point index of this lane's row
    float Ri_ub = row_ok ? INFINITY : -INFINITY;             // row threshold from the tile-local upper bounds (padded rows: never)
    float* wnbm = s_wnbm[warp];
    float* wcj = s_wcj[warp];
    unsigned short* wq = s_queue[warp];
    int evals_w = 0;
    bool alive = mbar_wait(bar_a, 0);
    const long long t_loop = tick();
    prof(19, t_loop - t_setup);
    long long p_wx = 0, p_wmm = 0, p_ld = 0, p_prep = 0, p_fil = 0, p_ev = 0;
    int p_tiles = 0;
    // column snapshot (best | point index) of the NEXT tile, fetched while the current one is processed when the scheduler
    // has already published it (pf_tile = tile the prefetch belongs to, -1 = none)
    int pf_tile = -1;
    unsigned long long pf_cb = ~0ull;
    unsigned pf_ob = 0;
    // shared per-tile column maxima (scheduler warp): after a tile is evaluated its 32 column bests are read again, and one tile
    // later (the load has landed) their max refreshes tcm4[tile][cq]
    unsigned* __restrict__ tcmw = tile_cmax + ((size_t)pair * (V >> 7)) * 4 + cq;
    int rr_tile = -1;
    unsigned rr_val = 0;
    auto rr_flush = [&]() {
      if (rr_tile < 0) return;
      const unsigned m = __reduce_max_sync(0xffffffffu, rr_val);
      if (lane == 0) atomicMin(tcmw + (size_t)rr_tile * 4, m);
      rr_tile = -1;
    };
    for (int k = 0;; ++k) {
      const int ts = k & (kTcAcc - 1), st = k & (kTcStages - 1);
      const long long q0 = tick();
      if (alive) alive = mbar_wait(bar_fullx0 + 8 * st, (uint32_t)((k >> 2) & 1));  // the exact image is read below
      const long long q1 = tick();
      p_wx += q1 - q0;
      alive = __all_sync(0xffffffffu, alive);
      if (!alive) { *v_dead = 1; break; }
      const int jt = v_seq[k & 7];
      if (jt < 0) break;
      const int c0 = jt * kTcN;
      // snapshot of the column's best and its point index (lane = column); the load overlaps the wait for the MMAs
      unsigned long long cb_cur = ~0ull;
      unsigned ob = 0;
      if (pf_tile == jt) {
        cb_cur = pf_cb; ob = pf_ob;
      } else if (c0 + cb + lane < nB) {
        cb_cur = __ldcg(cbg + c0 + cb + lane);
        ob = permB[c0 + cb + lane];
      }
      pf_tile = -1;
      {  // the next position is normally decided already (the scheduler runs 3 ahead): start its snapshot loads now
        if (__all_sync(0xffffffffu, *v_ndec > k + 1)) {  // (volatile shared loads of one warp are performed in order: no fence,
          const int jn = v_seq[(k + 1) & 7];               //  which would wait for this warp's global loads in flight)
          if (jn >= 0) {
            pf_tile = jn; pf_cb = ~0ull; pf_ob = 0;
            if (jn * kTcN + cb + lane < nB) {
              pf_cb = __ldcg(cbg + jn * kTcN + cb + lane);
              pf_ob = permB[jn * kTcN + cb + lane];
            }
          }
        }
      }
      const long long q2 = tick();
      alive = mbar_wait(bar_mma0 + 8 * ts, (uint32_t)((k >> 2) & 1));
      const long long q3 = tick();
      p_wmm += q3 - q2;
      p_prep += q2 - q1;
      ++p_tiles;
      alive = __all_sync(0xffffffffu, alive);
      if (!alive) { *v_dead = 1; break; }
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      const bool skip = __any_sync(0xffffffffu, (*v_abort | *v_dead) != 0);
      uint32_t v[32];
      if (!skip) tc_ld32(tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(ts * kTcN + cb), v);
      asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tfree0 + 8 * ts);  // accumulators are in registers: the stage may be overwritten
      const long long q4 = tick();
      p_ld += q4 - q3;
      if (!skip) {
        const float4* __restrict__ bex = reinterpret_cast<const float4*>(smem + kABytes + 2 * kHLBytes + st * kXBytes);  // exact image
        // ---- per-column filter data of this warp's 32 columns (lane = column)
        const float nbm = bex[9 * 128 + cb + lane].x;  // kLow |b'_j|^2, +inf for padded columns (split_desc_kernel)
        const float dbest = cb_cur == ~0ull ? INFINITY : __uint_as_float((unsigned)(cb_cur >> 32));
        float cj = nbm == INFINITY ? -INFINITY : dbest;  // column threshold: +inf while the column has no exact distance yet
        const unsigned long long rb = v_rbest[row];
        float Ri = fminf(Ri_ub, rb == ~0ull ? INFINITY : __uint_as_float((unsigned)(rb >> 32)));
        wnbm[lane] = nbm;
        __syncwarp();
        const float4* __restrict__ wn4 = reinterpret_cast<const float4*>(wnbm);
        const float4* __restrict__ wc4 = reinterpret_cast<const float4*>(wcj);
        // ---- warm-up: a row / column without any exact distance would let every entry through.  Upper bounds of the
        // exact minima come from the tile itself: UB_ij = d~_ij + e_ij = (1+c/2)(na'+nb') - 2 dot >= d_ij.
        if (__any_sync(0xffffffffu, (row_ok && Ri == INFINITY) || (cb_cur == ~0ull && nbm != INFINITY))) {
          if (lane == 0) atomicAdd(&s_warm, 1);
          float rowub = INFINITY;
          unsigned mine = 0xFFFFFFFFu;
#pragma unroll
          for (int c4 = 0; c4 < 8; ++c4) {
            const float4 w = wn4[c4];
            const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int c = 4 * c4 + e;
              const float ub = fmaf(kW, nam + wv[e], __uint_as_float(v[c]));
              rowub = fminf(rowub, ub);
              // min over the warp's 32 rows; an upper bound of a distance is >= 0 (bit patterns order them), anything else is dropped
              const unsigned mn = __reduce_min_sync(0xffffffffu, ub >= 0.0f ? __float_as_uint(ub) : 0xFFFFFFFFu);
              if (lane == c) mine = mn;
            }
          }
          if (mine != 0xFFFFFFFFu && nbm != INFINITY) cj = fminf(cj, __uint_as_float(mine));
          if (row_ok) { Ri_ub = fminf(Ri_ub, rowub); Ri = fminf(Ri, Ri_ub); }
        }
        wcj[lane] = cj;
        __syncwarp();
        const long long q5 = tick();
        p_prep += q5 - q4;
        // ---- branch-free filter of this lane's row over the 32 columns: two compares and a predicated OR per entry
        uint32_t mask = 0;
#pragma unroll
        for (int c4 = 0; c4 < 8; ++c4) {
          const float4 x = wc4[c4];
          const float xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int c = 4 * c4 + e;
            const float lbv = __uint_as_float(v[c]);
            asm("{\n\t.reg .pred p, q;\n\t"
                "setp.le.f32 p, %1, %2;\n\t"
                "setp.le.or.f32 q, %1, %3, p;\n\t"
                "@q or.b32 %0, %0, %4;\n\t}\n"
                : "+r"(mask)
                : "f"(lbv), "f"(xv[e]), "f"(Ri), "r"(1u << c));
            if (kDbg) {
              const unsigned oc = __shfl_sync(0xffffffffu, ob, c);
              const float nbc = wnbm[c];
              if (stripe == 0 && k == 0 && row_ok && c0 + cb + c < nB) dbg_tile[(size_t)oa * kTcN + oc] = fmaf(0.5f * kW, nam + nbc, lbv);
            }
          }
        }
        // padding never competes: a padded row (-inf thresholds) would pass the column test of a column without a best
        // (-inf <= -inf), a padded column (+inf terms) the row test of a row without one -- and their zero images look
        // like the all-zero descriptor of an isolated point
        mask &= __ballot_sync(0xffffffffu, c0 + cb + lane < nB);
        if (!row_ok) mask = 0;
        // ---- exact evaluation, one candidate per lane, 32 per round
        int remaining = __reduce_add_sync(0xffffffffu, __popc(mask));
        const long long q6 = tick();
        p_fil += q6 - q5;
        evals_w += remaining;
        while (remaining > 0) {  // warp-uniform
          int tot;
          int p = warp_excl_scan(__popc(mask), &tot);
          while (mask && p < 32) {
            const int c = __ffs(mask) - 1;
            mask &= mask - 1;
            wq[p++] = (unsigned short)((lane << 5) | c);
          }
          __syncwarp();
          const int nb = tot < 32 ? tot : 32;
          const int q = lane < nb ? wq[lane] : 0;
          const int rl = q >> 5, c = q & 31;
          const int pcol = cb + c;
          float acc = 0.0f;
#pragma unroll
          for (int kc = 0; kc < (kDescDim + 3) / 4; ++kc) {
            const float4 a = aex[kc * 128 + rl], b = bex[kc * 128 + pcol];
            float diff = a.x - b.x;
            acc = __fmaf_rn(diff, diff, acc);
            if (4 * kc + 1 < kDescDim) { diff = a.y - b.y; acc = __fmaf_rn(diff, diff, acc); }
            if (4 * kc + 2 < kDescDim) { diff = a.z - b.z; acc = __fmaf_rn(diff, diff, acc); }
            if (4 * kc + 3 < kDescDim) { diff = a.w - b.w; acc = __fmaf_rn(diff, diff, acc); }
          }
          const float dbc = __shfl_sync(0xffffffffu, dbest, c);
          const unsigned oc = __shfl_sync(0xffffffffu, ob, c);   // point index of the candidate's column
          const unsigned orow = __shfl_sync(0xffffffffu, oa, rl);  // ... and of its row
          if (lane < nb && acc == acc) {  // NaN never wins
            const int rr = quad * 32 + rl;
            const unsigned long long pr = tc_pack(acc, (int)oc);
            if (pr < v_rbest[rr]) atomicMin(&s_rbest[rr], pr);
            if (acc <= dbc) atomicMin(cbg + c0 + pcol, tc_pack(acc, (int)orow));
          }
          __syncwarp();
          remaining = tot - nb;
        }
        // massive ties: once more than half of the entries seen needed the exact chain, hand the pair to the exact kernel
        if (lane == 0 && evals_w) {
          const int seen = atomicAdd(&s_evals, evals_w) + evals_w;
          if (k >= 7 && seen > (k + 1) * (kTcM * kTcN / 2)) *v_abort = 1;
        }
        evals_w = 0;
        p_ev += tick() - q6;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_sfree0 + 8 * st);  // the exact-image stage may be refilled
      rr_flush();
      if (!skip) {
        rr_tile = jt;
        rr_val = 0;  // padding columns do not count
        if (c0 + cb + lane < nB) {
          const unsigned long long cbn = __ldcg(cbg + c0 + cb + lane);
          rr_val = cbn == ~0ull ? 0x7F800000u : (unsigned)(cbn >> 32);
        }
      }
    }
    rr_flush();
    prof(20, p_wx); prof(21, p_wmm); prof(22, p_ld); prof(23, p_prep); prof(24, p_fil); prof(25, p_ev);
    prof(26, tick() - t_loop); prof(27, p_tiles);
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  const bool aborted = (s_abort | s_dead) != 0;
  if (kProf && threadIdx.x == 0) atomicAdd(stats + 9, (unsigned long long)(tick() - t_begin));
  if (threadIdx.x == 0) {
    atomicAdd(stats + 0, (unsigned long long)s_evals);
    atomicAdd(stats + 1, (unsigned long long)s_npos);
    atomicAdd(stats + 2, (unsigned long long)s_warm);
    if (aborted) {
      atomicAdd(stats + 3, 1ull);
      fallback[pair] = 1;
    }
  }
  if (!aborted && threadIdx.x < kTcM && r0 + (int)threadIdx.x < nA) rowbest[(size_t)pair * V + r0 + threadIdx.x] = s_rbest[threadIdx.x];
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(kTcAcc * kTcN) : "memory");
}

static size_t tc_smem_bytes() { return (size_t)kTcImages * kTcTileBytes + 2 * 2 * (size_t)kTcTileBytes + (size_t)kTcStages * kTcTileBytes; }  // 220 KB

// norm keys -> one radix sort for all clouds (h->val_b = rank -> point) -> duplicate classes.
// Scratch (free once the sort consumed its inputs): val_a = unique rank -> point, key_a = [class of rank | unique counts].
static int sort_and_dedup(qb200_handle* h, int n_clouds, int dedup) {
  const int V = h->V;
  const dim3 g((V + 255) / 256, n_clouds);
  norm_key_kernel<<<g, 256, 0, h->stream>>>(h->desc_t, h->ctr.n_vox, V, h->key_a, h->val_a);
  h->launches += 1;
  int bits = 0;
  while ((1 << bits) < n_clouds) ++bits;
  int rc = launch_cloud_sort(h, n_clouds, h->ctr.n_vox, 32, 32);  // per-cloud shared-memory sort; device-wide radix sort for very large clouds
  if (rc == QB200_ERR_UNSUPPORTED) rc = sort_pairs(h, n_clouds * V, 32 + bits);
  if (rc) return rc;
  uint32_t* class_of = reinterpret_cast<uint32_t*>(h->key_a);
  int* n_unique = reinterpret_cast<int*>(class_of + (size_t)2 * h->S * V);
  dedup_kernel<<<n_clouds, 1024, 0, h->stream>>>(h->desc_t, h->ctr.n_vox, V, h->key_b, h->val_b, dedup, h->val_a, class_of, n_unique);
  h->launches += 1;
  QB_CUDA_TRY(h, cudaGetLastError());
  return QB200_OK;
}

int launch_match_nn(qb200_handle* h, int n_pairs) {
  const int V = h->V;
  const size_t smem = tc_smem_bytes();
  if (int rc = ensure_dyn_smem(h, (const void*)tc_nn_kernel<false>, smem)) return rc;
  // triage switches (results are identical either way): QB200_TC_NODEDUP=1 keeps duplicate descriptors, QB200_TC_NOPRUNE=1
  // visits every column tile
  static const int no_dedup = (getenv("QB200_TC_NODEDUP") && getenv("QB200_TC_NODEDUP")[0] == '1') ? 1 : 0;
  static const int no_prune = (getenv("QB200_TC_NOPRUNE") && getenv("QB200_TC_NOPRUNE")[0] == '1') ? 1 : 0;
  int rc = sort_and_dedup(h, 2 * n_pairs, no_dedup ? 0 : 1);
  if (rc) return rc;
  const uint32_t* uperm = h->val_a;
  const uint32_t* class_of = reinterpret_cast<const uint32_t*>(h->key_a);
  const int* n_unique = reinterpret_cast<const int*>(class_of + (size_t)2 * h->S * V);
  // class results, indexed by unique rank (colpart is scratch of the exact kernel, which runs later)
  unsigned long long* colbest_u = h->colpart;
  unsigned long long* rowbest_u = h->colpart + (size_t)h->S * V;
  unsigned* tile_cmax = reinterpret_cast<unsigned*>(h->colpart + (size_t)2 * h->S * V);  // [S][V/128][4] float bits, start above "+inf"
  QB_CUDA_TRY(h, cudaMemsetAsync(h->rowbest, 0xFF, (size_t)n_pairs * V * 8, h->stream));
  QB_CUDA_TRY(h, cudaMemsetAsync(h->colbest, 0xFF, (size_t)n_pairs * V * 8, h->stream));
  QB_CUDA_TRY(h, cudaMemsetAsync(h->colpart, 0xFF, ((size_t)2 * h->S * V + (size_t)h->S * (V >> 7) * 2 + 2) * 8, h->stream));  // 0xFFFFFFFF > +inf bits
  QB_CUDA_TRY(h, cudaMemsetAsync(h->tc_fallback, 0, (size_t)n_pairs * sizeof(int), h->stream));
  const dim3 gsplit((V + 255) / 256, 2 * n_pairs);
  split_desc_kernel<<<gsplit, 256, 0, h->stream>>>(h->desc_t, n_unique, V, uperm, h->desc_tiles, h->desc_norm);
  const dim3 g(h->NS, n_pairs);
  cudaEventRecord(h->kev[0], h->stream);
  static const int tc_prof = (getenv("QB200_TC_PROF") && getenv("QB200_TC_PROF")[0] == '1') ? 1 : 0;
  if (tc_prof) {
    if (int rc2 = ensure_dyn_smem(h, (const void*)tc_nn_kernel<false, true>, smem)) return rc2;
    tc_nn_kernel<false, true><<<g, kTcThreads, smem, h->stream>>>(h->desc_tiles, h->desc_norm, n_unique, V, uperm, rowbest_u, colbest_u,
                                                                  tile_cmax, h->tc_fallback, h->tc_stats, nullptr, no_prune);
  } else {
    tc_nn_kernel<false><<<g, kTcThreads, smem, h->stream>>>(h->desc_tiles, h->desc_norm, n_unique, V, uperm, rowbest_u, colbest_u,
                                                            tile_cmax, h->tc_fallback, h->tc_stats, nullptr, no_prune);
  }
  cudaEventRecord(h->kev[1], h->stream);
  h->kev_armed[0] = 1;
  broadcast_best_kernel<<<gsplit, 256, 0, h->stream>>>(rowbest_u, colbest_u, h->ctr.n_vox, V, h->val_b, class_of, h->rowbest, h->colbest);
  h->launches += 3;
  QB_CUDA_TRY(h, cudaGetLastError());
  return launch_match_exact(h, n_pairs, h->tc_fallback);
}

// debug/validation hook: approximate distances d~ of the first 128 x 128 tile of pair 0 (descriptors already in desc_t;
// duplicates are kept so that every (row, column) of the dump is filled)
int launch_tc_debug_tile(qb200_handle* h, float* d_out) {
  const size_t smem = tc_smem_bytes();
  if (int rc0 = ensure_dyn_smem(h, (const void*)tc_nn_kernel<true>, smem)) return rc0;
  int rc = sort_and_dedup(h, 2, 0);
  if (rc) return rc;
  const uint32_t* uperm = h->val_a;
  const int* n_unique = reinterpret_cast<const int*>(reinterpret_cast<const uint32_t*>(h->key_a) + (size_t)2 * h->S * h->V);
  QB_CUDA_TRY(h, cudaMemsetAsync(h->colpart, 0xFF, ((size_t)2 * h->S * h->V + (size_t)h->S * (h->V >> 7) * 2 + 2) * 8, h->stream));
  QB_CUDA_TRY(h, cudaMemsetAsync(h->tc_fallback, 0, sizeof(int), h->stream));
  const dim3 gsplit((h->V + 255) / 256, 2);
  split_desc_kernel<<<gsplit, 256, 0, h->stream>>>(h->desc_t, n_unique, h->V, uperm, h->desc_tiles, h->desc_norm);
  const dim3 g(1, 1);
  tc_nn_kernel<true><<<g, kTcThreads, smem, h->stream>>>(h->desc_tiles, h->desc_norm, n_unique, h->V, uperm, h->colpart + (size_t)h->S * h->V,
                                                         h->colpart, reinterpret_cast<unsigned*>(h->colpart + (size_t)2 * h->S * h->V), h->tc_fallback, h->tc_stats, d_out, 0);
  h->launches += 2;
  QB_CUDA_TRY(h, cudaGetLastError());
  return QB200_OK;
}

}  // namespace qb
