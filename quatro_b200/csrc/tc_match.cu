// tc_match.cu -- K6 on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
// The N_src x N_tgt x 33 descriptor-distance matrix is the one genuinely dense contraction of the
// path (north_star): d(i,j) = |a_i|^2 + |b_j|^2 - 2 a_i.b_j.  The reference does two exact 1-NN
// searches (FLANN kd-trees, src/teaser_utils/feature_matcher.cc:97-125); the exact answer here is
// defined by the fp32 fma chain of match.cu.  Tensor cores cannot reproduce that rounding, so they
// are used as a FILTER with a rigorous error bound and the winners are re-ranked exactly:
//
//   split_desc_kernel   x -> hi (top 19 bits = a TF32 value), lo = TF32(x - hi); |x|^2; per-cloud max norm
//   tc_match_kernel<0>  upper bounds  U_i = min over every 2nd column tile of d~(i,j)   (both orientations)
//   tc_match_kernel<1>  every (i,j) with d~(i,j) <= U_i + margin_i or <= U_j + margin_j is queued
//   rerank_kernel       exact fp32 chain distance of the queued pairs -> packed atomicMin into the
//                       row / column minima (distance bits << 32 | index  => lowest-index ties)
//
// d~ uses three TF32 MMAs per K block (hi.hi + hi.lo + lo.hi, fp32 accumulation in TMEM), i.e.
// |d~ - d| <= ~2.3e-5 (|a|^2 + |b|^2); margin = 1e-4 (|a_i|^2 + max_j |b_j|^2) therefore always keeps
// the exact arg-min in the queue (tests/test_gpu_parity.py::test_tc_filter_error_bound measures the
// slack).  If a queue overflows (thousands of near-identical descriptors) the pair is redone by the
// exact CUDA-core kernel of match.cu, so results never depend on the filter.
//
// Kernel anatomy (one CTA = 128 source rows, one CTA per SM):
//   operands  : 128-point blocks are stored in global memory as ready-made shared-memory images (canonical K-major,
//               no-swizzle UMMA layout), so a tile is ONE cp.async.bulk (TMA bulk copy) completing on an mbarrier;
//               B tiles stream through a 2-stage ring
//   MMA       : one thread issues 15 x tcgen05.mma.cta_group::1.kind::tf32 (M=128, N=128, K=8) into one of two
//               128-column TMEM accumulator stages, tcgen05.commit -> mbarrier
//   epilogue  : 8 warps drain the OTHER TMEM stage meanwhile: tcgen05.ld.32x32b.x32 (lane = row), fused
//               nb_j - 2 dot  + min / threshold test
#include "handle.cuh"

namespace qb {

constexpr int kTcM = 128, kTcN = 128;
constexpr int kTcKB = kDescK / 8;                 // K blocks of 8 (TF32 MMA K)
constexpr int kTcTileBytes = kDescK * 128 * 4;    // one operand tile (128 points x 40 dims) = 20480 B
constexpr int kTcThreads = 256;
constexpr float kTcKappa = 1.0e-4f;               // margin = kappa * (|a_i|^2 + max_j |b_j|^2)
constexpr int kSpinLimit = 400000;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Operand tiles use the canonical K-major / no-swizzle UMMA layout (validated by tools/tc_probe.cu on B200; MN-major
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

// ------------------------------------------------------------------------------------------------
// split_desc_kernel: per point  |x|^2 (fp32 fma chain), per-cloud max norm, and two re-layouts of the descriptor:
//   rows  [cloud][V][40]            exact fp32, point-major (the re-rank reads 160 contiguous bytes per point)
//   tiles [cloud][V/128][2][5120]   TF32 hi / lo images of each 128-point block, stored EXACTLY in the shared-memory
//                                   operand layout (float index kc*512 + p*4 + e), so a tile is one 20/40 KB bulk copy
// ------------------------------------------------------------------------------------------------
constexpr int kTileFloats = kDescK * 128;  // 5120

__global__ void __launch_bounds__(256) split_desc_kernel(const float* __restrict__ desc_t, const int* __restrict__ n_vox, int V,
                                                         float* __restrict__ tiles, float* __restrict__ rows, float* __restrict__ norm,
                                                         unsigned* __restrict__ norm_max) {
  const int cloud = blockIdx.y;
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = n_vox[cloud];
  const int NB = V >> 7;
  float acc = 0.0f;
  if (q < ((n + 127) & ~127)) {  // pad the last block with zeros so stale data never reaches the tensor core
    const size_t base = (size_t)cloud * kDescK * V + q;
    const int blk = q >> 7, p = q & 127;
    float4* __restrict__ th = reinterpret_cast<float4*>(tiles + ((size_t)(cloud * NB + blk) * 2 + 0) * kTileFloats) + p;
    float4* __restrict__ tl = reinterpret_cast<float4*>(tiles + ((size_t)(cloud * NB + blk) * 2 + 1) * kTileFloats) + p;
    float4* __restrict__ rw = reinterpret_cast<float4*>(rows + ((size_t)cloud * V + q) * kDescK);
#pragma unroll
    for (int kc = 0; kc < kDescK / 4; ++kc) {
      float xv[4], hv[4], lv[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int d = 4 * kc + e;
        const float x = (d < kDescDim && q < n) ? desc_t[base + (size_t)d * V] : 0.0f;
        xv[e] = x;
        hv[e] = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
        lv[e] = __uint_as_float(__float_as_uint(x - hv[e]) & 0xFFFFE000u);
        if (d < kDescDim) acc = __fmaf_rn(x, x, acc);
      }
      th[kc * 128] = make_float4(hv[0], hv[1], hv[2], hv[3]);
      tl[kc * 128] = make_float4(lv[0], lv[1], lv[2], lv[3]);
      rw[kc] = make_float4(xv[0], xv[1], xv[2], xv[3]);
    }
    if (q < n) norm[(size_t)cloud * V + q] = acc;
  }
  float m = (q < n && acc == acc) ? acc : 0.0f;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane_id() == 0 && m > 0.0f) atomicMax(norm_max + cloud, __float_as_uint(m));
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

// MODE 0: approx_min[cloudA][i] = min over the SAMPLED column tiles of d~(i,j)  (an upper bound of the row minimum).
// MODE 1: queue every (i,j) with d~ <= bound_i + margin_i or <= bound_j + margin_j   (rows = source cloud, all tiles).
// MODE 2: MODE 0 + dump of the first 128 x 128 tile of d~ (validation hook).
// Pipeline per CTA (128 rows): B tiles arrive by bulk copy into a 2-stage ring; MMA(k) into TMEM stage k&1 runs
// while all 8 warps drain TMEM stage (k-1)&1.
constexpr int kTcStages = 3;      // B-tile ring (prefetch distance 2)
constexpr int kTcQueue = 8192;    // per-CTA shared-memory candidate queue (entries)

template <int MODE>
__global__ void __launch_bounds__(kTcThreads, 1)
tc_match_kernel(int swap, int tile_step, const float* __restrict__ tiles, const float* __restrict__ norm, const unsigned* __restrict__ norm_max,
                const int* __restrict__ n_vox, int V, float* __restrict__ approx_min, unsigned* __restrict__ cand_q,
                int* __restrict__ cand_n, int qcap, float* __restrict__ dbg_tile) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t s_full[kTcStages], s_mma[2], s_afull;
  __shared__ uint32_t s_tmem;
  __shared__ float s_nb[kTcStages][kTcN], s_cj[kTcStages][kTcN], s_part[kTcM];
  __shared__ int s_dead, s_qn, s_qbase;

  const int pair = blockIdx.y, stripe = blockIdx.x;
  const int cloudA = swap ? 2 * pair + 1 : 2 * pair, cloudB = swap ? 2 * pair : 2 * pair + 1;
  const int nA = n_vox[cloudA], nB = n_vox[cloudB];
  const int r0 = stripe * kTcM;
  if (r0 >= nA || nB <= 0) return;  // uniform for the CTA, before any barrier / allocation

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int NB = V >> 7;
  constexpr uint32_t kPairBytes = 2 * kTcTileBytes;  // hi + lo image of one 128-point block
  const uint32_t sA = smem_u32(smem), sB0 = sA + kPairBytes;
  unsigned* s_q = reinterpret_cast<unsigned*>(smem + (1 + kTcStages) * kPairBytes);  // MODE 1 only
  const float* __restrict__ tA = tiles + (size_t)cloudA * NB * 2 * kTileFloats;
  const float* __restrict__ tB = tiles + (size_t)cloudB * NB * 2 * kTileFloats;
  const float* __restrict__ nA_ = norm + (size_t)cloudA * V;
  const float* __restrict__ nB_ = norm + (size_t)cloudB * V;
  const uint32_t bar_full0 = smem_u32(&s_full[0]), bar_mma0 = smem_u32(&s_mma[0]), bar_a = smem_u32(&s_afull);

  const int n_tiles = (nB + kTcN - 1) / kTcN;
  const int step = tile_step < n_tiles ? tile_step : n_tiles;      // sampled passes visit every step-th tile
  const int first = step > 1 ? (stripe % step) : 0;
  const int ntl = (n_tiles - first + step - 1) / step;             // >= 1

  if (warp == 0) {  // TMEM: 2 accumulator stages x 128 fp32 columns
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&s_tmem)), "r"(2 * kTcN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < kTcStages; ++i) mbar_init(bar_full0 + 8 * i, 1);
    mbar_init(bar_mma0, 1); mbar_init(bar_mma0 + 8, 1); mbar_init(bar_a, 1);
    s_dead = 0; s_qn = 0;
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = s_tmem;

  // per-tile column data (norms, and in MODE 1 the column thresholds) for ring stage st
  const float nmaxA = __uint_as_float(norm_max[cloudA]), nmaxB = __uint_as_float(norm_max[cloudB]);
  auto load_cols = [&](int st, int jt) {
    if (threadIdx.x < kTcN) {
      const int j = jt * kTcN + threadIdx.x;
      const float nb = j < nB ? nB_[j] : INFINITY;  // +inf: padded columns never win and never qualify
      s_nb[st][threadIdx.x] = nb;
      if (MODE == 1) s_cj[st][threadIdx.x] = j < nB ? nb - (approx_min[(size_t)cloudB * V + j] + kTcKappa * (nb + nmaxA)) : INFINITY;
    }
  };
  auto issue_tile = [&](int st, int jt) {  // one thread: 40 KB bulk copy of block jt of cloud B into ring stage st
    mbar_expect_tx(bar_full0 + 8 * st, kPairBytes);
    bulk_g2s(sB0 + st * kPairBytes, tB + (size_t)jt * 2 * kTileFloats, kPairBytes, bar_full0 + 8 * st);
  };
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar_a, kPairBytes);
    bulk_g2s(sA, tA + (size_t)stripe * 2 * kTileFloats, kPairBytes, bar_a);
    issue_tile(0, first);
    if (ntl > 1) issue_tile(1, first + step);
  }
  load_cols(0, first);
  if (ntl > 1) load_cols(1, first + step);

  // this thread's accumulator row and column half
  const int quad = warp & 3, chalf = warp >> 2;
  const int row = quad * 32 + lane, gi = r0 + row;
  const bool row_ok = gi < nA;
  const float na_i = row_ok ? nA_[gi] : 0.0f;
  float m = INFINITY;  // MODE 0/2: running min of nb_j - 2 dot
  float Ri = 0.0f;     // MODE 1: row threshold on nb_j - 2 dot
  if (MODE == 1) Ri = row_ok ? (approx_min[(size_t)cloudA * V + gi] + kTcKappa * (na_i + nmaxB)) - na_i : -INFINITY;
  const float negna = row_ok ? -na_i : -INFINITY;
  // instruction descriptor: D=F32 (bits 4-5), A=B=TF32 (bits 7-9, 10-12), both K-major (bits 15,16 = 0), N>>3 (17-22), M>>4 (24-28)
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kTcN >> 3) << 17) | ((uint32_t)(kTcM >> 4) << 24);
  __syncthreads();  // s_nb / s_cj of the first two tiles visible

  auto epilogue = [&](int k) {  // drain TMEM stage k&1 (tile index first + k*step, ring stage k%3)
    const int ts = k & 1, st = k % kTcStages, jt = first + k * step;
    if (!mbar_wait(bar_mma0 + 8 * ts, (uint32_t)((k >> 1) & 1))) s_dead = 1;
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const int c0 = jt * kTcN;
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
      const int cb = chalf * 64 + ch * 32;
      uint32_t v[32];
      tc_ld32(tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(ts * kTcN + cb), v);
      if (MODE != 1) {
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const float t = fmaf(-2.0f, __uint_as_float(v[c]), s_nb[st][cb + c]);
          m = fminf(m, t);
          if (MODE == 2 && stripe == 0 && jt == 0) dbg_tile[(size_t)row * kTcN + cb + c] = na_i + t;
        }
      } else {
        // branch-free candidate mask of this lane's row over the 32 columns, then one warp-aggregated queue reservation
        uint32_t mask = 0;
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const float dot = __uint_as_float(v[c]);
          const float t = fmaf(-2.0f, dot, s_nb[st][cb + c]);
          const float u = fmaf(-2.0f, dot, s_cj[st][cb + c]);
          mask |= ((t <= Ri) || (u <= negna) ? 1u : 0u) << c;
        }
        if (__any_sync(0xffffffffu, mask != 0)) {
          int tot;
          int off = warp_excl_scan(__popc(mask), &tot);
          int base = 0;
          if (lane == 0) base = atomicAdd(&s_qn, tot);
          off += __shfl_sync(0xffffffffu, base, 0);
          while (mask) {
            const int c = __ffs(mask) - 1;
            mask &= mask - 1;
            const unsigned e = ((unsigned)gi << 16) | (unsigned)(c0 + cb + c);
            if (off < kTcQueue) s_q[off] = e;
            else {  // shared queue full (pathological ties): straight to the global queue
              const int slot = atomicAdd(cand_n + pair, 1);
              if (slot < qcap) cand_q[(size_t)pair * qcap + slot] = e;
            }
            ++off;
          }
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  };

  for (int k = 0; k < ntl; ++k) {
    const int ts = k & 1, st = k % kTcStages;
    if (threadIdx.x == 0) {
      bool ok = true;
      if (k == 0) ok = mbar_wait(bar_a, 0);
      ok = ok && mbar_wait(bar_full0 + 8 * st, (uint32_t)((k / kTcStages) & 1));
      if (!ok) s_dead = 1;
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      const uint32_t aH = sA, aL = sA + kTcTileBytes, bH = sB0 + st * kPairBytes, bL = bH + kTcTileBytes, d = tmem + (uint32_t)(ts * kTcN);
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
    if (k >= 1) epilogue(k - 1);  // overlaps the MMAs just issued
    __syncthreads();              // TMEM stage ts^1 drained; MMA k-1 complete => ring stage (k-1)%3 and its column data are free
    if (k + 2 < ntl) {            // prefetch distance 2: tile k+2 goes into the stage tile k-1 just left
      const int st2 = (k + 2) % kTcStages;
      if (threadIdx.x == 0) issue_tile(st2, first + (k + 2) * step);
      load_cols(st2, first + (k + 2) * step);  // read again only after the next barriers
    }
    if (s_dead) break;  // uniform: written before the barrier above
  }
  if (!s_dead) epilogue(ntl - 1);
  __syncthreads();
  const bool dead = s_dead != 0;

  if (MODE != 1) {
    if (chalf == 1) s_part[row] = m;
    __syncthreads();
    // a dead wait (should never happen) yields +inf: everything qualifies, the queue overflows, the exact kernel takes over
    if (chalf == 0 && row_ok) approx_min[(size_t)cloudA * V + gi] = dead ? INFINITY : fminf(m, s_part[row]) + na_i;
  } else {
    // flush the shared-memory candidate queue with ONE global reservation
    const int nq = s_qn < kTcQueue ? s_qn : kTcQueue;
    if (threadIdx.x == 0) s_qbase = atomicAdd(cand_n + pair, dead ? qcap + 1 : nq);
    __syncthreads();
    const int gbase = s_qbase;
    for (int t = threadIdx.x; t < nq; t += kTcThreads)
      if (gbase + t < qcap) cand_q[(size_t)pair * qcap + gbase + t] = s_q[t];
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(2 * kTcN) : "memory");
}

__device__ __forceinline__ unsigned long long tc_pack(float d, int idx) {
  return ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)idx;
}

// exact canonical distance of every queued (src, tgt) pair; folds both directions
__global__ void __launch_bounds__(256) rerank_kernel(const float* __restrict__ rows, const int* __restrict__ n_vox, int V,
                                                     const unsigned* __restrict__ cand_q, const int* __restrict__ cand_n, int qcap,
                                                     int* __restrict__ fallback, unsigned long long* __restrict__ rowbest,
                                                     unsigned long long* __restrict__ colbest) {
  const int pair = blockIdx.y;
  const int n = cand_n[pair];
  if (n > qcap) {  // overflow: the exact CUDA-core kernel redoes this pair
    if (blockIdx.x == 0 && threadIdx.x == 0) fallback[pair] = 1;
    return;
  }
  const float4* __restrict__ A = reinterpret_cast<const float4*>(rows + (size_t)(2 * pair) * V * kDescK);
  const float4* __restrict__ B = reinterpret_cast<const float4*>(rows + (size_t)(2 * pair + 1) * V * kDescK);
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n; c += gridDim.x * blockDim.x) {
    const unsigned e = cand_q[(size_t)pair * qcap + c];
    const int i = (int)(e >> 16), j = (int)(e & 0xFFFFu);
    float acc = 0.0f;
#pragma unroll
    for (int c4 = 0; c4 < (kDescDim + 3) / 4; ++c4) {  // same d = 0..32 fma chain as the exact kernel
      const float4 a = __ldg(A + (size_t)i * (kDescK / 4) + c4), b = __ldg(B + (size_t)j * (kDescK / 4) + c4);
      float diff = a.x - b.x;
      acc = __fmaf_rn(diff, diff, acc);
      if (4 * c4 + 1 < kDescDim) { diff = a.y - b.y; acc = __fmaf_rn(diff, diff, acc); }
      if (4 * c4 + 2 < kDescDim) { diff = a.z - b.z; acc = __fmaf_rn(diff, diff, acc); }
      if (4 * c4 + 3 < kDescDim) { diff = a.w - b.w; acc = __fmaf_rn(diff, diff, acc); }
    }
    if (acc == acc) {
      atomicMin(rowbest + (size_t)pair * V + i, tc_pack(acc, j));
      atomicMin(colbest + (size_t)pair * V + j, tc_pack(acc, i));
    }
  }
}

// implemented in match.cu: exact kernels restricted to the pairs flagged in `only`
int launch_match_exact(qb200_handle* h, int n_pairs, const int* only);

constexpr int kTcSampleStep = 2;  // the bound passes visit every 2nd column tile

int launch_match_nn(qb200_handle* h, int n_pairs) {
  const int V = h->V;
  static bool attr_set = false;
  const size_t smem = (size_t)(1 + kTcStages) * (2 * kTcTileBytes) + (size_t)kTcQueue * 4 + 1024;  // A + B ring (hi+lo images) + queue
  if (!attr_set) {
    QB_CUDA_TRY(h, cudaFuncSetAttribute(tc_match_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    QB_CUDA_TRY(h, cudaFuncSetAttribute(tc_match_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    QB_CUDA_TRY(h, cudaFuncSetAttribute(tc_match_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  QB_CUDA_TRY(h, cudaMemsetAsync(h->rowbest, 0xFF, (size_t)n_pairs * V * 8, h->stream));
  QB_CUDA_TRY(h, cudaMemsetAsync(h->colbest, 0xFF, (size_t)n_pairs * V * 8, h->stream));
  QB_CUDA_TRY(h, cudaMemsetAsync(h->cand_n, 0, (size_t)n_pairs * sizeof(int), h->stream));
  QB_CUDA_TRY(h, cudaMemsetAsync(h->tc_fallback, 0, (size_t)n_pairs * sizeof(int), h->stream));
  QB_CUDA_TRY(h, cudaMemsetAsync(h->norm_max, 0, (size_t)2 * n_pairs * sizeof(unsigned), h->stream));
  const dim3 gsplit((V + 255) / 256, 2 * n_pairs);
  split_desc_kernel<<<gsplit, 256, 0, h->stream>>>(h->desc_t, h->ctr.n_vox, V, h->desc_tiles, h->desc_rows, h->desc_norm, h->norm_max);
  const dim3 g(h->NS, n_pairs);
  cudaEventRecord(h->kev[0], h->stream);
  tc_match_kernel<0><<<g, kTcThreads, smem, h->stream>>>(0, kTcSampleStep, h->desc_tiles, h->desc_norm, h->norm_max, h->ctr.n_vox, V, h->approx_min,
                                                         nullptr, nullptr, 0, nullptr);
  tc_match_kernel<0><<<g, kTcThreads, smem, h->stream>>>(1, kTcSampleStep, h->desc_tiles, h->desc_norm, h->norm_max, h->ctr.n_vox, V, h->approx_min,
                                                         nullptr, nullptr, 0, nullptr);
  tc_match_kernel<1><<<g, kTcThreads, smem, h->stream>>>(0, 1, h->desc_tiles, h->desc_norm, h->norm_max, h->ctr.n_vox, V, h->approx_min, h->cand_q,
                                                         h->cand_n, h->qcap, nullptr);
  cudaEventRecord(h->kev[1], h->stream);
  h->kev_armed[0] = 1;
  const dim3 gr(64, n_pairs);
  rerank_kernel<<<gr, 256, 0, h->stream>>>(h->desc_rows, h->ctr.n_vox, V, h->cand_q, h->cand_n, h->qcap, h->tc_fallback, h->rowbest, h->colbest);
  h->launches += 5;
  QB_CUDA_TRY(h, cudaGetLastError());
  return launch_match_exact(h, n_pairs, h->tc_fallback);
}

// debug/validation hook: approximate distances of the first 128 x 128 tile of pair 0 (descriptors already in desc_t)
int launch_tc_debug_tile(qb200_handle* h, float* d_out) {
  const size_t smem = (size_t)(1 + kTcStages) * (2 * kTcTileBytes) + (size_t)kTcQueue * 4 + 1024;
  QB_CUDA_TRY(h, cudaFuncSetAttribute(tc_match_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  QB_CUDA_TRY(h, cudaMemsetAsync(h->norm_max, 0, 2 * sizeof(unsigned), h->stream));
  const dim3 gsplit((h->V + 255) / 256, 2);
  split_desc_kernel<<<gsplit, 256, 0, h->stream>>>(h->desc_t, h->ctr.n_vox, h->V, h->desc_tiles, h->desc_rows, h->desc_norm, h->norm_max);
  const dim3 g(1, 1);
  tc_match_kernel<2><<<g, kTcThreads, smem, h->stream>>>(0, 1, h->desc_tiles, h->desc_norm, h->norm_max, h->ctr.n_vox, h->V, h->approx_min, nullptr,
                                                         nullptr, 0, d_out);
  h->launches += 2;
  QB_CUDA_TRY(h, cudaGetLastError());
  return QB200_OK;
}

}  // namespace qb
