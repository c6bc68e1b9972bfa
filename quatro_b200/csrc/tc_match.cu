// tc_match.cu -- K6 on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
// The N_src x N_tgt x 33 descriptor-distance matrix is the one genuinely dense contraction of the
// path (north_star): d(i,j) = |a_i|^2 + |b_j|^2 - 2 a_i.b_j.  The reference does two exact 1-NN
// searches (FLANN kd-trees, src/teaser_utils/feature_matcher.cc:97-125); the exact answer here is
// defined by the fp32 fma chain of match.cu.  Tensor cores cannot reproduce that rounding, so they
// are used as a FILTER with a rigorous error bound and the winners are re-ranked exactly:
//
//   split_desc_kernel   x -> hi (top 19 bits = a TF32 value), lo = TF32(x - hi); |x|^2; per-cloud max norm
//   tc_match_kernel<0>  approximate row minima  m~_i = min_j d~(i,j)      (run for both orientations)
//   tc_match_kernel<1>  every (i,j) with d~(i,j) <= m~_i + margin_i or <= m~_j + margin_j is queued
//   rerank_kernel       exact fp32 chain distance of the queued pairs -> packed atomicMin into the
//                       row / column minima (distance bits << 32 | index  => lowest-index ties)
//
// d~ uses three TF32 MMAs per K block (hi.hi + hi.lo + lo.hi, fp32 accumulation in TMEM), i.e.
// |d~ - d| <= ~2.3e-5 (|a|^2 + |b|^2); margin = 1e-4 (|a_i|^2 + max_j |b_j|^2) therefore always keeps
// the exact arg-min in the queue (tests/test_gpu_parity.py::test_tc_filter_error_bound measures the
// slack).  If a queue overflows (thousands of near-identical descriptors) the pair is redone by the
// exact CUDA-core kernel of match.cu, so results never depend on the filter.
//
// Kernel anatomy (one CTA = 128 source rows, 2 CTAs per SM so one CTA's epilogue overlaps the other's
// loads and MMAs):
//   operands  : point-major TF32 hi/lo rows -> shared memory in the canonical K-major, no-swizzle UMMA layout
//               (8 points x 16 B core matrices) by 16-byte cp.async; fence.proxy.async
//   MMA       : one thread issues 15 x tcgen05.mma.cta_group::1.kind::tf32 (M=128, N=128, K=8) into a
//               128-column TMEM accumulator, tcgen05.commit -> mbarrier
//   epilogue  : 8 warps, tcgen05.ld.32x32b.x32 (lane = row), fused  nb_j - 2 dot  + min / threshold test
#include "handle.cuh"

namespace qb {

constexpr int kTcM = 128, kTcN = 128;
constexpr int kTcKB = kDescK / 8;                 // K blocks of 8 (TF32 MMA K)
constexpr int kTcTileBytes = kDescK * 128 * 4;    // one operand tile (128 points x 40 dims) = 20480 B
constexpr int kTcThreads = 256;
constexpr float kTcKappa = 1.0e-4f;               // margin = kappa * (|a_i|^2 + max_j |b_j|^2)
constexpr int kSpinLimit = 400000;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tc_cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(src));
}

// Canonical K-major / no-swizzle operand tile (validated by tools/tc_probe.cu on B200; MN-major TF32 without swizzle
// yields zeros): core matrix = 8 points x 16 B (4 consecutive K values); [kc = K/4][point group = p/8][p % 8][4 floats],
// i.e. byte offset kc*2048 + p*16.  The source is point-major (row p = 40 contiguous floats), so every 16-byte chunk
// is one cp.async and consecutive threads write consecutive shared-memory addresses.
__device__ __forceinline__ void tc_fill_tile(uint32_t dst_base, const float* __restrict__ src_rows /* [V][kDescK] */, int p0) {
  for (int t = threadIdx.x; t < 128 * (kDescK / 4); t += kTcThreads) {
    const int p = t & 127, kc = t >> 7;
    tc_cp_async16(dst_base + kc * 2048 + p * 16, src_rows + (size_t)(p0 + p) * kDescK + kc * 4);
  }
}

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

// squared norms, TF32 hi/lo split, per-cloud max norm
__global__ void __launch_bounds__(256) split_desc_kernel(const float* __restrict__ desc_t, const int* __restrict__ n_vox, int V,
                                                         float* __restrict__ hi, float* __restrict__ lo, float* __restrict__ norm,
                                                         unsigned* __restrict__ norm_max) {
  const int cloud = blockIdx.y;
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = n_vox[cloud];
  float acc = 0.0f;
  if (q < n) {
    const size_t base = (size_t)cloud * kDescK * V + q;
    float4* __restrict__ ho = reinterpret_cast<float4*>(hi + ((size_t)cloud * V + q) * kDescK);  // point-major rows of 40 floats
    float4* __restrict__ lw = reinterpret_cast<float4*>(lo + ((size_t)cloud * V + q) * kDescK);
#pragma unroll
    for (int c4 = 0; c4 < kDescK / 4; ++c4) {
      float hv[4], lv[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int d = 4 * c4 + e;
        const float x = d < kDescDim ? desc_t[base + (size_t)d * V] : 0.0f;
        hv[e] = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
        lv[e] = __uint_as_float(__float_as_uint(x - hv[e]) & 0xFFFFE000u);
        if (d < kDescDim) acc = __fmaf_rn(x, x, acc);
      }
      ho[c4] = make_float4(hv[0], hv[1], hv[2], hv[3]);
      lw[c4] = make_float4(lv[0], lv[1], lv[2], lv[3]);
    }
    norm[(size_t)cloud * V + q] = acc;
  }
  float m = (q < n && acc == acc) ? acc : 0.0f;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane_id() == 0 && m > 0.0f) atomicMax(norm_max + cloud, __float_as_uint(m));
}

// MODE 0: approx_min[cloudA][i] = min_j d~(i,j).      MODE 1: queue candidates (rows = source cloud).
// MODE 2: MODE 0 + dump of the first 128 x 128 tile of d~ (validation hook).
template <int MODE>
__global__ void __launch_bounds__(kTcThreads, 2)
tc_match_kernel(int swap, const float* __restrict__ hi, const float* __restrict__ lo, const float* __restrict__ norm,
                const unsigned* __restrict__ norm_max, const int* __restrict__ n_vox, int V, float* __restrict__ approx_min,
                unsigned* __restrict__ cand_q, int* __restrict__ cand_n, int qcap, float* __restrict__ dbg_tile) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t s_bar;
  __shared__ uint32_t s_tmem;
  __shared__ float s_nb[kTcN], s_cj[kTcN], s_part[kTcM];

  const int pair = blockIdx.y, stripe = blockIdx.x;
  const int cloudA = swap ? 2 * pair + 1 : 2 * pair, cloudB = swap ? 2 * pair : 2 * pair + 1;
  const int nA = n_vox[cloudA], nB = n_vox[cloudB];
  const int r0 = stripe * kTcM;
  if (r0 >= nA || nB <= 0) return;  // uniform for the CTA, before any barrier / allocation

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t sA_hi = smem_u32(smem), sA_lo = sA_hi + kTcTileBytes, sB_hi = sA_lo + kTcTileBytes, sB_lo = sB_hi + kTcTileBytes;
  const float* __restrict__ Ahi = hi + (size_t)cloudA * V * kDescK;  // point-major [V][40]
  const float* __restrict__ Alo = lo + (size_t)cloudA * V * kDescK;
  const float* __restrict__ Bhi = hi + (size_t)cloudB * V * kDescK;
  const float* __restrict__ Blo = lo + (size_t)cloudB * V * kDescK;
  const float* __restrict__ nA_ = norm + (size_t)cloudA * V;
  const float* __restrict__ nB_ = norm + (size_t)cloudB * V;

  if (warp == 0) {  // TMEM: 128 fp32 columns x 128 lanes
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&s_tmem)), "r"(kTcN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&s_bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = s_tmem;

  tc_fill_tile(sA_hi, Ahi, r0);
  tc_fill_tile(sA_lo, Alo, r0);

  // this thread's accumulator row and column half
  const int quad = warp & 3, chalf = warp >> 2;
  const int row = quad * 32 + lane, gi = r0 + row;
  const bool row_ok = gi < nA;
  const float na_i = row_ok ? nA_[gi] : 0.0f;
  const float nmaxA = __uint_as_float(norm_max[cloudA]), nmaxB = __uint_as_float(norm_max[cloudB]);
  float m = INFINITY;  // MODE 0: running min of nb_j - 2 dot
  float Ri = 0.0f;     // MODE 1: row threshold on nb_j - 2 dot
  if (MODE == 1) Ri = row_ok ? (approx_min[(size_t)cloudA * V + gi] + kTcKappa * (na_i + nmaxB)) - na_i : -INFINITY;
  const float negna = -na_i;

  // instruction descriptor: D=F32 (bits 4-5), A=B=TF32 (bits 7-9, 10-12), both K-major (bits 15,16 = 0), N>>3 (17-22), M>>4 (24-28)
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kTcN >> 3) << 17) | ((uint32_t)(kTcM >> 4) << 24);
  uint32_t parity = 0;
  bool dead = false;
  const int n_tiles = (nB + kTcN - 1) / kTcN;

  for (int jt = 0; jt < n_tiles; ++jt) {
    const int c0 = jt * kTcN;
    tc_fill_tile(sB_hi, Bhi, c0);
    tc_fill_tile(sB_lo, Blo, c0);
    if (threadIdx.x < kTcN) {
      const int j = c0 + threadIdx.x;
      const float nb = j < nB ? nB_[j] : INFINITY;  // +inf: padded columns never win and never qualify
      s_nb[threadIdx.x] = nb;
      if (MODE == 1) s_cj[threadIdx.x] = j < nB ? nb - (approx_min[(size_t)cloudB * V + j] + kTcKappa * (nb + nmaxA)) : INFINITY;
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");  // generic-proxy writes -> visible to the tensor core
    __syncthreads();

    if (threadIdx.x == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      uint32_t acc = 0;
#pragma unroll
      for (int kb = 0; kb < kTcKB; ++kb) {  // small cross terms first, then hi.hi
        tc_mma_tf32(tmem, tc_smem_desc(sA_hi + kb * 4096), tc_smem_desc(sB_lo + kb * 4096), idesc, acc);
        acc = 1;
        tc_mma_tf32(tmem, tc_smem_desc(sA_lo + kb * 4096), tc_smem_desc(sB_hi + kb * 4096), idesc, 1);
      }
#pragma unroll
      for (int kb = 0; kb < kTcKB; ++kb) tc_mma_tf32(tmem, tc_smem_desc(sA_hi + kb * 4096), tc_smem_desc(sB_hi + kb * 4096), idesc, 1);
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&s_bar)) : "memory");
    }
    // wait for the accumulator (bounded spin: a descriptor bug must not hang the box)
    {
      int spin = 0;
      uint32_t ok = 0;
      while (!ok) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(ok)
            : "r"(smem_u32(&s_bar)), "r"(parity)
            : "memory");
        if (!ok && ++spin > kSpinLimit) { dead = true; break; }
      }
    }
    parity ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");

    if (!dead) {
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        const int cb = chalf * 64 + ch * 32;
        uint32_t v[32];
        tc_ld32(tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)cb, v);
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const float dot = __uint_as_float(v[c]);
          const float t = fmaf(-2.0f, dot, s_nb[cb + c]);
          if (MODE != 1) {
            m = fminf(m, t);
            if (MODE == 2 && stripe == 0 && jt == 0) dbg_tile[(size_t)row * kTcN + cb + c] = na_i + t;
          } else {
            const float u = fmaf(-2.0f, dot, s_cj[cb + c]);
            if (row_ok && ((t <= Ri) || (u <= negna))) {
              const int slot = atomicAdd(cand_n + pair, 1);
              if (slot < qcap) cand_q[(size_t)pair * qcap + slot] = ((unsigned)gi << 16) | (unsigned)(c0 + cb + c);
            }
          }
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();  // TMEM and the B tiles are free again
  }

  if (MODE != 1) {
    if (chalf == 1) s_part[row] = m;
    __syncthreads();
    // a dead wait (should never happen) yields +inf: everything qualifies, the queue overflows, the exact kernel takes over
    if (chalf == 0 && row_ok) approx_min[(size_t)cloudA * V + gi] = dead ? INFINITY : fminf(m, s_part[row]) + na_i;
  } else if (dead && threadIdx.x == 0) {
    atomicAdd(cand_n + pair, qcap + 1);
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(kTcN) : "memory");
}

__device__ __forceinline__ unsigned long long tc_pack(float d, int idx) {
  return ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)idx;
}

// exact canonical distance of every queued (src, tgt) pair; folds both directions
__global__ void __launch_bounds__(256) rerank_kernel(const float* __restrict__ desc_t, const int* __restrict__ n_vox, int V,
                                                     const unsigned* __restrict__ cand_q, const int* __restrict__ cand_n, int qcap,
                                                     int* __restrict__ fallback, unsigned long long* __restrict__ rowbest,
                                                     unsigned long long* __restrict__ colbest) {
  const int pair = blockIdx.y;
  const int n = cand_n[pair];
  if (n > qcap) {  // overflow: the exact CUDA-core kernel redoes this pair
    if (blockIdx.x == 0 && threadIdx.x == 0) fallback[pair] = 1;
    return;
  }
  const float* __restrict__ A = desc_t + (size_t)(2 * pair) * kDescK * V;
  const float* __restrict__ B = desc_t + (size_t)(2 * pair + 1) * kDescK * V;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n; c += gridDim.x * blockDim.x) {
    const unsigned e = cand_q[(size_t)pair * qcap + c];
    const int i = (int)(e >> 16), j = (int)(e & 0xFFFFu);
    float acc = 0.0f;
#pragma unroll
    for (int d = 0; d < kDescDim; ++d) {
      const float diff = A[(size_t)d * V + i] - B[(size_t)d * V + j];
      acc = __fmaf_rn(diff, diff, acc);
    }
    if (acc == acc) {
      atomicMin(rowbest + (size_t)pair * V + i, tc_pack(acc, j));
      atomicMin(colbest + (size_t)pair * V + j, tc_pack(acc, i));
    }
  }
}

// implemented in match.cu: exact kernels restricted to the pairs flagged in `only`
int launch_match_exact(qb200_handle* h, int n_pairs, const int* only);

int launch_match_nn(qb200_handle* h, int n_pairs) {
  const int V = h->V;
  static bool attr_set = false;
  const size_t smem = 4 * (size_t)kTcTileBytes + 1024;
  if (!attr_set) {
    QB_CUDA_TRY(h, cudaFuncSetAttribute(tc_match_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    QB_CUDA_TRY(h, cudaFuncSetAttribute(tc_match_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  QB_CUDA_TRY(h, cudaMemsetAsync(h->rowbest, 0xFF, (size_t)n_pairs * V * 8, h->stream));
  QB_CUDA_TRY(h, cudaMemsetAsync(h->colbest, 0xFF, (size_t)n_pairs * V * 8, h->stream));
  QB_CUDA_TRY(h, cudaMemsetAsync(h->cand_n, 0, (size_t)n_pairs * sizeof(int), h->stream));
  QB_CUDA_TRY(h, cudaMemsetAsync(h->tc_fallback, 0, (size_t)n_pairs * sizeof(int), h->stream));
  QB_CUDA_TRY(h, cudaMemsetAsync(h->norm_max, 0, (size_t)2 * n_pairs * sizeof(unsigned), h->stream));
  const dim3 gsplit((V + 255) / 256, 2 * n_pairs);
  split_desc_kernel<<<gsplit, 256, 0, h->stream>>>(h->desc_t, h->ctr.n_vox, V, h->desc_hi, h->desc_lo, h->desc_norm, h->norm_max);
  const dim3 g(h->NS, n_pairs);
  cudaEventRecord(h->kev[0], h->stream);
  tc_match_kernel<0><<<g, kTcThreads, smem, h->stream>>>(0, h->desc_hi, h->desc_lo, h->desc_norm, h->norm_max, h->ctr.n_vox, V, h->approx_min,
                                                         nullptr, nullptr, 0, nullptr);
  tc_match_kernel<0><<<g, kTcThreads, smem, h->stream>>>(1, h->desc_hi, h->desc_lo, h->desc_norm, h->norm_max, h->ctr.n_vox, V, h->approx_min,
                                                         nullptr, nullptr, 0, nullptr);
  tc_match_kernel<1><<<g, kTcThreads, smem, h->stream>>>(0, h->desc_hi, h->desc_lo, h->desc_norm, h->norm_max, h->ctr.n_vox, V, h->approx_min,
                                                         h->cand_q, h->cand_n, h->qcap, nullptr);
  cudaEventRecord(h->kev[1], h->stream);
  h->kev_armed[0] = 1;
  const dim3 gr(64, n_pairs);
  rerank_kernel<<<gr, 256, 0, h->stream>>>(h->desc_t, h->ctr.n_vox, V, h->cand_q, h->cand_n, h->qcap, h->tc_fallback, h->rowbest, h->colbest);
  h->launches += 5;
  QB_CUDA_TRY(h, cudaGetLastError());
  return launch_match_exact(h, n_pairs, h->tc_fallback);
}

// debug/validation hook: approximate distances of the first 128 x 128 tile of pair 0 (after launch_match_nn inputs are in place)
int launch_tc_debug_tile(qb200_handle* h, float* d_out) {
  const size_t smem = 4 * (size_t)kTcTileBytes + 1024;
  QB_CUDA_TRY(h, cudaFuncSetAttribute(tc_match_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  QB_CUDA_TRY(h, cudaMemsetAsync(h->norm_max, 0, 2 * sizeof(unsigned), h->stream));
  const dim3 gsplit((h->V + 255) / 256, 2);
  split_desc_kernel<<<gsplit, 256, 0, h->stream>>>(h->desc_t, h->ctr.n_vox, h->V, h->desc_hi, h->desc_lo, h->desc_norm, h->norm_max);
  const dim3 g(1, 1);
  tc_match_kernel<2><<<g, kTcThreads, smem, h->stream>>>(0, h->desc_hi, h->desc_lo, h->desc_norm, h->norm_max, h->ctr.n_vox, h->V, h->approx_min,
                                                         nullptr, nullptr, 0, d_out);
  h->launches += 2;
  QB_CUDA_TRY(h, cudaGetLastError());
  return QB200_OK;
}

}  // namespace qb
