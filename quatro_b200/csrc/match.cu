// match.cu -- K6 (33-D all-pairs nearest neighbour, both directions) + K7 (mutual check, tuple
// test, dedupe/sort, packing of the matched point pairs).   sm_100a
//
// Replaces Matcher::calculateCorrespondences / normalizePoints / advancedMatching
// (include/teaser_utils/feature_matcher.h:42-74, src/teaser_utils/feature_matcher.cc:18-265, which
// builds two FLANN kd-trees) and the packing loop of FPFHManager::setFeaturePair
// (include/fpfh_manager.hpp:130-152).
//
// The two exact 1-NN searches are the row- and column-argmin of one N_src x N_tgt distance matrix
// that is never materialised: a CTA owns a 128-row stripe of source descriptors (staged once in
// shared memory, dimension-major), streams 128-column target tiles through a cp.async double
// buffer, keeps row minima in registers and emits per-stripe column minima that a second kernel
// folds.  Distances are the fused-multiply-add chain over d = 0..32 of (a_d - b_d)^2 and minima
// carry the candidate index in the low word, so ties resolve to the lowest index -- exactly the
// CPU oracle's arithmetic, hence bit-identical argmins.
#include <stdlib.h>

#include "handle.cuh"

namespace qb {

constexpr int kMT = kMatchTile;      // 128 x 128 tile
constexpr int kMatchThreads = 256;   // 16 x 16 threads, 8 x 8 distances each

__device__ __forceinline__ unsigned long long pack_dist(float d, int idx) {
  return ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)idx;
}
__device__ __forceinline__ unsigned long long umin64(unsigned long long a, unsigned long long b) { return a < b ? a : b; }
__device__ __forceinline__ unsigned long long shfl_xor_u64(unsigned long long v, int m) {
  const unsigned lo = __shfl_xor_sync(0xffffffffu, (unsigned)v, m), hi = __shfl_xor_sync(0xffffffffu, (unsigned)(v >> 32), m);
  return ((unsigned long long)hi << 32) | lo;
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// load rows d = 0..32 of a 128-wide column window [c0, c0+128) of a dimension-major matrix
__device__ __forceinline__ void load_tile_async(float (*dst)[kMT], const float* __restrict__ src, int V, int c0) {
  for (int ch = threadIdx.x; ch < kDescDim * (kMT / 4); ch += kMatchThreads) {
    const int d = ch / (kMT / 4), c4 = (ch % (kMT / 4)) * 4;
    cp_async16(&dst[d][c4], src + (size_t)d * V + c0 + c4);
  }
}

__global__ void __launch_bounds__(kMatchThreads, 2)
match_stripe_kernel(const float* __restrict__ desc_t, const int* __restrict__ n_vox, int V, int NS, const int* __restrict__ only,
                    unsigned long long* __restrict__ rowbest, unsigned long long* __restrict__ colpart) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float(*As)[kMT] = reinterpret_cast<float(*)[kMT]>(smem_raw);                                    // [33][128]
  float(*Bs0)[kMT] = reinterpret_cast<float(*)[kMT]>(smem_raw + sizeof(float) * kDescDim * kMT);  // [33][128]
  float(*Bs1)[kMT] = reinterpret_cast<float(*)[kMT]>(smem_raw + 2 * sizeof(float) * kDescDim * kMT);
  unsigned long long(*colred)[kMT] =
      reinterpret_cast<unsigned long long(*)[kMT]>(smem_raw + 3 * sizeof(float) * kDescDim * kMT);  // [8][128]

  const int pair = blockIdx.y, stripe = blockIdx.x;
  if (only != nullptr && only[pair] == 0) return;  // fallback mode: only the pairs whose tensor-core queue overflowed
  const int nA = n_vox[2 * pair], nB = n_vox[2 * pair + 1];
  const int r0 = stripe * kMT;
  if (r0 >= nA || nB <= 0) return;
  const float* __restrict__ A = desc_t + (size_t)(2 * pair) * kDescK * V;
  const float* __restrict__ B = desc_t + (size_t)(2 * pair + 1) * kDescK * V;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4, warp = threadIdx.x >> 5;

  load_tile_async(As, A, V, r0);
  load_tile_async(Bs0, B, V, 0);
  cp_async_commit();

  unsigned long long rbest[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) rbest[r] = ~0ull;

  const int n_tiles = (nB + kMT - 1) / kMT;
  for (int jt = 0; jt < n_tiles; ++jt) {
    float(*Bs)[kMT] = (jt & 1) ? Bs1 : Bs0;
    if (jt + 1 < n_tiles) {
      load_tile_async((jt & 1) ? Bs0 : Bs1, B, V, (jt + 1) * kMT);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();

    float acc[8][8];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[r][c] = 0.0f;
#pragma unroll 3
    for (int d = 0; d < kDescDim; ++d) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[d][ty * 8]), a1 = *reinterpret_cast<const float4*>(&As[d][ty * 8 + 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[d][tx * 8]), b1 = *reinterpret_cast<const float4*>(&Bs[d][tx * 8 + 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float diff = a[r] - b[c];
          acc[r][c] = __fmaf_rn(diff, diff, acc[r][c]);
        }
    }
    // epilogue: fold into row minima (registers) and this tile's column minima
    const int cbase = jt * kMT + tx * 8, rbase = r0 + ty * 8;
    unsigned long long cbest[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) cbest[c] = ~0ull;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const bool rv = rbase + r < nA;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float dd = acc[r][c];
        const bool ok = rv && (cbase + c < nB) && (dd == dd);  // NaN never wins
        const unsigned long long pr = ok ? pack_dist(dd, cbase + c) : ~0ull;
        const unsigned long long pc = ok ? pack_dist(dd, rbase + r) : ~0ull;
        rbest[r] = umin64(rbest[r], pr);
        cbest[c] = umin64(cbest[c], pc);
      }
    }
    // columns: the two ty rows of a warp, then the 8 warps through shared memory
#pragma unroll
    for (int c = 0; c < 8; ++c) cbest[c] = umin64(cbest[c], shfl_xor_u64(cbest[c], 16));
    if ((threadIdx.x & 31) < 16) {
#pragma unroll
      for (int c = 0; c < 8; ++c) colred[warp][tx * 8 + c] = cbest[c];
    }
    __syncthreads();
    if (threadIdx.x < kMT) {
      unsigned long long m = colred[0][threadIdx.x];
#pragma unroll
      for (int w = 1; w < 8; ++w) m = umin64(m, colred[w][threadIdx.x]);
      const int col = jt * kMT + threadIdx.x;
      if (col < nB) colpart[((size_t)pair * NS + stripe) * V + col] = m;
    }
    // the next iteration's first __syncthreads orders the colred reads above before its writes,
    // and the B buffer being overwritten next was last read two barriers ago.
  }
  // rows: fold the 16 tx lanes that share a row group
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    unsigned long long m = rbest[r];
    m = umin64(m, shfl_xor_u64(m, 1));
    m = umin64(m, shfl_xor_u64(m, 2));
    m = umin64(m, shfl_xor_u64(m, 4));
    m = umin64(m, shfl_xor_u64(m, 8));
    if (tx == 0 && r0 + ty * 8 + r < nA) rowbest[(size_t)pair * V + r0 + ty * 8 + r] = m;
  }
}

// fold per-stripe column minima
__global__ void __launch_bounds__(256) match_colfold_kernel(const unsigned long long* __restrict__ colpart, const int* __restrict__ n_vox, int V,
                                                            int NS, const int* __restrict__ only, unsigned long long* __restrict__ colbest) {
  const int pair = blockIdx.y;
  if (only != nullptr && only[pair] == 0) return;
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  const int nA = n_vox[2 * pair], nB = n_vox[2 * pair + 1];
  if (col >= nB) return;
  const int ns = (nA + kMT - 1) / kMT;
  unsigned long long m = ~0ull;
  for (int s = 0; s < ns; ++s) m = umin64(m, colpart[((size_t)pair * NS + s) * V + col]);
  colbest[(size_t)pair * V + col] = m;
}

// mutual nearest neighbours, listed by ascending index in the LARGER cloud (feature_matcher.cc:84-89,146-177).
// One CTA per pair: ordered compaction with a carried block scan.
__global__ void __launch_bounds__(1024) match_mutual_kernel(const unsigned long long* __restrict__ rowbest, const unsigned long long* __restrict__ colbest,
                                                            const int* __restrict__ n_vox, int V, int* __restrict__ mut_i, int* __restrict__ mut_j,
                                                            int* __restrict__ n_mutual, int* __restrict__ swapped_out, unsigned char* __restrict__ mark,
                                                            int* __restrict__ partner) {
  __shared__ int sm[33];
  const int pair = blockIdx.x;
  const int nA = n_vox[2 * pair], nB = n_vox[2 * pair + 1];
  const bool swapped = nB > nA;
  const unsigned long long* __restrict__ rb = rowbest + (size_t)pair * V;
  const unsigned long long* __restrict__ cb = colbest + (size_t)pair * V;
  const int nI = swapped ? nB : nA;
  int carry = 0;
  for (int base = 0; base < nI; base += blockDim.x) {
    const int i = base + threadIdx.x;
    int j = -1, keep = 0;
    if (i < nI && nA > 0 && nB > 0) {
      const unsigned long long bi = swapped ? cb[i] : rb[i];
      if (bi != ~0ull) {
        j = (int)(unsigned)bi;
        const unsigned long long bj = swapped ? rb[j] : cb[j];
        keep = (bj != ~0ull && (int)(unsigned)bj == i) ? 1 : 0;
      }
    }
    int tot;
    const int ex = block_excl_scan(keep, sm, &tot);
    if (keep) {
      mut_i[(size_t)pair * V + carry + ex] = i;
      mut_j[(size_t)pair * V + carry + ex] = j;
    }
    carry += tot;
  }
  for (int t = threadIdx.x; t < V; t += blockDim.x) {
    mark[(size_t)pair * V + t] = 0;
    partner[(size_t)pair * V + t] = -1;
  }
  if (threadIdx.x == 0) {
    n_mutual[pair] = carry;
    swapped_out[pair] = swapped ? 1 : 0;
  }
}

// Matcher::normalizePoints: float mean accumulated in index order.  One warp per cloud: chunks of 32 points are staged in
// shared memory with coalesced loads (two chunks ahead); lanes 0..2 each own one coordinate and run its serial addition
// chain from shared memory (a 4-cycle dependent add per point instead of a shuffle round trip).
__global__ void __launch_bounds__(32) cloud_mean_kernel(const float4* __restrict__ pts, const int* __restrict__ n_pts, int V, float* __restrict__ mean) {
  constexpr int kChunk = 4;                 // 32-point rows per round: 4 loads per lane in flight cover the L2 / DRAM latency
  __shared__ float buf[2][3][32 * kChunk];
  const int cloud = blockIdx.x, lane = (int)lane_id();
  const int n = n_pts[cloud];
  const float4* __restrict__ p = pts + (size_t)cloud * V;
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  float acc = 0.f;
  float4 nxt[kChunk];
#pragma unroll
  for (int c = 0; c < kChunk; ++c) nxt[c] = 32 * c + lane < n ? p[32 * c + lane] : zero;
  for (int base = 0, it = 0; base < n; base += 32 * kChunk, ++it) {
    const int b = it & 1;
#pragma unroll
    for (int c = 0; c < kChunk; ++c) {
      buf[b][0][32 * c + lane] = nxt[c].x; buf[b][1][32 * c + lane] = nxt[c].y; buf[b][2][32 * c + lane] = nxt[c].z;
      const int i = base + 32 * kChunk + 32 * c + lane;
      nxt[c] = i < n ? p[i] : zero;          // in flight while this round is summed
    }
    __syncwarp();
    const int lim = min(32 * kChunk, n - base);
    if (lane < 3) {
      const float* __restrict__ src = buf[b][lane];
      if (lim == 32 * kChunk) {
#pragma unroll
        for (int l = 0; l < 32 * kChunk; ++l) acc = acc + src[l];
      } else {
        for (int l = 0; l < lim; ++l) acc = acc + src[l];
      }
    }
    // the other buffer is rewritten next iteration; this one only after another __syncwarp
  }
  if (lane < 3 && n > 0) mean[cloud * 4 + lane] = acc / (float)n;
  if (lane == 3 && n > 0) mean[cloud * 4 + 3] = 0.f;
}

__device__ __forceinline__ void philox4x32_10(unsigned long long seed, unsigned long long ctr, unsigned out[4]) {
  unsigned c0 = (unsigned)ctr, c1 = (unsigned)(ctr >> 32), c2 = 0, c3 = 0;
  unsigned k0 = (unsigned)seed, k1 = (unsigned)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const unsigned hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const unsigned n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__device__ __forceinline__ float tri_side(const float4 a, const float4 b, const float* __restrict__ m) {
  // points are mean-centred copies: (a - mean) - (b - mean), every step rounded to float
  const float dx = (a.x - m[0]) - (b.x - m[0]), dy = (a.y - m[1]) - (b.y - m[1]), dz = (a.z - m[2]) - (b.z - m[2]);
  return sqrtf((dx * dx + dy * dy) + dz * dz);
}

// tuple (triangle side-ratio) test, feature_matcher.cc:187-247: one thread per trial, counter-based RNG.  Every trial reads six
// random matched points: a CTA first stages the pair's matched points (both clouds, xyz) in shared memory when they fit
// (<= kTupleStage mutual pairs, the usual case), so the random reads stay on chip.
constexpr int kTupleStage = 1536;
constexpr int kTupleThreads = 1024;
constexpr int kTupleCtasPerPair = 8;
__global__ void __launch_bounds__(kTupleThreads) tuple_test_kernel(const float4* __restrict__ vox_pts, int V, const int* __restrict__ mut_i,
                                                                   const int* __restrict__ mut_j, const int* __restrict__ n_mutual,
                                                                   const int* __restrict__ swapped, const float* __restrict__ mean, float scale,
                                                                   int trials_per_corr, unsigned long long seed, unsigned char* __restrict__ mark) {
  __shared__ float sp[2][kTupleStage][3];
  const int pair = blockIdx.y;
  const int ncorr = n_mutual[pair];
  if (ncorr <= 0) return;
  const long long trials = (long long)ncorr * trials_per_corr;
  const bool sw = swapped[pair] != 0;
  // fi = larger cloud, fj = smaller
  const int ci = sw ? 2 * pair + 1 : 2 * pair, cj = sw ? 2 * pair : 2 * pair + 1;
  const float4* __restrict__ pi = vox_pts + (size_t)ci * V;
  const float4* __restrict__ pj = vox_pts + (size_t)cj * V;
  const float* __restrict__ mi = mean + ci * 4;
  const float* __restrict__ mj = mean + cj * 4;
  const int* __restrict__ li = mut_i + (size_t)pair * V;
  const int* __restrict__ lj = mut_j + (size_t)pair * V;
  unsigned char* __restrict__ mk = mark + (size_t)pair * V;
  const bool staged = ncorr <= kTupleStage;  // uniform for the CTA
  if (staged) {
    for (int e = threadIdx.x; e < ncorr; e += blockDim.x) {
      const float4 a = pi[li[e]], b = pj[lj[e]];
      sp[0][e][0] = a.x; sp[0][e][1] = a.y; sp[0][e][2] = a.z;
      sp[1][e][0] = b.x; sp[1][e][1] = b.y; sp[1][e][2] = b.z;
    }
    __syncthreads();
  }
  auto point = [&](int side, int r) -> float4 {
    if (staged) return make_float4(sp[side][r][0], sp[side][r][1], sp[side][r][2], 0.f);
    return side == 0 ? pi[li[r]] : pj[lj[r]];
  };
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < trials; t += (long long)gridDim.x * blockDim.x) {
    unsigned r[4];
    philox4x32_10(seed, (unsigned long long)t, r);
    const int r0 = (int)(r[0] % (unsigned)ncorr), r1 = (int)(r[1] % (unsigned)ncorr), r2 = (int)(r[2] % (unsigned)ncorr);
    const float4 a0 = point(0, r0), a1 = point(0, r1), a2 = point(0, r2);
    const float4 b0 = point(1, r0), b1 = point(1, r1), b2 = point(1, r2);
    const float li0 = tri_side(a0, a1, mi), li1 = tri_side(a1, a2, mi), li2 = tri_side(a2, a0, mi);
    const float lj0 = tri_side(b0, b1, mj), lj1 = tri_side(b1, b2, mj), lj2 = tri_side(b2, b0, mj);
    if ((li0 * scale < lj0) && (lj0 < li0 / scale) && (li1 * scale < lj1) && (lj1 < li1 / scale) && (li2 * scale < lj2) &&
        (lj2 < li2 / scale)) {
      mk[r0] = 1; mk[r1] = 1; mk[r2] = 1;
    }
  }
}

// survivors -> partner[src] = tgt  (mutual NN is a bijection, so sorting by (src,tgt) = sorting by src)
__global__ void __launch_bounds__(256) scatter_partner_kernel(const int* __restrict__ mut_i, const int* __restrict__ mut_j,
                                                              const int* __restrict__ n_mutual, const int* __restrict__ swapped,
                                                              const unsigned char* __restrict__ mark, int use_tuple, int V,
                                                              int* __restrict__ partner) {
  const int pair = blockIdx.y;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_mutual[pair]) return;
  if (use_tuple && !mark[(size_t)pair * V + e]) return;
  const int i = mut_i[(size_t)pair * V + e], j = mut_j[(size_t)pair * V + e];
  const bool sw = swapped[pair] != 0;
  const int s = sw ? j : i, t = sw ? i : j;
  partner[(size_t)pair * V + s] = t;
}

// ordered compaction over source index -> correspondence list + matched point copies
__global__ void __launch_bounds__(1024) pack_corr_kernel(const int* __restrict__ partner, const float4* __restrict__ vox_pts, const int* __restrict__ n_vox,
                                                         int V, int Lc, int* __restrict__ corr_src, int* __restrict__ corr_tgt,
                                                         float4* __restrict__ ma, float4* __restrict__ mb, int* __restrict__ n_corr,
                                                         int* __restrict__ cloud_status) {
  __shared__ int sm[33];
  const int pair = blockIdx.x;
  const int nA = n_vox[2 * pair];
  const float4* __restrict__ ps = vox_pts + (size_t)(2 * pair) * V;
  const float4* __restrict__ pt = vox_pts + (size_t)(2 * pair + 1) * V;
  int carry = 0;
  for (int base = 0; base < nA; base += blockDim.x) {
    const int s = base + threadIdx.x;
    const int t = s < nA ? partner[(size_t)pair * V + s] : -1;
    const int keep = t >= 0 ? 1 : 0;
    int tot;
    const int ex = block_excl_scan(keep, sm, &tot);
    const int o = carry + ex;
    if (keep && o < Lc) {
      corr_src[(size_t)pair * Lc + o] = s;
      corr_tgt[(size_t)pair * Lc + o] = t;
      const float4 a = ps[s], b = pt[t];
      ma[(size_t)pair * Lc + o] = make_float4(a.x, a.y, a.z, 1.0f);
      mb[(size_t)pair * Lc + o] = make_float4(b.x, b.y, b.z, 1.0f);
    }
    carry += tot;
  }
  if (threadIdx.x == 0) {
    if (carry > Lc) {
      carry = Lc;
      cloud_status[2 * pair] = QB200_CAPACITY_EXCEEDED;
    }
    n_corr[pair] = carry;
  }
}

size_t match_smem_bytes() { return 3 * sizeof(float) * kDescDim * kMT + 8 * kMT * sizeof(unsigned long long); }

// exact fp32 CUDA-core nearest neighbours (both directions).  only == nullptr: every pair; otherwise just the
// pairs flagged in only[] (the tensor-core filter's overflow fallback).
int launch_match_exact(qb200_handle* h, int n_pairs, const int* only) {
  const int V = h->V;
  const size_t smem = match_smem_bytes();
  if (int rc = ensure_dyn_smem(h, (const void*)match_stripe_kernel, smem)) return rc;
  const dim3 gs(h->NS, n_pairs);
  if (only == nullptr) cudaEventRecord(h->kev[0], h->stream);
  match_stripe_kernel<<<gs, kMatchThreads, smem, h->stream>>>(h->desc_t, h->ctr.n_vox, V, h->NS, only, h->rowbest, h->colpart);
  if (only == nullptr) {
    cudaEventRecord(h->kev[1], h->stream);
    h->kev_armed[0] = 1;
  }
  const dim3 gf((V + 255) / 256, n_pairs);
  match_colfold_kernel<<<gf, 256, 0, h->stream>>>(h->colpart, h->ctr.n_vox, V, h->NS, only, h->colbest);
  h->launches += 2;
  QB_CUDA_TRY(h, cudaGetLastError());
  return QB200_OK;
}

// QB200_TC_VERIFY: count nearest-neighbour entries whose packed (distance bits, index) differ between the two K6 implementations
__global__ void match_verify_kernel(const unsigned long long* __restrict__ rb_tc, const unsigned long long* __restrict__ cb_tc,
                                    const unsigned long long* __restrict__ rb_ex, const unsigned long long* __restrict__ cb_ex,
                                    const int* __restrict__ n_vox, int V, unsigned long long* __restrict__ stats) {
  const int pair = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  const int ns = n_vox[2 * pair], nt = n_vox[2 * pair + 1];
  if (ns <= 0 || nt <= 0) return;
  const int nr = ns, ncol = nt;  // rowbest: best target of every source point; colbest: best source of every target point
  unsigned bad = 0, cnt = 0;
  if (i < nr) { ++cnt; bad += rb_tc[(size_t)pair * V + i] != rb_ex[(size_t)pair * V + i]; }
  if (i < ncol) { ++cnt; bad += cb_tc[(size_t)pair * V + i] != cb_ex[(size_t)pair * V + i]; }
  cnt = __reduce_add_sync(0xffffffffu, cnt);
  bad = __reduce_add_sync(0xffffffffu, bad);
  if ((threadIdx.x & 31) == 0 && cnt) {
    atomicAdd(stats + 4, (unsigned long long)cnt);
    if (bad) atomicAdd(stats + 5, (unsigned long long)bad);
  }
}

int launch_match(qb200_handle* h, int n_pairs, const qb200_params& p) {
  if (n_pairs <= 0) return QB200_OK;
  const int V = h->V;
  int rc = h->force_exact_match ? launch_match_exact(h, n_pairs, nullptr) : launch_match_nn(h, n_pairs);
  if (rc) return rc;
  static const int verify = (getenv("QB200_TC_VERIFY") && getenv("QB200_TC_VERIFY")[0] == '1') ? 1 : 0;
  if (verify && !h->force_exact_match) {
    // whole-batch self-check: keep the tensor-core results, redo every pair with the exact CUDA-core kernel, compare
    unsigned long long* rb_tc = reinterpret_cast<unsigned long long*>(h->key_b);                   // the sort workspace is idle here
    unsigned long long* cb_tc = rb_tc + (size_t)h->S * V;
    QB_CUDA_TRY(h, cudaMemcpyAsync(rb_tc, h->rowbest, (size_t)n_pairs * V * 8, cudaMemcpyDeviceToDevice, h->stream));
    QB_CUDA_TRY(h, cudaMemcpyAsync(cb_tc, h->colbest, (size_t)n_pairs * V * 8, cudaMemcpyDeviceToDevice, h->stream));
    if ((rc = launch_match_exact(h, n_pairs, nullptr))) return rc;
    const dim3 gv((V + 255) / 256, n_pairs);
    match_verify_kernel<<<gv, 256, 0, h->stream>>>(rb_tc, cb_tc, h->rowbest, h->colbest, h->ctr.n_vox, V, h->tc_stats);
    h->launches += 1;
  }
  match_mutual_kernel<<<n_pairs, 1024, 0, h->stream>>>(h->rowbest, h->colbest, h->ctr.n_vox, V, h->mut_i, h->mut_j, h->ctr.n_mutual,
                                                       h->ctr.swapped, h->mark, h->partner);
  h->launches += 1;
  const int use_tuple = (p.use_tuple_test && p.tuple_scale != 0.0f) ? 1 : 0;
  if (use_tuple) {
    cloud_mean_kernel<<<2 * n_pairs, 32, 0, h->stream>>>(h->vox_pts, h->ctr.n_vox, V, h->mean);
    const dim3 gt(kTupleCtasPerPair, n_pairs);
    tuple_test_kernel<<<gt, kTupleThreads, 0, h->stream>>>(h->vox_pts, V, h->mut_i, h->mut_j, h->ctr.n_mutual, h->ctr.swapped, h->mean, p.tuple_scale,
                                                 p.tuple_trials_per_corr, (unsigned long long)p.seed, h->mark);
    h->launches += 2;
  }
  const dim3 gsc((V + 255) / 256, n_pairs);
  scatter_partner_kernel<<<gsc, 256, 0, h->stream>>>(h->mut_i, h->mut_j, h->ctr.n_mutual, h->ctr.swapped, h->mark, use_tuple, V, h->partner);
  pack_corr_kernel<<<n_pairs, 1024, 0, h->stream>>>(h->partner, h->vox_pts, h->ctr.n_vox, V, h->Lc, h->corr_src, h->corr_tgt, h->ma, h->mb,
                                                    h->ctr.n_corr, h->ctr.cloud_status);
  h->launches += 2;
  QB_CUDA_TRY(h, cudaGetLastError());
  return QB200_OK;
}

}  // namespace qb
