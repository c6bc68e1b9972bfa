// graph.cu -- K8: translation-invariant-measurement (TIM) consistency graph.   sm_100a
//
// Replaces Quatro::computeTIMs + solveForScale + the inlier_graph_.addEdge loop
// (include/quatro.hpp:307-386, 784-789; include/teaser/graph.h:96-104).  The reference
// materialises 2 x 3 x M doubles of TIMs, M index pairs and an M-byte mask (M = L(L-1)/2, ~65 B per
// pair) and then inserts edges one by one; here a pair (i,j) is tested straight from
// the two matched point sets and only the bit-packed symmetric adjacency matrix is written.
//
// Work unit: one warp owns a 32 x 32 block of the upper triangle.  Lane = column j; the 32 rows are
// broadcast from shared memory.  __ballot_sync gives the row word of the block, and each lane ORs
// its own test bit into the word of the TRANSPOSED block, so both halves of the symmetric matrix
// come out of one evaluation of the M pair tests.
//
// Arithmetic: the reference's mask is abs(db/da - 1) <= beta/da  &&  abs(da/db - 1) <= beta/db in
// fp64 (quatro.hpp:363-385), i.e. |db - da| <= beta up to rounding.  The kernel first decides in fp32
// with the sqrt-free form  u = da^2 + db^2 - beta^2;  edge <=> u <= 0 or u^2 <= 4 da^2 db^2  and a
// rigorous rounding-error bound; only pairs inside the bound (~1e-4 of them) evaluate the literal
// fp64 expression, so the result is bit-identical to the fp64 reference while >99.9 % of the pairs
// cost ~30 fp32 instructions.
#include "handle.cuh"

namespace qb {

constexpr int kGraphWarps = 8;  // 8 column blocks (256 columns) per CTA, one row block

// the literal reference expression (fp64, no FMA contraction: library is built with -fmad=false)
__device__ __noinline__ bool tim_consistent_fp64(const float4 ai, const float4 aj, const float4 bi, const float4 bj, double beta) {
  const double ax = (double)aj.x - (double)ai.x, ay = (double)aj.y - (double)ai.y, az = (double)aj.z - (double)ai.z;
  const double bx = (double)bj.x - (double)bi.x, by = (double)bj.y - (double)bi.y, bz = (double)bj.z - (double)bi.z;
  const double v1 = sqrt(ax * ax + ay * ay + az * az);
  const double v2 = sqrt(bx * bx + by * by + bz * bz);
  const double alpha_f = beta * (1.0 / v1);
  const double raw_f = v2 / v1;
  const bool in_f = fabs(raw_f - 1.0) <= alpha_f;
  const double alpha_r = beta * (1.0 / v2);
  const double raw_r = v1 / v2;
  const bool in_r = fabs(raw_r - 1.0) <= alpha_r;
  return in_f && in_r;
}

__global__ void __launch_bounds__(kGraphWarps * 32) tim_graph_kernel(const float4* __restrict__ ma, const float4* __restrict__ mb,
                                                                      const int* __restrict__ n_corr, int Lc, int W, double beta,
                                                                      uint32_t* __restrict__ adj) {
  __shared__ float4 ra[32], rb[32];
  const int pair = blockIdx.y;
  const int L = n_corr[pair];
  if (L <= 0) return;
  const int nb = (L + 31) >> 5;                       // 32-wide blocks per side
  const int ng = (nb + kGraphWarps - 1) / kGraphWarps; // column groups of 8 blocks
  const float4* __restrict__ A = ma + (size_t)pair * Lc;
  const float4* __restrict__ B = mb + (size_t)pair * Lc;
  uint32_t* __restrict__ G = adj + (size_t)pair * Lc * W;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float beta2 = (float)(beta * beta);

  for (int tile = blockIdx.x; tile < nb * ng; tile += gridDim.x) {
    const int bi = tile / ng, g = tile % ng;
    if (g * kGraphWarps + kGraphWarps - 1 < bi) continue;  // entirely below the diagonal (uniform per CTA)
    __syncthreads();
    if (threadIdx.x < 32) {
      const int i = bi * 32 + threadIdx.x;
      ra[threadIdx.x] = i < L ? A[i] : make_float4(0.f, 0.f, 0.f, 0.f);
      rb[threadIdx.x] = i < L ? B[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
    const int bj = g * kGraphWarps + warp;
    if (bj < bi || bj >= nb) continue;  // warp-uniform
    const int j = bj * 32 + lane;
    const bool vj = j < L;
    const float4 ca = vj ? A[j] : make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 cb = vj ? B[j] : make_float4(0.f, 0.f, 0.f, 0.f);
    uint32_t mine = 0, roww = 0;
#pragma unroll 4
    for (int r = 0; r < 32; ++r) {
      const int i = bi * 32 + r;
      const float4 pa = ra[r], pb = rb[r];
      const float dax = ca.x - pa.x, day = ca.y - pa.y, daz = ca.z - pa.z;
      const float dbx = cb.x - pb.x, dby = cb.y - pb.y, dbz = cb.z - pb.z;
      const float da2 = fmaf(daz, daz, fmaf(day, day, dax * dax));
      const float db2 = fmaf(dbz, dbz, fmaf(dby, dby, dbx * dbx));
      const float u = (da2 + db2) - beta2;
      const float v4 = 4.0f * da2 * db2;
      const float uu = u * u;
      const float S = u + 2.0f * beta2;  // = da2 + db2 + beta2
      bool edge = (u <= 0.0f) || (uu <= v4);
      // rigorous fp32 error bound of (uu - v4) and of u; inside it (or for coincident points) ask fp64
      const bool amb = (fabsf(uu - v4) <= 1.0e-6f * fmaf(fabsf(u), S, v4)) || (fabsf(u) <= 1.0e-6f * S) || (da2 == 0.0f) || (db2 == 0.0f);
      const bool live = vj && (i < L) && (i != j);
      if (__any_sync(0xffffffffu, amb && live)) {
        if (amb && live) edge = tim_consistent_fp64(pa, ca, pb, cb, beta);
      }
      edge = edge && live;
      const uint32_t word = __ballot_sync(0xffffffffu, edge);
      if (lane == r) roww = word;
      mine |= (edge ? 1u : 0u) << r;
    }
    // row-major half: row (bi*32 + lane), word bj.   transposed half: row j, word bi.
    const int irow = bi * 32 + lane;
    if (irow < L) G[(size_t)irow * W + bj] = roww;
    if (bj != bi && vj) G[(size_t)j * W + bi] = mine;
  }
}

// degrees + edge count + clearing of the words beyond ceil(L/32) (one warp per row)
__global__ void __launch_bounds__(256) degree_kernel(uint32_t* __restrict__ adj, const int* __restrict__ n_corr, int Lc, int W,
                                                     int* __restrict__ deg, long long* __restrict__ n_edges) {
  const int pair = blockIdx.y;
  const int L = n_corr[pair];
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= L) return;
  const int nb = (L + 31) >> 5;
  uint32_t* __restrict__ G = adj + ((size_t)pair * Lc + row) * W;
  int d = 0;
  for (int w = lane_id(); w < W; w += 32) {
    if (w < nb) d += __popc(G[w]);
    else G[w] = 0;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
  if (lane_id() == 0) {
    deg[(size_t)pair * Lc + row] = d;
    atomicAdd((unsigned long long*)(n_edges + pair), (unsigned long long)d);  // 2E; halved by the reader
  }
}

int launch_degree(qb200_handle* h, int n_pairs) {
  if (n_pairs <= 0) return QB200_OK;
  const dim3 gd((h->Lc + 7) / 8, n_pairs);
  degree_kernel<<<gd, 256, 0, h->stream>>>(h->adj, h->ctr.n_corr, h->Lc, h->W, h->deg, h->ctr.n_edges);
  h->launches++;
  QB_CUDA_TRY(h, cudaGetLastError());
  return QB200_OK;
}

int launch_graph(qb200_handle* h, int n_pairs, double noise_bound, double cbar2) {
  if (n_pairs <= 0) return QB200_OK;
  const double beta = 2 * noise_bound * sqrt(cbar2);  // quatro.hpp:367
  const dim3 g(148, n_pairs);
  cudaEventRecord(h->kev[2], h->stream);
  tim_graph_kernel<<<g, kGraphWarps * 32, 0, h->stream>>>(h->ma, h->mb, h->ctr.n_corr, h->Lc, h->W, beta, h->adj);
  cudaEventRecord(h->kev[3], h->stream);
  h->kev_armed[1] = 1;
  const dim3 gd((h->Lc + 7) / 8, n_pairs);
  degree_kernel<<<gd, 256, 0, h->stream>>>(h->adj, h->ctr.n_corr, h->Lc, h->W, h->deg, h->ctr.n_edges);
  h->launches += 2;
  QB_CUDA_TRY(h, cudaGetLastError());
  return QB200_OK;
}

}  // namespace qb
