// clique.cu -- K9: k-core decomposition + PMC heuristic maximum clique on the bit adjacency.  sm_100a
//
// Replaces teaser::MaxCliqueSolver::findMaxClique (src/graph.cc:12-130) and the pmc routines it
// calls ([EXT] pmc_graph::compute_cores, pmc_heu::search with heu_strat = "kcore").  pmc is
// downloaded by the reference's build and absent here; the semantics restated are:
//   * Batagelj-Zaversnik peeling with pmc's bucket mechanics (vertices in id order inside a bin;
//     a decremented vertex is swapped with the first vertex of its bin) -> kcore[v] = core(v)+1 and
//     the peel order;
//   * for start vertices in reverse peel order: P = {u in adj(v): kcore[u] > mc}; if |P| > mc the
//     greedy descent pops the candidate with the largest (kcore, id) and intersects P with its
//     neighbourhood until P is empty; a longer chain replaces the incumbent; stop at mc >= max_core+1.
// The reference runs the start vertices on 12 racy OpenMP threads; like the CPU oracle this kernel
// executes them in sequence, one WARP per registration pair -- the batch of pairs is
// the parallel axis -- with warp-wide bit-set operations inside a step:
//   - the peel loop expands a row to its neighbour list with ballot/popc prefix sums and applies the
//     bucket swaps of 32 neighbours at a time, serialising only neighbours of equal degree;
//   - vertices are renumbered by (kcore, id) rank (permute_adj_kernel), so "largest (kcore,id)
//     candidate" is the highest set bit of P and a descent step is one 128-word AND.
#include "handle.cuh"

namespace qb {

constexpr int kMaxWordsPerLane = 4;  // Lc <= 4096  ->  W <= 128 words per row

// one REDUX instead of five dependent shuffles: these sit on the serial chain of the single-warp kernels below
__device__ __forceinline__ int warp_max(int v) { return __reduce_max_sync(0xffffffffu, v); }
__device__ __forceinline__ int warp_sum(int v) { return __reduce_add_sync(0xffffffffu, v); }

// stable counting sort of vertices by key[v] (ids ascending inside a bucket); bin[d] ends up as the
// START of bucket d (d = 0..maxkey), bin[maxkey+1] = n.  One warp.
__device__ void warp_bucket_sort(const unsigned short* __restrict__ key, int n, int maxkey, int* __restrict__ bin,
                                 unsigned short* __restrict__ pos, unsigned short* __restrict__ vert) {
  const int lane = lane_id();
  for (int d = lane; d <= maxkey + 1; d += 32) bin[d] = 0;
  __syncwarp();
  for (int v = lane; v < n; v += 32) atomicAdd(&bin[key[v]], 1);
  __syncwarp();
  int carry = 0;
  for (int base = 0; base <= maxkey + 1; base += 32) {  // exclusive scan
    const int d = base + lane;
    const int c = d <= maxkey + 1 ? bin[d] : 0;
    int tot;
    const int ex = warp_excl_scan(c, &tot);
    if (d <= maxkey + 1) bin[d] = carry + ex;
    carry += tot;
  }
  __syncwarp();
  for (int base = 0; base < n; base += 32) {  // placement, ids ascending inside a bucket
    const int v = base + lane;
    const bool act = v < n;
    const int d = act ? key[v] : -1 - lane;
    const unsigned peers = __match_any_sync(0xffffffffu, d);
    const int rank = __popc(peers & ((1u << lane) - 1));
    const int b = act ? bin[d] : 0;
    __syncwarp();
    if (act) {
      pos[v] = (unsigned short)(b + rank);
      vert[b + rank] = (unsigned short)v;
      if (rank == 0) bin[d] = b + __popc(peers);
    }
    __syncwarp();
  }
  // bin[d] is now the END of bucket d: shift down to starts (descending chunks)
  for (int base = ((maxkey + 1) / 32) * 32; base >= 0; base -= 32) {
    const int d = base + lane;
    int prev = 0;
    if (d >= 1 && d <= maxkey + 1) prev = bin[d - 1];
    __syncwarp();
    if (d >= 1 && d <= maxkey + 1) bin[d] = prev;
    __syncwarp();
  }
  if (lane == 0) bin[0] = 0;
  __syncwarp();
}

// One warp per pair.  smem layout: deg, pos, vert, nbl (u16 x Lc each), bin (int x (Lc + 2)).
__global__ void __launch_bounds__(32) kcore_kernel(const uint32_t* __restrict__ adj, const int* __restrict__ deg_in, const int* __restrict__ n_corr,
                                                   int Lc, int W, int cache_words, int* __restrict__ kcore, int* __restrict__ korder,
                                                   int* __restrict__ rank_of, int* __restrict__ by_rank, int* __restrict__ kbin,
                                                   int* __restrict__ max_core_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned short* deg = reinterpret_cast<unsigned short*>(smem_raw);
  unsigned short* pos = deg + Lc;
  unsigned short* vert = pos + Lc;
  unsigned short* nbl = vert + Lc;
  int* bin = reinterpret_cast<int*>(nbl + Lc);
  uint32_t* cache = reinterpret_cast<uint32_t*>(bin + Lc + 2);  // [cache_words] adjacency rows when the graph fits
  const int pair = blockIdx.x, lane = lane_id();
  const int L = n_corr[pair];
  int* __restrict__ kc = kcore + (size_t)pair * (Lc + 2);
  int* __restrict__ ko = korder + (size_t)pair * (Lc + 2);
  int* __restrict__ ro = rank_of + (size_t)pair * (Lc + 2);
  int* __restrict__ br = by_rank + (size_t)pair * (Lc + 2);
  int* __restrict__ kb = kbin + (size_t)pair * (Lc + 2);
  if (L <= 0) {
    if (lane == 0) max_core_out[pair] = 0;
    return;
  }
  const uint32_t* __restrict__ G = adj + (size_t)pair * Lc * W;
  const int nbw = (L + 31) >> 5;                 // adjacency words per row
  const int nwl = (nbw + 31) >> 5;               // adjacency words per lane (<= kMaxWordsPerLane)
  // The peel is a chain of dependent row loads; a graph that fits is staged in shared memory once.
  const bool cached = (long long)L * nbw <= (long long)cache_words;
  if (cached) {
    for (int idx = lane; idx < L * nbw; idx += 32) cache[idx] = G[(size_t)(idx / nbw) * W + (idx % nbw)];
    __syncwarp();
  }
  auto row_word = [&](int v, int wi) -> uint32_t {
    if (wi >= nbw) return 0u;
    return cached ? cache[v * nbw + wi] : G[(size_t)v * W + wi];
  };

  int md = 0;
  for (int v = lane; v < L; v += 32) {
    const int d = deg_in[(size_t)pair * Lc + v];
    deg[v] = (unsigned short)d;
    md = max(md, d);
  }
  md = warp_max(md);
  __syncwarp();
  warp_bucket_sort(deg, L, md, bin, pos, vert);

  // ---- peel ----
  uint32_t wn[kMaxWordsPerLane];
  int guess = vert[0];
#pragma unroll
  for (int k = 0; k < kMaxWordsPerLane; ++k) wn[k] = (k < nwl) ? row_word(guess, lane + 32 * k) : 0u;
  for (int i = 0; i < L; ++i) {
    const int v = vert[i];
    const int dv = deg[v];
    uint32_t w[kMaxWordsPerLane];
    if (v == guess) {
#pragma unroll
      for (int k = 0; k < kMaxWordsPerLane; ++k) w[k] = wn[k];
    } else {
#pragma unroll
      for (int k = 0; k < kMaxWordsPerLane; ++k) w[k] = (k < nwl) ? row_word(v, lane + 32 * k) : 0u;
    }
    // speculative prefetch of the next row (the vertex at position i+1 rarely changes while v is processed)
    guess = (i + 1 < L) ? vert[i + 1] : v;
#pragma unroll
    for (int k = 0; k < kMaxWordsPerLane; ++k) wn[k] = (k < nwl) ? row_word(guess, lane + 32 * k) : 0u;
    if (dv == 0) continue;  // current degree 0: no unprocessed neighbour left, processed ones have degree <= 0
    // expand the row to an ascending neighbour list
    int cnt = 0;
#pragma unroll
    for (int k = 0; k < kMaxWordsPerLane; ++k) {
      if (k < nwl) {
        int tot;
        int off = cnt + warp_excl_scan(__popc(w[k]), &tot);
        uint32_t x = w[k];
        const int basebit = (lane + 32 * k) * 32;
        while (x) {
          const int b = __ffs(x) - 1;
          x &= x - 1;
          nbl[off++] = (unsigned short)(basebit + b);
        }
        cnt += tot;
      }
    }
    __syncwarp();
    // bucket updates, 32 neighbours at a time; neighbours of equal degree are serialised in id order
    for (int c0 = 0; c0 < cnt; c0 += 32) {
      const int e = c0 + lane;
      const int u = e < cnt ? nbl[e] : 0;
      const int du = e < cnt ? deg[u] : 0;
      const bool act = e < cnt && du > dv;
      const unsigned amask = __ballot_sync(0xffffffffu, act);
      if (amask == 0) continue;
      const unsigned grp = __match_any_sync(0xffffffffu, act ? du : -1 - lane);
      const int rank = __popc(grp & ((1u << lane) - 1));
      const int rounds = warp_max(act ? __popc(grp) : 0);
      for (int t = 0; t < rounds; ++t) {
        if (act && rank == t) {
          const int pu = pos[u], pw = bin[du], wv = vert[pw];
          if (u != wv) {
            pos[u] = (unsigned short)pw; vert[pu] = (unsigned short)wv;
            pos[wv] = (unsigned short)pu; vert[pw] = (unsigned short)u;
          }
          bin[du] = pw + 1;
          deg[u] = (unsigned short)(du - 1);
        }
        __syncwarp();
      }
    }
    __syncwarp();
  }
  // ---- outputs: kcore = core + 1, peel order, max core ----
  const int max_core = deg[vert[L - 1]];
  for (int v = lane; v < L; v += 32) {
    kc[v] = (int)deg[v] + 1;
    ko[v] = vert[v];
  }
  if (lane == 0) max_core_out[pair] = max_core;
  __syncwarp();
  // ---- (kcore, id) ranks for the clique search: stable bucket sort by core number ----
  warp_bucket_sort(deg, L, max_core, bin, pos, vert);
  for (int v = lane; v < L; v += 32) {
    ro[v] = pos[v];
    br[v] = vert[v];
  }
  for (int d = lane; d <= max_core + 1; d += 32) kb[d] = bin[d];
}

// adjacency rows/columns renumbered by rank: adjp[rank(v)] bit rank(u) = adj[v] bit u.  One warp per row.
__global__ void __launch_bounds__(256) permute_adj_kernel(const uint32_t* __restrict__ adj, const int* __restrict__ n_corr, int Lc, int W,
                                                          const int* __restrict__ rank_of, uint32_t* __restrict__ adjp) {
  extern __shared__ uint32_t rows[];  // [8][W]
  const int pair = blockIdx.y;
  const int L = n_corr[pair];
  const int wib = threadIdx.x >> 5, lane = lane_id();
  const int v = blockIdx.x * 8 + wib;
  if (v >= L) return;
  uint32_t* row = rows + wib * W;
  const int nb = (L + 31) >> 5;
  for (int w = lane; w < W; w += 32) row[w] = 0;
  __syncwarp();
  const int* __restrict__ ro = rank_of + (size_t)pair * (Lc + 2);
  const uint32_t* __restrict__ src = adj + ((size_t)pair * Lc + v) * W;
  for (int w = lane; w < nb; w += 32) {
    uint32_t x = src[w];
    while (x) {
      const int b = __ffs(x) - 1;
      x &= x - 1;
      const int r = ro[w * 32 + b];
      atomicOr(&row[r >> 5], 1u << (r & 31));
    }
  }
  __syncwarp();
  uint32_t* __restrict__ dst = adjp + ((size_t)pair * Lc + ro[v]) * W;
  for (int w = lane; w < W; w += 32) dst[w] = row[w];
}

// One warp per pair: PMC heuristic in rank space.  smem: cl (u16 x Lc) current chain, ids bitset (W words).
__global__ void __launch_bounds__(32) clique_kernel(const uint32_t* __restrict__ adjp, const int* __restrict__ n_corr, int Lc, int W,
                                                    const int* __restrict__ kcore, const int* __restrict__ korder, const int* __restrict__ rank_of,
                                                    const int* __restrict__ by_rank, const int* __restrict__ kbin, const int* __restrict__ max_core_in,
                                                    int mode, double kcore_thr, int cache_words, int* __restrict__ clique,
                                                    int* __restrict__ n_clique) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned short* chain = reinterpret_cast<unsigned short*>(smem_raw);      // [Lc]
  uint32_t* idbits = reinterpret_cast<uint32_t*>(chain + Lc);               // [W]
  uint32_t* cache = idbits + W;                                              // [cache_words]
  const int pair = blockIdx.x, lane = lane_id();
  const int L = n_corr[pair];
  int* __restrict__ out = clique + (size_t)pair * Lc;
  if (L <= 0) {
    if (lane == 0) n_clique[pair] = 0;
    return;
  }
  const int* __restrict__ kc = kcore + (size_t)pair * (Lc + 2);
  const int* __restrict__ ko = korder + (size_t)pair * (Lc + 2);
  const int* __restrict__ ro = rank_of + (size_t)pair * (Lc + 2);
  const int* __restrict__ br = by_rank + (size_t)pair * (Lc + 2);
  const int* __restrict__ kb = kbin + (size_t)pair * (Lc + 2);
  const uint32_t* __restrict__ G = adjp + (size_t)pair * Lc * W;
  const int max_core = max_core_in[pair];
  const int nbw = (L + 31) >> 5;
  const int nwl = (nbw + 31) >> 5;
  // every descent step is a dependent row load: keep the rank-space adjacency in shared memory when it fits
  const bool cached = (long long)L * nbw <= (long long)cache_words;
  if (cached) {
    for (int idx = lane; idx < L * nbw; idx += 32) cache[idx] = G[(size_t)(idx / nbw) * W + (idx % nbw)];
  }
  auto row_word = [&](int r, int wi) -> uint32_t {
    if (wi >= nbw) return 0u;
    return cached ? cache[r * nbw + wi] : G[(size_t)r * W + wi];
  };
  for (int w = lane; w < W; w += 32) idbits[w] = 0;
  __syncwarp();
  int csize = 0;

  if (mode == QB200_KCORE_HEU && kcore_thr != 1.0 && max_core > (int)(kcore_thr * (double)L)) {
    // src/graph.cc:67-82: keep every vertex whose k_cores entry (core + 1) is >= max_core
    for (int v = lane; v < L; v += 32)
      if (kc[v] >= max_core) atomicOr(&idbits[v >> 5], 1u << (v & 31));
  } else {
    const int ub = max_core + 1;  // src/graph.cc:84-86
    int mc = 0;
    for (int i = L - 1; i >= 0; --i) {
      if (mc >= ub) break;
      const int v = ko[i];
      if (kc[v] <= mc) break;  // kcore is non-increasing along the reversed peel order and mc only grows
      // candidates: neighbours of v with kcore > mc  <=>  rank >= kb[mc]  (first rank whose core >= mc)
      const int thr = kb[min(mc, max_core + 1)];
      const int rv = ro[v];
      uint32_t P[kMaxWordsPerLane];
      int psize = 0;
#pragma unroll
      for (int k = 0; k < kMaxWordsPerLane; ++k) {
        P[k] = 0;
        if (k < nwl) {
          const int wi = lane + 32 * k;
          uint32_t x = row_word(rv, wi);
          const int lo = wi * 32;
          if (thr >= lo + 32) x = 0;
          else if (thr > lo) x &= ~0u << (thr - lo);
          P[k] = x;
          psize += __popc(x);
        }
      }
      psize = warp_sum(psize);
      if (psize <= mc) continue;
      int sz = 1;
      if (nwl == 1) {
        // L <= 1024 (the usual case): one adjacency word per lane, the step is clz -> REDUX -> one row word -> AND
        uint32_t p0 = P[0];
        for (;;) {
          const int top = warp_max(p0 ? lane * 32 + 31 - __clz(p0) : -1);  // highest set bit of P across the warp
          if (top < 0) break;
          if (lane == 0) chain[sz - 1] = (unsigned short)top;
          ++sz;
          p0 &= row_word(top, lane);
        }
      } else {
        for (;;) {
          // highest set bit of P across the warp
          int top = -1;
#pragma unroll
          for (int k = 0; k < kMaxWordsPerLane; ++k)
            if (k < nwl && P[k]) top = max(top, (lane + 32 * k) * 32 + 31 - __clz(P[k]));
          top = warp_max(top);
          if (top < 0) break;
          if (lane == 0) chain[sz - 1] = (unsigned short)top;
          ++sz;
#pragma unroll
          for (int k = 0; k < kMaxWordsPerLane; ++k)
            if (k < nwl) P[k] &= row_word(top, lane + 32 * k);
        }
      }
      if (sz > mc) {
        mc = sz;
        __syncwarp();
        for (int w = lane; w < W; w += 32) idbits[w] = 0;
        __syncwarp();
        for (int t = lane; t < sz - 1; t += 32) {
          const int id = br[chain[t]];
          atomicOr(&idbits[id >> 5], 1u << (id & 31));
        }
        if (lane == 0) atomicOr(&idbits[v >> 5], 1u << (v & 31));
        __syncwarp();
      }
    }
  }
  __syncwarp();
  // ascending ids (std::sort(max_clique_), quatro.hpp:806)
  for (int base = 0; base < W; base += 32) {
    const int w = base + lane;
    uint32_t x = w < W ? idbits[w] : 0;
    int tot;
    int off = csize + warp_excl_scan(__popc(x), &tot);
    while (x) {
      const int b = __ffs(x) - 1;
      x &= x - 1;
      out[off++] = w * 32 + b;
    }
    csize += tot;
  }
  if (lane == 0) n_clique[pair] = csize;
}

int launch_clique(qb200_handle* h, int n_pairs, int mode, double kcore_thr) {
  if (n_pairs <= 0) return QB200_OK;
  if (mode == QB200_PMC_EXACT) return QB200_ERR_UNSUPPORTED;
  const int Lc = h->Lc, W = h->W;
  // shared-memory adjacency cache: 14336 words (56 KB) hold graphs up to L ~ 660; two CTAs per SM still fit
  const int cache_words = 14336;
  const size_t sm_kcore = (size_t)4 * Lc * sizeof(unsigned short) + (size_t)(Lc + 2) * sizeof(int) + (size_t)cache_words * 4;
  const size_t sm_clique = (size_t)Lc * sizeof(unsigned short) + (size_t)W * sizeof(uint32_t) + (size_t)cache_words * 4;
  if (!(h->func_attr_set & 1u)) {  // per handle: the opt-in is a per-device property of the function
    QB_CUDA_TRY(h, cudaFuncSetAttribute(kcore_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_kcore));
    QB_CUDA_TRY(h, cudaFuncSetAttribute(clique_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_clique));
    h->func_attr_set |= 1u;
  }
  kcore_kernel<<<n_pairs, 32, sm_kcore, h->stream>>>(h->adj, h->deg, h->ctr.n_corr, Lc, W, cache_words, h->kcore, h->korder, h->rank_of,
                                                     h->by_rank, h->kbin, h->ctr.max_core);
  const dim3 gp((Lc + 7) / 8, n_pairs);
  permute_adj_kernel<<<gp, 256, 8 * W * sizeof(uint32_t), h->stream>>>(h->adj, h->ctr.n_corr, Lc, W, h->rank_of, h->adjp);
  clique_kernel<<<n_pairs, 32, sm_clique, h->stream>>>(h->adjp, h->ctr.n_corr, Lc, W, h->kcore, h->korder, h->rank_of, h->by_rank, h->kbin,
                                                       h->ctr.max_core, mode, kcore_thr, cache_words, h->clique, h->ctr.n_clique);
  h->launches += 3;
  QB_CUDA_TRY(h, cudaGetLastError());
  return QB200_OK;
}

}  // namespace qb
