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
// The reference runs the start vertices on 12 racy OpenMP threads; like the CPU oracle these kernels
// reproduce the SEQUENTIAL semantics bit for bit.  One CTA per registration pair:
//
//   kcore_cta_kernel   The peel stays a sequence of L steps (the order is the output), but a step is CTA-wide:
//     thread t owns adjacency word t of the row being peeled; `above` (bitset of vertices whose current degree exceeds
//     the current level) turns "neighbour with deg[u] > deg[v]" into one AND; the surviving neighbours are expanded
//     to an ascending list with one block scan and every neighbour's bucket move is done by its own thread.  Moves
//     into different buckets commute; the members of one bucket (same current degree) must be applied in id order:
//     their ranks come from __match_any_sync inside a warp and a packed per-warp size word across warps, and with the
//     ranks known the moves of a group have a closed form (member t lands on slot bin+t, the displaced vertex takes
//     the member's old slot) unless a member already sits inside the target slots -- then that group is replayed
//     serially by one thread while the other groups proceed in parallel.  The next row is prefetched while the current one is processed.
//   clique_cta_kernel  The start vertices are tried SPECULATIVELY, one per warp, against the current incumbent size
//     mc; results are committed in sequential order (the first warp that beats mc wins, later warps are discarded and
//     redone), which is exactly the sequential outcome because a descent only depends on mc at its start.  Vertices
//     are renumbered by (kcore, id) rank (permute_adj_kernel), so "largest (kcore,id) candidate" is the highest set
//     bit of P and a descent step is one multi-word AND.
#include "handle.cuh"

namespace qb {

__device__ __forceinline__ int warp_max(int v) { return __reduce_max_sync(0xffffffffu, v); }
__device__ __forceinline__ int warp_sum(int v) { return __reduce_add_sync(0xffffffffu, v); }

// per-warp group sizes (0..32, 6 bits each) packed into one word: warp k owns bits [6k, 6k+6)
template <int NW, typename Acc>
__device__ __forceinline__ void unpack_sizes(Acc a, int warp, int& before, int& total) {
  before = 0; total = 0;
#pragma unroll
  for (int k = 0; k < NW; ++k) {
    const int f = (int)((a >> (6 * k)) & (Acc)63);
    if (k < warp) before += f;
    total += f;
  }
}

// Stable counting sort of vertices by key[v] (ids ascending inside a bucket) by the whole CTA.  On exit bin[d] is the
// START of bucket d (d = 0..maxkey) and bin[maxkey+1] = n.  acc[] must be all zero on entry and is all zero on exit.
// tmp: n ints of scratch.  Every thread of the block calls it.
template <int NW, typename Acc>
__device__ void block_bucket_sort(const unsigned short* __restrict__ key, int n, int maxkey, int* __restrict__ bin, Acc* __restrict__ acc,
                                  unsigned short* __restrict__ pos, unsigned short* __restrict__ vert, int* __restrict__ tmp, int* scan_smem) {
  constexpr int NT = NW * 32;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int d = tid; d <= maxkey + 1; d += NT) bin[d] = 0;
  __syncthreads();
  for (int v = tid; v < n; v += NT) atomicAdd(&bin[key[v]], 1);
  __syncthreads();
  int carry = 0;
  for (int base = 0; base <= maxkey + 1; base += NT) {
    const int d = base + tid;
    const int c = d <= maxkey + 1 ? bin[d] : 0;
    int tot;
    const int ex = block_excl_scan(c, scan_smem, &tot);
    if (d <= maxkey + 1) bin[d] = carry + ex;
    carry += tot;
  }
  __syncthreads();
  for (int base = 0; base < n; base += NT) {
    const int v = base + tid;
    const bool act = v < n;
    const int d = act ? key[v] : -1 - lane;
    const unsigned peers = __match_any_sync(0xffffffffu, d);
    const int rank_w = __popc(peers & ((1u << lane) - 1));
    if (act && rank_w == 0) atomicAdd(&acc[d], (Acc)__popc(peers) << (6 * warp));
    __syncthreads();
    int before = 0, total = 0, b = 0;
    if (act) {
      unpack_sizes<NW, Acc>(acc[d], warp, before, total);
      b = bin[d];
      const int p = b + before + rank_w;
      pos[v] = (unsigned short)p;
      vert[p] = (unsigned short)v;
    }
    __syncthreads();
    if (act && before + rank_w == 0) { bin[d] = b + total; acc[d] = 0; }
    __syncthreads();
  }
  // bin[d] is now the END of bucket d: shift down to starts
  for (int d = tid; d <= maxkey; d += NT) tmp[d] = bin[d];
  __syncthreads();
  for (int d = tid; d <= maxkey; d += NT) bin[d + 1] = tmp[d];
  if (tid == 0) bin[0] = 0;
  __syncthreads();
}

// One CTA (NW warps) per pair.  Requires W <= NW * 32.
template <int NW, typename Acc>
__global__ void __launch_bounds__(NW * 32) kcore_cta_kernel(const uint32_t* __restrict__ adj, const int* __restrict__ deg_in,
                                                            const int* __restrict__ n_corr, int Lc, int W, int* __restrict__ kcore,
                                                            int* __restrict__ korder, int* __restrict__ rank_of, int* __restrict__ by_rank,
                                                            int* __restrict__ kbin, int* __restrict__ max_core_out) {
  constexpr int NT = NW * 32;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Acc* acc = reinterpret_cast<Acc*>(smem_raw);                       // [Lc + 2] packed group sizes per degree (zero between uses)
  int* bin = reinterpret_cast<int*>(acc + Lc + 2);                   // [Lc + 2]
  uint32_t* above = reinterpret_cast<uint32_t*>(bin + Lc + 2);       // [W] vertices with current degree > current level
  unsigned short* deg = reinterpret_cast<unsigned short*>(above + W);
  unsigned short* pos = deg + Lc;
  unsigned short* vert = pos + Lc;
  unsigned short* nbl = vert + Lc;                                   // ascending list of the current step's live neighbours
  unsigned short* slot = nbl + Lc;                                   // slot[q] = member that lands on position q (serial replay)
  unsigned char* gflag = reinterpret_cast<unsigned char*>(slot + Lc);  // [2][Lc] group (keyed by its first slot) needs the serial replay
  __shared__ int s_wtot[2][NW];
  __shared__ int s_scan[33];
  __shared__ int s_red[NW];

  const int pair = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int L = n_corr[pair];
  int* __restrict__ kc = kcore + (size_t)pair * (Lc + 2);
  int* __restrict__ ko = korder + (size_t)pair * (Lc + 2);
  int* __restrict__ ro = rank_of + (size_t)pair * (Lc + 2);
  int* __restrict__ br = by_rank + (size_t)pair * (Lc + 2);
  int* __restrict__ kb = kbin + (size_t)pair * (Lc + 2);
  if (L <= 0) {
    if (tid == 0) max_core_out[pair] = 0;
    return;
  }
  const uint32_t* __restrict__ G = adj + (size_t)pair * Lc * W;
  const int nbw = (L + 31) >> 5;  // adjacency words per row (<= NT)

  int md = 0;
  for (int v = tid; v < L; v += NT) {
    const int d = deg_in[(size_t)pair * Lc + v];
    deg[v] = (unsigned short)d;
    md = max(md, d);
  }
  for (int d = tid; d < Lc + 2; d += NT) acc[d] = 0;
  for (int d = tid; d < 2 * Lc; d += NT) gflag[d] = 0;
  md = warp_max(md);
  if (lane == 0) s_red[warp] = md;
  __syncthreads();
  md = 0;
#pragma unroll
  for (int k = 0; k < NW; ++k) md = max(md, s_red[k]);
  block_bucket_sort<NW, Acc>(deg, L, md, bin, acc, pos, vert, reinterpret_cast<int*>(nbl), s_scan);

  // ---- peel ----
  for (int w = tid; w < W; w += NT) {
    const int lo = w * 32;
    above[w] = lo + 32 <= L ? ~0u : (lo < L ? (1u << (L - lo)) - 1u : 0u);
  }
  int cur = -1;  // current level: every vertex with degree <= cur has its bit in `above` cleared
  int par = 0, pend = -1;  // gflag buffer of the current pass; flag this thread still has to clear in the other buffer
  int guess = vert[0];
  uint32_t wn = tid < nbw ? G[(size_t)guess * W + tid] : 0u;
  __syncthreads();
  for (int i = 0; i < L; ++i) {
    const int v = vert[i];
    const int dv = deg[v];
    uint32_t w = wn;
    if (v != guess) w = tid < nbw ? G[(size_t)v * W + tid] : 0u;  // the speculation failed (block-uniform)
    if (i + 1 < L) {  // positions inside the current bucket are final: the vertex at i+1 rarely changes during this step
      guess = vert[i + 1];
      wn = tid < nbw ? G[(size_t)guess * W + tid] : 0u;
    }
    if (dv > cur) {  // level rise (block-uniform): bucket dv = positions [i, bin[dv+1]) leaves `above`
      const int end = bin[dv + 1];
      for (int p = i + tid; p < end; p += NT) {
        const int x = vert[p];
        atomicAnd(&above[x >> 5], ~(1u << (x & 31)));
      }
      cur = dv;
      __syncthreads();
    }
    uint32_t aw = tid < nbw ? (w & above[tid]) : 0u;
    const int c = __popc(aw);
    int inc = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int nb = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += nb;
    }
    if (lane == 31) s_wtot[i & 1][warp] = inc;
    __syncthreads();  // (A)
    int base = 0, cnt = 0;
#pragma unroll
    for (int k = 0; k < NW; ++k) {
      const int t = s_wtot[i & 1][k];
      if (k < warp) base += t;
      cnt += t;
    }
    if (cnt == 0) continue;  // block-uniform; s_wtot is double-buffered, (A) of the next step orders the reuse
    int off = base + inc - c;
    while (aw) {
      const int b = __ffs(aw) - 1;
      aw &= aw - 1;
      nbl[off++] = (unsigned short)(tid * 32 + b);
    }
    __syncthreads();  // (B)
    for (int c0 = 0; c0 < cnt; c0 += NT) {
      const int e = c0 + tid;
      const bool act = e < cnt;
      const int u = act ? nbl[e] : 0;
      const int du = act ? deg[u] : -1 - lane;  // every listed neighbour has du > dv
      const unsigned grp = __match_any_sync(0xffffffffu, du);
      const int rank_w = __popc(grp & ((1u << lane) - 1));
      if (act && rank_w == 0) atomicAdd(&acc[du], (Acc)__popc(grp) << (6 * warp));
      if (pend >= 0) { gflag[(par ^ 1) * Lc + pend] = 0; pend = -1; }
      __syncthreads();  // (b)
      int rank = 0, m = 0, b0 = 0, pu = 0, q = 0, wv = 0;
      if (act) {
        int before;
        unpack_sizes<NW, Acc>(acc[du], warp, before, m);
        rank = before + rank_w;
        b0 = bin[du];
        pu = pos[u];
        q = b0 + rank;
        wv = vert[q];
        slot[q] = (unsigned short)u;
        if (m > 1 && pu < b0 + m) gflag[par * Lc + b0] = 1;  // a member already sits inside the target slots: replay this group serially
      }
      __syncthreads();  // (d)
      if (act) {
        const bool serial = gflag[par * Lc + b0] != 0;
        if (!serial) {
          if (pu != q) {  // member `rank` lands on slot b0+rank, the displaced vertex takes its old slot
            vert[q] = (unsigned short)u; pos[u] = (unsigned short)q;
            vert[pu] = (unsigned short)wv; pos[wv] = (unsigned short)pu;
          }
        } else if (rank == 0) {
          for (int t = 0; t < m; ++t) {
            const int uu = slot[b0 + t], pw = b0 + t, p2 = pos[uu], w2 = vert[pw];
            if (uu != w2) {
              pos[uu] = (unsigned short)pw; vert[p2] = (unsigned short)w2;
              pos[w2] = (unsigned short)p2; vert[pw] = (unsigned short)uu;
            }
          }
        }
        deg[u] = (unsigned short)(du - 1);
        if (du - 1 == dv) atomicAnd(&above[u >> 5], ~(1u << (u & 31)));
        if (rank == 0) { bin[du] = b0 + m; acc[du] = 0; pend = b0; }
      }
      __syncthreads();  // (f)
      par ^= 1;
    }
  }
  __syncthreads();
  // ---- outputs: kcore = core + 1, peel order, max core ----
  const int max_core = deg[vert[L - 1]];
  for (int v = tid; v < L; v += NT) {
    kc[v] = (int)deg[v] + 1;
    ko[v] = vert[v];
  }
  if (tid == 0) max_core_out[pair] = max_core;
  __syncthreads();
  // ---- (kcore, id) ranks for the clique search: stable bucket sort by core number ----
  block_bucket_sort<NW, Acc>(deg, L, max_core, bin, acc, pos, vert, reinterpret_cast<int*>(nbl), s_scan);
  for (int v = tid; v < L; v += NT) {
    ro[v] = pos[v];
    br[v] = vert[v];
  }
  for (int d = tid; d <= max_core + 1; d += NT) kb[d] = bin[d];
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

constexpr int kCliqueWarps = 8;

// One greedy descent by one warp in rank space.  P = candidates of start vertex rv with rank >= thr; returns the chain
// length + 1 (the start vertex), or 0 when |P| <= mc.  WPL = adjacency words per lane.
template <int WPL>
__device__ __forceinline__ int clique_descent(const uint32_t* __restrict__ rows, int stride, int nbw, int rv, int thr, int mc,
                                              unsigned short* __restrict__ chain) {
  const int lane = lane_id();
  uint32_t P[WPL];
  int psize = 0;
#pragma unroll
  for (int k = 0; k < WPL; ++k) {
    const int wi = lane + 32 * k;
    uint32_t x = wi < nbw ? rows[(size_t)rv * stride + wi] : 0u;
    const int lo = wi * 32;
    if (thr >= lo + 32) x = 0;
    else if (thr > lo) x &= ~0u << (thr - lo);
    P[k] = x;
    psize += __popc(x);
  }
  psize = warp_sum(psize);
  if (psize <= mc) return 0;
  int sz = 1;
  for (;;) {
    int top = -1;  // highest set bit of P across the warp = largest (kcore, id)
#pragma unroll
    for (int k = 0; k < WPL; ++k)
      if (P[k]) top = max(top, (lane + 32 * k) * 32 + 31 - __clz(P[k]));
    top = warp_max(top);
    if (top < 0) break;
    if (lane == 0) chain[sz - 1] = (unsigned short)top;
    ++sz;
#pragma unroll
    for (int k = 0; k < WPL; ++k) {
      const int wi = lane + 32 * k;
      P[k] &= wi < nbw ? rows[(size_t)top * stride + wi] : 0u;
    }
  }
  return sz;
}

// One CTA (8 warps) per pair: PMC heuristic in rank space, start vertices tried speculatively by the warps and committed
// in sequential order.  smem: chains [8][Lc] u16, ids bitset [W], adjacency cache [cache_words].
__global__ void __launch_bounds__(kCliqueWarps * 32) clique_cta_kernel(const uint32_t* __restrict__ adjp, const int* __restrict__ n_corr, int Lc, int W,
                                                                       const int* __restrict__ kcore, const int* __restrict__ korder,
                                                                       const int* __restrict__ rank_of, const int* __restrict__ by_rank,
                                                                       const int* __restrict__ kbin, const int* __restrict__ max_core_in, int mode,
                                                                       double kcore_thr, int cache_words, int* __restrict__ clique,
                                                                       int* __restrict__ n_clique) {
  constexpr int NT = kCliqueWarps * 32;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned short* chains = reinterpret_cast<unsigned short*>(smem_raw);       // [8][Lc]
  uint32_t* idbits = reinterpret_cast<uint32_t*>(chains + (size_t)kCliqueWarps * Lc);  // [W]
  uint32_t* cache = idbits + W;                                                // [cache_words]
  __shared__ int s_sz[kCliqueWarps];
  __shared__ int s_scan[33];
  const int pair = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int L = n_corr[pair];
  int* __restrict__ out = clique + (size_t)pair * Lc;
  if (L <= 0) {
    if (tid == 0) n_clique[pair] = 0;
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
  for (int w = tid; w < W; w += NT) idbits[w] = 0;

  if (mode == QB200_KCORE_HEU && kcore_thr != 1.0 && max_core > (int)(kcore_thr * (double)L)) {
    // src/graph.cc:67-82: keep every vertex whose k_cores entry (core + 1) is >= max_core
    __syncthreads();
    for (int v = tid; v < L; v += NT)
      if (kc[v] >= max_core) atomicOr(&idbits[v >> 5], 1u << (v & 31));
  } else {
    // every descent step is a dependent row load: keep the rank-space adjacency in shared memory when it fits
    const bool cached = (long long)L * nbw <= (long long)cache_words;
    if (cached)
      for (int idx = tid; idx < L * nbw; idx += NT) cache[idx] = G[(size_t)(idx / nbw) * W + (idx % nbw)];
    const uint32_t* rows = cached ? cache : G;
    const int stride = cached ? nbw : W;
    unsigned short* chain = chains + (size_t)warp * Lc;
    const int ub = max_core + 1;  // src/graph.cc:84-86
    int mc = 0, i = L - 1;
    __syncthreads();
    while (i >= 0 && mc < ub) {
      // warp k tries the start vertex at peel position i - k against the current incumbent size mc
      const int my_i = i - warp;
      int sz = -1;  // -1: kcore[v] <= mc here, hence for every later start vertex too (kcore is non-increasing along the
                    //     reversed peel order and mc only grows)
      if (my_i >= 0) {
        const int v = ko[my_i];
        if (kc[v] > mc) {
          // candidates: neighbours of v with kcore > mc  <=>  rank >= kb[mc]  (first rank whose core >= mc)
          const int thr = kb[min(mc, max_core + 1)];
          const int rv = ro[v];
          if (nwl == 1) sz = clique_descent<1>(rows, stride, nbw, rv, thr, mc, chain);
          else if (nwl <= 4) sz = clique_descent<4>(rows, stride, nbw, rv, thr, mc, chain);
          else sz = clique_descent<8>(rows, stride, nbw, rv, thr, mc, chain);
        }
      }
      if (lane == 0) s_sz[warp] = sz;
      __syncthreads();
      int win = -1;
      bool stop = false;
#pragma unroll
      for (int k = 0; k < kCliqueWarps; ++k) {
        if (win < 0 && !stop) {
          const int s = s_sz[k];
          if (s < 0) stop = true;
          else if (s > mc) win = k;
        }
      }
      if (win >= 0) {  // commit the first improvement in sequential order; later warps saw a stale mc and are redone
        const int wsz = s_sz[win];
        for (int w = tid; w < W; w += NT) idbits[w] = 0;
        __syncthreads();
        if (warp == win) {
          for (int t = lane; t < wsz - 1; t += 32) {
            const int id = br[chain[t]];
            atomicOr(&idbits[id >> 5], 1u << (id & 31));
          }
          if (lane == 0) {
            const int v = ko[i - win];
            atomicOr(&idbits[v >> 5], 1u << (v & 31));
          }
        }
        mc = wsz;
        i = i - win - 1;
      } else if (stop) {
        i = -1;
      } else {
        i -= kCliqueWarps;
      }
      __syncthreads();
    }
  }
  __syncthreads();
  // ascending ids (std::sort(max_clique_), quatro.hpp:806)
  int csize = 0;
  for (int base = 0; base < W; base += NT) {
    const int w = base + tid;
    uint32_t x = w < W ? idbits[w] : 0;
    int tot;
    int off = csize + block_excl_scan(__popc(x), s_scan, &tot);
    while (x) {
      const int b = __ffs(x) - 1;
      x &= x - 1;
      out[off++] = w * 32 + b;
    }
    csize += tot;
  }
  if (tid == 0) n_clique[pair] = csize;
}

template <int NW, typename Acc>
static size_t kcore_smem_bytes(int Lc, int W) {
  return (size_t)(Lc + 2) * sizeof(Acc) + (size_t)(Lc + 2) * sizeof(int) + (size_t)W * sizeof(uint32_t) + (size_t)5 * Lc * sizeof(unsigned short) +
         (size_t)2 * Lc;
}

int launch_clique(qb200_handle* h, int n_pairs, int mode, double kcore_thr) {
  if (n_pairs <= 0) return QB200_OK;
  if (mode == QB200_PMC_EXACT) return QB200_ERR_UNSUPPORTED;
  const int Lc = h->Lc, W = h->W;
  // shared-memory adjacency cache of the descent: 14336 words (56 KB) hold graphs up to L ~ 660
  const int cache_words = 14336;
  const bool wide = W > 128;
  const size_t sm_kcore = wide ? kcore_smem_bytes<8, unsigned long long>(Lc, W) : kcore_smem_bytes<4, uint32_t>(Lc, W);
  const size_t sm_clique = (size_t)kCliqueWarps * Lc * sizeof(unsigned short) + (size_t)W * sizeof(uint32_t) + (size_t)cache_words * 4;
  if (!(h->func_attr_set & 1u)) {  // per handle: the opt-in is a per-device property of the function
    if (wide)
      QB_CUDA_TRY(h, cudaFuncSetAttribute(kcore_cta_kernel<8, unsigned long long>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_kcore));
    else
      QB_CUDA_TRY(h, cudaFuncSetAttribute(kcore_cta_kernel<4, uint32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_kcore));
    QB_CUDA_TRY(h, cudaFuncSetAttribute(clique_cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_clique));
    h->func_attr_set |= 1u;
  }
  if (wide)
    kcore_cta_kernel<8, unsigned long long><<<n_pairs, 256, sm_kcore, h->stream>>>(h->adj, h->deg, h->ctr.n_corr, Lc, W, h->kcore, h->korder,
                                                                                   h->rank_of, h->by_rank, h->kbin, h->ctr.max_core);
  else
    kcore_cta_kernel<4, uint32_t><<<n_pairs, 128, sm_kcore, h->stream>>>(h->adj, h->deg, h->ctr.n_corr, Lc, W, h->kcore, h->korder, h->rank_of,
                                                                         h->by_rank, h->kbin, h->ctr.max_core);
  const dim3 gp((Lc + 7) / 8, n_pairs);
  permute_adj_kernel<<<gp, 256, 8 * W * sizeof(uint32_t), h->stream>>>(h->adj, h->ctr.n_corr, Lc, W, h->rank_of, h->adjp);
  clique_cta_kernel<<<n_pairs, kCliqueWarps * 32, sm_clique, h->stream>>>(h->adjp, h->ctr.n_corr, Lc, W, h->kcore, h->korder, h->rank_of, h->by_rank,
                                                                          h->kbin, h->ctr.max_core, mode, kcore_thr, cache_words, h->clique,
                                                                          h->ctr.n_clique);
  h->launches += 3;
  QB_CUDA_TRY(h, cudaGetLastError());
  return QB200_OK;
}

}  // namespace qb
