// pgbart_predict.cuh — posterior prediction from the forest HISTORY.
//
// Replaces what the reference gets from bartrs' native `PosteriorSampler` (pymc_bart/utils.py:60-71,93-107,124-127:
// `from_history(batches, baseline_forest, m, n_outputs)` / `sample_posterior(X, draw_indices, excluded)`), for all
// chains of an op at once.
//
// Device layout (built once per op by bk_history_* from the `(baseline_forest, batches)` entries of op.all_trees):
//   nodes    bk_node[total]          every tree VERSION ever stored, compacted to its n_nodes (24 B per node)
//   ver_off  int32[n_versions + 1]   first node of version v
//   ver_tbl  int32[n_forests][m]     version id of tree t in forest f; forests are (chain, draw, output group) in that
//                                    order, so a global draw index d of `_MultiChainSampler` (utils.py:74-107) is forest
//                                    row d * n_outputs + g: the draw -> chain routing is the identity on this table
// A draw costs m ints here instead of a dense [m][255] copy of the forest (1.2 MB per draw at m = 200).
//
// Kernel: one CTA per (row tile, selected forest, exclusion mask).  The forest's trees are staged into shared memory
// (compact, usually m * ~7 nodes); every thread then walks all m trees for its rows and accumulates the leaf values in
// double in tree order (same order as oracle/pgbart_oracle.c: bko_predict).  With an exclusion mask the descent is
// weighted: at a split on an excluded variable both children are visited with weights n_left/n and 1 - n_left/n
// (SURVEY.md App. A.10); the explicit stack holds depth + 2 <= 130 entries (255-node trees), so it cannot overflow.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "bk_spec.h"
#include "pgbart_b200.h"

#define BKP_THREADS 256
#define BKP_ROWS_PER_THREAD 4
#define BKP_TILE (BKP_THREADS * BKP_ROWS_PER_THREAD)
#define BKP_SMEM_NODES 4096      // at most 96 KB of staged nodes; larger forests are read through L1/L2
#define BKP_MAX_TREES_SMEM 2048  // tree offsets kept in shared memory
#define BKP_STACK 130

struct PredictArgs {
  const bk_node* nodes;
  const int32_t* ver_off;
  const int32_t* ver_tbl;
  int m;
  const float* X;   // [n][p] row-major
  int n, p;
  const int32_t* sel;   // forest rows: [n_sel] shared by all masks (sel_stride 0) or [n_masks][n_sel] (sel_stride n_sel)
  int n_sel, sel_stride;
  const uint8_t* excl;  // [n_masks][p] or nullptr
  int n_masks;
  const int32_t* rules;  // [p] or nullptr
  const float* vals;     // nullptr, or shared-tree multi-output: leaf values of every node [total][K]
  int K;                 // outputs per forest (1 without vals)
  float* out;            // [n_masks][n_sel][K][n]
  int smem_nodes;        // nodes of shared memory available for staging (host: min(largest forest, BKP_SMEM_NODES))
  int32_t* err;          // device flag: 1 = stack overflow (cannot happen for trees of <= 255 nodes), 2 = bad version id
};

// acc[j] += value_j of the tree at x (K = 1: the node's own value; K > 1: vals[(node0 + k) * K + j])
template <bool EXCL>
__device__ __forceinline__ void bkp_tree_value(const bk_node* __restrict__ nodes, const float* __restrict__ x,
                                               const uint8_t* __restrict__ excl, const int32_t* __restrict__ rules, int32_t* err,
                                               const float* __restrict__ vals, int K, int node0, double* __restrict__ acc) {
  if (!EXCL) {
    int k = 0;
    for (;;) {
      const bk_node nd = nodes[k];
      if (nd.var < 0) {
        if (!vals) acc[0] = BK_DADD(acc[0], (double)nd.value);
        else for (int j = 0; j < K; ++j) acc[j] = BK_DADD(acc[j], (double)vals[(size_t)(node0 + k) * K + j]);
        return;
      }
      const float xv = x[nd.var];
      const int rule = rules ? rules[nd.var] : BK_RULE_CONTINUOUS;
      const bool left = rule == BK_RULE_SUBSET ? (bk_subset_left(xv, nd.split) != 0) : (rule == BK_RULE_ONEHOT ? (xv == nd.split) : (xv <= nd.split));
      k = left ? nd.left : nd.left + 1;
    }
  } else {
    // post-order evaluation with an explicit stack of (node, weight); the right child is pushed first so the left one
    // is visited first (the oracle's order)
    int sn[BKP_STACK];
    double sw[BKP_STACK];
    int sp = 1;
    sn[0] = 0; sw[0] = 1.0;
    double tv[BK_MAX_OUTPUTS];
    for (int j = 0; j < K; ++j) tv[j] = 0.0;
    while (sp > 0) {
      --sp;
      const int k = sn[sp];
      const double w = sw[sp];
      const bk_node nd = nodes[k];
      if (nd.var < 0) {
        if (!vals) tv[0] = BK_DFMA(w, (double)nd.value, tv[0]);
        else for (int j = 0; j < K; ++j) tv[j] = BK_DFMA(w, (double)vals[(size_t)(node0 + k) * K + j], tv[j]);
        continue;
      }
      const int l = nd.left, r = nd.left + 1;
      if (excl[nd.var]) {
        const double tot = (double)nodes[l].n + (double)nodes[r].n;
        if (!(tot > 0.0)) continue;
        if (sp + 2 > BKP_STACK) { *err = 1; continue; }
        const double wl = BK_DDIV((double)nodes[l].n, tot);
        const double wr = BK_DSUB(1.0, wl);
        sn[sp] = r; sw[sp] = BK_DMUL(w, wr); ++sp;
        sn[sp] = l; sw[sp] = BK_DMUL(w, wl); ++sp;
      } else {
        const float xv = x[nd.var];
        const int rule = rules ? rules[nd.var] : BK_RULE_CONTINUOUS;
        const bool left = rule == BK_RULE_SUBSET ? (bk_subset_left(xv, nd.split) != 0) : (rule == BK_RULE_ONEHOT ? (xv == nd.split) : (xv <= nd.split));
        sn[sp] = left ? l : r; sw[sp] = w; ++sp;
      }
    }
    for (int j = 0; j < K; ++j) acc[j] = BK_DADD(acc[j], tv[j]);
  }
}

template <bool EXCL>
__global__ void __launch_bounds__(BKP_THREADS) pgbart_predict_hist_kernel(const PredictArgs A) {
  extern __shared__ __align__(16) unsigned char bkp_smem[];
  int32_t* s_off = reinterpret_cast<int32_t*>(bkp_smem);                                   // [m_s + 1] offsets into s_nodes
  const int m_s = A.m <= BKP_MAX_TREES_SMEM ? A.m : 0;
  bk_node* s_nodes = reinterpret_cast<bk_node*>(bkp_smem + (((size_t)(m_s + 1) * 4 + 15) & ~(size_t)15));
  __shared__ int s_total, s_warp[BKP_THREADS / 32];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int f = A.sel[(size_t)blockIdx.z * A.sel_stride + blockIdx.y];
  const int32_t* vrow = A.ver_tbl + (size_t)f * A.m;
  const uint8_t* excl = EXCL ? A.excl + (size_t)blockIdx.z * A.p : nullptr;

  // ---- stage the forest: block-wide exclusive scan of the trees' node counts, then a flat copy
  bool staged = m_s > 0;
  if (staged) {
    int run = 0;   // (uniform) nodes before the current chunk of BKP_THREADS trees
    for (int t0 = 0; t0 < A.m; t0 += BKP_THREADS) {
      const int t = t0 + tid;
      int nn = 0;
      if (t < A.m) { const int v = vrow[t]; nn = A.ver_off[v + 1] - A.ver_off[v]; }
      int incl = nn;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int nb = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += nb; }
      if (lane == 31) s_warp[wid] = incl;
      __syncthreads();
      int before = 0, chunk = 0;
#pragma unroll
      for (int k = 0; k < BKP_THREADS / 32; ++k) { const int c = s_warp[k]; if (k < wid) before += c; chunk += c; }
      if (t < A.m) s_off[t] = run + before + incl - nn;
      run += chunk;
      __syncthreads();
    }
    if (tid == 0) { s_off[A.m] = run; s_total = run; }
    __syncthreads();
    staged = s_total <= A.smem_nodes;
    if (staged) {
      // one warp per tree: 24-byte nodes copied as 8-byte words
      for (int t = wid; t < A.m; t += BKP_THREADS / 32) {
        const int v = vrow[t];
        const int src = A.ver_off[v], nn = A.ver_off[v + 1] - src;
        const uint2* g = reinterpret_cast<const uint2*>(A.nodes + src);
        uint2* d = reinterpret_cast<uint2*>(s_nodes + s_off[t]);
        for (int i = lane; i < nn * 3; i += 32) d[i] = __ldg(g + i);
      }
      __syncthreads();
    }
  }

  const int K = A.vals ? A.K : 1;
  const size_t out_base = ((size_t)blockIdx.z * A.n_sel + blockIdx.y) * (size_t)K * (size_t)A.n;
#pragma unroll 1
  for (int r = 0; r < BKP_ROWS_PER_THREAD; ++r) {
    const long long i = (long long)blockIdx.x * BKP_TILE + (long long)r * BKP_THREADS + tid;
    if (i >= A.n) break;
    const float* x = A.X + (size_t)i * A.p;
    double acc[BK_MAX_OUTPUTS];
    for (int j = 0; j < K; ++j) acc[j] = 0.0;
    for (int t = 0; t < A.m; ++t) {
      const int node0 = A.ver_off[vrow[t]];
      const bk_node* nodes = staged ? s_nodes + s_off[t] : A.nodes + node0;
      bkp_tree_value<EXCL>(nodes, x, excl, A.rules, A.err, A.vals, K, node0, acc);
    }
    for (int j = 0; j < K; ++j) A.out[out_base + (size_t)j * A.n + (size_t)i] = (float)acc[j];
  }
}

// Squared Pearson correlation between the full-model predictions and every excluded-subset prediction, per posterior
// sample (pymc_bart/utils.py:1339-1346 `pearsonr2`, called per sample at :1003-1005 and :1038-1040): one CTA per
// (sample j, subset s) reduces the five moments over the n * n_outputs values in double.
//   a: [n_samples][len]   b: [n_subsets][n_samples][len]   out: [n_subsets][n_samples] double
__global__ void __launch_bounds__(256) pgbart_pearson_r2_kernel(const float* __restrict__ a, const float* __restrict__ b, int len,
                                                                 int n_samples, double* __restrict__ out) {
  const int j = blockIdx.x, s = blockIdx.y;
  const float* pa = a + (size_t)j * len;
  const float* pb = b + ((size_t)s * n_samples + j) * len;
  // two passes (means first) like the reference, so that the centred sums do not cancel
  __shared__ double sh[5][8];
  __shared__ double s_ma, s_mb;
  double sa = 0.0, sb = 0.0;
  for (int i = threadIdx.x; i < len; i += 256) { sa += (double)pa[i]; sb += (double)pb[i]; }
  for (int o = 16; o > 0; o >>= 1) { sa += __shfl_xor_sync(0xffffffffu, sa, o); sb += __shfl_xor_sync(0xffffffffu, sb, o); }
  if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = sa; sh[1][threadIdx.x >> 5] = sb; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ta = 0.0, tb = 0.0;
    for (int k = 0; k < 8; ++k) { ta += sh[0][k]; tb += sh[1][k]; }
    s_ma = ta / (double)len; s_mb = tb / (double)len;
  }
  __syncthreads();
  const double ma = s_ma, mb = s_mb;
  double ab = 0.0, aa = 0.0, bb = 0.0;
  for (int i = threadIdx.x; i < len; i += 256) {
    const double da = (double)pa[i] - ma, db = (double)pb[i] - mb;
    ab += da * db; aa += da * da; bb += db * db;
  }
  for (int o = 16; o > 0; o >>= 1) {
    ab += __shfl_xor_sync(0xffffffffu, ab, o); aa += __shfl_xor_sync(0xffffffffu, aa, o); bb += __shfl_xor_sync(0xffffffffu, bb, o);
  }
  if ((threadIdx.x & 31) == 0) { sh[2][threadIdx.x >> 5] = ab; sh[3][threadIdx.x >> 5] = aa; sh[4][threadIdx.x >> 5] = bb; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double tab = 0.0, taa = 0.0, tbb = 0.0;
    for (int k = 0; k < 8; ++k) { tab += sh[2][k]; taa += sh[3][k]; tbb += sh[4][k]; }
    out[(size_t)s * n_samples + j] = (tab * tab) / (taa * tbb);
  }
}
