// pgbart_b200.cu — B200 (sm_100a) PGBART step: one persistent cooperative kernel
// per step + the C ABI of include/pgbart_b200.h.
//
// Hot path (pymc-bart's PGBART.astep; reference sites cited in include/pgbart_b200.h
// and SURVEY.md §8a rows B1-B10).  One launch runs the whole step for every chain
// batched on this GPU as a per-chain DATAFLOW (no grid-wide barrier): CTA c (c < chains) is
// chain c's control CTA, every other CTA is a worker CTA split into BK_NGROUPS independent groups.
// A control CTA publishes an "epoch" = one batch of data units (a 16-byte descriptor behind a
// release fence); the groups serving that chain split the epoch statically, run their units and
// release-add a done counter the control CTA polls.  Chains are independent: group g serves chain
// g mod BK_NGROUPS, so the epochs of different chains overlap on every SM, and one chain's scalar
// control overlaps the other chains' streaming work.
//
//   CONTROL  (one CTA per chain, 16 warps; O(P) per round): leaf values, log weights, fixed-point
//            systematic resampling, queue pops, split-variable and split-index draws, k-th-member
//            selection (two-level search over per-tile counts), row allocation, job descriptors.
//            Random draws and the particle copy run in the SHADOW of the epoch just published.
//   DATA     (worker groups; the O(P*N) streams):
//     ROUND  one warp walks 256 rows x its jobs: the tile's fixed-point residual/sum-of-trees stay
//            in registers, per job it reads 8 leaf ids/lane (one 64-bit load) and the split column
//            (two 128-bit loads), routes members left/right with byte-parallel logic, writes the
//            new leaf-id row, forms the left child's (n, sum q_st, sum q_r, sum q_r^2) as REDUX
//            partial sums = 32-bit limbs added to shared memory with ONE atomic instruction, and
//            counts the members of the particle's next queue node per tile (feeds the selection).
//     LL     Bernoulli likelihood: per-row log-likelihood terms of the rows of freshly made leaves.
//     SWEEP  fused commit of tree t (sum_trees = noi + new prediction, leaf-id row, Welford running
//            sd) and prologue of tree t+1 (residual, fixed-point copies, per-leaf statistics).
//
// All reductions are integer (bk_spec.h), so results do not depend on the grid size, the
// reduction order, or the number of GPUs; every float op is an explicit round-to-nearest
// intrinsic in a fixed order.  No tensor cores: there is no dense contraction on this path
// (HBM/L2-bound integer and compare work).
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <new>

#include "pgbart_device.cuh"

// Bring-up aid: per-warp progress markers in device memory (compile with -DBK_DEBUG_MARKS and
// run with BK_DEBUG_MARKERS=1; tests/gpu_debug.py dumps them when a step hangs).
#ifdef BK_DEBUG_MARKS
#define MARK(code)                                                                                  \
  do {                                                                                              \
    if (P.marker && (threadIdx.x & 31) == 0)                                                        \
      ((volatile int*)P.marker)[blockIdx.x * 33 + (threadIdx.x >> 5)] = (code) | (phase << 12);    \
  } while (0)
#else
#define MARK(code) do { } while (0)
#endif

// Block barrier = PTX `barrier.sync` WITHOUT .aligned.  The control phase runs long scalar
// sections in thread 0 (weights, resampling, row allocation) while lanes 1..31 of warp 0 idle;
// `bar.sync`/__syncthreads() is the .aligned form, which requires every warp to reach the SAME
// barrier instruction converged — with this kernel's control flow ptxas emitted paths where lane 0
// arrived apart from its warp, the hardware counted warp 0 twice and the block dead-locked
// (compute-sanitizer synccheck: "Divergent thread(s) in warp" at a __syncthreads).  The non-aligned
// form has per-thread arrival semantics and is legal under intra-warp divergence.
#define BLOCK_SYNC() asm volatile("barrier.sync 0;" ::: "memory")
// The control CTA of a chain runs with its first 256 threads only (the other warps exit at once):
// barrier 1 with an explicit count, so a control-stage barrier waits for 8 warps, not 32.
#define BK_CTRL_THREADS 512
// Warp 0 of a control CTA is the SCALAR warp: thread 0 runs the sequential sections, lanes 1..31 only meet barriers.
// (Lane 0's long solo sections leave that warp split across the non-aligned barriers; a split warp executes any
// shared work twice and every *_sync collective through the WARPSYNC slow path, and the CTA waits for it.)  All
// per-particle / per-row / strided work therefore runs on threads 32.. : BK_WTID is the index among those.
#define BK_WTID ((int)threadIdx.x - 32)
#define BK_WTHREADS (BK_CTRL_THREADS - 32)
#define CTRL_SYNC() asm volatile("barrier.sync 1, 512;" ::: "memory")
// worker group g uses barrier 2+g with BK_GROUP_THREADS arrivals
#define GROUP_SYNC(g) asm volatile("barrier.sync %0, %1;" ::"r"(2 + (g)), "n"(BK_GROUP_THREADS) : "memory")
#ifndef BK_PROFILE_CTRL
#define TSUB(i) do { } while (0)
#else
#define TSUB(i)                                                              \
  do {                                                                       \
    if (threadIdx.x == 0) {                                                  \
      unsigned long long now_ = globaltimer_ns();                            \
      hot->t_sub[i] += now_ - hot->t_sub_last;                               \
      hot->t_sub_last = now_;                                                \
    }                                                                        \
  } while (0)
#endif


// Every device function that takes `const Params& P` is forced inline: the structure is a kernel parameter (constant
// bank); a real call would make each thread copy its 350 bytes to the stack and read them back from local memory.
// ------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ void red_release_add_u32(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_volatile_i32(const int* p) {
  int v;
  asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ void red_add_u64(unsigned long long* p, unsigned long long v) {
  atomicAdd(p, v);  // result unused -> RED.E.ADD.64
}

// ~4 s at 1.9 GHz: a barrier that is not reached means a bug; every CTA bails out
#define BK_BARRIER_TIMEOUT_CYCLES (8000000000LL)

#ifdef BK_PROFILE_CTRL
// worker-side latency sums (ns) per CTA (group 0): [0] publish->epoch seen, [2] ->units done, [3] ->done added, [4] epochs
__device__ unsigned long long g_wdbg[256][16];
// cycle stamp that waits for `dep` (a freshly loaded / computed register)
// (the clock read is predicated on `dep`, so it cannot issue before the value has arrived)
#define UTICK(dep) ({ long long t_ = 0; asm volatile("{ .reg .pred p; setp.ne.s32 p, %1, 0x7fffffff; @p mov.u64 %0, %%clock64; }" : "+l"(t_) : "r"((int)(dep)) : "memory"); t_; })
#define UACC(i, v) do { if ((threadIdx.x & 31) == 0) atomicAdd(&g_wdbg[blockIdx.x][i], (unsigned long long)(v)); } while (0)
// control-side cycle split (thread 0 of chain 0's control CTA): CT0 starts a section, CT(i, dep) closes part i
// (accumulated in shared memory, flushed to g_cdbg once per step, so that a stamp costs a few cycles)
__device__ unsigned long long g_cdbg[32];
__shared__ unsigned long long s_cdbg[32];
#define CT0() long long ct_last_ = clock64()
#define CT(i, dep) do { if (threadIdx.x == 32) { const long long n_ = UTICK(dep); s_cdbg[i] += (unsigned long long)(n_ - ct_last_); ct_last_ = n_; } } while (0)
#define WDBG(i, t0)                                                              \
  do { if (threadIdx.x == 0) g_wdbg[blockIdx.x][i] += globaltimer_ns() - (t0); } while (0)
// running cycle stamps of thread 32 (lane 0 of the first particle warp) across the round path: CTS(i, dep) adds the
// cycles since the previous stamp to slot i once `dep` is available
__shared__ long long s_ct_last;
#define CTS(i, dep) do { if (threadIdx.x == 32) { const long long n_ = UTICK(dep); s_cdbg[i] += (unsigned long long)(n_ - s_ct_last); s_ct_last = n_; } } while (0)
#define CTN(i) do { if (threadIdx.x == 32) s_cdbg[i] += 1ull; } while (0)
#else
#define CTS(i, dep) do { } while (0)
#define CTN(i) do { } while (0)
#define WDBG(i, t0) do { } while (0)
#define UTICK(dep) 0ll
#define UACC(i, v) do { } while (0)
#define CT0() do { } while (0)
#define CT(i, dep) do { } while (0)
#endif

// a node with fewer than N / BK_SPARSE_DIV members loads the split column only in lanes that hold members
// -DBK_NO_MISSING compiles the missing-covariate paths out (A/B of their cost on data without NaNs; tests/gpu_variants.sh)
// The step kernel exists in three MODES, chosen once per CTA at kernel entry and carried as a template argument, so that
// the plain path never executes, fetches or allocates registers for the other two (compiled into one body, the
// multi-output and missing-value code cost the plain C2 / C5 steps 8 %: twice the code for the instruction cache, and
// calls in the hot loops):
//   BK_MODE_PLAIN    single output per forest, no NaN in X
//   BK_MODE_MISSING  X holds NaNs (Params::has_nan) and / or some column uses the SubsetSplit rule (Params::n_subset)
//   BK_MODE_MULTI    shared-tree multi-output (Params::K > 1)
#define BK_MODE_PLAIN 0
#define BK_MODE_MISSING 1
#define BK_MODE_MULTI 2
#define BK_IS_MULTI(P) (MODE == BK_MODE_MULTI)
#define BK_MISSING_ENABLED (MODE == BK_MODE_MISSING)
#ifndef BK_SPARSE_DIV
#define BK_SPARSE_DIV 8
#endif
#define BK_CUM_SMEM 1024
// per-launch arguments passed by value (no H2D copy, no memset on the step's stream)
struct StepArgs {
  float sigma[64];       // likelihood scale of every (chain, group)
  unsigned epoch_base;   // epoch ids of this launch start above it
  int n_steps;           // steps of every chain in this launch (chains do not wait for each other between steps)
};
__device__ __forceinline__ bk_step_stats* step_stats(const Params& P, int step, int c) {
  return reinterpret_cast<bk_step_stats*>(reinterpret_cast<char*>(P.stats) + (size_t)step * P.rec_stride) + c;
}
__device__ __forceinline__ int32_t* step_vi(const Params& P, int step, int c) {
  return reinterpret_cast<int32_t*>(reinterpret_cast<char*>(P.vi) + (size_t)step * P.rec_stride) + (size_t)c * P.p;
}
struct CtlShared {
  double lw[BK_MAX_PARTICLES];
  int anc[BK_MAX_PARTICLES];
  int s_kind[BK_MAX_PARTICLES];   // 0 nothing, 1 grow, 2 only needs a count
  int s_j[BK_MAX_PARTICLES];
  int s_v[BK_MAX_PARTICLES];
  unsigned s_k[BK_MAX_PARTICLES];
  int s_row[BK_MAX_PARTICLES];
  float s_split[BK_MAX_PARTICLES];
  int row_cnt_node[2 * BK_MAX_PARTICLES];  // node whose per-tile counts a pool row holds (persists across phases)
  Job jobs[BK_MAX_PARTICLES];               // staged here, copied to global by the whole CTA
  unsigned long long cum_s[BK_MAX_PARTICLES];   // running sums of the fixed-point weights
  // Random draws computed in the SHADOW of a data epoch (while the workers stream): Philox counters do not depend on
  // the particle state, so the next round's proposal draws, this round's leaf-value normals and resampling uniform
  // are ready when the epoch completes.  Tagged with (tree, round); a consumer that finds another tag computes inline.
  double pre_u1[BK_MAX_PARTICLES];
  int pre_v[BK_MAX_PARTICLES];
  uint32_t pre_u3[BK_MAX_PARTICLES];
  double pre_zl[BK_MAX_PARTICLES], pre_zr[BK_MAX_PARTICLES];
  uint32_t pre_ures;
  int pre_prop_tree, pre_prop_round;   // tag of pre_u1 / pre_v / pre_u3
  int pre_z_tree, pre_z_round;         // tag of pre_zl / pre_zr / pre_ures
  // Deferred particle copy: after resampling, open_round() reads every slot's state THROUGH src_slot (its ancestor in
  // the current buffer) and only records the new queue head; the copy into the other buffer happens in the shadow
  // of the epoch it published (apply_pending_copy).
  int copy_pending;
  int src_slot[BK_MAX_PARTICLES];
  int s_qh[BK_MAX_PARTICLES];
  double cum_prior[BK_CUM_SMEM];   // normalised cumulative split prior (first BK_CUM_SMEM columns)
  double p_leaf[64];               // depth prior table, depths 0..63 (deeper: global)
  signed char rules[BK_CUM_SMEM];  // split rule of the first BK_CUM_SMEM columns
  int live;
  int win;
  unsigned pick;
  // round path (open_round): job index of every slot, used-row bitmap, per-warp grower counts, count-job cursor,
  // selection cursor, error bits, slices of the resampling scan
  int s_jobidx[BK_MAX_PARTICLES];
  unsigned used_rows[8];
  int warp_grow[4];
  int n_cnt_jobs;
  int next_sel;
  int err_bits;
  unsigned long long r_slice[4];
  int warp_grow_root[4];
  float job_vals[BK_MAX_PARTICLES][2][BK_MAX_OUTPUTS];   // shared-tree multi-output: leaf values of the LL jobs (staged, copied to global)
  int n_nan_fail, n_fail_root;   // slots whose split value could not be drawn (only members with a missing covariate)
  // last value read from every accumulator word of accL: the workers only ever ADD (RED), the control CTA takes
  // differences (exact in wrapping 64-bit arithmetic), so no accumulator is ever zeroed by a store that would have
  // to be ordered before the next epoch's adds
  unsigned long long acc_prev[BK_MAX_PARTICLES][BK_ACC_STRIDE];
};

// A worker CTA is split into BK_NGROUPS independent groups of BK_GROUP_THREADS threads.  Each group serves its own
// chain(s) with its own poller, barrier and shared-memory area, so the epochs of different chains run CONCURRENTLY on
// every SM (one chain's unit is a ~500-instruction dependent stream; several chains interleaved hide that latency).
struct Work {
  int chain;       // -1: nothing new
  int cmd, njobs, total;
  unsigned lo, hi; // this group's range of (tile, job) pairs / first sweep tile and stride
  int exit_now;
};

struct GroupShared {
  // partial sums of a ROUND / LL epoch: warps add here (shared-memory atomics), the group then issues ONE global
  // atomic per (job, statistic) — the L2 sees groups x jobs x 5 atomics per epoch instead of tiles x jobs x 5 on a
  // handful of lines
  unsigned acc[BK_MAX_PARTICLES * 16];   // [job][BK_LIMBS] 32-bit limbs, see round_unit
  Job jobs[BK_MAX_PARTICLES];   // the epoch's job list, one copy per group
  float old_vals[256];
  float new_vals[256];
  float pro_vals[256];
  unsigned long long leaf_acc[256];      // per-leaf sum of the old tree's q(r) (Bernoulli: log-likelihood terms)
  unsigned long long tot_acc[8];
  unsigned long long sd_acc[8];          // K > 1: running-sd sums per output of a SWEEP
  // shared-tree multi-output scratch, one use per epoch kind (ROUND / LL / SWEEP): zeroed / staged at the start of each
  union {
    unsigned acck[BK_MAX_PARTICLES][BK_MAX_OUTPUTS][4];      // ROUND: (lo, hi) limbs of sum q(sum_trees[j]) for the left / right child
    float job_vals[BK_MAX_PARTICLES][2][BK_MAX_OUTPUTS];     // LL: leaf values of the two new leaves of every job
    float pro_vals_k[BK_MAX_OUTPUTS][256];                   // SWEEP: leaf values of the tree the prologue removes
  } u;
  Work work;
  unsigned seen[64];            // last epoch of each chain this group has executed
  unsigned char fin[64];
};
static_assert(sizeof(GroupShared) * BK_NGROUPS <= 160 * 1024, "worker groups must fit the dynamic shared memory of the launch");

struct __align__(16) KernelShared {   // static shared memory of a control CTA (workers use the dynamic area)
  CtlShared ctl;
};

// ------------------------------------------------------------------ dataflow sync helpers
__device__ __forceinline__ uint4 ld_acquire_v4(const uint4* p) {
  uint4 v;
  asm volatile("ld.acquire.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint4 ld_volatile_v4(const uint4* p) {
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_volatile_v4(uint4* p, uint4 v) {
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__device__ __forceinline__ int4 ld_ca_v4(const int4* p) {   // ordinary (weak, L1-cached) 16-byte load, never the .nc path
  int4 v;
  asm volatile("ld.global.ca.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ int ld_ca_s32(const int* p) {   // ordinary (weak, L1-cached) 4-byte load, never the .nc path
  int v;
  asm volatile("ld.global.ca.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// ------------------------------------------------------------------ small device utils
// Particle states of the chain a control CTA owns live in its dynamic shared memory: header + the
// first P.fastF nodes of every particle (both ping-pong buffers); nodes beyond fastF (rare deep trees)
// overflow to the global `parts` array.  All control-phase particle traffic is then ~30-cycle shared
// memory instead of ~600-cycle L2 round trips.
extern __shared__ __align__(16) unsigned char bk_dyn_smem[];
struct PHdr { int32_t n_nodes, q_head, row, pad; double gain, lw; };   // gain: Gaussian sum of bk_leaf_gain over the leaves; Bernoulli: integer log-likelihood sum
static_assert(sizeof(PHdr) == 32, "PHdr layout");
struct PRef {
  PHdr* h;
  DNode* fast;
  DNode* slow;
  int F;
  __device__ __forceinline__ DNode& node(int k) const { return k < F ? fast[k] : slow[k]; }
};
__device__ __forceinline__ PRef pref(const Params& P, int c, int buf, int q) {
  unsigned char* base = bk_dyn_smem + ((size_t)buf * P.P + q) * P.fast_stride;
  PRef r;
  r.h = reinterpret_cast<PHdr*>(base);
  r.fast = reinterpret_cast<DNode*>(base + sizeof(PHdr));
  r.slow = (P.parts + ((size_t)c * 2 + buf) * P.P + q)->nodes;
  r.F = P.fastF;
  return r;
}
__device__ __forceinline__ bk_stats node_stats(const DNode& nd) {
  bk_stats s; s.n = nd.n; s.sst = nd.sst; s.sr = nd.sr; return s;
}
__device__ __forceinline__ void set_node_stats(DNode& nd, const bk_stats& s) {
  nd.n = s.n; nd.sst = s.sst; nd.sr = s.sr;
}
__device__ __forceinline__ bk_trace_rec* trace_at(const Params& P, int c, int pos) {
  if (P.trace_cap <= 0 || pos < 0 || pos >= P.trace_cap) return nullptr;
  return P.trace + (size_t)c * P.trace_cap + pos;
}

__device__ __forceinline__ unsigned warp_incl_scan_u32(unsigned v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const unsigned nb = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += nb; }
  return v;
}
// position of the k-th unit in an array of counts: returns the entry index and leaves the remaining offset in `off`
// (-1 when the counts hold fewer than off + 1 units).  n4 = entries / 4 (the arrays are padded to 16 bytes with zeros).
__device__ __forceinline__ int find_in_counts(const uint4* __restrict__ cnt4, int n4, unsigned& off, int lane) {
  const int per = (n4 + 31) >> 5;                   // 4-entry groups per lane
  int idx = -1;
  if (per <= 4) {
    // a lane keeps its (at most 16) counts in registers, so the lane that owns the k-th unit walks them itself —
    // one L2 round trip for the whole search
    uint4 v[4];
    unsigned sum = 0u;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i4 = lane * per + u;
      v[u] = (u < per && i4 < n4) ? __ldcg(cnt4 + i4) : make_uint4(0u, 0u, 0u, 0u);
      sum += v[u].x + v[u].y + v[u].z + v[u].w;
    }
    const unsigned incl = warp_incl_scan_u32(sum, lane), excl = incl - sum;
    const unsigned b = __ballot_sync(0xffffffffu, off >= excl && off < incl);
    if (b != 0u) {
      const int src = __ffs(b) - 1;
      unsigned o2 = off - excl;     // valid in lane src
      int t = 0;
      bool found = false;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const unsigned cs[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if (!found) { if (o2 < cs[e]) { found = true; t = u * 4 + e; } else o2 -= cs[e]; }
        }
      }
      idx = __shfl_sync(0xffffffffu, src * per * 4 + t, src);
      off = __shfl_sync(0xffffffffu, o2, src);
    }
  } else {
    unsigned sum = 0u;
    for (int j0 = 0; j0 < per; j0 += 8) {   // eight independent 128-bit loads in flight per lane
      uint4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i4 = lane * per + j0 + u;
        v[u] = (j0 + u < per && i4 < n4) ? __ldcg(cnt4 + i4) : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) sum += v[u].x + v[u].y + v[u].z + v[u].w;
    }
    const unsigned incl = warp_incl_scan_u32(sum, lane), excl = incl - sum;
    const unsigned b = __ballot_sync(0xffffffffu, off >= excl && off < incl);
    if (b != 0u) {
      const int src = __ffs(b) - 1;
      off -= __shfl_sync(0xffffffffu, excl, src);
      // second level: the `per` groups of lane src's chunk, 32 at a time
      for (int j0 = 0; j0 < per && idx < 0; j0 += 32) {
        const int i4 = src * per + j0 + lane;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (j0 + lane < per && i4 < n4) v = __ldcg(cnt4 + i4);
        const unsigned s4 = v.x + v.y + v.z + v.w;
        const unsigned in2 = warp_incl_scan_u32(s4, lane), ex2 = in2 - s4;
        const unsigned b2 = __ballot_sync(0xffffffffu, off >= ex2 && off < in2);
        if (b2 != 0u) {
          const int l2 = __ffs(b2) - 1;
          unsigned o2 = off - ex2;          // valid in lane l2
          int t = 0;
          if (o2 >= v.x) { o2 -= v.x; t = 1; if (o2 >= v.y) { o2 -= v.y; t = 2; if (o2 >= v.z) { o2 -= v.z; t = 3; } } }
          idx = __shfl_sync(0xffffffffu, (src * per + j0 + l2) * 4 + t, l2);
          off = __shfl_sync(0xffffffffu, o2, l2);
        } else {
          off -= __shfl_sync(0xffffffffu, in2, 31);
        }
      }
    }
  }
  return idx;
}

// k-th member (ascending row index) of `node` in pool row `row`; executed by one warp.  Search over the member counts
// the workers left behind: [large N only: one count per bucket of 32 tiles, then the bucket's 32 tile counts] or the
// per-tile counts directly; then the tile's leaf ids AND the tile of the split column are requested together, so the
// split value is there when the member's position is known: two dependent L2 round trips at N = 100k, three at N = 1M.
__device__ __forceinline__ float select_split(const Params& P, int c, int row, int node, unsigned k, int var, int* err) {
  const int lane = threadIdx.x & 31;
  CTS(20, 0);   // (time between the end of the job assembly / previous selection and this one: cursor traffic)
  const unsigned* cnt = P.rowcnt + ((size_t)c * P.R + row) * P.cnt_stride;
  unsigned off = k;
  int tile = -1;
  if (P.nb > 0) {
    const int bucket = find_in_counts(reinterpret_cast<const uint4*>(P.coarse + ((size_t)c * P.R + row) * P.nb_stride), P.nb_stride >> 2, off, lane);
    if (bucket >= 0) {
      const int t = bucket * BK_COARSE_TILES + lane;
      const unsigned v = t < P.ntiles ? __ldcg(cnt + t) : 0u;
      const unsigned incl = warp_incl_scan_u32(v, lane), excl = incl - v;
      const unsigned b = __ballot_sync(0xffffffffu, off >= excl && off < incl);
      if (b != 0u) {
        const int src = __ffs(b) - 1;
        tile = bucket * BK_COARSE_TILES + src;
        off -= __shfl_sync(0xffffffffu, excl, src);
      }
    }
  } else {
    tile = find_in_counts(reinterpret_cast<const uint4*>(cnt), P.cnt_stride >> 2, off, lane);
  }
  if (tile < 0) { if (lane == 0) *err |= 2; return 0.0f; }
  CTS(26, tile);
  const size_t base = (size_t)tile * BK_WARP_TILE + (size_t)lane * BK_ROWS_PER_LANE;
  const unsigned long long ids = __ldcg(reinterpret_cast<const unsigned long long*>(P.rows + ((size_t)c * P.R + row) * P.Npad + base));
  const float4* xp = reinterpret_cast<const float4*>(P.X + (size_t)var * P.Npad + base);
  const float4 x0 = __ldg(xp), x1 = __ldg(xp + 1);       // (speculative: 1 KB per selection buys a whole round trip)
  unsigned mm = 0;
#pragma unroll
  for (int e = 0; e < 8; ++e) mm |= (((unsigned)(ids >> (8 * e)) & 255u) == (unsigned)node) ? (1u << e) : 0u;
  const unsigned v = __popc(mm);
  CTS(27, v);
  const unsigned incl = warp_incl_scan_u32(v, lane), excl = incl - v;
  const unsigned b = __ballot_sync(0xffffffffu, off >= excl && off < incl);
  if (b == 0) { if (lane == 0) *err |= 4; return 0.0f; }
  const int src = __ffs(b) - 1;
  float val = 0.0f;
  if (lane == src) {
    const int pos = (int)__fns(mm, 0, (int)(off - excl) + 1);   // position of the (off - excl)-th member among the lane's 8 rows
    const float xs[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
    for (int e = 0; e < 8; ++e) if (e == pos) val = xs[e];
  }
  val = __shfl_sync(0xffffffffu, val, src);
  CTS(28, __float_as_int(val));
  return val;
}

// Fixed-point weights of sh.lw[first..first+count) and systematic resampling into sh.anc[0..count) (bk_spec.h:
// bk_weight_fix / bk_resample_*).  Runs in warps 1..4 (count <= 128: element i lives in warp 1 + i / 32, lane i % 32).
// Everything is integer after the float exponential, so the warp-parallel scan gives the oracle's sequential sums
// bit for bit.  Ends with a block barrier.
// Never on warp 0: lane 0 of warp 0 executes the chain's scalar sections alone and (non-aligned barriers) can leave
// that warp split, and a split warp takes the WARPSYNC slow path on EVERY shuffle (~200 cycles each, measured);
// warps 1.. never diverge across a barrier, so their collectives stay on the fast path.
__device__ __forceinline__ unsigned long long shfl_up_u64(unsigned long long v, int o) {
  const unsigned lo = __shfl_up_sync(0xffffffffu, (unsigned)v, o), hi = __shfl_up_sync(0xffffffffu, (unsigned)(v >> 32), o);
  return ((unsigned long long)hi << 32) | lo;
}
__device__ __forceinline__ void normalise_and_resample(const Params& P, CtlShared& sh, int first, int count, uint32_t u32) {
  // warps 1..4, one 32-element slice each (count <= 128), meeting on their own named barrier
  const int wq = (int)(threadIdx.x >> 5) - 1;
  if (wq >= 0 && wq < 4) {
#define R_SYNC() asm volatile("barrier.sync 13, 128;" ::: "memory")
    __shared__ unsigned long long s_slice[4];
    const int lane = threadIdx.x & 31;
    CT0();
    double mx = -1.7976931348623157e308;   // every warp forms the same maximum
    for (int i = lane; i < count; i += 32) { double v = sh.lw[first + i]; mx = v > mx ? v : mx; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { double ov = __shfl_xor_sync(0xffffffffu, mx, o); mx = ov > mx ? ov : mx; }
    CT(0, __double2loint(mx));
    const int i = wq * 32 + lane;
    unsigned long long s = i < count ? bk_weight_fix(sh.lw[first + i], mx) : 0ull;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned long long nb = shfl_up_u64(s, o); if (lane >= o) s += nb; }
    if (lane == 31) s_slice[wq] = s;
    R_SYNC();
    unsigned long long before = 0ull, s_last = 0ull;     // sum of the earlier slices; S[count - 1]
#pragma unroll
    for (int k = 0; k < 4; ++k) { const unsigned long long t = s_slice[k]; if (k < wq) before += t; s_last += t; }
    s += before;
    if (i < count) sh.cum_s[i] = s;
    R_SYNC();
    CT(1, (int)s_last);
    if (i < count) {
      // first index whose running sum reaches the point (= the walk `while (point > c[idx]) idx++`)
      const bk_u128 point = bk_resample_point((uint32_t)i, u32, s_last);
      int lo = 0, hi = count - 1;
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (!bk_resample_le(point, sh.cum_s[mid], (uint32_t)count)) lo = mid + 1; else hi = mid; }
      sh.anc[i] = lo;
    }
    CT(4, sh.anc[lane < count ? lane : 0]);
#undef R_SYNC
  }
  CTRL_SYNC();
}

// Resampling by the particle threads themselves (round path).  Thread q = BK_WTID of warps 1..4 owns particle q; the
// weight vector covers particles first..first+count-1 (element i = q - first).  Every warp forms the maximum of all log
// weights from the particle headers (two 32-bit REDUX on an order-preserving key instead of ten 64-bit shuffles), the
// fixed-point weights are scanned in particle order (warp scan + one slice exchange), and thread q searches the
// ancestor of ITS OWN point, so the result stays in a register of the thread that pops the slot next: no block barrier
// between resampling and the queue pops.  Same integers as normalise_and_resample / the oracle's sequential sums.
// Must be called by all threads of warps 1..4 (two named barriers); returns the ancestor index in [0, count).
__device__ __forceinline__ int resample_own(const Params& P, int c, int buf, CtlShared& sh, int first, int count, uint32_t u32) {
#define R_SYNC() asm volatile("barrier.sync 13, 128;" ::: "memory")
  const int q = BK_WTID, lane = threadIdx.x & 31, wq = q >> 5;
  const int i = q - first;
  const bool in = i >= 0 && i < count;
  unsigned mh = 0u, ml = 0u;
  {
    unsigned long long best = 0ull;
    for (int j = lane; j < count; j += 32) {
      const unsigned long long b = bk_d2bits(pref(P, c, buf, first + j).h->lw);
      const unsigned long long key = (b >> 63) ? ~b : (b | 0x8000000000000000ull);   // unsigned order = double order
      best = key > best ? key : best;
    }
    mh = __reduce_max_sync(0xffffffffu, (unsigned)(best >> 32));
    ml = __reduce_max_sync(0xffffffffu, (unsigned)(best >> 32) == mh ? (unsigned)best : 0u);
  }
  const unsigned long long mk = ((unsigned long long)mh << 32) | ml;
  const double mx = bk_bits2d((mk >> 63) ? (mk & 0x7FFFFFFFFFFFFFFFull) : ~mk);
  unsigned long long s = in ? bk_weight_fix(pref(P, c, buf, q).h->lw, mx) : 0ull;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const unsigned long long nb = shfl_up_u64(s, o); if (lane >= o) s += nb; }
  if (lane == 31) sh.r_slice[wq] = s;
  R_SYNC();
  unsigned long long before = 0ull, s_last = 0ull;
#pragma unroll
  for (int k = 0; k < 4; ++k) { const unsigned long long t = sh.r_slice[k]; if (k < wq) before += t; s_last += t; }
  s += before;
  if (in) sh.cum_s[i] = s;
  R_SYNC();
  int lo = 0;
  if (in) {
    const bk_u128 point = bk_resample_point((uint32_t)i, u32, s_last);
    int hi = count - 1;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (!bk_resample_le(point, sh.cum_s[mid], (uint32_t)count)) lo = mid + 1; else hi = mid; }
  }
#undef R_SYNC
  return lo;
}

__device__ __forceinline__ void zero_acc0(const Params& P, int c) {
  unsigned long long* a = P.acc0 + (size_t)c * BK_ACC0_WORDS;
  for (int i = BK_WTID; i >= 0 && i < BK_ACC0_WORDS; i += BK_WTHREADS) a[i] = 0ull;
}

__device__ __forceinline__ void rebuild_cum_dev(const Params& P, int c) {  // thread 0 only
  double* av = P.alpha_vec + (size_t)c * P.p;
  double* cum = P.cum + (size_t)c * P.p;
  double tot = 0.0;
  for (int v = 0; v < P.p; ++v) tot = BK_DADD(tot, av[v]);
  double run = 0.0;
  for (int v = 0; v < P.p; ++v) { run = BK_DADD(run, av[v]); cum[v] = BK_DDIV(run, tot); }
}

// ------------------------------------------------------------------ control phase
__device__ __forceinline__ void init_particles(const Params& P, int c, ChainCtl* ctl, ChainHot* hot, CtlShared& sh) {
  const int t = hot->cur_tree;
  const unsigned long long* a0 = P.acc0 + (size_t)c * BK_ACC0_WORDS;
  const DNode* ft = P.forest + ((size_t)c * P.m + t) * BK_MAX_NODES;
  const int nn = P.forest_nn[(size_t)c * P.m + t];
  const PRef p0 = pref(P, c, 0, 0);
  for (int k = BK_WTID; k >= 0 && k < nn; k += BK_WTHREADS) {
    DNode nd = ft[k];
    nd.sst = 0;
    nd.sr = (int64_t)__ldcg(a0 + (size_t)k * BK_ACC0_STRIDE + 0);
    p0.node(k) = nd;
  }
  for (int r = BK_WTID; r >= 0 && r < P.R; r += BK_WTHREADS) sh.row_cnt_node[r] = -1;
  for (int v = BK_WTID; v >= 0 && v < P.p && v < BK_CUM_SMEM; v += BK_WTHREADS) sh.cum_prior[v] = P.cum[(size_t)c * P.p + v];
  CTRL_SYNC();
  const bool bern = P.lik != BK_LIK_NORMAL;   // every family without a sufficient statistic: weights come from the LL epoch
  if (threadIdx.x == 0) {
    // Bernoulli: integer sum of the leaves' quantised log-likelihood terms (exact in double), plus the terms of the rows
    // the tree dropped for a missing covariate (they predict 0)
    double gain = bern ? (double)(long long)__ldcg(a0 + (size_t)256 * BK_ACC0_STRIDE + 1) : 0.0;
    for (int k = 0; k < nn; ++k) {
      const DNode& nd = p0.node(k);
      if (nd.var < 0) gain = BK_DADD(gain, bern ? (double)nd.sr : bk_leaf_gain(node_stats(nd), nd.value, P.inv_qscale));
    }
    const double r2_total = bk_total_r2(bk_u128_from_split(__ldcg(a0 + (size_t)255 * BK_ACC0_STRIDE + 2), __ldcg(a0 + (size_t)255 * BK_ACC0_STRIDE + 1)), P.inv_qscale);
    hot->r2_total = r2_total;
    p0.h->n_nodes = nn; p0.h->q_head = nn; p0.h->row = BK_ROW_FOREST;
    p0.h->gain = gain; p0.h->lw = bern ? bk_bern_loglik(gain) : bk_normal_loglik_pre(bk_ssq_from_gain(r2_total, gain), hot->ll_inv2s2, hot->ll_c);
    hot->buf = 0; hot->round = 0; sh.live = 0;
  }
  const int q = BK_WTID;
  if (q >= 1 && q < P.P) {
    bk_stats tot;
    tot.n = P.N;
    tot.sr = (int64_t)__ldcg(a0 + (size_t)255 * BK_ACC0_STRIDE + 0);
    const double r2_total = bk_total_r2(bk_u128_from_split(__ldcg(a0 + (size_t)255 * BK_ACC0_STRIDE + 2), __ldcg(a0 + (size_t)255 * BK_ACC0_STRIDE + 1)), P.inv_qscale);
    tot.sst = (int64_t)__ldcg(a0 + (size_t)255 * BK_ACC0_STRIDE + 3);
    const PRef S = pref(P, c, 0, q);
    DNode nd;
    nd.var = -1; nd.split = 0.0f; nd.left = -1; nd.depth = 0; nd.value = P.init_leaf; nd.aux[0] = 0; nd.aux[1] = 0; nd.aux[2] = 0;
    for (int j = 1; j < P.K; ++j) set_node_val(nd, j, P.init_leaf);
    set_node_stats(nd, tot);
    S.node(0) = nd;
    S.h->n_nodes = 1; S.h->q_head = 0; S.h->row = BK_ROW_VIRTUAL;
    S.h->gain = bern ? (double)tot.sr : bk_leaf_gain(tot, P.init_leaf, P.inv_qscale);
    S.h->lw = bern ? bk_bern_loglik(S.h->gain) : bk_normal_loglik_pre(bk_ssq_from_gain(r2_total, S.h->gain), hot->ll_inv2s2, hot->ll_c);
  }
  CTRL_SYNC();
}

// first index with u2 < cum[idx], else p-1 (split variable ~ split prior; bart.py:139,155)
__device__ __forceinline__ int draw_variable_dev(const Params& P, int c, const CtlShared& sh, double u2) {
  const double* cum = P.cum + (size_t)c * P.p;
  int lo = 0, hi = P.p - 1;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    const double cv = mid < BK_CUM_SMEM ? sh.cum_prior[mid] : cum[mid];
    if (u2 < cv) hi = mid; else lo = mid + 1;
  }
  return lo;
}

__device__ __forceinline__ void copy_particles(const Params& P, int c, int buf, const int* anc_of_slot);

// Executes a deferred particle copy (see CtlShared::copy_pending): buffer `buf` -> `buf ^ 1` through src_slot, then
// the queue heads open_round() recorded; flips the buffer.  All control threads.
__device__ __forceinline__ void apply_pending_copy(const Params& P, int c, ChainHot* hot, CtlShared& sh) {
  if (!sh.copy_pending) return;     // (uniform: shared flag, read after a barrier)
  const int buf = hot->buf;
  copy_particles(P, c, buf, sh.src_slot);
  if (BK_WTID >= 1 && BK_WTID < P.P) pref(P, c, buf ^ 1, BK_WTID).h->q_head = sh.s_qh[BK_WTID];
  if (threadIdx.x == 0) { hot->buf = buf ^ 1; sh.copy_pending = 0; }
  CTRL_SYNC();
}

// Runs right after a ROUND epoch has been published (all control threads), overlapping the workers.
__device__ __forceinline__ void shadow_round(const Params& P, int c, ChainHot* hot, CtlShared& sh) {
  const int round = hot->round, t = hot->cur_tree;
  const uint32_t S0 = P.seed, C0 = P.chain_base + (uint32_t)(c / P.G), G0 = (uint32_t)(c % P.G), D0 = (uint32_t)hot->draw;
  apply_pending_copy(P, c, hot, sh);
  const int q = BK_WTID;
  if (q >= 1 && q < P.P) {
    // leaf-value normals of this round (used by finalize_own when the epoch is done) for the slots that grow
    if (sh.s_kind[q] == 1) {
      sh.pre_zl[q] = bk_normal(bk_rng(S0, C0, D0, G0, (uint32_t)t, (uint32_t)round, (uint32_t)q, BK_Z_LEFT));
      sh.pre_zr[q] = bk_normal(bk_rng(S0, C0, D0, G0, (uint32_t)t, (uint32_t)round, (uint32_t)q, BK_Z_RIGHT));
    }
    // proposal draws of the NEXT round for every slot
    const uint32_t r1 = (uint32_t)(round + 1);
    sh.pre_u1[q] = bk_u01(bk_rng(S0, C0, D0, G0, (uint32_t)t, r1, (uint32_t)q, BK_U_LEAF).v[0]);
    sh.pre_v[q] = draw_variable_dev(P, c, sh, bk_u01(bk_rng(S0, C0, D0, G0, (uint32_t)t, r1, (uint32_t)q, BK_U_VAR).v[0]));
    sh.pre_u3[q] = bk_rng(S0, C0, D0, G0, (uint32_t)t, r1, (uint32_t)q, BK_U_VAL).v[0];
  }
  if (threadIdx.x == 0) {
    sh.pre_ures = bk_rng(S0, C0, D0, G0, (uint32_t)t, (uint32_t)round, 0, BK_U_RESAMPLE).v[0];
    sh.pre_prop_tree = t; sh.pre_prop_round = round + 1;
    sh.pre_z_tree = t; sh.pre_z_round = round;
  }
  CTRL_SYNC();
}

// n-th (0-based) zero bit of the used-row bitmap = n-th free pool row; -1 if there are fewer
__device__ __forceinline__ int nth_free_row(const CtlShared& sh, int R, int n) {
#pragma unroll
  for (int w = 0; w < 8; ++w) {
    const int rows_here = R - 32 * w;
    if (rows_here <= 0) break;
    const unsigned valid = rows_here >= 32 ? 0xFFFFFFFFu : ((1u << rows_here) - 1u);
    const unsigned freeb = ~sh.used_rows[w] & valid;
    const int cfree = __popc(freeb);
    if (n < cfree) return 32 * w + (int)__fns(freeb, 0, n + 1);
    n -= cfree;
  }
  return -1;
}

// Shared-tree multi-output: K leaf values per child from the per-output sums of both children (accK, zeroed here: the
// release fence of the next publication orders the stores before the workers' next adds), K normals per child from the
// Philox blocks whose `group` word is the output index; the values also go to the LL job of the slot.
__device__ __forceinline__ void finalize_multi(const Params& P, int c, ChainHot* hot, CtlShared& sh, int buf, int q, int ji, int round, int t,
                                            int rbase) {
  const uint32_t S0 = P.seed, C0 = P.chain_base + (uint32_t)c, D0 = (uint32_t)hot->draw;
  const Job jb = sh.jobs[ji];
  const PRef S = pref(P, c, buf, q);
  int n_left;
  {
    unsigned long long* acc = P.accL + ((size_t)c * P.P + q) * BK_ACC_STRIDE;
    const unsigned long long a_n = __ldcg(acc + BK_ACC_N);
    n_left = (int)(a_n - sh.acc_prev[q][BK_ACC_N]);
    sh.acc_prev[q][BK_ACC_N] = a_n;
  }
  DNode parent = S.node(jb.node);
  const int nn = S.h->n_nodes, n_right = parent.n - n_left;
  const float split = sh.s_split[q];
  parent.var = jb.var; parent.split = split; parent.left = nn;
  S.node(jb.node) = parent;
  DNode nl; memset(&nl, 0, sizeof(nl));
  nl.var = -1; nl.left = -1; nl.depth = parent.depth + 1;
  DNode nr = nl;
  nl.n = n_left; nr.n = n_right;
  unsigned long long* ak = P.accK + ((size_t)c * P.P + q) * P.K * 2;
  for (int j = 0; j < P.K; ++j) {
    const long long sst_l = (long long)__ldcg(ak + 2 * j), sst_r = (long long)__ldcg(ak + 2 * j + 1);
    ak[2 * j] = 0ull; ak[2 * j + 1] = 0ull;
    const double zl = bk_normal(bk_rng(S0, C0, D0, (uint32_t)j, (uint32_t)t, (uint32_t)round, (uint32_t)q, BK_Z_LEFT));
    const double zr = bk_normal(bk_rng(S0, C0, D0, (uint32_t)j, (uint32_t)t, (uint32_t)round, (uint32_t)q, BK_Z_RIGHT));
    const float vl = bk_leaf_value(n_left, sst_l, P.inv_qm, zl, hot->leaf_sdk[j]);
    const float vr = bk_leaf_value(n_right, sst_r, P.inv_qm, zr, hot->leaf_sdk[j]);
    set_node_val(nl, j, vl); set_node_val(nr, j, vr);
    sh.job_vals[ji][0][j] = vl; sh.job_vals[ji][1][j] = vr;
    if (j == 0) { nl.sst = sst_l; nr.sst = sst_r; }
  }
  S.node(nn) = nl; S.node(nn + 1) = nr;
  S.h->n_nodes = nn + 2;
  Job lj = jb;   // the partition job becomes the LL job of the same particle (same list position)
  lj.kind = BK_JOB_LL; lj.src_row = jb.dst_row;
  sh.jobs[ji] = lj;
  S.h->row = jb.dst_row;
  bk_trace_rec* rec = trace_at(P, c, rbase + q - 1);
  if (rec) { rec->var = jb.var; rec->split = split; rec->n_left = n_left; rec->n_right = n_right; rec->val_left = nl.value; rec->val_right = nr.value; }
}

// apply the statistics of the finished ROUND to slot q's particle if it grew (thread q = BK_WTID); no barrier inside
template <int MODE>
__device__ __forceinline__ void finalize_own(const Params& P, int c, ChainHot* hot, CtlShared& sh, int buf, int round, int t, int rbase) {
  const int q = BK_WTID;
  if (q < 1 || q >= P.P) return;
  const int ji = sh.s_jobidx[q];
  if (ji < 0) return;
  const uint32_t S0 = P.seed, C0 = P.chain_base + (uint32_t)(c / P.G), G0 = (uint32_t)(c % P.G), D0 = (uint32_t)hot->draw;
  if (BK_IS_MULTI(P)) { finalize_multi(P, c, hot, sh, buf, q, ji, round, t, rbase); return; }
  const bool bern = P.lik != BK_LIK_NORMAL;   // every family without a sufficient statistic: weights come from the LL epoch
  const Job jb = sh.jobs[ji];   // the job list staged by open_round() is still in shared memory
  const PRef S = pref(P, c, buf, q);
  unsigned long long* acc = P.accL + ((size_t)c * P.P + q) * BK_ACC_STRIDE;
  bk_stats sl;
  {
    const unsigned long long a_n = __ldcg(acc + BK_ACC_N), a_st = __ldcg(acc + BK_ACC_SST), a_sr = __ldcg(acc + BK_ACC_SR);
    unsigned long long* prev = sh.acc_prev[q];
    sl.n = (int32_t)(a_n - prev[BK_ACC_N]);
    sl.sst = (int64_t)(a_st - prev[BK_ACC_SST]);
    sl.sr = (int64_t)(a_sr - prev[BK_ACC_SR]);
    prev[BK_ACC_N] = a_n; prev[BK_ACC_SST] = a_st; prev[BK_ACC_SR] = a_sr;
  }
  DNode parent = S.node(jb.node);
  const bk_stats sp = node_stats(parent);
  bk_stats sr = bk_stats_sub(sp, sl);
  if (BK_MISSING_ENABLED && jb.pad[1]) {   // the split column has missing values: the rows dropped from the node count for neither child
    const unsigned long long a_n = __ldcg(acc + BK_ACC_ND), a_st = __ldcg(acc + BK_ACC_SSTD), a_sr = __ldcg(acc + BK_ACC_SRD);
    unsigned long long* prev = sh.acc_prev[q];
    bk_stats sd;
    sd.n = (int32_t)(a_n - prev[BK_ACC_ND]); sd.sst = (int64_t)(a_st - prev[BK_ACC_SSTD]); sd.sr = (int64_t)(a_sr - prev[BK_ACC_SRD]);
    prev[BK_ACC_ND] = a_n; prev[BK_ACC_SSTD] = a_st; prev[BK_ACC_SRD] = a_sr;
    if (bern) { S.h->gain = BK_DADD(S.h->gain, (double)sd.sr); sd.sr = 0; }   // their log-likelihood terms at the value 0 they now carry
    sr = bk_stats_sub(sr, sd);
  }
  if (bern) { sl.sr = 0; sr.sr = 0; }   // the children's log-likelihood sums arrive with the LL epoch
  const bool pre = sh.pre_z_tree == t && sh.pre_z_round == round;
  const double zl = pre ? sh.pre_zl[q] : bk_normal(bk_rng(S0, C0, D0, G0, (uint32_t)t, (uint32_t)round, (uint32_t)q, BK_Z_LEFT));
  const double zr = pre ? sh.pre_zr[q] : bk_normal(bk_rng(S0, C0, D0, G0, (uint32_t)t, (uint32_t)round, (uint32_t)q, BK_Z_RIGHT));
  const float vl = bk_leaf_value(sl.n, sl.sst, P.inv_qm, zl, hot->leaf_sd);
  const float vr = bk_leaf_value(sr.n, sr.sst, P.inv_qm, zr, hot->leaf_sd);
  const double g_parent = bk_leaf_gain(sp, parent.value, P.inv_qscale);
  const int nn = S.h->n_nodes;
  const float split = sh.s_split[q];   // (the staged job carries the split value only in its global copy)
  parent.var = jb.var; parent.split = split; parent.left = nn;
  S.node(jb.node) = parent;
  DNode nl; nl.var = -1; nl.split = 0.0f; nl.left = -1; nl.depth = parent.depth + 1; nl.value = vl; nl.aux[0] = 0; nl.aux[1] = 0; nl.aux[2] = 0;
  set_node_stats(nl, sl);
  DNode nr = nl; nr.value = vr; set_node_stats(nr, sr);
  S.node(nn) = nl; S.node(nn + 1) = nr;
  S.h->n_nodes = nn + 2;
  if (!bern) {
    const double gain = BK_DADD(BK_DADD(BK_DSUB(S.h->gain, g_parent), bk_leaf_gain(sl, vl, P.inv_qscale)), bk_leaf_gain(sr, vr, P.inv_qscale));
    S.h->gain = gain;
    S.h->lw = bk_normal_loglik_pre(bk_ssq_from_gain(hot->r2_total, gain), hot->ll_inv2s2, hot->ll_c);
  } else {   // turn the partition job into the LL job of the same particle (same list position)
    Job lj = jb;
    lj.kind = BK_JOB_LL; lj.src_row = jb.dst_row; lj.split = vl; lj.rule = __float_as_int(vr);
    sh.jobs[ji] = lj;
  }
  S.h->row = jb.dst_row;
  bk_trace_rec* rec = trace_at(P, c, rbase + q - 1);
  if (rec) { rec->var = jb.var; rec->split = split; rec->n_left = sl.n; rec->n_right = sr.n; rec->val_left = vl; rec->val_right = vr; }
}

// Bernoulli: the LL epoch summed the quantised log-likelihood terms of the rows of slot q's new leaf pair
__device__ __forceinline__ void finalize_ll_own(const Params& P, int c, CtlShared& sh, int buf) {
  const int q = BK_WTID;
  if (q < 1 || q >= P.P) return;
  const int ji = sh.s_jobidx[q];
  if (ji < 0 || sh.jobs[ji].kind != BK_JOB_LL) return;
  const Job jb = sh.jobs[ji];
  const PRef S = pref(P, c, buf, q);
  unsigned long long* acc = P.accL + ((size_t)c * P.P + q) * BK_ACC_STRIDE;
  const unsigned long long a_l = __ldcg(acc + BK_ACC_LLL), a_r = __ldcg(acc + BK_ACC_LLR);
  const long long ll_l = (long long)(a_l - sh.acc_prev[q][BK_ACC_LLL]), ll_r = (long long)(a_r - sh.acc_prev[q][BK_ACC_LLR]);
  sh.acc_prev[q][BK_ACC_LLL] = a_l; sh.acc_prev[q][BK_ACC_LLR] = a_r;
  const long long ll_parent = S.node(jb.node).sr;
  S.node(jb.left_id).sr = ll_l;
  S.node(jb.left_id + 1).sr = ll_r;
  const double llq = BK_DADD(BK_DSUB(S.h->gain, (double)ll_parent), (double)(ll_l + ll_r));   // integers: exact
  S.h->gain = llq;
  S.h->lw = bk_bern_loglik(llq);
}

// bk_spec.h BK_SPLIT_TRIES: candidates 1..3 of a split-value draw whose first candidate member has no value in the split
// column (NaN).  Only the BK_MODE_MISSING instantiation of the round path contains it.
__device__ __forceinline__ float retry_split_value(const Params& P, int c, CtlShared& sh, int buf, int sl, uint32_t S0, uint32_t C0,
                                                uint32_t D0, uint32_t G0, int t, int round) {
  const int lane = threadIdx.x & 31;
  const PRef S = pref(P, c, buf, sh.src_slot[sl]);
  const bk_u32x4 wv = bk_rng(S0, C0, D0, G0, (uint32_t)t, (uint32_t)round, (uint32_t)sl, BK_U_VAL);
  const uint32_t n = (uint32_t)S.node(sh.s_j[sl]).n;
  float sv = bk_bits2f(0x7FC00000u);
  for (int tr = 1; tr < BK_SPLIT_TRIES; ++tr) {
    const unsigned kk = bk_index(wv.v[tr], n);
    if (sh.s_row[sl] == BK_ROW_VIRTUAL) {
      sv = __ldg(P.X + (size_t)sh.s_v[sl] * P.Npad + kk);
    } else {
      int err = 0;
      sv = select_split(P, c, sh.s_row[sl], sh.s_j[sl], kk, sh.s_v[sl], &err);
      if (lane == 0 && err) atomicOr(&sh.err_bits, err);
    }
    if (!(sv != sv)) break;
  }
  return sv;
}

// Opens round `round` of tree t: queue pops and grow decisions of every slot (through its ancestor `src` when a
// resampling has just happened: the particle copy itself is deferred into the epoch's shadow), row allocation, job
// list, split values (k-th member).  Called by every control thread; `src` is the calling particle thread's own
// ancestor slot.  Two block barriers inside; returns (uniformly) the job count.  The caller's barrier orders the
// global job list before thread 0 publishes the epoch.
template <int MODE>
__device__ __forceinline__ int open_round(const Params& P, int c, ChainCtl* ctl, ChainHot* hot, CtlShared& sh, const int buf, const int round,
                          const int rbase, const bool deferred, const int src) {
  const int t = hot->cur_tree;
  const uint32_t S0 = P.seed, C0 = P.chain_base + (uint32_t)(c / P.G), G0 = (uint32_t)(c % P.G), D0 = (uint32_t)hot->draw;
  const int q = BK_WTID, lane = threadIdx.x & 31, w = q >> 5;
  const bool is_p = q >= 1 && q < P.P;
  int kind = 0, j = -1, v = -1, next = -1, row = -3, nn = 0, sparse = 0;
  unsigned k = 0;
  if (is_p) {
    const PRef S = pref(P, c, buf, deferred ? src : q);   // the slot's state-to-be = its ancestor's state
    int qh = S.h->q_head;
    nn = S.h->n_nodes; row = S.h->row;
    if (qh < nn) {
      j = qh; qh += 1;
      if (!deferred) S.h->q_head = qh;   // (deferred: several slots may share this ancestor; written after the copy)
      const int depth = S.node(j).depth;
      const int n = S.node(j).n;
      const double pl = depth < 64 ? sh.p_leaf[depth] : (depth < BK_MAX_DEPTH_TABLE ? P.p_leaf[depth] : 1.0);
      const bool pre = sh.pre_prop_tree == t && sh.pre_prop_round == round;   // draws made in the shadow of the last epoch
      const double u1 = pre ? sh.pre_u1[q] : bk_u01(bk_rng(S0, C0, D0, G0, (uint32_t)t, (uint32_t)round, (uint32_t)q, BK_U_LEAF).v[0]);
      if (u1 > pl && nn + 2 <= BK_MAX_NODES) {
        v = pre ? sh.pre_v[q]
                : draw_variable_dev(P, c, sh, bk_u01(bk_rng(S0, C0, D0, G0, (uint32_t)t, (uint32_t)round, (uint32_t)q, BK_U_VAR).v[0]));
        if (n >= 2) {
          const uint32_t u3 = pre ? sh.pre_u3[q] : bk_rng(S0, C0, D0, G0, (uint32_t)t, (uint32_t)round, (uint32_t)q, BK_U_VAL).v[0];
          // SubsetSplit columns draw a set of categories from the raw uniform (bk_subset_draw), not a member index
          const bool sub = BK_MISSING_ENABLED && (v < BK_CUM_SMEM ? (int)sh.rules[v] : P.rules[v]) == BK_RULE_SUBSET;
          k = sub ? u3 : bk_index(u3, (uint32_t)n);
          kind = 1;
        }
      }
      if (kind == 1) next = qh;            // a queued node, or one of the two children being made
      else { next = qh < nn ? qh : -1; if (next >= 0) kind = 2; }
      sparse = ((long long)n * BK_SPARSE_DIV < (long long)P.N) ? 1 : 0;
    }
    sh.src_slot[q] = deferred ? src : q;
    sh.s_qh[q] = qh;
    sh.s_kind[q] = kind; sh.s_j[q] = j; sh.s_v[q] = v; sh.s_k[q] = k; sh.s_row[q] = row;
    sh.s_jobidx[q] = -1;
    if (row >= 0) atomicOr(&sh.used_rows[row >> 5], 1u << (row & 31));
    if (kind == 1 && row != BK_ROW_VIRTUAL && sh.row_cnt_node[row] != j) atomicOr(&sh.err_bits, 1);
    bk_trace_rec* rec = trace_at(P, c, rbase + q - 1);
    if (rec) {
      bk_trace_rec r; memset(&r, 0, sizeof(r));
      r.kind = 1; r.tree = t; r.round = round; r.particle = q; r.node = j; r.var = -1; r.ancestor = -1;
      *rec = r;
    }
  }
  if (q == 0) { sh.src_slot[0] = 0; sh.s_kind[0] = 0; sh.s_jobidx[0] = -1; }
  unsigned bg = 0u;
  if (w >= 0 && w < 4) {   // the particle warps (converged: every lane of warps 1..4 gets here)
    bg = __ballot_sync(0xffffffffu, kind == 1);
    const unsigned bv = __ballot_sync(0xffffffffu, kind == 1 && row == BK_ROW_VIRTUAL);
    if (lane == 0) { sh.warp_grow[w] = __popc(bg); sh.warp_grow_root[w] = __popc(bv); }
  }
  CTRL_SYNC();
  CTS(19, 0);
  TSUB(4);
  // ---- jobs (particle threads) beside the k-th member selections (every warp but the scalar warp 0)
  int tot_g = 0;
  if (w >= 0 && w < 4) {
    int off_g = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) { const int cw = sh.warp_grow[i]; if (i < w) off_g += cw; tot_g += cw; }
    if (kind == 1) {
      const int rank_g = off_g + __popc(bg & ((1u << lane) - 1u));
      const int dst = nth_free_row(sh, P.R, rank_g);
      if (dst < 0) atomicOr(&sh.err_bits, 8);
      Job jb;
      jb.kind = BK_JOB_PARTITION; jb.slot = q; jb.src_row = row; jb.dst_row = dst < 0 ? 0 : dst;
      jb.node = j; jb.var = v; jb.split = 0.0f; jb.left_id = nn;
      jb.next_node = next; jb.rule = v < BK_CUM_SMEM ? (int)sh.rules[v] : P.rules[v]; jb.pad[0] = sparse; jb.pad[1] = P.col_nan[v];
      if (dst >= 0) sh.row_cnt_node[dst] = next;
      sh.jobs[rank_g] = jb;
      sh.s_jobidx[q] = rank_g;
    } else if (kind == 2 && row >= 0) {
      // count-only: the members of the next queue node per tile, once per shared row (slots that share a row have the
      // same queue)
      const int old = atomicExch(&sh.row_cnt_node[row], next);
      if (old != next) {
        const int pos = tot_g + atomicAdd(&sh.n_cnt_jobs, 1);   // (order among count jobs is immaterial)
        Job jb;
        jb.kind = BK_JOB_COUNT; jb.slot = q; jb.src_row = row; jb.dst_row = row; jb.node = -1; jb.var = 0;
        jb.split = 0.0f; jb.left_id = 0; jb.next_node = next; jb.rule = 0; jb.pad[0] = 0; jb.pad[1] = 0;
        sh.jobs[pos] = jb;
      }
    }
  }
  CTS(20, 0);
  if (threadIdx.x >= 32) {   // split values: warps 1.. take growing slots from the shared cursor
    for (;;) {
      int sl = 0;
      if (lane == 0) sl = atomicAdd(&sh.next_sel, 1);
      sl = __shfl_sync(0xffffffffu, sl, 0);
      if (sl >= P.P) break;
      if (sh.s_kind[sl] != 1) continue;
      float sv;
      const int vv = sh.s_v[sl];
      if (BK_MISSING_ENABLED && (vv < BK_CUM_SMEM ? (int)sh.rules[vv] : P.rules[vv]) == BK_RULE_SUBSET) {
        // SubsetSplit: the categories present among the node's members were left by the workers beside the member
        // counts (a stump's root: the whole column's); the set is drawn from them.  Fewer than two present: no split
        // (NaN = the cancelled-split marker of the missing-value path; a set's float is an integer < 2^24).
        const int rw = sh.s_row[sl];
        const unsigned present = rw == BK_ROW_VIRTUAL ? __ldg(P.col_cats + vv)
                                                      : __ldcg(P.present + ((size_t)c * P.R + rw) * BK_MAX_SUBSET_COLS + __ldg(P.subset_idx + vv));
        const unsigned set = bk_subset_draw(present, sh.s_k[sl]);
        sv = set ? (float)set : bk_bits2f(0x7FC00000u);
      } else {
        if (sh.s_row[sl] == BK_ROW_VIRTUAL) {
          sv = __ldg(P.X + (size_t)vv * P.Npad + sh.s_k[sl]);
        } else {
          int err = 0;
          sv = select_split(P, c, sh.s_row[sl], sh.s_j[sl], sh.s_k[sl], vv, &err);
          if (lane == 0 && err) atomicOr(&sh.err_bits, err);
          CTN(29);
        }
        if (BK_MISSING_ENABLED && sv != sv) sv = retry_split_value(P, c, sh, buf, sl, S0, C0, D0, G0, t, round);   // the candidate's covariate is missing
      }
      if (lane == 0) { sh.s_split[sl] = sv; if (sv != sv) atomicAdd(&sh.n_nan_fail, 1); }
    }
  }
  CTS(21, 0);
  CTRL_SYNC();
  CTS(22, 0);
  TSUB(5);
  if (BK_MISSING_ENABLED && sh.n_nan_fail) {
    // Some slot found only members with a missing covariate (BK_SPLIT_TRIES misses): its node stays a leaf.  The
    // partition job becomes the count-only job the slot would have got (or nothing), and finalize skips the slot.
    if (is_p && kind == 1 && sh.s_split[q] != sh.s_split[q]) {
      const int ji = sh.s_jobidx[q];
      const int next2 = sh.s_qh[q] < nn ? sh.s_qh[q] : -1;
      Job jb = sh.jobs[ji];
      jb.kind = BK_JOB_NOP; jb.next_node = -1;
      if (row >= 0 && next2 >= 0 && atomicExch(&sh.row_cnt_node[row], next2) != next2) {
        jb.kind = BK_JOB_COUNT; jb.src_row = row; jb.dst_row = row; jb.next_node = next2;
      }
      sh.jobs[ji] = jb;
      sh.s_jobidx[q] = -1; sh.s_kind[q] = 0;
      if (row == BK_ROW_VIRTUAL) atomicAdd(&sh.n_fail_root, 1);
    }
    CTRL_SYNC();
  }
  // ---- the list goes to global memory with the split values patched in (48-byte descriptors, 16-byte pieces)
  tot_g = sh.warp_grow[0] + sh.warp_grow[1] + sh.warp_grow[2] + sh.warp_grow[3];
  const int nj = tot_g + sh.n_cnt_jobs;
  for (int i = q; i >= 0 && i < nj * 3; i += BK_WTHREADS) {
    uint4 piece = reinterpret_cast<const uint4*>(sh.jobs)[i];
    if (i % 3 == 1 && i / 3 < tot_g) piece.z = __float_as_uint(sh.s_split[sh.jobs[i / 3].slot]);
    reinterpret_cast<uint4*>(ctl->jobs[0])[i] = piece;
  }
  const bool zero_present = BK_MISSING_ENABLED && P.n_subset > 0;
  if ((P.nb > 0 || zero_present) && threadIdx.x >= 32) {
    // large N: the bucket counts of every row that gets new member counts this epoch start from zero (the workers add);
    // SubsetSplit: so do the row's category presence masks (the workers OR)
    const int wj = (int)(threadIdx.x >> 5) - 1, nw = (BK_CTRL_THREADS >> 5) - 1;
    for (int ji = wj; ji < nj; ji += nw) {
      const Job& jb = sh.jobs[ji];
      if (jb.next_node < 0) continue;
      const size_t rr = (size_t)c * P.R + (jb.kind == BK_JOB_PARTITION ? jb.dst_row : jb.src_row);
      if (P.nb > 0) {
        unsigned* cz = P.coarse + rr * P.nb_stride;
        for (int i = lane; i < P.nb_stride; i += 32) cz[i] = 0u;
      }
      if (zero_present && lane < BK_MAX_SUBSET_COLS) P.present[rr * BK_MAX_SUBSET_COLS + lane] = 0u;
    }
  }
  if (threadIdx.x == 0) {
    hot->n_jobs = nj; hot->n_grow = tot_g;
    hot->c_grow += tot_g - sh.n_nan_fail;   // (a partition job grows its particle unless its split value was cancelled)
    hot->c_grow_root += sh.warp_grow_root[0] + sh.warp_grow_root[1] + sh.warp_grow_root[2] + sh.warp_grow_root[3] - sh.n_fail_root;
    hot->c_count_passes += sh.n_cnt_jobs;
    if (sh.err_bits) hot->c_err |= sh.err_bits;
  }
  TSUB(6);
  return nj;
}

__device__ __forceinline__ void copy_particles(const Params& P, int c, int buf, const int* anc_of_slot /* smem, [P] */) {
  const int warp = (int)(threadIdx.x >> 5) - 1, nwarps = (BK_CTRL_THREADS >> 5) - 1, lane = threadIdx.x & 31;   // warps 1..15
  for (int s = warp; s >= 0 && s < P.P; s += nwarps) {
    const PRef src = pref(P, c, buf, anc_of_slot[s]);
    const PRef dst = pref(P, c, buf ^ 1, s);
    const int nn = src.h->n_nodes;
    const int nfast = nn < P.fastF ? nn : P.fastF;
    const uint4* s4 = reinterpret_cast<const uint4*>(src.h);
    uint4* d4 = reinterpret_cast<uint4*>(dst.h);
    const int words = 2 + nfast * 4;   // 32-byte header + 64-byte nodes, contiguous in shared memory
    for (int i = lane; i < words; i += 32) d4[i] = s4[i];
    if (nn > P.fastF) {                // overflow nodes live in global memory
      const uint4* g4 = reinterpret_cast<const uint4*>(src.slow + P.fastF);
      uint4* h4 = reinterpret_cast<uint4*>(dst.slow + P.fastF);
      for (int i = lane; i < (nn - P.fastF) * 4; i += 32) h4[i] = g4[i];
    }
  }
  CTRL_SYNC();
}

template <int MODE>
__device__ __forceinline__ void finish_tree(const Params& P, int c, ChainCtl* ctl, ChainHot* hot, CtlShared& sh) {
  const int buf = hot->buf, t = hot->cur_tree;
  const uint32_t S0 = P.seed, C0 = P.chain_base + (uint32_t)(c / P.G), G0 = (uint32_t)(c % P.G), D0 = (uint32_t)hot->draw;
  if (BK_WTID >= 0 && BK_WTID < P.P) sh.lw[BK_WTID] = pref(P, c, buf, BK_WTID).h->lw;
  CTRL_SYNC();
  const uint32_t uf = bk_rng(S0, C0, D0, G0, (uint32_t)t, 0xFFFFu, 0, BK_U_FINAL).v[0];
  normalise_and_resample(P, sh, 0, P.P, uf);
  if (threadIdx.x == 0) {
    unsigned pick = bk_index(bk_rng(S0, C0, D0, G0, (uint32_t)t, 0xFFFFu, 0, BK_U_PICK).v[0], (uint32_t)P.P);
    sh.pick = pick; sh.win = sh.anc[pick];
  }
  CTRL_SYNC();
  const int win = sh.win;
  const PRef W = pref(P, c, buf, win);
  DNode* ft = P.forest + ((size_t)c * P.m + t) * BK_MAX_NODES;
  const int old_nn = P.forest_nn[(size_t)c * P.m + t];
  const int new_nn = W.h->n_nodes;
  for (int k = BK_WTID; k >= 0 && k < 256; k += BK_WTHREADS) {
    const int ov = k < BK_MAX_NODES ? __ldcg(&ft[k].var) : 0;        // independent loads (slots beyond the tree are masked)
    const float of = k < BK_MAX_NODES ? __ldcg(&ft[k].value) : 0.0f;
    ctl->old_vals[k] = (k < old_nn && ov < 0) ? of : 0.0f;
    ctl->new_vals[k] = (k < new_nn && W.node(k).var < 0) ? W.node(k).value : 0.0f;
    if (BK_IS_MULTI(P)) {
      for (int j = 0; j < P.K; ++j) {
        const float ofj = (k < BK_MAX_NODES && j > 0) ? __ldcg(reinterpret_cast<const float*>(ft[k].aux) + (j - 1)) : of;
        ctl->old_vals_k[j][k] = (k < old_nn && ov < 0) ? ofj : 0.0f;
        ctl->new_vals_k[j][k] = (k < new_nn && W.node(k).var < 0) ? node_val(W.node(k), j) : 0.0f;
      }
    }
  }
  CTRL_SYNC();
  {
    uint4* d4 = reinterpret_cast<uint4*>(ft);
    for (int i = BK_WTID; i >= 0 && i < new_nn * 4; i += BK_WTHREADS) d4[i] = reinterpret_cast<const uint4*>(&W.node(i >> 2))[i & 3];
  }
  zero_acc0(P, c);
  {
    // split-variable usage of the new tree.  The cumulative prior is rebuilt from the counts BEFORE this tree's
    // (oracle order); the running sums stay sequential in thread 0, the p divisions run in parallel.  The counts
    // themselves are exact integer increments, so parallel atomics give the oracle's sequential result.
    double* av = P.alpha_vec + (size_t)c * P.p;
    const bool rebuild = hot->tune && hot->iter > P.m;
    __shared__ double s_tot;
    if (rebuild) {
      if (P.p <= BK_CUM_SMEM) {
        if (threadIdx.x == 0) {
          double run = 0.0;
          for (int v = 0; v < P.p; ++v) { run = BK_DADD(run, av[v]); sh.cum_prior[v] = run; }
          s_tot = run;
        }
        CTRL_SYNC();
        double* cum = P.cum + (size_t)c * P.p;
        for (int v = BK_WTID; v >= 0 && v < P.p; v += BK_WTHREADS) cum[v] = BK_DDIV(sh.cum_prior[v], s_tot);
      } else if (threadIdx.x == 0) {
        rebuild_cum_dev(P, c);
      }
      CTRL_SYNC();   // the sums above read the counts before anybody bumps them
    }
    for (int k = BK_WTID; k >= 0 && k < new_nn; k += BK_WTHREADS) {
      const int v = W.node(k).var;
      if (v >= 0) { if (hot->tune) atomicAdd(&av[v], 1.0); else atomicAdd(step_vi(P, hot->step_in_launch, c) + v, 1); }
    }
  }
  if (threadIdx.x == 0) {
    P.forest_nn[(size_t)c * P.m + t] = new_nn;
    SweepJob sj; memset(&sj, 0, sizeof(sj));
    sj.do_commit = 1; sj.commit_tree = t; sj.new_row = W.h->row; sj.do_welford = hot->tune ? 1 : 0;
    sj.wf_count = hot->wf_count + (hot->tune ? 1 : 0);
    sj.do_prologue = (t + 1 < hot->tree_hi) ? 1 : 0; sj.prologue_tree = t + 1;
    sj.draw_slot = (P.draws_out && !sj.do_prologue) ? hot->step_in_launch + 1 : 0;   // the step's last commit also keeps the draw
    ctl->sweep = sj;
    hot->cmd = BK_CMD_SWEEP; hot->stage_next = BK_ST_WAIT_SWEEP;
    // kind-2 trace record is completed after the sweep (leaf_sd); stash its fields now
    bk_trace_rec* rec = trace_at(P, c, hot->trace_round_base);
    if (rec) {
      bk_trace_rec r; memset(&r, 0, sizeof(r));
      r.kind = 2; r.tree = t; r.round = hot->round; r.particle = win; r.node = new_nn; r.var = -1;
      r.ancestor = (int32_t)sh.pick; r.log_w = W.h->lw;
      *rec = r;
    }
  }
  CTRL_SYNC();
}

template <int MODE>
__device__ __forceinline__ void control_step(const Params& P, int c, int phase, int tune, const StepArgs& A, ChainHot* hot, CtlShared& sh) {
  const int first_phase = phase == 0;
  ChainCtl* ctl = P.ctl + c;
  if (first_phase) {
    if (threadIdx.x == 0) {
      hot->stage = BK_ST_START; hot->stage_next = BK_ST_START; hot->tune = tune; hot->sigma = A.sigma[c]; hot->step_in_launch = 0;
      hot->ll_inv2s2 = bk_normal_inv2s2(hot->sigma); hot->ll_c = bk_normal_const(hot->sigma, (double)P.N);
    }
    CTRL_SYNC();
  }
  // hot lives in shared memory.  Every thread reads the stage here; thread 0 records the NEXT stage in stage_next and
  // control_loop moves it into `stage` between two barriers, so no thread can see it change under its feet.
  const int stage = hot->stage;
  if (threadIdx.x == 0) { hot->t_sub_last = globaltimer_ns(); if (first_phase) for (int i = 0; i < 8; ++i) hot->t_sub[i] = 0; }
  MARK(100 + stage);
  if (stage == BK_ST_DONE) return;
  if (threadIdx.x == 0) hot->c_phases += 1;

  // first phase of a step (of the launch, or right after the previous step's last commit): clear the step's counters,
  // pick the tree batch, publish the prologue of its first tree
  auto begin_step = [&]() {
    int32_t* vis = step_vi(P, hot->step_in_launch, c);
    for (int v = BK_WTID; v >= 0 && v < P.p; v += BK_WTHREADS) vis[v] = 0;
    zero_acc0(P, c);
    if (threadIdx.x == 0) {
      int T = hot->tune ? P.batch_tune : P.batch_post;
      int lo = hot->lower, hi = lo + T < P.m ? lo + T : P.m;
      hot->tree_lo = lo; hot->tree_hi = hi; hot->cur_tree = lo;
      hot->c_tree_updates = 0; hot->c_rounds = 0; hot->c_grow = 0; hot->c_grow_root = 0;
      hot->c_count_passes = 0; hot->c_phases = 1; hot->c_err = 0;
      hot->trace_len = 0; hot->trace_round_base = 0;
      SweepJob sj; memset(&sj, 0, sizeof(sj));
      sj.do_prologue = 1; sj.prologue_tree = lo;
      ctl->sweep = sj; hot->cmd = BK_CMD_SWEEP; hot->stage_next = BK_ST_WAIT_SWEEP;
    }
    CTRL_SYNC();
  };
  if (stage == BK_ST_START) { begin_step(); return; }

  // `closing` = a round of the current tree has just completed (its epoch is done); false right after init_particles
  bool closing = false;
  if (stage == BK_ST_WAIT_SWEEP) {
    __shared__ int s_more;
    if (threadIdx.x == 0) {
      const SweepJob sj = ctl->sweep;
      if (sj.do_commit) {
        if (hot->tune) {
          hot->wf_count = sj.wf_count;
          if (hot->iter > 2) {
            if (BK_IS_MULTI(P)) {
              for (int j = 0; j < P.K; ++j) {
                const long long sd_j = (long long)__ldcg(P.acc_sd + (size_t)c * 8 + j);
                hot->leaf_sdk[j] = (float)BK_DDIV(BK_DMUL((double)sd_j, P.inv_qscale), (double)P.N);
              }
              hot->leaf_sd = hot->leaf_sdk[0];
            } else {
              long long sd_sum = (long long)__ldcg(P.acc0 + (size_t)c * BK_ACC0_WORDS + (size_t)256 * BK_ACC0_STRIDE);
              hot->leaf_sd = (float)BK_DDIV(BK_DMUL((double)sd_sum, P.inv_qscale), (double)P.N);
            }
          }
          if (BK_IS_MULTI(P)) for (int j = 0; j < 8; ++j) P.acc_sd[(size_t)c * 8 + j] = 0ull;   // (ordered before the next SWEEP by the publication's release fence)
        }
        hot->c_tree_updates += 1;
        bk_trace_rec* rec = trace_at(P, c, hot->trace_round_base);
        if (rec) rec->aux = (double)(BK_IS_MULTI(P) ? hot->leaf_sdk[P.K - 1] : hot->leaf_sd);
        hot->trace_round_base += 1;
      }
      if (sj.do_prologue) {
        hot->iter += 1; hot->cur_tree = sj.prologue_tree; s_more = 1;
      } else {
        hot->lower = hot->tree_hi < P.m ? hot->tree_hi : 0;
        hot->draw += 1;
        s_more = (hot->step_in_launch + 1 < A.n_steps) ? 2 : 0;     // 2: this chain's next step starts right away
        if (!s_more) { hot->cmd = BK_CMD_DONE; hot->stage_next = BK_ST_DONE; }
        bk_step_stats st; memset(&st, 0, sizeof(st));
        st.tree_updates = hot->c_tree_updates; st.rounds = hot->c_rounds; st.grow_events = hot->c_grow;
        st.grow_root = hot->c_grow_root; st.count_passes = hot->c_count_passes; st.phases = hot->c_phases;
        st.trace_len = hot->trace_round_base; st.error_flags = hot->c_err | (hot->trace_round_base > P.trace_cap && P.trace_cap > 0 ? 1 : 0);
        st.leaf_sd = hot->leaf_sd; st.iter = hot->iter;
        *step_stats(P, hot->step_in_launch, c) = st;
        if (s_more) hot->step_in_launch += 1;
      }
    }
    CTRL_SYNC();
    if (!s_more) return;
    if (s_more == 2) { begin_step(); return; }
    MARK(110);
    init_particles(P, c, ctl, hot, sh);
    TSUB(7);
    MARK(111);
  } else if (stage == BK_ST_WAIT_ROUND) {
    MARK(120);
    finalize_own<MODE>(P, c, hot, sh, hot->buf, hot->round, hot->cur_tree, hot->trace_round_base);
    TSUB(0);
    MARK(121);
    if (P.lik != BK_LIK_NORMAL && hot->n_grow > 0) {
      // leaf values are known now: publish the LL jobs (the first n_grow list entries) and wait for their sums
      CTRL_SYNC();
      const int ng = hot->n_grow;
      const uint4* s4 = reinterpret_cast<const uint4*>(sh.jobs);
      for (int i = BK_WTID; i >= 0 && i < ng * 3; i += BK_WTHREADS) reinterpret_cast<uint4*>(ctl->jobs[0])[i] = s4[i];
      if (BK_IS_MULTI(P))
        for (int i = BK_WTID; i >= 0 && i < ng * 2 * BK_MAX_OUTPUTS; i += BK_WTHREADS) (&ctl->job_vals[0][0][0])[i] = (&sh.job_vals[0][0][0])[i];
      if (threadIdx.x == 0) { hot->n_jobs = ng; hot->cmd = BK_CMD_LL; hot->stage_next = BK_ST_WAIT_LL; }
      CTRL_SYNC();
      return;
    }
    closing = true;
  } else {  // BK_ST_WAIT_LL
    finalize_ll_own(P, c, sh, hot->buf);
    TSUB(0);
    closing = true;
  }

  for (;;) {
    // Every thread works on values it read BEFORE the barrier below; thread 0 moves the shared scalars on after it.
    const int buf = hot->buf, t = hot->cur_tree;
    const int round_done = hot->round, rbase = hot->trace_round_base;
    const int q = BK_WTID;
    const bool is_p = q >= 1 && q < P.P;
    if (closing && is_p) {   // the round `round_done` is complete: log weights and liveness (own particle, finalized above)
      const PRef S = pref(P, c, buf, q);
      if (S.h->q_head < S.h->n_nodes) sh.live = 1;
      bk_trace_rec* rec = trace_at(P, c, rbase + q - 1);
      if (rec) rec->log_w = S.h->lw;
    }
    if (q >= 0 && q < 8) sh.used_rows[q] = 0u;
    if (q == 0) { sh.n_cnt_jobs = 0; sh.next_sel = 1; sh.err_bits = 0; sh.n_nan_fail = 0; sh.n_fail_root = 0; }
    CTRL_SYNC();
    CTS(17, 0);
    const int live = sh.live;
    int round = 0, rb = rbase, src = q;
    bool deferred = false;
    if (closing) {
      if (threadIdx.x == 0) { hot->c_rounds += 1; hot->trace_round_base = rbase + (P.P - 1); }
      MARK(130 + live);
      TSUB(1);
      if (!live) { CTRL_SYNC(); finish_tree<MODE>(P, c, ctl, hot, sh); TSUB(7); MARK(139); return; }
      round = round_done + 1; rb = rbase + (P.P - 1); deferred = true;
      if (q >= 0 && q < 128) {   // warps 1..4: the particle threads resample for themselves
        const uint32_t u = (sh.pre_z_tree == t && sh.pre_z_round == round_done)
                               ? sh.pre_ures
                               : bk_rng(P.seed, P.chain_base + (uint32_t)(c / P.G), (uint32_t)hot->draw, (uint32_t)(c % P.G),
                                        (uint32_t)t, (uint32_t)round_done, 0, BK_U_RESAMPLE).v[0];
        const int anc = resample_own(P, c, buf, sh, 1, P.P - 1, u);
        src = is_p ? anc + 1 : 0;
        if (is_p) { bk_trace_rec* rec = trace_at(P, c, rbase + q - 1); if (rec) rec->ancestor = src; }   // (record of the round just closed)
        CTS(18, src);
      }
      TSUB(2);
    }
    MARK(140);
    const int nj = open_round<MODE>(P, c, ctl, hot, sh, buf, round, rb, deferred, src);
    MARK(141);
    if (threadIdx.x == 0) {   // (every thread took its copies of these before the barriers inside open_round)
      hot->round = round; sh.live = 0;
      if (deferred) sh.copy_pending = 1;
    }
    if (nj > 0) {
      if (threadIdx.x == 0) { hot->cmd = BK_CMD_ROUND; hot->stage_next = BK_ST_WAIT_ROUND; }
      return;   // (control_loop's barrier follows)
    }
    CTRL_SYNC();
    apply_pending_copy(P, c, hot, sh);   // no epoch to hide behind: the next round starts right away
    closing = true;
  }
}

// ------------------------------------------------------------------ data phase: ROUND
// Byte-parallel helpers: leaf ids are one byte, a lane holds 8 of them in two 32-bit words.
// 0x80 in every byte of w that equals the corresponding byte of pat4
__device__ __forceinline__ unsigned bytes_eq_msb(unsigned w, unsigned pat4) {
  const unsigned x = w ^ pat4;
  const unsigned t = ((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x;   // MSB set where the byte differs
  return ~t & 0x80808080u;
}
// 0xFF in every such byte (PRMT replicates the MSB over the byte)
__device__ __forceinline__ unsigned bytes_eq(unsigned w, unsigned pat4) {
  unsigned d;   // prmt selector nibble 8+i = byte i with its sign bit replicated (__byte_perm ignores that mode bit)
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(bytes_eq_msb(w, pat4)), "r"(0u), "r"(0xBA98u));
  return d;
}
// Shared-memory accumulator of one job: 32-bit limbs so that a warp adds ALL its statistics with one native
// shared-memory atomic instruction (lane k adds limb k; 64-bit shared atomics would be CAS loops).  The limbs are the
// REDUX partial sums themselves: a signed 48-bit sum v is (hi, lo) = (sum of v >> 16, sum of v & 0xFFFF), the
// 64-bit sum of squares is four 16-bit chunks.  A limb absorbs < 2^11 warp contributions per epoch without overflow
// (bk_create caps n_rows accordingly); the flush recombines them exactly.
#define BK_LIMB_N 0
#define BK_LIMB_ST_LO 1
#define BK_LIMB_ST_HI 2
#define BK_LIMB_SR_LO 3
#define BK_LIMB_SR_HI 4
#define BK_LIMB_ND 5        // rows dropped from the node by a missing covariate: count, sum q(sum_trees), sum q(r)
#define BK_LIMB_STD_LO 6
#define BK_LIMB_STD_HI 7
#define BK_LIMB_SRD_LO 8
#define BK_LIMB_SRD_HI 13
#define BK_LIMB_LLL_LO 9
#define BK_LIMB_LLL_HI 10
#define BK_LIMB_LLR_LO 11
#define BK_LIMB_LLR_HI 12
#define BK_LIMBS 16   // per job (64 bytes)
// statistic k (BK_ACC_* index) of a job from its limbs; the limbs of a statistic belong to it alone and are cleared
__device__ __forceinline__ unsigned long long take_stat(unsigned* L, int k) {
  int lo, hi;   // limb indices; hi < 0: single limb
  bool sgn = true;
  switch (k) {
    case BK_ACC_N: lo = BK_LIMB_N; hi = -1; break;
    case BK_ACC_SST: lo = BK_LIMB_ST_LO; hi = BK_LIMB_ST_HI; break;
    case BK_ACC_SR: lo = BK_LIMB_SR_LO; hi = BK_LIMB_SR_HI; break;
    case BK_ACC_ND: lo = BK_LIMB_ND; hi = -1; break;
    case BK_ACC_SSTD: lo = BK_LIMB_STD_LO; hi = BK_LIMB_STD_HI; break;
    case BK_ACC_SRD: lo = BK_LIMB_SRD_LO; hi = BK_LIMB_SRD_HI; break;
    case BK_ACC_LLL: lo = BK_LIMB_LLL_LO; hi = BK_LIMB_LLL_HI; break;
    case BK_ACC_LLR: lo = BK_LIMB_LLR_LO; hi = BK_LIMB_LLR_HI; break;
    default: return 0ull;
  }
  const unsigned vlo = L[lo];
  L[lo] = 0u;
  if (hi < 0) return (unsigned long long)vlo;
  const unsigned vhi = L[hi];
  L[hi] = 0u;
  return sgn ? (unsigned long long)((long long)(int)vhi * 65536ll + (long long)vlo)
             : (unsigned long long)vlo + ((unsigned long long)vhi << 16);
}
// byte e (0..3) of a 0x00/0xFF byte mask widened to a 32-bit mask
#define BK_ROWMASK(m, e) __byte_perm((m), 0u, 0x1111u * (e))

// MISSING: X holds NaNs somewhere (Params::has_nan).  A separate instantiation so that the path without missing values
// keeps its register allocation (the extra masks and sums cost the C5 step 12 % when compiled into one body).
// MULTI: shared-tree multi-output (Params::K > 1): per job the sums of q(sum_trees[j]) of BOTH children for every
// output j go to `acck` (no parent statistics are kept per output); no Gaussian residual statistics (the weight comes
// from the LL epoch).
// ------------------------------------------------------------------ staged loads of the ROUND epoch (cp.async)
// A (tile, job) unit is a dependent chain: leaf ids + column tile (an L2 / HBM round trip), then ~250 instructions of
// routing and sums.  With 32 resident warps per SM at 64 registers each, the loads of job j+1 cannot live in registers
// beside the work of job j (tried: the spills cost more than the latency), so they go to SHARED memory instead: while a
// warp works on job j, the leaf ids and the column tile of job j+1 are in flight as asynchronous copies
// (cp.async = LDGSTS) into a slot the same lane reads back itself — every lane copies and consumes its own 40 bytes,
// so no cross-lane synchronisation is needed, only cp.async.wait_group.  Two slots per lane (ping-pong): the slot being
// refilled was consumed a whole job earlier.  Per warp 2 x 1280 B, per group of 8 warps 20 KB.  Worker CTAs have that
// memory idle: groups 0-1 use the static area that holds the control state in control CTAs, groups 2-3 the tail of the
// dynamic area.  Leaf ids are copied with .ca (8-byte copies exist only in that form): like the residual tiles they are
// read through an L1 that the epoch's acquire poll has just invalidated, and a row read in an epoch is never written
// in it; the column tile bypasses L1 (.cg).
// MEASURED, NOT ENABLED: compiled only with -DBK_STAGE.  Bit-identical results (40 parity tests), but no gain on a B200:
// C2 6283 -> 6136 draws/s (-2.3 %), C5 at 8 chains per GPU 341 -> 344 (+0.9 %) (profiles/r2_ab_stage_cpasync.txt).  With
// 8 warps per scheduler issuing ~52 % of the cycles, one job's 250 instructions already take longer than the load
// latency they would hide; the units are bound by their instruction count, not by the round trip.
#define BK_STAGE_SLOT_BYTES 1280                                  // [32 lanes x 16 B x0][32 x 16 B x1][32 x 8 B ids]
#define BK_STAGE_WARP_BYTES (2 * BK_STAGE_SLOT_BYTES)
#define BK_STAGE_GROUP_BYTES ((BK_GROUP_THREADS / 32) * BK_STAGE_WARP_BYTES)
__device__ __forceinline__ void cp_async_8(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
// start the copies of one job's inputs into `slot` (this lane's part): leaf ids unless the source is a stump, the
// column tile for a partition of a dense node (sparse nodes load it only in lanes that hold members, after the ids)
__device__ __forceinline__ void stage_issue(unsigned char* slot, const Job* __restrict__ jb, const uint8_t* rows_c, const float* x_b,
                                            unsigned uNpad, int lane) {
  const int4* jp = reinterpret_cast<const int4*>(jb);
  const int4 j0 = jp[0];
  const int kind = j0.x, src_row = j0.z;
  if (kind == BK_JOB_PARTITION || kind == BK_JOB_COUNT) {
    if (src_row != BK_ROW_VIRTUAL) cp_async_8(slot + 1024 + lane * 8, rows_c + (unsigned long long)(unsigned)src_row * uNpad);
    if (kind == BK_JOB_PARTITION && jp[2].z == 0) {
      const float* xp = x_b + (unsigned long long)(unsigned)jp[1].y * uNpad;
      cp_async_16(slot + lane * 16, xp);
      cp_async_16(slot + 512 + lane * 16, xp + 4);
    }
  }
  cp_async_commit();
}

// SubsetSplit (feature instantiation only).  subset_left: bk_subset_left with the node's set already decoded.
__device__ __forceinline__ bool subset_left(float x, unsigned set) {
  const int cd = bk_subset_code(x);
  return cd >= 0 && ((set >> cd) & 1u);
}
// The categories present among the members of a row's counted node (m0 / m1: 0x80 in the bytes of the member rows of
// this lane), for every subset column: OR over the tile, one global atomic per (tile, column) that found any.  The
// control CTA zeroed the row's masks before it published the epoch and reads them when it pops that node.
__device__ __forceinline__ void note_present(const Params& P, unsigned uCR, int row, unsigned m0, unsigned m1, const float* x_b, unsigned uNpad,
                                             int lane) {
  const bool anym = (m0 | m1) != 0u;
  for (int sc = 0; sc < P.n_subset; ++sc) {
    unsigned bits = 0u;
    if (anym) {
      const float4* xp = reinterpret_cast<const float4*>(x_b + (unsigned long long)(unsigned)__ldg(P.subset_cols + sc) * uNpad);
      const float4 a = __ldg(xp), b = __ldg(xp + 1);
      const float xs[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (m0 & (0x80u << (8 * e))) { const int cd = bk_subset_code(xs[e]); if (cd >= 0) bits |= 1u << cd; }
        if (m1 & (0x80u << (8 * e))) { const int cd = bk_subset_code(xs[4 + e]); if (cd >= 0) bits |= 1u << cd; }
      }
    }
    bits = __reduce_or_sync(0xffffffffu, bits);
    if (lane == 0 && bits) atomicOr(P.present + (unsigned long long)(uCR + (unsigned)row) * BK_MAX_SUBSET_COLS + sc, bits);
  }
}

template <bool MISSING, bool MULTI>
__device__ __forceinline__ void round_unit(const Params& P, int c, int tile, int job_lo, int job_hi, const Job* __restrict__ sjobs,
                                           unsigned* __restrict__ sacc, unsigned (*__restrict__ acck)[BK_MAX_OUTPUTS][4],
                                           unsigned char* __restrict__ stage) {
  const int lane = threadIdx.x & 31;
  const size_t base = (size_t)tile * BK_WARP_TILE + (size_t)lane * BK_ROWS_PER_LANE;
  const bool gauss = !MULTI && P.lik == BK_LIK_NORMAL;
#ifdef BK_STAGE
  // the first job's inputs start before the residual tiles are asked for (one round trip for both)
  {
    const unsigned uNp = (unsigned)P.Npad;
    stage_issue(stage, &sjobs[job_lo], P.rows + (unsigned long long)((unsigned)c * (unsigned)P.R) * uNp + base, P.X + base, uNp, lane);
  }
  int cur = 0;
#endif
  int q_r[8], q_s[8];
  if (MULTI) {
#pragma unroll
    for (int e = 0; e < 8; ++e) { q_r[e] = 0; q_s[e] = 0; }
  } else {
    const int4* a = reinterpret_cast<const int4*>(P.qr + (size_t)c * P.Npad + base);
    const int4* b = reinterpret_cast<const int4*>(P.qst + (size_t)c * P.Npad + base);
    // L1-cached on purpose: the warps of this CTA share a few tiles (fresh after the epoch's acquire fence)
    int4 a0 = ld_ca_v4(a), a1 = ld_ca_v4(a + 1), b0 = ld_ca_v4(b), b1 = ld_ca_v4(b + 1);
    q_r[0] = a0.x; q_r[1] = a0.y; q_r[2] = a0.z; q_r[3] = a0.w; q_r[4] = a1.x; q_r[5] = a1.y; q_r[6] = a1.z; q_r[7] = a1.w;
    q_s[0] = b0.x; q_s[1] = b0.y; q_s[2] = b0.z; q_s[3] = b0.w; q_s[4] = b1.x; q_s[5] = b1.y; q_s[6] = b1.z; q_s[7] = b1.w;
  }
  // per-unit address bases (the per-job part is one multiply-add)
  // Addresses as base pointer + (unsigned 32 x 32 -> 64) products: one IMAD.WIDE.U32 each.  With signed `int * size_t` the
  // compiler spent ~70 of the ~250 instructions of a (tile, job) unit on 64-bit address arithmetic (sign extensions, two
  // 64-bit multiplies per address), recomputed per job because the 64-register cap leaves no room to keep them.
  const unsigned uNpad = (unsigned)P.Npad, uCnt = (unsigned)P.cnt_stride, uCR = (unsigned)c * (unsigned)P.R;
  const uint8_t* rows_c = P.rows + (unsigned long long)uCR * uNpad + base;
  const float* x_b = P.X + base;
  unsigned* cnt_c = P.rowcnt + (unsigned long long)uCR * uCnt + (unsigned)tile;
  // leaf ids of a stump: 0 for real rows, 0xFF (limbo) for the padding rows of the last tile
  unsigned vw0 = 0u, vw1 = 0u;
  if (base + 8 > (size_t)P.N) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (base + e >= (size_t)P.N) vw0 |= 0xFFu << (8 * e);
      if (base + 4 + e >= (size_t)P.N) vw1 |= 0xFFu << (8 * e);
    }
  }
  for (int ji = job_lo; ji < job_hi; ++ji) {
    const int4* jp = reinterpret_cast<const int4*>(&sjobs[ji]);   // the group's shared-memory copy of the job list
    const int4 j0 = jp[0], j1 = jp[1], j2 = jp[2];
    const int kind = j0.x, src_row = j0.z, dst_row = j0.w;
    const int node = j1.x, var = j1.y; const float split = __int_as_float(j1.z); const int left_id = j1.w;
    const int next_node = j2.x, rule = j2.y, sparse = j2.z;
    const bool has_nan = MISSING && j2.w != 0;
    unsigned w0 = vw0, w1 = vw1;
#ifdef BK_STAGE
    // this job's inputs were copied into the current slot while the previous job was worked on; the next job's go
    // into the other slot now
    float4 xs0 = make_float4(0.f, 0.f, 0.f, 0.f), xs1 = xs0;
    {
      cp_async_wait_all();
      const unsigned char* sl = stage + cur * BK_STAGE_SLOT_BYTES;
      if ((kind == BK_JOB_PARTITION || kind == BK_JOB_COUNT) && src_row != BK_ROW_VIRTUAL) {
        const uint2 v = *reinterpret_cast<const uint2*>(sl + 1024 + lane * 8);
        w0 = v.x; w1 = v.y;
      }
      if (kind == BK_JOB_PARTITION && !sparse) {
        xs0 = *reinterpret_cast<const float4*>(sl + lane * 16);
        xs1 = *reinterpret_cast<const float4*>(sl + 512 + lane * 16);
      }
      cur ^= 1;
      if (ji + 1 < job_hi) stage_issue(stage + cur * BK_STAGE_SLOT_BYTES, &sjobs[ji + 1], rows_c, x_b, uNpad, lane);
    }
#else
    if (src_row != BK_ROW_VIRTUAL) {
      const uint2 v = __ldcg(reinterpret_cast<const uint2*>(rows_c + (unsigned long long)(unsigned)src_row * uNpad));
      w0 = v.x; w1 = v.y;
    }
#endif
    const unsigned next4 = (unsigned)next_node * 0x01010101u;
    if (kind == BK_JOB_PARTITION) {
      const unsigned node4 = (unsigned)node * 0x01010101u;
      const unsigned mem0 = bytes_eq(w0, node4), mem1 = bytes_eq(w1, node4);
      unsigned lb0 = 0u, lb1 = 0u;
      // dense nodes: the column load is issued together with the leaf-id load (one L2/HBM round trip
      // per job); sparse nodes (few members) keep it dependent on the ids to save the bytes
      float4 x0 = make_float4(0.f, 0.f, 0.f, 0.f), x1 = x0;
      if (!sparse || (mem0 | mem1)) {
#ifdef BK_STAGE
        if (!sparse) { x0 = xs0; x1 = xs1; }
        else
#endif
        {
          const float4* xp = reinterpret_cast<const float4*>(x_b + (unsigned long long)(unsigned)var * uNpad);
          x0 = __ldg(xp); x1 = __ldg(xp + 1);
        }
        if (MISSING && rule == BK_RULE_SUBSET) {   // the node's set of left-going categories travels as the float of its bit mask
          const unsigned set = (unsigned)split;
          if (subset_left(x0.x, set)) lb0 |= 0x000000FFu; if (subset_left(x0.y, set)) lb0 |= 0x0000FF00u;
          if (subset_left(x0.z, set)) lb0 |= 0x00FF0000u; if (subset_left(x0.w, set)) lb0 |= 0xFF000000u;
          if (subset_left(x1.x, set)) lb1 |= 0x000000FFu; if (subset_left(x1.y, set)) lb1 |= 0x0000FF00u;
          if (subset_left(x1.z, set)) lb1 |= 0x00FF0000u; if (subset_left(x1.w, set)) lb1 |= 0xFF000000u;
        } else if (rule == BK_RULE_ONEHOT) {
          if (x0.x == split) lb0 |= 0x000000FFu; if (x0.y == split) lb0 |= 0x0000FF00u;
          if (x0.z == split) lb0 |= 0x00FF0000u; if (x0.w == split) lb0 |= 0xFF000000u;
          if (x1.x == split) lb1 |= 0x000000FFu; if (x1.y == split) lb1 |= 0x0000FF00u;
          if (x1.z == split) lb1 |= 0x00FF0000u; if (x1.w == split) lb1 |= 0xFF000000u;
        } else {
          if (x0.x <= split) lb0 |= 0x000000FFu; if (x0.y <= split) lb0 |= 0x0000FF00u;
          if (x0.z <= split) lb0 |= 0x00FF0000u; if (x0.w <= split) lb0 |= 0xFF000000u;
          if (x1.x <= split) lb1 |= 0x000000FFu; if (x1.y <= split) lb1 |= 0x0000FF00u;
          if (x1.z <= split) lb1 |= 0x00FF0000u; if (x1.w <= split) lb1 |= 0xFF000000u;
        }
      }
      const unsigned lm0 = lb0 & mem0, lm1 = lb1 & mem1;           // 0xFF where the row goes to the left child
      const unsigned L4 = (unsigned)left_id * 0x01010101u, R4 = L4 + 0x01010101u;
      unsigned n0 = (w0 & ~mem0) | (mem0 & ((lm0 & L4) | (~lm0 & R4)));
      unsigned n1 = (w1 & ~mem1) | (mem1 & ((lm1 & L4) | (~lm1 & R4)));
      if (has_nan) {
        // Missing covariates: a member whose x is NaN leaves the tree (leaf id BK_LIMBO, predicts 0)
        // and counts for neither child; its statistics go to the job's "dropped" accumulators so that the right
        // child is parent - left - dropped.  (NaN compares false, so such a row never went left above.)
        unsigned d0 = 0u, d1 = 0u;
        if (!sparse || (mem0 | mem1)) {
          if (x0.x != x0.x) d0 |= 0x000000FFu; if (x0.y != x0.y) d0 |= 0x0000FF00u;
          if (x0.z != x0.z) d0 |= 0x00FF0000u; if (x0.w != x0.w) d0 |= 0xFF000000u;
          if (x1.x != x1.x) d1 |= 0x000000FFu; if (x1.y != x1.y) d1 |= 0x0000FF00u;
          if (x1.z != x1.z) d1 |= 0x00FF0000u; if (x1.w != x1.w) d1 |= 0xFF000000u;
        }
        d0 &= mem0; d1 &= mem1;
        n0 |= d0; n1 |= d1;                                          // 0xFF = BK_LIMBO
        if (__any_sync(0xffffffffu, (d0 | d1) != 0u)) {
          int s_a = 0, r_a = 0;
          long long ll = 0;
          if (gauss) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int ma = (int)BK_ROWMASK(d0, e), mb = (int)BK_ROWMASK(d1, e);
              s_a += (q_s[e] & ma) + (q_s[4 + e] & mb); r_a += (q_r[e] & ma) + (q_r[4 + e] & mb);
            }
          } else {   // Bernoulli: q_r holds noi; the dropped rows' terms at the value 0 (rare path: y is read here)
            const float4* yp = reinterpret_cast<const float4*>(P.y + (size_t)(c % P.G) * P.Npad + base);
            const float4 y0 = __ldg(yp), y1 = __ldg(yp + 1);
            const float yy[8] = {y0.x, y0.y, y0.z, y0.w, y1.x, y1.y, y1.z, y1.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              s_a += (q_s[e] & (int)BK_ROWMASK(d0, e)) + (q_s[4 + e] & (int)BK_ROWMASK(d1, e));
              if (BK_ROWMASK(d0, e)) ll += (long long)bk_bern_q(yy[e], __int_as_float(q_r[e]), 0.0f);
              if (BK_ROWMASK(d1, e)) ll += (long long)bk_bern_q(yy[4 + e], __int_as_float(q_r[4 + e]), 0.0f);
            }
          }
          // (|q| < 2^29 and at most 8 rows per lane: the per-lane sums fit 32 bits; Bernoulli terms < 2^29 each as well)
          const long long st = (long long)s_a, sr = gauss ? (long long)r_a : ll;
          const unsigned cd = (unsigned)(__popc(d0) + __popc(d1)) >> 3;
          unsigned v = __reduce_add_sync(0xffffffffu, cd);
          const unsigned st_lo = __reduce_add_sync(0xffffffffu, (unsigned)st & 0xFFFFu);
          const int st_hi = __reduce_add_sync(0xffffffffu, (int)(st >> 16));
          const unsigned sr_lo = __reduce_add_sync(0xffffffffu, (unsigned)sr & 0xFFFFu);
          const int sr_hi = __reduce_add_sync(0xffffffffu, (int)(sr >> 16));
          v = lane == 1 ? st_lo : v; v = lane == 2 ? (unsigned)st_hi : v; v = lane == 3 ? sr_lo : v; v = lane == 4 ? (unsigned)sr_hi : v;
          if (lane < 5) atomicAdd(sacc + ji * BK_LIMBS + (lane < 4 ? BK_LIMB_ND + lane : BK_LIMB_SRD_HI), v);
        }
      }
      __stcg(reinterpret_cast<uint2*>(const_cast<uint8_t*>(rows_c) + (unsigned long long)(unsigned)dst_row * uNpad), make_uint2(n0, n1));
      if (MULTI) {
        if (__any_sync(0xffffffffu, (mem0 | mem1) != 0u)) {
          const unsigned rm0 = mem0 & ~lm0, rm1 = mem1 & ~lm1;     // rows of the right child
          const unsigned cl = (unsigned)(__popc(lm0) + __popc(lm1)) >> 3;
          const unsigned nl_tot = __reduce_add_sync(0xffffffffu, cl);
          if (lane == 0 && nl_tot) atomicAdd(sacc + ji * BK_LIMBS + BK_LIMB_N, nl_tot);
          for (int j = 0; j < P.K; ++j) {
            const int4* b = reinterpret_cast<const int4*>(P.qst + ((size_t)c * P.K + j) * P.Npad + base);
            const int4 b0 = ld_ca_v4(b), b1 = ld_ca_v4(b + 1);
            const int qs[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
            int sl_a = 0, sr_a = 0;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              sl_a += (qs[e] & (int)BK_ROWMASK(lm0, e)) + (qs[4 + e] & (int)BK_ROWMASK(lm1, e));
              sr_a += (qs[e] & (int)BK_ROWMASK(rm0, e)) + (qs[4 + e] & (int)BK_ROWMASK(rm1, e));
            }
            // (8 rows of |q| < 2^29 fit a 32-bit lane sum; limbs as in the single-output path)
            const unsigned l_lo = __reduce_add_sync(0xffffffffu, (unsigned)sl_a & 0xFFFFu), r_lo = __reduce_add_sync(0xffffffffu, (unsigned)sr_a & 0xFFFFu);
            const int l_hi = __reduce_add_sync(0xffffffffu, sl_a >> 16), r_hi = __reduce_add_sync(0xffffffffu, sr_a >> 16);
            unsigned v = l_lo;
            v = lane == 1 ? (unsigned)l_hi : v; v = lane == 2 ? r_lo : v; v = lane == 3 ? (unsigned)r_hi : v;
            if (lane < 4) atomicAdd(&acck[ji][j][lane], v);
          }
        }
      } else if (__any_sync(0xffffffffu, (lm0 | lm1) != 0u)) {
        // masked per-lane sums: 4 rows fit 32 bits (|q| < 2^29)
        int s_a = 0, s_b = 0, r_a = 0, r_b = 0;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int ma = (int)BK_ROWMASK(lm0, e), mb = (int)BK_ROWMASK(lm1, e);
          s_a += q_s[e] & ma; s_b += q_s[4 + e] & mb;
          if (gauss) { r_a += q_r[e] & ma; r_b += q_r[4 + e] & mb; }
        }
        const unsigned cl = (unsigned)(__popc(lm0) + __popc(lm1)) >> 3;
        const long long st = (long long)s_a + (long long)s_b, sr = (long long)r_a + (long long)r_b;
        // REDUX partial sums = the limbs
        unsigned v = __reduce_add_sync(0xffffffffu, cl);                                          // lane 0: BK_LIMB_N
        const unsigned st_lo = __reduce_add_sync(0xffffffffu, (unsigned)st & 0xFFFFu);
        const int st_hi = __reduce_add_sync(0xffffffffu, (int)(st >> 16));
        v = lane == BK_LIMB_ST_LO ? st_lo : v;
        v = lane == BK_LIMB_ST_HI ? (unsigned)st_hi : v;
        int nl = 3;
        if (gauss) {   // Gaussian sufficient statistics of the residual (Bernoulli: q_r holds noi bits)
          const unsigned sr_lo = __reduce_add_sync(0xffffffffu, (unsigned)sr & 0xFFFFu);
          const int sr_hi = __reduce_add_sync(0xffffffffu, (int)(sr >> 16));
          v = lane == BK_LIMB_SR_LO ? sr_lo : v;
          v = lane == BK_LIMB_SR_HI ? (unsigned)sr_hi : v;
          nl = 5;
        }
        if (lane < nl) atomicAdd(sacc + ji * BK_LIMBS + lane, v);   // one shared-memory atomic instruction per job
      }
      if (next_node >= 0) {
        const unsigned nm0 = bytes_eq_msb(n0, next4), nm1 = bytes_eq_msb(n1, next4);
        const unsigned cnt = (unsigned)(__popc(nm0) + __popc(nm1));
        const unsigned tot = __reduce_add_sync(0xffffffffu, cnt);
        if (lane == 0) {
          cnt_c[(unsigned long long)(unsigned)dst_row * uCnt] = tot;
          if (P.nb > 0 && tot) atomicAdd(P.coarse + (unsigned long long)(uCR + (unsigned)dst_row) * (unsigned)P.nb_stride + ((unsigned)tile / BK_COARSE_TILES), tot);
        }
        if (MISSING && P.n_subset > 0 && tot) note_present(P, uCR, dst_row, nm0, nm1, x_b, uNpad, lane);
      }
    } else if (kind == BK_JOB_COUNT) {
      const unsigned nm0 = bytes_eq_msb(w0, next4), nm1 = bytes_eq_msb(w1, next4);
      const unsigned cnt = (unsigned)(__popc(nm0) + __popc(nm1));
      const unsigned tot = __reduce_add_sync(0xffffffffu, cnt);
      if (lane == 0) {
        cnt_c[(unsigned long long)(unsigned)src_row * uCnt] = tot;
        if (P.nb > 0 && tot) atomicAdd(P.coarse + (unsigned long long)(uCR + (unsigned)src_row) * (unsigned)P.nb_stride + ((unsigned)tile / BK_COARSE_TILES), tot);
      }
      if (MISSING && P.n_subset > 0 && tot) note_present(P, uCR, src_row, nm0, nm1, x_b, uNpad, lane);
    }
  }
}

// ------------------------------------------------------------------ data phase: LL (Bernoulli)
// Same unit shape as ROUND: one warp, 256 rows, a group of jobs.  noi (float bits kept in the qr array) and
// y stay in registers; per job the particle's new leaf-id row is read and the rows of the two new leaves
// contribute their quantised log-likelihood term at the leaf's value (+9N bytes per grow event).
__device__ __forceinline__ void ll_unit(const Params& P, int c, int tile, int job_lo, int job_hi, const Job* __restrict__ sjobs,
                                        unsigned* __restrict__ sacc) {
  const int lane = threadIdx.x & 31;
  const size_t base = (size_t)tile * BK_WARP_TILE + (size_t)lane * BK_ROWS_PER_LANE;
  float noi[8], yv[8];
  {
    const int4* a = reinterpret_cast<const int4*>(P.qr + (size_t)c * P.Npad + base);
    const float4* b = reinterpret_cast<const float4*>(P.y + (size_t)(c % P.G) * P.Npad + base);
    const int4 a0 = __ldcg(a), a1 = __ldcg(a + 1);
    const float4 b0 = __ldg(b), b1 = __ldg(b + 1);
    noi[0] = __int_as_float(a0.x); noi[1] = __int_as_float(a0.y); noi[2] = __int_as_float(a0.z); noi[3] = __int_as_float(a0.w);
    noi[4] = __int_as_float(a1.x); noi[5] = __int_as_float(a1.y); noi[6] = __int_as_float(a1.z); noi[7] = __int_as_float(a1.w);
    yv[0] = b0.x; yv[1] = b0.y; yv[2] = b0.z; yv[3] = b0.w; yv[4] = b1.x; yv[5] = b1.y; yv[6] = b1.z; yv[7] = b1.w;
  }
  for (int ji = job_lo; ji < job_hi; ++ji) {
    const int4* jp = reinterpret_cast<const int4*>(&sjobs[ji]);
    const int4 j0 = jp[0], j1 = jp[1], j2 = jp[2];
    if (j0.x != BK_JOB_LL) continue;   // (a cancelled partition job keeps its list position)
    const int src_row = j0.z;
    const float vl = __int_as_float(j1.z), vr = __int_as_float(j2.y);
    const unsigned left_id = (unsigned)j1.w;
    const unsigned long long ids = __ldcg(reinterpret_cast<const unsigned long long*>(P.rows + (unsigned long long)((unsigned)c * (unsigned)P.R + (unsigned)src_row) * (unsigned)P.Npad + base));
    long long s_l = 0, s_r = 0;
    bool any = false;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const unsigned id = (unsigned)(ids >> (8 * e)) & 255u;
      if (id == left_id) { s_l += (long long)bk_bern_q(yv[e], noi[e], vl); any = true; }
      else if (id == left_id + 1u) { s_r += (long long)bk_bern_q(yv[e], noi[e], vr); any = true; }
    }
    if (__any_sync(0xffffffffu, any)) {   // |s| <= 8 * 2^29 per lane: (hi, lo) limbs as in round_unit
      const unsigned l_lo = __reduce_add_sync(0xffffffffu, (unsigned)s_l & 0xFFFFu), r_lo = __reduce_add_sync(0xffffffffu, (unsigned)s_r & 0xFFFFu);
      const int l_hi = __reduce_add_sync(0xffffffffu, (int)(s_l >> 16)), r_hi = __reduce_add_sync(0xffffffffu, (int)(s_r >> 16));
      unsigned v = l_lo;
      v = lane == 1 ? (unsigned)l_hi : v;
      v = lane == 2 ? r_lo : v;
      v = lane == 3 ? (unsigned)r_hi : v;
      if (lane < 4) atomicAdd(sacc + ji * BK_LIMBS + BK_LIMB_LLL_LO + lane, v);
    }
  }
}

// Shared-tree multi-output: the rows of the two new leaves contribute bk_lik_q(y, noi[.] + leaf values[.]); the K
// linear predictors without the tree are read per job (L1 hits after the first job of the tile), y stays in registers.
__device__ __forceinline__ void ll_unit_multi(const Params& P, int c, int tile, int job_lo, int job_hi, const Job* __restrict__ sjobs,
                                              unsigned* __restrict__ sacc, const float (*__restrict__ jvals)[2][BK_MAX_OUTPUTS]) {
  const int lane = threadIdx.x & 31;
  const size_t base = (size_t)tile * BK_WARP_TILE + (size_t)lane * BK_ROWS_PER_LANE;
  float yv[8];
  {
    const float4* b = reinterpret_cast<const float4*>(P.y + base);
    const float4 b0 = __ldg(b), b1 = __ldg(b + 1);
    yv[0] = b0.x; yv[1] = b0.y; yv[2] = b0.z; yv[3] = b0.w; yv[4] = b1.x; yv[5] = b1.y; yv[6] = b1.z; yv[7] = b1.w;
  }
  for (int ji = job_lo; ji < job_hi; ++ji) {
    const int4* jp = reinterpret_cast<const int4*>(&sjobs[ji]);
    const int4 j0 = jp[0], j1 = jp[1];
    if (j0.x != BK_JOB_LL) continue;
    const int src_row = j0.z;
    const unsigned left_id = (unsigned)j1.w;
    const unsigned long long ids = __ldcg(reinterpret_cast<const unsigned long long*>(P.rows + ((size_t)c * P.R + src_row) * P.Npad + base));
    long long s_l = 0, s_r = 0;
    bool any = false;
    for (int e = 0; e < 8; ++e) {
      const unsigned id = (unsigned)(ids >> (8 * e)) & 255u;
      const int side = id == left_id ? 0 : (id == left_id + 1u ? 1 : -1);
      if (side < 0) continue;
      float f[BK_MAX_OUTPUTS];
      for (int j = 0; j < P.K; ++j)
        f[j] = BK_FADD(__int_as_float(ld_ca_s32(P.qr + ((size_t)c * P.K + j) * P.Npad + base + e)), jvals[ji][side][j]);
      const long long q = (long long)bk_lik_q(P.lik, P.K, yv[e], f);
      if (side == 0) s_l += q; else s_r += q;
      any = true;
    }
    if (__any_sync(0xffffffffu, any)) {
      const unsigned l_lo = __reduce_add_sync(0xffffffffu, (unsigned)s_l & 0xFFFFu), r_lo = __reduce_add_sync(0xffffffffu, (unsigned)s_r & 0xFFFFu);
      const int l_hi = __reduce_add_sync(0xffffffffu, (int)(s_l >> 16)), r_hi = __reduce_add_sync(0xffffffffu, (int)(s_r >> 16));
      unsigned v = l_lo;
      v = lane == 1 ? (unsigned)l_hi : v;
      v = lane == 2 ? r_lo : v;
      v = lane == 3 ? (unsigned)r_hi : v;
      if (lane < 4) atomicAdd(sacc + ji * BK_LIMBS + BK_LIMB_LLL_LO + lane, v);
    }
  }
}

// ------------------------------------------------------------------ data phase: SWEEP
// Group-wide: BK_GROUP_THREADS threads x 4 rows.  Fuses commit of tree A with prologue of tree B.
template <int MODE>
__device__ __forceinline__ void sweep_unit(const Params& P, int c, int ctile, GroupShared& sh, const int g, const int tid) {
  const ChainCtl* ctl = P.ctl + c;
  const int4 s0 = __ldcg(reinterpret_cast<const int4*>(&ctl->sweep));
  const int4 s1 = __ldcg(reinterpret_cast<const int4*>(&ctl->sweep) + 1);
  const int do_commit = s0.x, commit_tree = s0.y, new_row = s0.z, do_wf = s0.w;
  const int do_pro = s1.x, pro_tree = s1.y, wf_count = s1.z;
  // stage leaf-value tables
  for (int k = tid; k < 256; k += BK_GROUP_THREADS) {
    sh.old_vals[k] = do_commit ? __ldcg(&ctl->old_vals[k]) : 0.0f;
    sh.new_vals[k] = do_commit ? __ldcg(&ctl->new_vals[k]) : 0.0f;
    float pv = 0.0f;
    if (do_pro && k < 255) {   // three independent loads: one L2 round trip (slots beyond the tree hold stale nodes, masked by nn)
      const DNode* nd = P.forest + ((size_t)c * P.m + pro_tree) * BK_MAX_NODES + k;
      const int nn = __ldcg(P.forest_nn + (size_t)c * P.m + pro_tree);
      const int var = __ldcg(&nd->var);
      const float val = __ldcg(&nd->value);
      pv = (k < nn && var < 0) ? val : 0.0f;
    }
    sh.pro_vals[k] = pv;
  }
  for (int k = tid; k < 256; k += BK_GROUP_THREADS) sh.leaf_acc[k] = 0ull;
  if (tid < 8) sh.tot_acc[tid] = 0ull;
  GROUP_SYNC(g);

  const size_t base = (size_t)ctile * BK_COMMIT_TILE + (size_t)tid * 4;
  long long t_sst = 0, t_sr = 0, t_sd = 0, t_limbo = 0;
  unsigned long long t_r2 = 0;
  unsigned pro_pid4 = 0xFFFFFFFFu;
  int pro_valid = 0;
  int pro_q[4] = {0, 0, 0, 0};
  if (base < (size_t)P.Npad) {
    float* stp = P.st + (size_t)c * P.Npad + base;
    float4 st4 = __ldcg(reinterpret_cast<const float4*>(stp));
    float stv[4] = {st4.x, st4.y, st4.z, st4.w};
    if (do_commit) {
      uint8_t* idp = P.ids_tree + ((size_t)c * P.m + commit_tree) * P.Npad + base;
      unsigned oid4 = __ldcg(reinterpret_cast<const unsigned*>(idp));
      unsigned nid4;
      if (new_row == BK_ROW_FOREST) nid4 = oid4;
      else if (new_row == BK_ROW_VIRTUAL) {
        nid4 = 0u;
#pragma unroll
        for (int e = 0; e < 4; ++e) if (base + e >= (size_t)P.N) nid4 |= 0xFFu << (8 * e);
      } else nid4 = __ldcg(reinterpret_cast<const unsigned*>(P.rows + ((size_t)c * P.R + new_row) * P.Npad + base));
      float4 mean4 = make_float4(0.f, 0.f, 0.f, 0.f), m24 = make_float4(0.f, 0.f, 0.f, 0.f);
      float* mp = P.wf_mean + (size_t)c * P.Npad + base;
      float* m2p = P.wf_m2 + (size_t)c * P.Npad + base;
      if (do_wf) { mean4 = __ldcg(reinterpret_cast<const float4*>(mp)); m24 = __ldcg(reinterpret_cast<const float4*>(m2p)); }
      float mean[4] = {mean4.x, mean4.y, mean4.z, mean4.w}, m2[4] = {m24.x, m24.y, m24.z, m24.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        unsigned oid = (oid4 >> (8 * e)) & 255u, nid = (nid4 >> (8 * e)) & 255u;
        float oldp = sh.old_vals[oid], newp = sh.new_vals[nid];  // entry 255 (limbo) is 0
        float noi = BK_FSUB(stv[e], oldp);
        stv[e] = BK_FADD(noi, newp);
        if (do_wf && base + e < (size_t)P.N) {
          float cntf = (float)wf_count;
          float delta = BK_FSUB(newp, mean[e]);
          float mn = BK_FADD(mean[e], BK_FDIV(delta, cntf));
          float delta2 = BK_FSUB(newp, mn);
          float mm2 = BK_FFMA(delta, delta2, m2[e]);
          mean[e] = mn; m2[e] = mm2;
          float sd = BK_FSQRT(BK_FDIV(mm2, cntf));
          t_sd += (long long)bk_quant(sd, P.qscale);
        }
      }
      __stcg(reinterpret_cast<float4*>(stp), make_float4(stv[0], stv[1], stv[2], stv[3]));
      if (s1.w > 0) {   // the step's last commit: keep the draw (bk_run_launch's draws_out), here and on the peer GPUs
        float* dp = P.draws_out + ((size_t)(s1.w - 1) * P.C + c) * P.Npad + base;
        __stcs(reinterpret_cast<float4*>(dp), make_float4(stv[0], stv[1], stv[2], stv[3]));
        for (int pe = 0; pe < P.n_draw_peers; ++pe)   // P2P stores over NVLink: the all-gather of the draws, fused
          __stcs(reinterpret_cast<float4*>(reinterpret_cast<char*>(dp) + P.draw_peer_delta[pe]), make_float4(stv[0], stv[1], stv[2], stv[3]));
      }
      if (new_row != BK_ROW_FOREST) __stcg(reinterpret_cast<unsigned*>(idp), nid4);
      if (do_wf) {
        __stcg(reinterpret_cast<float4*>(mp), make_float4(mean[0], mean[1], mean[2], mean[3]));
        __stcg(reinterpret_cast<float4*>(m2p), make_float4(m2[0], m2[1], m2[2], m2[3]));
      }
    }
    if (do_pro) {
      unsigned pid4 = __ldcg(reinterpret_cast<const unsigned*>(P.ids_tree + ((size_t)c * P.m + pro_tree) * P.Npad + base));
      float4 y4 = __ldg(reinterpret_cast<const float4*>(P.y + (size_t)(c % P.G) * P.Npad + base));
      float yv[4] = {y4.x, y4.y, y4.z, y4.w};
      int qrv[4], qsv[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        unsigned pid = (pid4 >> (8 * e)) & 255u;
        float oldp = sh.pro_vals[pid];
        float noi = BK_FSUB(stv[e], oldp);
        const bool real = base + e < (size_t)P.N;
        int b = real ? bk_quant(stv[e], P.qscale) : 0;
        qsv[e] = b;
        if (P.lik == BK_LIK_NORMAL) {
          float r = BK_FSUB(yv[e], noi);
          int a = real ? bk_quant(r, P.qscale) : 0;
          qrv[e] = a; pro_q[e] = a;
          if (real) {
            unsigned long long sq = (unsigned long long)((long long)a * (long long)a);
            t_sst += b; t_sr += a; t_r2 += sq;
          }
        } else {   // Bernoulli: keep noi itself; terms of the old tree's leaf (per leaf) and of the stump (total)
          qrv[e] = real ? __float_as_int(noi) : 0;
          pro_q[e] = real ? bk_bern_q(yv[e], noi, oldp) : 0;
          if (real) { t_sst += b; t_sr += bk_bern_q(yv[e], noi, P.init_leaf); }
          if (BK_MISSING_ENABLED && real && pid == BK_LIMBO) t_limbo += pro_q[e];   // rows the old tree dropped (missing covariate) predict 0: oldp = 0
        }
      }
      pro_pid4 = pid4; pro_valid = 1;
      __stcg(reinterpret_cast<int4*>(P.qr + (size_t)c * P.Npad + base), make_int4(qrv[0], qrv[1], qrv[2], qrv[3]));
      __stcg(reinterpret_cast<int4*>(P.qst + (size_t)c * P.Npad + base), make_int4(qsv[0], qsv[1], qsv[2], qsv[3]));
    }
  }
  // per-leaf statistics of the old tree: lanes holding the same leaf id are grouped with
  // match.any and summed with redux.sync on 16-bit chunks (exact), one shared-memory atomic
  // per (warp, leaf) instead of one per row
  if (do_pro) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const unsigned pid = (pro_pid4 >> (8 * e)) & 255u;
      const bool ok = pro_valid && pid != BK_LIMBO && base + e < (size_t)P.N;
      const unsigned key = ok ? pid : 0x100u;
      const unsigned grp = __match_any_sync(0xffffffffu, key);
      const int a = ok ? pro_q[e] : 0;
      const unsigned r_lo = __reduce_add_sync(grp, (unsigned)a & 0xFFFFu);
      const int r_hi = __reduce_add_sync(grp, a >> 16);
      if (ok && (int)(tid & 31) == __ffs(grp) - 1) {
        const long long sr = (long long)r_hi * 65536ll + (long long)r_lo;
        atomicAdd(&sh.leaf_acc[pid], (unsigned long long)sr);
      }
    }
  }
  // block totals
  {
    unsigned long long v0 = warp_sum_u64((unsigned long long)t_sr);
    unsigned long long v1 = warp_sum_u64(t_r2 & 0xFFFFFFFFull);
    unsigned long long v2 = warp_sum_u64(t_r2 >> 32);
    unsigned long long v3 = warp_sum_u64((unsigned long long)t_sst);
    unsigned long long v4 = warp_sum_u64((unsigned long long)t_sd);
    unsigned long long v5 = BK_MISSING_ENABLED ? warp_sum_u64((unsigned long long)t_limbo) : 0ull;
    if ((tid & 31) == 0) {
      atomicAdd(&sh.tot_acc[0], v0); atomicAdd(&sh.tot_acc[1], v1); atomicAdd(&sh.tot_acc[2], v2);
      atomicAdd(&sh.tot_acc[3], v3); atomicAdd(&sh.tot_acc[4], v4);
      if (v5) atomicAdd(&sh.tot_acc[5], v5);
    }
  }
  GROUP_SYNC(g);
  unsigned long long* a0 = P.acc0 + (size_t)c * BK_ACC0_WORDS;
  if (do_pro) {
    for (int k = tid; k < 255; k += BK_GROUP_THREADS) {
      unsigned long long v = sh.leaf_acc[k];
      if (v) red_add_u64(a0 + (size_t)k * BK_ACC0_STRIDE, v);
    }
    if (tid < 4) { unsigned long long v = sh.tot_acc[tid]; if (v) red_add_u64(a0 + (size_t)255 * BK_ACC0_STRIDE + tid, v); }
    if (tid == 4) { unsigned long long v = sh.tot_acc[5]; if (v) red_add_u64(a0 + (size_t)256 * BK_ACC0_STRIDE + 1, v); }   // Bernoulli terms of the limbo rows
  }
  if (do_commit && do_wf && tid == 0) { unsigned long long v = sh.tot_acc[4]; if (v) red_add_u64(a0 + (size_t)256 * BK_ACC0_STRIDE, v); }
  GROUP_SYNC(g);
}

// Shared-tree multi-output SWEEP (Params::K > 1): same fusion as sweep_unit — commit of tree A and prologue of tree
// B — looped over the K outputs (the leaf-value tables of one output at a time in shared memory), then ONE pass of
// per-row log-likelihood terms with all K linear predictors: the old tree's leaf of the row (per-leaf sums, particle 0)
// and the root-only stump (total).
__device__ __forceinline__ void sweep_unit_multi(const Params& P, int c, int ctile, GroupShared& sh, const int g, const int tid) {
  const ChainCtl* ctl = P.ctl + c;
  const int4 s0 = __ldcg(reinterpret_cast<const int4*>(&ctl->sweep));
  const int4 s1 = __ldcg(reinterpret_cast<const int4*>(&ctl->sweep) + 1);
  const int do_commit = s0.x, commit_tree = s0.y, new_row = s0.z, do_wf = s0.w;
  const int do_pro = s1.x, pro_tree = s1.y, wf_count = s1.z;
  const int K = P.K;
  for (int idx = tid; idx < K * 256; idx += BK_GROUP_THREADS) {
    const int j = idx >> 8, k = idx & 255;
    float pv = 0.0f;
    if (do_pro && k < 255) {
      const DNode* nd = P.forest + ((size_t)c * P.m + pro_tree) * BK_MAX_NODES + k;
      const int nn = __ldcg(P.forest_nn + (size_t)c * P.m + pro_tree);
      const int var = __ldcg(&nd->var);
      const float val = j == 0 ? __ldcg(&nd->value) : __ldcg(reinterpret_cast<const float*>(nd->aux) + (j - 1));
      pv = (k < nn && var < 0) ? val : 0.0f;
    }
    sh.u.pro_vals_k[j][k] = pv;
  }
  for (int k = tid; k < 256; k += BK_GROUP_THREADS) sh.leaf_acc[k] = 0ull;
  if (tid < 8) { sh.tot_acc[tid] = 0ull; sh.sd_acc[tid] = 0ull; }
  GROUP_SYNC(g);

  const size_t base = (size_t)ctile * BK_COMMIT_TILE + (size_t)tid * 4;
  const bool in = base < (size_t)P.Npad;
  unsigned oid4 = 0u, nid4 = 0u, pid4 = 0xFFFFFFFFu;
  if (in && do_commit) {
    uint8_t* idp = P.ids_tree + ((size_t)c * P.m + commit_tree) * P.Npad + base;
    oid4 = __ldcg(reinterpret_cast<const unsigned*>(idp));
    if (new_row == BK_ROW_FOREST) nid4 = oid4;
    else if (new_row == BK_ROW_VIRTUAL) {
      nid4 = 0u;
#pragma unroll
      for (int e = 0; e < 4; ++e) if (base + e >= (size_t)P.N) nid4 |= 0xFFu << (8 * e);
    } else nid4 = __ldcg(reinterpret_cast<const unsigned*>(P.rows + ((size_t)c * P.R + new_row) * P.Npad + base));
  }
  if (in && do_pro) pid4 = __ldcg(reinterpret_cast<const unsigned*>(P.ids_tree + ((size_t)c * P.m + pro_tree) * P.Npad + base));
  float noi_k[BK_MAX_OUTPUTS][4];
  long long t_sst0 = 0;
  for (int j = 0; j < K; ++j) {
    // this output's tables of the committed tree
    for (int k = tid; k < 256; k += BK_GROUP_THREADS) {
      sh.old_vals[k] = do_commit ? __ldcg(&ctl->old_vals_k[j][k]) : 0.0f;
      sh.new_vals[k] = do_commit ? __ldcg(&ctl->new_vals_k[j][k]) : 0.0f;
    }
    GROUP_SYNC(g);
    long long t_sd = 0;
    if (in) {
      float* stp = P.st + ((size_t)c * K + j) * P.Npad + base;
      const float4 st4 = __ldcg(reinterpret_cast<const float4*>(stp));
      float stv[4] = {st4.x, st4.y, st4.z, st4.w};
      if (do_commit) {
        float4 mean4 = make_float4(0.f, 0.f, 0.f, 0.f), m24 = make_float4(0.f, 0.f, 0.f, 0.f);
        float* mp = P.wf_mean + ((size_t)c * K + j) * P.Npad + base;
        float* m2p = P.wf_m2 + ((size_t)c * K + j) * P.Npad + base;
        if (do_wf) { mean4 = __ldcg(reinterpret_cast<const float4*>(mp)); m24 = __ldcg(reinterpret_cast<const float4*>(m2p)); }
        float mean[4] = {mean4.x, mean4.y, mean4.z, mean4.w}, m2[4] = {m24.x, m24.y, m24.z, m24.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const unsigned oid = (oid4 >> (8 * e)) & 255u, nid = (nid4 >> (8 * e)) & 255u;
          const float oldp = sh.old_vals[oid], newp = sh.new_vals[nid];
          const float noi = BK_FSUB(stv[e], oldp);
          stv[e] = BK_FADD(noi, newp);
          if (do_wf && base + e < (size_t)P.N) {
            const float cntf = (float)wf_count;
            const float delta = BK_FSUB(newp, mean[e]);
            const float mn = BK_FADD(mean[e], BK_FDIV(delta, cntf));
            const float delta2 = BK_FSUB(newp, mn);
            const float mm2 = BK_FFMA(delta, delta2, m2[e]);
            mean[e] = mn; m2[e] = mm2;
            t_sd += (long long)bk_quant(BK_FSQRT(BK_FDIV(mm2, cntf)), P.qscale);
          }
        }
        __stcg(reinterpret_cast<float4*>(stp), make_float4(stv[0], stv[1], stv[2], stv[3]));
        if (s1.w > 0) {
          float* dp = P.draws_out + (((size_t)(s1.w - 1) * P.C + c) * K + j) * P.Npad + base;
          __stcs(reinterpret_cast<float4*>(dp), make_float4(stv[0], stv[1], stv[2], stv[3]));
          for (int pe = 0; pe < P.n_draw_peers; ++pe)
            __stcs(reinterpret_cast<float4*>(reinterpret_cast<char*>(dp) + P.draw_peer_delta[pe]), make_float4(stv[0], stv[1], stv[2], stv[3]));
        }
        if (do_wf) {
          __stcg(reinterpret_cast<float4*>(mp), make_float4(mean[0], mean[1], mean[2], mean[3]));
          __stcg(reinterpret_cast<float4*>(m2p), make_float4(m2[0], m2[1], m2[2], m2[3]));
        }
      }
      if (do_pro) {
        int qrv[4], qsv[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const unsigned pid = (pid4 >> (8 * e)) & 255u;
          const float noi = BK_FSUB(stv[e], sh.u.pro_vals_k[j][pid]);
          const bool real = base + e < (size_t)P.N;
          noi_k[j][e] = noi;
          qrv[e] = real ? __float_as_int(noi) : 0;
          qsv[e] = real ? bk_quant(stv[e], P.qscale) : 0;
          if (j == 0 && real) t_sst0 += qsv[e];
        }
        __stcg(reinterpret_cast<int4*>(P.qr + ((size_t)c * K + j) * P.Npad + base), make_int4(qrv[0], qrv[1], qrv[2], qrv[3]));
        __stcg(reinterpret_cast<int4*>(P.qst + ((size_t)c * K + j) * P.Npad + base), make_int4(qsv[0], qsv[1], qsv[2], qsv[3]));
      }
    }
    if (do_commit && do_wf) {
      const unsigned long long v = warp_sum_u64((unsigned long long)t_sd);
      if ((tid & 31) == 0 && v) atomicAdd(&sh.sd_acc[j], v);
    }
    GROUP_SYNC(g);   // (the tables are overwritten for the next output)
  }
  if (in && do_commit && new_row != BK_ROW_FOREST)
    __stcg(reinterpret_cast<unsigned*>(P.ids_tree + ((size_t)c * P.m + commit_tree) * P.Npad + base), nid4);
  // per-row log-likelihood terms with all K outputs
  long long t_sr = 0;
  int pro_q[4] = {0, 0, 0, 0};
  if (in && do_pro) {
    const float4 y4 = __ldg(reinterpret_cast<const float4*>(P.y + base));
    const float yv[4] = {y4.x, y4.y, y4.z, y4.w};
    for (int e = 0; e < 4; ++e) {
      if (base + e >= (size_t)P.N) continue;
      const unsigned pid = (pid4 >> (8 * e)) & 255u;
      float f_old[BK_MAX_OUTPUTS], f_init[BK_MAX_OUTPUTS];
      for (int j = 0; j < K; ++j) {
        f_old[j] = BK_FADD(noi_k[j][e], sh.u.pro_vals_k[j][pid]);
        f_init[j] = BK_FADD(noi_k[j][e], P.init_leaf);
      }
      pro_q[e] = bk_lik_q(P.lik, K, yv[e], f_old);
      t_sr += (long long)bk_lik_q(P.lik, K, yv[e], f_init);
    }
  }
  if (do_pro) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const unsigned pid = (pid4 >> (8 * e)) & 255u;
      const bool ok = in && pid != BK_LIMBO && base + e < (size_t)P.N;
      const unsigned key = ok ? pid : 0x100u;
      const unsigned grp = __match_any_sync(0xffffffffu, key);
      const int a = ok ? pro_q[e] : 0;
      const unsigned r_lo = __reduce_add_sync(grp, (unsigned)a & 0xFFFFu);
      const int r_hi = __reduce_add_sync(grp, a >> 16);
      if (ok && (int)(tid & 31) == __ffs(grp) - 1) atomicAdd(&sh.leaf_acc[pid], (unsigned long long)((long long)r_hi * 65536ll + (long long)r_lo));
    }
    const unsigned long long v0 = warp_sum_u64((unsigned long long)t_sr), v3 = warp_sum_u64((unsigned long long)t_sst0);
    if ((tid & 31) == 0) { atomicAdd(&sh.tot_acc[0], v0); atomicAdd(&sh.tot_acc[3], v3); }
  }
  GROUP_SYNC(g);
  unsigned long long* a0 = P.acc0 + (size_t)c * BK_ACC0_WORDS;
  if (do_pro) {
    for (int k = tid; k < 255; k += BK_GROUP_THREADS) {
      const unsigned long long v = sh.leaf_acc[k];
      if (v) red_add_u64(a0 + (size_t)k * BK_ACC0_STRIDE, v);
    }
    if (tid == 0) { unsigned long long v = sh.tot_acc[0]; if (v) red_add_u64(a0 + (size_t)255 * BK_ACC0_STRIDE + 0, v); }
    if (tid == 1) { unsigned long long v = sh.tot_acc[3]; if (v) red_add_u64(a0 + (size_t)255 * BK_ACC0_STRIDE + 3, v); }
  }
  if (do_commit && do_wf && tid < K) { const unsigned long long v = sh.sd_acc[tid]; if (v) red_add_u64(P.acc_sd + (size_t)c * 8 + tid, v); }
  GROUP_SYNC(g);
}

// ------------------------------------------------------------------ dataflow scheduling
// Which groups serve chain c: with C >= BK_NGROUPS chains, group g serves the chains c = g (mod BK_NGROUPS); with
// fewer chains, group g serves chain g mod C, so chain c has servers_of(c) groups per worker CTA.
__device__ __forceinline__ int servers_of(int C, int c) { return C >= BK_NGROUPS ? 1 : (BK_NGROUPS - c + C - 1) / C; }

// An epoch of a chain is split STATICALLY: the (tile, job) pairs of a ROUND / LL epoch are flattened tile-major,
// worker CTA w of W takes the contiguous range [w*T/W, (w+1)*T/W) (a few whole tiles: its warps re-read the same
// residual tiles through L1), the chain's groups in that CTA split the range, and each warp takes a contiguous
// chunk.  No claim atomics.  Every serving group reports each epoch exactly once (release add), so the control CTA
// waits for `workers x servers_of(chain)` per epoch.
// Cold start: while the control CTAs run their first phase, every worker thread asks the L2 for a slice of what the
// step is going to touch (the covariates, the responses, each chain's sum of trees, the leaf-id rows of the trees of
// this batch, the running-sd arrays while tuning).  A step then meets DRAM latency once, in parallel, instead of once
// per first touch inside the latency-bound rounds (bench.py flushes the L2 between steps: 0.24 ms of a 0.91 ms C2 step
// were cold misses).  When the working set is still resident the prefetches are L2 hits and cost nothing.
__device__ __forceinline__ void prefetch_l2_range(const void* base, size_t bytes, size_t first, size_t stride) {
  const char* b = reinterpret_cast<const char*>(base);
  for (size_t o = first * 128; o < bytes; o += stride * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(b + o));
}
__device__ __forceinline__ void cold_start_prefetch(const Params& P, const int tune) {
  const size_t W = gridDim.x - P.C, first = (size_t)(blockIdx.x - P.C) * BK_CTA_THREADS + threadIdx.x, stride = W * BK_CTA_THREADS;
  const size_t col = (size_t)P.Npad * 4;
  if ((size_t)P.p * col <= ((size_t)48 << 20)) prefetch_l2_range(P.X, (size_t)P.p * col, first, stride);   // (larger X does not fit the L2 anyway)
  prefetch_l2_range(P.y, (size_t)P.G * col, first, stride);
  for (int c = 0; c < P.C; ++c) {
    prefetch_l2_range(P.st + (size_t)c * P.Npad, col, first, stride);
    if (tune) {
      prefetch_l2_range(P.wf_mean + (size_t)c * P.Npad, col, first, stride);
      prefetch_l2_range(P.wf_m2 + (size_t)c * P.Npad, col, first, stride);
    }
    const int lo = P.ctl[c].hot.lower, T = tune ? P.batch_tune : P.batch_post;   // (the control CTA writes `hot` back only when the step is done)
    const int hi = lo + T < P.m ? lo + T : P.m;
    prefetch_l2_range(P.ids_tree + ((size_t)c * P.m + lo) * P.Npad, (size_t)(hi - lo) * P.Npad, first, stride);
  }
}

template <int MODE>
__device__ __forceinline__ void worker_loop(const Params& P, GroupShared& sh, const int g, const unsigned epoch_base, unsigned char* stage_group) {
  const int tid = (int)threadIdx.x - g * BK_GROUP_THREADS, warp = tid >> 5;
  unsigned char* const stage = stage_group + warp * BK_STAGE_WARP_BYTES;   // this warp's two staging slots (see stage_issue)
  const int W = gridDim.x - P.C, w = blockIdx.x - P.C;
  const int c_first = P.C >= BK_NGROUPS ? g : g % P.C, c_step = P.C >= BK_NGROUPS ? BK_NGROUPS : P.C;
  const int my_rank = P.C >= BK_NGROUPS ? 0 : g / P.C;      // index of this group among the servers of its chain
  int n_mine = 0;
  for (int c = c_first; c < P.C; c += c_step) n_mine++;
  if (P.C < BK_NGROUPS) n_mine = 1;
  if (tid < 64) { sh.fin[tid] = 0; sh.seen[tid] = epoch_base; }
  for (int i = tid; i < BK_MAX_PARTICLES * BK_LIMBS; i += BK_GROUP_THREADS) sh.acc[i] = 0u;
  GROUP_SYNC(g);
  int next = 0, n_finished = 0;
  long long t_idle0 = clock64();
#ifdef BK_PROFILE_CTRL
  unsigned long long t_pub = 0;
#endif
  for (;;) {
    if (warp == 0) {
      // The group's poller is its whole first warp, in lock step: every lane issues the same loads (one transaction)
      // and takes the same decisions, so the warp never splits.  (A single polling lane leaves its warp split after
      // the barrier below, and a split warp pays a WARPSYNC slow path on every REDUX / vote of its units.)
      const int lane = tid;
      Work wk; wk.chain = -1; wk.exit_now = 0; wk.cmd = 0; wk.njobs = 0; wk.total = 0; wk.lo = 0u; wk.hi = 0u;
      unsigned spins = 0;
      while (wk.chain < 0 && !wk.exit_now) {
        for (int k = 0; k < n_mine && wk.chain < 0; ++k) {
          const int idx = (next + k) % n_mine;
          const int c = P.C >= BK_NGROUPS ? c_first + idx * c_step : c_first;
          if (sh.fin[c]) continue;
          ChainSync* sy = P.sync + c;
          // the descriptor {epoch, cmd, jobs, units} is ONE aligned 16-byte word, stored and loaded whole: a single
          // L2 round trip tells the group that an epoch started and what it is
          // The poll IS the acquire (ld.acquire.gpu = LDG.STRONG.GPU + CCTL.IVALL, no MEMBAR): it pairs with the control
          // CTA's release fence before the descriptor store, and it drops stale L1 lines, which is what lets the residual
          // tiles be read through L1.  The other threads of the group are ordered behind it by the group barrier.
          // -DBK_POLL_FENCE: the earlier form, a volatile poll + fence.acq_rel once the epoch is seen (A/B in
          // profiles/r2_ab_acquire_poll.txt: the acquire polls are +4.8 % on C2, +1.4 % on C5).
#ifndef BK_POLL_FENCE
          const uint4 d = ld_acquire_v4(&sy->desc);
#else
          const uint4 d = ld_volatile_v4(&sy->desc);
#endif
          if ((int)(d.x - sh.seen[c]) <= 0) continue;   // (a descriptor left by an earlier launch is older than epoch_base)
#ifdef BK_POLL_FENCE
          if (lane == 0) fence_acq_rel_gpu();
#endif
          __syncwarp();
          if (lane == 0) { sh.seen[c] = d.x; if ((d.y & 0xFFu) == BK_CMD_DONE) sh.fin[c] = 1; }
          __syncwarp();
          if ((d.y & 0xFFu) == BK_CMD_DONE) { n_finished++; continue; }
          wk.chain = c; wk.cmd = (int)(d.y & 0xFFu); wk.njobs = (int)d.z; wk.total = (int)d.w;
          const unsigned ns = (unsigned)servers_of(P.C, c);
          if (wk.cmd == BK_CMD_SWEEP) {
            wk.lo = (unsigned)w * ns + (unsigned)my_rank;             // first row tile of this group; stride W * ns
            wk.hi = (unsigned)W * ns;
          } else {
            // floor(w * T / W) in 32-bit arithmetic: T = q * W + r
            const unsigned Tq = d.w / (unsigned)W, Tr = d.w - Tq * (unsigned)W;
            const unsigned clo = (unsigned)w * Tq + ((unsigned)w * Tr) / (unsigned)W;            // the CTA's pairs ...
            const unsigned chi = (unsigned)(w + 1) * Tq + ((unsigned)(w + 1) * Tr) / (unsigned)W;
            wk.lo = clo + ((chi - clo) * (unsigned)my_rank) / ns;                                    // ... split over its groups
            wk.hi = clo + ((chi - clo) * (unsigned)(my_rank + 1)) / ns;
          }
          next = (idx + 1) % n_mine;
#ifdef BK_PROFILE_CTRL
          t_pub = *reinterpret_cast<volatile unsigned long long*>(&sy->pad[1]);
          if (wk.cmd == BK_CMD_ROUND && g == 0 && lane == 0) { g_wdbg[blockIdx.x][0] += globaltimer_ns() - t_pub; g_wdbg[blockIdx.x][4] += 1; }
#endif
        }
        if (wk.chain < 0) {
          if (n_finished == n_mine) wk.exit_now = 1;
          else if ((++spins & 15u) == 0) {
            int stop = ld_volatile_i32(P.abort_flag) != 0;
            if (!stop && clock64() - t_idle0 > BK_BARRIER_TIMEOUT_CYCLES) { atomicExch(P.abort_flag, 1); stop = 1; }
            if (__any_sync(0xffffffffu, stop)) wk.exit_now = 1;
          }
        }
      }
      if (lane == 0) sh.work = wk;
    }
    GROUP_SYNC(g);
    const Work wk = sh.work;
    if (wk.exit_now) return;
    if (wk.cmd == BK_CMD_ROUND || wk.cmd == BK_CMD_LL) {
      {
        const uint4* g4 = reinterpret_cast<const uint4*>(P.ctl[wk.chain].jobs[(w * BK_NGROUPS + g) % BK_JOB_COPIES]);   // readers spread over the copies
        uint4* d4 = reinterpret_cast<uint4*>(sh.jobs);
        for (int i = tid; i < wk.njobs * 3; i += BK_GROUP_THREADS) d4[i] = __ldcg(g4 + i);
        if (BK_IS_MULTI(P)) {   // shared-tree multi-output: per-output accumulators (ROUND) / the jobs' leaf values (LL)
          if (wk.cmd == BK_CMD_ROUND) {
            unsigned* z = &sh.u.acck[0][0][0];
            for (int i = tid; i < wk.njobs * BK_MAX_OUTPUTS * 4; i += BK_GROUP_THREADS) z[i] = 0u;
          } else {
            const float* gv = &P.ctl[wk.chain].job_vals[0][0][0];
            float* dv = &sh.u.job_vals[0][0][0];
            for (int i = tid; i < wk.njobs * 2 * BK_MAX_OUTPUTS; i += BK_GROUP_THREADS) dv[i] = __ldcg(gv + i);
          }
        }
      }
      GROUP_SYNC(g);
      const Job* jobs = sh.jobs;
      const unsigned nwarp = BK_GROUP_THREADS >> 5;
      const unsigned chunk = (wk.hi - wk.lo + nwarp - 1u) / nwarp;   // pairs per warp
      unsigned lo = wk.lo + (unsigned)warp * chunk;
      const unsigned hi = lo + chunk < wk.hi ? lo + chunk : wk.hi;
      while (lo < hi) {     // one call per tile segment of this warp's range
        const unsigned tile = lo / (unsigned)wk.njobs, j0 = lo - tile * (unsigned)wk.njobs;
        const unsigned seg = (unsigned)wk.njobs - j0 < hi - lo ? (unsigned)wk.njobs - j0 : hi - lo;
        if (wk.cmd == BK_CMD_ROUND) {
          if (BK_IS_MULTI(P)) round_unit<false, true>(P, wk.chain, (int)tile, (int)j0, (int)(j0 + seg), jobs, sh.acc, sh.u.acck, stage);
          else if (BK_MISSING_ENABLED) round_unit<true, false>(P, wk.chain, (int)tile, (int)j0, (int)(j0 + seg), jobs, sh.acc, nullptr, stage);
          else round_unit<false, false>(P, wk.chain, (int)tile, (int)j0, (int)(j0 + seg), jobs, sh.acc, nullptr, stage);
        }
        else if (BK_IS_MULTI(P)) ll_unit_multi(P, wk.chain, (int)tile, (int)j0, (int)(j0 + seg), jobs, sh.acc, sh.u.job_vals);
        else ll_unit(P, wk.chain, (int)tile, (int)j0, (int)(j0 + seg), jobs, sh.acc);
        lo += seg;
      }
      GROUP_SYNC(g);
      // flush the group's partial sums: one global atomic per non-zero (job, statistic); leaves sh.acc zeroed
      for (int i = tid; i < wk.njobs * BK_ACC_STRIDE; i += BK_GROUP_THREADS) {
        const int ji = i / BK_ACC_STRIDE, k = i % BK_ACC_STRIDE;
        const unsigned long long v = take_stat(sh.acc + ji * BK_LIMBS, k);
        if (v) red_add_u64(P.accL + ((size_t)wk.chain * P.P + jobs[ji].slot) * BK_ACC_STRIDE + k, v);
      }
      if (BK_IS_MULTI(P) && wk.cmd == BK_CMD_ROUND) {   // per-output sums of both children: (lo, hi) limbs -> one global atomic each
        for (int i = tid; i < wk.njobs * P.K * 2; i += BK_GROUP_THREADS) {
          const int ji = i / (P.K * 2), j = (i / 2) % P.K, side = i & 1;
          const unsigned lo = sh.u.acck[ji][j][2 * side], hi = sh.u.acck[ji][j][2 * side + 1];
          const long long v = (long long)(int)hi * 65536ll + (long long)lo;
          if (v) red_add_u64(P.accK + (((size_t)wk.chain * P.P + jobs[ji].slot) * P.K + j) * 2 + side, (unsigned long long)v);
        }
      }
    } else {  // BK_CMD_SWEEP: group-wide row tiles, round robin over all serving groups
      for (unsigned u = wk.lo; u < (unsigned)wk.total; u += wk.hi) {
        if (BK_IS_MULTI(P)) sweep_unit_multi(P, wk.chain, (int)u, sh, g, tid);
        else sweep_unit<MODE>(P, wk.chain, (int)u, sh, g, tid);
      }
    }
    GROUP_SYNC(g);   // every warp's stores are ordered before the release below
    if (wk.cmd == BK_CMD_ROUND && g == 0) WDBG(2, t_pub);
    if (tid == 0) red_release_add_u32(&P.sync[wk.chain].done, 1u);
    t_idle0 = clock64();
    if (wk.cmd == BK_CMD_ROUND && g == 0) WDBG(3, t_pub);
  }
}

// ---- control CTA of chain c: wait for the previous epoch, run the state machine, publish the next
template <int MODE>
__device__ __forceinline__ bool control_loop(const Params& P, int c, int tune, const StepArgs& A, int max_phases, CtlShared& sh) {
  __shared__ int s_abort, s_fin;
  __shared__ ChainHot s_hot;
  ChainCtl* ctl = P.ctl + c;
  ChainHot* hot = &s_hot;
  ChainSync* sy = P.sync + c;
  if (threadIdx.x == 0) { s_hot = ctl->hot; sh.pre_prop_tree = -1; sh.pre_z_tree = -1; sh.copy_pending = 0; }   // persistent scalars -> shared memory for the whole step
  // read-only tables of the round path: the control CTA's L1 is dropped by every acquire fence, so a global table
  // costs an L2 round trip per phase
  if (threadIdx.x < 64) sh.p_leaf[threadIdx.x] = P.p_leaf[threadIdx.x];
  for (int i = threadIdx.x; i < P.P * BK_ACC_STRIDE; i += BK_CTRL_THREADS)
    sh.acc_prev[i / BK_ACC_STRIDE][i % BK_ACC_STRIDE] = __ldcg(P.accL + (size_t)c * P.P * BK_ACC_STRIDE + i);
  for (int v = threadIdx.x; v < P.p && v < BK_CUM_SMEM; v += BK_CTRL_THREADS) sh.rules[v] = (signed char)P.rules[v];
#ifdef BK_PROFILE_CTRL
  if (threadIdx.x < 32) s_cdbg[threadIdx.x] = 0ull;
#endif
  CTRL_SYNC();
  // epoch ids and the done counter run on across launches (no per-launch memset of the sync lines): this launch's epochs are
  // A.epoch_base + 1, + 2, ...; the workers' share of an epoch is counted from the value `done` had when the launch began
  unsigned issued = ld_relaxed_u32(&sy->done), epoch = A.epoch_base;
  const int n_workers = (gridDim.x - P.C) * servers_of(P.C, c);   // serving groups: each reports every epoch once
  const int sweep_tiles = (P.Npad + BK_COMMIT_TILE - 1) / BK_COMMIT_TILE;
  unsigned long long t_wait = 0, t_ctrl = 0, t_pub = 0, t_begin = 0, t_wait_sweep = 0;
  int last_cmd = BK_CMD_IDLE;
  if (threadIdx.x == 0) t_begin = globaltimer_ns();
  for (int phase = 0; phase < max_phases; ++phase) {
    unsigned long long q0 = 0, q1 = 0, q2 = 0;
    if (threadIdx.x == 0) {
      q0 = globaltimer_ns();
      int ab = 0;
      long long t0 = clock64();
      unsigned spins = 0;
      // acquire poll of the done counter (pairs with the serving groups' red.release); -DBK_POLL_FENCE: relaxed poll + fence
#ifndef BK_POLL_FENCE
      while ((int)(ld_acquire_u32(&sy->done) - issued) < 0) {
#else
      while ((int)(ld_relaxed_u32(&sy->done) - issued) < 0) {
#endif
        if ((++spins & 63u) == 0) {
          if (ld_volatile_i32(P.abort_flag)) { ab = 1; break; }
          if (clock64() - t0 > BK_BARRIER_TIMEOUT_CYCLES) { atomicExch(P.abort_flag, 1); ab = 1; break; }
        }
      }
#ifdef BK_POLL_FENCE
      fence_acq_rel_gpu();
#endif
      if (!ab && ld_volatile_i32(P.abort_flag)) ab = 1;
      s_abort = ab;
      q1 = globaltimer_ns();
    }
    CTRL_SYNC();
    if (s_abort) return false;
    CTS(16, 0);
    control_step<MODE>(P, c, phase, tune, A, hot, sh);
    CTRL_SYNC();
    CTS(23, 0);
    if (threadIdx.x == 0) {
      q2 = globaltimer_ns();
      hot->stage = hot->stage_next;   // (every thread read the old stage before the barrier above)
      const int cmd = hot->cmd;
      int fin = 0;
      if (cmd == BK_CMD_DONE) {
        fin = 1;
        epoch += 1;
        st_volatile_v4(&sy->desc, make_uint4(epoch, (unsigned)BK_CMD_DONE, 0u, 0u));   // tells the workers this chain is finished
      } else {
        const int nj = hot->n_jobs;
        const int total = (cmd == BK_CMD_ROUND || cmd == BK_CMD_LL) ? nj * P.ntiles : sweep_tiles;
        issued += (unsigned)n_workers; epoch += 1;
#ifdef BK_PROFILE_CTRL
        *reinterpret_cast<volatile unsigned long long*>(&sy->pad[1]) = globaltimer_ns();
#endif
        // jobs, accumulators, rows bookkeeping written by this CTA happen-before the descriptor: release fence
        // (cumulative over the CTA barrier above), then one 16-byte store
        fence_acq_rel_gpu();
        st_volatile_v4(&sy->desc, make_uint4(epoch, (unsigned)cmd, (unsigned)nj, (unsigned)total));
      }
      s_fin = fin;
      const unsigned long long q3 = globaltimer_ns();
      t_wait += q1 - q0; t_ctrl += q2 - q1; t_pub += q3 - q2;
      if (last_cmd == BK_CMD_SWEEP) t_wait_sweep += q1 - q0;
      last_cmd = cmd;
      if (fin) {
        ctl->hot = s_hot;                   // write the scalar state back
#ifdef BK_PROFILE_CTRL
        if (c == 0) for (int i = 0; i < 32; ++i) g_cdbg[i] += s_cdbg[i];
#endif
        bk_step_stats* st = step_stats(P, A.n_steps - 1, c);   // (launch-wide timers go with the last step's record)
        st->us_control = (int32_t)(t_ctrl / 1000ull); st->us_data = (int32_t)(t_wait / 1000ull);
        st->us_sync = (int32_t)(t_pub / 1000ull); st->us_total = (int32_t)((q3 - t_begin) / 1000ull);
        st->reserved[0] = (int32_t)(t_wait_sweep / 1000ull);   // part of us_data spent waiting for SWEEP epochs
        __threadfence();
      }
    }
    CTRL_SYNC();
    CTS(24, 0);
    if (s_fin) return true;
    if (hot->cmd == BK_CMD_ROUND) shadow_round(P, c, hot, sh);   // overlaps the epoch just published
    CTS(25, 0);
  }
  if (threadIdx.x == 0) { atomicExch(P.abort_flag, 1); }
  return false;
}

// ------------------------------------------------------------------ the step kernel
__global__ void __launch_bounds__(BK_CTA_THREADS, 1)
pgbart_step_kernel(const Params P, const int tune, const StepArgs A, const int max_phases) {
  const int mode = P.K > 1 ? BK_MODE_MULTI : ((P.has_nan || P.n_subset > 0) ? BK_MODE_MISSING : BK_MODE_PLAIN);   // (uniform over the grid)
  __shared__ KernelShared sh;   // control CTAs: the chain's control state; worker CTAs: staging slots of groups 0-1
  if ((int)blockIdx.x < P.C) {
    if (threadIdx.x < BK_CTRL_THREADS) {
      if (mode == BK_MODE_PLAIN) control_loop<BK_MODE_PLAIN>(P, blockIdx.x, tune, A, max_phases, sh.ctl);
      else if (mode == BK_MODE_MISSING) control_loop<BK_MODE_MISSING>(P, blockIdx.x, tune, A, max_phases, sh.ctl);
      else control_loop<BK_MODE_MULTI>(P, blockIdx.x, tune, A, max_phases, sh.ctl);
    }
    return;
  }
  const int g = threadIdx.x / BK_GROUP_THREADS;
#ifndef BK_NO_COLD_PREFETCH
  cold_start_prefetch(P, tune);
#endif
  GroupShared& gs = reinterpret_cast<GroupShared*>(bk_dyn_smem)[g];
  static_assert(BK_NGROUPS == 4, "staging regions are laid out for four worker groups");
  static_assert(sizeof(KernelShared) >= 2 * BK_STAGE_GROUP_BYTES, "groups 0-1 stage in the static control area");
  static_assert(sizeof(GroupShared) % 16 == 0, "the dynamic staging tail must stay 16-byte aligned");
  static_assert(sizeof(GroupShared) * BK_NGROUPS + 2 * BK_STAGE_GROUP_BYTES + sizeof(KernelShared) + 2048 <= 227 * 1024,
                "worker groups + staging tail + static area must fit the 227 KB of shared memory of an SM");
  unsigned char* const stage_group = g < 2 ? reinterpret_cast<unsigned char*>(&sh) + g * BK_STAGE_GROUP_BYTES
                                           : bk_dyn_smem + sizeof(GroupShared) * BK_NGROUPS + (g - 2) * BK_STAGE_GROUP_BYTES;
  if (mode == BK_MODE_PLAIN) worker_loop<BK_MODE_PLAIN>(P, gs, g, A.epoch_base, stage_group);
  else if (mode == BK_MODE_MISSING) worker_loop<BK_MODE_MISSING>(P, gs, g, A.epoch_base, stage_group);
  else worker_loop<BK_MODE_MULTI>(P, gs, g, A.epoch_base, stage_group);
}

// ------------------------------------------------------------------ init kernel
__global__ void pgbart_init_kernel(const Params P, const float init_sum, const float leaf_sd_init,
                                   const double* __restrict__ split_prior) {
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t nth = (size_t)gridDim.x * blockDim.x;
  for (size_t i = tid; i < (size_t)P.C * P.K * P.Npad; i += nth) {
    size_t r = i % P.Npad;
    P.st[i] = r < (size_t)P.N ? init_sum : 0.0f;
    P.wf_mean[i] = 0.0f; P.wf_m2[i] = 0.0f; P.qr[i] = 0; P.qst[i] = 0;
  }
  for (size_t i = tid; i < (size_t)P.C * P.P * P.K * 2; i += nth) P.accK[i] = 0ull;
  for (size_t i = tid; i < (size_t)P.C * 8; i += nth) P.acc_sd[i] = 0ull;
  for (size_t i = tid; i < (size_t)P.C * P.m * P.Npad; i += nth) {
    size_t r = i % P.Npad;
    P.ids_tree[i] = r < (size_t)P.N ? 0 : BK_LIMBO;
  }
  for (size_t i = tid; i < (size_t)P.C * P.R * P.cnt_stride; i += nth) P.rowcnt[i] = 0u;
  for (size_t i = tid; i < (size_t)P.C * P.R * P.nb_stride; i += nth) P.coarse[i] = 0u;
  for (size_t i = tid; i < (size_t)P.C * P.R * BK_MAX_SUBSET_COLS; i += nth) P.present[i] = 0u;
  for (size_t i = tid; i < (size_t)P.C * P.P * BK_ACC_STRIDE; i += nth) P.accL[i] = 0ull;
  for (size_t i = tid; i < (size_t)P.C * BK_ACC0_WORDS; i += nth) P.acc0[i] = 0ull;
  for (size_t i = tid; i < (size_t)P.C * P.m; i += nth) {
    P.forest_nn[i] = 1;
    DNode nd; memset(&nd, 0, sizeof(nd));
    nd.var = -1; nd.left = -1; nd.value = P.init_leaf; nd.n = P.N;
    for (int j = 1; j < P.K; ++j) set_node_val(nd, j, P.init_leaf);
    P.forest[i * BK_MAX_NODES] = nd;
  }
  for (size_t i = tid; i < (size_t)P.C * P.p; i += nth) P.alpha_vec[i] = split_prior[i % P.p];
  for (size_t i = tid; i < (size_t)BK_MAX_STEPS_PER_LAUNCH * P.rec_stride / 4; i += nth) reinterpret_cast<int32_t*>(P.stats)[i] = 0;
  for (size_t c = tid; c < (size_t)P.C; c += nth) {
    ChainCtl* ctl = P.ctl + c;
    ctl->hot.tune = 1; ctl->hot.sigma = 1.0f; ctl->hot.iter = 0; ctl->hot.lower = 0; ctl->hot.draw = 0; ctl->hot.wf_count = 0;
    ctl->hot.leaf_sd = leaf_sd_init; ctl->hot.stage = BK_ST_DONE;
    for (int j = 0; j < BK_MAX_OUTPUTS; ++j) ctl->hot.leaf_sdk[j] = leaf_sd_init; ctl->hot.cmd = BK_CMD_DONE; ctl->hot.n_jobs = 0;
    ctl->hot.c_err = 0; ctl->hot.trace_round_base = 0;
  }
  if (tid == 0) *P.abort_flag = 0;
  for (size_t i = tid; i < (size_t)P.C * (sizeof(ChainSync) / 4); i += nth) reinterpret_cast<unsigned int*>(P.sync)[i] = 0u;
}
// One block per column: does it hold missing values (bit 0 of col_nan)?  SubsetSplit columns: the categories present in the
// column (col_cats) and whether every value is a category code (bit 1 of col_nan = some value is not).
__global__ void pgbart_nan_scan_kernel(const Params P) {
  const float* x = P.X + (size_t)blockIdx.x * P.Npad;
  const bool subset = P.rules[blockIdx.x] == BK_RULE_SUBSET;
  __shared__ unsigned s_cats;
  __shared__ int s_bad;
  if (threadIdx.x == 0) { s_cats = 0u; s_bad = 0; }
  __syncthreads();
  int any = 0, bad = 0;
  unsigned cats = 0u;
  for (int i = threadIdx.x; i < P.N; i += blockDim.x) {
    const float xv = x[i];
    if (xv != xv) any = 1;
    else if (subset) { const int cd = bk_subset_code(xv); if (cd < 0) bad = 1; else cats |= 1u << cd; }
  }
  if (cats) atomicOr(&s_cats, cats);
  if (bad) s_bad = 1;
  any = __syncthreads_or(any);
  if (threadIdx.x == 0) { P.col_nan[blockIdx.x] = (any ? 1 : 0) | (s_bad ? 2 : 0); P.col_cats[blockIdx.x] = s_cats; }
}
__global__ void pgbart_init_cum_kernel(const Params P) {
  int c = blockIdx.x;
  if (threadIdx.x == 0 && c < P.C) rebuild_cum_dev(P, c);
}

// ==================================================================== host side / C ABI
static thread_local char g_err[512] = "";
static void set_err(const char* fmt, const char* a = "", const char* b = "") { snprintf(g_err, sizeof(g_err), fmt, a, b); }
#define CK(call)                                                                         \
  do {                                                                                   \
    cudaError_t e_ = (call);                                                             \
    if (e_ != cudaSuccess) { set_err("CUDA error %s at %s", cudaGetErrorString(e_), #call); return BK_ERR_CUDA; } \
  } while (0)

// Every entry point runs on the handle's device and puts the caller's current device back (the host is a torch
// process whose current device must not change under its feet).
struct DeviceGuard {
  int prev = -1;
  cudaError_t err;
  explicit DeviceGuard(int dev) {
    err = cudaGetDevice(&prev);
    if (err == cudaSuccess && prev != dev) err = cudaSetDevice(dev); else if (err == cudaSuccess) prev = -1;
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
#define ON_DEVICE(dev)                                                                                              \
  DeviceGuard guard_(dev);                                                                                          \
  if (guard_.err != cudaSuccess) { set_err("CUDA error %s selecting the device", cudaGetErrorString(guard_.err)); return BK_ERR_CUDA; }

struct bk_handle_s {
  bk_settings s;
  Params P;
  cudaStream_t stream;
  int grid;
  int max_phases;
  size_t dyn_smem;
  double* split_prior_dev;
  char* workspace;
  int32_t* vi_pinned;
  bk_step_stats* stats_pinned;
  // Two steps may be in flight (bk_step_launch twice, then bk_step_wait): every per-step host buffer exists twice and
  // slot k & 1 belongs to launch k; the *_pinned pointers above / below follow the last step waited for
  unsigned char* out_slot[2];  // pinned block [vi | stats | abort flag] of a launch: ONE D2H copy per step
  float* st_slot[2];           // pinned [C][n_rows] sum of trees (bk_set_host_output)
  DNode* hist_nodes_slot[2];   // pinned [steps][C][Tmax][255] raw nodes of the trees the post-tuning steps rewrote (bk_set_history)
  int32_t* hist_nn_slot[2];    // pinned [steps][C][Tmax]
  int hist_cap[2];             // steps the slot's buffers hold
  int hist_first[2][BK_MAX_STEPS_PER_LAUNCH], hist_count[2][BK_MAX_STEPS_PER_LAUNCH];
  cudaEvent_t done_ev[2];
  long long n_launched, n_waited;
  int last_slot;
  int history;                 // capture the rewritten trees of post-tuning steps
  int host_lower;              // host mirror of the chains' `lower` (first tree of the next batch)
  size_t out_off_dev, out_bytes, out_stats_off, out_abort_off, rec_stride;
  int steps_in_slot[2];        // steps of the launch that owns the slot
  StepArgs args;
  int32_t* abort_pinned;
  int host_output;
  int poisoned;           // a step timed out: trees / sum of trees may be half rewritten, the handle refuses further steps
  int32_t* marker_host;
  int marker_count;
  float* y_stage;         // pinned [G][n_rows] staging of bk_set_response
};

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct Layout {
  size_t qr, qst, ids_tree, rows, rowcnt, coarse, wf_mean, wf_m2, parts, forest, forest_nn, ctl, accL, acc0, accK, acc_sd, alpha_vec, cum,
      p_leaf, rules, col_nan, subset_cols, subset_idx, col_cats, present, stats, trace, sync, abort_flag, split_prior, total;
  int Npad, ntiles, R, nb;
  size_t rec_stride;
};

static int make_layout(const bk_settings* s, Layout* L) {
  if (!s || s->abi_version != BK_ABI_VERSION) { set_err("bk_settings.abi_version mismatch"); return BK_ERR_ARG; }
  if (s->n_rows < 1 || s->n_cols < 1 || s->n_trees < 1 || s->n_chains < 1) { set_err("empty problem"); return BK_ERR_ARG; }
  if (s->n_particles < 2 || s->n_particles > BK_MAX_PARTICLES) { set_err("n_particles must be in [2,128]"); return BK_ERR_ARG; }
  if ((long long)s->n_chains * (s->n_groups > 1 ? s->n_groups : 1) > 64) { set_err("at most 64 chains x groups per handle"); return BK_ERR_ARG; }
  if (s->n_groups > 0xFFFF) { set_err("n_groups must fit the 16-bit Philox counter field"); return BK_ERR_ARG; }
  if (s->n_trees > 65535) { set_err("n_trees must fit the 16-bit Philox counter field"); return BK_ERR_ARG; }
  if (s->n_rows > (1 << 25)) { set_err("n_rows above 2^25 would overflow the 32-bit partial-sum limbs of a worker group"); return BK_ERR_ARG; }
  if (s->n_outputs > 1) {
    if (s->n_outputs > BK_MAX_OUTPUTS || s->n_groups > 1) { set_err("at most 7 shared-tree outputs, and not together with separate trees"); return BK_ERR_UNSUPPORTED; }
    if (s->likelihood != BK_LIK_NORMAL_HETERO && s->likelihood != BK_LIK_CATEGORICAL) { set_err("shared-tree multi-output needs the heteroscedastic Normal or the Categorical likelihood"); return BK_ERR_UNSUPPORTED; }
    if (s->likelihood == BK_LIK_NORMAL_HETERO && s->n_outputs != 2) { set_err("the heteroscedastic Normal likelihood takes two outputs"); return BK_ERR_ARG; }
  } else if (s->likelihood != BK_LIK_NORMAL && s->likelihood != BK_LIK_BERNOULLI_LOGIT) { set_err("likelihood not implemented on the device (no CPU fallback)"); return BK_ERR_UNSUPPORTED; }
  if (!s->p_leaf || !s->split_prior) { set_err("p_leaf and split_prior are required"); return BK_ERR_ARG; }
  if (s->split_rules) {
    int n_sub = 0;
    for (int v = 0; v < s->n_cols; ++v) {
      if (s->split_rules[v] < BK_RULE_CONTINUOUS || s->split_rules[v] > BK_RULE_SUBSET) { set_err("unknown split rule code"); return BK_ERR_ARG; }
      n_sub += s->split_rules[v] == BK_RULE_SUBSET ? 1 : 0;
    }
    if (n_sub > BK_MAX_SUBSET_COLS) { set_err("at most 8 columns may use the SubsetSplit rule"); return BK_ERR_UNSUPPORTED; }
    if (n_sub > 0 && s->n_outputs > 1) { set_err("SubsetSplit is not available together with shared-tree multi-output"); return BK_ERR_UNSUPPORTED; }
  }
  const size_t C = (size_t)s->n_chains * (s->n_groups > 1 ? s->n_groups : 1), P = s->n_particles, m = s->n_trees, p = s->n_cols;
  const size_t K = s->n_outputs > 1 ? (size_t)s->n_outputs : 1;
  L->Npad = (int)align_up((size_t)s->n_rows, BK_WARP_TILE);
  L->ntiles = L->Npad / BK_WARP_TILE;
  L->R = 2 * s->n_particles;
  const size_t Npad = L->Npad, R = L->R;
  size_t o = 0;
#define CARVE(name, bytes) L->name = o; o = align_up(o + (bytes), 256)
  CARVE(qr, C * K * Npad * 4);
  CARVE(qst, C * K * Npad * 4);
  CARVE(ids_tree, C * m * Npad);
  CARVE(rows, C * R * Npad);
  CARVE(rowcnt, C * R * (size_t)((L->ntiles + 3) & ~3) * 4);   // rows padded to 16 bytes for 128-bit loads
  L->nb = L->ntiles > BK_COARSE_MIN_TILES ? (L->ntiles + BK_COARSE_TILES - 1) / BK_COARSE_TILES : 0;   // bucket counts only where the tile counts no longer fit one load per lane
  CARVE(coarse, C * R * (size_t)((L->nb + 3) & ~3) * 4);
  CARVE(wf_mean, C * K * Npad * 4);
  CARVE(wf_m2, C * K * Npad * 4);
  CARVE(parts, C * 2 * P * sizeof(DParticle));
  CARVE(forest, C * m * BK_MAX_NODES * sizeof(DNode));
  CARVE(forest_nn, C * m * 4);
  CARVE(ctl, C * sizeof(ChainCtl));
  CARVE(accL, C * P * BK_ACC_STRIDE * 8);
  CARVE(acc0, C * BK_ACC0_WORDS * 8);
  CARVE(accK, C * P * K * 2 * 8);
  CARVE(acc_sd, C * 8 * 8);
  CARVE(alpha_vec, C * p * 8);
  CARVE(cum, C * p * 8);
  CARVE(p_leaf, 256 * 8);
  CARVE(rules, p * 4);
  CARVE(col_nan, p * 4);
  CARVE(subset_cols, BK_MAX_SUBSET_COLS * 4);
  CARVE(subset_idx, p * 4);
  CARVE(col_cats, p * 4);
  CARVE(present, C * R * BK_MAX_SUBSET_COLS * 4);
  // abort flag | step records [stats [C] | vi [C][p]] x BK_MAX_STEPS_PER_LAUNCH: ONE D2H copy per launch brings the flag and the
  // records of the steps that ran
  CARVE(abort_flag, 256);
  L->rec_stride = align_up(C * sizeof(bk_step_stats) + C * p * 4, 16);
  CARVE(stats, (size_t)BK_MAX_STEPS_PER_LAUNCH * L->rec_stride);
  CARVE(trace, C * (size_t)(s->trace_capacity > 0 ? s->trace_capacity : 0) * sizeof(bk_trace_rec));
  CARVE(sync, C * sizeof(ChainSync));
  CARVE(split_prior, p * 8);
#undef CARVE
  L->total = o;
  return BK_OK;
}

static int create_body(bk_handle* h, const bk_settings* s, const Layout& L, const float* X_dev, const float* y_dev,
                       float* sum_trees_dev, void* workspace_dev);

extern "C" {

int bk_abi_version(void) { return BK_ABI_VERSION; }
void bk_set_error_message(const char* msg) { set_err("%s", msg); }   // (used by the prediction translation unit)
int bk_padded_rows(int n_rows) { return n_rows < 1 ? 0 : (int)align_up((size_t)n_rows, BK_WARP_TILE); }
const char* bk_last_error(void) { return g_err; }

int bk_query_bytes(const bk_settings* s, size_t* workspace_bytes) {
  Layout L;
  int rc = make_layout(s, &L);
  if (rc != BK_OK) return rc;
  if (!workspace_bytes) { set_err("workspace_bytes is NULL"); return BK_ERR_ARG; }
  *workspace_bytes = L.total;
  return BK_OK;
}

int bk_create(const bk_settings* s, const float* X_dev, const float* y_dev, float* sum_trees_dev, void* workspace_dev,
              bk_handle** out) {
  Layout L;
  int rc = make_layout(s, &L);
  if (rc != BK_OK) return rc;
  if (!X_dev || !y_dev || !sum_trees_dev || !workspace_dev || !out) { set_err("NULL device pointer"); return BK_ERR_ARG; }
  if (((uintptr_t)X_dev | (uintptr_t)y_dev | (uintptr_t)sum_trees_dev | (uintptr_t)workspace_dev) & 15) {
    set_err("device pointers must be 16-byte aligned"); return BK_ERR_ARG;
  }
  ON_DEVICE(s->device);
  bk_handle* h = new (std::nothrow) bk_handle();
  if (!h) { set_err("out of host memory"); return BK_ERR_ARG; }
  memset(h, 0, sizeof(*h));
  h->s = *s;
  rc = create_body(h, s, L, X_dev, y_dev, sum_trees_dev, workspace_dev);
  if (rc != BK_OK) { bk_destroy(h); return rc; }   // stream, pinned buffers and the handle go with it
  *out = h;
  return BK_OK;
}

}  // extern "C"

static int create_body(bk_handle* h, const bk_settings* s, const Layout& L, const float* X_dev, const float* y_dev,
                       float* sum_trees_dev, void* workspace_dev) {
  char* w = (char*)workspace_dev;
  Params& P = h->P;
  P.N = s->n_rows; P.Npad = L.Npad; P.p = s->n_cols; P.m = s->n_trees; P.P = s->n_particles;
  P.C = s->n_chains * (s->n_groups > 1 ? s->n_groups : 1);
  P.G = s->n_groups > 1 ? s->n_groups : 1;
  P.K = s->n_outputs > 1 ? s->n_outputs : 1;
  P.R = L.R; P.ntiles = L.ntiles; P.cnt_stride = (L.ntiles + 3) & ~3; P.lik = s->likelihood; P.trace_cap = s->trace_capacity > 0 ? s->trace_capacity : 0;
  P.batch_tune = s->batch_tune < 1 ? 1 : s->batch_tune; P.batch_post = s->batch_post < 1 ? 1 : s->batch_post;
  P.qscale = ldexpf(1.0f, s->qshift); P.inv_qscale = ldexp(1.0, -s->qshift); P.init_leaf = s->init_leaf;
  P.inv_qm = P.inv_qscale / (double)s->n_trees;
  P.seed = s->seed; P.chain_base = s->chain_base;
  P.X = X_dev; P.y = y_dev; P.st = sum_trees_dev;
  P.qr = (int32_t*)(w + L.qr); P.qst = (int32_t*)(w + L.qst); P.ids_tree = (uint8_t*)(w + L.ids_tree);
  P.rows = (uint8_t*)(w + L.rows); P.rowcnt = (uint32_t*)(w + L.rowcnt);
  P.coarse = (uint32_t*)(w + L.coarse); P.nb = L.nb; P.nb_stride = (L.nb + 3) & ~3;
  P.wf_mean = (float*)(w + L.wf_mean); P.wf_m2 = (float*)(w + L.wf_m2);
  P.parts = (DParticle*)(w + L.parts); P.forest = (DNode*)(w + L.forest); P.forest_nn = (int32_t*)(w + L.forest_nn);
  P.ctl = (ChainCtl*)(w + L.ctl); P.accL = (unsigned long long*)(w + L.accL); P.acc0 = (unsigned long long*)(w + L.acc0);
  P.accK = (unsigned long long*)(w + L.accK); P.acc_sd = (unsigned long long*)(w + L.acc_sd);
  P.alpha_vec = (double*)(w + L.alpha_vec); P.cum = (double*)(w + L.cum); P.p_leaf = (double*)(w + L.p_leaf);
  P.rules = (int32_t*)(w + L.rules); P.col_nan = (int32_t*)(w + L.col_nan); P.stats = (bk_step_stats*)(w + L.stats);
  P.subset_cols = (int32_t*)(w + L.subset_cols); P.subset_idx = (int32_t*)(w + L.subset_idx);
  P.col_cats = (uint32_t*)(w + L.col_cats); P.present = (uint32_t*)(w + L.present); P.n_subset = 0;
  P.vi = (int32_t*)(w + L.stats + (size_t)P.C * sizeof(bk_step_stats)); P.rec_stride = (int32_t)L.rec_stride; P.draws_out = nullptr; P.n_draw_peers = 0;
  P.trace = (bk_trace_rec*)(w + L.trace); P.sync = (ChainSync*)(w + L.sync); P.abort_flag = (int32_t*)(w + L.abort_flag);
  h->split_prior_dev = (double*)(w + L.split_prior);

  CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  CK(cudaMemcpyAsync(P.p_leaf, s->p_leaf, 256 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(h->split_prior_dev, s->split_prior, (size_t)P.p * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  {
    int32_t* rules_h = (int32_t*)calloc((size_t)P.p, sizeof(int32_t));
    int32_t* sidx_h = (int32_t*)malloc((size_t)P.p * sizeof(int32_t));
    if (!rules_h || !sidx_h) { free(rules_h); free(sidx_h); set_err("out of host memory"); return BK_ERR_ARG; }
    if (s->split_rules) memcpy(rules_h, s->split_rules, (size_t)P.p * sizeof(int32_t));
    int32_t scols_h[BK_MAX_SUBSET_COLS];
    for (int i = 0; i < BK_MAX_SUBSET_COLS; ++i) scols_h[i] = 0;
    for (int v = 0; v < P.p; ++v) {
      sidx_h[v] = -1;
      if (rules_h[v] == BK_RULE_SUBSET && P.n_subset < BK_MAX_SUBSET_COLS) { sidx_h[v] = P.n_subset; scols_h[P.n_subset++] = v; }
    }
    cudaError_t e = cudaMemcpyAsync(P.rules, rules_h, (size_t)P.p * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(P.subset_idx, sidx_h, (size_t)P.p * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(P.subset_cols, scols_h, sizeof(scols_h), cudaMemcpyHostToDevice, h->stream);
    cudaStreamSynchronize(h->stream);
    free(rules_h); free(sidx_h);
    CK(e);
  }
  // vi, stats and the abort flag are adjacent in the workspace: one D2H copy brings all three back
  h->out_off_dev = L.abort_flag; h->out_stats_off = L.stats - L.abort_flag; h->out_abort_off = 0;
  h->out_bytes = h->out_stats_off + (size_t)BK_MAX_STEPS_PER_LAUNCH * L.rec_stride;
  h->rec_stride = L.rec_stride;
  for (int k = 0; k < 2; ++k) {
    CK(cudaMallocHost(&h->out_slot[k], h->out_bytes));
    memset(h->out_slot[k], 0, h->out_bytes);
    CK(cudaEventCreateWithFlags(&h->done_ev[k], cudaEventDisableTiming));
  }
  h->stats_pinned = (bk_step_stats*)(h->out_slot[0] + h->out_stats_off);
  h->vi_pinned = (int32_t*)(h->out_slot[0] + h->out_stats_off + (size_t)P.C * sizeof(bk_step_stats));
  h->abort_pinned = (int32_t*)(h->out_slot[0] + h->out_abort_off);
  h->workspace = (char*)workspace_dev;
  memset(&h->args, 0, sizeof(h->args));

  int dev = s->device, n_sm = 0, coop = 0, occ = 0;
  CK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  CK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
  if (!coop) { set_err("device lacks cooperative launch"); return BK_ERR_UNSUPPORTED; }
  
  {
    // dynamic shared memory of the control CTAs: both particle buffers, header + fastF nodes each
    const int budget = 168 * 1024;
    int F = (budget / (2 * P.P) - (int)sizeof(PHdr)) / (int)sizeof(DNode);
    F = F > 63 ? 63 : F;
    if (F < 3) { set_err("too many particles for the shared-memory particle store"); return BK_ERR_ARG; }
    P.fastF = F; P.fast_stride = (int)sizeof(PHdr) + F * (int)sizeof(DNode);
    h->dyn_smem = (size_t)2 * P.P * P.fast_stride;
    // worker CTAs carve their groups out of it (-DBK_STAGE: then the staging slots of groups 2-3, round_unit's cp.async ping-pong)
#ifdef BK_STAGE
    if (h->dyn_smem < sizeof(GroupShared) * BK_NGROUPS + 2 * BK_STAGE_GROUP_BYTES) h->dyn_smem = sizeof(GroupShared) * BK_NGROUPS + 2 * BK_STAGE_GROUP_BYTES;
#else
    if (h->dyn_smem < sizeof(GroupShared) * BK_NGROUPS) h->dyn_smem = sizeof(GroupShared) * BK_NGROUPS;
#endif
    CK(cudaFuncSetAttribute(pgbart_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->dyn_smem));
  }
  h->grid = n_sm;  // one persistent CTA per SM (148 on B200): chains control CTAs + workers
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pgbart_step_kernel, BK_CTA_THREADS, h->dyn_smem));
  if (occ < 1) { set_err("step kernel does not fit on an SM"); return BK_ERR_CUDA; }
  if (h->grid <= P.C) { set_err("more chains than SMs minus one"); return BK_ERR_ARG; }
  if (getenv("BK_DEBUG_MARKERS")) {
    h->marker_count = 4736 + n_sm * 64 + 64;
    h->marker_host = (int32_t*)calloc((size_t)h->marker_count, sizeof(int32_t));
    int32_t* dptr = nullptr;
    CK(cudaMalloc(&dptr, (size_t)h->marker_count * sizeof(int32_t)));
    CK(cudaMemset(dptr, 0, (size_t)h->marker_count * sizeof(int32_t)));
    P.marker = dptr;
  }
  h->max_phases = 1 << 20;

  {
    // which columns hold missing values (NaN): one block per column; the answer also sizes the code paths (P.has_nan)
    pgbart_nan_scan_kernel<<<P.p, 256, 0, h->stream>>>(P);
    int32_t* flags = (int32_t*)malloc((size_t)P.p * sizeof(int32_t));
    if (!flags) { set_err("out of host memory"); return BK_ERR_ARG; }
    cudaError_t e = cudaMemcpyAsync(flags, P.col_nan, (size_t)P.p * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    P.has_nan = 0;
    int bad_codes = 0;
    for (int v = 0; v < P.p && e == cudaSuccess; ++v) { P.has_nan |= (flags[v] & 1) ? 1 : 0; bad_codes |= (flags[v] & 2) ? 1 : 0; }
    free(flags);
    CK(e);
    if (bad_codes) { set_err("a SubsetSplit column holds values that are not category codes (integers 0..23)"); return BK_ERR_ARG; }
  }
  pgbart_init_kernel<<<n_sm * 2, 512, 0, h->stream>>>(P, s->init_sum, s->leaf_sd_init, h->split_prior_dev);
  pgbart_init_cum_kernel<<<P.C, 32, 0, h->stream>>>(P);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(h->stream));
  return BK_OK;
}

extern "C" {

void bk_destroy(bk_handle* h) {
  if (!h) return;
  DeviceGuard guard_(h->s.device);
  if (h->stream) { cudaStreamSynchronize(h->stream); cudaStreamDestroy(h->stream); }
  for (int k = 0; k < 2; ++k) {
    if (h->out_slot[k]) cudaFreeHost(h->out_slot[k]);
    if (h->st_slot[k]) cudaFreeHost(h->st_slot[k]);
    if (h->hist_nodes_slot[k]) cudaFreeHost(h->hist_nodes_slot[k]);
    if (h->hist_nn_slot[k]) cudaFreeHost(h->hist_nn_slot[k]);
    if (h->done_ev[k]) cudaEventDestroy(h->done_ev[k]);
  }
  if (h->P.marker) cudaFree(h->P.marker);
  if (h->y_stage) cudaFreeHost(h->y_stage);
  free(h->marker_host);
  delete h;
}

int bk_step_launch(bk_handle* h, int tune, const float* sigma_host) { return bk_run_launch(h, 1, tune, sigma_host, nullptr); }

int bk_run_launch(bk_handle* h, int n_steps, int tune, const float* sigma_host, float* draws_dev) {
  if (!h) { set_err("NULL handle"); return BK_ERR_ARG; }
  if (n_steps < 1 || n_steps > BK_MAX_STEPS_PER_LAUNCH) { set_err("n_steps must be in [1, 16]"); return BK_ERR_ARG; }
  if (n_steps > 1 && h->s.trace_capacity > 0) { set_err("the trace needs one step per launch"); return BK_ERR_STATE; }
  if (n_steps > 1 && h->history && !tune && (long long)n_steps * h->P.batch_post > h->P.m) {
    set_err("with the tree history on, the steps of one launch must not rewrite a tree twice (n_steps * trees per step <= n_trees)"); return BK_ERR_ARG;
  }
  if (h->poisoned) { set_err("an earlier step timed out inside the kernel; the sampler state is undefined: create a new handle"); return BK_ERR_STATE; }
  if (h->n_launched - h->n_waited >= 2) { set_err("two steps are already in flight: call bk_step_wait first"); return BK_ERR_STATE; }
  ON_DEVICE(h->s.device);
  Params& P = h->P;
  for (int c = 0; c < P.C; ++c) {
    float sg = sigma_host ? sigma_host[c] : 1.0f;
    if (!(sg > 0.0f)) { set_err("sigma must be positive"); return BK_ERR_ARG; }
    h->args.sigma[c] = sg;
  }
  int tune_i = tune ? 1 : 0;
  int maxp = h->max_phases;
  {
    const char* dbg = getenv("BK_DEBUG");
    P.debug = dbg ? atoi(dbg) : 0;
  }
  const int slot = (int)(h->n_launched & 1);
  const int lo = h->host_lower, T = tune ? P.batch_tune : P.batch_post, hi = lo + T < P.m ? lo + T : P.m;
  h->args.n_steps = n_steps;
  P.draws_out = draws_dev;
  h->steps_in_slot[slot] = n_steps;
  // ONE kernel and ONE small D2H copy per step: the likelihood scales travel as kernel arguments, the epoch ids and
  // done counters of the dataflow run on across launches (no memset), the per-step outputs are adjacent
  void* args[] = {(void*)&P, (void*)&tune_i, (void*)&h->args, (void*)&maxp};
  CK(cudaLaunchCooperativeKernel((const void*)pgbart_step_kernel, dim3(h->grid), dim3(BK_CTA_THREADS), args, h->dyn_smem, h->stream));
  h->args.epoch_base += 1u << 20;   // (max_phases = 2^20 epochs per launch at most)
  CK(cudaMemcpyAsync(h->out_slot[slot], h->workspace + h->out_off_dev, h->out_stats_off + (size_t)n_steps * h->rec_stride, cudaMemcpyDeviceToHost, h->stream));
  if (h->host_output)   // the value handed back to PyMC: strided device rows -> dense pinned host rows, behind the kernel
    CK(cudaMemcpy2DAsync(h->st_slot[slot], (size_t)P.N * sizeof(float), P.st, (size_t)P.Npad * sizeof(float), (size_t)P.N * sizeof(float),
                         (size_t)P.C * P.K, cudaMemcpyDeviceToHost, h->stream));
  for (int sidx = 0; sidx < BK_MAX_STEPS_PER_LAUNCH; ++sidx) h->hist_count[slot][sidx] = 0;
  if (h->history && !tune) {
    // the trees every step of the launch rewrote (op.all_trees batches, pymc_bart/utils.py:117-127): two strided copies per
    // step behind the kernel, no stall (a tree is rewritten at most once per launch, so its nodes are still the step's
    // when the launch ends); bk_history_batch_at compacts them on the host
    const size_t Tmax = (size_t)(P.batch_tune > P.batch_post ? P.batch_tune : P.batch_post);
    if (h->hist_cap[slot] < n_steps) {   // (the slot is idle: its previous launch has been waited for)
      if (h->hist_nodes_slot[slot]) cudaFreeHost(h->hist_nodes_slot[slot]);
      if (h->hist_nn_slot[slot]) cudaFreeHost(h->hist_nn_slot[slot]);
      h->hist_nodes_slot[slot] = nullptr; h->hist_nn_slot[slot] = nullptr; h->hist_cap[slot] = 0;
      CK(cudaMallocHost(&h->hist_nodes_slot[slot], (size_t)n_steps * P.C * Tmax * BK_MAX_NODES * sizeof(DNode)));
      CK(cudaMallocHost(&h->hist_nn_slot[slot], (size_t)n_steps * P.C * Tmax * sizeof(int32_t)));
      h->hist_cap[slot] = n_steps;
    }
    int l = lo;
    for (int sidx = 0; sidx < n_steps; ++sidx) {
      const int hh = l + T < P.m ? l + T : P.m;
      CK(cudaMemcpy2DAsync(h->hist_nn_slot[slot] + (size_t)sidx * P.C * Tmax, Tmax * sizeof(int32_t), P.forest_nn + l, (size_t)P.m * sizeof(int32_t),
                           (size_t)(hh - l) * sizeof(int32_t), (size_t)P.C, cudaMemcpyDeviceToHost, h->stream));
      CK(cudaMemcpy2DAsync(h->hist_nodes_slot[slot] + (size_t)sidx * P.C * Tmax * BK_MAX_NODES, Tmax * BK_MAX_NODES * sizeof(DNode),
                           P.forest + (size_t)l * BK_MAX_NODES, (size_t)P.m * BK_MAX_NODES * sizeof(DNode),
                           (size_t)(hh - l) * BK_MAX_NODES * sizeof(DNode), (size_t)P.C, cudaMemcpyDeviceToHost, h->stream));
      h->hist_first[slot][sidx] = l; h->hist_count[slot][sidx] = hh - l;
      l = hh < P.m ? hh : 0;
    }
  }
  CK(cudaEventRecord(h->done_ev[slot], h->stream));
  {
    int l = lo;
    for (int sidx = 0; sidx < n_steps; ++sidx) { const int hh = l + T < P.m ? l + T : P.m; l = hh < P.m ? hh : 0; }
    h->host_lower = l;
  }
  h->n_launched += 1;
  return BK_OK;
}

int bk_set_draw_peers(bk_handle* h, int n_peers, const void* const* peer_bases, const void* local_base) {
  if (!h || n_peers < 0 || n_peers > BK_MAX_DRAW_PEERS || (n_peers > 0 && (!peer_bases || !local_base))) { set_err("bad argument"); return BK_ERR_ARG; }
  if (h->n_launched != h->n_waited) { set_err("steps in flight"); return BK_ERR_STATE; }
  for (int i = 0; i < n_peers; ++i) {
    if (!peer_bases[i]) { set_err("NULL peer buffer"); return BK_ERR_ARG; }
    h->P.draw_peer_delta[i] = (long long)((const char*)peer_bases[i] - (const char*)local_base);
  }
  h->P.n_draw_peers = n_peers;
  return BK_OK;
}

int bk_set_response(bk_handle* h, const float* y_host) {
  if (!h || !y_host) { set_err("bad argument"); return BK_ERR_ARG; }
  if (h->n_launched != h->n_waited) { set_err("steps in flight"); return BK_ERR_STATE; }
  ON_DEVICE(h->s.device);
  const Params& P = h->P;
  const size_t row = (size_t)P.N * sizeof(float);
  if (!h->y_stage) CK(cudaMallocHost(&h->y_stage, (size_t)P.G * row));
  else CK(cudaStreamSynchronize(h->stream));   // (an earlier copy out of the staging buffer has completed)
  memcpy(h->y_stage, y_host, (size_t)P.G * row);
  CK(cudaMemcpy2DAsync(const_cast<float*>(P.y), (size_t)P.Npad * sizeof(float), h->y_stage, row, row, (size_t)P.G, cudaMemcpyHostToDevice, h->stream));
  return BK_OK;
}

int bk_set_host_output(bk_handle* h, int enable) {
  if (!h) { set_err("NULL handle"); return BK_ERR_ARG; }
  ON_DEVICE(h->s.device);
  if (h->n_launched != h->n_waited) { set_err("steps in flight"); return BK_ERR_STATE; }
  for (int k = 0; k < 2 && enable; ++k)
    if (!h->st_slot[k]) CK(cudaMallocHost(&h->st_slot[k], (size_t)h->P.C * h->P.K * h->P.N * sizeof(float)));
  h->host_output = enable ? 1 : 0;
  return BK_OK;
}

const float* bk_sum_trees_host(bk_handle* h) { return (h && h->host_output) ? h->st_slot[h->last_slot] : nullptr; }

int bk_set_history(bk_handle* h, int enable) {
  if (!h) { set_err("NULL handle"); return BK_ERR_ARG; }
  ON_DEVICE(h->s.device);
  if (h->n_launched != h->n_waited) { set_err("steps in flight"); return BK_ERR_STATE; }
  const Params& P = h->P;
  const size_t Tmax = (size_t)(P.batch_tune > P.batch_post ? P.batch_tune : P.batch_post);
  if (enable > BK_MAX_STEPS_PER_LAUNCH) enable = BK_MAX_STEPS_PER_LAUNCH;
  for (int k = 0; k < 2 && enable > 0; ++k) {   // enable = the steps per launch the pinned buffers are sized for (they grow on demand)
    if (h->hist_cap[k] < enable) {
      if (h->hist_nodes_slot[k]) cudaFreeHost(h->hist_nodes_slot[k]);
      if (h->hist_nn_slot[k]) cudaFreeHost(h->hist_nn_slot[k]);
      h->hist_nodes_slot[k] = nullptr; h->hist_nn_slot[k] = nullptr; h->hist_cap[k] = 0;
      CK(cudaMallocHost(&h->hist_nodes_slot[k], (size_t)enable * P.C * Tmax * BK_MAX_NODES * sizeof(DNode)));
      CK(cudaMallocHost(&h->hist_nn_slot[k], (size_t)enable * P.C * Tmax * sizeof(int32_t)));
      h->hist_cap[k] = enable;
    }
  }
  h->history = enable ? 1 : 0;
  return BK_OK;
}

int bk_history_batch(bk_handle* h, int32_t* first_tree, int32_t* n_nodes_host, bk_node* nodes_host, int64_t* total_nodes) {
  return bk_history_batch_at(h, 0, first_tree, n_nodes_host, nodes_host, total_nodes);
}
int bk_history_values(bk_handle* h, float* values_host) { return bk_history_values_at(h, 0, values_host); }

int bk_history_batch_at(bk_handle* h, int step, int32_t* first_tree, int32_t* n_nodes_host, bk_node* nodes_host, int64_t* total_nodes) {
  if (!h || !first_tree || !n_nodes_host || !nodes_host || !total_nodes || step < 0 || step >= BK_MAX_STEPS_PER_LAUNCH) { set_err("bad argument"); return BK_ERR_ARG; }
  const Params& P = h->P;
  const int slot = h->last_slot, T = h->hist_count[slot][step];
  *first_tree = h->hist_first[slot][step]; *total_nodes = 0;
  if (!h->history || T <= 0) return 0;
  const size_t Tmax = (size_t)(P.batch_tune > P.batch_post ? P.batch_tune : P.batch_post);
  const size_t sbase = (size_t)step * P.C * Tmax;
  int64_t tot = 0;
  for (int c = 0; c < P.C; ++c)
    for (int t = 0; t < T; ++t) {
      const int nn = h->hist_nn_slot[slot][sbase + (size_t)c * Tmax + t];
      if (nn < 1 || nn > BK_MAX_NODES) { set_err("history batch holds a malformed tree"); return BK_ERR_STATE; }
      n_nodes_host[(size_t)c * T + t] = nn;
      const DNode* src = h->hist_nodes_slot[slot] + (sbase + (size_t)c * Tmax + t) * BK_MAX_NODES;
      for (int k = 0; k < nn; ++k) {
        bk_node* d = &nodes_host[tot + k];
        d->var = src[k].var; d->split = src[k].split; d->left = src[k].left; d->value = src[k].var < 0 ? src[k].value : 0.0f;
        d->n = src[k].n; d->depth = src[k].depth;
      }
      tot += nn;
    }
  *total_nodes = tot;
  return T;
}

/* leaf values of every output of the last history batch, [total_nodes][n_outputs] in the batch's node order */
int bk_history_values_at(bk_handle* h, int step, float* values_host) {
  if (!h || !values_host || step < 0 || step >= BK_MAX_STEPS_PER_LAUNCH) { set_err("bad argument"); return BK_ERR_ARG; }
  const Params& P = h->P;
  const int slot = h->last_slot, T = h->hist_count[slot][step];
  if (!h->history || T <= 0) return 0;
  const size_t Tmax = (size_t)(P.batch_tune > P.batch_post ? P.batch_tune : P.batch_post);
  const size_t sbase = (size_t)step * P.C * Tmax;
  size_t tot = 0;
  for (int c = 0; c < P.C; ++c)
    for (int t = 0; t < T; ++t) {
      const int nn = h->hist_nn_slot[slot][sbase + (size_t)c * Tmax + t];
      const DNode* src = h->hist_nodes_slot[slot] + (sbase + (size_t)c * Tmax + t) * BK_MAX_NODES;
      for (int k = 0; k < nn; ++k, ++tot)
        for (int j = 0; j < P.K; ++j) values_host[tot * P.K + j] = src[k].var < 0 ? node_val(src[k], j) : 0.0f;
    }
  return T;
}

/* leaf values of every output of a chain's current forest: values_host [n_trees][255][n_outputs] */
int bk_export_leaf_values(bk_handle* h, int chain, float* values_host) {
  if (!h || chain < 0 || chain >= h->P.C || !values_host) { set_err("bad argument"); return BK_ERR_ARG; }
  ON_DEVICE(h->s.device);
  CK(cudaStreamSynchronize(h->stream));
  const Params& P = h->P;
  const size_t cnt = (size_t)P.m * BK_MAX_NODES;
  DNode* tmp = (DNode*)malloc(cnt * sizeof(DNode));
  int32_t* nn = (int32_t*)malloc((size_t)P.m * sizeof(int32_t));
  if (!tmp || !nn) { free(tmp); free(nn); set_err("out of host memory"); return BK_ERR_ARG; }
  cudaError_t e = cudaMemcpy(tmp, P.forest + (size_t)chain * cnt, cnt * sizeof(DNode), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemcpy(nn, P.forest_nn + (size_t)chain * P.m, (size_t)P.m * sizeof(int32_t), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) { free(tmp); free(nn); set_err("CUDA error %s in bk_export_leaf_values", cudaGetErrorString(e)); return BK_ERR_CUDA; }
  for (int t = 0; t < P.m; ++t)
    for (int k = 0; k < BK_MAX_NODES; ++k)
      for (int j = 0; j < P.K; ++j) {
        const DNode& nd = tmp[(size_t)t * BK_MAX_NODES + k];
        values_host[((size_t)t * BK_MAX_NODES + k) * P.K + j] = (k < nn[t] && nd.var < 0) ? node_val(nd, j) : 0.0f;
      }
  free(tmp); free(nn);
  return BK_OK;
}

int bk_step_wait(bk_handle* h, int32_t* vi_counts_host, bk_step_stats* stats_host) { return bk_run_wait(h, vi_counts_host, stats_host); }

int bk_run_wait(bk_handle* h, int32_t* vi_counts_host, bk_step_stats* stats_host) {
  if (!h) { set_err("NULL handle"); return BK_ERR_ARG; }
  if (h->n_waited >= h->n_launched) { set_err("no step in flight"); return BK_ERR_STATE; }
  ON_DEVICE(h->s.device);
  Params& P = h->P;
  const int slot = (int)(h->n_waited & 1);
  CK(cudaEventSynchronize(h->done_ev[slot]));
  h->n_waited += 1;
  h->last_slot = slot;
  const int n_steps = h->steps_in_slot[slot];
  h->abort_pinned = (int32_t*)(h->out_slot[slot] + h->out_abort_off);
  if (*h->abort_pinned) { h->poisoned = 1; set_err("a dataflow wait timed out inside the step kernel"); return BK_ERR_TIMEOUT; }
  for (int sidx = 0; sidx < n_steps; ++sidx) {   // records of the launch's steps, in order: [n_steps][C] stats, [n_steps][C][p] counts
    unsigned char* rec = h->out_slot[slot] + h->out_stats_off + (size_t)sidx * h->rec_stride;
    h->stats_pinned = (bk_step_stats*)rec;                                      // (the last step's stay current for the trace readers)
    h->vi_pinned = (int32_t*)(rec + (size_t)P.C * sizeof(bk_step_stats));
    if (vi_counts_host) memcpy(vi_counts_host + (size_t)sidx * P.C * P.p, h->vi_pinned, (size_t)P.C * P.p * sizeof(int32_t));
    if (stats_host) memcpy(stats_host + (size_t)sidx * P.C, h->stats_pinned, (size_t)P.C * sizeof(bk_step_stats));
    for (int c = 0; c < P.C; ++c)
      if (h->stats_pinned[c].error_flags & ~1) {
        char buf[64]; snprintf(buf, sizeof(buf), "chain %d flags 0x%x", c, h->stats_pinned[c].error_flags);
        set_err("device-side consistency check failed: %s", buf); return BK_ERR_STATE;
      }
  }
  return BK_OK;
}

int bk_step(bk_handle* h, int tune, const float* sigma_host, int32_t* vi_counts_host, bk_step_stats* stats_host) {
  int rc = bk_step_launch(h, tune, sigma_host);
  if (rc != BK_OK) return rc;
  return bk_step_wait(h, vi_counts_host, stats_host);
}

void* bk_stream(bk_handle* h) { return h ? (void*)h->stream : nullptr; }

/* debug only: control sub-step timers (ns) of the last step */
int bk_debug_timers(bk_handle* h, int chain, unsigned long long* out8) {
  if (!h || chain < 0 || chain >= h->P.C) return BK_ERR_ARG;
  ChainCtl* c = h->P.ctl + chain;
  return cudaMemcpy(out8, (char*)c + offsetof(ChainCtl, hot) + offsetof(ChainHot, t_sub), 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost) == cudaSuccess ? BK_OK : BK_ERR_CUDA;
}

/* debug only: worker-side latency sums of the last steps, [grid][8] u64 (zeroed after reading); needs -DBK_PROFILE_CTRL */
int bk_debug_worker_timers(bk_handle* h, unsigned long long* out, int n_cta) {
#ifdef BK_PROFILE_CTRL
  if (!h || n_cta > 256) return BK_ERR_ARG;
  if (cudaMemcpyFromSymbol(out, g_wdbg, (size_t)n_cta * 16 * sizeof(unsigned long long)) != cudaSuccess) return BK_ERR_CUDA;
  if (cudaMemcpyFromSymbol(out + (size_t)n_cta * 16, g_cdbg, sizeof(g_cdbg)) != cudaSuccess) return BK_ERR_CUDA;
  { static unsigned long long z32[32]; cudaMemcpyToSymbol(g_cdbg, z32, sizeof(z32)); }
  static unsigned long long zeros[256 * 16];
  return cudaMemcpyToSymbol(g_wdbg, zeros, sizeof(zeros)) == cudaSuccess ? BK_OK : BK_ERR_CUDA;
#else
  (void)h; (void)out; (void)n_cta; return BK_ERR_UNSUPPORTED;
#endif
}

/* debug only (not part of the public header): host view of the per-warp progress markers */
int32_t* bk_debug_markers(bk_handle* h, int* count) {
  if (!h) return nullptr;
  if (count) *count = h->marker_count;
  if (!h->P.marker) return nullptr;
  cudaStream_t s2;
  cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking);
  cudaMemcpyAsync(h->marker_host, h->P.marker, (size_t)h->marker_count * sizeof(int32_t), cudaMemcpyDeviceToHost, s2);
  cudaStreamSynchronize(s2);
  cudaStreamDestroy(s2);
  return h->marker_host;
}

int bk_read_trace(bk_handle* h, int chain, bk_trace_rec* out_host, int capacity) {
  if (!h || chain < 0 || chain >= h->P.C || !out_host) { set_err("bad argument"); return BK_ERR_ARG; }
  if (h->n_launched != h->n_waited) { set_err("bk_read_trace: a step is in flight (the trace is the last step's)"); return BK_ERR_STATE; }
  ON_DEVICE(h->s.device);
  int n = h->stats_pinned[chain].trace_len;
  if (n > h->P.trace_cap) n = h->P.trace_cap;
  if (n > capacity) n = capacity;
  if (n <= 0) return 0;
  cudaError_t e = cudaMemcpy(out_host, h->P.trace + (size_t)chain * h->P.trace_cap, (size_t)n * sizeof(bk_trace_rec), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) { set_err("CUDA error %s in bk_read_trace", cudaGetErrorString(e)); return BK_ERR_CUDA; }
  return n;
}

int bk_export_trees(bk_handle* h, int chain, int first, int count, bk_node* nodes_host, int32_t* n_nodes_host) {
  if (!h || chain < 0 || chain >= h->P.C || !nodes_host || !n_nodes_host || first < 0 || count < 0 || first + count > h->P.m) {
    set_err("bad argument"); return BK_ERR_ARG;
  }
  if (count == 0) return BK_OK;
  ON_DEVICE(h->s.device);
  CK(cudaStreamSynchronize(h->stream));   // (steps still in flight settle first)
  const Params& P = h->P;
  CK(cudaMemcpy(n_nodes_host, P.forest_nn + (size_t)chain * P.m + first, (size_t)count * sizeof(int32_t), cudaMemcpyDeviceToHost));
  DNode* tmp = (DNode*)malloc((size_t)BK_MAX_NODES * sizeof(DNode));
  if (!tmp) { set_err("out of host memory"); return BK_ERR_ARG; }
  for (int t = 0; t < count; ++t) {
    int nn = n_nodes_host[t];
    cudaError_t e = cudaMemcpy(tmp, P.forest + ((size_t)chain * P.m + first + t) * BK_MAX_NODES, (size_t)nn * sizeof(DNode), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { free(tmp); set_err("CUDA error %s in bk_export_trees", cudaGetErrorString(e)); return BK_ERR_CUDA; }
    for (int k = 0; k < BK_MAX_NODES; ++k) {
      bk_node* d = &nodes_host[(size_t)t * BK_MAX_NODES + k];
      memset(d, 0, sizeof(*d));
      if (k < nn) { d->var = tmp[k].var; d->split = tmp[k].split; d->left = tmp[k].left; d->value = tmp[k].var < 0 ? tmp[k].value : 0.0f; d->n = tmp[k].n; d->depth = tmp[k].depth; }
    }
  }
  free(tmp);
  return BK_OK;
}

int bk_export_forest(bk_handle* h, int chain, bk_node* nodes_host, int32_t* n_nodes_host) {
  if (!h || chain < 0 || chain >= h->P.C || !nodes_host || !n_nodes_host) { set_err("bad argument"); return BK_ERR_ARG; }
  ON_DEVICE(h->s.device);
  CK(cudaStreamSynchronize(h->stream));
  const Params& P = h->P;
  size_t cnt = (size_t)P.m * BK_MAX_NODES;
  DNode* tmp = (DNode*)malloc(cnt * sizeof(DNode));
  if (!tmp) { set_err("out of host memory"); return BK_ERR_ARG; }
  cudaError_t e = cudaMemcpy(tmp, P.forest + (size_t)chain * cnt, cnt * sizeof(DNode), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemcpy(n_nodes_host, P.forest_nn + (size_t)chain * P.m, (size_t)P.m * sizeof(int32_t), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) { free(tmp); set_err("CUDA error %s in bk_export_forest", cudaGetErrorString(e)); return BK_ERR_CUDA; }
  for (int t = 0; t < P.m; ++t)
    for (int k = 0; k < BK_MAX_NODES; ++k) {
      bk_node* d = &nodes_host[(size_t)t * BK_MAX_NODES + k];
      memset(d, 0, sizeof(*d));
      if (k < n_nodes_host[t]) {
        const DNode& sN = tmp[(size_t)t * BK_MAX_NODES + k];
        d->var = sN.var; d->split = sN.split; d->left = sN.left; d->value = sN.var < 0 ? sN.value : 0.0f; d->n = sN.n; d->depth = sN.depth;
      }
    }
  free(tmp);
  return BK_OK;
}

int bk_export_leaf_ids(bk_handle* h, int chain, uint8_t* ids_host) {
  if (!h || chain < 0 || chain >= h->P.C || !ids_host) { set_err("bad argument"); return BK_ERR_ARG; }
  ON_DEVICE(h->s.device);
  CK(cudaStreamSynchronize(h->stream));
  const Params& P = h->P;
  CK(cudaMemcpy2D(ids_host, (size_t)P.N, P.ids_tree + (size_t)chain * P.m * P.Npad, (size_t)P.Npad, (size_t)P.N, (size_t)P.m,
                  cudaMemcpyDeviceToHost));
  return BK_OK;
}

}  // extern "C"
