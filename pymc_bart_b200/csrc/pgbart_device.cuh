// pgbart_device.cuh — device-side data layout of the B200 PGBART sampler.
//
// Everything a chain needs lives in HBM inside ONE workspace allocation that the
// Python host owns as a torch.uint8 tensor (see bk_query_bytes / bk_create in
// pgbart_b200.cu).  Names follow the reference's domain: forest, tree, node,
// particle, leaf id, split variable, not ML vocabulary.
#pragma once

#include <stdint.h>

#include "bk_spec.h"
#include "pgbart_b200.h"

#define BK_MAX_PARTICLES 128
#define BK_WARP_TILE 256           // rows per warp pass: 32 lanes x 8 rows
#define BK_ROWS_PER_LANE 8
#define BK_COARSE_TILES 32          // tiles per bucket of the coarse member counts (k-th member search at large N)
#define BK_COARSE_MIN_TILES 512    // more tiles than this: the per-tile counts no longer fit 16 per lane, use buckets
#ifndef BK_CTA_THREADS
#define BK_CTA_THREADS 1024
#endif
#ifndef BK_NGROUPS
#define BK_NGROUPS 4
#endif
//                       // independent worker groups per worker CTA
#define BK_GROUP_THREADS (BK_CTA_THREADS / BK_NGROUPS)
#define BK_COMMIT_TILE (BK_GROUP_THREADS * 4)  // rows per group pass in the commit/prologue sweep

// leaf-id row references
#define BK_ROW_VIRTUAL (-1)  // stump: every real row is in node 0
#define BK_ROW_FOREST (-2)   // particle 0: the current tree's row in ids_tree

// per-chain command of an epoch (what the worker groups execute next)
#define BK_CMD_IDLE 0
#define BK_CMD_ROUND 1
#define BK_CMD_SWEEP 2   // commit of tree A and/or prologue of tree B, fused
#define BK_CMD_DONE 3
#define BK_CMD_LL 4      // Bernoulli: log-likelihood terms of the rows of freshly made leaves

// chain state machine
#define BK_ST_START 0
#define BK_ST_WAIT_SWEEP 1
#define BK_ST_WAIT_ROUND 2
#define BK_ST_DONE 3
#define BK_ST_WAIT_LL 4

#define BK_MAX_DRAW_PEERS 7   // the other GPUs of one NVSwitch domain of 8
#ifndef BK_JOB_COPIES
#define BK_JOB_COPIES 1
#endif
#define BK_JOB_PARTITION 1
#define BK_JOB_COUNT 2
#define BK_MAX_STEPS_PER_LAUNCH 16
#define BK_JOB_NOP 0  // a partition job whose split value could not be drawn (only members with a missing covariate)
#define BK_JOB_LL 3   // src_row = the particle's new row, left_id, split = left leaf value, rule = bits of the right leaf value

struct __align__(16) DNode {
  int32_t var;   // -1 = leaf
  float split;
  int32_t left;  // right = left + 1
  int32_t depth;
  float value;
  int32_t n;
  int64_t sst;
  int64_t sr;
  int64_t aux[3];   // shared-tree multi-output: leaf values of outputs 1..6 as floats (output 0 is `value`)
};
static_assert(BK_MAX_OUTPUTS <= 7, "a device node holds 1 + 6 leaf values");
__host__ __device__ __forceinline__ float node_val(const DNode& nd, int j) { return j == 0 ? nd.value : reinterpret_cast<const float*>(nd.aux)[j - 1]; }
__host__ __device__ __forceinline__ void set_node_val(DNode& nd, int j, float v) { if (j == 0) nd.value = v; else reinterpret_cast<float*>(nd.aux)[j - 1] = v; }
static_assert(sizeof(DNode) == 64, "DNode must be 64 bytes");

struct __align__(16) DParticle {
  int32_t n_nodes;
  int32_t q_head;  // expansion queue = nodes [q_head, n_nodes)
  int32_t row;     // leaf-id row reference
  int32_t pad;
  double gain;
  double lw;
  DNode nodes[BK_MAX_NODES];
};
static_assert(sizeof(DParticle) == 32 + 64 * BK_MAX_NODES, "DParticle layout");

struct __align__(16) Job {
  int32_t kind;
  int32_t slot;       // particle slot (accumulator index)
  int32_t src_row;
  int32_t dst_row;
  int32_t node;       // node being split (partition) / unused (count)
  int32_t var;
  float split;
  int32_t left_id;    // id of the new left child; right = left_id + 1
  int32_t next_node;  // node whose members are counted per tile (-1: none)
  int32_t rule;
  int32_t pad[2];
};
static_assert(sizeof(Job) == 48, "Job layout");

struct __align__(16) SweepJob {
  int32_t do_commit;
  int32_t commit_tree;
  int32_t new_row;      // row holding the winning particle's leaf ids
  int32_t do_welford;
  int32_t do_prologue;
  int32_t prologue_tree;
  int32_t wf_count;
  int32_t draw_slot;    // 1 + step index whose final sum of trees this commit also writes to Params::draws_out (0: none)
};

// Scalar state of a chain.  The persistent part lives in global memory between steps; during a
// step the chain's control CTA works on a shared-memory copy (every access is ~30 cycles instead
// of an L2 round trip) and writes it back when the step is done.
struct __align__(16) ChainHot {
  int32_t tune;
  float sigma;
  int32_t iter;      // tree updates so far                  (persistent)
  int32_t lower;     // first tree of the next batch          (persistent)
  int32_t draw;      // steps so far (Philox counter word 0)  (persistent)
  int32_t step_in_launch;   // step index inside the current launch
  int32_t pad_step;
  int32_t wf_count;  // Welford count                         (persistent)
  float leaf_sd;     //                                       (persistent)
  float leaf_sdk[BK_MAX_OUTPUTS];   // shared-tree multi-output: running leaf sd per output (persistent; [0] mirrors leaf_sd)
  int32_t pad_sdk;
  int32_t stage;       // state machine position, read by every control thread at the start of a phase
  int32_t stage_next;  // written by thread 0 during the phase, moved into `stage` by control_loop
  int32_t tree_lo, tree_hi, cur_tree;
  int32_t round;
  int32_t buf;       // particle ping-pong index
  int32_t trace_len;
  int32_t trace_round_base;
  int32_t cmd;
  int32_t n_jobs;
  int32_t n_grow;    // partition jobs among them (listed first)
  int32_t c_tree_updates, c_rounds, c_grow, c_grow_root, c_count_passes, c_phases, c_err;
  double ll_inv2s2, ll_c;   // per-step constants of the Gaussian log-likelihood
  double r2_total;          // Gaussian: sum of squares of all rows' residuals for the tree being updated (bk_total_r2)
  unsigned long long t_sub_last;
  unsigned long long t_sub[8];  // optional control sub-step timers (ns), -DBK_PROFILE_CTRL
};

struct __align__(16) ChainCtl {
  ChainHot hot;         // global home of the scalar state
  SweepJob sweep;       // descriptor of the next SWEEP epoch (read by the workers)
  float old_vals[256];  // leaf values of the tree being replaced
  float new_vals[256];  // leaf values of the winning particle
  // shared-tree multi-output (K > 1): the same two tables per output, and the leaf values of the LL jobs
  float old_vals_k[BK_MAX_OUTPUTS][256];
  float new_vals_k[BK_MAX_OUTPUTS][256];
  float job_vals[BK_MAX_PARTICLES][2][BK_MAX_OUTPUTS];   // [job][left/right][output]
  Job jobs[BK_JOB_COPIES][BK_MAX_PARTICLES];   // the epoch's job list (optionally replicated; A/B: one copy is fastest)
};

// accumulator slots per particle (u64 each)
#define BK_ACC_N 0
#define BK_ACC_SST 1
#define BK_ACC_SR 2
#define BK_ACC_ND 3     // rows dropped from the split node by a missing covariate: count, sum q(sum_trees), sum q(r) / log-lik terms
#define BK_ACC_SSTD 4
#define BK_ACC_SRD 7
#define BK_ACC_LLL 5    // Bernoulli: quantised log-likelihood of the new left / right leaf
#define BK_ACC_LLR 6
#define BK_ACC_STRIDE 8

// acc0 layout per chain: [256][4]: leaf k -> (sr, -, -, -); entry 255 = totals
// (sr, sr2lo, sr2hi, sst); then [4]: (wf sd sum, Bernoulli terms of the rows in limbo, -, -)
#define BK_ACC0_STRIDE 4
#define BK_ACC0_WORDS (257 * BK_ACC0_STRIDE)

// Per-chain dataflow synchronisation (own 128-byte line each).
//   desc:   {epoch id, cmd, n_jobs, total units}: ONE aligned 16-byte store by the chain's control CTA after a release
//           fence; the worker groups poll it with one 16-byte load (a single L2 round trip tells a group that an epoch
//           started and what it is).
//   done:   serving groups that have finished their share of the step's epochs (workers: release add; the control
//           CTA polls, then fences).
//   pad[1..2] hold the publication time stamp of the profiling build (-DBK_PROFILE_CTRL).
struct __align__(128) ChainSync {
  uint4 desc;
  unsigned long long reserved0;   // (was the claim ticket of the dynamic scheduler)
  unsigned int done;
  unsigned int pad[25];
};
static_assert(sizeof(ChainSync) == 128, "ChainSync layout");

struct Params {
  int32_t N, Npad, p, m, P, C, R, ntiles;   // C = chains * groups ("virtual chains": one forest + one sum-of-trees row each)
  int32_t G;                               // output groups per chain (separate trees); y has G rows
  int32_t K;                               // leaf values per leaf (shared-tree multi-output); per-row arrays marked [C][K] then hold K rows per chain
  int32_t cnt_stride;                      // ntiles rounded up to a multiple of 4 (row stride of rowcnt)
  int32_t fastF, fast_stride;   // nodes per particle kept in the control CTA's shared memory; bytes per particle there
  int32_t lik, trace_cap, batch_tune, batch_post;
  int32_t has_nan;   // some column of X holds missing values
  float qscale, init_leaf;
  double inv_qscale;
  double inv_qm;     // 2^-qshift / m (leaf mean scale)
  uint32_t seed, chain_base;
  const float* X;   // [p][Npad]
  const float* y;   // [Npad]
  float* st;        // [C][K][Npad] sum of trees
  int32_t* qr;      // [C][K][Npad]  (Gaussian: q(r); other likelihoods: bits of the linear predictor without the tree)
  int32_t* qst;     // [C][K][Npad]
  uint8_t* ids_tree;   // [C][m][Npad]
  uint8_t* rows;       // [C][R][Npad]
  uint32_t* rowcnt;    // [C][R][cnt_stride]
  uint32_t* coarse;    // [C][R][nb_stride]: members of the row's counted node per bucket of BK_COARSE_TILES tiles (nb > 0 only)
  int32_t nb, nb_stride;
  float* wf_mean;      // [C][K][Npad]
  float* wf_m2;        // [C][K][Npad]
  DParticle* parts;    // [C][2][P]
  DNode* forest;       // [C][m][255]
  int32_t* forest_nn;  // [C][m]
  ChainCtl* ctl;       // [C]
  unsigned long long* accL;  // [C][P][BK_ACC_STRIDE]
  unsigned long long* acc0;  // [C][BK_ACC0_WORDS]
  unsigned long long* accK;  // [C][P][K][2]: sum q(sum_trees[j]) of the new left / right child (K > 1; zeroed by the reader)
  unsigned long long* acc_sd;  // [C][8]: running-sd sums per output (K > 1)
  double* alpha_vec;   // [C][p]
  double* cum;         // [C][p]
  double* p_leaf;      // [256]
  int32_t* rules;      // [p]
  int32_t* col_nan;    // [p] 1 = the column holds missing values (NaN)
  // SubsetSplit columns (bk_spec.h bk_subset_*; served by the same instantiation as the missing values)
  int32_t n_subset;          // columns with the subset rule
  int32_t* subset_cols;      // [BK_MAX_SUBSET_COLS] their column indices
  int32_t* subset_idx;       // [p] index of a column among them, -1 otherwise
  uint32_t* col_cats;        // [p] categories present in the whole column (the presence mask of a root node)
  uint32_t* present;         // [C][R][BK_MAX_SUBSET_COLS] categories present among the members of the node the row's
                             // per-tile counts belong to (the particle's next queue node), per subset column
  // Per-step outputs live in RECORDS [stats [C] | vi [C][p]], rec_stride bytes apart, one per step of a launch
  // (bk_run_launch runs up to BK_MAX_STEPS_PER_LAUNCH steps per launch); `vi` / `stats` point into record 0.
  int32_t* vi;         // [C][p] of record 0
  bk_step_stats* stats;  // [C] of record 0
  int32_t rec_stride;
  float* draws_out;    // per launch: [n_steps][C*K][Npad] the sum of trees after every step, or nullptr
  // Draws also stored into the same place of peer GPUs' buffers (bk_set_draw_peers): byte distance from this GPU's
  // buffer to each peer's mapping of its own.  The posterior all-gather of a multi-GPU run happens inside the commit sweep.
  int32_t n_draw_peers;
  long long draw_peer_delta[BK_MAX_DRAW_PEERS];
  bk_trace_rec* trace;   // [C][trace_cap]
  ChainSync* sync;   // [C]
  int32_t* abort_flag;
  int32_t debug;
  int32_t* marker;  // debug: mapped host memory, one int per warp
};
