/*
 * pgbart_b200.h — C ABI of the B200-native PGBART sampler core
 * (libpgbart_b200.so, built from pymc_bart_b200/csrc/ by __graft_entry__.build()).
 *
 * Drop-in boundary.  pymc-bart @4daa2e2 delegates its sampler to the un-vendored
 * PyO3 package `bartrs` (requirements.txt:6).  The entry points below are what a
 * binding for that path has to provide; each cites the reference site it stands
 * in for (paths relative to /root/reference):
 *
 *   bk_query_bytes/bk_create  <- bartrs.PGBART([rv], num_particles=..) constructor
 *                                (tests/test_bart.py:4,231-232) reading the op
 *                                attributes X, Y, m, alpha, beta, split_prior,
 *                                split_rules (pymc_bart/bart.py:141-158) and the
 *                                native `PySampler(PyBartSettings)` it builds
 *                                (pymc_bart/pymc_bart.py:2)
 *   bk_step                   <- PGBART.astep: one particle-Gibbs sweep over a batch
 *                                of trees; returns the new sum-of-trees value of the
 *                                BART variable (tests/test_bart.py:121-123,197) and
 *                                the per-draw variable-inclusion counts that the
 *                                shell encodes with _encode_vi (pymc_bart/utils.py:1387-1398)
 *   bk_export_forest          <- the (baseline_forest, batches) history published in
 *                                op.all_trees (pymc_bart/utils.py:117,124-127)
 *   bk_set_history/bk_history_batch <- the per-draw batches of that history, appended while sampling
 *                                (pymc_bart/bart.py:133-146)
 *   bk_predict_history        <- PosteriorSampler.from_history(...).sample_posterior(X, draw_indices,
 *                                excluded) (pymc_bart/utils.py:60-71,93-107,124-127)
 *   bk_pearson_r2             <- pearsonr2 of compute_variable_importance (pymc_bart/utils.py:1339-1346)
 *
 * Conventions: plain pointers and sizes only; 0 = success, negative = error
 * (message via bk_last_error); nothing throws across the boundary.  Device
 * pointers are BORROWED (the Python host owns them as torch tensors); one handle
 * drives `n_chains` independent chains on one GPU from one host thread.
 */
#ifndef PGBART_B200_H
#define PGBART_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BK_ABI_VERSION 3

#define BK_OK 0
#define BK_ERR_ARG (-1)
#define BK_ERR_CUDA (-2)
#define BK_ERR_STATE (-3)
#define BK_ERR_TIMEOUT (-4)
#define BK_ERR_UNSUPPORTED (-5)

#define BK_RULE_CONTINUOUS 0 /* "ContinuousSplit": x <= s  (tests/test_bart.py:143) */
#define BK_RULE_ONEHOT 1     /* "OneHotSplit":     x == s  (tests/test_bart.py:144) */
#define BK_RULE_SUBSET 2     /* "SubsetSplit":     category(x) in S (docs/api_reference.rst:16, pymc_bart/bart.py:103); the column holds
                                integer category codes 0..23, S travels as the float (mask of S) — bk_spec.h bk_subset_* */

typedef struct bk_handle_s bk_handle;

typedef struct {
  int32_t abi_version;   /* BK_ABI_VERSION */
  int32_t n_rows;        /* N observations */
  int32_t n_cols;        /* p covariates */
  int32_t n_trees;       /* m (bart.py:120) */
  int32_t n_particles;   /* P, PGBART(num_particles=) (tests/test_bart.py:231) */
  int32_t n_chains;      /* chains batched on this GPU */
  int32_t likelihood;    /* BK_LIK_NORMAL | BK_LIK_BERNOULLI_LOGIT (bk_spec.h) */
  int32_t qshift;        /* fixed-point scale 2^qshift for the sufficient statistics */
  int32_t batch_tune;    /* trees per step while tuning: max(1,int(m*batch[0])) */
  int32_t batch_post;    /* trees per step after tuning */
  uint32_t seed;         /* Philox key word 0 */
  uint32_t chain_base;   /* global index of local chain 0 (Philox key word 1) */
  float init_sum;        /* initial sum of trees = Y.mean() (bart.py:148) */
  float init_leaf;       /* initial leaf value Y.mean()/m */
  float leaf_sd_init;    /* Y.std()/sqrt(m), or 3/sqrt(m) for 0/1 data */
  int32_t device;        /* CUDA device ordinal */
  int32_t trace_capacity;/* trace records per chain per step (0 = off) */
  int32_t n_groups;      /* output groups with separate trees (BART(shape=(k,n), separate_trees=True)); 0/1 = single output */
  int32_t n_outputs;     /* leaf values per leaf with SHARED trees (BART(shape=(k,n)), tests/test_bart.py:107-123,140-164);
                            0/1 = single output; needs likelihood BK_LIK_NORMAL_HETERO / BK_LIK_CATEGORICAL, n_groups <= 1 */
  int32_t reserved0;
  const double* p_leaf;        /* [256] P(node at depth d stays a leaf) (bart.py:107-109) */
  const double* split_prior;   /* [n_cols] positive weights (bart.py:139,155) */
  const int32_t* split_rules;  /* [n_cols] BK_RULE_* (bart.py:156), NULL = all continuous */
} bk_settings;

/* per chain, per step */
typedef struct {
  int32_t tree_updates;   /* trees rewritten this step */
  int32_t rounds;         /* grow rounds summed over the trees */
  int32_t grow_events;    /* successful particle growths (G of BASELINE.md) */
  int32_t grow_root;      /* ... of which at the root (no leaf-id read) */
  int32_t count_passes;   /* member-count-only passes */
  int32_t phases;         /* control phases (= epochs published + 1) */
  int32_t trace_len;      /* trace records written */
  int32_t error_flags;    /* 0 = clean */
  float leaf_sd;          /* running leaf sd after the step */
  int32_t iter;           /* tree updates since creation */
  int32_t us_control;     /* wall time (us) this chain's control CTA spent in control phases */
  int32_t us_data;        /* ... waiting for the worker groups to finish its epochs */
  int32_t us_sync;        /* ... publishing epochs (release fence + descriptor store) */
  int32_t us_total;       /* kernel wall time seen by the control CTA */
  int32_t reserved[2];    /* [0]: part of us_data spent on SWEEP epochs */
} bk_step_stats;

/* one record per (tree update, round, particle>=1) plus one per tree update (kind 2) */
typedef struct {
  int32_t kind;       /* 1 = particle round record, 2 = tree committed */
  int32_t tree;       /* tree id */
  int32_t round;
  int32_t particle;   /* slot index in this round (kind 2: winning slot) */
  int32_t node;       /* popped node (-1: empty queue) */
  int32_t var;        /* split variable (-1: stayed leaf / no growth) */
  int32_t n_left;
  int32_t n_right;
  float split;
  float val_left;
  float val_right;
  int32_t ancestor;   /* slot copied into this slot by the resampling that follows (-1: none) */
  double log_w;       /* log-weight after the round */
  double aux;         /* kind 2: leaf_sd after commit */
} bk_trace_rec;

/* flat forest node, 24 bytes (export + prediction format) */
typedef struct {
  int32_t var;      /* split variable, -1 = leaf */
  float split;      /* split value */
  int32_t left;     /* index of the left child; right = left + 1 */
  float value;      /* leaf value (0 for split nodes) */
  int32_t n;        /* training rows that reached the node ("nvalue") */
  int32_t depth;
} bk_node;

int bk_abi_version(void);
const char* bk_last_error(void);

/* bytes of device workspace the host must allocate (as one torch.uint8 tensor) */
int bk_query_bytes(const bk_settings* s, size_t* workspace_bytes);

/* Row padding: every per-row array uses a leading dimension of
 * ld = bk_padded_rows(n_rows) (n_rows rounded up to a multiple of 256) so that
 * each column / chain starts 1 KiB-aligned for 128-bit loads. */
int bk_padded_rows(int n_rows);

/* X: [n_cols][ld] float32 (column-major copy of op.X, bart.py:209-210; padding
 * rows may hold anything); y: [n_groups][ld] float32 (padding 0); sum_trees:
 * [n_chains][max(n_groups, n_outputs)][ld] float32 (written by every step: the value handed back to PyMC); workspace:
 * bk_query_bytes bytes.  All four are device pointers that stay owned by the caller. */
int bk_create(const bk_settings* s, const float* X_dev, const float* y_dev,
              float* sum_trees_dev, void* workspace_dev, bk_handle** out);
void bk_destroy(bk_handle* h);

/* One PGBART step for every chain.  Per-chain arrays below have n_chains*n_groups entries
 * (index chain*n_groups + group).  sigma_host: likelihood scale of
 * the current point (ignored for Bernoulli).  vi_counts_host: [..][n_cols]
 * split-variable usage of the trees rewritten by this step (the vector that
 * pymc_bart/utils.py:1387 encodes).  stats_host: [n_chains] or NULL.
 * Launches one persistent kernel on the handle's stream and waits for it. */
int bk_step(bk_handle* h, int tune, const float* sigma_host, int32_t* vi_counts_host,
            bk_step_stats* stats_host);

/* Asynchronous form: bk_step_launch enqueues the step kernel and the D2H of the per-step outputs
 * on the handle's stream and returns; bk_step_wait blocks until the OLDEST step in flight is done
 * and copies its outputs out (same outputs as bk_step).  Up to two steps may be in flight (launch,
 * launch, wait, launch, wait, ...): the host prepares step k+1 while step k runs.  bk_stream returns the handle's
 * cudaStream_t (as void*) so the host can record events on it or order its own copies after it. */
int bk_step_launch(bk_handle* h, int tune, const float* sigma_host);
int bk_step_wait(bk_handle* h, int32_t* vi_counts_host, bk_step_stats* stats_host);

/* Several steps in ONE launch (1 <= n_steps <= 16) for drivers whose likelihood parameters do not change between steps
 * (the reference's `pm.sample` with a fixed scale; not usable when another step method updates sigma between draws):
 * every chain runs its n_steps back to back inside the persistent kernel, so chains do not wait for each other at step
 * boundaries and no launch gap separates the steps.  draws_dev: NULL, or device memory [n_steps][n_chains*max(n_groups,
 * n_outputs)][ld] float32 that receives the sum of trees after every step (the posterior draws stay on the device).
 * bk_run_wait returns the steps' outputs in order: vi_counts_host [n_steps][..][n_cols], stats_host [n_steps][..].
 * With the tree history on (bk_set_history) the steps of a launch must not rewrite a tree twice (n_steps * trees per
 * step <= n_trees); the trace needs one step per launch.  bk_step_launch/wait = n_steps 1. */
int bk_run_launch(bk_handle* h, int n_steps, int tune, const float* sigma_host, float* draws_dev);
int bk_run_wait(bk_handle* h, int32_t* vi_counts_host, bk_step_stats* stats_host);
/* Multi-GPU runs (chains are independent replicas, one process per GPU): the run's one collective is the
 * all-gather of the posterior draws.  With peers set, the commit sweep that writes a draw into draws_dev also stores it
 * into the same place of every peer's buffer (P2P stores over NVLink): local_base is this GPU's buffer, peer_bases[i]
 * this process's mapping of peer i's buffer of the same layout (e.g. torch symmetric memory), draws_dev of
 * bk_run_launch points into local_base.  After the last launch a barrier across the ranks is all that is left of the
 * gather.  n_peers = 0 switches it off.  At most 7 peers. */
int bk_set_draw_peers(bk_handle* h, int n_peers, const void* const* peer_bases, const void* local_base);
void* bk_stream(bk_handle* h);

/* Several BART variables in one likelihood (tests/test_bart.py:167-241, `pm.Normal("y", mu1 + mu2, sigma, observed=Y)`): the
 * step of one variable sees the response `observed - the other terms of the location` at the current point, which the
 * reference gets by re-evaluating its compiled datalogp with the other variables as shared inputs.  bk_set_response
 * replaces the response rows (y_host [n_groups][n_rows] float32, any host memory; staged through pinned memory, copied
 * on the handle's stream before the next launch; no step may be in flight).  The kernels read y afresh in every step.
 * Normal likelihood: the response is the data; other families would need the offset inside the linear predictor. */
int bk_set_response(bk_handle* h, const float* y_host);

/* Host copy of the value (tests/test_bart.py:121-123,197: the step returns the new value of the BART variable as a
 * host array).  With bk_set_host_output(h, 1) every step also copies the sum of trees into a pinned host buffer
 * [n_chains*n_groups][n_rows] behind the kernel on the handle's stream; bk_sum_trees_host returns that buffer (valid
 * after bk_step / bk_step_wait until the next launch; NULL when disabled). */
int bk_set_host_output(bk_handle* h, int enable);
const float* bk_sum_trees_host(bk_handle* h);

/* trace of the last step of one chain (host copy); returns records copied */
int bk_read_trace(bk_handle* h, int chain, bk_trace_rec* out_host, int capacity);

/* Current forest of a chain as flat nodes: nodes_host [n_trees][255], n_nodes_host [n_trees] */
int bk_export_forest(bk_handle* h, int chain, bk_node* nodes_host, int32_t* n_nodes_host);

/* `count` consecutive trees starting at `first` (the batch a step rewrote), same layout */
int bk_export_trees(bk_handle* h, int chain, int first, int count, bk_node* nodes_host, int32_t* n_nodes_host);

/* leaf assignment of every training row in every tree: ids_host [n_trees][n_rows] uint8 */
int bk_export_leaf_ids(bk_handle* h, int chain, uint8_t* ids_host);

/* Tree history (the `(baseline_forest, batches)` entries of op.all_trees, pymc_bart/utils.py:117,124-127; bart.py:133-146).
 * With bk_set_history(h, n >= 1) every post-tuning step also copies the trees it rewrote into pinned host memory behind
 * the kernel (asynchronous, the step does not stall); n = the steps per launch the pinned buffers are sized for up
 * front (1 for bk_step_launch; a larger bk_run_launch grows them when it comes).  0 switches the history off.  bk_history_batch hands out the batch of the last step waited
 * for: *first_tree, n_nodes_host [n_chains*n_groups][T] and the trees' nodes compacted back to back in the same order
 * (capacity n_chains*n_groups*T*255), *total_nodes of them.  Returns T, 0 for a tuning step / history off, < 0 on error. */
int bk_set_history(bk_handle* h, int enable);
int bk_history_batch(bk_handle* h, int32_t* first_tree, int32_t* n_nodes_host, bk_node* nodes_host, int64_t* total_nodes);
/* Shared-tree multi-output (n_outputs > 1; BART(shape=(k, n)), tests/test_bart.py:107-123,140-164): every leaf carries one
 * value per output.  bk_history_values: the values of the last batch's nodes, [total_nodes][n_outputs] in the order of
 * bk_history_batch; bk_export_leaf_values: the current forest's, [n_trees][255][n_outputs] (0 for split nodes). */
int bk_history_values(bk_handle* h, float* values_host);
int bk_export_leaf_values(bk_handle* h, int chain, float* values_host);
/* The same for step `step` of a launch of several steps (bk_run_launch); the plain forms are step 0. */
int bk_history_batch_at(bk_handle* h, int step, int32_t* first_tree, int32_t* n_nodes_host, bk_node* nodes_host, int64_t* total_nodes);
int bk_history_values_at(bk_handle* h, int step, float* values_host);

/* Posterior prediction from the forest history (row N1): what bartrs' PosteriorSampler.sample_posterior(X, draw_indices,
 * excluded) does for the shell (pymc_bart/utils.py:60-71,93-107), for all chains of an op in ONE launch.
 * nodes_dev: every stored tree version, compacted; ver_off_dev [n_versions + 1]: first node of a version;
 * ver_tbl_dev [n_forests][n_trees]: version of every tree of a forest (forest = chain, draw, output group);
 * max_forest_nodes: largest node total of a forest (sizes the shared-memory staging); X_dev [n][n_cols] float32
 * ROW-major new data; sel_dev: forest rows to evaluate, [n_sel] (shared by all masks) or, with sel_per_mask,
 * [n_masks][n_sel]; excluded_masks_dev [n_masks][n_cols] uint8 (1 = variable excluded: weighted descent by the
 * children's training counts) or NULL with n_masks = 0; split_rules_dev [n_cols] BK_RULE_* or NULL;
 * leaf_values_dev: NULL, or for shared-tree multi-output the values of every node, [total nodes][n_values] (then one
 * tree walk yields n_values outputs); out_dev [max(n_masks,1)][n_sel][max(n_values,1)][n] float32; err_dev: int32
 * device flag the caller cleared (non-zero afterwards = malformed history).  Runs on `stream` (cudaStream_t as void*). */
int bk_predict_history(int device, void* stream, const bk_node* nodes_dev, const int32_t* ver_off_dev,
                       const int32_t* ver_tbl_dev, int n_trees, int max_forest_nodes, const float* X_dev, int n,
                       int n_cols, const int32_t* sel_dev, int n_sel, int sel_per_mask,
                       const uint8_t* excluded_masks_dev, int n_masks, const int32_t* split_rules_dev,
                       const float* leaf_values_dev, int n_values, float* out_dev, int32_t* err_dev);

/* Squared Pearson correlation of the variable-importance search (row N4; pymc_bart/utils.py:1339-1346 `pearsonr2`,
 * called per posterior sample at :1003-1005 and :1038-1040).  a_dev [n_samples][len] (full model),
 * b_dev [n_subsets][n_samples][len]; out_dev [n_subsets][n_samples] float64. */
int bk_pearson_r2(int device, void* stream, const float* a_dev, const float* b_dev, int len, int n_samples,
                  int n_subsets, double* out_dev);

#ifdef __cplusplus
}
#endif
#endif /* PGBART_B200_H */
