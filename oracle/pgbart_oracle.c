/*
 * pgbart_oracle.c — CPU restatement of the PGBART step.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this; the product (pymc_bart_b200/) never does.
 *
 * PARITY UNPINNED.  pymc-bart @4daa2e2 holds no golden vector, known-answer test
 * or fixture for the sampler, and the sampler arithmetic itself lives in the
 * un-vendored Rust dependency `bartrs>=0.4.0` (requirements.txt:6, imported at
 * pymc_bart/__init__.py:15, pymc_bart/pymc_bart.py:2, tests/test_bart.py:4) which
 * cannot be built or imported offline.  This file therefore restates the published
 * particle-Gibbs BART algorithm (Quiroga et al., arXiv:2206.03619; the pure-Python
 * pymc-bart <= 0.12 `PGBART.astep`, summarised in SURVEY.md Appendix A) and is
 * anchored on the reference's own call sites and statistical tests:
 *   - inputs read from the op ............ pymc_bart/bart.py:141-158
 *   - depth prior alpha*(1+d)^-beta ...... pymc_bart/bart.py:107-109 (table passed in)
 *   - initial value Y.mean() ............. pymc_bart/bart.py:148
 *   - split rules "ContinuousSplit"/"OneHotSplit" tests/test_bart.py:143-145; "SubsetSplit" docs/api_reference.rst:16
 *   - variable_inclusion counts .......... pymc_bart/utils.py:1387-1398, tests/test_bart.py:59-64
 *   - VI dominance / prediction self-consistency: tests/test_bart.py:44-64, tests/test_utils.py:24-32
 * What does pin it: oracle/model_float.py, an independent float64 restatement without bk_spec.h, takes the same decisions
 * and agrees within 1e-5 on every value (tests/test_independent_model.py).
 * The only exact fixture the reference has for this path — the varint/base64 codec
 * round trip (tests/test_utils.py:101-113) — is checked in tests/test_codec.py.
 *
 * Structure: a deliberately plain, sequential, one-chain implementation with
 * explicit per-particle leaf-id arrays that are deep-copied on resampling.  It
 * shares with the kernels only include/bk_spec.h (RNG, fixed point, scalar
 * closed forms).  Everything order-dependent is written out here independently.
 *
 * Algorithm steps, per tree update:
 *   B1  r = y - (sum_trees - predict(old tree)), quantise r and sum_trees
 *   B2  particle 0 = old tree (never grows); 1..P-1 = stumps
 *   B3  pop one node per particle and round: stay leaf w.p. p_leaf[depth];
 *       variable ~ split prior; split value = X[k-th member, var]
 *   B4  partition the node's rows (x <= s / x == s / category(x) in S)
 *   B5  leaf value = mean(sum_trees over members)/m + z * leaf_sd
 *   B6  log-weight = Gaussian log-likelihood from per-leaf (n, sum r) and the tree-independent total sum r^2, or Bernoulli-logit
 *       log-likelihood from a second pass over the rows of the two new leaves (fixed-point terms)
 *   B7  w = exp(lw - max) + 1e-12 in fixed point; exact integer running sums
 *   B8  systematic resampling of particles 1..P-1
 *   B9  final systematic resampling over all P + uniform pick; commit
 *   B10 batch of max(1,int(m*batch)) trees per step, round robin
 *
 * Build: see oracle/Makefile (gcc -O3 -funroll-loops -march=x86-64-v3 -ffp-contract=off).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "bk_spec.h"
#include "pgbart_b200.h"

/* The particles of a round are independent (every random number is addressed by its counter, a particle touches only its
 * own nodes and leaf-id array), so a chain may spread them over host threads (bko_set_threads): same results bit for bit,
 * and the CPU baseline of bench.py uses every core the box has (chains x particle threads).  Plain pthreads: a handful of
 * threads per round pull particle indices from an atomic counter (libgomp is not in the image). */
#include <pthread.h>

typedef struct {
  void (*fn)(void* ctx, int index);
  void* ctx;
  int next, end;   /* next index to hand out (atomic), one past the last */
} pf_job;
static void* pf_worker(void* arg) {
  pf_job* j = (pf_job*)arg;
  for (;;) {
    int i = __atomic_fetch_add(&j->next, 1, __ATOMIC_RELAXED);
    if (i >= j->end) return NULL;
    j->fn(j->ctx, i);
  }
}
/* fn(ctx, i) for i in [lo, hi) on up to `threads` threads (the caller is one of them); returns when all are done */
static void parallel_for(int threads, int lo, int hi, void (*fn)(void*, int), void* ctx) {
  pf_job job; job.fn = fn; job.ctx = ctx; job.next = lo; job.end = hi;
  int extra = threads - 1;
  if (extra > hi - lo - 1) extra = hi - lo - 1;
  pthread_t th[64];
  int started = 0;
  if (extra > 64) extra = 64;
  for (int k = 0; k < extra; ++k) if (pthread_create(&th[started], NULL, pf_worker, &job) == 0) started++;
  pf_worker(&job);
  for (int k = 0; k < started; ++k) pthread_join(th[k], NULL);
}

typedef struct {
  int32_t var;   /* -1 leaf */
  float split;
  int32_t left;  /* right = left + 1 */
  int32_t depth;
  float value;
  float vals[BK_MAX_OUTPUTS]; /* shared-tree multi-output: one value per output (vals[0] == value) */
  bk_stats st;
  int64_t ll;    /* Bernoulli / multi-output: sum over member rows of the quantised log-likelihood terms at the leaf's value(s) */
} o_node;

typedef struct {
  int32_t n_nodes;
  int32_t q_head; /* expansion queue = nodes [q_head, n_nodes) in creation order */
  double gain;    /* Gaussian: sum over leaves of bk_leaf_gain */
  int64_t llq;    /* Bernoulli: sum of the leaves' ll */
  double lw;
  o_node nodes[BK_MAX_NODES];
  uint8_t* ids; /* [N] leaf id per row */
} o_particle;

typedef struct bko_s {
  bk_settings s;
  int chain, group;
  int N, p, m, P;
  const float* X; /* [p][N] borrowed */
  const float* y;
  double p_leaf[BK_MAX_DEPTH_TABLE];
  double* alpha_vec; /* split prior + tuned usage counts */
  double* cum;       /* normalised cumulative split prior in use */
  int32_t* rules;
  float qscale;
  double inv_qscale;
  double inv_qm;  /* 2^-qshift / m */
  float* st;      /* sum of trees */
  int32_t* qr;
  int32_t* qst;
  float* noi;
  o_particle* forest; /* m trees, ids = leaf assignment */
  o_particle* parts;  /* P */
  o_particle* tmp;    /* P scratch for resampling */
  float* wf_mean;
  float* wf_m2;
  int32_t wf_count;
  float leaf_sd;
  int32_t iter;
  int32_t lower;
  int32_t draw;
  bk_trace_rec* trace;
  int32_t trace_len, trace_cap;
  uint64_t* w;   /* running sums of the fixed-point weights */
  int32_t* anc;
  double r2_total;  /* Gaussian: sum of squares of all rows' residuals for the tree being updated */
  long long bytes_touched; /* rough algorithmic byte counter for the CPU baseline */
  int threads;             /* host threads of this chain's particle loops (bko_set_threads; 1 = sequential) */
  /* shared-tree multi-output (K = n_outputs > 1): per-output copies of the row arrays and of the running leaf sd */
  int K;
  float* stk;      /* [K][N] sum of trees */
  float* noik;     /* [K][N] */
  int32_t* qstk;   /* [K][N] */
  float* wf_meank; /* [K][N] */
  float* wf_m2k;   /* [K][N] */
  float leaf_sdk[BK_MAX_OUTPUTS];
} bko;

static void part_alloc(o_particle* q, int N) { q->ids = (uint8_t*)malloc((size_t)N); }
static void part_copy(o_particle* dst, const o_particle* src, int N) {
  uint8_t* keep = dst->ids;
  dst->n_nodes = src->n_nodes; dst->q_head = src->q_head; dst->gain = src->gain; dst->llq = src->llq; dst->lw = src->lw;
  memcpy(dst->nodes, src->nodes, sizeof(o_node) * (size_t)src->n_nodes);
  dst->ids = keep;
  memcpy(dst->ids, src->ids, (size_t)N);
}

static void rebuild_cum(bko* o) {
  double tot = 0.0;
  for (int v = 0; v < o->p; ++v) tot = BK_DADD(tot, o->alpha_vec[v]);
  double run = 0.0;
  for (int v = 0; v < o->p; ++v) {
    run = BK_DADD(run, o->alpha_vec[v]);
    o->cum[v] = BK_DDIV(run, tot);
  }
}

/* y: [n_groups][n_rows]; (chain_local, group) selects one chain and one output group (separate trees) */
int bko_create(const bk_settings* s, const float* X, const float* y, int chain_local, int group, bko** out) {
  if (!s || !X || !y || !out) return BK_ERR_ARG;
  if (s->n_particles < 2 || s->n_rows < 1 || s->n_trees < 1) return BK_ERR_ARG;
  bko* o = (bko*)calloc(1, sizeof(bko));
  o->s = *s; o->chain = chain_local; o->threads = 1;
  o->N = s->n_rows; o->p = s->n_cols; o->m = s->n_trees; o->P = s->n_particles;
  o->X = X; o->y = y + (size_t)group * (size_t)s->n_rows; o->group = group;
  memcpy(o->p_leaf, s->p_leaf, sizeof(double) * BK_MAX_DEPTH_TABLE);
  o->alpha_vec = (double*)malloc(sizeof(double) * (size_t)o->p);
  o->cum = (double*)malloc(sizeof(double) * (size_t)o->p);
  o->rules = (int32_t*)calloc((size_t)o->p, sizeof(int32_t));
  for (int v = 0; v < o->p; ++v) {
    o->alpha_vec[v] = s->split_prior[v];
    if (s->split_rules) o->rules[v] = s->split_rules[v];
  }
  rebuild_cum(o);
  o->qscale = ldexpf(1.0f, s->qshift);
  o->inv_qscale = ldexp(1.0, -s->qshift);
  o->inv_qm = BK_DDIV(o->inv_qscale, (double)o->m);
  int N = o->N;
  o->st = (float*)malloc(sizeof(float) * (size_t)N);
  o->noi = (float*)malloc(sizeof(float) * (size_t)N);
  o->qr = (int32_t*)malloc(sizeof(int32_t) * (size_t)N);
  o->qst = (int32_t*)malloc(sizeof(int32_t) * (size_t)N);
  o->wf_mean = (float*)calloc((size_t)N, sizeof(float));
  o->wf_m2 = (float*)calloc((size_t)N, sizeof(float));
  for (int i = 0; i < N; ++i) o->st[i] = s->init_sum;
  o->forest = (o_particle*)calloc((size_t)o->m, sizeof(o_particle));
  for (int t = 0; t < o->m; ++t) {
    o_particle* f = &o->forest[t];
    part_alloc(f, N);
    memset(f->ids, 0, (size_t)N);
    f->n_nodes = 1; f->q_head = 1;
    f->nodes[0].var = -1; f->nodes[0].left = -1; f->nodes[0].depth = 0;
    f->nodes[0].value = s->init_leaf; f->nodes[0].st.n = N;
  }
  o->parts = (o_particle*)calloc((size_t)o->P, sizeof(o_particle));
  o->tmp = (o_particle*)calloc((size_t)o->P, sizeof(o_particle));
  for (int q = 0; q < o->P; ++q) { part_alloc(&o->parts[q], N); part_alloc(&o->tmp[q], N); }
  o->leaf_sd = s->leaf_sd_init;
  o->K = s->n_outputs > 1 ? s->n_outputs : 1;
  if (o->K > 1) {
    if (o->K > BK_MAX_OUTPUTS || s->n_groups > 1) return BK_ERR_UNSUPPORTED;
    for (int v = 0; v < o->p; ++v) if (o->rules[v] == BK_RULE_SUBSET) return BK_ERR_UNSUPPORTED;   /* (device: the same) */
    const size_t KN = (size_t)o->K * (size_t)N;
    o->stk = (float*)malloc(sizeof(float) * KN);
    o->noik = (float*)malloc(sizeof(float) * KN);
    o->qstk = (int32_t*)malloc(sizeof(int32_t) * KN);
    o->wf_meank = (float*)calloc(KN, sizeof(float));
    o->wf_m2k = (float*)calloc(KN, sizeof(float));
    for (size_t i = 0; i < KN; ++i) o->stk[i] = s->init_sum;
    for (int j = 0; j < o->K; ++j) o->leaf_sdk[j] = s->leaf_sd_init;
    for (int t = 0; t < o->m; ++t) for (int j = 0; j < o->K; ++j) o->forest[t].nodes[0].vals[j] = s->init_leaf;
  }
  o->trace_cap = s->trace_capacity;
  if (o->trace_cap > 0) o->trace = (bk_trace_rec*)calloc((size_t)o->trace_cap, sizeof(bk_trace_rec));
  o->w = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)o->P);
  o->anc = (int32_t*)malloc(sizeof(int32_t) * (size_t)o->P);
  *out = o;
  return BK_OK;
}

void bko_destroy(bko* o) {
  if (!o) return;
  for (int t = 0; t < o->m; ++t) free(o->forest[t].ids);
  for (int q = 0; q < o->P; ++q) { free(o->parts[q].ids); free(o->tmp[q].ids); }
  free(o->forest); free(o->parts); free(o->tmp);
  free(o->alpha_vec); free(o->cum); free(o->rules);
  free(o->st); free(o->noi); free(o->qr); free(o->qst); free(o->wf_mean); free(o->wf_m2);
  free(o->trace); free(o->w); free(o->anc);
  free(o->stk); free(o->noik); free(o->qstk); free(o->wf_meank); free(o->wf_m2k);
  free(o);
}

static bk_trace_rec* trace_slot(bko* o) {
  if (!o->trace || o->trace_len >= o->trace_cap) { if (o->trace) o->trace_len++; return NULL; }
  bk_trace_rec* r = &o->trace[o->trace_len++];
  memset(r, 0, sizeof(*r));
  return r;
}

/* Gaussian log-likelihood of a whole particle from its per-leaf statistics,
 * leaves visited in node-index order */
static double particle_gain(const bko* o, const o_particle* q) {
  double g = 0.0;
  for (int k = 0; k < q->n_nodes; ++k)
    if (q->nodes[k].var < 0) g = BK_DADD(g, bk_leaf_gain(q->nodes[k].st, q->nodes[k].value, o->inv_qscale));
  return g;
}

/* Fixed-point weights W_i (bk_weight_fix), their exact integer running sums S_j, and systematic resampling:
 * ancestor of point (u + i)/L = first j with (i*2^32 + u32) * S_last <= S_j * L * 2^32, capped at L-1
 * (the inverse-CDF walk of App. A.7 without roundings; see bk_spec.h). */
static void weight_sums(const o_particle* parts, int first, int count, uint64_t* S) {
  double mx = parts[first].lw;
  for (int i = 1; i < count; ++i) if (parts[first + i].lw > mx) mx = parts[first + i].lw;
  uint64_t run = 0;
  for (int i = 0; i < count; ++i) { run += bk_weight_fix(parts[first + i].lw, mx); S[i] = run; }
}

static void systematic(const uint64_t* S, int L, uint32_t u32, int32_t* idx_out) {
  int idx = 0;
  for (int i = 0; i < L; ++i) {
    bk_u128 point = bk_resample_point((uint32_t)i, u32, S[L - 1]);
    while (!bk_resample_le(point, S[idx], (uint32_t)L) && idx < L - 1) idx += 1;
    idx_out[i] = idx;
  }
}

static int draw_variable(const bko* o, double u) {
  for (int v = 0; v < o->p; ++v) if (u < o->cum[v]) return v;
  return o->p - 1;
}

/* one grow attempt of particle slot `pi` at round `round`; returns 1 if it grew */
static int grow(bko* o, int tree, int round, int pi, float sigma, bk_trace_rec* rec, long long* bytes) {
  o_particle* q = &o->parts[pi];
  const int N = o->N;
  if (rec) { rec->node = -1; rec->var = -1; }
  if (q->q_head >= q->n_nodes) return 0;
  int j = q->q_head++;
  if (rec) rec->node = j;
  o_node* nd = &q->nodes[j];
  uint32_t S = o->s.seed, C = o->s.chain_base + (uint32_t)o->chain, D = (uint32_t)o->draw;
  int depth = nd->depth;
  double pl = depth < BK_MAX_DEPTH_TABLE ? o->p_leaf[depth] : 1.0;
  double u1 = bk_u01(bk_rng(S, C, D, (uint32_t)o->group, (uint32_t)tree, (uint32_t)round, (uint32_t)pi, BK_U_LEAF).v[0]);
  if (!(u1 > pl)) return 0;                       /* stays a leaf */
  if (q->n_nodes + 2 > BK_MAX_NODES) return 0;    /* node budget of one byte ids */
  double u2 = bk_u01(bk_rng(S, C, D, (uint32_t)o->group, (uint32_t)tree, (uint32_t)round, (uint32_t)pi, BK_U_VAR).v[0]);
  int v = draw_variable(o, u2);
  int n = nd->st.n;
  if (n < 2) return 0;                            /* fewer than two candidate split values */
  /* Split value = covariate of a uniformly drawn member with a value (members with a missing
   * covariate are not candidates).  bk_spec.h BK_SPLIT_TRIES: the four words of ONE Philox block give up to four
   * candidate members k_t = floor(x_t * n / 2^32) among ALL members of the node (ascending row index); the first one
   * whose covariate is not NaN supplies the value (rejection sampling: uniform over the members that have one); four
   * misses in a row leave the node a leaf.  Without missing values the first word decides, as before. */
  const bk_u32x4 wv = bk_rng(S, C, D, (uint32_t)o->group, (uint32_t)tree, (uint32_t)round, (uint32_t)pi, BK_U_VAL);
  const float* xc = o->X + (size_t)v * (size_t)N;
  float s = 0.0f;
  int have = 0;
  const int subset = o->rules[v] == BK_RULE_SUBSET;
  if (subset) {
    /* SubsetSplit (bk_spec.h bk_subset_draw): the categories present among the node's members
     * that have a value, then a uniformly drawn non-empty subset of them without the largest; fewer than two
     * categories present: no split.  The set travels as the float whose integer value is its bit mask. */
    uint32_t present = 0u;
    for (int i = 0; i < N; ++i) {
      if (q->ids[i] != (uint8_t)j) continue;
      const int code = bk_subset_code(xc[i]);
      if (code >= 0) present |= 1u << code;
    }
    const uint32_t mask = bk_subset_draw(present, wv.v[0]);
    s = (float)mask;
    have = mask != 0u;
  }
  for (int t = 0; t < BK_SPLIT_TRIES && !have && !subset; ++t) {
    uint32_t k = bk_index(wv.v[t], (uint32_t)n), seen = 0;
    for (int i = 0; i < N; ++i) {
      if (q->ids[i] == (uint8_t)j) { if (seen == k) { s = xc[i]; break; } seen++; }
    }
    have = !(s != s);
  }
  if (!have) return 0;
  int L = q->n_nodes, R = q->n_nodes + 1;
  bk_stats sl, sr;
  memset(&sl, 0, sizeof(sl)); memset(&sr, 0, sizeof(sr));
  int64_t ll_dropped = 0;   /* Bernoulli: terms of the rows dropped here, at the value 0 they now carry */
  const int onehot = o->rules[v] == BK_RULE_ONEHOT;
  for (int i = 0; i < N; ++i) {
    if (q->ids[i] != (uint8_t)j) continue;
    float x = xc[i];
    if (x != x) {   /* missing covariate: the row leaves the tree (limbo, predicts 0) and counts for neither child */
      q->ids[i] = BK_LIMBO;
      if (o->s.likelihood == BK_LIK_BERNOULLI_LOGIT) ll_dropped += (int64_t)bk_bern_q(o->y[i], o->noi[i], 0.0f);
      continue;
    }
    int left = subset ? bk_subset_left(x, s) : (onehot ? (x == s) : (x <= s));
    bk_stats* t = left ? &sl : &sr;
    q->ids[i] = (uint8_t)(left ? L : R);
    int64_t a = (int64_t)o->qr[i];
    t->n += 1; t->sst += (int64_t)o->qst[i]; t->sr += a;
  }
  *bytes += (long long)N * 14;
  double zl = bk_normal(bk_rng(S, C, D, (uint32_t)o->group, (uint32_t)tree, (uint32_t)round, (uint32_t)pi, BK_Z_LEFT));
  double zr = bk_normal(bk_rng(S, C, D, (uint32_t)o->group, (uint32_t)tree, (uint32_t)round, (uint32_t)pi, BK_Z_RIGHT));
  float vl = bk_leaf_value(sl.n, sl.sst, o->inv_qm, zl, o->leaf_sd);
  float vr = bk_leaf_value(sr.n, sr.sst, o->inv_qm, zr, o->leaf_sd);
  double g_parent = bk_leaf_gain(nd->st, nd->value, o->inv_qscale);
  nd->var = v; nd->split = s; nd->left = L;
  o_node* nl = &q->nodes[L]; o_node* nr = &q->nodes[R];
  nl->var = -1; nl->split = 0.0f; nl->left = -1; nl->depth = depth + 1; nl->value = vl; nl->st = sl;
  nr->var = -1; nr->split = 0.0f; nr->left = -1; nr->depth = depth + 1; nr->value = vr; nr->st = sr;
  q->n_nodes += 2;
  if (o->s.likelihood == BK_LIK_BERNOULLI_LOGIT) {
    /* no sufficient statistic: second pass over the rows of the two new leaves (+9N bytes) */
    int64_t ll_l = 0, ll_r = 0;
    for (int i = 0; i < N; ++i) {
      if (q->ids[i] == (uint8_t)L) ll_l += (int64_t)bk_bern_q(o->y[i], o->noi[i], vl);
      else if (q->ids[i] == (uint8_t)R) ll_r += (int64_t)bk_bern_q(o->y[i], o->noi[i], vr);
    }
    *bytes += (long long)N * 9;
    nl->ll = ll_l; nr->ll = ll_r;
    q->llq = q->llq - nd->ll + ll_l + ll_r + ll_dropped;
    q->lw = bk_bern_loglik((double)q->llq);
  } else {
    q->gain = BK_DADD(BK_DADD(BK_DSUB(q->gain, g_parent), bk_leaf_gain(sl, vl, o->inv_qscale)), bk_leaf_gain(sr, vr, o->inv_qscale));
    q->lw = bk_normal_loglik(bk_ssq_from_gain(o->r2_total, q->gain), sigma, (double)N);
  }
  if (rec) { rec->var = v; rec->split = s; rec->n_left = sl.n; rec->n_right = sr.n; rec->val_left = vl; rec->val_right = vr; }
  return 1;
}

static int bko_step_multi(bko* o, int tune, int32_t* vi_counts, bk_step_stats* stats);

/* one round of the particles 1..P-1 and the deep copies of a resampling, as parallel_for bodies */
typedef struct {
  bko* o; int tree, round; float sigma;
  bk_trace_rec** recs;
  int grew[128], was_root[128];
  long long bytes[128];
} round_ctx;
static void grow_one(void* ctx, int q) {
  round_ctx* rc = (round_ctx*)ctx;
  rc->grew[q] = grow(rc->o, rc->tree, rc->round, q, rc->sigma, rc->recs[q], &rc->bytes[q]);
  if (rc->recs[q]) rc->recs[q]->log_w = rc->o->parts[q].lw;
}
static void copy_one(void* ctx, int q) {
  bko* o = (bko*)ctx;
  part_copy(&o->tmp[q], &o->parts[o->anc[q - 1] + 1], o->N);
}

int bko_step(bko* o, int tune, float sigma, int32_t* vi_counts, bk_step_stats* stats) {
  if (o->K > 1) return bko_step_multi(o, tune, vi_counts, stats);
  const int bern = o->s.likelihood == BK_LIK_BERNOULLI_LOGIT;
  if (!bern && o->s.likelihood != BK_LIK_NORMAL) return BK_ERR_UNSUPPORTED;
  const int N = o->N, P = o->P, m = o->m;
  bk_step_stats loc; memset(&loc, 0, sizeof(loc));
  o->trace_len = 0;
  if (vi_counts) memset(vi_counts, 0, sizeof(int32_t) * (size_t)o->p);
  int T = tune ? o->s.batch_tune : o->s.batch_post;
  int upper = o->lower + T < m ? o->lower + T : m;
  uint32_t S = o->s.seed, C = o->s.chain_base + (uint32_t)o->chain, D = (uint32_t)o->draw;
  for (int t = o->lower; t < upper; ++t) {
    o->iter += 1;
    o_particle* old = &o->forest[t];
    /* B1: residual without tree t, fixed-point copies */
    bk_stats tot; memset(&tot, 0, sizeof(tot));
    bk_u128 tot_sr2 = bk_u128_make(0, 0);
    for (int k = 0; k < old->n_nodes; ++k) { bk_stats z; memset(&z, 0, sizeof(z)); z.n = old->nodes[k].st.n; old->nodes[k].st = z; old->nodes[k].ll = 0; }
    int64_t tot_ll = 0, limbo_ll = 0;
    for (int i = 0; i < N; ++i) {
      float oldp = old->ids[i] == BK_LIMBO ? 0.0f : old->nodes[old->ids[i]].value;
      float noi = BK_FSUB(o->st[i], oldp);
      float r = BK_FSUB(o->y[i], noi);
      o->noi[i] = noi;
      if (bern) {   /* per-row terms of the old tree's leaf and of the root-only stump */
        if (old->ids[i] != BK_LIMBO) old->nodes[old->ids[i]].ll += (int64_t)bk_bern_q(o->y[i], noi, oldp);
        else limbo_ll += (int64_t)bk_bern_q(o->y[i], noi, 0.0f);   /* rows the old tree dropped predict 0 */
        tot_ll += (int64_t)bk_bern_q(o->y[i], noi, o->s.init_leaf);
        r = 0.0f;   /* the Gaussian residual statistics are not used */
      }
      int32_t a = bk_quant(r, o->qscale), b = bk_quant(o->st[i], o->qscale);
      o->qr[i] = a; o->qst[i] = b;
      bk_u128 sq = bk_u128_make(0, (uint64_t)((int64_t)a * (int64_t)a));
      tot.n += 1; tot.sst += b; tot.sr += a; tot_sr2 = bk_u128_add(tot_sr2, sq);
      if (old->ids[i] != BK_LIMBO) old->nodes[old->ids[i]].st.sr += a;
    }
    o->bytes_touched += (long long)N * 23;
    o->r2_total = bk_total_r2(tot_sr2, o->inv_qscale);
    /* B2: particles */
    part_copy(&o->parts[0], old, N);
    o->parts[0].q_head = o->parts[0].n_nodes;
    if (bern) {
      int64_t llq = limbo_ll;
      for (int k = 0; k < old->n_nodes; ++k) if (old->nodes[k].var < 0) llq += old->nodes[k].ll;
      o->parts[0].gain = 0.0; o->parts[0].llq = llq; o->parts[0].lw = bk_bern_loglik((double)llq);
    } else {
      o->parts[0].gain = particle_gain(o, &o->parts[0]);
      o->parts[0].llq = 0;
      o->parts[0].lw = bk_normal_loglik(bk_ssq_from_gain(o->r2_total, o->parts[0].gain), sigma, (double)N);
    }
    for (int q = 1; q < P; ++q) {
      o_particle* pq = &o->parts[q];
      pq->n_nodes = 1; pq->q_head = 0;
      pq->nodes[0].var = -1; pq->nodes[0].split = 0.0f; pq->nodes[0].left = -1; pq->nodes[0].depth = 0;
      pq->nodes[0].value = o->s.init_leaf; pq->nodes[0].st = tot; pq->nodes[0].ll = tot_ll;
      memset(pq->ids, 0, (size_t)N);
      if (bern) { pq->gain = 0.0; pq->llq = tot_ll; pq->lw = bk_bern_loglik((double)tot_ll); }
      else {
        pq->llq = 0;
        pq->gain = bk_leaf_gain(tot, o->s.init_leaf, o->inv_qscale);
        pq->lw = bk_normal_loglik(bk_ssq_from_gain(o->r2_total, pq->gain), sigma, (double)N);
      }
    }
    /* B3-B8: grow rounds */
    int round = 0;
    for (;; ++round) {
      int tr0 = o->trace_len;
      bk_trace_rec* recs[128];   /* (P <= 128) record slots are handed out in particle order first */
      for (int q = 1; q < P; ++q) {
        bk_trace_rec* rec = trace_slot(o);
        if (rec) { rec->kind = 1; rec->tree = t; rec->round = round; rec->particle = q; rec->ancestor = -1; }
        recs[q] = rec;
      }
      round_ctx rc; rc.o = o; rc.tree = t; rc.round = round; rc.sigma = sigma; rc.recs = recs;
      memset(rc.grew, 0, sizeof(rc.grew)); memset(rc.bytes, 0, sizeof(rc.bytes));
      for (int q = 1; q < P; ++q) rc.was_root[q] = o->parts[q].q_head == 0;
      parallel_for(o->threads, 1, P, grow_one, &rc);
      for (int q = 1; q < P; ++q) {   /* (per-particle results are gathered in particle order: no shared counter in the loop) */
        if (rc.grew[q]) { loc.grow_events++; if (rc.was_root[q]) loc.grow_root++; }
        o->bytes_touched += rc.bytes[q];
      }
      loc.rounds++;
      int live = 0;
      for (int q = 1; q < P; ++q) if (o->parts[q].q_head < o->parts[q].n_nodes) live = 1;
      if (!live) break;
      weight_sums(o->parts, 1, P - 1, o->w);
      uint32_t u = bk_rng(S, C, D, (uint32_t)o->group, (uint32_t)t, (uint32_t)round, 0, BK_U_RESAMPLE).v[0];
      systematic(o->w, P - 1, u, o->anc);
      parallel_for(o->threads, 1, P, copy_one, o);
      for (int q = 1; q < P; ++q) {
        o_particle sw = o->parts[q]; o->parts[q] = o->tmp[q]; o->tmp[q] = sw;
        if (o->trace && tr0 + q - 1 < o->trace_cap) o->trace[tr0 + q - 1].ancestor = o->anc[q - 1] + 1;
      }
    }
    /* B9: final selection */
    weight_sums(o->parts, 0, P, o->w);
    uint32_t uf = bk_rng(S, C, D, (uint32_t)o->group, (uint32_t)t, 0xFFFFu, 0, BK_U_FINAL).v[0];
    systematic(o->w, P, uf, o->anc);
    uint32_t pick = bk_index(bk_rng(S, C, D, (uint32_t)o->group, (uint32_t)t, 0xFFFFu, 0, BK_U_PICK).v[0], (uint32_t)P);
    int win = o->anc[pick];
    o_particle* nw = &o->parts[win];
    /* commit: sum_trees = noi + predict(new) */
    double sd_acc_q = 0.0; int64_t sd_sum = 0;
    int do_wf = tune;
    if (do_wf) o->wf_count += 1;
    for (int i = 0; i < N; ++i) {
      float newp = nw->ids[i] == BK_LIMBO ? 0.0f : nw->nodes[nw->ids[i]].value;
      o->st[i] = BK_FADD(o->noi[i], newp);
      if (do_wf) {
        float cnt = (float)o->wf_count;
        float delta = BK_FSUB(newp, o->wf_mean[i]);
        float mean = BK_FADD(o->wf_mean[i], BK_FDIV(delta, cnt));
        float delta2 = BK_FSUB(newp, mean);
        float m2 = BK_FFMA(delta, delta2, o->wf_m2[i]);
        o->wf_mean[i] = mean; o->wf_m2[i] = m2;
        float sd = BK_FSQRT(BK_FDIV(m2, cnt));
        sd_sum += (int64_t)bk_quant(sd, o->qscale);
      }
    }
    (void)sd_acc_q;
    o->bytes_touched += (long long)N * (do_wf ? 26 : 10);
    if (tune) {
      if (o->iter > m) rebuild_cum(o);
      for (int k = 0; k < nw->n_nodes; ++k) if (nw->nodes[k].var >= 0) o->alpha_vec[nw->nodes[k].var] = BK_DADD(o->alpha_vec[nw->nodes[k].var], 1.0);
      if (o->iter > 2) o->leaf_sd = (float)BK_DDIV(BK_DMUL((double)sd_sum, o->inv_qscale), (double)N);
    } else if (vi_counts) {
      for (int k = 0; k < nw->n_nodes; ++k) if (nw->nodes[k].var >= 0) vi_counts[nw->nodes[k].var] += 1;
    }
    bk_trace_rec* rec = trace_slot(o);
    if (rec) { rec->kind = 2; rec->tree = t; rec->round = round; rec->particle = win; rec->node = nw->n_nodes; rec->var = -1; rec->ancestor = (int32_t)pick; rec->log_w = nw->lw; rec->aux = (double)o->leaf_sd; }
    part_copy(old, nw, N);
    old->q_head = old->n_nodes;
    loc.tree_updates++;
  }
  o->lower = upper < m ? upper : 0;
  o->draw += 1;
  loc.trace_len = o->trace_len; loc.leaf_sd = o->leaf_sd; loc.iter = o->iter;
  if (o->trace && o->trace_len > o->trace_cap) loc.error_flags |= 1;
  if (stats) *stats = loc;
  return BK_OK;
}

/* ------------------------------------------------------------------------
 * Shared-tree multi-output step (BART(shape=(k, n)) without separate trees: the reference's tested multi-output mode,
 * tests/test_bart.py:107-123 heteroscedastic Normal, :140-164 Categorical-softmax).  Every leaf carries K values;
 * the particle weight is the full-model log-likelihood of the (K, N) value, which has no
 * sufficient statistic: like the Bernoulli path, the rows of the two new leaves are revisited once the leaf values
 * are known and contribute quantised per-row terms (bk_lik_q); sums are exact integers.
 *   leaf value of output j: mean(sum_trees[j] over members)/m + z_j * leaf_sd[j], z_j = normal of the Philox block
 *   whose `group` word is j; running leaf sd per output.
 */
static int64_t multi_row_term(const bko* o, int i, const float* leaf_vals) {
  float f[BK_MAX_OUTPUTS];
  for (int j = 0; j < o->K; ++j) f[j] = BK_FADD(o->noik[(size_t)j * (size_t)o->N + (size_t)i], leaf_vals[j]);
  return (int64_t)bk_lik_q(o->s.likelihood, o->K, o->y[i], f);
}

static int grow_multi(bko* o, int tree, int round, int pi, bk_trace_rec* rec) {
  o_particle* q = &o->parts[pi];
  const int N = o->N, K = o->K;
  if (rec) { rec->node = -1; rec->var = -1; }
  if (q->q_head >= q->n_nodes) return 0;
  int j = q->q_head++;
  if (rec) rec->node = j;
  o_node* nd = &q->nodes[j];
  uint32_t S = o->s.seed, C = o->s.chain_base + (uint32_t)o->chain, D = (uint32_t)o->draw;
  int depth = nd->depth;
  double pl = depth < BK_MAX_DEPTH_TABLE ? o->p_leaf[depth] : 1.0;
  double u1 = bk_u01(bk_rng(S, C, D, 0u, (uint32_t)tree, (uint32_t)round, (uint32_t)pi, BK_U_LEAF).v[0]);
  if (!(u1 > pl)) return 0;
  if (q->n_nodes + 2 > BK_MAX_NODES) return 0;
  double u2 = bk_u01(bk_rng(S, C, D, 0u, (uint32_t)tree, (uint32_t)round, (uint32_t)pi, BK_U_VAR).v[0]);
  int v = draw_variable(o, u2);
  int n = nd->st.n;
  if (n < 2) return 0;
  uint32_t k = bk_index(bk_rng(S, C, D, 0u, (uint32_t)tree, (uint32_t)round, (uint32_t)pi, BK_U_VAL).v[0], (uint32_t)n);
  const float* xc = o->X + (size_t)v * (size_t)N;
  float s = 0.0f;
  {
    uint32_t seen = 0;
    for (int i = 0; i < N; ++i) if (q->ids[i] == (uint8_t)j) { if (seen == k) { s = xc[i]; break; } seen++; }
  }
  int L = q->n_nodes, R = q->n_nodes + 1;
  int32_t n_l = 0, n_r = 0;
  int64_t sst_l[BK_MAX_OUTPUTS], sst_r[BK_MAX_OUTPUTS];
  for (int jj = 0; jj < K; ++jj) { sst_l[jj] = 0; sst_r[jj] = 0; }
  const int onehot = o->rules[v] == BK_RULE_ONEHOT;
  for (int i = 0; i < N; ++i) {
    if (q->ids[i] != (uint8_t)j) continue;
    float x = xc[i];
    int left = onehot ? (x == s) : (x <= s);
    q->ids[i] = (uint8_t)(left ? L : R);
    if (left) n_l += 1; else n_r += 1;
    for (int jj = 0; jj < K; ++jj) {
      int64_t b = (int64_t)o->qstk[(size_t)jj * (size_t)N + (size_t)i];
      if (left) sst_l[jj] += b; else sst_r[jj] += b;
    }
  }
  nd->var = v; nd->split = s; nd->left = L;
  o_node* nl = &q->nodes[L]; o_node* nr = &q->nodes[R];
  memset(nl, 0, sizeof(*nl)); memset(nr, 0, sizeof(*nr));
  nl->var = -1; nl->left = -1; nl->depth = depth + 1; nl->st.n = n_l;
  nr->var = -1; nr->left = -1; nr->depth = depth + 1; nr->st.n = n_r;
  for (int jj = 0; jj < K; ++jj) {
    double zl = bk_normal(bk_rng(S, C, D, (uint32_t)jj, (uint32_t)tree, (uint32_t)round, (uint32_t)pi, BK_Z_LEFT));
    double zr = bk_normal(bk_rng(S, C, D, (uint32_t)jj, (uint32_t)tree, (uint32_t)round, (uint32_t)pi, BK_Z_RIGHT));
    nl->vals[jj] = bk_leaf_value(n_l, sst_l[jj], o->inv_qm, zl, o->leaf_sdk[jj]);
    nr->vals[jj] = bk_leaf_value(n_r, sst_r[jj], o->inv_qm, zr, o->leaf_sdk[jj]);
  }
  nl->value = nl->vals[0]; nr->value = nr->vals[0];
  nl->st.sst = sst_l[0]; nr->st.sst = sst_r[0];
  q->n_nodes += 2;
  int64_t ll_l = 0, ll_r = 0;
  for (int i = 0; i < N; ++i) {
    if (q->ids[i] == (uint8_t)L) ll_l += multi_row_term(o, i, nl->vals);
    else if (q->ids[i] == (uint8_t)R) ll_r += multi_row_term(o, i, nr->vals);
  }
  nl->ll = ll_l; nr->ll = ll_r;
  q->llq = q->llq - nd->ll + ll_l + ll_r;
  q->lw = bk_bern_loglik((double)q->llq);
  o->bytes_touched += (long long)N * (6 + 4 * K) + (long long)N * (5 + 4 * K);
  if (rec) { rec->var = v; rec->split = s; rec->n_left = n_l; rec->n_right = n_r; rec->val_left = nl->vals[0]; rec->val_right = nr->vals[0]; }
  return 1;
}

static int bko_step_multi(bko* o, int tune, int32_t* vi_counts, bk_step_stats* stats) {
  const int lik = o->s.likelihood;
  if (lik != BK_LIK_NORMAL_HETERO && lik != BK_LIK_CATEGORICAL) return BK_ERR_UNSUPPORTED;
  const int N = o->N, P = o->P, m = o->m, K = o->K;
  bk_step_stats loc; memset(&loc, 0, sizeof(loc));
  o->trace_len = 0;
  if (vi_counts) memset(vi_counts, 0, sizeof(int32_t) * (size_t)o->p);
  int T = tune ? o->s.batch_tune : o->s.batch_post;
  int upper = o->lower + T < m ? o->lower + T : m;
  uint32_t S = o->s.seed, C = o->s.chain_base + (uint32_t)o->chain, D = (uint32_t)o->draw;
  float init_vals[BK_MAX_OUTPUTS];
  for (int j = 0; j < K; ++j) init_vals[j] = o->s.init_leaf;
  for (int t = o->lower; t < upper; ++t) {
    o->iter += 1;
    o_particle* old = &o->forest[t];
    /* B1: linear predictors without tree t, fixed-point copies of the sums of trees, per-leaf terms of the old tree */
    for (int k = 0; k < old->n_nodes; ++k) old->nodes[k].ll = 0;
    int64_t tot_ll = 0, tot_sst0 = 0;
    for (int i = 0; i < N; ++i) {
      const o_node* on = &old->nodes[old->ids[i]];
      for (int j = 0; j < K; ++j) {
        const size_t ix = (size_t)j * (size_t)N + (size_t)i;
        o->noik[ix] = BK_FSUB(o->stk[ix], on->vals[j]);
        o->qstk[ix] = bk_quant(o->stk[ix], o->qscale);
      }
      tot_sst0 += (int64_t)o->qstk[i];
      old->nodes[old->ids[i]].ll += multi_row_term(o, i, on->vals);
      tot_ll += multi_row_term(o, i, init_vals);
    }
    /* B2: particles */
    part_copy(&o->parts[0], old, N);
    o->parts[0].q_head = o->parts[0].n_nodes;
    {
      int64_t llq = 0;
      for (int k = 0; k < old->n_nodes; ++k) if (old->nodes[k].var < 0) llq += old->nodes[k].ll;
      o->parts[0].gain = 0.0; o->parts[0].llq = llq; o->parts[0].lw = bk_bern_loglik((double)llq);
    }
    for (int q = 1; q < P; ++q) {
      o_particle* pq = &o->parts[q];
      pq->n_nodes = 1; pq->q_head = 0;
      memset(&pq->nodes[0], 0, sizeof(o_node));
      pq->nodes[0].var = -1; pq->nodes[0].left = -1;
      pq->nodes[0].value = o->s.init_leaf;
      for (int j = 0; j < K; ++j) pq->nodes[0].vals[j] = o->s.init_leaf;
      pq->nodes[0].st.n = N; pq->nodes[0].st.sst = tot_sst0; pq->nodes[0].ll = tot_ll;
      memset(pq->ids, 0, (size_t)N);
      pq->gain = 0.0; pq->llq = tot_ll; pq->lw = bk_bern_loglik((double)tot_ll);
    }
    /* B3-B8: grow rounds */
    int round = 0;
    for (;; ++round) {
      int tr0 = o->trace_len;
      for (int q = 1; q < P; ++q) {
        bk_trace_rec* rec = trace_slot(o);
        if (rec) { rec->kind = 1; rec->tree = t; rec->round = round; rec->particle = q; rec->ancestor = -1; }
        int was_root = o->parts[q].q_head == 0;
        if (grow_multi(o, t, round, q, rec)) { loc.grow_events++; if (was_root) loc.grow_root++; }
        if (rec) rec->log_w = o->parts[q].lw;
      }
      loc.rounds++;
      int live = 0;
      for (int q = 1; q < P; ++q) if (o->parts[q].q_head < o->parts[q].n_nodes) live = 1;
      if (!live) break;
      weight_sums(o->parts, 1, P - 1, o->w);
      uint32_t u = bk_rng(S, C, D, 0u, (uint32_t)t, (uint32_t)round, 0, BK_U_RESAMPLE).v[0];
      systematic(o->w, P - 1, u, o->anc);
      for (int q = 1; q < P; ++q) part_copy(&o->tmp[q], &o->parts[o->anc[q - 1] + 1], N);
      for (int q = 1; q < P; ++q) {
        o_particle sw = o->parts[q]; o->parts[q] = o->tmp[q]; o->tmp[q] = sw;
        if (o->trace && tr0 + q - 1 < o->trace_cap) o->trace[tr0 + q - 1].ancestor = o->anc[q - 1] + 1;
      }
    }
    /* B9: final selection and commit, per output */
    weight_sums(o->parts, 0, P, o->w);
    uint32_t uf = bk_rng(S, C, D, 0u, (uint32_t)t, 0xFFFFu, 0, BK_U_FINAL).v[0];
    systematic(o->w, P, uf, o->anc);
    uint32_t pick = bk_index(bk_rng(S, C, D, 0u, (uint32_t)t, 0xFFFFu, 0, BK_U_PICK).v[0], (uint32_t)P);
    int win = o->anc[pick];
    o_particle* nw = &o->parts[win];
    if (tune) o->wf_count += 1;
    for (int j = 0; j < K; ++j) {
      int64_t sd_sum = 0;
      for (int i = 0; i < N; ++i) {
        const size_t ix = (size_t)j * (size_t)N + (size_t)i;
        float newp = nw->nodes[nw->ids[i]].vals[j];
        o->stk[ix] = BK_FADD(o->noik[ix], newp);
        if (tune) {
          float cnt = (float)o->wf_count;
          float delta = BK_FSUB(newp, o->wf_meank[ix]);
          float mean = BK_FADD(o->wf_meank[ix], BK_FDIV(delta, cnt));
          float delta2 = BK_FSUB(newp, mean);
          float m2 = BK_FFMA(delta, delta2, o->wf_m2k[ix]);
          o->wf_meank[ix] = mean; o->wf_m2k[ix] = m2;
          sd_sum += (int64_t)bk_quant(BK_FSQRT(BK_FDIV(m2, cnt)), o->qscale);
        }
      }
      if (tune && o->iter > 2) o->leaf_sdk[j] = (float)BK_DDIV(BK_DMUL((double)sd_sum, o->inv_qscale), (double)N);
    }
    o->leaf_sd = o->leaf_sdk[0];
    if (tune) {
      if (o->iter > m) rebuild_cum(o);
      for (int k = 0; k < nw->n_nodes; ++k) if (nw->nodes[k].var >= 0) o->alpha_vec[nw->nodes[k].var] = BK_DADD(o->alpha_vec[nw->nodes[k].var], 1.0);
    } else if (vi_counts) {
      for (int k = 0; k < nw->n_nodes; ++k) if (nw->nodes[k].var >= 0) vi_counts[nw->nodes[k].var] += 1;
    }
    bk_trace_rec* rec = trace_slot(o);
    if (rec) { rec->kind = 2; rec->tree = t; rec->round = round; rec->particle = win; rec->node = nw->n_nodes; rec->var = -1; rec->ancestor = (int32_t)pick; rec->log_w = nw->lw; rec->aux = (double)o->leaf_sdk[K - 1]; }
    part_copy(old, nw, N);
    old->q_head = old->n_nodes;
    loc.tree_updates++;
  }
  o->lower = upper < m ? upper : 0;
  o->draw += 1;
  loc.trace_len = o->trace_len; loc.leaf_sd = o->leaf_sdk[0]; loc.iter = o->iter;
  if (o->trace && o->trace_len > o->trace_cap) loc.error_flags |= 1;
  if (stats) *stats = loc;
  return BK_OK;
}

/* out: [K][N] for shared-tree multi-output, [N] otherwise */
int bko_sum_trees(const bko* o, float* out) {
  if (o->K > 1) memcpy(out, o->stk, sizeof(float) * (size_t)o->K * (size_t)o->N);
  else memcpy(out, o->st, sizeof(float) * (size_t)o->N);
  return BK_OK;
}

/* leaf values of every output: vals [n_trees][255][K] (shared-tree multi-output) */
int bko_export_leaf_values(const bko* o, float* vals) {
  for (int t = 0; t < o->m; ++t)
    for (int k = 0; k < BK_MAX_NODES; ++k)
      for (int j = 0; j < o->K; ++j)
        vals[((size_t)t * BK_MAX_NODES + k) * (size_t)o->K + j] =
            (k < o->forest[t].n_nodes && o->forest[t].nodes[k].var < 0) ? (o->K > 1 ? o->forest[t].nodes[k].vals[j] : o->forest[t].nodes[k].value) : 0.0f;
  return BK_OK;
}

int bko_read_trace(const bko* o, bk_trace_rec* out, int capacity) {
  int n = o->trace_len < o->trace_cap ? o->trace_len : o->trace_cap;
  if (n > capacity) n = capacity;
  if (n > 0) memcpy(out, o->trace, sizeof(bk_trace_rec) * (size_t)n);
  return n;
}

int bko_export_forest(const bko* o, bk_node* nodes, int32_t* n_nodes) {
  for (int t = 0; t < o->m; ++t) {
    const o_particle* f = &o->forest[t];
    n_nodes[t] = f->n_nodes;
    for (int k = 0; k < BK_MAX_NODES; ++k) {
      bk_node* d = &nodes[(size_t)t * BK_MAX_NODES + k];
      memset(d, 0, sizeof(*d));
      if (k < f->n_nodes) {
        d->var = f->nodes[k].var; d->split = f->nodes[k].split; d->left = f->nodes[k].left;
        d->value = f->nodes[k].var < 0 ? f->nodes[k].value : 0.0f; d->n = f->nodes[k].st.n; d->depth = f->nodes[k].depth;
      }
    }
  }
  return BK_OK;
}

int bko_export_leaf_ids(const bko* o, uint8_t* ids) {
  for (int t = 0; t < o->m; ++t) memcpy(ids + (size_t)t * (size_t)o->N, o->forest[t].ids, (size_t)o->N);
  return BK_OK;
}

long long bko_bytes_touched(const bko* o) { return o->bytes_touched; }
/* host threads for this chain's particle loops (single-output step); returns the number in effect */
int bko_set_threads(bko* o, int n) {
  o->threads = n < 1 ? 1 : (n > 65 ? 65 : n);
  return o->threads;
}
float bko_leaf_sd(const bko* o) { return o->leaf_sd; }

/* ------------------------------------------------------------------------
 * Posterior prediction restatement (boundary:
 * pymc_bart/utils.py:60-71).  Row-major X_new [n][p]; weighted descent when the
 * split variable is excluded; NaN compares false (goes right).
 */
static double predict_tree(const bk_node* nodes, const float* x, const uint8_t* excl, const int32_t* rules) {
  /* explicit stack of (node, weight); the right child is pushed first so the left is visited first */
  int sn[130]; double sw[130]; int sp = 1;   /* depth + 2 <= 130 entries for trees of at most 255 nodes: cannot overflow */
  sn[0] = 0; sw[0] = 1.0;
  double tv = 0.0;
  while (sp > 0) {
    --sp;
    int k = sn[sp]; double w = sw[sp];
    const bk_node* nd = &nodes[k];
    if (nd->var < 0) { tv = BK_DFMA(w, (double)nd->value, tv); continue; }
    int l = nd->left, r = nd->left + 1;
    if (excl && excl[nd->var]) {
      double tot = (double)nodes[l].n + (double)nodes[r].n;
      if (!(tot > 0.0) || sp + 2 > 130) continue;
      double wl = BK_DDIV((double)nodes[l].n, tot);
      double wr = BK_DSUB(1.0, wl);
      sn[sp] = r; sw[sp] = BK_DMUL(w, wr); ++sp;
      sn[sp] = l; sw[sp] = BK_DMUL(w, wl); ++sp;
    } else {
      float xv = x[nd->var];
      const int rule = rules ? rules[nd->var] : BK_RULE_CONTINUOUS;
      int left = rule == BK_RULE_SUBSET ? bk_subset_left(xv, nd->split) : (rule == BK_RULE_ONEHOT ? (xv == nd->split) : (xv <= nd->split));
      sn[sp] = left ? l : r; sw[sp] = w; ++sp;
    }
  }
  return tv;
}

int bko_predict(const bk_node* forests, const int32_t* n_nodes, int n_trees, const float* X, int n, int n_cols,
                const int32_t* draw_idx, int n_idx, const uint8_t* excluded_mask, const int32_t* rules, float* out) {
  (void)n_nodes;
  for (int d = 0; d < n_idx; ++d) {
    const bk_node* f = forests + (size_t)draw_idx[d] * (size_t)n_trees * BK_MAX_NODES;
    for (int i = 0; i < n; ++i) {
      double acc = 0.0;
      for (int t = 0; t < n_trees; ++t)
        acc = BK_DADD(acc, predict_tree(f + (size_t)t * BK_MAX_NODES, X + (size_t)i * (size_t)n_cols, excluded_mask, rules));
      out[(size_t)d * (size_t)n + i] = (float)acc;
    }
  }
  return BK_OK;
}
