"""An INDEPENDENT statement of the PGBART step in plain Python / numpy float64.  TEST INFRASTRUCTURE ONLY.

Why it exists: the C oracle (pgbart_oracle.c) and the CUDA kernels share include/bk_spec.h — the fixed-point sums, the
polynomial exp / log / cos and the integer resampling are written once and used by both, so a mistake in that header is
invisible to the GPU-vs-oracle parity tests ("what's weak" 3).  This module restates SURVEY.md Appendix A
a second time with NONE of that header: leaf means, log-likelihoods and particle weights in float64 with numpy's own
exp / log / cos, the textbook systematic resampling (`w = exp(lw - max) + 1e-12`, normalise, walk the running sum), no
fixed point anywhere, its own Philox4x32-10.  Only the random-number ADDRESSING (key = (seed, chain), counter = (draw,
group<<16 | tree, round<<16 | particle, purpose), App. A.9) and the depth-prior table are taken as given.

What it pins (tests/test_independent_model.py): for the same seed the C oracle makes the SAME decisions — popped nodes,
split variables, split values, child sizes, resampling ancestors, the selected particle, variable-inclusion counts, leaf
ids — and its leaf values, log-weights and running leaf sd agree within 1e-5 (north_star's stated tolerance), i.e. the
fixed-point / polynomial arithmetic of bk_spec.h changes no decision and no number beyond rounding.

Pure Python loops: small cases only (N of a few hundred).  Normal and Bernoulli-logit likelihoods, the two shared-tree
multi-output families (K values per leaf: heteroscedastic Normal, Categorical-softmax; tests/test_bart.py:107-123,140-164),
Continuous / OneHot / Subset rules, missing covariates (App. A.4: up to four candidate members per split-value draw, rows without the split
covariate leave the tree).  Reference anchors as in pgbart_oracle.c: inputs pymc_bart/bart.py:141-158, depth prior :107-109, initial
value :148, split rules tests/test_bart.py:143-145 and docs/api_reference.rst:16.
"""
from __future__ import annotations

import math

import numpy as np

U_LEAF, U_VAR, U_VAL, Z_LEFT, Z_RIGHT, U_RESAMPLE, U_FINAL, U_PICK = range(8)
RULE_CONTINUOUS, RULE_ONEHOT, RULE_SUBSET = 0, 1, 2
M32 = 0xFFFFFFFF


def philox4x32_10(k0, k1, c0, c1, c2, c3):
    """Philox4x32-10 (Salmon et al. 2011) on Python integers."""
    for _ in range(10):
        p0 = 0xD2511F53 * c0
        p1 = 0xCD9E8D57 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & M32, p1 & M32, ((p0 >> 32) ^ c3 ^ k1) & M32, p0 & M32
        k0 = (k0 + 0x9E3779B9) & M32
        k1 = (k1 + 0xBB67AE85) & M32
    return c0, c1, c2, c3


class _Node:
    __slots__ = ("var", "split", "left", "depth", "value", "n")

    def __init__(self, depth, value, n):     # value: one float32 per output (K of them; K = 1 unless the trees are shared)
        self.var, self.split, self.left, self.depth, self.value, self.n = -1, 0.0, -1, depth, np.array(value, dtype=np.float32), int(n)

    def copy(self):
        c = _Node(self.depth, self.value, self.n)
        c.var, c.split, c.left = self.var, self.split, self.left
        return c


class _Particle:
    def __init__(self, nodes, ids, q_head, lw):
        self.nodes, self.ids, self.q_head, self.lw = nodes, ids, q_head, lw

    def copy(self):
        return _Particle([n.copy() for n in self.nodes], self.ids.copy(), self.q_head, self.lw)

    def predict(self):
        """(K, N) in-sample prediction of the tree; rows that left the tree (id 255) predict 0."""
        K = self.nodes[0].value.size
        vals = np.zeros((256, K), dtype=np.float32)
        for k, n in enumerate(self.nodes):
            vals[k] = n.value
        return vals[self.ids].T


class FloatModelChain:
    def __init__(self, X, y, m, num_particles, p_leaf, seed=0, chain=0, batch=(0.1, 0.1), split_prior=None, split_rules=None,
                 likelihood="normal", n_outputs=1, group=0):
        assert likelihood in ("normal", "bernoulli", "normal_hetero", "categorical")
        self.lik, self.K = likelihood, int(n_outputs)
        self.group = int(group)      # output group of a BART(shape=(k, n), separate_trees=True) variable: its own forest and Philox word
        self.X = np.asarray(X, dtype=np.float32)                  # (N, p)
        self.y = np.asarray(y, dtype=np.float32)
        self.N, self.p = self.X.shape
        self.m, self.P = int(m), int(num_particles)
        self.p_leaf = np.asarray(p_leaf, dtype=np.float64)
        self.seed, self.chain = int(seed) & M32, int(chain)
        self.bt, self.bp = max(1, int(m * batch[0])), max(1, int(m * batch[1]))
        self.alpha_vec = np.ones(self.p) if split_prior is None else np.asarray(split_prior, dtype=np.float64).copy()
        self.rules = np.zeros(self.p, dtype=int) if split_rules is None else np.asarray(split_rules, dtype=int)
        self._rebuild_cum()
        ymean = float(np.asarray(y, dtype=np.float64).mean())
        self.init_leaf = np.full(self.K, np.float32(ymean / m), dtype=np.float32)
        uniq = np.unique(np.asarray(y, dtype=np.float64))
        sd0 = 3.0 / math.sqrt(m) if (uniq.size == 2 and set(uniq.tolist()) == {0.0, 1.0}) else float(np.asarray(y, dtype=np.float64).std()) / math.sqrt(m)
        self.leaf_sd = np.full(self.K, float(np.float32(sd0)))
        self.st = np.full((self.K, self.N), np.float32(ymean), dtype=np.float32)
        self.forest = [_Particle([_Node(0, self.init_leaf, self.N)], np.zeros(self.N, dtype=np.uint8), 1, 0.0) for _ in range(m)]
        self.wf_mean = np.zeros((self.K, self.N), dtype=np.float32)
        self.wf_m2 = np.zeros((self.K, self.N), dtype=np.float32)
        self.wf_count = self.iter = self.lower = self.draw = 0
        self.trace = []

    # ---- random numbers: addressing as App. A.9
    def _rng(self, tree, rnd, particle, purpose, group=None):
        group = self.group if group is None else group
        return philox4x32_10(self.seed, self.chain, self.draw, ((group << 16) | (tree & 0xFFFF)) & M32,
                             ((rnd << 16) | (particle & 0xFFFF)) & M32, purpose)

    def _normal(self, tree, rnd, particle, purpose, group=None):
        w = self._rng(tree, rnd, particle, purpose, group)
        u1 = (w[0] + 1.0) / 4294967296.0
        return math.sqrt(-2.0 * math.log(u1)) * math.cos(2.0 * math.pi * (w[1] / 4294967296.0))

    def _rebuild_cum(self):
        run, cum = 0.0, []
        tot = 0.0
        for v in range(self.p):
            tot += self.alpha_vec[v]
        for v in range(self.p):
            run += self.alpha_vec[v]
            cum.append(run / tot)
        self.cum = cum

    def _loglik(self, part, r, sigma):
        """Full-model data log-likelihood with this particle's tree in place (App. A.6)."""
        y = self.y.astype(np.float64)
        if self.lik == "normal":
            d = r.astype(np.float64) - part.predict()[0].astype(np.float64)
            return -0.5 * float(np.dot(d, d)) / (sigma * sigma) - self.N * (math.log(sigma) + 0.5 * math.log(2.0 * math.pi))
        f = (self._noi + part.predict()).astype(np.float64)      # linear predictors: (sum of trees without the tree) + this tree
        if self.lik == "bernoulli":            # sum_i y_i f_i - softplus(f_i)
            term = y * f[0] - np.logaddexp(0.0, f[0])
        elif self.lik == "normal_hetero":      # y ~ Normal(f0, |f1|)
            a = np.maximum(np.abs(f[1]), 1e-20)
            with np.errstate(over="ignore"):
                term = -0.5 * ((y - f[0]) / a) ** 2 - np.log(a) - 0.5 * math.log(2.0 * math.pi)
        else:                                  # y ~ Categorical(softmax(f, axis=0))
            mx = f.max(axis=0)
            term = f[self.y.astype(int), np.arange(self.N)] - (mx + np.log(np.exp(f - mx).sum(axis=0)))
        # the one property of the spec taken over here: a row's term saturates at +-512 (the range of the 2^-20 fixed-point
        # terms; a row more than 32 scales from its mean under the heteroscedastic Normal) — a DEFINED divergence, DESIGN.md 2
        return float(np.sum(np.clip(term, -512.0, 512.0)))

    def _leaf_value(self, rows, tree, rnd, q, purpose):
        """mean(sum of trees over the members)/m + z * leaf_sd per output; the normal of output j comes from the Philox block
        whose group word is j; an empty leaf carries 0."""
        out = np.zeros(self.K, dtype=np.float32)
        if rows.size == 0:
            return out
        for j in range(self.K):
            z = self._normal(tree, rnd, q, purpose, group=j if self.K > 1 else None)
            mean = float(self.st[j, rows].astype(np.float64).sum()) / self.m / rows.size
            out[j] = np.float32(mean + z * self.leaf_sd[j])
        return out

    def _grow(self, part, tree, rnd, q, r, sigma, rec):
        if part.q_head >= len(part.nodes):
            return False
        j = part.q_head
        part.q_head += 1
        rec["node"] = j
        nd = part.nodes[j]
        pl = self.p_leaf[nd.depth] if nd.depth < 256 else 1.0
        u1 = self._rng(tree, rnd, q, U_LEAF)[0] / 4294967296.0
        if not (u1 > pl) or len(part.nodes) + 2 > 255:
            return False
        u2 = self._rng(tree, rnd, q, U_VAR)[0] / 4294967296.0
        v = self.p - 1
        for i, c in enumerate(self.cum):
            if u2 < c:
                v = i
                break
        if nd.n < 2:
            return False
        members = np.nonzero(part.ids == j)[0]
        w = self._rng(tree, rnd, q, U_VAL)
        xc = self.X[members, v]
        missing = np.isnan(xc)
        if self.rules[v] == RULE_SUBSET:
            cats = sorted({int(c) for c in xc[~missing].tolist()})
            if len(cats) < 2:
                return False
            cand = cats[:-1]
            pick = ((w[0] * ((1 << len(cand)) - 1)) >> 32) + 1
            chosen = [c for i, c in enumerate(cand) if (pick >> i) & 1]
            s = np.float32(sum(1 << c for c in chosen))
            left = np.isin(np.where(missing, -1, xc).astype(int), chosen)
        else:
            s = None
            for t in range(4):          # a candidate member whose covariate is missing is skipped; four misses: no split
                cand = xc[(w[t] * nd.n) >> 32]
                if not np.isnan(cand):
                    s = cand
                    break
            if s is None:
                return False
            with np.errstate(invalid="ignore"):
                left = (xc == s) if self.rules[v] == RULE_ONEHOT else (xc <= s)
        L = len(part.nodes)
        rows_l, rows_r = members[left & ~missing], members[~left & ~missing]
        part.ids[rows_l] = L
        part.ids[rows_r] = L + 1
        part.ids[members[missing]] = 255       # no value in the split covariate: the row leaves the tree and predicts 0
        vl = self._leaf_value(rows_l, tree, rnd, q, Z_LEFT)
        vr = self._leaf_value(rows_r, tree, rnd, q, Z_RIGHT)
        nd.var, nd.split, nd.left = v, s, L
        part.nodes.append(_Node(nd.depth + 1, vl, rows_l.size))
        part.nodes.append(_Node(nd.depth + 1, vr, rows_r.size))
        part.lw = self._loglik(part, r, sigma)
        rec.update(var=v, split=float(s), n_left=int(rows_l.size), n_right=int(rows_r.size), val_left=float(vl[0]), val_right=float(vr[0]))
        return True

    @staticmethod
    def _systematic(lws, u):
        lws = np.asarray(lws, dtype=np.float64)
        w = np.exp(lws - lws.max()) + 1e-12
        cum = np.cumsum(w / w.sum())
        L, idx, out = len(lws), 0, []
        for i in range(L):
            point = (u + i) / L
            while point > cum[idx] and idx < L - 1:
                idx += 1
            out.append(idx)
        return out

    def step(self, tune, sigma=1.0):
        self.trace = []
        vi = np.zeros(self.p, dtype=np.int64)
        T = self.bt if tune else self.bp
        upper = min(self.lower + T, self.m)
        grow_events = 0
        for t in range(self.lower, upper):
            self.iter += 1
            old = self.forest[t]
            noi = (self.st - old.predict()).astype(np.float32)
            self._noi = noi
            r = (self.y - noi[0]).astype(np.float32)
            parts = [old.copy()]
            parts[0].q_head = len(parts[0].nodes)
            parts[0].lw = self._loglik(parts[0], r, sigma)
            for q in range(1, self.P):
                pq = _Particle([_Node(0, self.init_leaf, self.N)], np.zeros(self.N, dtype=np.uint8), 0, 0.0)
                pq.lw = self._loglik(pq, r, sigma)
                parts.append(pq)
            rnd = 0
            while True:
                recs = []
                for q in range(1, self.P):
                    rec = dict(kind=1, tree=t, round=rnd, particle=q, node=-1, var=-1, n_left=0, n_right=0, split=0.0, val_left=0.0,
                               val_right=0.0, ancestor=-1)
                    if self._grow(parts[q], t, rnd, q, r, sigma, rec):
                        grow_events += 1
                    rec["log_w"] = parts[q].lw
                    recs.append(rec)
                self.trace.extend(recs)
                if not any(pq.q_head < len(pq.nodes) for pq in parts[1:]):
                    break
                u = self._rng(t, rnd, 0, U_RESAMPLE)[0] / 4294967296.0
                anc = self._systematic([pq.lw for pq in parts[1:]], u)
                parts = [parts[0]] + [parts[a + 1].copy() for a in anc]
                for rec, a in zip(recs, anc):
                    rec["ancestor"] = a + 1
                rnd += 1
            uf = self._rng(t, 0xFFFF, 0, U_FINAL)[0] / 4294967296.0
            anc = self._systematic([pq.lw for pq in parts], uf)
            pick = (self._rng(t, 0xFFFF, 0, U_PICK)[0] * self.P) >> 32
            win = anc[pick]
            new = parts[win]
            newp = new.predict()
            self.st = (noi + newp).astype(np.float32)
            if tune:
                self.wf_count += 1
                cnt = np.float32(self.wf_count)
                delta = newp - self.wf_mean
                self.wf_mean = (self.wf_mean + delta / cnt).astype(np.float32)
                self.wf_m2 = (self.wf_m2 + delta * (newp - self.wf_mean)).astype(np.float32)
                if self.iter > self.m:
                    self._rebuild_cum()
                for nd in new.nodes:
                    if nd.var >= 0:
                        self.alpha_vec[nd.var] += 1.0
                if self.iter > 2:
                    self.leaf_sd = np.sqrt(self.wf_m2.astype(np.float64) / self.wf_count).mean(axis=1)
            else:
                for nd in new.nodes:
                    if nd.var >= 0:
                        vi[nd.var] += 1
            self.trace.append(dict(kind=2, tree=t, round=rnd, particle=win, node=len(new.nodes), var=-1, ancestor=int(pick),
                                   log_w=new.lw, aux=float(self.leaf_sd[self.K - 1])))
            kept = new.copy()
            kept.q_head = len(kept.nodes)
            self.forest[t] = kept
        self.lower = upper if upper < self.m else 0
        self.draw += 1
        return vi, grow_events


def predict_tree_float(nodes, x, excluded=None, rules=None):
    """Out-of-sample value of one tree at the covariate row x (App. A.10), recursively in float64: at a split on an
    excluded variable both children count, weighted by their shares of the training rows; otherwise the split rule
    decides (a missing covariate compares false and goes right).  nodes: NODE_DTYPE records of the tree."""
    def walk(k):
        nd = nodes[k]
        v = int(nd["var"])
        if v < 0:
            return float(nd["value"])
        l = int(nd["left"])
        if excluded is not None and excluded[v]:
            tot = float(nodes[l]["n"]) + float(nodes[l + 1]["n"])
            if not tot > 0.0:
                return 0.0
            wl = float(nodes[l]["n"]) / tot
            return wl * walk(l) + (1.0 - wl) * walk(l + 1)
        xv, s = float(x[v]), float(nd["split"])
        rule = RULE_CONTINUOUS if rules is None else int(rules[v])
        if rule == RULE_SUBSET:
            left = (not math.isnan(xv)) and xv == int(xv) and 0 <= int(xv) < 24 and bool((int(s) >> int(xv)) & 1)
        elif rule == RULE_ONEHOT:
            left = xv == s
        else:
            left = xv <= s
        return walk(l if left else l + 1)

    return walk(0)
