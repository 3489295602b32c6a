"""PGBART step method: host side of the B200 sampler (reference: ``bartrs.PGBART``,
constructed as ``PGBART([rv], num_particles=...)`` at tests/test_bart.py:231-232 and
driven by ``pm.sample`` one ``step(point)`` per draw).

State lives on the GPU (pymc_bart_b200.core.DeviceSampler); this class keeps the step protocol:

* ``astep`` returns the new value of the BART variable (the sum of trees) and the per-draw stats
  ``{"variable_inclusion": <base64 varint>, "tune": bool}`` (pymc_bart/utils.py:1387-1398, consumers :778-790);
* ``stop_tuning()`` ends adaptation.  From the first post-tuning draw on, the chain has ONE entry
  ``(baseline_forest, batches)`` in ``op.all_trees`` (pymc_bart/utils.py:117,124-127) and every draw appends the trees it
  rewrote to that entry's ``batches`` — a nested ``Manager().list()`` proxy when ``op.all_trees`` is one
  (pymc_bart/bart.py:133-146), so a worker process ships ~T trees per draw and the parent can predict at any time;
  ``op.n_outputs`` is set on the op class (utils.py:125);
* the object pickles WITHOUT device state (PyMC pickles the step into one worker process per chain when
  ``cores > 1``): the CUDA context, the device buffers and the native handle are created on first use in whichever
  process runs the chain, on ``device = chain % n_gpus`` unless a device is given;
* when ``tune`` goes back to True after post-tuning draws (PyMC re-using one step object for the next chain,
  ``cores=1``), the sampler starts a fresh chain with the next chain index.

Several BART variables in one model (tests/test_bart.py:167-241, ``pm.Normal("y", mu1 + mu2, sigma, observed=Y)``): one step
object per variable, as in the reference; ``observed=`` is the likelihood's data and ``offset_names=`` the point entries
that make up the rest of the location (the other BART variables), so that each step weighs its particles with
``Normal(observed - offset | value, sigma)`` at the current point — what the reference's compiled ``datalogp`` does with
the other variables as shared inputs.  Normal likelihood only.

Extensions: ``chains=C`` batches C independent chains in one launch (the reference runs one step object per chain).
``lookahead=n`` (posterior phase only, and only when the likelihood parameters are FIXED — no ``sigma_name``, nobody
assigns ``step.sigma`` between draws): ``astep`` is served from launches of ``n`` steps (``bk_run_launch``); the next launch
runs on the GPU while the caller consumes the draws of the previous one, each draw still arrives as a host array with its
stats and its batch of rewritten trees.  With another step method updating the scale between draws (the reference's
``sigma ~ HalfNormal`` models) keep ``lookahead=1``: a step needs the scale of the current point.  A step method is not
told how long tuning lasts (PyMC calls ``stop_tuning()`` when it is over), so tuning steps are one launch per call unless
the caller says it: with ``tune_draws=k`` the first k calls (made with ``tune=True``) are served ahead as well, and the
launches change to posterior steps exactly there; a call whose ``tune`` flag disagrees with that count raises.
"""
from __future__ import annotations

import numpy as np

from . import _cabi
from .core import DeviceSampler
from .settings import encode_subset_columns, make_settings, subset_category_tables
from .utils import _encode_vi

# Closed likelihood families of the device path.  The two multi-output ones are the models the reference's own tests
# build on shared trees: y ~ Normal(w[0], |w[1]|) with w = BART(shape=(2, n)) (tests/test_bart.py:107-123) and
# y ~ Categorical(softmax(mu, axis=0)) with mu = BART(shape=(k, n)) (tests/test_bart.py:140-164).
LIKELIHOODS = {"normal": _cabi.BK_LIK_NORMAL, "bernoulli": _cabi.BK_LIK_BERNOULLI_LOGIT,
               "normal_hetero": _cabi.BK_LIK_NORMAL_HETERO, "categorical": _cabi.BK_LIK_CATEGORICAL}
SHARED_TREE_LIKELIHOODS = ("normal_hetero", "categorical")


class PGBART:
    name = "pgbart"
    default_blocked = False
    generates_stats = True
    stats_dtypes_shapes = {"variable_inclusion": (object, []), "tune": (bool, [])}

    def __init__(self, vars=None, num_particles=10, batch=(0.1, 0.1), model=None, *, likelihood="normal", sigma=1.0,
                 chains=1, chain_base=0, seed=0, device=None, depth_offset=0, store_history=True, trace_capacity=0,
                 sigma_name=None, sigma_transform=None, lookahead=1, tune_draws=None, observed=None, offset_names=None,
                 **kwargs):
        if vars is None or len(vars) != 1:
            raise ValueError("PGBART takes exactly one BART variable: PGBART([rv], num_particles=...)")
        rv = vars[0]
        op = rv.owner.op if hasattr(rv, "owner") else rv
        if getattr(op, "name", None) != "BART" or not hasattr(op, "all_trees"):
            raise TypeError("PGBART can only sample BART variables")
        if op.response != "constant":
            raise NotImplementedError(f"response={op.response!r} has no device implementation (constant leaves only)")
        if likelihood not in LIKELIHOODS:
            raise NotImplementedError(f"likelihood {likelihood!r} has no device implementation")
        # multi-output, BART(shape=(k, n)):
        #  * separate_trees=True -> k output groups, each with its own forest and its own response row (independent
        #    likelihood per output; BASELINE.json config 4);
        #  * otherwise SHARED trees (the reference's mode at this commit, tests/test_bart.py:107-123,140-164): one
        #    forest, k values per leaf, the weight is the likelihood of the whole (k, n) value -> needs one of the
        #    multi-output families
        shape = tuple(getattr(rv, "shape", (np.asarray(op.X).shape[0],)))
        k_out = int(shape[0]) if len(shape) == 2 else 1
        separate = bool(getattr(op, "separate_trees", False))
        groups = k_out if separate else 1
        outputs = 1 if separate else k_out
        if outputs > 1 and likelihood not in SHARED_TREE_LIKELIHOODS:
            raise NotImplementedError("shared-tree multi-output BART needs likelihood='normal_hetero' (shape=(2, n)) or "
                                      "'categorical' (shape=(k, n)); pass separate_trees=True for independent outputs")
        if outputs == 1 and likelihood in SHARED_TREE_LIKELIHOODS:
            raise ValueError(f"likelihood={likelihood!r} needs a multi-output BART variable with shared trees (shape=(k, n))")
        Yarr = np.asarray(op.Y, dtype=np.float64)
        if groups > 1 and Yarr.ndim == 1:
            Yarr = np.broadcast_to(Yarr, (groups, Yarr.shape[0]))
        # Several BART variables in one model (tests/test_bart.py:167-241: pm.Normal("y", mu1 + mu2, sigma, observed=Y)): the
        # likelihood of this variable's step is Normal(observed - offset | value, sigma), offset = the other terms of the
        # location at the current point.  `observed` is the likelihood's data (default: the Y handed to BART, which the
        # reference uses for the initial value and the leaf scale only); `offset_names` are the point entries added up into
        # the offset by step(point); set_offset() does the same for a caller that drives astep() itself.
        self.offset_names = list(offset_names) if offset_names else None
        self._observed = None if observed is None else np.asarray(observed, dtype=np.float64)
        self._offset_dirty = False
        if self._observed is not None or self.offset_names:
            if likelihood != "normal":
                raise NotImplementedError("observed= / offset_names= (this variable as one term of the location) are implemented for "
                                          "likelihood='normal' only")
            if self._observed is None:
                self._observed = np.asarray(op.Y, dtype=np.float64)
            if self._observed.shape != Yarr.shape:
                self._observed = np.broadcast_to(self._observed, Yarr.shape)
            if lookahead and int(lookahead) > 1 and self.offset_names:
                raise ValueError("offset_names needs lookahead=1: the offset changes with every point")
        self.groups = groups
        self.outputs = outputs
        self.op = op
        self.vars = [rv]
        self.num_particles = int(num_particles)
        self.batch = tuple(batch)
        self.chains = int(chains)
        self.chain_base = int(chain_base)
        self.seed = seed
        self.device = device
        self.tune = True
        self.sigma = sigma
        self.sigma_name = sigma_name
        self.sigma_transform = sigma_transform   # e.g. np.exp when sigma_name is the log-transformed value variable
        self.store_history = bool(store_history)
        self.lookahead = int(lookahead)
        self.tune_draws = None if tune_draws is None else int(tune_draws)
        self._Y = Yarr
        self._settings_kw = dict(
            m=op.m, alpha=op.alpha, beta=op.beta, split_prior=op.split_prior, split_rules=op.split_rules,
            num_particles=num_particles, batch=batch, n_chains=chains, seed=seed, likelihood=LIKELIHOODS[likelihood],
            depth_offset=depth_offset, trace_capacity=trace_capacity, n_groups=groups, n_outputs=outputs,
            value_range=0.0 if self._observed is None else float(np.abs(self._observed).max()))
        self.settings = make_settings(op.X, Yarr, chain_base=self.chain_base, device=0 if device is None else device,
                                      **self._settings_kw)
        # SubsetSplit columns (docs/api_reference.rst:16): the device works on category codes; the tables that map a
        # column's values to codes stay with the op, so that prediction on new data encodes it the same way
        self.subset_tables = subset_category_tables(np.asarray(op.X), self.settings.split_rules)
        (op if isinstance(op, type) else type(op)).subset_tables = self.subset_tables
        self.n_rows, self.n_cols, self.m = self.settings.n_rows, self.settings.n_cols, self.settings.n_trees
        self.core = None                # device state: created lazily, never pickled
        self._offset = None
        self._pub_thread = None
        self._pub_queue = None
        self._reset_chain_state()
        self.last_stats = None
        self.history_bytes_per_step = 0
        # read back through the CLASS by BARTRV.rng_fn -> _get_posterior_sampler(cls) (bart.py:65, utils.py:125)
        (op if isinstance(op, type) else type(op)).n_outputs = k_out

    # ---- device state ---------------------------------------------------------
    def _reset_chain_state(self):
        self._lower = 0
        self._post_draws = 0
        self._batches = None        # per chain: the `batches` list published in op.all_trees
        self._served = []           # lookahead: draws of the last collected launch not handed out yet
        self._inflight = []         # lookahead: the launches queued on the GPU, oldest first (at most two)
        self._ring = None
        self._tune_launched = 0     # tuning steps launched ahead (tune_draws given)

    def _pick_device(self):
        if self.device is not None:
            return int(self.device)
        import torch

        n = torch.cuda.device_count() if torch.cuda.is_available() else 0
        return (self.chain_base // max(1, self.chains)) % n if n else 0

    def _ensure_core(self):
        if self.core is None:
            dev = self._pick_device()
            self.settings = make_settings(self.op.X, self._Y, chain_base=self.chain_base, device=dev, **self._settings_kw)
            self.core = DeviceSampler(self.settings, encode_subset_columns(np.asarray(self.op.X), self.subset_tables), self._Y)
            if self._observed is not None:
                self.set_offset(self._offset if getattr(self, "_offset", None) is not None else 0.0)
            self.core.enable_host_output(True)   # astep returns a host array every draw (the trace stores it)
            ahead = self.core.MAX_STEPS_PER_LAUNCH if self._serves_ahead(True) else (
                self._steps_ahead(self.core) if self._serves_ahead(False) else 1)
            ahead = min(ahead, max(1, self.lookahead))
            if self.store_history:
                self.core.enable_history(True, steps_per_launch=ahead)
                self.history_bytes_per_step = getattr(self.core, "history_bytes_per_step", 0)
            if ahead > 1:
                self._make_ring(self.core, ahead)     # (pinning host memory takes ~1 ms per MB: part of the set-up, not of a draw)
        return self.core

    def set_offset(self, offset):
        """The other terms of the likelihood's location at the current point (array [n] / [groups][n] or a number): the next
        steps see the response `observed - offset`."""
        if self._observed is None:
            raise RuntimeError("set_offset needs observed= (or offset_names=) at construction")
        self._offset = np.broadcast_to(np.asarray(offset, dtype=np.float64), self._observed.shape)
        if self.core is not None:
            if self._served or self._inflight:
                raise RuntimeError("the offset cannot change while draws computed ahead are pending (use lookahead=1)")
            self.core.set_response(self._observed - self._offset)

    def prepare(self):
        """Create the device state now (X/Y upload, workspace, pinned buffers) instead of at the first astep()."""
        self._ensure_core()
        return self

    def _serves_ahead(self, tune):
        """Are calls of this phase served from launches of several steps?"""
        if self.lookahead <= 1 or self.sigma_name is not None:
            return False
        return (self.tune_draws is not None and self.tune_draws > 0) if tune else True

    def _steps_ahead(self, core):
        """Steps of a posterior launch: the history needs every tree rewritten at most once per launch."""
        return max(1, min(self.lookahead, core.MAX_STEPS_PER_LAUNCH, self.m // max(1, self.settings.batch_post)))

    def _make_ring(self, core, n):
        torch = core.torch
        with torch.cuda.device(core.device):
            self._ring = {"n": n, "slot": 0,
                          # three slots: two launches queued on the GPU + the one whose draws are being handed out
                          "dev": [torch.empty((n, core.rows, core.ld), dtype=torch.float32, device=core.device) for _ in range(3)],
                          "host": [torch.empty((n, core.rows, core.ld), dtype=torch.float32).pin_memory() for _ in range(3)],
                          "ev": [torch.cuda.Event() for _ in range(3)], "kernel_done": [torch.cuda.Event() for _ in range(3)],
                          "copy_stream": torch.cuda.Stream(device=core.device)}

    def __getstate__(self):
        state = dict(self.__dict__)
        state["core"] = None            # no CUDA context / native handle crosses a process boundary
        state["_batches"] = None
        state["last_stats"] = None
        state["_pub_thread"] = None
        state["_pub_queue"] = None
        state["_served"] = []
        state["_inflight"] = []
        state["_ring"] = None
        return state

    def __setstate__(self, state):
        self.__dict__.update(state)
        self.__dict__.setdefault("_offset", None)

    def next_chain(self):
        """Start a fresh chain on this step object (PyMC with cores=1 runs the chains one after the other)."""
        self.close()
        self.chain_base += self.chains
        self._reset_chain_state()
        self.tune = True

    # ---- step-method protocol -------------------------------------------------
    @staticmethod
    def competence(var, has_grad=False):
        op = getattr(getattr(var, "owner", None), "op", None)
        return 3 if getattr(op, "name", None) == "BART" and hasattr(op, "all_trees") else 0  # 3 == Competence.IDEAL

    def stop_tuning(self):
        self.tune = False

    # ---- history publication ---------------------------------------------------
    # Appending to a Manager proxy is an IPC round trip (~0.1 ms per chain and draw): a writer thread does it, so the
    # step does not wait for the parent process.  Appends keep their order (one queue, one thread); flush_history()
    # waits until everything queued so far is in op.all_trees (close() and sample() call it).
    def _publish(self, target, item):
        from multiprocessing.managers import BaseProxy

        if not isinstance(target, BaseProxy):
            target.append(item)
            return
        if self._pub_thread is None:
            import queue
            import threading

            self._pub_queue = queue.Queue()

            def drain(q=self._pub_queue):
                while True:
                    job = q.get()
                    try:
                        if job is not None:
                            job[0].append(job[1])
                    finally:
                        q.task_done()
                    if job is None:
                        return

            self._pub_thread = threading.Thread(target=drain, name="pgbart-history", daemon=False)
            self._pub_thread.start()
        self._pub_queue.put((target, item))

    def flush_history(self):
        """Block until every batch queued so far has reached op.all_trees."""
        if self._pub_thread is not None:
            self._pub_queue.join()

    def _stop_publisher(self):
        if self._pub_thread is not None:
            self._pub_queue.put(None)
            self._pub_thread.join()
            self._pub_thread = None
            self._pub_queue = None

    def _new_batches(self):
        """A list of the same kind as op.all_trees: a nested Manager proxy when the history crosses processes."""
        from .bart import sibling_list

        return sibling_list(self.op.all_trees)

    def _chain_slice(self, per_vc, c):
        """(chain c's output groups) of a per-(chain, group) list of (nodes, n_nodes[, leaf values]) -> one entry."""
        parts = per_vc[c * self.groups:(c + 1) * self.groups]
        return tuple(np.concatenate([p[i] for p in parts]) for i in range(len(parts[0])))

    # ---- lookahead: posterior draws served from launches of several steps -----------------------------------------------
    def _publish_first_entry(self, core):
        """First posterior draw: publish the chain's entry (baseline = the forest tuning ended with); every later draw
        appends its rewritten trees to the entry's `batches`."""
        base = core.baseline()
        self._batches = [self._new_batches() for _ in range(self.chains)]
        for c in range(self.chains):
            self._publish(self.op.all_trees, (self._chain_slice(base, c), self._batches[c]))

    def _publish_batch(self, batch):
        first, nn, *arrays = batch      # nodes[, leaf values of every output]
        off = np.concatenate([[0], np.cumsum(nn.sum(axis=1))])
        G = self.groups
        for c in range(self.chains):
            self._publish(self._batches[c], (first, nn[c * G:(c + 1) * G].copy(), *(a[off[c * G]: off[(c + 1) * G]].copy() for a in arrays)))

    def _launch_ahead(self, core):
        torch = core.torch
        left = 0 if self.tune_draws is None else self.tune_draws - self._tune_launched
        tune = left > 0                     # (without tune_draws only posterior steps come here)
        n = min(left, core.MAX_STEPS_PER_LAUNCH, max(1, self.lookahead)) if tune else self._steps_ahead(core)
        if self._ring is None or self._ring["n"] < n:
            self._make_ring(core, n)
        if not tune and self.store_history and self._batches is None:
            self._publish_first_entry(core)   # the forest tuning ended with (nothing is in flight at this point)
        r = self._ring
        k = r["slot"]
        core.run_launch(n, tune, self.sigma, draws_out=r["dev"][k][:n])
        # the draws follow the kernel to pinned host memory on a second stream: the next launch does not wait for 4*n*rows*N
        # bytes to cross PCIe (the ring slot is not written again before this copy has been waited for)
        r["kernel_done"][k].record(core.stream())
        with torch.cuda.stream(r["copy_stream"]):
            r["copy_stream"].wait_event(r["kernel_done"][k])
            r["host"][k][:n].copy_(r["dev"][k][:n], non_blocking=True)
            r["ev"][k].record()
        if tune:
            self._tune_launched += n
        self._inflight.append({"slot": k, "n": n, "tune": tune, "sigma": np.array(self.sigma, dtype=np.float64, copy=True)})
        r["slot"] = (k + 1) % 3

    def _collect_ahead(self, core):
        fl = self._inflight.pop(0)
        vi, stats = core.run_wait()
        self._ring["ev"][fl["slot"]].synchronize()
        host = self._ring["host"][fl["slot"]].numpy()[:, :, : self.n_rows]
        keep = self.store_history and not fl["tune"]
        hist = [core.history_batch(k) for k in range(fl["n"])] if keep else [None] * fl["n"]
        self._served = [(host[k], vi[k], stats[k], hist[k], fl["sigma"], fl["tune"]) for k in range(fl["n"])]

    def _astep_ahead(self, core, tune):
        if not self._served:
            while len(self._inflight) < 2:    # two launches queued: the GPU goes from one to the next without the host
                self._launch_ahead(core)
            self._collect_ahead(core)         # (waits for the older one)
            self._launch_ahead(core)          # queued behind the one that is running now
        value, vi, stats, hist, sigma, was_tune = self._served.pop(0)
        if was_tune != tune:
            raise RuntimeError(f"tune_draws={self.tune_draws} does not match the calls: a step computed with tune={was_tune} "
                               f"was asked for with tune={tune} (stop_tuning() must come after exactly tune_draws calls)")
        if not np.array_equal(np.asarray(self.sigma, dtype=np.float64), sigma):
            raise RuntimeError("lookahead > 1 needs fixed likelihood parameters: step.sigma changed while draws computed with the "
                               "old value were waiting (use lookahead=1 when another step method updates the scale)")
        self.last_stats = stats
        if hist is not None:
            self._publish_batch(hist)
        return value, vi

    def astep(self, _q=None):
        tune = bool(self.tune)
        if tune and self._post_draws > 0:     # tuning again after posterior draws: the next chain starts
            self.next_chain()
        core = self._ensure_core()
        if self._serves_ahead(tune):
            value, vi = self._astep_ahead(core, tune)
            self._post_draws += 0 if tune else 1
            return self._pack(value, vi, tune)
        if self._served or self._inflight:
            raise RuntimeError("draws computed ahead are pending: lookahead cannot be switched off in the middle of a chain")
        T = self.settings.batch_tune if tune else self.settings.batch_post
        lo = self._lower
        hi = min(lo + T, self.m)
        if not tune and self.store_history and self._batches is None:
            self._publish_first_entry(core)
        vi, stats = core.step(tune, self.sigma)
        self.last_stats = stats
        self._lower = hi if hi < self.m else 0
        value = core.sum_trees_host()
        if not tune:
            self._post_draws += 1
            if self.store_history:
                self._publish_batch(core.history_batch())
        return self._pack(value, vi, tune)

    def _pack(self, value, vi, tune):
        """(value, stats) in the shapes the step protocol hands back."""
        vic = vi.reshape(self.chains, self.groups, -1).sum(axis=1)       # one inclusion vector per BART variable
        out_stats = [{"variable_inclusion": _encode_vi(vic[c].tolist()), "tune": tune} for c in range(self.chains)]
        value = value.reshape(self.chains, self.groups * self.outputs, -1)    # (chain, output, row)
        if self.groups * self.outputs == 1:
            value = value[:, 0]
        if self.chains == 1:
            return value[0].copy(), [out_stats[0]]
        return value.copy(), out_stats

    def step(self, point):
        """PyMC-style: reads the likelihood scale from the point when ``sigma_name`` is set."""
        if self.sigma_name is not None:
            if self.sigma_name not in point:   # a silently kept default would sample the wrong posterior
                raise KeyError(f"sigma_name={self.sigma_name!r} is not in the point (keys: {sorted(point)}); PyMC points are keyed "
                               "by value-variable names (e.g. 'sigma_log__'): pass sigma_transform= to map it back")
            v = point[self.sigma_name]
            self.sigma = self.sigma_transform(v) if self.sigma_transform is not None else v
        if self.offset_names:
            missing = [n for n in self.offset_names if n not in point]
            if missing:
                raise KeyError(f"offset_names {missing} are not in the point (keys: {sorted(point)})")
            self.set_offset(sum(np.asarray(point[n], dtype=np.float64) for n in self.offset_names))
        value, stats = self.astep(None)
        new_point = dict(point)
        new_point[self.vars[0].name] = value
        return new_point, stats

    def publish_history(self):
        """Kept for callers of the round-1 API: the history is published while sampling (see astep)."""
        return None

    def close(self):
        while self._inflight and self.core is not None:      # let the launches that ran ahead settle
            self._inflight.pop(0)
            try:
                self.core.run_wait()
            except Exception:
                break
        self._inflight = []
        self._served = []
        self._ring = None
        self._stop_publisher()
        if self.core is not None:
            self.core.close()
            self.core = None

    def __del__(self):
        try:
            self._stop_publisher()
        except Exception:
            pass
