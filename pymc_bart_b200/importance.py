"""Variable-importance search on the device.

Numeric core of ``pymc_bart.utils.compute_variable_importance`` (pymc_bart/utils.py:868-1090) without the plotting
and without arviz: the ranking methods "VI", "backward" and "backward_VI", the per-sample squared Pearson correlation
between the full-model prediction and every excluded-subset prediction (``pearsonr2``, utils.py:1339-1346), and
``get_variable_inclusion`` (utils.py:747-806).

What changes against the reference is only where the work runs: the reference calls ``_sample_posterior`` once per
candidate subset (O(p) calls for "VI", O(p^2) for "backward", utils.py:996-1002,1034-1036); here every level of the
search is ONE prediction launch over all candidate exclusion masks (X uploaded once) followed by ONE fused
reduction launch for the correlations.  The random draw indices are generated in the reference's order
(one ``rng.integers(0, n_draws, samples)`` per subset, utils.py:63), so a given ``random_seed`` selects the same
posterior draws as the sequential formulation.
"""
from __future__ import annotations

import numpy as np

from .utils import _decode_vi, _get_posterior_sampler


def generate_sequences(n_vars, i_var, include):
    """Candidate exclusion sets of one backward-search level (pymc_bart/utils.py:1330-1336)."""
    if i_var:
        return [tuple(include + [i]) for i in range(n_vars) if i not in include]
    return [()]


def hdi(x, prob=0.94):
    """Narrowest interval holding `prob` of the sample (what arviz_stats' array_stats.hdi returns for a 1-D sample)."""
    x = np.sort(np.asarray(x, dtype=np.float64))
    n = x.size
    k = int(np.floor(prob * n))
    if k < 1 or k >= n:
        return np.array([x[0], x[-1]])
    widths = x[k:] - x[: n - k]
    i = int(np.argmin(widths))
    return np.array([x[i], x[i + k]])


def variable_inclusion_counts(stats, n_vars):
    """Sum of the decoded ``variable_inclusion`` strings of a list of per-draw stats (utils.py:778-790)."""
    tot = np.zeros(n_vars, dtype=np.int64)
    for s in stats:
        v = s["variable_inclusion"] if isinstance(s, dict) else s
        tot += np.asarray(_decode_vi(v, n_vars), dtype=np.int64)
    return tot


def get_variable_inclusion(stats, X, labels=None, to_kulprit=False):
    """Normalised variable inclusion and labels, most included first (pymc_bart/utils.py:747-806).
    `stats`: per-draw stats dicts / strings of ONE BART variable (what idata.sample_stats holds)."""
    n_vars = X.shape[1]
    VIs = variable_inclusion_counts(stats, n_vars)
    VI_norm = VIs / VIs.sum()
    indices = np.argsort(VI_norm)[::-1]
    if hasattr(X, "columns") and hasattr(X, "to_numpy"):
        labels = list(X.columns[indices])
    if labels is None:
        labels = [str(i) for i in indices]
    if to_kulprit:
        return [labels[:idx] for idx in range(n_vars + 1)]
    return VI_norm[indices], labels


def _masks(subsets, n_vars):
    mk = np.zeros((len(subsets), n_vars), dtype=np.uint8)
    for i, s in enumerate(subsets):
        if s is not None and len(s):
            mk[i, np.asarray(list(s), dtype=np.int64)] = 1
    return mk


def _predict_subsets(sampler, Xd, rng, samples, subsets, n_vars):
    """One launch for all `subsets`: returns (device tensor [K][samples][G][n], nothing else).  Draw indices are taken
    from `rng` subset by subset, as the sequential reference does."""
    draws = [rng.integers(0, sampler.n_draws, size=samples) for _ in subsets]
    return sampler.predict_subsets(Xd, draws, _masks(subsets, n_vars))


def compute_variable_importance(stats, bartrv, X, method="VI", fixed=0, samples=50, random_seed=None, ci_prob=0.94):
    """pymc_bart/utils.py:868-1090 for one BART variable; `stats` = its per-draw sample stats (VI methods only).

    Returns the reference's dict: indices, labels, r2_mean, r2_hdi, preds ``(n_vars, samples, n, shape)`` squeezed,
    preds_all."""
    if method not in ["VI", "backward", "backward_VI"]:
        raise ValueError("method must be 'VI', 'backward' or 'backward_VI'")
    rng = np.random.default_rng(random_seed)
    op = bartrv.owner.op if hasattr(bartrv, "owner") else bartrv
    sampler = _get_posterior_sampler(op)
    n_vars = X.shape[1]
    if hasattr(X, "columns") and hasattr(X, "to_numpy"):
        labels = np.asarray(X.columns)
        X = X.to_numpy()
    else:
        labels = np.arange(n_vars).astype(str)
    shape = sampler.n_outputs
    n = X.shape[0]
    Xd = sampler.upload(X)                      # X goes to the device once for the whole search
    r2_mean = np.zeros(n_vars)
    r2_hdi = np.zeros((n_vars, 2))
    preds = np.zeros((n_vars, samples, n, shape))

    def to_host(t):      # [samples][G][n] -> (samples, n, G) like _sample_posterior
        return t.permute(0, 2, 1).cpu().numpy().astype(np.float64)

    if method == "backward_VI":
        if fixed >= n_vars:
            raise ValueError("fixed must be less than the number of variables")
        elif fixed < 1:
            raise ValueError("fixed must be greater than 0")
        init = fixed + 1
    else:
        fixed = 0
        init = 0
    # (the reference leaves predicted_all undefined for backward_VI, utils.py:948-959; it is needed by every method)
    all_dev = sampler.predict_subsets(Xd, [rng.integers(0, sampler.n_draws, size=samples)], _masks([None], n_vars))[0]
    predicted_all = to_host(all_dev)

    indices = list(range(n_vars))
    if method in ["VI", "backward_VI"]:
        idxs = np.argsort(variable_inclusion_counts(stats, n_vars))
        subsets = [list(idxs[:-i]) for i in range(1, len(idxs))]
        subsets.append(None)
        if method == "backward_VI":
            subsets = subsets[-init:]
        indices = list(idxs[::-1])
        sub_dev = _predict_subsets(sampler, Xd, rng, samples, subsets, n_vars)          # ONE launch for all subsets
        r2 = sampler.pearson_r2(all_dev, sub_dev)                                          # [K][samples]
        for idx in range(len(subsets)):
            r2_mean[idx] = np.mean(r2[idx])
            r2_hdi[idx] = hdi(r2[idx], ci_prob)
            preds[idx] = to_host(sub_dev[idx])

    if method in ["backward", "backward_VI"]:
        if method == "backward_VI":
            least_important_vars = [int(v) for v in indices[-fixed:]]
            r2_mean_vi, r2_hdi_vi, preds_vi = r2_mean[:init], r2_hdi[:init], preds[:init]
            r2_mean = np.zeros(n_vars - fixed - 1)
            r2_hdi = np.zeros((n_vars - fixed - 1, 2))
            preds = np.zeros((n_vars - fixed - 1, samples, n, shape))
        else:
            least_important_vars = []
        for i_var in range(init, n_vars):
            subsets = generate_sequences(n_vars, i_var, least_important_vars)
            sub_dev = _predict_subsets(sampler, Xd, rng, samples, subsets, n_vars)      # one launch per search level
            r2 = sampler.pearson_r2(all_dev, sub_dev)
            means = r2.mean(axis=1)
            best = int(np.argmax(means))          # first maximum, like the reference's strict `>` scan
            r2_mean[i_var - init] = means[best]
            r2_hdi[i_var - init] = hdi(r2[best], ci_prob)
            preds[i_var - init] = to_host(sub_dev[best])
            for var_i in subsets[best]:
                if var_i not in least_important_vars:
                    least_important_vars.append(int(var_i))
        for var_i in range(n_vars):
            if var_i not in least_important_vars:
                least_important_vars.append(var_i)
        if method == "backward_VI":
            r2_mean = np.concatenate((r2_mean[::-1], r2_mean_vi))
            r2_hdi = np.concatenate((r2_hdi[::-1], r2_hdi_vi))
            preds = np.concatenate((preds[::-1], preds_vi))
        else:
            r2_mean, r2_hdi, preds = r2_mean[::-1], r2_hdi[::-1], preds[::-1]
        indices = least_important_vars[::-1]

    labels = np.array(["+ " + ele if index != 0 else ele for index, ele in enumerate(labels[np.asarray(indices)])])
    return {"indices": np.asarray(indices), "labels": labels, "r2_mean": r2_mean, "r2_hdi": r2_hdi,
            "preds": preds.squeeze(), "preds_all": predicted_all.squeeze()}


def vi_to_kulprit(vi_results: dict):
    """pymc_bart/utils.py:1093-1108."""
    clean_labels = [label.strip("+ ") for label in vi_results["labels"]]
    return [clean_labels[:idx] for idx in range(len(clean_labels))]
