"""CPU models of two index computations of the step kernel, checked exhaustively against brute force.

They restate (in numpy / plain Python) the arithmetic of pymc_bart_b200/csrc/pgbart_b200.cu so that the edge
cases the GPU parity tests cannot reach in seconds (N = 1M and beyond, very few or very many units) are covered:

* `worker_loop`: static split of an epoch's (tile, job) pairs over worker CTAs, groups and warps — every pair must
  be executed exactly once;
* `select_split`: the search of the k-th member over the per-tile counts (single level for small N, two levels
  otherwise) — must return the tile and the offset inside it.
"""
import numpy as np
import pytest

NGROUPS, WARPS_PER_GROUP = 4, 8


def servers_of(C, c):
    return 1 if C >= NGROUPS else (NGROUPS - c + C - 1) // C


def group_range(T, W, w, ns, rank):
    """[lo, hi) of worker CTA w, serving group `rank` of `ns` (32-bit arithmetic of the kernel)."""
    Tq, Tr = T // W, T - (T // W) * W
    clo = w * Tq + (w * Tr) // W
    chi = (w + 1) * Tq + ((w + 1) * Tr) // W
    return clo + ((chi - clo) * rank) // ns, clo + ((chi - clo) * (rank + 1)) // ns


def warp_segments(lo_g, hi_g, njobs):
    """(tile, first job, last job + 1) segments executed by the warps of one group."""
    segs = []
    chunk = (hi_g - lo_g + WARPS_PER_GROUP - 1) // WARPS_PER_GROUP
    for warp in range(WARPS_PER_GROUP):
        lo = lo_g + warp * chunk
        hi = min(lo + chunk, hi_g)
        while lo < hi:
            tile, j0 = divmod(lo, njobs)
            seg = min(njobs - j0, hi - lo)
            segs.append((tile, j0, j0 + seg))
            lo += seg
    return segs


@pytest.mark.parametrize("ntiles,njobs,W,C", [(1, 1, 147, 1), (3, 2, 147, 4), (391, 9, 147, 4), (391, 39, 144, 1), (3907, 59, 147, 1),
                                               (8, 127, 140, 2), (40000, 3, 147, 3), (5, 1, 3, 8)])
def test_every_pair_of_an_epoch_runs_exactly_once(ntiles, njobs, W, C):
    T = ntiles * njobs
    for c in range(min(C, NGROUPS)):
        ns = servers_of(C, c)
        seen = np.zeros((ntiles, njobs), dtype=np.int32)
        sizes = []
        for w in range(W):
            for rank in range(ns):
                lo, hi = group_range(T, W, w, ns, rank)
                sizes.append(hi - lo)
                for tile, a, b in warp_segments(lo, hi, njobs):
                    seen[tile, a:b] += 1
        assert np.all(seen == 1)
        assert max(sizes) - min(sizes) <= 1 + (1 if ns > 1 else 0)      # balanced to a pair per CTA (and per group)
    # the groups that serve the chains cover all of a CTA's groups
    if C < NGROUPS:
        assert sum(servers_of(C, c) for c in range(C)) == NGROUPS


def scan32(v):
    return np.cumsum(v)


def select_model(counts, k):
    """Mirror of select_split's search: counts per tile (padded to a multiple of 4) -> (tile, offset in tile)."""
    stride = (len(counts) + 3) & ~3
    cnt = np.zeros(stride, dtype=np.int64)
    cnt[: len(counts)] = counts
    n4 = stride >> 2
    per = (n4 + 31) >> 5
    g = cnt.reshape(n4, 4)
    off = k
    if per <= 4:      # one level: a lane keeps its (at most 16) counts in registers
        lane_sum = np.zeros(32, dtype=np.int64)
        for lane in range(32):
            for u in range(per):
                i4 = lane * per + u
                if i4 < n4:
                    lane_sum[lane] += g[i4].sum()
        incl = scan32(lane_sum); excl = incl - lane_sum
        hit = [l for l in range(32) if excl[l] <= off < incl[l]]
        if not hit:
            return None
        src = hit[0]; o2 = off - excl[src]; t = 0; found = False
        for u in range(4):
            i4 = src * per + u
            cs = g[i4] if (u < per and i4 < n4) else np.zeros(4, dtype=np.int64)
            for e in range(4):
                if not found:
                    if o2 < cs[e]:
                        found, t = True, u * 4 + e
                    else:
                        o2 -= cs[e]
        return src * per * 4 + t, int(o2)
    lane_sum = np.array([g[l * per: min((l + 1) * per, n4)].sum() if l * per < n4 else 0 for l in range(32)], dtype=np.int64)
    incl = scan32(lane_sum); excl = incl - lane_sum
    hit = [l for l in range(32) if excl[l] <= off < incl[l]]
    if not hit:
        return None
    src = hit[0]; off -= excl[src]
    for j0 in range(0, per, 32):
        s4 = np.zeros(32, dtype=np.int64); vals = np.zeros((32, 4), dtype=np.int64)
        for lane in range(32):
            i4 = src * per + j0 + lane
            if j0 + lane < per and i4 < n4:
                vals[lane] = g[i4]; s4[lane] = g[i4].sum()
        in2 = scan32(s4); ex2 = in2 - s4
        hit2 = [l for l in range(32) if ex2[l] <= off < in2[l]]
        if hit2:
            l2 = hit2[0]; o2 = off - ex2[l2]; t = 0
            v = vals[l2]
            if o2 >= v[0]:
                o2 -= v[0]; t = 1
                if o2 >= v[1]:
                    o2 -= v[1]; t = 2
                    if o2 >= v[2]:
                        o2 -= v[2]; t = 3
            return (src * per + j0 + l2) * 4 + t, int(o2)
        off -= in2[31]
    return None


@pytest.mark.parametrize("ntiles", [1, 3, 4, 5, 127, 128, 129, 391, 512, 513, 3907, 4100, 40001])
def test_kth_member_search_matches_brute_force(ntiles):
    rng = np.random.default_rng(ntiles)
    for density in (0.02, 0.5, 1.0):
        counts = (rng.integers(0, 257, size=ntiles) * (rng.uniform(size=ntiles) < density)).astype(np.int64)
        total = int(counts.sum())
        if total == 0:
            assert select_model(counts, 0) is None
            continue
        cum = np.cumsum(counts)
        ks = sorted(set([0, total - 1] + rng.integers(0, total, size=60).tolist()))
        for k in ks:
            tile = int(np.searchsorted(cum, k, side="right"))
            want = (tile, k - (int(cum[tile - 1]) if tile else 0))
            assert select_model(counts, k) == want, (ntiles, density, k)
        assert select_model(counts, total) is None      # k out of range is reported, never mis-indexed
