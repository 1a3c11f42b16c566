"""Pins the CPU oracle's alias table and SkipGram / CBOW update (CPU only, no GPU).

The reference holds no numeric fixtures for this path (parity unpinned against
Ensmallen), so the C oracle is cross-checked here against an independent float64 numpy
restatement of the normative recipe in DESIGN.md (kwarg semantics from
/root/reference/embiggen/embedders/ensmallen_embedders/node2vec_skipgram.py:37-119; model
structure from /root/reference/embiggen/embedders/tensorflow_embedders/skipgram.py:28-61 and
cbow.py:28-60) and against the closed forms the domain offers.
"""
import numpy as np
import pytest
from scipy import stats

import oracle
from embiggen_b200.graph import philox4x32

PAD = oracle.PAD_TOKEN
TAG_NEG = 3


# ---------------------------------------------------------------- alias table
@pytest.mark.parametrize("alpha", [0.0, 0.5, 0.75, 1.0, 0.6])
def test_alias_table_reproduces_target_pmf(rmat_graph, alpha):
    degrees = np.diff(rmat_graph.indptr).astype(np.float64)
    target = np.where(degrees > 0, degrees ** alpha, 0.0)
    target /= target.sum()
    thr, alias = oracle.alias_build(rmat_graph.indptr, alpha)
    n = len(degrees)
    keep = (thr.astype(np.float64) + (thr == 0xFFFFFFFF)) / 2.0 ** 32  # 0xFFFFFFFF means "always"
    pmf = keep / n
    np.add.at(pmf, alias, (1.0 - keep) / n)
    assert np.abs(pmf - target).max() < 2e-9
    assert pmf[degrees == 0].max(initial=0.0) < 1e-9  # isolated nodes are (almost) never drawn


def closed_form_alias(indptr, alpha):
    """The table of oracle/alias.c in closed form (what the GPU builder computes, alias_build.cu):
    light node k is topped up by heavy node #{j : S_j < D_k}; heavy node j is exhausted by the
    first light node whose running deficit exceeds S_j and keeps c + S_j - that running deficit."""
    import ctypes
    lib = oracle.lib()
    lib.orc_alias_fraction_bits.restype = ctypes.c_int
    lib.orc_alias_fraction_bits.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_double]
    deg = np.diff(indptr).astype(np.uint64)
    n = len(deg)
    d = deg.astype(np.float64)
    w = {0.0: np.ones_like(d), 1.0: d, 0.5: np.sqrt(d), 0.75: np.sqrt(np.sqrt(d * d * d))}.get(alpha)
    w = np.where(deg > 0, d ** alpha if w is None else w, 0.0)
    bits = lib.orc_alias_fraction_bits(n, int(deg.max()), alpha)
    mass = np.floor(np.ldexp(w, bits)).astype(np.uint64)
    total = int(mass.sum())
    c = (total + n - 1) // n
    nz = deg > 0
    each, first = divmod(c * n - total, int(nz.sum()))
    rank = np.cumsum(nz) - 1
    mass[nz] += np.uint64(each) + (rank[nz] < first).astype(np.uint64)

    def threshold(m):
        r, q = int(m), 0
        for _ in range(32):
            r, q = r << 1, q << 1
            if r >= c:
                r, q = r - c, q | 1
        return q

    light, heavy = np.flatnonzero(mass < c), np.flatnonzero(mass >= c)
    thr = np.full(n, 0xFFFFFFFF, dtype=np.uint32)
    alias = np.arange(n, dtype=np.uint32)
    deficit, surplus = np.uint64(c) - mass[light], mass[heavy] - np.uint64(c)
    d_sum, s_sum = np.cumsum(deficit), np.cumsum(surplus)
    thr[light] = [threshold(m) for m in mass[light]]
    alias[light] = heavy[np.searchsorted(s_sum, d_sum - deficit, side="left")]
    k = np.searchsorted(d_sum, s_sum, side="right")
    exhausted = np.flatnonzero(k < len(light))
    thr[heavy[exhausted]] = [threshold(c + int(s_sum[j]) - int(d_sum[k[j]])) for j in exhausted]
    alias[heavy[exhausted]] = heavy[exhausted + 1]
    return thr, alias


@pytest.mark.parametrize("alpha", [0.0, 0.5, 0.75, 1.0, 0.6])
def test_alias_sweep_equals_its_closed_form(rmat_graph, er_graph, small_ppi, alpha):
    """oracle/alias.c sweeps the light and heavy nodes sequentially; the product builds the same
    table from two prefix sums.  Here the closed form is restated in numpy and compared with the
    sweep, so the derivation is checked on the CPU before the GPU test compares the kernels."""
    for graph in (rmat_graph, er_graph, small_ppi):
        thr, alias = oracle.alias_build(graph.indptr, alpha)
        c_thr, c_alias = closed_form_alias(graph.indptr, alpha)
        assert np.array_equal(thr, c_thr) and np.array_equal(alias, c_alias)


def test_alias_sampling_chi_square(small_ppi):
    thr, alias = oracle.alias_build(small_ppi.indptr, 0.75)
    n = len(thr)
    draws = 400_000
    r0, r1, _, _ = philox4x32(99, np.arange(draws, dtype=np.uint64), 0, 0, 0)
    idx = ((r0.astype(np.uint64) * np.uint64(n)) >> np.uint64(32)).astype(np.int64)
    node = np.where(r1 < thr[idx], idx, alias[idx])
    degrees = np.diff(small_ppi.indptr).astype(np.float64)
    target = degrees ** 0.75 / (degrees ** 0.75).sum()
    counts = np.bincount(node, minlength=n)
    order = np.argsort(target)
    # pool the many degree-1 nodes so every bin expects >= 50 draws
    bins = np.array_split(order, 60)
    observed = np.array([counts[b].sum() for b in bins])
    expected = np.array([target[b].sum() for b in bins]) * draws
    assert stats.chisquare(observed, expected).pvalue > 1e-3


# ------------------------------------------------- numpy restatement of the update
def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def draw_negatives(seed, wid, site, K, n, thr, alias, centre, context):
    negs, valid = [], []
    for k in range(K):
        r0, r1, _, _ = philox4x32(seed, wid & 0xFFFFFFFF, wid >> 32, site, (TAG_NEG << 24) | k)
        idx = (int(r0) * n) >> 32
        node = idx
        if thr is not None:
            node = idx if int(r1) < int(thr[idx]) else int(alias[idx])
        ok = node != centre and node != context and node not in negs
        negs.append(node)
        valid.append(ok)
    return negs, valid


def apply_targets(h, t1, targets, valid, lr, clip, scale, stats_out):
    rows = [t1[u].copy() for u in targets]
    acc = np.zeros_like(h)
    for k, (u, ok) in enumerate(zip(targets, valid)):
        if not ok:
            continue
        stats_out["targets"] += 1
        f = float(h @ rows[k]) * scale
        if abs(f) > clip:
            continue
        label = 1.0 if k == 0 else 0.0
        g = (label - sigmoid(f)) * lr
        stats_out["loss"] += np.log1p(np.exp(-f if k == 0 else f))
        acc += g * rows[k]
        t1[u] = rows[k] + g * h
    return acc


def reference_train(model, walks, t0, t1, seed, n, D, w, K, lr, clip=6.0, first_walk=0, thr=None,
                    alias=None, indptr=None, normalize=False, scale_by_sqrt_dim=False, downsample=False,
                    shared=False):
    bound = int(np.diff(indptr).max()) + 1 if downsample else 0
    t0, t1 = t0.astype(np.float64), t1.astype(np.float64)
    out = {"pairs": 0, "targets": 0, "loss": 0.0}
    scale = 1.0 / np.sqrt(D) if scale_by_sqrt_dim else 1.0
    L = walks.shape[1]
    for row_index, walk in enumerate(walks):
        wid = first_walk + row_index
        for i in range(L):
            c = int(walk[i])
            if c == PAD:
                break
            if bound and (oracle.philox(seed, wid & 0xFFFFFFFF, wid >> 32, i, 6 << 24)[0] * bound) >> 32 \
                    < int(indptr[c + 1] - indptr[c]):
                continue  # stochastic_downsample_by_degree
            step = lr / float(indptr[c + 1] - indptr[c]) if normalize else lr
            window = [j for j in range(max(0, i - w), min(L - 1, i + w) + 1)
                      if j != i and int(walk[j]) != PAD and int(walk[j]) != c]
            if model == "SkipGram" and shared:  # one set of negatives per centre (sgns.c: train_centre_shared)
                if not window:
                    continue
                ctx = [int(walk[j]) for j in window]
                h = t0[c].copy()
                negs, valid = draw_negatives(seed, wid, (i << 16) | 0xFFFF, K, n, thr, alias, c, c)
                valid = [ok and u not in ctx for u, ok in zip(negs, valid)]
                before = t1.copy()
                acc = np.zeros_like(h)
                for u, ok in zip(negs, valid):
                    if not ok:
                        continue
                    out["targets"] += 1
                    f = float(h @ before[u]) * scale
                    if abs(f) > clip:
                        continue
                    g = (0.0 - 1.0 / (1.0 + np.exp(-f))) * step * len(ctx)
                    out["loss"] += len(ctx) * np.log1p(np.exp(f))
                    acc += g * before[u]
                    t1[u] = t1[u] + g * h
                for o in ctx:
                    out["targets"] += 1
                    f = float(h @ before[o]) * scale
                    if abs(f) > clip:
                        continue
                    g = (1.0 - 1.0 / (1.0 + np.exp(-f))) * step
                    out["loss"] += np.log1p(np.exp(-f))
                    acc += g * before[o]
                    t1[o] = t1[o] + g * h
                t0[c] = h + acc
                out["pairs"] += len(ctx)
            elif model == "SkipGram":
                h = t0[c].copy()
                for j in window:
                    o = int(walk[j])
                    negs, valid = draw_negatives(seed, wid, (i << 16) | j, K, n, thr, alias, c, o)
                    h = h + apply_targets(h, t1, [o] + negs, [True] + valid, step, clip, scale, out)
                    out["pairs"] += 1
                t0[c] = h
            else:
                if not window:
                    continue
                ctx = [int(walk[j]) for j in window]
                h = t0[ctx].sum(axis=0) / len(ctx)
                negs, valid = draw_negatives(seed, wid, (i << 16) | 0xFFFF, K, n, thr, alias, c, c)
                acc = apply_targets(h, t1, [c] + negs, [True] + valid, step, clip, scale, out)
                for o in ctx:
                    t0[o] = t0[o] + acc
                out["pairs"] += len(ctx)
    return t0, t1, out


CASES = [
    ("SkipGram", 16, 5, 2, dict()),
    ("CBOW", 16, 5, 2, dict()),
    ("SkipGram", 7, 3, 1, dict(use_alias=False)),
    ("CBOW", 12, 4, 3, dict(normalize=True, lr=0.5)),
    ("SkipGram", 20, 6, 3, dict(scale_by_sqrt_dim=True, lr=0.2)),
    ("SkipGram", 8, 10, 4, dict(clip=0.02, lr=0.5)),
    ("CBOW", 8, 0, 2, dict()),
    ("SkipGram", 9, 4, 3, dict(downsample=True)),
    ("CBOW", 9, 4, 3, dict(downsample=True)),
    ("SkipGram", 16, 5, 2, dict(shared=True)),
    ("SkipGram", 12, 10, 4, dict(shared=True, lr=0.1, normalize=True)),
    ("SkipGram", 8, 6, 3, dict(shared=True, clip=0.02, lr=0.5, scale_by_sqrt_dim=True)),
    ("SkipGram", 9, 4, 3, dict(shared=True, downsample=True, use_alias=False)),
]


@pytest.mark.parametrize("model,D,K,w,options", CASES)
def test_c_oracle_matches_numpy_restatement(small_ppi, model, D, K, w, options):
    seed, L, n_walks, first = 17, 12, 40, 1000
    lr = options.get("lr", 0.05)
    clip = options.get("clip", 6.0)
    n = small_ppi.get_number_of_nodes()
    walks, _ = oracle.walks(small_ppi.indptr, small_ppi.indices, seed, first, n_walks, L, 0.25, 4.0)
    thr = alias = None
    if options.get("use_alias", True):
        thr, alias = oracle.alias_build(small_ppi.indptr, 0.75)
    t0, t1 = oracle.init_tables(n, D, seed)
    # make the rows large enough for the sigmoid to leave its linear range
    t0 *= 30.0
    t1 *= 30.0
    e0, e1, expected = reference_train(
        model, walks, t0[:, :D], t1[:, :D], seed, n, D, w, K, lr, clip, first, thr, alias,
        small_ppi.indptr, options.get("normalize", False), options.get("scale_by_sqrt_dim", False),
        options.get("downsample", False), options.get("shared", False))
    got = oracle.train(model, walks, t0, t1, seed, n, D, w, K, lr, clip, first_walk=first, thr=thr,
                       alias=alias, indptr=small_ppi.indptr,
                       normalize_learning_rate_by_degree=options.get("normalize", False),
                       scale_by_sqrt_dim=options.get("scale_by_sqrt_dim", False),
                       stochastic_downsample_by_degree=options.get("downsample", False),
                       shared_negatives=options.get("shared", False))
    assert got["pairs"] == expected["pairs"] and got["targets"] == expected["targets"]
    assert got["pairs"] > 0
    assert np.isclose(got["loss_sum"], expected["loss"], rtol=1e-5)
    # float32 rounding of the C oracle against float64: a few ulps of the largest entries
    assert np.allclose(t0[:, :D], e0, rtol=0, atol=2e-6 * max(1.0, np.abs(e0).max() / 4))
    assert np.allclose(t1[:, :D], e1, rtol=0, atol=2e-6 * max(1.0, np.abs(e1).max() / 4))
    assert (t0[:, D:] == 0).all() and (t1[:, D:] == 0).all()  # row padding stays zero


def test_pairs_per_walk_closed_form(er_graph):
    """P(L, w) = 2wL - w(w+1) positives per walk when no token repeats inside a window."""
    from embiggen_b200.engine import pairs_per_walk
    n = er_graph.get_number_of_nodes()
    for L, w in [(16, 2), (32, 4), (9, 5)]:
        walks = np.arange(3 * L, dtype=np.uint32).reshape(3, L)  # distinct tokens
        t0, t1 = oracle.init_tables(n, 8, 1)
        r = oracle.train("SkipGram", walks, t0, t1, 1, n, 8, w, 2, 0.01)
        assert r["pairs"] == 3 * pairs_per_walk(L, w)
        t0, t1 = oracle.init_tables(n, 8, 1)
        r = oracle.train("CBOW", walks, t0, t1, 1, n, 8, w, 2, 0.01)
        assert r["pairs"] == 3 * pairs_per_walk(L, w)


def test_init_tables_distribution_and_padding():
    t0, t1 = oracle.init_tables(5000, 10, 3)
    assert t0.shape == (5000, 12) and (t0[:, 10:] == 0).all() and (t1[:, 10:] == 0).all()
    for t in (t0, t1):
        values = t[:, :10].ravel().astype(np.float64) * 10
        assert values.min() >= -0.5 and values.max() < 0.5
        assert stats.kstest(values + 0.5, "uniform").pvalue > 1e-3
    assert not np.array_equal(t0, t1)
    again0, _ = oracle.init_tables(5000, 10, 3)
    assert np.array_equal(again0, t0)


def test_sigmoid_and_dot_primitives():
    xs = np.linspace(-20, 20, 4001, dtype=np.float32)
    got = np.array([oracle.sigmoid(float(x)) for x in xs])
    assert np.abs(got - 1.0 / (1.0 + np.exp(-xs.astype(np.float64)))).max() < 2e-7
    rng = np.random.default_rng(0)
    for length in (4, 8, 100, 128, 300):
        a = rng.standard_normal(length).astype(np.float32)
        b = rng.standard_normal(length).astype(np.float32)
        stride = (length + 3) // 4 * 4
        pa, pb = np.zeros(stride, np.float32), np.zeros(stride, np.float32)
        pa[:length], pb[:length] = a, b
        assert abs(oracle.dot(pa, pb) - float(a.astype(np.float64) @ b.astype(np.float64))) < 1e-4


def test_training_lowers_the_objective(small_ppi):
    o0, o1, losses = oracle.fit("SkipGram", small_ppi.indptr, small_ppi.indices, 42, 16, 3, 1, 24, 3,
                                5, 0.05, 0.9, return_weight=0.25, explore_weight=4.0)
    assert losses[0] > losses[-1] and np.isfinite(o0).all() and np.isfinite(o1).all()


def test_stochastic_downsample_skips_centres_in_proportion_to_degree():
    """A star: the hub (degree 40 = max) is skipped with probability 40/41, a leaf with 1/41.
    With window 1 and no negatives every surviving centre contributes exactly its context
    count, so the pair count measures the skip rate (node2vec_skipgram.py:97-98)."""
    from conftest import tiny_graphs
    star = tiny_graphs()["star"]
    n = star.get_number_of_nodes()
    deg = np.diff(star.indptr)
    hub = int(deg.argmax())
    bound = int(deg.max()) + 1
    walks, _ = oracle.walks(star.indptr, star.indices, 3, 0, 4000, 16)
    t0, t1 = oracle.init_tables(n, 4, 3)
    full = oracle.train("SkipGram", walks, t0.copy(), t1.copy(), 3, n, 4, 1, 0, 0.01, indptr=star.indptr)
    kept = oracle.train("SkipGram", walks, t0, t1, 3, n, 4, 1, 0, 0.01, indptr=star.indptr,
                        stochastic_downsample_by_degree=True)
    # expected surviving pairs: sum over centres of contexts * (1 - deg/bound)
    L = walks.shape[1]
    contexts = np.full(walks.shape, 2); contexts[:, 0] = 1; contexts[:, -1] = 1
    keep_p = 1.0 - deg[walks] / bound
    expected = (contexts * keep_p).sum()
    assert full["pairs"] == contexts.sum()
    sd = np.sqrt((contexts ** 2 * keep_p * (1 - keep_p)).sum())
    assert abs(kept["pairs"] - expected) < 5 * sd
    assert kept["pairs"] < 0.7 * full["pairs"] and (walks == hub).mean() > 0.4


# ---- Walklets (walklets.py:7-149): scale k = pairs exactly k hops apart ----
def test_walklet_split_keeps_exactly_the_pairs_k_hops_apart(er_graph):
    L = 23
    walks, _ = oracle.walks(er_graph.indptr, er_graph.indices, 4, 0, 50, L)
    walks[7, 15:] = oracle.PAD_TOKEN  # a walk that ended early
    for k in (1, 2, 3, 5, 22):
        sub = oracle.walklet_split(walks, k)
        assert sub.shape == (k, 50, (L + k - 1) // k)
        pairs = set()
        for r in range(k):
            for w in range(50):
                row = sub[r, w]
                for m in range(len(row) - 1):
                    if row[m] != oracle.PAD_TOKEN and row[m + 1] != oracle.PAD_TOKEN:
                        pairs.add((w, r + m * k, r + (m + 1) * k))
        expected = {(w, i, i + k) for w in range(50) for i in range(L - k)
                    if walks[w, i] != oracle.PAD_TOKEN and walks[w, i + k] != oracle.PAD_TOKEN}
        assert pairs == expected


def test_walklet_training_counts_the_pairs_of_its_scale(small_ppi):
    n, D, L = small_ppi.get_number_of_nodes(), 8, 20
    walks, _ = oracle.walks(small_ppi.indptr, small_ppi.indices, 9, 0, 64, L, 0.25, 4.0)
    for k in (2, 3):
        t0, t1 = oracle.init_tables(n, D, 9)
        stats = oracle.train_walklets("SkipGram", walks, k, t0, t1, 9, n, D, 1, 0, 0.05)
        # every ordered pair (i, i +- k) with two different tokens
        a, b = walks[:, :-k], walks[:, k:]
        assert stats["pairs"] == 2 * int((a != b).sum())


# ---- GloVe (oracle/glove.c) ----
def test_cooccurrence_against_a_python_loop(small_ppi):
    L, w = 12, 3
    walks, _ = oracle.walks(small_ppi.indptr, small_ppi.indices, 2, 0, 60, L, 0.25, 4.0)
    walks[5, 7:] = oracle.PAD_TOKEN
    expected = {}
    for walk in walks:
        for i in range(L):
            for j in range(max(0, i - w), min(L - 1, i + w) + 1):
                a, b = int(walk[i]), int(walk[j])
                if j != i and a != oracle.PAD_TOKEN and b != oracle.PAD_TOKEN and a != b:
                    expected[(a, b)] = expected.get((a, b), 0) + 1
    centre, context, count = oracle.cooccurrence(walks, w)
    got = {(int(a), int(b)): int(c) for a, b, c in zip(centre, context, count)}
    assert got == expected
    keys = centre.astype(np.uint64) << np.uint64(32) | context
    assert (np.diff(keys.astype(np.int64)) > 0).all()


def test_log_det_and_glove_step_against_float64(small_ppi):
    for x in (1.0, 2.0, 3.0, 7.0, 1000.0, 0.25, 1e-4, 123456.0):
        assert abs(oracle.log_det(x) - np.log(x)) <= 3e-7 * max(1.0, abs(np.log(x)))
    n, D, alpha, lr = small_ppi.get_number_of_nodes(), 12, 0.75, 0.05
    walks, _ = oracle.walks(small_ppi.indptr, small_ppi.indices, 3, 0, 200, 16)
    centre, context, count = oracle.cooccurrence(walks, 3)
    t0, t1 = oracle.init_tables(n, D, 3)
    t0 *= 25.0
    t1 *= 25.0
    e0, e1 = t0[:, :D].astype(np.float64), t1[:, :D].astype(np.float64)
    xmax, loss = float(count.max()), 0.0
    for c, o, x in zip(centre, context, count):
        f = e0[c] @ e1[o]
        if abs(f) > 6.0:
            continue
        weight, diff = (x / xmax) ** alpha, f - np.log(x)
        g = 2.0 * weight * diff * lr
        loss += weight * diff * diff
        e0[c], e1[o] = e0[c] - g * e1[o], e1[o] - g * e0[c]
    got = oracle.glove_train(centre, context, count, t0, t1, D, alpha, lr)
    assert got["trained"] == len(centre) and np.isclose(got["loss_sum"], loss, rtol=1e-5)
    assert np.allclose(t0[:, :D], e0, atol=3e-6) and np.allclose(t1[:, :D], e1, atol=3e-6)
    assert (t0[:, D:] == 0).all() and (t1[:, D:] == 0).all()
    _, _, losses = oracle.glove_fit(small_ppi.indptr, small_ppi.indices, 5, 16, 6, 32, 4, 0.75, 0.05, 0.9)
    assert losses[-1] < 0.8 * losses[0]


def test_shared_negatives_mode_counts(small_ppi):
    """The opt-in estimator sees the same (centre, context) pairs as the per-pair model and scores
    at most one row per context position plus K negatives per centre; a centre with no valid
    context draws nothing (the site key is the centre's, like CBOW's)."""
    seed, L, w, K, D = 5, 24, 3, 6, 8
    n = small_ppi.get_number_of_nodes()
    walks, _ = oracle.walks(small_ppi.indptr, small_ppi.indices, seed, 0, 200, L, 2.0, 0.5)
    thr, alias = oracle.alias_build(small_ppi.indptr, 0.75)
    stats = {}
    for shared in (False, True):
        t0, t1 = oracle.init_tables(n, D, seed)
        stats[shared] = oracle.train("SkipGram", walks, t0, t1, seed, n, D, w, K, 0.05, thr=thr, alias=alias,
                                     shared_negatives=shared)
        assert np.isfinite(t0).all() and np.isfinite(t1).all()
    centres = int((walks != oracle.PAD_TOKEN).sum())
    assert stats[True]["pairs"] == stats[False]["pairs"] > 0
    assert stats[True]["pairs"] < stats[True]["targets"] <= stats[True]["pairs"] + K * centres
    assert stats[True]["targets"] < stats[False]["targets"] / 2
    with pytest.raises(ValueError):  # a SkipGram option
        t0, t1 = oracle.init_tables(n, D, seed)
        oracle.train("CBOW", walks, t0, t1, seed, n, D, w, K, 0.05, thr=thr, alias=alias, shared_negatives=True)
