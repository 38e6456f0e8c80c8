"""CPU tests of the host-side logic in front of the hot path: ensemble-member preprocessing (shared recipe
cores, vectorised fingerprint), task sharding, and the algorithmic FLOP model bench.py reports against."""
import numpy as np

import bench
from multimodalpfn_b200.preprocessing import (_fingerprint, _quantile_uniform, fit_transform_all, make_members,
                                              transform_all)
from multimodalpfn_b200.synth import make_dataset
from multimodalpfn_b200.tasks import shard_tasks


def _members(n=8, F=21, n_cls=6, seed=0):
    return make_members(n, F, n_cls, np.random.default_rng(seed))


def test_shared_cores_equal_member_by_member():
    d = make_dataset("pad_ufes_small", 0)
    a, b = _members(), _members()
    one = [m.fit_transform(d["X_train"], d["y_train"]) for m in a]
    both = fit_transform_all(b, d["X_train"], d["y_train"])
    for (x1, y1), (x2, y2) in zip(one, both):
        assert np.array_equal(x1, x2, equal_nan=True) and np.array_equal(y1, y2)
    t1 = [m.transform(d["X_test"]) for m in a]
    t2 = transform_all(b, d["X_test"])
    assert all(np.array_equal(x, y, equal_nan=True) for x, y in zip(t1, t2))
    # members of one recipe share one fitted core; the two recipes give the two feature widths
    assert len({id(m.core) for m in b}) == 2
    assert sorted({x.shape[1] for x in t2}) == [22, 35]
    # transforming the train rows again reproduces what fit returned (fit and predict use the same arithmetic)
    again = transform_all(b, d["X_train"])
    assert all(np.array_equal(x, y[0], equal_nan=True) for x, y in zip(again, both))


def test_member_structure():
    ms = _members(8, 21, 6, seed=3)
    assert [m.recipe for m in ms] == ["quantile_svd"] * 4 + ["none"] * 4
    for m in ms:
        assert sorted(m.class_perm.tolist()) == list(range(6))
    assert len({m.feature_shift_seed for m in ms}) == 8


def test_fingerprint_is_deterministic_and_row_local():
    rng = np.random.default_rng(1)
    X = rng.standard_normal((200, 17)).astype(np.float32)
    X[3, 4] = np.nan
    f1, f2 = _fingerprint(X), _fingerprint(X.copy())
    assert np.array_equal(f1, f2) and f1.dtype == np.float32
    assert ((f1 >= 0) & (f1 < 1)).all() and len(np.unique(f1)) > 190
    perm = rng.permutation(200)
    assert np.array_equal(_fingerprint(X[perm]), f1[perm])       # a row's value does not depend on its neighbours
    Y = X.copy()
    Y[7, 0] += 1.0
    g = _fingerprint(Y)
    assert g[7] != f1[7] and np.array_equal(np.delete(g, 7), np.delete(f1, 7))


def test_quantile_transform_matches_sklearn():
    from sklearn.preprocessing import QuantileTransformer
    rng = np.random.default_rng(2)
    Xtr = rng.lognormal(size=(500, 3)).astype(np.float32)
    Xte = rng.lognormal(size=(120, 3)).astype(np.float32) * 1.5
    Xte[5, 1] = np.nan
    qt = QuantileTransformer(n_quantiles=50, output_distribution="uniform", random_state=0).fit(Xtr)
    ref = qt.transform(Xte)
    got = _quantile_uniform(Xte, np.asarray(qt.quantiles_, np.float64), np.asarray(qt.references_, np.float64))
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    assert np.nanmax(np.abs(got - ref)) < 1e-6


def test_task_sharding():
    for world in (1, 2, 3, 8):
        owned = [shard_tasks(256, r, world) for r in range(world)]
        assert sorted(sum(owned, [])) == list(range(256))
        assert max(map(len, owned)) - min(map(len, owned)) <= 1


def test_flop_model_matches_survey_totals():
    # cfg2 T=27 2.06 TFLOP, T=20 1.53 TFLOP per estimator (12 layers)
    assert abs(bench.flops_estimator(2000, 300, 27) / 1e12 - 2.03) < 0.02
    assert abs(bench.flops_estimator(2000, 300, 20) / 1e12 - 1.50) < 0.02
    total, n_tr, n_te = bench.full_flops()
    assert (n_tr, n_te) == (2000, 300) and abs(total / 1e12 - 14.09) < 0.05
    assert bench.flops_item_attention(2000, 2000, 27, 4) == 4.0 * 4 * 27 * 6 * 2000 * 2000 * 32
