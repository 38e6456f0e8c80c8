"""Replay of the reference's fitted preprocessors (multimodalpfn_b200/ref_transform.py) against the reference's own
``transform`` — bit for bit, on the real reference classifier / regressor objects (host code only: runs without a
GPU).  Skipped where neither /root/reference nor the oracle/_ref snapshot exists."""
import numpy as np
import pytest
import torch

from oracle import ref_compat

pytestmark = pytest.mark.skipif(not ref_compat.reference_available(), reason="reference not present")


def _fitted(tmp_path, dataset, n_est, regression=False):
    from multimodalpfn_b200.synth import Geometry, make_checkpoint_config, make_dataset, make_state_dict
    ref_compat.install()
    geom = Geometry(mgm_heads=2, cap_heads=4, n_out=64 if regression else 10)
    sd = make_state_dict(geom, seed=11, regression=regression)
    path = str(tmp_path / "m.ckpt")
    torch.save({"state_dict": {k: torch.as_tensor(v) for k, v in sd.items()},
                "config": make_checkpoint_config(geom, regression=regression)}, path)
    d = make_dataset(dataset, 0)
    kw = dict(mixer_type="MGM+CAP", mgm_heads=2, cap_heads=4, features_per_group=2, n_estimators=n_est, model_path=path,
              device="cpu", ignore_pretraining_limits=True, random_state=0)
    if regression:
        import mmpfn.models.mmpfn.regressor as R
        y = (d["X_train"][:, 0] * 0.7 + d["y_train"]).astype(np.float32)
        est = R.MMPFNRegressor(**kw).fit(d["X_train"], d["img_train"], y)
    else:
        import mmpfn.models.mmpfn.classifier as C
        est = C.MMPFNClassifier(**kw).fit(d["X_train"], d["img_train"], d["y_train"])
    return est, d


def _boundary_table(est, X):
    """The table ``predict_proba`` hands to ``iter_outputs`` (classifier.py:527-530)."""
    from mmpfn.models.mmpfn.utils import _fix_dtypes, validate_X_predict
    X = validate_X_predict(X, est)
    X = _fix_dtypes(X, cat_indices=est.categorical_features_indices)
    return est.preprocessor_.transform(X)


@pytest.mark.parametrize("dataset", ["tiny", "pad_ufes_small"])
def test_replay_is_bit_identical_classifier(tmp_path, dataset):
    from multimodalpfn_b200 import ref_transform as RT
    clf, d = _fitted(tmp_path, dataset, 8)
    X = _boundary_table(clf, d["X_test"])
    probe = RT.make_probe(X)
    recipes = set()
    for pre in clf.executor_.preprocessors:
        fast = RT.compile_preprocessor(pre)
        assert RT.verify(pre, fast, X) and RT.verify(pre, fast, probe)
        ref = pre.transform(probe).X
        got = fast(probe)
        assert got.dtype == ref.dtype and np.array_equal(ref, got, equal_nan=True)
        assert np.isnan(ref).any()                       # the probes did reach the NaN / unknown-category branches
        recipes.add(ref.shape[1])
    assert len(recipes) == 2                             # both default recipes (quantile + SVD, and "none") were covered
    # a single row, and the input untouched
    before = X.copy()
    one = X[:1]
    for pre in clf.executor_.preprocessors:
        assert np.array_equal(pre.transform(one).X, RT.compile_preprocessor(pre)(one), equal_nan=True)
    assert np.array_equal(before, X, equal_nan=True)


def test_shared_replay_of_all_members(tmp_path):
    """``replay_all``: members share equal fitted nodes (quantile transform, scaling chain, ordinal encoder) through a
    per-call memo.  Same tables as member-by-member, on two different tables in a row (the memo does not outlive a
    call), inputs untouched, and the memo is actually hit."""
    import time
    from multimodalpfn_b200 import ref_transform as RT
    clf, d = _fitted(tmp_path, "pad_ufes_small", 8)
    pres = clf.executor_.preprocessors
    fasts = [RT.compile_preprocessor(p) for p in pres]
    X1 = _boundary_table(clf, d["X_test"])
    X2 = RT.make_probe(_boundary_table(clf, d["X_test"][::-1].copy()), seed=3)
    for X in (X1, X2, X1):
        before = X.copy()
        got = RT.replay_all(fasts, X)
        for pre, g in zip(pres, got):
            ref = pre.transform(X).X
            assert g.dtype == ref.dtype and np.array_equal(ref, g, equal_nan=True)
        assert np.array_equal(before, X, equal_nan=True) and RT._MEMO is None
    class Counting(dict):
        lookups = 0

        def get(self, key, default=None):
            Counting.lookups += 1
            return dict.get(self, key, default)
    try:
        RT._MEMO = Counting()
        for f in fasts:
            f(X1)
        entries, lookups = len(RT._MEMO), Counting.lookups
    finally:
        RT._MEMO = None
    assert 0 < entries < lookups, (entries, lookups)          # fewer evaluations than lookups: nodes were shared
    t0 = time.perf_counter()
    for _ in range(10):
        RT.replay_all(fasts, X1)
    t_shared = (time.perf_counter() - t0) / 10
    t0 = time.perf_counter()
    for _ in range(10):
        [f(X1) for f in fasts]
    t_each = (time.perf_counter() - t0) / 10
    t0 = time.perf_counter()
    for _ in range(3):
        [p.transform(X1) for p in pres]
    t_ref = (time.perf_counter() - t0) / 3
    print(f"8 members, {len(X1)} rows: reference transforms {t_ref * 1e3:.1f} ms, replays {t_each * 1e3:.2f} ms, "
          f"shared replays {t_shared * 1e3:.2f} ms ({entries} memo entries for {lookups} lookups)")


def test_replay_refuses_what_the_reference_refuses(tmp_path):
    from multimodalpfn_b200 import ref_transform as RT
    clf, d = _fitted(tmp_path, "tiny", 2)
    X = _boundary_table(clf, d["X_test"]).copy()
    X[3, -1] = np.inf                                    # sklearn's validation raises on infinities in a numeric column
    for pre in clf.executor_.preprocessors:
        fast = RT.compile_preprocessor(pre)
        try:
            pre.transform(X)
            raised = None
        except Exception as exc:
            raised = type(exc)
        if raised is None:
            assert np.array_equal(pre.transform(X).X, fast(X), equal_nan=True)
        else:
            with pytest.raises(raised):
                fast(X)
        assert RT.verify(pre, fast, X)
    wide = np.concatenate([X[:, :-1], X[:, :2]], axis=1)[:, : X.shape[1] + 1]
    with pytest.raises(Exception):
        RT.compile_preprocessor(clf.executor_.preprocessors[0])(wide)


def test_unknown_steps_fall_back(tmp_path):
    """The regressor's default recipes contain a power transform the replay does not reproduce: ``Unsupported``, and
    the plug-in engine keeps the reference's transforms."""
    from multimodalpfn_b200 import ref_transform as RT
    reg, _ = _fitted(tmp_path, "tiny", 4, regression=True)
    outcomes = []
    for pre in reg.executor_.preprocessors:
        try:
            RT.compile_preprocessor(pre)
            outcomes.append("compiled")
        except RT.Unsupported as exc:
            outcomes.append(str(exc))
    print(outcomes)
    assert any(o != "compiled" for o in outcomes)
