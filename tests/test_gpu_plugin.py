"""The literal drop-in on hardware: the reference's OWN ``MMPFNClassifier`` (validation, ordinal encoding,
``EnsembleConfig`` generation, numpy/sklearn preprocessing, probability tail — all the reference's code,
``classifier.py:364-576``, ``inference.py:282-351``) with this repo's CUDA model swapped in by
``multimodalpfn_b200.plugin.install`` at ``base.py:168-257``, against

* the reference's own ``predict_proba`` on the CPU in fp32 (the anchor north_star names), and
* the reference's own ``predict_proba`` on the same GPU in fp32 (torch eager + SDPA),

in the same process, so both sides see the same fingerprint hash seed (SURVEY.md gotcha 3).  Needs the
unmodified reference snapshot under ``oracle/_ref`` (``python -m oracle.snapshot_ref``; it travels to the
GPU box with the working tree) or ``/root/reference``.

Tolerances (BASELINE.json north_star): max |dp| <= 1e-5 in fp32, <= 2e-3 in bf16."""
import numpy as np
import pytest
import torch

from oracle import ref_compat
from tests.cases import check_proba

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref_compat.reference_available(), reason="reference snapshot not present")]


def _checkpoint(tmp_path, geom, seed):
    from multimodalpfn_b200.synth import make_checkpoint_config, make_state_dict
    sd = make_state_dict(geom, seed=seed)
    path = str(tmp_path / "m.ckpt")
    torch.save({"state_dict": {k: torch.as_tensor(v) for k, v in sd.items()}, "config": make_checkpoint_config(geom)}, path)
    return path


@pytest.mark.parametrize("dataset,n_est,heads", [("tiny", 4, (2, 4)), ("pad_ufes_small", 8, (8, 8))])
def test_plugin_real_model_vs_reference(tmp_path, dataset, n_est, heads):
    from multimodalpfn_b200 import _lib, plugin
    from multimodalpfn_b200.model import B200PerFeatureTransformer
    from multimodalpfn_b200.synth import Geometry, make_dataset
    ref_compat.install()
    import mmpfn.models.mmpfn.classifier as C

    geom = Geometry(mgm_heads=heads[0], cap_heads=heads[1])
    path = _checkpoint(tmp_path, geom, seed=11)
    d = make_dataset(dataset, 0)
    kw = dict(mixer_type="MGM+CAP", mgm_heads=heads[0], cap_heads=heads[1], features_per_group=2, n_estimators=n_est,
              model_path=path, ignore_pretraining_limits=True, random_state=0)

    def run(**over):
        clf = C.MMPFNClassifier(**{**kw, **over}).fit(d["X_train"], d["img_train"], d["y_train"])
        return clf, clf.predict_proba(d["X_test"], d["img_test"])

    with ref_compat.without_diagnostic_loop():          # result-neutral (verified bit-exact); keeps the CPU run short
        _, ref_cpu = run(device="cpu")
        _, ref_gpu = run(device="cuda", inference_precision=torch.float32)
    print(f"reference CUDA fp32 vs reference CPU fp32: max|dp| {np.abs(ref_gpu - ref_cpu).max():.3e} "
          "(different positional-noise streams on the two devices, SURVEY.md section 7 probe 2)")

    for precision, mode, tol in (("fp32", "model", 1e-5), ("fp32", "engine", 1e-5), ("bf16", "engine", 2e-3),
                                 ("bf16", "model", 2e-3)):
        for pos_dev, ref, what in (("cpu", ref_cpu, "reference CPU fp32"), ("cuda", ref_gpu, "reference CUDA fp32")):
            l0 = _lib.launch_count()
            uninstall = plugin.install(precision=precision, pos_emb_device=pos_dev, mode=mode)
            try:
                clf, got = run(device="cuda")
            finally:
                uninstall()
            assert isinstance(clf.executor_.model, B200PerFeatureTransformer)      # the swap happened
            assert isinstance(clf.executor_, plugin.B200PluginEngine) == (mode == "engine")
            if mode == "engine":      # the fitted preprocessors were replayed, after the bit-for-bit check at the first table
                assert clf.executor_.replay_state == "on", clf.executor_.replay_note
            assert _lib.launch_count() > l0                                          # ... and our kernels ran
            assert got.shape == ref.shape and got.dtype == ref.dtype
            assert np.allclose(got.sum(1), 1.0, atol=1e-5)
            # the GPU reference itself is not bit-equal to the CPU one (cuBLAS / SDPA summation order): allow for it
            slack = 0.0 if pos_dev == "cpu" else 2e-6
            check_proba(got, ref, tol + slack, f"plug-in[{mode}] {dataset} {precision} vs {what}")
    assert C.create_inference_engine.__name__ == "create_inference_engine"
    assert not getattr(C.create_inference_engine, "_mmpfn_b200", False)              # uninstalled


def test_plugin_replay_equals_reference_transforms(tmp_path):
    """Engine mode with the fitted preprocessors replayed (one batched pass) against engine mode on the reference's own
    ``transform`` calls (pipelined sub-batches): the tables that reach the model are bit-identical and the kernels are
    batching-invariant, so the probabilities must be EQUAL — also on a second table (no verification any more) with
    NaNs and category codes the fit never saw."""
    from multimodalpfn_b200 import plugin
    from multimodalpfn_b200.synth import Geometry, make_dataset
    ref_compat.install()
    import mmpfn.models.mmpfn.classifier as C

    geom = Geometry(mgm_heads=2, cap_heads=4)
    path = _checkpoint(tmp_path, geom, seed=11)
    d = make_dataset("pad_ufes_small", 0)
    X2 = d["X_test"].copy()
    rng = np.random.default_rng(3)
    X2[rng.random(X2.shape) < 0.05] = np.nan
    X2[::7, 15] = 9.0                                  # an unseen level of a categorical column
    X2[::5, 20] *= 4.0                                 # beyond the fitted quantile range
    kw = dict(mixer_type="MGM+CAP", mgm_heads=2, cap_heads=4, features_per_group=2, n_estimators=8, model_path=path,
              ignore_pretraining_limits=True, random_state=0, device="cuda")
    out = {}
    for replay in (True, False):
        uninstall = plugin.install(precision="bf16", mode="engine", replay=replay)
        try:
            clf = C.MMPFNClassifier(**kw).fit(d["X_train"], d["img_train"], d["y_train"])
            out[replay] = (clf.predict_proba(d["X_test"], d["img_test"]), clf.predict_proba(X2, d["img_test"]))
        finally:
            uninstall()
        assert clf.executor_.replay_state == ("on" if replay else "off"), clf.executor_.replay_note
    assert np.array_equal(out[True][0], out[False][0]) and np.array_equal(out[True][1], out[False][1])


@pytest.mark.parametrize("mode", ["engine", "model"])
def test_plugin_regressor_vs_reference(tmp_path, mode):
    """the reference's own ``MMPFNRegressor`` (target transforms, border translation,
    bar-distribution mean / median / quantiles: regressor.py:390-730, model/bar_distribution.py — all its code) on a
    regression checkpoint (y-encoder without the class-rank step, 64-bucket decoder) with this repo's engine plugged
    in, against the reference's CPU fp32 ``predict``."""
    from multimodalpfn_b200 import plugin
    from multimodalpfn_b200.synth import Geometry, make_checkpoint_config, make_dataset, make_state_dict
    ref_compat.install()
    import mmpfn.models.mmpfn.regressor as R

    geom = Geometry(mgm_heads=2, cap_heads=4, n_out=64)
    sd = make_state_dict(geom, seed=9, regression=True)
    path = str(tmp_path / "r.ckpt")
    torch.save({"state_dict": {k: torch.as_tensor(v) for k, v in sd.items()},
                "config": make_checkpoint_config(geom, regression=True)}, path)
    d = make_dataset("tiny", 0)
    rng = np.random.default_rng(0)
    ytr = (d["X_train"][:, 0] * 0.7 + 0.3 * rng.standard_normal(len(d["X_train"]))).astype(np.float32)
    kw = dict(mixer_type="MGM+CAP", mgm_heads=2, cap_heads=4, features_per_group=2, n_estimators=4, model_path=path,
              ignore_pretraining_limits=True, random_state=0)

    def run(**over):
        reg = R.MMPFNRegressor(**{**kw, **over}).fit(d["X_train"], d["img_train"], ytr)
        return reg, reg.predict(d["X_test"], d["img_test"], output_type="full")

    with ref_compat.without_diagnostic_loop():
        _, ref = run(device="cpu")
    for precision, tol_logit, tol_mean in (("fp32", 2e-4, 2e-5), ("bf16", 0.05, 5e-3)):
        uninstall = plugin.install(precision=precision, pos_emb_device="cpu", mode=mode)
        try:
            reg, got = run(device="cuda")
        finally:
            uninstall()
        assert isinstance(reg.executor_, plugin.B200PluginEngine) == (mode == "engine")
        if mode == "engine":
            print(f"[plug-in regressor] preprocessor replay: {reg.executor_.replay_state} ({reg.executor_.replay_note})")
        dl = float((got["logits"] - ref["logits"]).abs().max())
        dm = float(np.abs(got["mean"] - ref["mean"]).max())
        dq = max(float(np.abs(a - b).max()) for a, b in zip(got["quantiles"], ref["quantiles"]))
        print(f"[plug-in[{mode}] regressor {precision}] max |d log-prob| {dl:.3e}, |d mean| {dm:.3e}, |d quantile| {dq:.3e} "
              f"(targets span {float(ytr.min()):.2f} .. {float(ytr.max()):.2f})")
        assert dl < tol_logit and dm < tol_mean, (precision, dl, dm)
