"""Plug-in mode wiring (multimodalpfn_b200/plugin.py) against the REAL reference classifier, with
a recording stand-in for the CUDA model (this container has the reference but no GPU).  Skipped
where the reference is absent (the GPU box)."""
import numpy as np
import pytest
import torch

from oracle import ref_compat

pytestmark = pytest.mark.skipif(not ref_compat.reference_available(), reason="reference not present")


def test_plugin_swaps_only_the_model(tmp_path, monkeypatch):
    from multimodalpfn_b200 import plugin
    from multimodalpfn_b200.synth import Geometry, make_checkpoint_config, make_dataset, make_state_dict
    from oracle import forward_ref as R
    ref_compat.install()
    import mmpfn.models.mmpfn.classifier as C

    geom = Geometry(mgm_heads=2, cap_heads=4)
    sd = make_state_dict(geom, seed=11)
    path = str(tmp_path / "m.ckpt")
    torch.save({"state_dict": {k: torch.as_tensor(v) for k, v in sd.items()},
                "config": make_checkpoint_config(geom)}, path)
    calls = []
    LayerShim = type("S", (), {"layers": [None] * 12})

    class Recorder:
        """Stands in for B200PerFeatureTransformer: same ctor/call contract, oracle arithmetic."""
        def __init__(self, state_dict, g, *, device, precision, seed, outlier_std, pos_emb_device):
            assert g == geom and outlier_std == 12.0 and seed == 0
            self.sd = R.as_torch_state_dict({k: v.numpy() for k, v in state_dict.items()})
            self.g = g

        def to(self, *a, **k): return self
        def type(self, *a, **k): return self
        def cpu(self): return self
        ninp, features_per_group = 192, 2
        transformer_encoder = LayerShim()
        def parameters(self): return iter(self.sd.values())
        def reset_save_peak_mem_factor(self, f=None): pass

        def __call__(self, style, x, image, y, *, only_return_standard_out, categorical_inds, single_eval_pos):
            calls.append((tuple(x.shape), tuple(image.shape), tuple(y.shape), single_eval_pos))
            lg = R.forward_joint(x[:, 0], image, y, self.sd, self.g, seed=0)
            return lg[:, None, :]

    d = make_dataset("tiny", 0)
    kw = dict(mixer_type="MGM+CAP", mgm_heads=2, cap_heads=4, features_per_group=2, n_estimators=2,
              model_path=path, device="cpu", ignore_pretraining_limits=True, random_state=0)
    ref = C.MMPFNClassifier(**kw).fit(d["X_train"], d["img_train"], d["y_train"]).predict_proba(d["X_test"], d["img_test"])

    # the patch only converts on CUDA devices; force the conversion for this CPU wiring test
    orig_convert = plugin.convert
    monkeypatch.setattr(plugin, "convert", lambda m, *, device, precision, model_cls, pos_emb_device="cuda": orig_convert(
        m, device=device, precision=precision, model_cls=model_cls, pos_emb_device=pos_emb_device))
    uninstall = plugin.install(precision="fp32", model_cls=Recorder)
    try:
        clf = C.MMPFNClassifier(**kw).fit(d["X_train"], d["img_train"], d["y_train"])
        clf.executor_.model = plugin.convert(clf.executor_.model, device="cpu", precision="fp32", model_cls=Recorder)
        got = clf.predict_proba(d["X_test"], d["img_test"])
    finally:
        uninstall()
    assert len(calls) == 2 and calls[0][3] == len(d["y_train"])
    assert np.abs(got - ref).max() < 1e-5        # same probabilities through the swapped model
    assert C.create_inference_engine.__name__ == "create_inference_engine"
