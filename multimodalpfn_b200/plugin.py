"""Plug-in mode: keep the reference's own ``MMPFNClassifier`` in charge (validation, ordinal
encoding, ``EnsembleConfig`` generation, numpy/sklearn preprocessing) and swap only the object its
engine calls at ``inference.py:343-348`` for ``B200PerFeatureTransformer``.

    import multimodalpfn_b200.plugin as plugin
    plugin.install(precision="bf16")          # after `mmpfn` is importable
    clf = mmpfn.models.mmpfn.MMPFNClassifier(..., device="cuda").fit(X, img, y)
    clf.predict_proba(X_te, img_te)           # forward runs on libmmpfn_b200.so

The patch wraps ``create_inference_engine`` (``base.py:168-257``, called from
``classifier.py:485-500``): the engine the reference builds keeps its per-estimator preprocessors
and tables; its ``.model`` attribute is replaced.  Nothing under the reference tree is edited.
Needs the reference and an sm_100 GPU in the same process (neither CI box here has both: the GPU
box has no reference; this container has no GPU), so it is exercised by ``tests/test_plugin_wiring.py``
with a recording stand-in for the CUDA model.
"""
from __future__ import annotations

import functools

from .synth import Geometry

_MIXERS = {"MGM": "MGM", "MGM+CAP": "MGM+CAP", "MoE": "MoE"}


def geometry_of(module) -> Geometry:
    """Read the geometry off a reference ``PerFeatureTransformer`` (``model/transformer.py:292-409``)."""
    mixer = getattr(module, "mixer_type", "MGM+CAP")
    mgm_heads = len(module.mgm.projs) if hasattr(module, "mgm") else (
        len(module.moe.experts) if hasattr(module, "moe") else 1)
    cap_heads = module.cap.queries.shape[0] if hasattr(module, "cap") else None
    dec = module.decoder_dict["standard"]
    return Geometry(emsize=module.ninp, nhead=module.nhead, nhid_factor=module.nhid // module.ninp,
                    nlayers=len(module.transformer_encoder.layers), n_out=dec[2].weight.shape[0],
                    features_per_group=module.features_per_group, img_dim=module.nhid, mgm_heads=mgm_heads,
                    cap_heads=cap_heads, mixer_type=_MIXERS[mixer])


def outlier_std_of(module):
    """``update_encoder_outlier_params`` (``utils.py:703-745``) stored it on the normalisation step."""
    for step in module.encoder:
        if "InputNormalizationEncoderStep" in str(step.__class__):
            return float(step.remove_outliers_sigma) if step.remove_outliers else None
    return None


def convert(module, *, device, precision="bf16", model_cls=None):
    """reference nn.Module -> B200PerFeatureTransformer with the same weights and seed."""
    if model_cls is None:
        from .model import B200PerFeatureTransformer as model_cls
    sd = {k: v.detach().cpu() for k, v in module.state_dict().items()}
    return model_cls(sd, geometry_of(module), device=device, precision=precision, seed=module.seed,
                     outlier_std=outlier_std_of(module),
                     pos_emb_device="cuda")     # the reference draws on the model's device (transformer.py:887-892)


def install(precision: str = "bf16", model_cls=None):
    """Patch ``mmpfn.models.mmpfn.classifier.create_inference_engine``; returns an ``uninstall()``."""
    import mmpfn.models.mmpfn.classifier as C

    orig = C.create_inference_engine
    if getattr(orig, "_mmpfn_b200", False):
        return lambda: None

    @functools.wraps(orig)
    def create_inference_engine(**kw):
        engine = orig(**kw)
        dev = kw["device_"]
        if getattr(dev, "type", str(dev)) == "cuda" and hasattr(engine, "model"):
            engine.model = convert(engine.model, device=dev, precision=precision, model_cls=model_cls)
        return engine

    create_inference_engine._mmpfn_b200 = True
    C.create_inference_engine = create_inference_engine

    def uninstall():
        C.create_inference_engine = orig
    return uninstall
