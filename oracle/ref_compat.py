"""ORACLE support (test infrastructure): make the reference importable in this container.

The reference (``/root/reference``, read-only, absent on the GPU box) imports plotting packages
it never uses on the hot path (``model/transformer.py:19-20``) and calls scikit-learn APIs that
newer releases renamed (``utils.py:484-495, 538-544, 557-567``).  This shim is what SURVEY.md
Appendix C describes; it edits nothing under ``/root/reference``.

Used only by ``oracle/make_golden.py`` and by tests that are skipped when the reference is
absent.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("MMPFN_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "mmpfn", "models", "mmpfn"))


def install() -> None:
    """Idempotent: stub missing plotting modules, patch sklearn renames, extend sys.path."""
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    for name in ("matplotlib", "matplotlib.pyplot", "seaborn"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    if "matplotlib" in sys.modules and not hasattr(sys.modules["matplotlib"], "pyplot"):
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]

    import sklearn.base
    import sklearn.utils.validation as skv

    def _rename(kw):
        kw = dict(kw)
        if "force_all_finite" in kw:
            kw["ensure_all_finite"] = kw.pop("force_all_finite")
        kw.pop("estimator", None)
        return kw

    if not hasattr(sklearn.base.BaseEstimator, "_validate_data"):
        def _validate_data(self, X="no_validation", y="no_validation", reset=True, **kw):
            return skv.validate_data(self, X, y, reset=reset, **_rename(kw))
        sklearn.base.BaseEstimator._validate_data = _validate_data

    if not getattr(sklearn.base, "_mmpfn_b200_patched", False):
        _orig_check_array = skv.check_array

        def check_array(*a, **kw):
            kw = dict(kw)
            if "force_all_finite" in kw:
                kw["ensure_all_finite"] = kw.pop("force_all_finite")
            return _orig_check_array(*a, **kw)

        sklearn.base.check_array = check_array
        sklearn.base._mmpfn_b200_patched = True

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def load_reference_model(state_dict, config, *, mixer_type="MGM+CAP", mgm_heads=8, cap_heads=8,
                         features_per_group=2, model_seed=0, outlier_std=12.0, tmp_dir="/tmp"):
    """Write a ``{"state_dict","config"}`` checkpoint (the reference's format,
    ``model/loading.py:427-444``) and load it with the reference's own loader; returns the
    reference ``PerFeatureTransformer`` in eval mode with every parameter taken from
    ``state_dict`` (asserted) and the outlier step switched on as ``fit`` does
    (``classifier.py:396-406``)."""
    import torch
    install()
    from mmpfn.models.mmpfn.model.loading import load_model
    from mmpfn.models.mmpfn.utils import update_encoder_outlier_params

    path = os.path.join(tmp_dir, f"mmpfn_b200_oracle_{os.getpid()}.ckpt")
    sd = {k: torch.as_tensor(v) for k, v in state_dict.items()}
    torch.save({"state_dict": sd, "config": dict(config)}, path)
    try:
        model, _, cfg = load_model(path=path, model_seed=model_seed, mixer_type=mixer_type,
                                   mgm_heads=mgm_heads, cap_heads=cap_heads,
                                   features_per_group=features_per_group)
    finally:
        pass
    have = {k for k, _ in model.named_parameters()}
    missing = have - set(sd)
    assert not missing, f"synthetic state_dict does not cover: {sorted(missing)[:8]}"
    if outlier_std is not None:
        update_encoder_outlier_params(model, outlier_std, model_seed, inplace=True)
    return model, path
