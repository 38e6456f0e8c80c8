"""ORACLE support: generate ``tests/golden/*.npz`` by running the REAL reference.

Run here (the reference is absent on the GPU box):

    PYTHONHASHSEED=0 python -m oracle.make_golden

Two kinds of fixture, both from ``/root/reference`` code on seeded synthetic inputs
(``multimodalpfn_b200.synth``; weights are regenerated from the seed at test time, never stored):

* ``model_<case>.npz`` — the reference ``PerFeatureTransformer`` called exactly like
  ``inference.py:343-348`` does: logits, a few rows of the token state after the stem and
  after each layer, and the reference's own per-layer head-0 K/V cache for the cached path.
* ``clf_<case>.npz`` — the reference ``MMPFNClassifier.fit / predict_proba``
  (``classifier.py:364-576``) with its default preprocessing: the tensors that crossed the
  model boundary for each estimator (``X_full``, ``y_train``), the class permutations, the
  per-estimator logits and the final probabilities.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from multimodalpfn_b200.synth import (Geometry, make_checkpoint_config, make_dataset,  # noqa: E402
                                     make_state_dict)
from oracle import ref_compat  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# case name -> (geometry kwargs, dataset, weight seed, input mutation)
MODEL_CASES = {
    "mgmcap_tiny": (dict(mgm_heads=2, cap_heads=4), "tiny", 1, None),
    "mgmcap_8x8_small": (dict(mgm_heads=8, cap_heads=8), "pad_ufes_small", 1, None),
    "mgm_only_tiny": (dict(mgm_heads=3, cap_heads=None, mixer_type="MGM"), "tiny", 2, None),
    "moe_tiny": (dict(mgm_heads=4, cap_heads=2, mixer_type="MoE"), "tiny", 3, None),
    "noimage_tiny": (dict(mgm_heads=2, cap_heads=4), "tiny", 4, "noimage"),
    "imageonly_tiny": (dict(mgm_heads=2, cap_heads=4), "tiny", 5, "imageonly"),
    "edge_tiny": (dict(mgm_heads=2, cap_heads=4), "tiny", 6, "edge"),
    "stress_tiny": (dict(mgm_heads=2, cap_heads=4), "tiny", 7, "stress"),
    # a checkpoint with two_sets_of_queries (multi_head_attention.py:216-260): _w_q [2,..] + _w_kv for the item attention
    "twosets_tiny": (dict(mgm_heads=2, cap_heads=4), "tiny", 8, "twosets"),
    # the head counts of the authors' best PAD-UFES cell is (256, 24) (mmpfn/charts/pad_ufes_20.csv:36); (64, 24) is the
    # geometry of their fine-tuning log (50.2 M parameters, mmpfn/logs/finetune_tabpfn.log:6): CAP head_dim = 8,
    # 64 MGM tokens per embedding
    "mgm64_cap24_tiny": (dict(mgm_heads=64, cap_heads=24), "tiny", 10, None),
    # a regression checkpoint (max_num_classes = 0): y-encoder without the class-rank step, 64-bucket decoder
    "regression_tiny": (dict(mgm_heads=2, cap_heads=4, n_out=64), "tiny", 9, "regression"),
}


def case_state_dict(geom, wseed, mut):
    """The seeded synthetic weights + checkpoint config of a model case."""
    extra = dict(residual_std=0.2, decoder_gain=20.0) if mut == "stress" else {}
    if mut == "twosets":
        extra = dict(two_sets_of_queries=True)
    if mut == "regression":
        extra = dict(regression=True)
    sd = make_state_dict(geom, seed=wseed, **extra)
    cfg = make_checkpoint_config(geom, two_sets_of_queries=(mut == "twosets"), regression=(mut == "regression"))
    return sd, cfg


def case_targets(mut, d, X):
    """Training targets of a model case: class ids, or (regression) a standardised continuous target."""
    if mut == "regression":
        rng = np.random.default_rng(77)
        n = len(d["y_train"])
        y = d["X_train"][:, 0] * 0.7 + d["y_train"] * 0.5 + 0.3 * rng.standard_normal(n)
        return ((y - y.mean()) / y.std()).astype(np.float32)
    return d["y_train"].astype(np.float32)


def mutate_inputs(kind, X, img):
    """Edge cases the stem must survive (encoders.py:461-491, :515, :615): constant columns,
    a heavy-NaN column, +-inf cells, heavy outliers, odd feature count."""
    if kind == "noimage":
        return X, None
    if kind == "imageonly":
        return None, img
    if kind == "edge":
        X = X.copy()
        X = np.concatenate([X, X[:, :2] * 0 + 3.25], axis=1)        # two constant columns -> F'=23 (odd)
        X[:, 1] = 7.0                                               # constant inside a mixed group
        X[100::7, 18] = np.inf                                      # inf only in test rows: in train rows the
        X[103::11, 19] = -np.inf                                    # reference itself raises (nanmean keeps inf)
        X[5, 20] = 1e6                                              # outlier beyond 12 sigma
        X[::3, 5] = np.nan                                          # heavy-NaN column
        X[:, 4] = 2.0
        X[100:, 4] = 3.0                                            # constant in train, varies in test
        return X, img
    return X, img


def _snap(state):
    """Keep 2 train + 2 test rows of a [1,S,T,E] state (fixtures stay small)."""
    s = state[0]
    return torch.cat([s[:2], s[-2:]], 0).numpy().copy()


def run_model_case(name):
    gkw, ds, wseed, mut = MODEL_CASES[name]
    geom = Geometry(**{k: v for k, v in gkw.items() if v is not None or k == "cap_heads"})
    sd, cfg = case_state_dict(geom, wseed, mut)
    model, _ = ref_compat.load_reference_model(
        sd, cfg, mixer_type=geom.mixer_type, mgm_heads=geom.mgm_heads,
        cap_heads=geom.cap_heads, features_per_group=geom.features_per_group, model_seed=0)
    d = make_dataset(ds, 0)
    X = np.concatenate([d["X_train"], d["X_test"]])
    img = np.concatenate([d["img_train"], d["img_test"]])
    X, img = mutate_inputs(mut, X, img)
    y = case_targets(mut, d, X)
    n_tr = len(y)

    snaps = {}
    hooks = []
    enc = model.transformer_encoder
    hooks.append(enc.register_forward_pre_hook(
        lambda m, a: snaps.__setitem__("state_stem", _snap(a[0]))))
    for li, layer in enumerate(enc.layers):
        if li not in (0, len(enc.layers) - 1):
            continue
        hooks.append(layer.register_forward_hook(
            lambda m, a, o, li=li: snaps.__setitem__(f"state_l{li}", _snap(o))))

    def call(x, im, yy, sep):
        with torch.inference_mode():
            return model(None, None if x is None else torch.tensor(x)[:, None],
                         None if im is None else torch.tensor(im),
                         None if yy is None else torch.tensor(yy),
                         only_return_standard_out=True, categorical_inds=[],
                         single_eval_pos=sep)

    model.cache_trainset_representation = False
    logits = call(X, img, y, n_tr).squeeze(1).numpy()
    for h in hooks:
        h.remove()
    out = dict(logits=logits, **snaps)

    # the reference's own cached path
    if X is not None:   # the reference's cached call needs x for its device lookup (transformer.py:616)
        model.cache_trainset_representation = True
        call(X[:n_tr], None if img is None else img[:n_tr], y, n_tr)
        kv = [l.self_attn_between_items._kv_cache.squeeze(-2).numpy().copy() for l in enc.layers]
        out["kv_l0"] = kv[0][:, :8]            # [T, first 8 train rows, 2, 32]
        out["kv_l11"] = kv[-1][:, :8]
        out["logits_cached"] = call(X[n_tr:], None if img is None else img[n_tr:],
                                    None, None).squeeze(1).numpy()
        model.empty_trainset_representation_cache()
    np.savez_compressed(os.path.join(OUT, f"model_{name}.npz"), **out)
    dc = np.abs(out["logits_cached"] - logits).max() if "logits_cached" in out else float("nan")
    print(f"model_{name}: logits {logits.shape} |joint-cached| {dc:.2e}")


CLF_CASES = {
    # name: (dataset, n_estimators, clf kwargs)
    "default_tiny": ("tiny", 4, {}),
    "avg_before_softmax_tiny": ("tiny", 2, dict(average_before_softmax=True, softmax_temperature=1.0,
                                               balance_probabilities=True)),
}


def run_clf_case(name):
    ds, n_est, kw = CLF_CASES[name]
    geom = Geometry(mgm_heads=2, cap_heads=4)
    sd = make_state_dict(geom, seed=11)
    ref_compat.install()
    path = os.path.join("/tmp", f"mmpfn_b200_clf_{os.getpid()}.ckpt")
    torch.save({"state_dict": {k: torch.as_tensor(v) for k, v in sd.items()},
                "config": make_checkpoint_config(geom)}, path)
    from mmpfn.models.mmpfn import MMPFNClassifier
    from mmpfn.models.mmpfn.model.transformer import PerFeatureTransformer

    d = make_dataset(ds, 0)
    clf = MMPFNClassifier(mixer_type="MGM+CAP", mgm_heads=2, cap_heads=4, features_per_group=2,
                          n_estimators=n_est, model_path=path, device="cpu",
                          ignore_pretraining_limits=True, random_state=0, **kw)
    clf.fit(d["X_train"], d["img_train"], d["y_train"])
    rec = []
    orig = PerFeatureTransformer._forward

    def spy(self, x, image, y, **k):
        xin = x.clone()
        yin = y.clone()
        o = orig(self, x, image, y, **k)
        rec.append((xin[:, 0].numpy().copy(), yin.numpy().copy(), o.squeeze(1).numpy().copy()))
        return o

    PerFeatureTransformer._forward = spy
    try:
        proba = clf.predict_proba(d["X_test"], d["img_test"])
    finally:
        PerFeatureTransformer._forward = orig
    out = dict(proba=proba, n_estimators=n_est, n_classes=clf.n_classes_,
               class_counts=clf.class_counts_)
    for e, ((x, y, lg), cfg) in enumerate(zip(rec, clf.executor_.ensemble_configs)):
        out[f"X_full_{e}"] = x.astype(np.float32)
        out[f"y_train_{e}"] = y.astype(np.float32)
        out[f"logits_{e}"] = lg
        out[f"class_perm_{e}"] = (np.arange(clf.n_classes_) if cfg.class_permutation is None
                                  else np.asarray(cfg.class_permutation))
    np.savez_compressed(os.path.join(OUT, f"clf_{name}.npz"), **out)
    print(f"clf_{name}: proba {proba.shape}, F' per estimator "
          f"{[r[0].shape[1] for r in rec]}")


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    assert os.environ.get("PYTHONHASHSEED") == "0", "run with PYTHONHASHSEED=0 (SURVEY.md gotcha 3)"
    only = sys.argv[1:]
    for name in MODEL_CASES:
        if not only or name in only:
            run_model_case(name)
    for name in CLF_CASES:
        if not only or name in only:
            run_clf_case(name)


if __name__ == "__main__":
    main()
