"""Shared builders for the golden cases: regenerate the seeded inputs/weights a fixture was
made from (``oracle/make_golden.py``) without touching the reference."""
import os

import numpy as np
import torch

from multimodalpfn_b200.synth import Geometry, make_dataset, make_state_dict
from oracle.make_golden import CLF_CASES, MODEL_CASES, mutate_inputs

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def model_case(name):
    """-> geom, state_dict(np), X_full(np|None), img_full(np|None), y_train(np f32), n_train"""
    gkw, ds, wseed, mut = MODEL_CASES[name]
    geom = Geometry(**{k: v for k, v in gkw.items() if v is not None or k == "cap_heads"})
    extra = dict(residual_std=0.2, decoder_gain=20.0) if mut == "stress" else {}
    sd = make_state_dict(geom, seed=wseed, **extra)
    d = make_dataset(ds, 0)
    X = np.concatenate([d["X_train"], d["X_test"]])
    img = np.concatenate([d["img_train"], d["img_test"]])
    X, img = mutate_inputs(mut, X, img)
    y = d["y_train"].astype(np.float32)
    return geom, sd, X, img, y, len(y)


def clf_case(name):
    ds, n_est, kw = CLF_CASES[name]
    geom = Geometry(mgm_heads=2, cap_heads=4)
    sd = make_state_dict(geom, seed=11)
    d = make_dataset(ds, 0)
    return geom, sd, d, n_est, kw


def t(x):
    return None if x is None else torch.as_tensor(x)
