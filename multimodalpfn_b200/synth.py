"""Synthetic checkpoints and datasets for the MMPFN in-context inference path.

There is no network for real checkpoints or datasets, so every test, the bench and the
golden-vector generator draw from here.  Everything is `np.random.default_rng(seed)`
(bit-stable across numpy versions and machines), never torch RNG.

* ``make_state_dict`` — a state_dict with the reference's parameter names and shapes
  (reference ``model/loading.py:470-538``,
  ``model/transformer.py:33-88,392-409``).  The tensors the reference zero-initialises
  (``_w_out``: ``model/multi_head_attention.py:204-205``; ``mlp.linear2``:
  ``model/mlp.py:88-89``) are drawn from N(0, 0.05^2) so that the 12 layers are not
  identities (SURVEY.md gotcha 2).  QKV weights follow the reference's own init law
  (``model/multi_head_attention.py:149-162``): uniform with std sqrt(2/(nhead*d+E)).
* ``make_dataset`` — the shapes BASELINE.json's configs name.
"""
from __future__ import annotations

import dataclasses
import math

import numpy as np

__all__ = ["Geometry", "make_state_dict", "make_checkpoint_config", "make_dataset", "DATASETS"]


@dataclasses.dataclass(frozen=True)
class Geometry:
    """TabPFN-v2 classifier geometry (reference ``model/config.py:18-83``)."""

    emsize: int = 192
    nhead: int = 6
    nhid_factor: int = 4
    nlayers: int = 12
    n_out: int = 10
    features_per_group: int = 2
    img_dim: int = 768
    mgm_heads: int = 8
    cap_heads: int = 8
    mixer_type: str = "MGM+CAP"

    @property
    def d_k(self) -> int:
        return self.emsize // self.nhead

    @property
    def nhid(self) -> int:
        return self.emsize * self.nhid_factor


def _uniform(rng, shape, bound):
    return rng.uniform(-bound, bound, size=shape).astype(np.float32)


def _normal(rng, shape, std, mean=0.0):
    return (mean + std * rng.standard_normal(size=shape)).astype(np.float32)


def make_state_dict(geom: Geometry = Geometry(), seed: int = 1, *, residual_std: float = 0.05,
                    qkv_gain: float = 1.0, decoder_gain: float = 1.0,
                    two_sets_of_queries: bool = False, regression: bool = False) -> dict[str, np.ndarray]:
    """Random weights under the reference's state_dict names (numpy float32).

    ``residual_std`` scales ``_w_out`` / ``linear2`` (0.05 = SURVEY.md section 8(d) weights,
    logits within about +-0.5); ``decoder_gain`` and a larger ``residual_std`` give the
    "trained-like" confident regime of SURVEY.md gotcha 9.
    """
    rng = np.random.default_rng(seed)
    E, H, D, Hid = geom.emsize, geom.nhead, geom.d_k, geom.nhid
    I = geom.img_dim
    sd: dict[str, np.ndarray] = {}
    if geom.mixer_type in ("MGM", "MGM+CAP"):
        for h in range(geom.mgm_heads):
            p = f"mgm.projs.{h}."
            sd[p + "0.weight"] = _normal(rng, (I,), 0.1, 1.0)
            sd[p + "0.bias"] = _normal(rng, (I,), 0.1)
            sd[p + "1.weight"] = _uniform(rng, (I, I), 1 / math.sqrt(I))
            sd[p + "1.bias"] = _uniform(rng, (I,), 1 / math.sqrt(I))
            sd[p + "4.weight"] = _uniform(rng, (E, I // 2), 1 / math.sqrt(I // 2))
            sd[p + "4.bias"] = _uniform(rng, (E,), 1 / math.sqrt(I // 2))
    if geom.mixer_type == "MGM+CAP":
        C = geom.cap_heads
        sd["cap.queries"] = _normal(rng, (C, E), 1e-2)
        sd["cap.q_proj.weight"] = _uniform(rng, (E, E), 1 / math.sqrt(E))
        for n in ("k_norm", "q_norm", "out_norm"):
            sd[f"cap.{n}.weight"] = _normal(rng, (E,), 0.1, 1.0)
            sd[f"cap.{n}.bias"] = _normal(rng, (E,), 0.1)
        sd["cap.mha.in_proj_weight"] = _uniform(rng, (3 * E, E), math.sqrt(6 / (4 * E)))
        sd["cap.mha.in_proj_bias"] = _normal(rng, (3 * E,), 0.02)
        sd["cap.mha.out_proj.weight"] = _uniform(rng, (E, E), 1 / math.sqrt(E))
        sd["cap.mha.out_proj.bias"] = _normal(rng, (E,), 0.02)
        sd["cap.ffn.0.weight"] = _uniform(rng, (2 * E, E), 1 / math.sqrt(E))
        sd["cap.ffn.0.bias"] = _uniform(rng, (2 * E,), 1 / math.sqrt(E))
        sd["cap.ffn.3.weight"] = _uniform(rng, (E, 2 * E), 1 / math.sqrt(2 * E))
        sd["cap.ffn.3.bias"] = _uniform(rng, (E,), 1 / math.sqrt(2 * E))
    if geom.mixer_type == "MoE":
        for h in range(geom.mgm_heads):
            p = f"moe.experts.{h}."
            sd[p + "0.weight"] = _normal(rng, (I,), 0.1, 1.0)
            sd[p + "0.bias"] = _normal(rng, (I,), 0.1)
            sd[p + "1.weight"] = _uniform(rng, (I // 2, I), 1 / math.sqrt(I))
            sd[p + "1.bias"] = _uniform(rng, (I // 2,), 1 / math.sqrt(I))
            sd[p + "4.weight"] = _uniform(rng, (E, I // 2), 1 / math.sqrt(I // 2))
            sd[p + "4.bias"] = _uniform(rng, (E,), 1 / math.sqrt(I // 2))
        sd["moe.gate.weight"] = _uniform(rng, (geom.mgm_heads, I), 1 / math.sqrt(I))
        sd["moe.gate.bias"] = _uniform(rng, (geom.mgm_heads,), 1 / math.sqrt(I))
    fpg = geom.features_per_group
    sd["encoder.5.layer.weight"] = _uniform(rng, (E, 2 * fpg), 1 / math.sqrt(2 * fpg))
    # the regression y-encoder has no class-rank step (model/loading.py:374-398: NaN handling, Linear), so its
    # Linear is step 1 instead of 2; the checkpoint also carries the bar-distribution borders of the criterion
    ykey = "y_encoder.1.layer." if regression else "y_encoder.2.layer."
    sd[ykey + "weight"] = _uniform(rng, (E, 2), 1 / math.sqrt(2))
    sd[ykey + "bias"] = _uniform(rng, (E,), 1 / math.sqrt(2))
    if regression:
        # borders of geom.n_out buckets over a standardised target: normal quantiles (regressor.py:390-540 rescales them)
        qs = np.linspace(0.0, 1.0, geom.n_out + 1)[1:-1]
        from statistics import NormalDist
        inner = np.array([NormalDist().inv_cdf(float(q)) for q in qs])
        sd["criterion.borders"] = np.concatenate([[inner[0] - 1.0], inner, [inner[-1] + 1.0]]).astype(np.float32)
        sd["criterion.losses_per_bucket"] = np.zeros(geom.n_out, dtype=np.float32)      # bar_distribution.py:460-461
    a_qkv = math.sqrt(3.0) * math.sqrt(2.0 / (H * D + E)) * qkv_gain
    for l in range(geom.nlayers):
        p = f"transformer_encoder.layers.{l}."
        for att in ("self_attn_between_features", "self_attn_between_items"):
            if two_sets_of_queries and att == "self_attn_between_items":
                # multi_head_attention.py:216-260: a second query set for the test rows replaces the fused tensor
                sd[p + att + "._w_q"] = _uniform(rng, (2, H, D, E), a_qkv)
                sd[p + att + "._w_kv"] = _uniform(rng, (2, H, D, E), a_qkv)
            else:
                sd[p + att + "._w_qkv"] = _uniform(rng, (3, H, D, E), a_qkv)
            sd[p + att + "._w_out"] = _normal(rng, (H, D, E), residual_std)
        sd[p + "mlp.linear1.weight"] = _uniform(rng, (Hid, E), 1 / math.sqrt(E))
        sd[p + "mlp.linear2.weight"] = _normal(rng, (E, Hid), residual_std)
    sd["decoder_dict.standard.0.weight"] = _uniform(rng, (Hid, E), 1 / math.sqrt(E))
    sd["decoder_dict.standard.0.bias"] = _uniform(rng, (Hid,), 1 / math.sqrt(E))
    sd["decoder_dict.standard.2.weight"] = _uniform(rng, (geom.n_out, Hid), decoder_gain / math.sqrt(Hid))
    sd["decoder_dict.standard.2.bias"] = _uniform(rng, (geom.n_out,), decoder_gain / math.sqrt(Hid))
    sd["feature_positional_embedding_embeddings.weight"] = _uniform(rng, (E, E // 4), 1 / math.sqrt(E // 4))
    sd["feature_positional_embedding_embeddings.bias"] = _uniform(rng, (E,), 1 / math.sqrt(E // 4))
    return sd


def make_checkpoint_config(geom: Geometry = Geometry(), *, two_sets_of_queries: bool = False,
                           regression: bool = False) -> dict:
    """The minimal ``config`` dict the reference's loader accepts (SURVEY.md Appendix B;
    reference ``model/loading.py:253-305``, ``model/config.py:18-108``)."""
    return dict(
        adaptive_max_seq_len_to_max_full_table_size=150000, batch_size=4, emsize=geom.emsize,
        features_per_group=geom.features_per_group, max_num_classes=0 if regression else geom.n_out, nhead=geom.nhead,
        remove_duplicate_features=False, seq_len=4000, task_type="regression" if regression else "multiclass",
        num_buckets=geom.n_out if regression else 1000,
        max_num_features=85, aggregate_k_gradients=1, nlayers=geom.nlayers,
        nhid_factor=geom.nhid_factor, **({"two_sets_of_queries": True} if two_sets_of_queries else {}),
    )


# ------------------------------------------------------------------------------------------
# datasets
# ------------------------------------------------------------------------------------------
DATASETS = {
    # name: (n_train, n_test, n_tok, n_classes)
    "tiny": (96, 40, 1, 3),
    "pad_ufes_small": (400, 120, 1, 6),
    "pad_ufes": (2000, 300, 1, 6),
    "img_text_10k": (10_000, 10_000, 2, 10),
    "large_ctx_50k": (50_000, 50_000, 0, 10),
    "small_task": (800, 200, 1, 4),
}


def _pad_like_table(rng, n):
    """21 PAD-UFES-20-like columns: 14 {0,1,2} ints, 4 categoricals (3/3/2/14 levels),
    3 numerics with 2 % NaN."""
    cols = [rng.integers(0, 3, size=n).astype(np.float32) for _ in range(14)]
    for levels in (3, 3, 2, 14):
        cols.append(rng.integers(0, levels, size=n).astype(np.float32))
    age = rng.normal(60, 15, size=n)
    d1 = rng.lognormal(1.5, 0.6, size=n)
    d2 = rng.lognormal(1.2, 0.7, size=n)
    for c in (age, d1, d2):
        c = c.astype(np.float32)
        c[rng.random(n) < 0.02] = np.nan
        cols.append(c)
    return np.stack(cols, axis=1)


def make_dataset(name: str, seed: int = 0, *, img_dim: int = 768):
    """Returns dict(X_train, y_train, img_train, X_test, y_test, img_test) of numpy arrays.

    ``img_*`` is ``[N, n_tok, img_dim]`` float32 (``None`` when the config has no image).
    Labels depend weakly on the features and the embedding so that the task is learnable.
    """
    n_tr, n_te, n_tok, n_cls = DATASETS[name]
    rng = np.random.default_rng(seed)
    n = n_tr + n_te
    if name in ("pad_ufes", "pad_ufes_small", "tiny"):
        X = _pad_like_table(rng, n)
        prior = np.array([0.37, 0.32, 0.10, 0.10, 0.08, 0.03])[:n_cls]
    elif name == "img_text_10k":
        X = rng.standard_normal((n, 64)).astype(np.float32)
        for j in range(8):
            X[:, j] = rng.integers(0, 5 + j, size=n)
        prior = np.full(n_cls, 1.0 / n_cls)
    elif name == "large_ctx_50k":
        X = rng.standard_normal((n, 100)).astype(np.float32)
        prior = np.full(n_cls, 1.0 / n_cls)
    elif name == "small_task":
        X = rng.standard_normal((n, 32)).astype(np.float32)
        prior = np.full(n_cls, 1.0 / n_cls)
    else:
        raise KeyError(name)
    prior = prior / prior.sum()
    img = rng.standard_normal((n, n_tok, img_dim)).astype(np.float32) if n_tok > 0 else None
    # class scores: prior + a linear readout of a few columns (+ the embedding)
    Xz = np.nan_to_num(X, nan=0.0)
    Xz = (Xz - Xz.mean(0)) / (Xz.std(0) + 1e-6)
    W = rng.standard_normal((X.shape[1], n_cls)) * 0.6
    score = Xz @ W + np.log(prior)[None]
    if img is not None:
        Wi = rng.standard_normal((img_dim, n_cls)) * (0.8 / math.sqrt(img_dim))
        score = score + img[:, 0] @ Wi
    score = score + rng.gumbel(size=score.shape)
    y = score.argmax(1).astype(np.int64)
    # every class must appear in the training split
    for c in range(n_cls):
        if not (y[:n_tr] == c).any():
            y[c] = c
    return dict(
        X_train=X[:n_tr], y_train=y[:n_tr], img_train=None if img is None else img[:n_tr],
        X_test=X[n_tr:], y_test=y[n_tr:], img_test=None if img is None else img[n_tr:],
        n_classes=n_cls,
    )
