"""Host-side ensemble preprocessing in front of the hot path.

Written from scratch; it plays the role of the reference's ``EnsembleConfig`` generation and
``fit_preprocessing`` (``preprocessing.py:228-335, 501-633``; transforms in
``model/preprocessing.py``) with the same structure — ``n_estimators`` members alternate between
two recipes, each with its own feature permutation and class permutation — but it is NOT a
bit-level replica of that CPU pipeline (which is out of the hot path and would be reused
unchanged in the plug-in integration).  Recipes:

* ``"none"``            — columns as they are (what the authors run: ``run.py:101-104``);
* ``"quantile_svd"``    — originals + uniform quantile transform of the numeric columns +
                          truncated-SVD components (the shape of the reference's first default
                          recipe, ``preprocessing.py:141-149``).

A deterministic row-fingerprint column is appended when ``fingerprint`` is on (the reference
uses Python ``hash`` — process dependent, SURVEY.md gotcha 3; here blake2b).
"""
from __future__ import annotations

import dataclasses
from typing import Optional

import numpy as np

__all__ = ["EnsembleMember", "RecipeCore", "make_members", "fit_transform_all", "transform_all", "RECIPES"]

RECIPES = ("quantile_svd", "none")


_FP_MULT = None


def _fingerprint(X: np.ndarray) -> np.ndarray:
    """Deterministic per-row fingerprint in [0, 1): a multiplicative hash of the row's float32 bit patterns
    (vectorised: one pass over the table; the reference hashes row by row with Python ``hash``, which is
    process dependent — SURVEY.md gotcha 3)."""
    global _FP_MULT
    Xc = np.ascontiguousarray(X, dtype=np.float32)
    w = Xc.view(np.uint32).astype(np.uint64)
    if _FP_MULT is None or len(_FP_MULT) < w.shape[1]:
        _FP_MULT = (np.random.default_rng(0x5EED).integers(1, 2**62, size=max(64, w.shape[1]), dtype=np.uint64) | np.uint64(1))
    with np.errstate(over="ignore"):
        h = (w * _FP_MULT[: w.shape[1]]).sum(axis=1, dtype=np.uint64)
        h ^= h >> np.uint64(33)
        h *= np.uint64(0xFF51AFD7ED558CCD)
        h ^= h >> np.uint64(33)
    return ((h >> np.uint64(40)).astype(np.float64) / float(1 << 24)).astype(np.float32)


def _quantile_uniform(Xn: np.ndarray, quantiles: np.ndarray, references: np.ndarray) -> np.ndarray:
    """Uniform-output quantile transform from fitted quantiles — the arithmetic of sklearn's
    ``QuantileTransformer._transform_col`` (forward/backward interpolation averaged, bounds clipped, NaN
    kept) without the estimator plumbing, which costs more than the interpolation at 300 rows."""
    out = np.empty(Xn.shape, dtype=np.float64)
    r = references
    for j in range(Xn.shape[1]):
        x = Xn[:, j].astype(np.float64)
        q = quantiles[:, j]
        ok = ~np.isnan(x)
        res = x.copy()
        res[ok] = 0.5 * (np.interp(x[ok], q, r) - np.interp(-x[ok], -q[::-1], -r[::-1]))
        res[x + 1e-7 > q[-1]] = 1.0
        res[x - 1e-7 < q[0]] = 0.0
        out[:, j] = res
    return out.astype(np.float32)


class RecipeCore:
    """The fitted, member-independent part of a recipe (quantile transformer, SVD basis).  Members of
    the same recipe differ only in their feature / class permutations, so they share one core: it is
    fitted once and, at predict time, applied once per call instead of once per member."""

    def __init__(self, recipe: str):
        self.recipe = recipe
        self.numeric_cols = None
        self._qt = None
        self._svd = None
        self._svd_mean = None
        self._svd_scale = None
        self._q = self._r = self._basis = None
        self.fitted = False

    def features(self, X: np.ndarray, fit: bool) -> np.ndarray:
        """Recipe columns for the raw table X (before fingerprint column and feature permutation)."""
        X = np.asarray(X, dtype=np.float32)
        parts = [X]
        if self.recipe == "quantile_svd":
            from sklearn.decomposition import TruncatedSVD
            from sklearn.preprocessing import QuantileTransformer
            n, F = X.shape
            if fit:
                nun = np.array([len(np.unique(X[~np.isnan(X[:, j]), j])) for j in range(F)])
                self.numeric_cols = np.where(nun > 30)[0]
            if len(self.numeric_cols):
                Xn = X[:, self.numeric_cols]
                if fit:
                    self._qt = QuantileTransformer(n_quantiles=max(min(n // 10, 1000), 2),
                                                   output_distribution="uniform", random_state=0)
                    self._qt.fit(Xn)
                    self._q = np.asarray(self._qt.quantiles_, dtype=np.float64)
                    self._r = np.asarray(self._qt.references_, dtype=np.float64)
                parts.append(_quantile_uniform(Xn, self._q, self._r))
            k = max(1, min(n // 10 + 1, F // 2)) if fit else self._svd.n_components
            Xz = np.nan_to_num(np.concatenate(parts, 1), nan=0.0, posinf=0.0, neginf=0.0)
            if fit:
                self._svd_mean = Xz.mean(0)
                self._svd_scale = Xz.std(0) + 1e-6
                self._svd = TruncatedSVD(n_components=k, algorithm="arpack", random_state=0)
                self._svd.fit((Xz - self._svd_mean) / self._svd_scale)
                self._basis = np.asarray(self._svd.components_, dtype=np.float64).T.copy()
            # TruncatedSVD.transform is X @ components_.T
            parts.append((((Xz - self._svd_mean) / self._svd_scale).astype(np.float64) @ self._basis).astype(np.float32))
        if fit:
            self.fitted = True
        return np.concatenate(parts, 1)


@dataclasses.dataclass
class EnsembleMember:
    recipe: str
    feature_shift_seed: Optional[int]      # None = keep the column order
    class_perm: Optional[np.ndarray]       # logits are gathered with it (classifier.py:550-551)
    fingerprint: bool
    feature_perm: Optional[np.ndarray] = None
    core: Optional[RecipeCore] = None      # shared by the members of a recipe (make_members)

    def __post_init__(self):
        if self.core is None:
            self.core = RecipeCore(self.recipe)

    def fit_transform(self, X: Optional[np.ndarray], y: np.ndarray, *, base=None, fp=None):
        """-> (X_train' or None, y_train permuted)."""
        y_out = y if self.class_perm is None else self.class_perm[y]      # preprocessing.py:540-541
        if X is None:
            return None, y_out.astype(np.float32)
        return self._apply(X, fit=True, base=base, fp=fp), y_out.astype(np.float32)

    def transform(self, X: Optional[np.ndarray], *, base=None, fp=None):
        return None if X is None else self._apply(X, fit=False, base=base, fp=fp)

    def _apply(self, X: np.ndarray, fit: bool, base=None, fp=None) -> np.ndarray:
        """``base`` / ``fp``: the recipe columns and the fingerprint column of X when the caller has
        them already (fit_transform_all / transform_all compute each once per call)."""
        X = np.asarray(X, dtype=np.float32)
        out = self.core.features(X, fit and not self.core.fitted) if base is None else base
        if self.fingerprint:
            out = np.concatenate([out, (_fingerprint(X) if fp is None else fp)[:, None]], 1)
        if self.feature_shift_seed is not None:
            if fit:
                self.feature_perm = np.random.default_rng(self.feature_shift_seed).permutation(out.shape[1])
            out = out[:, self.feature_perm]
        return np.ascontiguousarray(out, dtype=np.float32)


def _shared_parts(members, X: np.ndarray, fit: bool):
    X = np.asarray(X, dtype=np.float32)
    bases = {}
    for m in members:
        if id(m.core) not in bases:
            bases[id(m.core)] = m.core.features(X, fit)
    fp = _fingerprint(X) if any(m.fingerprint for m in members) else None
    return X, bases, fp


def fit_transform_all(members, X: Optional[np.ndarray], y: np.ndarray):
    """All members at once: [(X_train', y_train')]; recipe cores are fitted once, the row fingerprint
    is computed once."""
    if X is None:
        return [m.fit_transform(None, y) for m in members]
    X, bases, fp = _shared_parts(members, X, fit=True)
    return [m.fit_transform(X, y, base=bases[id(m.core)], fp=fp) for m in members]


def transform_all(members, X: Optional[np.ndarray]):
    """The test table of every member; the member-independent work (quantile / SVD projection,
    per-row fingerprint hashing) is done once per call, not once per member."""
    if X is None:
        return [None for _ in members]
    X, bases, fp = _shared_parts(members, X, fit=False)
    return [m.transform(X, base=bases[id(m.core)], fp=fp) for m in members]


def make_members(n_estimators: int, n_features: int, n_classes: int, rng: np.random.Generator, *,
                 recipes=RECIPES, fingerprint: bool = True, feature_shift: bool = True, class_shift: bool = True):
    """Balanced mix of the recipes, each with its own feature / class permutation
    (the structure of ``EnsembleConfig.generate_for_classification``, ``preprocessing.py:228-335``)."""
    members = []
    cores = {}
    for e in range(n_estimators):
        recipe = recipes[e * len(recipes) // max(n_estimators, 1)] if n_estimators >= len(recipes) else recipes[e % len(recipes)]
        fseed = int(rng.integers(0, 2**31 - 1)) if feature_shift and n_features > 0 else None
        cperm = rng.permutation(n_classes) if class_shift else None
        members.append(EnsembleMember(recipe=recipe, feature_shift_seed=fseed, class_perm=cperm,
                                      fingerprint=fingerprint, core=cores.setdefault(recipe, RecipeCore(recipe))))
    return members
