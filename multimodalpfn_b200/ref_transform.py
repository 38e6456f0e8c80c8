"""Fast replay of a FITTED reference preprocessor at predict time.

The reference transforms the test table once per ensemble member with that member's fitted
``SequentialFeatureTransformer`` (``inference.py:303``, ``model/preprocessing.py:371-441``): a chain of steps whose
heavy ones wrap sklearn objects (``ColumnTransformer`` / ``Pipeline`` / ``FeatureUnion`` over ``QuantileTransformer``,
``SimpleImputer``, ``StandardScaler``, ``TruncatedSVD``, ``OrdinalEncoder``; ``model/preprocessing.py:579-996,
998-1200``).  For a 300-row table the arithmetic is a few tens of microseconds, but every sklearn ``transform`` call
re-validates its input, resolves estimator tags and goes through joblib: 1-8 ms per member, GIL-bound, which is what
is left of ``predict_proba`` once the forward takes milliseconds.

``compile_preprocessor`` walks the fitted objects ONCE and returns a closure that performs the same arithmetic on
the same fitted state — numpy calls in the same order, on arrays of the same dtype and memory layout (so that even
the BLAS call of the SVD projection is the same call), sklearn's own column kernels where they exist
(``QuantileTransformer._transform``) — without the per-call validation.  Anything it does not know raises
``Unsupported`` and the caller keeps the reference's ``transform`` for that member.  ``verify`` compares the replay
with the reference's own ``transform`` bit for bit (float64, NaN == NaN) on probe rows; the plug-in does that on the
first table it sees (plus perturbed copies: NaNs, out-of-range values, unseen categories) before it trusts a replay.

Nothing here is specific to the GPU path: it is host code that stands before the boundary, next to the reference's.
"""
from __future__ import annotations

import hashlib

import numpy as np

__all__ = ["Unsupported", "compile_preprocessor", "replay_all", "make_probe", "verify"]

_FLOATS = (np.dtype(np.float32), np.dtype(np.float64))


class Unsupported(Exception):
    """The fitted object tree contains something the replay does not reproduce."""


# Ensemble members share fitted state: with the reference's default recipes the four "quantile + SVD" members hold
# the same fitted QuantileTransformer, scaling chain and OrdinalEncoder (same data, same seeds) and differ only in the SVD
# basis, category shuffle, fingerprint salt and feature shift.  ``replay_all`` opens a per-call memo: a heavy node whose
# fitted parameters AND input bytes equal those of an earlier evaluation in the same call returns that result (marked
# read-only; every consumer here copies before it writes).
_MEMO = None


def _digest(*parts) -> bytes:
    h = hashlib.blake2b(digest_size=16)
    for p in parts:
        if isinstance(p, np.ndarray):
            h.update(str((p.shape, p.dtype.str)).encode())
            h.update(np.ascontiguousarray(p).tobytes())
        elif isinstance(p, bytes):
            h.update(p)
        else:
            h.update(repr(p).encode())
        h.update(b"|")
    return h.digest()


def _memo(sig: bytes, fn):
    """``fn`` with the per-call memo in front (key: node signature + the input's shape, dtype and bytes)."""
    def run(X):
        memo = _MEMO
        if memo is None:
            return fn(X)
        # input identity: shape, dtype, memory order and Python's keyed 64-bit hash of the bytes as they lie in memory
        # (SipHash, ~2.5 GB/s; a different layout of equal values is simply a miss)
        key = (sig, X.shape, X.dtype.str, X.strides, hash(X.tobytes(order="A")))
        hit = memo.get(key)
        if hit is None:
            hit = fn(X)
            if isinstance(hit, np.ndarray):
                hit.setflags(write=False)
            memo[key] = hit
        return hit
    run.signature = sig
    return run


def _float_copy(X):
    """What sklearn's ``validate_data(..., dtype=FLOAT_DTYPES, copy=True)`` hands to the arithmetic: a fresh array,
    float32 / float64 kept, everything else promoted to float64, memory order kept."""
    return np.array(X, dtype=X.dtype if X.dtype in _FLOATS else np.float64, copy=True)


def _int_columns(cols, n_features):
    if isinstance(cols, (list, tuple, np.ndarray)) and len(cols) > 0 and all(
            isinstance(c, (int, np.integer)) and not isinstance(c, (bool, np.bool_)) for c in cols):
        if max(cols) >= n_features:
            raise Unsupported("column index past the fitted width")
        return list(cols) if not isinstance(cols, np.ndarray) else cols
    raise Unsupported(f"column specifier {type(cols).__name__}")


def _compile_sk(t):
    """fitted sklearn node -> callable(ndarray [n, k]) -> ndarray"""
    if t is None or (isinstance(t, str) and t == "passthrough"):
        return lambda X: X
    name = type(t).__name__
    if name in ("FunctionTransformer", "NoneTransformer"):
        if getattr(t, "validate", False):
            raise Unsupported("FunctionTransformer(validate=True)")
        func, kw = t.func, dict(t.kw_args or {})
        if func is None:
            return lambda X: X

        def run_func(X):
            return func(X, **kw)
        if not kw and getattr(func, "__module__", "").endswith("model.preprocessing"):
            run_func.signature = _digest("func", func.__name__)         # the reference's own pure helpers (_inf_to_nan_func, ...)
        return run_func
    if name == "Pipeline":
        fs = [_compile_sk(s) for _, s in t.steps]

        def run_pipeline(X):
            for f in fs:
                X = f(X)
            return X
        sigs = [getattr(f, "signature", None) for f in fs]
        if len(fs) > 1 and all(sg is not None for sg in sigs):        # a chain of pure fitted nodes: one memo entry for all
            return _memo(_digest("pipeline", *sigs), run_pipeline)
        return run_pipeline
    if name == "FeatureUnion":
        if t.transformer_weights:
            raise Unsupported("FeatureUnion weights")
        fs = [_compile_sk(s) for _, s in t.transformer_list if not (isinstance(s, str) and s == "drop")]
        return lambda X: np.hstack([f(X) for f in fs])
    if name == "ColumnTransformer":
        if getattr(t, "sparse_output_", False) or t.transformer_weights:
            raise Unsupported("ColumnTransformer sparse output / weights")
        n_in = int(t.n_features_in_)
        parts = []
        for _, tr, cols in t.transformers_:
            if isinstance(tr, str) and tr == "drop":
                continue
            if isinstance(cols, (list, tuple, np.ndarray)) and len(cols) == 0:
                continue
            parts.append((_compile_sk(tr), _int_columns(cols, n_in)))
        if not parts:
            raise Unsupported("ColumnTransformer without output columns")

        def run_columns(X):
            if X.shape[1] != n_in:
                raise ValueError(f"X has {X.shape[1]} features, but ColumnTransformer is expecting {n_in} features as input.")
            return np.hstack([f(X[:, cols]) for f, cols in parts])
        return run_columns
    if name == "QuantileTransformer":
        if not hasattr(t, "_transform") or not hasattr(t, "quantiles_"):
            raise Unsupported("QuantileTransformer internals")

        def run_quantile(X):
            Xc = _float_copy(X)
            if np.isinf(Xc).any():        # sklearn's validation (ensure_all_finite="allow-nan") refuses these
                raise ValueError("Input X contains infinity or a value too large for dtype('float64').")
            return t._transform(Xc, inverse=False)      # sklearn's own per-column interpolation kernel
        return _memo(_digest("quantile", t.quantiles_, t.references_, t.output_distribution), run_quantile)
    if name == "SimpleImputer":
        stats = np.asarray(t.statistics_, dtype=np.float64)
        mv = t.missing_values
        if t.add_indicator or not (isinstance(mv, float) and np.isnan(mv)) or np.isnan(stats).any() \
                or t.strategy not in ("mean", "median", "most_frequent", "constant"):
            raise Unsupported("SimpleImputer configuration")

        def run_impute(X):
            Xc = _float_copy(X)
            if np.isinf(Xc).any():
                raise ValueError("Input X contains infinity or a value too large for dtype('float64').")
            mask = np.isnan(Xc)
            if mask.any():
                Xc[mask] = np.broadcast_to(stats.astype(Xc.dtype, copy=False), Xc.shape)[mask]
            return Xc
        run_impute.signature = _digest("impute", stats)
        return run_impute
    if name == "StandardScaler":
        mean = t.mean_ if t.with_mean else None
        scale = t.scale_ if t.with_std else None

        def run_scale(X):
            Xc = _float_copy(X)
            if mean is not None:
                Xc -= mean
            if scale is not None:
                Xc /= scale
            return Xc
        run_scale.signature = _digest("scale", "none" if mean is None else mean, "none" if scale is None else scale)
        return run_scale
    if name == "TruncatedSVD":
        comp_t = t.components_.T            # the view sklearn multiplies with (safe_sparse_dot(X, components_.T))

        def run_svd(X):
            if not np.isfinite(X).all():
                raise ValueError("Input X contains NaN.")
            return X @ comp_t
        return run_svd
    if name == "OrdinalEncoder":
        if t.handle_unknown != "use_encoded_value" or t.max_categories is not None or t.min_frequency is not None \
                or getattr(t, "_infrequent_enabled", False):
            raise Unsupported("OrdinalEncoder configuration")
        cats = [np.asarray(c) for c in t.categories_]
        if any(c.dtype.kind not in "fiu" or c.size == 0 for c in cats):
            raise Unsupported("OrdinalEncoder over non-numeric categories")
        cats = [c.astype(np.float64) for c in cats]
        has_nan = [bool(np.isnan(c[-1])) for c in cats]
        missing = dict(t._missing_indices)
        unknown, enc_missing, dtype = t.unknown_value, t.encoded_missing_value, t.dtype

        def run_ordinal(X):
            n, k = X.shape
            if k != len(cats):
                raise ValueError(f"X has {k} features, but OrdinalEncoder is expecting {len(cats)} features as input.")
            if X.dtype.kind not in "fiu":
                raise Unsupported("non-numeric table")
            out = np.empty((n, k), dtype=dtype)
            for j, c in enumerate(cats):
                x = X[:, j]
                idx = np.searchsorted(c, x)                     # NaN sorts last: its own slot if fitted, else len(c)
                hit = c[np.minimum(idx, c.size - 1)] == x
                if has_nan[j]:
                    hit |= np.isnan(x)
                col = idx.astype(dtype)
                if j in missing:
                    col[hit & (idx == missing[j])] = enc_missing
                col[~hit] = unknown
                out[:, j] = col
            return out
        return _memo(_digest("ordinal", *cats, sorted(missing.items()), unknown, enc_missing, np.dtype(dtype).str), run_ordinal)
    raise Unsupported(name)


_HASH_MOD = 10 ** 12


def _compile_step(step):
    name = type(step).__name__
    if name == "RemoveConstantFeaturesStep":                      # model/preprocessing.py:468-470
        sel = step.sel_
        return lambda X: X[:, sel]
    if name == "ShuffleFeaturesStep":                             # :566-571
        perm = step.index_permutation_

        def run_shuffle(X):
            assert len(perm) == X.shape[1], "The number of features must not change after fit"
            return X[:, perm]
        return run_shuffle
    if name == "ReshapeFeatureDistributionsStep":                 # :993-995
        sub = step.subsampled_features_
        f = _compile_sk(step.transformer_)
        return lambda X: f(X[:, sub])
    if name == "EncodeCategoricalFeaturesStep":                   # :1189-1200
        ct = step.categorical_transformer_
        if ct is None:
            return lambda X: X
        f = _compile_sk(ct)
        shuffled = step.categorical_transform_name.endswith("_shuffled")
        mappings = dict(step.random_mappings_) if shuffled else {}

        def run_encode(X):
            out = f(X)
            if mappings and not out.flags.writeable:
                out = out.copy()
            for col, mapping in mappings.items():
                column = out[:, col]
                keep = ~np.isnan(column)
                column[keep] = mapping[column[keep].astype(int)].astype(column.dtype)
            return out
        return run_encode
    if name == "AddFingerprintFeaturesStep":                      # :501-523, the is_test branch
        salt = step.rnd_salt_

        def run_fingerprint(X):
            rows = (X + salt) + salt          # the reference hashes row + salt of the table it already salted once
            raw = np.ascontiguousarray(rows).tobytes()            # the same bytes as row.tobytes(), row by row
            w = rows.shape[1] * rows.dtype.itemsize
            hs = np.array([hash(raw[i:i + w]) for i in range(0, w * rows.shape[0], w)], dtype=np.int64)
            # hash % 10**12 / 10**12 as the reference computes it on Python ints: floor modulo (numpy agrees for a positive
            # modulus), the remainder is below 2**53, so the true division is the same correctly rounded float64 quotient
            h = ((hs % _HASH_MOD) / _HASH_MOD).astype(X.dtype)
            return np.concatenate([X, h.reshape(-1, 1)], axis=1)
        return run_fingerprint
    raise Unsupported(name)


def compile_preprocessor(seq):
    """``SequentialFeatureTransformer`` (fitted) -> ``fast(X) -> ndarray``, the same table as ``seq.transform(X).X``.
    Raises ``Unsupported`` when a step or sklearn node is not reproduced."""
    steps = [_compile_step(s) for s in seq]
    if not steps:
        raise Unsupported("empty preprocessor")

    def fast(X):
        for f in steps:
            X = f(X)
        return X
    return fast


def replay_all(fasts, X):
    """All members' replays of one table with the shared per-call memo (see ``_memo``) -> list of arrays."""
    global _MEMO
    _MEMO = {}
    try:
        return [None if f is None else f(X) for f in fasts]
    finally:
        _MEMO = None


def make_probe(X, seed: int = 0):
    """Rows that exercise the edge branches of the transforms: the table itself, a copy with NaNs, a copy pushed out
    of the fitted range (beyond the outer quantiles, unseen category codes), a copy with a few exact repeats."""
    X = np.asarray(X)
    rng = np.random.default_rng(seed)
    Xf = X.astype(np.float64)
    holes = Xf.copy()
    holes[rng.random(X.shape) < 0.1] = np.nan
    spread = np.nanstd(Xf, axis=0, keepdims=True) + 1.0
    far = Xf + rng.choice([-3.0, 3.0], size=X.shape) * spread
    rounded = np.round(Xf * 1.5)
    return np.concatenate([Xf, holes, far, rounded, Xf[:1]]).astype(X.dtype if X.dtype in _FLOATS else np.float64)


def verify(seq, fast, X) -> bool:
    """Is the replay identical to the reference's own ``transform`` on ``X`` (float64 bits, NaN == NaN)?"""
    try:
        ref = np.asarray(seq.transform(X).X)
    except Exception as exc:               # the reference refuses this table: the replay must refuse it too
        try:
            fast(X)
        except Exception as exc2:
            return type(exc2) is type(exc)
        return False
    try:
        got = np.asarray(fast(X))
    except Exception:
        return False
    return ref.shape == got.shape and ref.dtype == got.dtype and bool(np.array_equal(ref, got, equal_nan=True))
