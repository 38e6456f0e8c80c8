"""Many independent small tasks packed per launch (BASELINE.json configs[4]).

The reference has a batch axis in the transformer (``x [S, B, F]``, state ``[B, S, T, E]``,
``layer.py:284-285``) but its engines always pass ``B = 1`` (``inference.py:305``) and the CAP stem
hard-codes it (``transformer.py:78-79``): 256 small datasets mean 256 x n_estimators sequential
forwards there.  Here the batch axis of every kernel carries *tasks*: all tasks with the same shape
run as ONE batched forward per estimator slot (their image tokens ride along per batch entry,
``mmpfn_stem_tokens(img_bstride)``), so the launches see ``B x S x T`` tokens instead of ``S x T``.
Tasks are independent until the end, so a multi-GPU run just deals them out round-robin and gathers
the probabilities — no collective on the data path.

The result for every task is what ``MMPFNClassifier(n_estimators=..., random_state=...)`` returns
for that task alone (same members, same reference-equivalent joint forward); the packing only
changes which rows share a kernel launch.
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch
from sklearn.preprocessing import LabelEncoder

from .engine import proba_from_logits
from .model import B200PerFeatureTransformer
from .preprocessing import RECIPES, fit_transform_all, make_members, transform_all

__all__ = ["predict_proba_tasks", "shard_tasks", "gather_task_results"]


def shard_tasks(n_tasks: int, rank: int, world: int) -> list:
    """Indices of the tasks rank ``rank`` owns (round-robin)."""
    return list(range(rank, n_tasks, world))


def gather_task_results(local: dict, n_tasks: int) -> list:
    """All-gather ``{task index: proba}`` dictionaries over the default process group -> list in task order."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [local.get(i) for i in range(n_tasks)]
    parts = [None] * dist.get_world_size()
    dist.all_gather_object(parts, local)
    merged = {}
    for p in parts:
        merged.update(p)
    return [merged.get(i) for i in range(n_tasks)]


def _prepare(task: dict, n_estimators: int, random_state, recipes):
    X = None if task.get("X_train") is None else np.asarray(task["X_train"], dtype=np.float32)
    Xte = None if task.get("X_test") is None else np.asarray(task["X_test"], dtype=np.float32)
    y = np.asarray(task["y_train"])
    enc = LabelEncoder()
    yi = enc.fit_transform(y)
    n_cls = len(enc.classes_)
    _, counts = np.unique(y, return_counts=True)
    rng = np.random.default_rng(random_state if isinstance(random_state, (int, np.integer)) else None)
    members = make_members(n_estimators, 0 if X is None else X.shape[1], n_cls, rng, recipes=tuple(recipes))
    train = fit_transform_all(members, X, yi)
    test = transform_all(members, Xte)

    def img(a):
        if a is None:
            return None
        a = np.asarray(a, dtype=np.float32)
        return a[:, None] if a.ndim == 2 else a
    itr, ite = img(task.get("img_train")), img(task.get("img_test"))
    n_te = len(Xte) if Xte is not None else len(ite)
    key = (len(y), n_te, None if itr is None else itr.shape[1:], tuple(None if t[0] is None else t[0].shape[1] for t in train))
    return dict(members=members, train=train, test=test, img_train=itr, img_test=ite, n_classes=n_cls, counts=counts,
                classes=enc.classes_, key=key)


def predict_proba_tasks(model: B200PerFeatureTransformer, tasks: Sequence[dict], *, n_estimators: int = 4,
                        random_state=0, softmax_temperature: float = 0.9, average_before_softmax: bool = False,
                        balance_probabilities: bool = False, recipes=RECIPES, indices: Optional[Sequence[int]] = None,
                        n_jobs: int = 1, timings: Optional[dict] = None):
    """``tasks``: dicts with ``X_train, img_train, y_train, X_test, img_test`` (tables or embeddings may be
    ``None`` like in ``MMPFNClassifier.fit``).  Returns ``{task index: float32 [Nte, n_classes]}`` for the
    tasks in ``indices`` (default: all).  Tasks of equal shape share launches.  The per-task host work
    (fitting the members' quantile / SVD transforms) can run in ``n_jobs`` threads (it is mostly GIL-bound
    sklearn glue: 16 threads measured slower than 1); ``timings`` (a dict) receives
    the seconds spent in ``host_prepare`` and ``device``."""
    import os
    import time
    from concurrent.futures import ThreadPoolExecutor
    indices = list(range(len(tasks))) if indices is None else list(indices)
    dev = model.device
    t0 = time.perf_counter()
    workers = (os.cpu_count() or 1) if n_jobs in (-1, None) else max(1, int(n_jobs))
    workers = min(workers, max(1, len(indices)))
    if workers > 1:
        with ThreadPoolExecutor(workers) as pool:
            prepared = dict(zip(indices, pool.map(lambda i: _prepare(tasks[i], n_estimators, random_state, recipes),
                                                  indices)))
    else:
        prepared = {i: _prepare(tasks[i], n_estimators, random_state, recipes) for i in indices}
    t1 = time.perf_counter()
    groups = {}
    for i in indices:
        groups.setdefault(prepared[i]["key"], []).append(i)
    logits = {i: [None] * n_estimators for i in indices}
    for key, idx in groups.items():
        n_tr, n_te, img_shape, widths = key
        tok_tr = tok_te = None
        if img_shape is not None:
            # one image-stem pass over the embeddings of every task of the group: [Bt * S, n_tok, I] -> [Bt, S, H, E]
            img_all = np.concatenate([np.concatenate([prepared[i]["img_train"], prepared[i]["img_test"]]) for i in idx])
            tok = model.stem_image(torch.from_numpy(img_all).to(dev))
            tok = tok.view(len(idx), n_tr + n_te, tok.shape[1], tok.shape[2])
            tok_tr, tok_te = tok[:, :n_tr].contiguous(), tok[:, n_tr:].contiguous()
        for e in range(n_estimators):
            ytr = torch.from_numpy(np.stack([prepared[i]["train"][e][1] for i in idx])).to(dev)
            if widths[e] is None:
                Xtr = Xte = X_full = None
            else:
                Xtr = torch.from_numpy(np.stack([prepared[i]["train"][e][0] for i in idx])).to(dev)
                Xte = torch.from_numpy(np.stack([prepared[i]["test"][e] for i in idx])).to(dev)
                X_full = torch.cat([Xtr, Xte], dim=1)
            # reference-equivalent joint forward (the stem's constant-column tests see train and test rows)
            flag = torch.zeros(1, dtype=torch.int32, device=dev)      # one NaN flag for train and test rows
            ctx = model.fit_context(Xtr, None, ytr, X_all=X_full, img_tok_train=tok_tr, check=False, nan_flag=flag)
            lg = model.predict_with_context(ctx, Xte, None, img_tok_test=tok_te, nan_flag=flag)
            for k, i in enumerate(idx):
                logits[i][e] = lg[k]
            del ctx
    out = {}
    for i in indices:
        pr = prepared[i]
        out[i] = proba_from_logits(torch.stack(logits[i]), [m.class_perm for m in pr["members"]],
                                   n_classes=pr["n_classes"], class_counts=pr["counts"],
                                   softmax_temperature=softmax_temperature,
                                   average_before_softmax=average_before_softmax,
                                   balance_probabilities=balance_probabilities)
    if timings is not None:
        torch.cuda.synchronize(dev)
        timings["host_prepare"] = t1 - t0
        timings["device"] = time.perf_counter() - t1
    return out
