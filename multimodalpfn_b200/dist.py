"""Multi-GPU partitioning of the hot path on one NVLink/NVSwitch box.

One process per GPU (``torch.distributed``, backend ``nccl``; ``gloo`` in the CPU tests).  The
path shards along its two independent axes:

* **ensemble estimators** — the train context of estimator ``e`` (its stem statistics and the
  12-layer K/V context of the train rows) is built once, on rank ``e mod W``;
* **test rows** — every rank classifies its own chunk of test rows against ALL estimators.

The one real exchange step sits between the two: each context is broadcast from its owner to every
rank (``ncclBroadcast`` over NVLink; cfg2: 83 MB per T=27 estimator in bf16), and the per-rank
probabilities are all-gathered at the end.  Nothing else crosses GPUs.

The stem's "constant column" tests see the train rows plus the OWNER's test chunk
(``encoders.py:515, 615``); after preprocessing no train column is constant, so every chunk
gives the same mask (SURVEY.md section 8(e) caveat).
"""
from __future__ import annotations

import dataclasses
from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist

__all__ = ["owner_of", "local_members", "broadcast_bundle", "all_gather_rows", "ShardedEngine"]


def owner_of(member: int, world: int) -> int:
    return member % world


def local_members(members: List[int], rank: int, world: int) -> List[int]:
    """Positions (within ``members``) of the estimators this rank owns."""
    return [k for k, i in enumerate(members) if owner_of(i, world) == rank]


def broadcast_bundle(tensors: Optional[List[torch.Tensor]], shapes: List[Tuple[torch.Size, torch.dtype]], src: int,
                     device, group=None) -> List[torch.Tensor]:
    """Broadcast a list of tensors from ``src``; non-owners allocate from ``shapes``."""
    out = []
    for k, (shape, dtype) in enumerate(shapes):
        t = tensors[k].contiguous() if tensors is not None else torch.empty(shape, dtype=dtype, device=device)
        dist.broadcast(t, src=src, group=group)
        out.append(t)
    return out


def all_gather_rows(x: torch.Tensor, group=None) -> torch.Tensor:
    """[rows, c] per rank (same shape on every rank) -> [world*rows, c], rank-major."""
    world = dist.get_world_size(group)
    parts = [torch.empty_like(x) for _ in range(world)]
    dist.all_gather(parts, x.contiguous(), group=group)
    return torch.cat(parts, dim=0)


@dataclasses.dataclass
class _Sub:
    group: int            # index into engine.groups
    owner: int
    pos: List[int]        # positions inside the group's batch
    members: List[int]    # estimator ids


class ShardedEngine:
    """Wraps a ``B200InferenceEngine``: contexts are built by their owner rank and broadcast; every
    rank runs its own test rows against all of them."""

    def __init__(self, engine, rank: int, world: int, group=None):
        self.engine = engine
        self.model = engine.model
        self.rank, self.world, self.group = rank, world, group
        self.groups = engine.groups
        self.members = engine.members
        self.subs: List[_Sub] = []
        for gi, g in enumerate(engine.groups):
            for r in range(world):
                pos = local_members(g["idx"], r, world)
                if pos:
                    self.subs.append(_Sub(gi, r, pos, [g["idx"][k] for k in pos]))

    def stage(self, X_test_per_member, image_test):
        return self.engine.stage(X_test_per_member, image_test)

    def logits_staged(self, staged) -> torch.Tensor:
        from .model import TrainContext
        eng, m = self.engine, self.model
        dev = m.device
        eng.nan_flag.zero_()
        img_test_dev = staged["img_test"]
        tok = None
        n_img_train = 0
        if img_test_dev is not None:
            n_img_train = eng.img_train_dev.shape[0]
            tok = m.stem_image(torch.cat([eng.img_train_dev, img_test_dev], dim=0))
        out = [None] * len(self.members)
        # 1. build the contexts this rank owns (several sub-batches of different token counts: one batched
        #    pass, model.fit_contexts, when the model offers it)
        mine: Dict[int, TrainContext] = {}
        owned = [(si, sub) for si, sub in enumerate(self.subs) if sub.owner == self.rank]
        specs = []
        for si, sub in owned:
            g = eng.groups[sub.group]
            Xtr = None if g["X_train"] is None else g["X_train"][sub.pos].contiguous()
            Xte = staged["X_test"][sub.group]
            Xte = None if Xte is None else Xte[sub.pos].contiguous()
            ytr = g["y_train"][sub.pos].contiguous()
            n_tr = ytr.shape[1]
            ls = g["label_stats"]
            specs.append(dict(X_train=Xtr, y_train=ytr, X_all=None if Xte is None else torch.cat([Xtr, Xte], dim=1),
                              img_tok_train=None if tok is None else tok[:n_tr],
                              label_stats=(ls[0][sub.pos].contiguous(), ls[1][sub.pos].contiguous())))
        multi = (getattr(eng, "multi_group", False) and hasattr(m, "fit_contexts") and m.precision == 1
                 and len(specs) <= 8 and len(eng.groups) <= 8      # MMPFN_MAX_SEGMENTS
                 and all(sp["X_train"] is not None for sp in specs) and all(g["F"] >= 0 for g in eng.groups))
        if multi and len(specs) > 1:
            for (si, _), c in zip(owned, m.fit_contexts(specs, nan_flag=eng.nan_flag)):
                mine[si] = c
        else:
            for (si, _), sp in zip(owned, specs):
                mine[si] = m.fit_context(sp["X_train"], None, sp["y_train"], X_all=sp["X_all"],
                                         img_tok_train=sp["img_tok_train"], check=False,
                                         label_stats=sp["label_stats"], nan_flag=eng.nan_flag)
        # 2. replicate every context (the exchange step)
        ctxs: Dict[int, TrainContext] = {}
        for si, sub in enumerate(self.subs):
            g = eng.groups[sub.group]
            B, n_tr = len(sub.pos), g["y_train"].shape[1]
            F = max(g["F"], 0)
            G = m._n_groups(F) if g["F"] >= 0 else 0
            T = G + (0 if tok is None else tok.shape[1]) + 1
            n_stats = m.lib.mmpfn_tab_stats_elems(m._g, G) if G else 0
            shapes = [((m.lib.mmpfn_kv_bytes(m._g, B, n_tr, T, m.precision),), torch.uint8),
                      ((B,), torch.float32), ((B,), torch.int64)]
            if G:
                shapes.append(((B, n_stats), torch.float32))
            src = None
            if sub.owner == self.rank:
                c = mine[si]
                src = [c.kv, c.y_mean, c.y_mask] + ([c.tab_stats] if G else [])
            got = broadcast_bundle(src, shapes, sub.owner, dev, self.group)
            ctxs[si] = TrainContext(B=B, n_train=n_tr, F=F, T=T, n_tok=0, kv=got[0], tab_stats=got[3] if G else None,
                                    y_mean=got[1], y_mask=got[2], pos_emb=m.positional_embeddings(T - 1),
                                    precision=m.precision)
        # 3. this rank's test rows against every context.  The contexts of a group's estimators arrive
        #    from different owners; they are merged back into ONE batched context per group so that the
        #    test pass launches once per group (B = 4 here) instead of once per owner (B = 1 at W = 8:
        #    8 x as many launches of kernels that are already latency bound at 300 rows).
        tok_test = None if tok is None else tok[n_img_train:]
        merged = []
        for gi, g in enumerate(eng.groups):
            subs = [(si, sub) for si, sub in enumerate(self.subs) if sub.group == gi]
            if len(subs) == 1:
                ctx = ctxs[subs[0][0]]
            else:
                Bg = len(g["idx"])
                c0 = ctxs[subs[0][0]]
                y_mean = torch.empty(Bg, dtype=c0.y_mean.dtype, device=dev)
                y_mask = torch.empty(Bg, dtype=c0.y_mask.dtype, device=dev)
                stats = None if c0.tab_stats is None else torch.empty((Bg,) + tuple(c0.tab_stats.shape[1:]),
                                                                      dtype=c0.tab_stats.dtype, device=dev)
                for si, sub in subs:
                    c = ctxs[si]
                    y_mean[sub.pos] = c.y_mean
                    y_mask[sub.pos] = c.y_mask
                    if stats is not None:
                        stats[sub.pos] = c.tab_stats
                kv = m.merge_kv([(ctxs[si].kv, sub.pos) for si, sub in subs], Bg, c0.n_train, c0.T)
                ctx = TrainContext(B=Bg, n_train=c0.n_train, F=c0.F, T=c0.T, n_tok=0, kv=kv, tab_stats=stats,
                                   y_mean=y_mean, y_mask=y_mask, pos_emb=c0.pos_emb, precision=m.precision)
            merged.append(ctx)
        if multi and len(merged) > 1:
            lgs = m.predict_with_contexts(merged, staged["X_test"], img_tok_test=tok_test, nan_flag=eng.nan_flag)
        else:
            lgs = [m.predict_with_context(ctx, staged["X_test"][gi], None, img_tok_test=tok_test, check=False,
                                          nan_flag=eng.nan_flag) for gi, ctx in enumerate(merged)]
        for g, lg in zip(eng.groups, lgs):
            for k, i in enumerate(g["idx"]):
                out[i] = lg[k]
        return torch.stack(out)

    def logits(self, X_test_per_member, image_test, *, graph: bool = False) -> torch.Tensor:
        # collectives inside: replayed eagerly (NCCL graph capture is left for a later round)
        return self.logits_staged(self.stage(X_test_per_member, image_test))

    def check_nan(self):
        self.engine.check_nan()
