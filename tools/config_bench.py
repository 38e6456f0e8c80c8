"""BASELINE.json configs[2] / [3] / [4] in their STATED form, through the public API, one JSON line per config.

    python tools/config_bench.py cfg3                                           # 1 GPU
    python -m torch.distributed.run --nproc-per-node 8 tools/config_bench.py cfg4 cfg5

* cfg3 — image+text: 10 000 train / 10 000 test rows, 64 features + [N,2,768] embeddings, 8 estimators
  (``MMPFNClassifier``, bf16, context rebuilt per call), one GPU.
* cfg4 — large context: 50 000 train / 50 000 test rows, 100 features, 8 estimators, test rows sharded over the
  ranks, every estimator's context built on its owner rank and all-gathered layer by layer (``dist.ShardedEngine``).
* cfg5 — 256 independent small tasks (800 train + 200 test rows, 32 features + one embedding, 8 estimators each)
  dealt round-robin to the ranks and packed per launch (``tasks.predict_proba_tasks``).

Times are CUDA events around the device-resident step (max over ranks) and wall clock around the public call
(max over ranks); token counts are printed as run."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import bench
from multimodalpfn_b200.classifier import MMPFNClassifier
from multimodalpfn_b200.model import B200PerFeatureTransformer
from multimodalpfn_b200.preprocessing import transform_all
from multimodalpfn_b200.synth import Geometry, make_dataset, make_state_dict

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
geom = Geometry(mgm_heads=8, cap_heads=8)
sd = make_state_dict(geom, seed=1)
N_EST = 8


def sync():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def rmax(x):
    if world == 1:
        return float(x)
    t = torch.tensor([float(x)], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def emit(**kw):
    if rank == 0:
        print(json.dumps(kw), flush=True)


def run_table(name, dataset, steps, n_est=N_EST, shard="estimators"):
    d = make_dataset(dataset, 0)
    n_tr, n_te_all = len(d["y_train"]), len(d["y_test"])
    per = n_te_all // world
    sl = slice(rank * per, (rank + 1) * per)
    X_test = d["X_test"][sl]
    img_test = None if d["img_test"] is None else d["img_test"][sl]
    t0 = time.perf_counter()
    clf = MMPFNClassifier(mixer_type="MGM+CAP", mgm_heads=8, cap_heads=8, features_per_group=2, n_estimators=n_est,
                          model_path=(sd, geom), device=f"cuda:{local}", inference_precision="bf16",
                          ignore_pretraining_limits=True, random_state=0)
    clf.fit(d["X_train"], d["img_train"], d["y_train"])
    t_fit = time.perf_counter() - t0
    eng = clf.executor_
    if world > 1:
        from multimodalpfn_b200.dist import ShardedEngine
        eng = ShardedEngine(eng, rank, world, shard=shard)
        clf.executor_ = eng
    H_img = 0 if d["img_train"] is None else 8
    Ts = sorted({(g["F"] + 1) // 2 + H_img + 1 for g in eng.groups}, reverse=True)
    t0 = time.perf_counter()
    X_tests = transform_all(clf.members_, X_test)
    t_host = time.perf_counter() - t0
    perms = [m.class_perm for m in clf.members_]
    staged = eng.stage(X_tests, img_test)

    def step():
        lg = eng.logits_staged(staged)
        if world > 1:
            return eng.proba_gathered(lg, perms, n_classes=clf.n_classes_)
        from multimodalpfn_b200.engine import proba_device
        return proba_device(lg, perms, n_classes=clf.n_classes_)
    p = step()
    sync()
    times = []
    for _ in range(steps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync()
        a.record()
        p = step()
        b.record()
        sync()
        times.append(a.elapsed_time(b))
    ms = rmax(np.mean(times))
    assert p.shape[0] == per * world and bool(torch.isfinite(p).all())
    assert torch.allclose(p.sum(1), torch.ones_like(p[:, 0]), atol=1e-4)
    fl = sum(len(g["idx"]) * bench.flops_estimator(n_tr, per * world, (g["F"] + 1) // 2 + H_img + 1) for g in eng.groups)
    # e2e through predict_proba with host buffers (host transform of this rank's rows + H2D + step + D2H)
    if world == 1:
        clf.predict_proba(X_test, img_test)           # (captures the CUDA graph of this shape: not part of the timing)
    sync()
    t0 = time.perf_counter()
    if world == 1:
        clf.predict_proba(X_test, img_test)
    else:
        st2 = eng.stage(transform_all(clf.members_, X_test), img_test)
        eng.proba_gathered(eng.logits_staged(st2), perms, n_classes=clf.n_classes_).cpu()
    torch.cuda.synchronize()
    e2e = rmax(time.perf_counter() - t0)
    emit(config=name, n_gpus=world, workload=f"{dataset}: {n_tr} train / {per * world} test rows ({per} per rank), "
         f"{d['X_train'].shape[1]} features" + ("" if d["img_train"] is None else f" + {d['img_train'].shape[1]} x 768-d embeddings")
         + f", {n_est} estimator(s), bf16, context rebuilt per call, sharding: {shard}", T=Ts, ms_per_step=ms, value=per * world / ms * 1e3,
         unit="test rows/s", e2e_ms=e2e * 1e3, e2e_value=per * world / e2e, host_transform_ms_per_rank=t_host * 1e3,
         fit_s=t_fit, algorithmic_tflop=fl / 1e12, achieved_tflops_per_gpu=fl / (ms * 1e-3) / 1e12 / world,
         exchange=getattr(eng, "exchange", None), peak_mem_gib=torch.cuda.max_memory_allocated() / 2**30,
         steps=steps)
    del clf, eng, staged
    torch.cuda.empty_cache()


def run_tasks(n_tasks, steps):
    from multimodalpfn_b200.tasks import gather_task_results, predict_proba_tasks, shard_tasks
    model = B200PerFeatureTransformer(sd, geom, precision="bf16", seed=0)
    mine = shard_tasks(n_tasks, rank, world)
    tasks = {}
    for k in mine:
        d = make_dataset("small_task", k)
        tasks[k] = dict(X_train=d["X_train"], img_train=d["img_train"], y_train=d["y_train"], X_test=d["X_test"],
                        img_test=d["img_test"])
    lst = [tasks.get(i) for i in range(n_tasks)]
    walls, devs, hosts = [], [], []
    for it in range(steps + 1):
        sync()
        t0 = time.perf_counter()
        tm = {}
        out = predict_proba_tasks(model, lst, n_estimators=N_EST, random_state=0, indices=mine, timings=tm)
        allp = gather_task_results(out, n_tasks)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if it:
            walls.append(dt)
            devs.append(tm["device"])
            hosts.append(tm["host_prepare"])
    assert all(p is not None and p.shape[0] == 200 for p in allp)
    wall, dv, hs = rmax(np.mean(walls)), rmax(np.mean(devs)), rmax(np.mean(hosts))
    emit(config="cfg5", n_gpus=world, workload=f"{n_tasks} tasks x {N_EST} estimators (800 train + 200 test rows, 32 features + "
         "one 768-d embedding), dealt round-robin to the ranks, tasks of one shape packed per launch, bf16",
         tasks_per_rank=len(mine), wall_ms=wall * 1e3, device_ms=dv * 1e3, host_member_fitting_ms=hs * 1e3,
         value=n_tasks * 200 / wall, unit="test rows/s (wall, incl. host member fitting and the result gather)",
         device_value=n_tasks * 200 / dv, tasks_per_s=n_tasks / wall, steps=steps)


for cfg in (sys.argv[1:] or ["cfg3"]):
    if cfg == "cfg3":
        run_table("cfg3", "img_text_10k", steps=2)
    elif cfg == "cfg4":
        run_table("cfg4", "large_ctx_50k", steps=2)
    elif cfg == "cfg4_rows":        # the same table with the TRAIN ROWS of every estimator split over the ranks
        run_table("cfg4_rows", "large_ctx_50k", steps=2, shard="rows")
    elif cfg == "cfg4_one":         # ONE estimator's 50 000-row context: estimator ownership leaves W - 1 GPUs idle
        run_table("cfg4_one", "large_ctx_50k", steps=2, n_est=1)
    elif cfg == "cfg4_one_rows":    # ... row sharding uses all of them
        run_table("cfg4_one_rows", "large_ctx_50k", steps=2, n_est=1, shard="rows")
    elif cfg == "cfg5":
        run_tasks(256, steps=2)
    elif cfg == "dry":              # small stand-ins that exercise the same code paths (script check before a long run)
        run_table("dry_table", "pad_ufes_small", steps=1)
        run_tasks(2 * world, steps=1)
    else:
        raise SystemExit(f"unknown config {cfg}")
if world > 1:
    dist.destroy_process_group()
