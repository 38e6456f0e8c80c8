"""BASELINE.json configs[4] on one GPU: the per-GPU share of 256 small tasks (800 train + 200 test rows,
32 features + one embedding, 8 estimators) packed per launch vs one task at a time.
usage: python tools/tasks_bench.py [n_tasks=32]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from multimodalpfn_b200.model import B200PerFeatureTransformer
from multimodalpfn_b200.synth import Geometry, make_dataset, make_state_dict
from multimodalpfn_b200.tasks import predict_proba_tasks

n_tasks = int(sys.argv[1]) if len(sys.argv) > 1 else 32
geom = Geometry(mgm_heads=8, cap_heads=8)
sd = make_state_dict(geom, seed=1)
model = B200PerFeatureTransformer(sd, geom, precision="bf16", seed=0)
tasks = []
for k in range(n_tasks):
    d = make_dataset("small_task", k)
    tasks.append(dict(X_train=d["X_train"], img_train=d["img_train"], y_train=d["y_train"], X_test=d["X_test"],
                      img_test=d["img_test"]))
for mode in ("packed", "one by one"):
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        tm = {"host_prepare": 0.0, "device": 0.0}
        if mode == "packed":
            out = predict_proba_tasks(model, tasks, n_estimators=8, random_state=0, timings=tm)
        else:
            out = {}
            for i, t in enumerate(tasks):
                t1 = {}
                out.update({i: predict_proba_tasks(model, [t], n_estimators=8, random_state=0, timings=t1)[0]})
                tm = {k: tm[k] + t1[k] for k in tm}
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    rows = sum(len(t["X_test"]) for t in tasks)
    print(f"{mode:10s}: {n_tasks} tasks x 8 estimators (800 train + 200 test rows, 32 features + 1 embedding): "
          f"{dt * 1e3:.0f} ms wall (host member fitting {tm['host_prepare'] * 1e3:.0f} ms, stem + 12 layers + tail "
          f"{tm['device'] * 1e3:.0f} ms) -> {n_tasks / dt:.1f} tasks/s, {rows / dt:.0f} test rows/s; "
          f"device part alone {rows / tm['device']:.0f} test rows/s", flush=True)
