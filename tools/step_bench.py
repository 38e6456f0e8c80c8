"""The device-resident step (8 estimators, context rebuilt per call, CUDA graph, L2 flushed between steps) alone:
    [MMPFN_DEBUG_LIB=1 <tuning switches>] python tools/step_bench.py [dataset]
Used for A/B timing of variants of the tuning build (the product library reads no environment), e.g. the
L2-resident chunked schedule of profiles/r02_chunked_schedule.txt (measured, no gain, removed)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from multimodalpfn_b200 import _lib
from multimodalpfn_b200.classifier import MMPFNClassifier
from multimodalpfn_b200.preprocessing import transform_all
from multimodalpfn_b200.synth import Geometry, make_dataset, make_state_dict

dev = torch.device("cuda", 0)
geom = Geometry(mgm_heads=8, cap_heads=8)
sd = make_state_dict(geom, seed=1)
name = sys.argv[1] if len(sys.argv) > 1 else "pad_ufes"
d = make_dataset(name, 0)
clf = MMPFNClassifier(mixer_type="MGM+CAP", mgm_heads=8, cap_heads=8, features_per_group=2, n_estimators=8,
                      model_path=(sd, geom), device="cuda:0", inference_precision="bf16", ignore_pretraining_limits=True,
                      random_state=0).fit(d["X_train"], d["img_train"], d["y_train"])
eng = clf.executor_
n_te = min(len(d["X_test"]), 300)
staged = eng.stage(transform_all(clf.members_, d["X_test"][:n_te]), None if d["img_test"] is None else d["img_test"][:n_te])
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
n0 = _lib.launch_count()
ref = eng.logits_staged(staged).clone()
launches = _lib.launch_count() - n0
for _ in range(3):
    out = eng.logits_graphed(staged)
torch.cuda.synchronize()
ts = []
for _ in range(10):
    flush.fill_(1)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    out = eng.logits_graphed(staged)
    b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
env = {k: v for k, v in os.environ.items() if k.startswith("MMPFN_")}
print(f"{name} {env}: step median {np.median(ts):.2f} ms (min {min(ts):.2f}), {launches} launches, "
      f"logits checksum {float(ref.double().abs().sum()):.10e}", flush=True)
