"""Host-side profile of MMPFNClassifier.predict_proba at the cfg2 shape (where do the milliseconds
between the device-timed step and the end-to-end call go?)."""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from multimodalpfn_b200.classifier import MMPFNClassifier
from multimodalpfn_b200.synth import Geometry, make_dataset, make_state_dict

geom = Geometry(mgm_heads=8, cap_heads=8)
sd = make_state_dict(geom, seed=1)
d = make_dataset("pad_ufes", 0)
clf = MMPFNClassifier(mixer_type="MGM+CAP", mgm_heads=8, cap_heads=8, features_per_group=2, n_estimators=8,
                      model_path=(sd, geom), device="cuda:0", inference_precision="bf16",
                      ignore_pretraining_limits=True, random_state=0)
clf.fit(d["X_train"], d["img_train"], d["y_train"])
for _ in range(3):
    clf.predict_proba(d["X_test"], d["img_test"])
torch.cuda.synchronize()
ts = []
for _ in range(10):
    t0 = time.perf_counter()
    clf.predict_proba(d["X_test"], d["img_test"])
    ts.append(time.perf_counter() - t0)
print("predict_proba ms:", [round(t * 1e3, 2) for t in ts])
pr = cProfile.Profile()
pr.enable()
for _ in range(10):
    clf.predict_proba(d["X_test"], d["img_test"])
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
