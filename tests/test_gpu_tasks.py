"""BASELINE.json configs[4]: independent small tasks packed on the batch axis of every launch give, task by
task, exactly what MMPFNClassifier gives for that task alone (the reference would loop 256 x n_estimators
B = 1 forwards)."""
import numpy as np
import pytest

from multimodalpfn_b200.classifier import MMPFNClassifier
from multimodalpfn_b200.model import B200PerFeatureTransformer
from multimodalpfn_b200.synth import Geometry, make_dataset, make_state_dict
from multimodalpfn_b200.tasks import predict_proba_tasks, shard_tasks

pytestmark = pytest.mark.gpu


def _task(name, seed, with_img=True, with_x=True):
    d = make_dataset(name, seed)
    return dict(X_train=d["X_train"] if with_x else None, img_train=d["img_train"] if with_img else None,
                y_train=d["y_train"], X_test=d["X_test"] if with_x else None,
                img_test=d["img_test"] if with_img else None)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_packed_tasks_equal_one_by_one(precision):
    geom = Geometry(mgm_heads=2, cap_heads=4)
    sd = make_state_dict(geom, seed=5)
    # five tasks of one shape, one of another shape, one without embeddings: three launch groups
    tasks = [_task("tiny", k) for k in range(5)] + [_task("pad_ufes_small", 9), _task("tiny", 11, with_img=False)]
    model = B200PerFeatureTransformer(sd, geom, precision=precision, seed=0)
    packed = predict_proba_tasks(model, tasks, n_estimators=4, random_state=0)
    assert sorted(packed) == list(range(len(tasks)))
    for i, t in enumerate(tasks):
        clf = MMPFNClassifier(mixer_type="MGM+CAP", mgm_heads=2, cap_heads=4, features_per_group=2, n_estimators=4,
                              model_path=model, device="cuda", inference_precision=precision,
                              ignore_pretraining_limits=True, random_state=0)
        clf.fit(t["X_train"], t["img_train"], t["y_train"])
        one = clf.predict_proba(t["X_test"], t["img_test"])
        assert packed[i].shape == one.shape
        assert np.allclose(packed[i].sum(1), 1.0, atol=1e-5)
        # every row's arithmetic is independent of which rows share its launch: bit-identical
        assert np.array_equal(packed[i], one), (i, np.abs(packed[i] - one).max())


def test_task_sharding_round_robin():
    assert shard_tasks(10, 1, 4) == [1, 5, 9]
    assert sorted(sum((shard_tasks(10, r, 4) for r in range(4)), [])) == list(range(10))
