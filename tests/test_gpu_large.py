"""Full-size property tests (BASELINE.json configs[2], configs[3]): sizes at which the oracle cannot run,
checked through what the domain guarantees: a test row's logits depend only on that
row and on the train context, so classifying the test rows in chunks, or in a different order, must give
the same logits bit for bit; probabilities are finite and normalised."""
import numpy as np
import pytest
import torch

from multimodalpfn_b200.model import B200PerFeatureTransformer
from multimodalpfn_b200.synth import Geometry, make_dataset, make_state_dict

pytestmark = pytest.mark.gpu


def _softmax(z):
    e = np.exp(z - z.max(-1, keepdims=True))
    return e / e.sum(-1, keepdims=True)


@pytest.mark.parametrize("name,n_cls", [("img_text_10k", 10), ("large_ctx_50k", 10)])
def test_test_row_chunking_and_order_invariance(name, n_cls):
    d = make_dataset(name, 0)
    has_img = d["img_train"] is not None
    geom = Geometry(mgm_heads=8, cap_heads=8)        # without embeddings the image stem is simply not run
    sd = make_state_dict(geom, seed=1)
    model = B200PerFeatureTransformer(sd, geom, precision="bf16", seed=0)
    Xtr, Xte = torch.as_tensor(d["X_train"]), torch.as_tensor(d["X_test"])
    itr = torch.as_tensor(d["img_train"]) if has_img else None
    ite = torch.as_tensor(d["img_test"]) if has_img else None
    y = torch.as_tensor(d["y_train"].astype(np.float32))
    ctx = model.fit_context(Xtr, itr, y)
    n = Xte.shape[0]
    full = model.predict_with_context(ctx, Xte, ite).float().cpu()
    assert full.shape == (1, n, 10) and torch.isfinite(full).all()
    # two chunks of uneven size (not multiples of the 128-row tile)
    cut = n // 2 + 37
    a = model.predict_with_context(ctx, Xte[:cut], None if ite is None else ite[:cut]).float().cpu()
    b = model.predict_with_context(ctx, Xte[cut:], None if ite is None else ite[cut:]).float().cpu()
    assert torch.equal(torch.cat([a, b], 1), full)
    # a permutation of the test rows permutes the logits
    perm = torch.as_tensor(np.random.default_rng(3).permutation(n))
    pp = model.predict_with_context(ctx, Xte[perm], None if ite is None else ite[perm]).float().cpu()
    assert torch.equal(pp, full[:, perm])
    p = _softmax(full[0, :, :n_cls].numpy() / 0.9)
    assert np.allclose(p.sum(1), 1.0, atol=1e-5)
    # the logits differ from row to row (random-init weights: the argmax need not)
    assert float(full[0].std(0).max()) > 0
