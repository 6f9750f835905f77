"""Device-side downstream classifier (tools/device_classifier.py, csrc/logreg.cu + the tcgen05 GEMM): the accuracy yardstick
(tools/graph_model_downstream_classification.py:85-148 of the reference) without moving the embedding matrix to the host."""
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN, cuda_device
from shallow_encoders import _native as nat

pytestmark = pytest.mark.gpu


def test_softmax_xent_kernel_matches_torch():
    dev = cuda_device()
    g = torch.Generator().manual_seed(0)
    for n, c in ((1000, 7), (513, 2), (100, 20)):
        logits = (torch.randn(n, c, generator=g) * 3).to(dev)
        labels = torch.randint(0, c, (n,), generator=g).to(dev)
        bias = torch.randn(c, generator=g).to(dev)
        ref = (logits + bias).double().requires_grad_(True)
        loss_ref = torch.nn.functional.cross_entropy(ref, labels, reduction='sum')
        loss_ref.backward()
        loss = torch.zeros(1, dtype=torch.float64, device=dev); ok = torch.zeros(1, dtype=torch.int32, device=dev)
        z = logits.clone()
        nat.softmax_xent(z, labels.to(torch.int32), bias, 0.5, loss, ok)
        assert abs(float(loss) - float(loss_ref)) < 1e-4 * float(loss_ref)
        assert float((z.double() - 0.5 * ref.grad).abs().max()) < 1e-5
        assert int(ok) == int(((logits + bias).argmax(1) == labels).sum())
    # binary form: one logit column
    s = (torch.randn(777, 1, generator=g) * 2).to(dev); y = torch.randint(0, 2, (777,), generator=g).to(dev)
    ref = s.double().clone().requires_grad_(True)
    l = torch.nn.functional.binary_cross_entropy_with_logits(ref[:, 0], y.double(), reduction='sum'); l.backward()
    loss = torch.zeros(1, dtype=torch.float64, device=dev)
    z = s.clone()
    nat.softmax_xent(z, y.to(torch.int32), None, 1.0, loss)
    assert abs(float(loss) - float(l)) < 1e-4 * float(l) and float((z.double() - ref.grad).abs().max()) < 1e-5


@pytest.mark.parametrize('n_classes', [2, 5])
def test_device_logistic_regression_reaches_sklearns_optimum(n_classes):
    """Same objective as sklearn's default LogisticRegression: the fitted coefficients and predictions agree."""
    from sklearn.linear_model import LogisticRegression
    from tools.device_classifier import DeviceLogisticRegression
    dev = cuda_device()
    rng = np.random.default_rng(n_classes)
    n, e = 3000, 32
    centres = rng.standard_normal((n_classes, e)) * 1.2
    y = rng.integers(0, n_classes, n)
    x = (centres[y] + rng.standard_normal((n, e))).astype(np.float32)
    sk = LogisticRegression(max_iter=2000, tol=1e-8).fit(x, y)
    dv = DeviceLogisticRegression(max_iter=300, tol=1e-5).fit(torch.from_numpy(x).to(dev), torch.from_numpy(y).to(dev))
    pred = dv.predict(torch.from_numpy(x).to(dev)).cpu().numpy()
    assert (pred == sk.predict(x)).mean() > 0.995
    coef = dv.coef_.cpu().numpy()
    want = sk.coef_
    assert coef.shape == want.shape
    assert np.abs(coef - want).max() < 0.02 * np.abs(want).max() + 1e-3
    assert abs(dv.score(torch.from_numpy(x).to(dev), torch.from_numpy(y).to(dev)) - sk.score(x, y)) < 0.003


def test_device_node_classification_matches_the_reference_experiment_by_experiment():
    """The reference's own perform_node_classification on a fixed karate-club embedding (fixture from oracle/make_golden.py: 20 experiments,
    sklearn on the host) against node_classification_device: same splits, the device classifier; accuracies are quantised to 1/34, so
    at least 18 of 20 experiments must be identical and none may differ by more than one node."""
    from shallow_encoders.split import TrainTestRatioSplit
    from tools.downstream import node_classification_device
    dev = cuda_device()
    z = np.load(os.path.join(GOLDEN, 'downstream_karate.npz'))
    itos = [str(x) for x in z['itos']]
    labels = {name: str(lab) for name, lab in zip(itos[1:], z['labels'])}
    mean_acc, best_acc, accs = node_classification_device(torch.from_numpy(z['embedding']).to(dev), itos, labels,
                                                          TrainTestRatioSplit(train_ratio=0.5, test_all=True), 20, None, return_all=True)
    want = z['node_accuracies']
    diff = np.abs(np.array(accs) - want)
    assert (diff < 1e-6).sum() >= 18 and diff.max() <= 1 / 34 + 1e-6, (accs, want.tolist())
    assert abs(mean_acc - want.mean()) < 0.005 and abs(best_acc - want.max()) < 1e-6
