"""GPU check of the training-time twin: BFMNet's vertex loss and its gradient
(voicepuppet/bfmnet/bfmnet.py:215-268) against the float64 numpy restatement (parity unpinned: TensorFlow is
not available to run the reference graph) -- value within 1e-5 relative, gradient against central differences
of the oracle and against the analytic float64 gradient."""
import numpy as np
import pytest

from oracle import shape_loss_oracle as slo
from voicepuppet_b200 import synthetic
from voicepuppet_b200.shape_loss import ExpressionShapeLoss

pytestmark = pytest.mark.gpu


def make_case(model, b, t, seed):
  rng = np.random.Generator(np.random.PCG64(seed))
  coeffs = np.stack([synthetic.make_coeffs(t, seed=seed + 10 * i) for i in range(b)])        # [B,T,257]
  pred = (coeffs[:, :, 80:144] + 0.3 * rng.standard_normal((b, t, 64))).astype(np.float32)
  nver = model.meanshape.size // 3
  mask = np.ones((nver, 3), np.float32)
  mask[rng.choice(nver, nver // 12, replace=False)] = 10.0                                    # bfmnet.py:134-137
  seq_len = np.array([t] + [int(x) for x in rng.integers(2, t + 1, b - 1)])
  return coeffs, pred, mask, seq_len


def analytic_grad(pred, coeffs, seq_len, model, mask):
  """d cost / d pred in float64 from the closed form (signs of D and of its temporal difference)."""
  b, t = pred.shape[:2]
  ex = np.asarray(model.exBase, dtype=np.float64)
  m = mask.reshape(-1).astype(np.float64)
  delta = coeffs[:, :, 80:144].astype(np.float64) - pred.astype(np.float64)
  d = delta @ ex.T                                                  # [B,T,3N]
  g = np.zeros_like(d)
  for i in range(b):
    n = int(seq_len[i])
    g[i, :n] += np.sign(d[i, :n])
    s = np.sign(d[i, :n - 1] - d[i, 1:n])
    g[i, :n - 1] += s
    g[i, 1:n] -= s
  g *= m / b
  return -(g @ ex)                                                  # d/d pred = - d/d delta


@pytest.mark.parametrize('b,t', [(2, 7), (3, 20)])
def test_loss_and_gradient_small_model(small_model, b, t):
  import torch
  coeffs, pred, mask, seq_len = make_case(small_model, b, t, seed=40 + b)
  want = slo.cost(pred.astype(np.float64), coeffs.astype(np.float64), seq_len, small_model, mask)
  loss_fn = ExpressionShapeLoss(small_model, mask)
  p = torch.tensor(pred, device='cuda:0', requires_grad=True)
  lab = torch.tensor(coeffs[:, :, 80:144], device='cuda:0')
  loss = loss_fn(p, lab, seq_len)
  assert abs(float(loss.detach()) - want) <= 1e-5 * abs(want), (float(loss.detach()), want)
  loss.backward()
  got = p.grad.cpu().numpy().astype(np.float64)
  ref = analytic_grad(pred, coeffs, seq_len, small_model, mask)
  assert np.max(np.abs(got - ref)) <= 1e-4 * np.max(np.abs(ref)), np.max(np.abs(got - ref)) / np.max(np.abs(ref))
  # frames beyond a sequence's length get no gradient
  for i in range(b):
    assert not got[i, int(seq_len[i]):].any()
  # central differences of the oracle on a few coordinates (the loss is piecewise linear: exact away from kinks)
  rng = np.random.Generator(np.random.PCG64(3))
  for _ in range(4):
    i, j, k = int(rng.integers(b)), int(rng.integers(min(seq_len))), int(rng.integers(64))
    e = np.zeros_like(pred, dtype=np.float64)
    e[i, j, k] = 1e-4
    fd = (slo.cost(pred + e, coeffs.astype(np.float64), seq_len, small_model, mask) -
          slo.cost(pred - e, coeffs.astype(np.float64), seq_len, small_model, mask)) / 2e-4
    assert abs(fd - got[i, j, k]) <= 2e-2 * max(1.0, abs(fd)), (fd, got[i, j, k])


def test_loss_full_model_matches_oracle_and_is_deterministic(full_model):
  import torch
  b, t = 2, 24                                                      # 48 frames: the tcgen05 basis path
  coeffs, pred, mask, seq_len = make_case(full_model, b, t, seed=77)
  want = slo.cost(pred.astype(np.float64), coeffs.astype(np.float64), seq_len, full_model, mask)
  loss_fn = ExpressionShapeLoss(full_model, mask)
  lab = torch.tensor(coeffs[:, :, 80:144], device='cuda:0')
  outs = []
  for _ in range(2):
    p = torch.tensor(pred, device='cuda:0', requires_grad=True)
    loss = loss_fn(p, lab, seq_len)
    loss.backward()
    outs.append((float(loss.detach()), p.grad.cpu().numpy()))
  assert abs(outs[0][0] - want) <= 1e-5 * abs(want), (outs[0][0], want)
  assert outs[0][0] == outs[1][0] and np.array_equal(outs[0][1], outs[1][1])       # fixed reduction order
  ref = analytic_grad(pred, coeffs, seq_len, full_model, mask)
  assert np.max(np.abs(outs[0][1] - ref)) <= 2e-4 * np.max(np.abs(ref))
