"""
L2-regularised logistic regression fitted ON THE DEVICE -- the classifier of the downstream yardstick
(tools/graph_model_downstream_classification.py:85-91 of the reference: `LogisticRegression(**classifier_params).fit`), for embedding
matrices that should not travel to the host (at 10 M x 128 the host fit is the bottleneck).

Same objective as sklearn's default (`penalty='l2', C=1.0, solver='lbfgs'`, intercept not penalised):
    minimise   C * sum_i loss_i(W, b)  +  1/2 |W|^2        loss = softmax cross-entropy (binary problems: ONE logistic weight vector)
minimised with L-BFGS (two-loop recursion, Armijo backtracking).  A gradient evaluation is two tensor-core GEMMs (`se_gemm_nt`:
logits = X W^T, dW = G^T X through pre-transposed operands) around the elementwise kernel `se_softmax_xent`; the L-BFGS vector
arithmetic is a handful of torch ops on (C x E)-sized vectors.  Checked against sklearn's per-experiment accuracies recorded by the
reference's own `perform_node_classification` (tests/golden/downstream_karate.npz) in tests/test_gpu_text.py... (test_gpu_classifier.py).
"""
from typing import Optional

import torch

from shallow_encoders import _native as nat


class DeviceLogisticRegression:
    def __init__(self, C: float = 1.0, max_iter: int = 100, tol: float = 1e-4, fit_intercept: bool = True, history: int = 10, **_ignored):
        self.C, self.max_iter, self.tol, self.fit_intercept, self.history = float(C), int(max_iter), float(tol), bool(fit_intercept), int(history)
        self.coef_: Optional[torch.Tensor] = None
        self.intercept_: Optional[torch.Tensor] = None
        self.classes_: Optional[torch.Tensor] = None
        self.n_iter_ = 0

    # -- objective ---------------------------------------------------------------------------------------------------
    def _loss_grad(self, x, xt, y, theta, n_cols, want_grad=True):
        n, e = x.shape
        w = theta[:n_cols * e].view(n_cols, e)
        b = theta[n_cols * e:] if self.fit_intercept else None
        logits = nat.gemm_nt(x, w.contiguous())                              # [n, n_cols] on the tensor cores
        loss = torch.zeros(1, dtype=torch.float64, device=x.device)
        nat.softmax_xent(logits, y, b.contiguous() if b is not None else None, self.C, loss_sum=loss)      # logits <- C * dL/dlogits
        value = self.C * float(loss.item()) + 0.5 * float((w.double() ** 2).sum().item())
        if not want_grad:
            return value, None
        gw = nat.gemm_nt(nat.transpose(logits), xt) + w                      # [n_cols, e] = G^T X + W
        grad = gw.reshape(-1)
        if self.fit_intercept:
            grad = torch.cat([grad, logits.sum(dim=0)])
        return value, grad

    def fit(self, x: torch.Tensor, y: torch.Tensor) -> 'DeviceLogisticRegression':
        """x: CUDA float32 [n, E]; y: integer class labels [n] (any integers)."""
        x = x.to(torch.float32).contiguous()
        self.classes_, y_idx = torch.unique(y.to(x.device), return_inverse=True)
        n_classes = int(self.classes_.numel())
        assert n_classes >= 2, 'need at least two classes'
        n_cols = 1 if n_classes == 2 else n_classes
        y_idx = y_idx.to(torch.int32).contiguous()
        xt = nat.transpose(x)
        e = x.shape[1]
        theta = torch.zeros(n_cols * e + (n_cols if self.fit_intercept else 0), dtype=torch.float32, device=x.device)
        f, g = self._loss_grad(x, xt, y_idx, theta, n_cols)
        s_hist, y_hist, rho = [], [], []
        self.n_iter_ = 0
        for it in range(self.max_iter):
            if float(g.abs().max().item()) <= self.tol:
                break
            q = g.clone()
            alphas = []
            for s, yv, r in zip(reversed(s_hist), reversed(y_hist), reversed(rho)):
                a = r * torch.dot(s, q)
                alphas.append(a)
                q -= a * yv
            if s_hist:
                q *= torch.dot(s_hist[-1], y_hist[-1]) / torch.dot(y_hist[-1], y_hist[-1])
            else:
                q *= 1.0 / max(float(g.norm().item()), 1e-12)
            for (s, yv, r), a in zip(zip(s_hist, y_hist, rho), reversed(alphas)):
                q += (a - r * torch.dot(yv, q)) * s
            d = -q
            slope = float(torch.dot(g, d).item())
            if slope >= 0:                                                   # not a descent direction (numerical noise): restart from steepest descent
                d, slope, s_hist, y_hist, rho = -g, -float(torch.dot(g, g).item()), [], [], []
            step, ok = 1.0, False
            for _ in range(30):
                f_new, _ = self._loss_grad(x, xt, y_idx, theta + step * d, n_cols, want_grad=False)
                if f_new <= f + 1e-4 * step * slope:
                    ok = True
                    break
                step *= 0.5
            if not ok:
                break
            theta_new = theta + step * d
            f_new, g_new = self._loss_grad(x, xt, y_idx, theta_new, n_cols)
            s_vec, y_vec = theta_new - theta, g_new - g
            sy = float(torch.dot(s_vec, y_vec).item())
            if sy > 1e-10:
                s_hist.append(s_vec); y_hist.append(y_vec); rho.append(1.0 / sy)
                if len(s_hist) > self.history:
                    s_hist.pop(0); y_hist.pop(0); rho.pop(0)
            converged = abs(f - f_new) <= 1e-12 * max(abs(f), abs(f_new), 1.0)
            theta, f, g = theta_new, f_new, g_new
            self.n_iter_ = it + 1
            if converged:
                break
        self.coef_ = theta[:n_cols * e].view(n_cols, e).clone()
        self.intercept_ = theta[n_cols * e:].clone() if self.fit_intercept else torch.zeros(n_cols, device=x.device)
        return self

    def decision_function(self, x: torch.Tensor) -> torch.Tensor:
        return nat.gemm_nt(x.to(torch.float32).contiguous(), self.coef_.contiguous()) + self.intercept_

    def predict(self, x: torch.Tensor) -> torch.Tensor:
        z = self.decision_function(x)
        idx = (z[:, 0] > 0).long() if z.shape[1] == 1 else z.argmax(dim=1)
        return self.classes_[idx]

    def score(self, x: torch.Tensor, y: torch.Tensor) -> float:
        return float((self.predict(x) == y.to(x.device)).float().mean().item())
