"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's TensorFlow graphs.

TensorFlow 1.x cannot be installed here (no network; the code needs tf.contrib / tf.placeholder),
and the reference ships no golden vectors for losses, scores or updates, so this half of the
oracle is PARITY UNPINNED by the reference: it restates the published TF-1.x op semantics and is
cross-checked only against itself in fp64.  Each function cites the reference lines it follows.

  batch layout ........ Model.py:55-74        (plane-major [1+k+kr, B] -> positives [B,1], negatives [B,k+kr])
  TransE .............. TransE.py:11-15, 26-58
  TransH .............. TransH.py:12-20, 33-82
  TransR .............. TransR.py:16-23, 36-87
  TransD .............. TransD.py:23-31, 46-98 (tf_resize is the identity: all tables use hidden_size)
  optimizer ........... distribute_training.py:94-101 (GradientDescentOptimizer / AdamOptimizer(alpha))

Third-party semantics encoded (TensorFlow 1.x, source not under /root/reference):
  tf.nn.l2_normalize(x) = x * rsqrt(max(sum x^2, 1e-12)); tf.abs' = sign; tf.maximum(x, 0)' = [x >= 0];
  reduce_mean over B*(k+kr); embedding_lookup gradients are IndexedSlices whose duplicate rows are
  summed; sparse SGD: row -= lr * sum; sparse Adam (tf.train.AdamOptimizer._apply_sparse_shared):
  m <- b1*m (dense), m[idx] += (1-b1) g; v <- b2*v (dense), v[idx] += (1-b2) g^2;
  var -= lr_t * m / (sqrt(v) + eps) (dense), lr_t = lr*sqrt(1-b2^t)/(1-b1^t), b1=.9 b2=.999 eps=1e-8.
"""
from __future__ import annotations

import numpy as np
import torch

NAMES = {
    "TransE": ("ent_embeddings", "rel_embeddings"),
    "TransH": ("ent_embeddings", "rel_embeddings", "normal_vectors"),
    "TransR": ("ent_embeddings", "rel_embeddings", "transfer_matrix"),
    "TransD": ("ent_embeddings", "rel_embeddings", "ent_transfer", "rel_transfer"),
}


def _l2n(x):
    ss = (x * x).sum(-1, keepdim=True)
    return x * torch.rsqrt(torch.clamp(ss, min=1e-12))


def _calc(h, t, r):
    return torch.abs(_l2n(h) + _l2n(r) - _l2n(t))


def _project(model, P, h, t, r, r_for_matrix=None):
    """Embedding lookups + the model's transfer; returns (h', t', r') ready for _calc."""
    ent, rel = P["ent_embeddings"], P["rel_embeddings"]
    he, te, re = ent[h], ent[t], rel[r]
    if model == "TransE":
        return he, te, re
    if model == "TransH":
        n = _l2n(P["normal_vectors"][r])
        return (he - (he * n).sum(-1, keepdim=True) * n, te - (te * n).sum(-1, keepdim=True) * n, re)
    if model == "TransD":
        et, rt = P["ent_transfer"], P["rel_transfer"][r]
        return (he + (he * et[h]).sum(-1, keepdim=True) * rt, te + (te * et[t]).sum(-1, keepdim=True) * rt, re)
    if model == "TransR":
        De, Dr = ent.shape[1], rel.shape[1]
        rm = r if r_for_matrix is None else r_for_matrix
        M = P["transfer_matrix"][rm].reshape(rm.shape + (De, Dr))
        return (torch.matmul(he.unsqueeze(-2), M).squeeze(-2), torch.matmul(te.unsqueeze(-2), M).squeeze(-2), re)
    raise ValueError(model)


def loss_fn(model, P, bh, bt, br, B, k, kr, margin):
    """loss_def of the four models on a plane-major batch (Model.py:62-69)."""
    K = k + kr
    bh, bt, br = (torch.as_tensor(np.asarray(x), dtype=torch.long) for x in (bh, bt, br))
    ph, pt, pr = bh[:B].reshape(B, 1), bt[:B].reshape(B, 1), br[:B].reshape(B, 1)
    nh, nt, nr = (x[B:].reshape(K, B).t() for x in (bh, bt, br))
    # TransR.py:57-60: with negative_rel == 0 the negatives are projected by the POSITIVE's matrix
    rm = pr.expand(B, K) if (model == "TransR" and kr == 0) else None
    p = _calc(*_project(model, P, ph, pt, pr)).sum(-1, keepdim=True)
    n = _calc(*_project(model, P, nh, nt, nr, rm)).sum(-1, keepdim=True)
    x = p - n + margin
    return torch.where(x >= 0, x, torch.zeros_like(x)).mean()


def predict_fn(model, P, h, t, r):
    """predict_def: TransE = mean over d -> [N]; H/R/D = sum keepdims -> [N,1]."""
    h, t, r = (torch.as_tensor(np.asarray(x), dtype=torch.long) for x in (h, t, r))
    if model == "TransR":
        rm = r[0].expand(r.shape)         # TransR.py:83 uses predict_r[0]'s matrix for every row
        s = _calc(*_project(model, P, h, t, r, rm))
    else:
        s = _calc(*_project(model, P, h, t, r))
    return s.mean(1) if model == "TransE" else s.sum(-1, keepdim=True)


class Trainer:
    """One worker's view of distribute_training.py:267-283: step(batch) = train_op + loss."""

    def __init__(self, model, params, margin=1.0, lr=0.001, opt="SGD", dtype=torch.float32):
        self.model, self.margin, self.lr, self.opt = model, margin, lr, opt
        self.dtype = dtype
        self.P = {k: torch.tensor(np.asarray(params[k]), dtype=dtype).requires_grad_(True) for k in NAMES[model]}
        self.t = 0
        if opt.lower() == "adam":
            self.m = {k: torch.zeros_like(v) for k, v in self.P.items()}
            self.v = {k: torch.zeros_like(v) for k, v in self.P.items()}
            self.b1p = torch.ones((), dtype=dtype)
            self.b2p = torch.ones((), dtype=dtype)

    def grads(self, bh, bt, br, B, k, kr):
        for v in self.P.values():
            v.grad = None
        loss = loss_fn(self.model, self.P, bh, bt, br, B, k, kr, self.margin)
        loss.backward()
        return float(loss.detach()), {n: (v.grad if v.grad is not None else torch.zeros_like(v)) for n, v in self.P.items()}

    def step(self, bh, bt, br, B, k, kr):
        loss, g = self.grads(bh, bt, br, B, k, kr)
        self.t += 1
        with torch.no_grad():
            if self.opt.lower() == "adam":
                b1, b2, eps = 0.9, 0.999, 1e-8
                # beta powers and lr_t are fp32 tensors in TF; keep them in the working dtype
                one = torch.ones((), dtype=self.dtype)
                # TF keeps beta1_power/beta2_power as variables multiplied once per step (AdamOptimizer._finish)
                self.b1p = self.b1p * torch.tensor(b1, dtype=self.dtype)
                self.b2p = self.b2p * torch.tensor(b2, dtype=self.dtype)
                b1p, b2p = self.b1p, self.b2p
                lr_t = torch.tensor(self.lr, dtype=self.dtype) * torch.sqrt(one - b2p) / (one - b1p)
                for n, p in self.P.items():
                    self.m[n].mul_(b1).add_(g[n] * (1 - b1))
                    self.v[n].mul_(b2).add_(g[n] * g[n] * (1 - b2))
                    p.sub_(lr_t * self.m[n] / (torch.sqrt(self.v[n]) + eps))
            else:
                for n, p in self.P.items():
                    p.sub_(self.lr * g[n])
        return loss

    def params(self):
        return {k: v.detach().numpy().copy() for k, v in self.P.items()}

    def predict(self, h, t, r):
        with torch.no_grad():
            return predict_fn(self.model, self.P, h, t, r).numpy()
