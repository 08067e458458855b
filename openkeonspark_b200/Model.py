"""Model descriptors — the host-side mirror of the reference's Model.py / TransX.py classes.

The reference classes build TensorFlow graphs (input_def / embedding_def / loss_def / predict_def,
/root/reference/Model.py:55-99).  Here a model class only declares WHAT the fused CUDA kernels need:
its id, its parameter tables (same names and shapes as `parameter_lists` in TransE.py:21-24,
TransH.py:26-31, TransR.py:29-35, TransD.py:37-44) and the shape of `predict`.  The math lives in
csrc/train.cu, csrc/score.cu and csrc/transr.cu.
"""
from __future__ import annotations

import math

import torch


class Model(object):
    name = None            # "TransE" | "TransH" | "TransR" | "TransD"
    predict_keepdims = True   # TransE.predict is [N] (reduce_mean, keep_dims=False); the others are [N,1]

    def get_config(self):
        return self.config

    def __init__(self, config, define=True, device=None, seed=None):
        self.config = config
        self.device = torch.device(device if device is not None else "cuda")
        self.parameter_lists = {}
        if define:
            self.embedding_def(seed)

    # tables: name -> (rows, cols)
    def table_shapes(self):
        raise NotImplementedError

    def embedding_def(self, seed=None):
        """Allocate the tables in HBM with tf.contrib.layers.xavier_initializer(uniform=False)
        statistics: truncated normal, stddev sqrt(1.3 * 2 / (rows + cols)).  (TF's random stream is
        not reproducible; parity tests inject parameters with Config.set_parameters.)"""
        gen = torch.Generator(device="cpu")
        gen.manual_seed(0 if seed is None else int(seed))
        for name, (rows, cols) in self.table_shapes().items():
            std = math.sqrt(1.3 * 2.0 / (rows + cols))
            w = torch.empty(rows, cols, dtype=torch.float32)
            torch.nn.init.trunc_normal_(w, mean=0.0, std=std, a=-2 * std, b=2 * std, generator=gen)
            self.parameter_lists[name] = w.to(self.device).contiguous()
