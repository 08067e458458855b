"""Config — the user API of the reference (/root/reference/Config.py), re-backed by libokb200.so.

Same class name, same setter names and positional arguments, same public attributes; the
TensorFlow session and the ctypes calls into release/Base.so are replaced by PyTorch-owned device
memory and the hand-written sm_100a kernels behind include/okb200.h.  What the reference only
implements inside distribute_training.py is provided here as methods:

    run() / train()            the train loop of distribute_training.py:267-283
    test()                     triple classification (Config.py:491-516) AND, when
                               set_test_link_prediction(True), the link-prediction evaluation of
                               distribute_training.py:464-607 + the tables of main_spark.py:447-469

There is no CPU fallback: a missing library or GPU raises.
"""
from __future__ import annotations

import ctypes
import json
import os
import time

import numpy as np
import torch

from . import _native, metrics
from ._native import MODEL_ID, OkbError, okb_hyper, okb_model

_vp = ctypes.c_void_p
_AUX_ENT = {"TransD": "ent_transfer"}
_AUX_REL = {"TransH": "normal_vectors", "TransR": "transfer_matrix", "TransD": "rel_transfer"}


def _addr(a):
    return _vp(a.__array_interface__["data"][0])


def _stream():
    return _vp(torch.cuda.current_stream().cuda_stream)


class Config(object):
    """Set the essential parameters, load the data, train and evaluate (reference Config.py:12)."""

    def __init__(self, cpp_lib_path=None, init_new_entities=False, private_context=False):
        self.init_new_entities = init_new_entities
        # `lib` is the raw C library like the reference's (Config.py:30); its Base.so-compatible
        # symbols act on the same process-global context this Config uses.
        self.lib = _native.load(cpp_lib_path)
        self.ctx = _native.Ctx(self.lib, default=not private_context)
        for env, flag, val in (("OKB200_PDL", 2, 0), ("OKB200_ADAM_TMA", 3, 1), ("OKB200_L2_PREFETCH", 4, 1), ("OKB200_ADAM_LEGACY", 5, 1), ("OKB200_GRAD_GENERIC", 6, 1), ("OKB200_DP_PULL", 7, 1), ("OKB200_GRAD_SINGLE_WARP", 8, 1), ("OKB200_PLAN_MULTI", 9, 1), ("OKB200_CHUNK_KERNEL", 10, 1), ("OKB200_TRANSR_FUSED", 12, 1)):
            if os.environ.get(env) == str(val):      # A/B switches for measurements (include/okb200.h OKB_FLAG_*)
                self.ctx.call("okb_set_flag", flag, val)
        if os.environ.get("OKB200_ADAM_VPT"):
            self.ctx.call("okb_set_flag", 11, int(os.environ["OKB200_ADAM_VPT"]))
        self.in_path = None
        self.out_path = None
        self.bern = 0
        self.hidden_size = 64
        self.ent_size = self.hidden_size
        self.rel_size = self.hidden_size
        self.train_times = 0
        self.margin = 1.0
        self.nbatches = 0
        self.negative_ent = 1
        self.negative_rel = 0
        self.workThreads = 8
        self.alpha = 0.001
        self.exportName = None
        self.importName = None
        self.opt_method = "SGD"
        self.test_link_prediction = False
        self.test_triple_classification = False
        self.valid_triple_classification = False
        # additions (not in the reference)
        self.test_head = 0            # main_spark.py --test_head
        self.log_every = 0            # read the loss back every n steps in run(); 0 = once per epoch
        self.seed = 0
        self.trainModel = None
        self.model = None
        self._step = 0
        self._adam = None
        self._world = None            # parallel.DataParallel when running under torch.distributed
        self.transr_tensor_cores = False   # TransR ranking: project candidates with tcgen05 (3xTF32) instead of canonical fp32
        self.plan_ahead = 64          # steps sampled + planned per launch chunk (integer work, parameter-independent)
        # early stopping (main_spark.py --early_stop_patience / --early_stop_start_step / --early_stop_stopping_step)
        self.early_stop_patience = 0  # 0 = off
        self.early_stop_start_step = 0        # epochs before the first check
        self.early_stop_stopping_step = 1     # epochs between checks
        self.early_stop = None        # after run(): {"reason", "best_step", "best_acc", "best_loss", "checks"} or None
        self._chunk_pos = self._chunk_len = 0

    # ------------------------------------------------------------------ init (Config.py:74-186)
    def init_link_prediction(self):
        self.ctx.call("okb_import_test_files")
        self.ctx.call("okb_import_type_files")
        self.ctx.call("okb_import_ontology_files")
        self.testTotal = self.ctx.total(4)
        self.validTotal = self.ctx.total(5)

    def _alloc_tc(self, with_test):
        n_t, n_v = self.ctx.total(4), self.ctx.total(5)
        self.testTotal, self.validTotal = n_t, n_v
        if with_test:
            for nm in ("test_pos_h", "test_pos_t", "test_pos_r", "test_neg_h", "test_neg_t", "test_neg_r"):
                setattr(self, nm, np.zeros(n_t, dtype=np.int64))
                setattr(self, nm + "_addr", getattr(self, nm).__array_interface__["data"][0])
        for nm in ("valid_pos_h", "valid_pos_t", "valid_pos_r", "valid_neg_h", "valid_neg_t", "valid_neg_r"):
            setattr(self, nm, np.zeros(n_v, dtype=np.int64))
            setattr(self, nm + "_addr", getattr(self, nm).__array_interface__["data"][0])
        self.relThresh = np.zeros(self.ctx.total(1), dtype=np.float32)
        self.relThresh_addr = self.relThresh.__array_interface__["data"][0]
        self.acc = np.zeros(1, dtype=np.float32)
        self.acc_addr = self.acc.__array_interface__["data"][0]

    def init_triple_classification(self):
        self.ctx.call("okb_import_test_files")
        self.ctx.call("okb_import_type_files")
        self._alloc_tc(True)

    def init_valid_triple_classification(self):
        self.ctx.call("okb_import_test_files")
        self.ctx.call("okb_import_type_files")
        self._alloc_tc(False)

    def init(self):
        """prepare for train and test (Config.py:153-186)"""
        if self.init_new_entities:
            return
        if not torch.cuda.is_available():
            raise OkbError("no CUDA device: openkeonspark_b200 has no CPU path")
        self.trainModel = None
        if self.in_path is not None:
            self.ctx.call("okb_set_device", torch.cuda.current_device())
            self.ctx.call("okb_set_in_path", self.in_path.encode())
            self.ctx.call("okb_set_bern", self.bern)
            self.ctx.call("okb_set_work_threads", self.workThreads)
            self.ctx.call("okb_rand_reset")
            self.ctx.call("okb_import_train_files")
            self._after_train_import()
        if self.test_link_prediction:
            self.init_link_prediction()
        if self.test_triple_classification:
            self.init_triple_classification()
        if self.valid_triple_classification:
            self.init_valid_triple_classification()

    def init_from_arrays(self, n_ent, n_rel, train, valid=None, test=None, new_batch=0):
        """Same as init() but from [n,3] int64 arrays with columns h, t, r (large synthetic graphs)."""
        if not torch.cuda.is_available():
            raise OkbError("no CUDA device: openkeonspark_b200 has no CPU path")
        cols = lambda a: [np.ascontiguousarray(a[:, k], dtype=np.int64) for k in range(3)]
        h, t, r = cols(train)
        self.ctx.call("okb_set_device", torch.cuda.current_device())
        self.ctx.call("okb_set_bern", self.bern)
        self.ctx.call("okb_set_work_threads", self.workThreads)
        self.ctx.call("okb_rand_reset")
        self.ctx.call("okb_import_train_arrays", n_ent, n_rel, _addr(h), _addr(t), _addr(r), h.size, new_batch)
        self._after_train_import()
        if test is not None:
            a, b = cols(test), cols(valid)
            self.ctx.call("okb_import_test_arrays", _addr(a[0]), _addr(a[1]), _addr(a[2]), a[0].size,
                          _addr(b[0]), _addr(b[1]), _addr(b[2]), b[0].size)
            self.testTotal, self.validTotal = self.ctx.total(4), self.ctx.total(5)
            self.ctx.call("okb_build_type_constraints")       # what main_spark.n_n() writes to type_constrain.txt

    def _after_train_import(self):
        self.relTotal = self.ctx.total(1)
        self.entTotal = self.ctx.total(0)
        self.trainTotal = self.ctx.total(2)
        self.testTotal = self.ctx.total(4)
        self.validTotal = self.ctx.total(5)
        self.bt = self.ctx.total(7)
        self.set_mini_batch()
        self._alloc_batch()

    def _alloc_batch(self):
        """(Re)allocate the caller-visible numpy batch buffers (Config.py:172-180)."""
        self.batch_seq_size = S = self.batch_size * (1 + self.negative_ent + self.negative_rel)
        # page-locked and contiguous, so sampling()/train_step() move the three id arrays with ONE DMA each way
        self._batch_pinned = torch.zeros(3 * S, dtype=torch.int64).pin_memory()
        self._batch_y_pinned = torch.zeros(S, dtype=torch.float32).pin_memory()
        ids = self._batch_pinned.numpy()
        self.batch_h, self.batch_t, self.batch_r = ids[:S], ids[S:2 * S], ids[2 * S:]
        self.batch_y = self._batch_y_pinned.numpy()
        self.batch_h_addr = self.batch_h.__array_interface__["data"][0]
        self.batch_t_addr = self.batch_t.__array_interface__["data"][0]
        self.batch_r_addr = self.batch_r.__array_interface__["data"][0]
        self.batch_y_addr = self.batch_y.__array_interface__["data"][0]

    def set_mini_batch(self):
        """Config.py:189-210 (note the truncating division: tot % nbatches samples are dropped)."""
        tot = self.bt if self.bt > 0 else self.trainTotal
        if self.nbatches > 0:
            self.batch_size = int(tot / self.nbatches)
        else:
            self.batch_size = tot
            while self.batch_size > 9999:
                self.batch_size = int(self.batch_size / 10)
            self.nbatches = int(tot / self.batch_size)
        print("Batch size is {}".format(self.batch_size))
        print("Number of batches: {}".format(self.nbatches))

    # ------------------------------------------------------------------ setters (Config.py:213-340)
    def get_ent_total(self):
        return self.entTotal

    def get_rel_total(self):
        return self.relTotal

    def set_opt_method(self, method):
        self.opt_method = method

    def set_test_link_prediction(self, flag):
        self.test_link_prediction = flag

    def set_test_triple_classification(self, flag):
        self.test_triple_classification = flag

    def set_valid_triple_classification(self, flag):
        self.valid_triple_classification = flag

    def set_alpha(self, alpha):
        self.alpha = alpha

    def set_in_path(self, path):
        self.in_path = path

    def set_out_files(self, path):
        self.out_path = path

    def set_bern(self, bern):
        self.bern = bern

    def set_dimension(self, dim):
        self.hidden_size = dim
        self.ent_size = dim
        self.rel_size = dim

    def set_ent_dimension(self, dim):
        self.ent_size = dim

    def set_rel_dimension(self, dim):
        self.rel_size = dim

    def set_train_times(self, times):
        self.train_times = times

    def set_nbatches(self, nbatches):
        self.nbatches = nbatches

    def set_margin(self, margin):
        self.margin = margin

    def set_ent_neg_rate(self, rate):
        self.negative_ent = rate

    def set_rel_neg_rate(self, rate):
        self.negative_rel = rate

    def set_import_files(self, path):
        self.importName = path

    def set_export_files(self, path):
        self.exportName = path

    def set_test_head(self, flag):
        self.test_head = int(flag)

    def set_early_stopping(self, patience, start_step=0, stopping_step=1):
        """Early stopping as in distribute_training.py:253-360: from epoch `start_step` on, every `stopping_step`
        epochs, fit the per-relation thresholds on the valid triples, measure their accuracy and look at the last
        batch loss; stop when either has not improved `patience` checks in a row and restore the best model
        (main_spark.py:347-376).  Needs set_valid_triple_classification(True) before init()."""
        self.early_stop_patience = int(patience)
        self.early_stop_start_step = int(start_step)
        self.early_stop_stopping_step = max(1, int(stopping_step))

    # ------------------------------------------------------------------ sampling (Config.py:343-347)
    def sampling(self):
        """One reference sampling() call on the GPU; results land in batch_h/t/r/y like the reference's."""
        # labels are the fixed pattern of Base.cpp:110,129,138 (+1 for the B positives, -1 for every negative plane):
        # written here instead of crossing PCIe as a second copy
        self.batch_y[:self.batch_size] = 1.0
        self.batch_y[self.batch_size:] = -1.0
        # one launch: the sample kernel stores the int64 batch straight into the page-locked batch_h/t/r block
        self.ctx.call("okb_sample_to_host", self.batch_size, self.negative_ent, self.negative_rel, 0, self.workThreads,
                      _vp(self.batch_h_addr), _vp(self.batch_t_addr), _vp(self.batch_r_addr), _stream())
        self._chunk_pos = self._chunk_len = 0

    def sampling_device(self, steps=1):
        """Sample `steps` consecutive batches, leaving them resident in HBM (no host copy)."""
        self.ctx.call("okb_sample", self.batch_size, self.negative_ent, self.negative_rel, steps, 0, self.workThreads, _stream())
        self._chunk_pos = self._chunk_len = 0

    def next_step_device(self):
        """One iteration of the train loop (sampling() + train_op, distribute_training.py:274-282).
        Sampling and gradient-row planning do not depend on the parameters, so they are done for
        `plan_ahead` consecutive steps in one launch each; the batches are exactly the ones that many
        consecutive sampling() calls would have produced."""
        if self._world is not None and self._world.mode == "owner":
            return self._world.next_step(self)
        if self._chunk_pos >= self._chunk_len:
            n = self.batch_size * (3 + self.negative_ent + self.negative_rel)
            C = max(1, min(int(self.plan_ahead), (1 << 24) // n))
            if self._world is None:
                # current chunk: a prefetched one if it matches, else sampled + planned now; then start the next one on the
                # library's side stream — it overlaps the train kernels of this chunk
                a = (self.batch_size, self.negative_ent, self.negative_rel, C, _stream())
                self.ctx.call("okb_chunk_begin", *a)
                self.ctx.call("okb_chunk_prefetch", *a)
            else:
                self.sampling_device(C)
                self.ctx.call("okb_plan_steps", 0, C, _stream())
            self._chunk_pos, self._chunk_len = 0, C
        loss = self.train_step_device(self._chunk_pos)
        self._chunk_pos += 1
        return loss

    # ------------------------------------------------------------------ data-parallel hygiene
    def _owner_mode(self):
        return self._world is not None and self._world.mode == "owner"

    def _rank(self):
        return 0 if self._world is None else self._world.rank

    def _settle(self, adam=False):
        """Owner-sharded data parallelism: before anything READS the tables, wait until every peer's last row updates have
        landed here (a rank's final update kernel stores into its peers' arenas).  adam=True also completes the Adam
        slots (a rank only maintains m / v for the rows it owns) — a collective, for checkpoints and snapshots."""
        if self._owner_mode():
            self._world.quiesce(self)
            if adam:
                self._world.sync_adam_slots(self)
        elif self._world is not None and self._world.mode == "relation":
            self._world.gather_relations(self, adam)      # TransR: relation rows live with their owners (a collective)

    def _barrier(self):
        if self._world is not None:
            import torch.distributed as dist
            dist.barrier(group=self._world.group)

    # ------------------------------------------------------------------ parameters (Config.py:379-422)
    def get_parameter_lists(self):
        return self.trainModel.parameter_lists

    def get_parameters_by_name(self, var_name):
        if var_name in self.trainModel.parameter_lists:
            return self.trainModel.parameter_lists[var_name].detach().cpu().numpy()
        return None

    def get_parameters(self, mode="numpy"):
        self._settle()                                # peers' last row updates have landed (not a collective)
        res = {}
        for var_name in self.get_parameter_lists():
            v = self.get_parameters_by_name(var_name)
            res[var_name] = v if mode == "numpy" else v.tolist()
        return res

    def save_parameters(self, path=None):
        if path is None:
            path = self.out_path
        with open(path, "w") as f:
            f.write(json.dumps(self.get_parameters("list")))

    def set_parameters_by_name(self, var_name, tensor, allow_partial_rows=False):
        """Config.py:414-418 (tf.assign: a shape mismatch raises).  allow_partial_rows=True is the incremental-batch
        case only: a model trained before new entities arrived lands in the leading rows of the grown table."""
        if var_name in self.trainModel.parameter_lists:
            dst = self.trainModel.parameter_lists[var_name]
            src = torch.as_tensor(np.asarray(tensor, dtype=np.float32))
            partial = (src.dim() == 2 and src.shape[1] == dst.shape[1] and src.shape[0] < dst.shape[0]
                       and var_name in ("ent_embeddings", "ent_transfer"))
            if partial and allow_partial_rows:
                # main_spark.py:71-75: scatter_update of the old rows into a tensor of the final shape; the new entities
                # keep their fresh Xavier-normal rows
                dst[:src.shape[0]].copy_(src)
            elif src.numel() != dst.numel():
                raise OkbError("set_parameters: %s has shape %s, the table is %s" % (var_name, tuple(src.shape), tuple(dst.shape)))
            else:
                dst.copy_(src.reshape(dst.shape))

    def set_parameters(self, lists, allow_partial_rows=False):
        for i in lists:
            self.set_parameters_by_name(i, lists[i], allow_partial_rows)

    def save_tensorflow(self):
        """Reference: Saver.save (Config.py:350-356).  Here: a torch checkpoint of tables + optimizer state.
        Under data parallelism this is a collective (every rank calls it); rank 0 writes the file."""
        self._settle(adam=True)
        if self._rank() == 0:
            state = {"params": {k: v.cpu() for k, v in self.trainModel.parameter_lists.items()}, "step": self._step,
                     "adam": None if self._adam is None else {k: v.cpu() if torch.is_tensor(v) else float(v) for k, v in self._adam.items()}}
            torch.save(state, self.exportName)
        self._barrier()

    def save_tensorflow_weights(self, export_name=None, write_meta_graph=False):
        """Config.py:358-364 (Saver.save to an explicit path; the meta graph has no counterpart here)."""
        keep, self.exportName = self.exportName, (self.exportName if export_name is None else export_name)
        try:
            self.save_tensorflow()
        finally:
            self.exportName = keep

    def import_model(self, ckpt):
        """Config.py:366-376 (Saver.restore from an explicit checkpoint path)."""
        self._ensure_model()
        keep, self.importName = self.importName, ckpt
        try:
            self.restore_tensorflow()
        finally:
            self.importName = keep

    def restore_tensorflow(self):
        """Saver.restore (Config.py:366-376).  Under data parallelism a collective: every rank loads the same file."""
        self._settle()                                # no peer may still be storing rows into the tables being overwritten
        self._barrier()
        state = torch.load(self.importName, map_location="cpu")
        self.set_parameters({k: v.numpy() for k, v in state["params"].items()}, allow_partial_rows=True)
        self._step = state.get("step", 0)
        if state.get("adam") is not None and self._adam is not None:
            for k, v in state["adam"].items():
                if torch.is_tensor(v):
                    if v.shape[0] < self._adam[k].shape[0]:       # grown entity table: new rows keep zero slots (main_spark.py:77-81)
                        self._adam[k].zero_()
                        self._adam[k][:v.shape[0]].copy_(v)
                    else:
                        self._adam[k].copy_(v)
                else:
                    self._adam[k] = np.float32(v)     # beta powers are fp32 running products (tf.train.AdamOptimizer)
        if self._world is not None:
            torch.cuda.synchronize()
            self._barrier()                           # every rank's tables are in place before anyone steps (and pushes rows)

    # ------------------------------------------------------------------ model (Config.py:425-461)
    def set_model(self, model):
        self.model = model

    def set_model_and_session(self, model):
        """Allocate the model's tables in HBM (the reference builds a TF graph + session here)."""
        self.model = model
        self.trainModel = self.model(config=self, define=True, seed=self.seed)
        self._step = 0
        self._adam = None
        if self.opt_method.lower() == "adam":
            self._adam = {"b1p": np.float32(1.0), "b2p": np.float32(1.0)}
            for k, v in self.trainModel.parameter_lists.items():
                self._adam["m_" + k] = torch.zeros_like(v)
                self._adam["v_" + k] = torch.zeros_like(v)
        self._loss_dev = torch.zeros(1, dtype=torch.float32, device=self.trainModel.device)
        self._model_struct = None

    def _ensure_model(self):
        if self.trainModel is None:
            if self.model is None:
                raise OkbError("call set_model / set_model_and_session first")
            self.set_model_and_session(self.model)

    def _cmodel(self):
        """okb_model for the current tables (device pointers stay valid: tables are updated in place)."""
        if self._model_struct is not None:
            return self._model_struct
        P = self.trainModel.parameter_lists
        name = self.trainModel.name
        # the kernels index the tables with the context's entity / relation ids: shapes must agree (a dataset re-imported
        # after growing needs grow_entities() first)
        E, R = self.ctx.total(0), self.ctx.total(1)
        if E and (P["ent_embeddings"].shape[0] != E or P["rel_embeddings"].shape[0] != R):
            raise OkbError("tables have %d entity / %d relation rows, the loaded dataset has %d / %d (grow_entities() after a "
                           "re-import?)" % (P["ent_embeddings"].shape[0], P["rel_embeddings"].shape[0], E, R))
        m = okb_model()
        m.model = MODEL_ID[name]
        m.ent_dim = P["ent_embeddings"].shape[1]
        m.rel_dim = P["rel_embeddings"].shape[1]
        m.optimizer = 1 if self._adam is not None else 0
        ptr = lambda t: t.data_ptr() if t is not None else None
        m.ent, m.rel = ptr(P["ent_embeddings"]), ptr(P["rel_embeddings"])
        ae, ar = _AUX_ENT.get(name), _AUX_REL.get(name)
        m.ent_aux = ptr(P[ae]) if ae else None
        m.rel_aux = ptr(P[ar]) if ar else None
        if self._adam is not None:
            A = self._adam
            m.m_ent, m.v_ent = ptr(A["m_ent_embeddings"]), ptr(A["v_ent_embeddings"])
            m.m_rel, m.v_rel = ptr(A["m_rel_embeddings"]), ptr(A["v_rel_embeddings"])
            if ae:
                m.m_ent_aux, m.v_ent_aux = ptr(A["m_" + ae]), ptr(A["v_" + ae])
            if ar:
                m.m_rel_aux, m.v_rel_aux = ptr(A["m_" + ar]), ptr(A["v_" + ar])
        self._model_struct = m
        return m

    def _hypers(self, n=1):
        """Hyper-parameters of the next n steps and the beta powers after them.  The powers are COMMITTED by the caller
        (_commit_powers) only once the native call has succeeded, so a failed call leaves the Adam state in step."""
        out = []
        powers = None
        if self._adam is not None:
            b1p, b2p = self._adam["b1p"], self._adam["b2p"]
        for _ in range(n):
            hp = okb_hyper()
            hp.margin = float(self.margin)
            hp.beta1, hp.beta2, hp.eps = 0.9, 0.999, 1e-8
            if self._adam is not None:
                # tf.train.AdamOptimizer: beta powers are fp32 variables multiplied once per step;
                # lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t), all in fp32.
                b1p = np.float32(b1p * np.float32(0.9))
                b2p = np.float32(b2p * np.float32(0.999))
                one = np.float32(1.0)
                hp.lr = float(np.float32(self.alpha) * np.sqrt(one - b2p, dtype=np.float32) / (one - b1p))
                powers = (b1p, b2p)
            else:
                hp.lr = float(self.alpha)
            out.append(hp)
        return out, powers

    def _commit_powers(self, powers, steps):
        if powers is not None:
            self._adam["b1p"], self._adam["b2p"] = powers
        self._step += steps

    # ------------------------------------------------------------------ training
    def train_step_device(self, step=0):
        """loss_def + optimizer on batch `step` of the last sampling_device(); loss stays on the GPU."""
        self._ensure_model()
        m = self._cmodel()
        (hp,), powers = self._hypers(1)
        if self._world is not None and self._world.mode == "owner":
            self.ctx.call("okb_dp_train_steps", ctypes.byref(m), ctypes.byref(hp), step, 1, _vp(self._loss_dev.data_ptr()), _stream())
        elif self._world is not None:
            self._world.train_step(self, m, hp, step)
        else:
            self.ctx.call("okb_train_step", ctypes.byref(m), ctypes.byref(hp), step, _vp(self._loss_dev.data_ptr()), _stream())
        self._commit_powers(powers, 1)
        return self._loss_dev

    def train_chunk_device(self, n=None, n_next=None):
        """`n` iterations of the train loop in ONE library call: sample n batches, plan them with one sort,
        then n x (grad + update).  Returns the device tensor of the n losses.  Same results as n calls of
        next_step_device(); exists because a Python round trip per step costs more than the step."""
        self._ensure_model()
        cap = max(1, (1 << 24) // (self.batch_size * (3 + self.negative_ent + self.negative_rel)))
        n = max(1, min(int(self.plan_ahead if n is None else n), cap))
        if self._world is not None and self._world.mode == "owner":
            return self._world.train_chunk(self, n)
        if self._world is not None:
            return torch.stack([self.next_step_device().clone() for _ in range(n)]).reshape(-1)
        a = (self.batch_size, self.negative_ent, self.negative_rel)
        self.ctx.call("okb_chunk_begin", *a, n, _stream())
        n_next = n if n_next is None else max(0, min(int(n_next), cap))
        if n_next > 0:                                # sample + plan the next chunk on the side stream while this one trains
            self.ctx.call("okb_chunk_prefetch", *a, n_next, _stream())
        m = self._cmodel()
        hl, powers = self._hypers(n)
        hps = (okb_hyper * n)(*hl)
        if getattr(self, "_loss_chunk", None) is None or self._loss_chunk.numel() < n:
            self._loss_chunk = torch.zeros(n, dtype=torch.float32, device=self.trainModel.device)
        self.ctx.call("okb_train_steps", ctypes.byref(m), hps, 0, n, _vp(self._loss_chunk.data_ptr()), _stream())
        self._commit_powers(powers, n)
        self._chunk_pos = self._chunk_len = 0
        return self._loss_chunk[:n]

    def train_step(self, batch_h, batch_t, batch_r, batch_y):
        """Perform a single training step on a caller-provided batch (Config.py:464-475)."""
        self._ensure_model()
        h = np.ascontiguousarray(batch_h, dtype=np.int64)
        t = np.ascontiguousarray(batch_t, dtype=np.int64)
        r = np.ascontiguousarray(batch_r, dtype=np.int64)
        if h.size != self.batch_seq_size:
            raise OkbError("batch has %d rows, expected batch_seq_size=%d" % (h.size, self.batch_seq_size))
        if self._world is not None:
            self.ctx.call("okb_batch_from_host", self.batch_size, self.negative_ent, self.negative_rel, _addr(h), _addr(t), _addr(r), _stream())
            self.ctx.call("okb_batch_check", _stream())       # ids outside the tables raise here, before any rank steps
            self._chunk_pos = self._chunk_len = 0
            return float(self.train_step_device(0).item())
        # one library call; the loss comes back through one page-locked host word the update kernel stores into (no
        # device->host copy) and the call returns as soon as it is there, while the table update is still running
        if getattr(self, "_loss_pin", None) is None:
            self._loss_pin = torch.zeros(1, dtype=torch.float32).pin_memory()
            self._loss_pin_np = self._loss_pin.numpy()
        m = self._cmodel()
        (hp,), powers = self._hypers(1)
        self._chunk_pos = self._chunk_len = 0
        self.ctx.call("okb_train_step_host", ctypes.byref(m), ctypes.byref(hp), self.batch_size, self.negative_ent, self.negative_rel,
                      _addr(h), _addr(t), _addr(r), _vp(self._loss_pin.data_ptr()), _stream())
        self._commit_powers(powers, 1)
        return float(self._loss_pin_np[0])

    def _snapshot(self):
        """Device-side copy of everything a reference checkpoint holds: tables, Adam slots, beta powers, step.
        (Collective under owner-sharded data parallelism: the Adam slots are completed first.)"""
        self._settle(adam=True)
        snap = {"params": {k: v.clone() for k, v in self.trainModel.parameter_lists.items()}, "step": self._step, "adam": None}
        if self._adam is not None:
            snap["adam"] = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in self._adam.items()}
        return snap

    def _restore_snapshot(self, snap):
        self._settle()                                # no peer may still be storing rows into the tables being overwritten
        self._barrier()
        for k, v in snap["params"].items():
            self.trainModel.parameter_lists[k].copy_(v)
        if snap["adam"] is not None:
            for k, v in snap["adam"].items():
                if torch.is_tensor(v):
                    self._adam[k].copy_(v)
                else:
                    self._adam[k] = v
        self._step = snap["step"]

    def valid_accuracy(self):
        """Thresholds fitted on the valid triples + their accuracy there (the early-stop check).  The valid negatives
        are generated once per run() like the reference's single getValidBatch call (distribute_training.py:262)."""
        if getattr(self, "_valid_batch_ready", False) is False:
            self.ctx.call("okb_tc_batch", 1, *[_vp(getattr(self, "valid_" + n + "_addr")) for n in ("pos_h", "pos_t", "pos_r", "neg_h", "neg_t", "neg_r")])
            self._valid_batch_ready = True
        th, vp, vn = self._fit_thresholds_device()
        self.ctx.call("okb_tc_counts_dev", _vp(th.data_ptr()), _vp(vp.data_ptr()), _vp(vn.data_ptr()), 1, None, _vp(self.acc_addr), _stream())
        return float(self.acc[0])

    def run(self):
        """The train loop of distribute_training.py:267-283: train_times x nbatches x [sampling, step], with the
        early-stop checks of :286-356 when set_early_stopping() was called."""
        self._ensure_model()
        losses = []
        es = self.early_stop_patience > 0
        self.early_stop = None
        if es:
            if not hasattr(self, "valid_pos_h"):
                raise OkbError("early stopping needs set_valid_triple_classification(True) before init()")
            self._valid_batch_ready = False
            best_acc, best_loss = float(np.finfo("float32").min), float(np.finfo("float32").max)
            wait_acc = wait_loss = 0
            best_snap_acc = best_snap_loss = None
            step0 = self._step
            stopping = self.early_stop_stopping_step * self.nbatches
            to_reach = self.early_stop_start_step * self.nbatches + step0
            checks = []
        stop = False
        for epoch in range(self.train_times):
            t0 = time.time()
            acc = torch.zeros(1, dtype=torch.float32, device=self._loss_dev.device)
            batch = 0
            while batch < self.nbatches and not stop:
                n = min(self.plan_ahead, self.nbatches - batch)
                if es and self._step < to_reach:
                    n = min(n, to_reach - self._step)  # a check falls right after step `to_reach`
                left = self.nbatches - batch - n
                nxt = min(self.plan_ahead, left) if left > 0 else (min(self.plan_ahead, self.nbatches) if epoch + 1 < self.train_times else 0)
                if es:
                    nxt = 0                           # early-stop checks cut chunks at data-dependent points: no look-ahead
                losses_dev = self.train_chunk_device(n, nxt)
                acc += losses_dev.sum()
                batch += losses_dev.numel()
                if self.log_every and (self._step // self.log_every != (self._step - losses_dev.numel()) // self.log_every):
                    print("Global step: {} Epoch: {} Batch: {} loss: {}".format(self._step, epoch, batch - 1, float(losses_dev[-1].item())))
                g = self._step
                if es and g >= to_reach and g < step0 + self.train_times * self.nbatches:
                    while g >= to_reach:
                        to_reach += stopping
                    a, l = self.valid_accuracy(), float(losses_dev[-1].item())
                    checks.append((g, a, l))
                    if a > best_acc:                                   # distribute_training.py:319-327
                        best_acc, wait_acc, best_snap_acc = a, 0, self._snapshot()
                    elif wait_acc < self.early_stop_patience:
                        wait_acc += 1
                    if wait_acc >= self.early_stop_patience:
                        print("Accuracy early stop. Accuracy has not been improved enough in {} times".format(self.early_stop_patience))
                        self.early_stop = {"reason": "accuracy", "best_step": best_snap_acc["step"]}
                        stop = True
                        break
                    if l < best_loss:                                  # :343-351
                        best_loss, wait_loss, best_snap_loss = l, 0, self._snapshot()
                    elif wait_loss < self.early_stop_patience:
                        wait_loss += 1
                    if wait_loss >= self.early_stop_patience:
                        print("Loss early stop. Losses has not been improved enough in {} times".format(self.early_stop_patience))
                        self.early_stop = {"reason": "loss", "best_step": best_snap_loss["step"]}
                        stop = True
                        break
            res = float(acc.item())
            losses.append(res)
            print("Epoch: {} loss: {} ({:.3f} s)".format(epoch, res, time.time() - t0))
            if stop:
                break
        if es and self.early_stop is not None:
            # main_spark.py:347-376 keeps the checkpoint nearest to the step written to stop.txt; checks fall on checkpoint
            # steps (both are multiples of nbatches), so that is the snapshot taken when the best value was seen
            self._restore_snapshot(best_snap_acc if self.early_stop["reason"] == "accuracy" else best_snap_loss)
            self.early_stop.update(best_acc=best_acc, best_loss=best_loss, checks=checks)
            if self.out_path is not None and self._rank() == 0:
                import os as _os
                with open(_os.path.join(_os.path.dirname(self.out_path) or ".", "stop.txt"), "w") as f:
                    f.write(str(self.early_stop["best_step"]) + "\n")
        elif es:
            self.early_stop = {"reason": None, "best_step": None, "best_acc": best_acc, "best_loss": best_loss, "checks": checks}
        if self.exportName is not None:
            self.save_tensorflow()                    # collective; rank 0 writes
        if self.out_path is not None:
            self._settle()
            if self._rank() == 0:
                self.save_parameters(self.out_path)
            self._barrier()
        return losses

    def grow_entities(self, n_new, seed=None):
        """New entities arrived with an incremental batch (main_spark.py:29-96 update_entities_and_model): the entity
        tables get `n_new` more rows drawn like a fresh Xavier-normal table of the FINAL shape, their Adam slots get
        zero rows, everything else is kept.  Call before init() re-imports the grown dataset or right after it."""
        self._ensure_model()
        if n_new <= 0:
            return self.trainModel.parameter_lists["ent_embeddings"].shape[0]
        import math
        gen = torch.Generator(device="cpu")
        gen.manual_seed(int(self.seed if seed is None else seed) + 7919)
        P = self.trainModel.parameter_lists
        for name in ("ent_embeddings", "ent_transfer"):
            if name not in P:
                continue
            old = P[name]
            rows, cols = old.shape[0] + n_new, old.shape[1]
            std = math.sqrt(1.3 * 2.0 / (rows + cols))
            fresh = torch.empty(n_new, cols, dtype=torch.float32)
            torch.nn.init.trunc_normal_(fresh, mean=0.0, std=std, a=-2 * std, b=2 * std, generator=gen)
            P[name] = torch.cat([old, fresh.to(old.device)], 0).contiguous()
            if self._adam is not None:
                for pre in ("m_", "v_"):
                    a = self._adam[pre + name]
                    self._adam[pre + name] = torch.cat([a, torch.zeros(n_new, cols, dtype=a.dtype, device=a.device)], 0).contiguous()
        self._model_struct = None
        return P["ent_embeddings"].shape[0]

    train = run

    # ------------------------------------------------------------------ scoring (Config.py:478-488)
    def test_step_device(self, test_h, test_t, test_r):
        """predict_def for arbitrary triples, scores left on the GPU: float32 device tensor [n]."""
        self._ensure_model()
        self._settle()
        dev = self.trainModel.device
        h = torch.as_tensor(np.ascontiguousarray(test_h, dtype=np.int64)).to(dev)
        t = torch.as_tensor(np.ascontiguousarray(test_t, dtype=np.int64)).to(dev)
        r = torch.as_tensor(np.ascontiguousarray(test_r, dtype=np.int64)).to(dev)
        out = torch.empty(h.numel(), dtype=torch.float32, device=dev)
        m = self._cmodel()
        self.ctx.call("okb_predict", ctypes.byref(m), _vp(h.data_ptr()), _vp(t.data_ptr()), _vp(r.data_ptr()), h.numel(),
                      _vp(out.data_ptr()), _stream())
        return out

    def test_step(self, test_h, test_t, test_r):
        """predict for arbitrary triples: float32 [n] (TransE) or [n,1] (TransH/R/D)."""
        res = self.test_step_device(test_h, test_t, test_r).cpu().numpy()
        return res.reshape(-1, 1) if self.trainModel.predict_keepdims else res

    def _fit_thresholds_device(self):
        """getValidBatch + predict + getBestThreshold (Config.py:503-507) with scores, threshold grid search and the result
        on the GPU; self.relThresh receives the fitted thresholds.  Returns (thresholds, valid pos scores, valid neg scores)
        as device tensors."""
        vp = self.test_step_device(self.valid_pos_h, self.valid_pos_t, self.valid_pos_r)
        vn = self.test_step_device(self.valid_neg_h, self.valid_neg_t, self.valid_neg_r)
        th = torch.as_tensor(self.relThresh).to(vp.device)
        self.ctx.call("okb_tc_thresholds_dev", _vp(vp.data_ptr()), _vp(vn.data_ptr()), _vp(th.data_ptr()), _stream())
        self.relThresh[:] = th.cpu().numpy()
        return th, vp, vn

    # ------------------------------------------------------------------ evaluation
    def link_prediction_records(self, q_lo=0, q_hi=None, cand_lo=0, cand_hi=None, reduce_fn=None):
        """8-int records of testHead/testTail for test triples [q_lo,q_hi): int64 [n, 2, 8] on the GPU."""
        self._ensure_model()
        dev = self.trainModel.device
        q_hi = self.testTotal if q_hi is None else q_hi
        cand_hi = self.entTotal if cand_hi is None else cand_hi
        n = q_hi - q_lo
        counts = torch.zeros(n * 8, dtype=torch.int64, device=dev)
        best = torch.full((n * 8,), -1, dtype=torch.int64, device=dev)       # all ones = "no candidate"
        m = self._cmodel()
        self.ctx.call("okb_set_flag", 1, int(bool(self.transr_tensor_cores)))
        self.ctx.call("okb_rank", ctypes.byref(m), q_lo, q_hi, int(bool(self.test_head)), cand_lo, cand_hi,
                      _vp(counts.data_ptr()), _vp(best.data_ptr()), _stream())
        if reduce_fn is not None:
            counts, best = reduce_fn(counts, best)
        out = torch.empty(n * 16, dtype=torch.int64, device=dev)
        self.ctx.call("okb_rank_finalize", q_lo, q_hi, _vp(counts.data_ptr()), _vp(best.data_ptr()), _vp(out.data_ptr()), _stream())
        return out.view(n, 2, 8)

    def test(self):
        """Triple classification (Config.py:491-516) and/or link prediction (distribute_training.py:464-607)."""
        self._ensure_model()
        if self.importName is not None:
            self.restore_tensorflow()
        t0 = time.time()
        if self.test_triple_classification:
            self.ctx.call("okb_tc_batch", 1, *[_vp(getattr(self, "valid_" + n + "_addr")) for n in ("pos_h", "pos_t", "pos_r", "neg_h", "neg_t", "neg_r")])
            th, _, _ = self._fit_thresholds_device()          # scores, threshold search and counts stay on the GPU (csrc/tc.cu)
            self.ctx.call("okb_tc_batch", 0, *[_vp(getattr(self, "test_" + n + "_addr")) for n in ("pos_h", "pos_t", "pos_r", "neg_h", "neg_t", "neg_r")])
            tp = self.test_step_device(self.test_pos_h, self.test_pos_t, self.test_pos_r)
            tn = self.test_step_device(self.test_neg_h, self.test_neg_t, self.test_neg_r)
            cnt = np.zeros(4, np.int64)
            self.ctx.call("okb_tc_counts_dev", _vp(th.data_ptr()), _vp(tp.data_ptr()), _vp(tn.data_ptr()), 0, _addr(cnt), _vp(self.acc_addr), _stream())
            TP, TN, FP, FN = [float(x) for x in cnt]
            prec, rec = TP / max(TP + FP, 1.0), TP / max(TP + FN, 1.0)
            self.tc_counts = cnt
            print("triple classification accuracy is %f" % self.acc[0])          # Test.h:381-384
            print("triple classification precision is %f" % prec)
            print("triple classification recall is %f" % rec)
            print("triple classification f-measure is %f" % ((2 * prec * rec) / max(prec + rec, 1e-30)))
        if self.test_link_prediction:
            if self._world is not None:
                rec = self._world.link_prediction(self)
            else:
                rec = self.link_prediction_records()
            rec = rec.cpu().numpy()
            self.lp_records = rec
            d = metrics.accumulate(rec, bool(self.test_head))
            self.lp_results = metrics.finalize(d, self.testTotal)
            print(metrics.format_table(self.lp_results, bool(self.test_head)))
        print("\nElapsed test time (seconds): {}".format(time.time() - t0))
        return getattr(self, "lp_results", None)

    def plot_roc(self, rel_index, fig_name=None):
        """ROC of one relation (Config.py:519-571).  matplotlib is optional: returns (FPR, TPR, auc)."""
        self._ensure_model()
        if self.importName is not None:
            self.restore_tensorflow()
        self.init_triple_classification()
        a = lambda pre: [_vp(getattr(self, pre + n + "_addr")) for n in ("pos_h", "pos_t", "pos_r", "neg_h", "neg_t", "neg_r")]
        self.ctx.call("okb_tc_batch", 1, *a("valid_"))
        pv = self.test_step(self.valid_pos_h, self.valid_pos_t, self.valid_pos_r)
        nv = self.test_step(self.valid_neg_h, self.valid_neg_t, self.valid_neg_r)
        self.ctx.call("okb_tc_batch", 0, *a("test_"))
        pt = self.test_step(self.test_pos_h, self.test_pos_t, self.test_pos_r)
        nt = self.test_step(self.test_neg_h, self.test_neg_t, self.test_neg_r)
        n_int = int(self.lib.okb_n_interval(self.ctx.h, rel_index, _addr(pv), _addr(nv)))
        self.lib.okb_tpfp.restype = ctypes.POINTER(ctypes.c_int64 * ((n_int + 1) * 2))
        ptr = self.lib.okb_tpfp(self.ctx.h, rel_index, _addr(pv), _addr(nv), _addr(pt), _addr(nt))
        if not ptr:       # Test.h:417 returns a null pointer for a relation without valid triples (the reference then crashes)
            raise OkbError("plot_roc: relation %d has no valid triples to fit a threshold grid on" % rel_index)
        res = [j for j in ptr.contents]
        TPR, FPR = [], []
        if res[0] != 0 or res[0 + n_int + 1] != 0:
            TPR.append(0)
            FPR.append(0)
        for i in range(0, n_int + 1):
            TPR.append(res[i])
            FPR.append(res[i + n_int + 1])
        if TPR[-1] != len(pt.flatten()) or FPR[-1] != len(nt.flatten()):
            TPR.append(len(pt.flatten()))
            FPR.append(len(nt.flatten()))
        TPR = [x / TPR[-1] for x in TPR]
        FPR = [x / FPR[-1] for x in FPR]
        auc = float(np.trapezoid(TPR, FPR)) if hasattr(np, "trapezoid") else float(np.trapz(TPR, FPR))
        try:
            import matplotlib.pyplot as plt
            plt.figure()
            plt.plot(FPR, TPR, color="darkorange", lw=2, label="ROC curve (area = %0.3f)" % auc)
            plt.plot([0, 1], [0, 1], color="navy", lw=2, linestyle="--")
            plt.xlabel("False Positive Rate (FPR)")
            plt.ylabel("True Positive Rate (TPR)")
            plt.legend(loc="lower right")
            plt.savefig(fig_name) if fig_name else plt.show()
        except ImportError:
            pass
        return FPR, TPR, auc

    # ------------------------------------------------------------------ predict_* (Config.py:574-663)
    def predict_head_entity(self, t, r, k):
        if self.importName is not None:
            self.restore_tensorflow()
        test_h = np.array(range(self.entTotal))
        test_r = np.array([r] * self.entTotal)
        test_t = np.array([t] * self.entTotal)
        res = self.test_step(test_h, test_t, test_r).reshape(-1).argsort()[:k]
        print(res)
        return res

    def predict_tail_entity(self, h, r, k):
        if self.importName is not None:
            self.restore_tensorflow()
        test_h = np.array([h] * self.entTotal)
        test_r = np.array([r] * self.entTotal)
        test_t = np.array(range(self.entTotal))
        res = self.test_step(test_h, test_t, test_r).reshape(-1).argsort()[:k]
        print(res)
        return res

    def predict_relation(self, h, t, k):
        if self.importName is not None:
            self.restore_tensorflow()
        test_h = np.array([h] * self.relTotal)
        test_r = np.array(range(self.relTotal))
        test_t = np.array([t] * self.relTotal)
        res = self.test_step(test_h, test_t, test_r).reshape(-1).argsort()[:k]
        print(res)
        return res

    def predict_triple(self, h, t, r, thresh=None):
        self.init_triple_classification()
        if self.importName is not None:
            self.restore_tensorflow()
        res = self.test_step(np.array([h]), np.array([t]), np.array([r]))
        if thresh is None:
            a = [_vp(getattr(self, "valid_" + n + "_addr")) for n in ("pos_h", "pos_t", "pos_r", "neg_h", "neg_t", "neg_r")]
            self.ctx.call("okb_tc_batch", 1, *a)
            self._fit_thresholds_device()
            thresh = self.relThresh[r]
        ok = bool(res.reshape(-1)[0] < thresh)
        print("triple (%d,%d,%d) is %s" % (h, t, r, "correct" if ok else "wrong"))
        return ok
