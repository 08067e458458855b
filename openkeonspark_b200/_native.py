"""ctypes binding of libokb200.so (include/okb200.h) — the only way the Python host reaches the GPU.

There is deliberately no fallback: if the library is missing or a call fails, an exception is
raised.  Nothing in this package imports the CPU oracle.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("OKB200_LIB") or os.path.join(HERE, "libokb200.so")    # OKB200_LIB: A/B builds (tools/)

_vp, _i64, _int = C.c_void_p, C.c_int64, C.c_int


class OkbError(RuntimeError):
    pass


class okb_model(C.Structure):
    _fields_ = [("model", C.c_int32), ("ent_dim", C.c_int32), ("rel_dim", C.c_int32), ("optimizer", C.c_int32),
                ("ent", _vp), ("rel", _vp), ("ent_aux", _vp), ("rel_aux", _vp),
                ("m_ent", _vp), ("v_ent", _vp), ("m_rel", _vp), ("v_rel", _vp),
                ("m_ent_aux", _vp), ("v_ent_aux", _vp), ("m_rel_aux", _vp), ("v_rel_aux", _vp)]


class okb_hyper(C.Structure):
    _fields_ = [("margin", C.c_float), ("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float)]


OKB_DP_MAX = 16


class okb_dp(C.Structure):
    _fields_ = [("rank", C.c_int32), ("world", C.c_int32), ("b_lo", _i64), ("b_hi", _i64), ("arena", _vp * OKB_DP_MAX),
                ("off_ent", _i64), ("off_ent_aux", _i64), ("off_rel", _i64), ("off_rel_aux", _i64),
                ("off_stage_ent", _i64), ("off_stage_rel", _i64), ("off_flags", _i64), ("arena_bytes", _i64),
                ("plan_steps", _i64), ("max_local", _i64), ("neg_ent", _i64), ("neg_rel", _i64),
                ("off_rowhead", _i64), ("off_perm", _i64), ("off_gent", _i64), ("off_grel", _i64), ("off_loss", _i64),
                ("off_partial", _i64), ("scatter", _i64), ("global_batch", _i64)]


MODEL_ID = {"TransE": 0, "TransH": 1, "TransR": 2, "TransD": 3}

# name -> (restype, argtypes); every okb_* symbol of include/okb200.h
_SIGS = {
    "okb_version": (_int, []),
    "okb_default_ctx": (_vp, []),
    "okb_create": (_int, [C.POINTER(_vp)]),
    "okb_destroy": (_int, [_vp]),
    "okb_last_error": (C.c_char_p, [_vp]),
    "okb_set_device": (_int, [_vp, _int]),
    "okb_set_in_path": (_int, [_vp, C.c_char_p]),
    "okb_set_bern": (_int, [_vp, _i64]),
    "okb_set_work_threads": (_int, [_vp, _i64]),
    "okb_import_train_files": (_int, [_vp]),
    "okb_import_test_files": (_int, [_vp]),
    "okb_import_type_files": (_int, [_vp]),
    "okb_import_ontology_files": (_int, [_vp]),
    "okb_import_train_arrays": (_int, [_vp, _i64, _i64, _vp, _vp, _vp, _i64, _i64]),
    "okb_import_test_arrays": (_int, [_vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _i64]),
    "okb_build_type_constraints": (_int, [_vp]),
    "okb_total": (_i64, [_vp, _int]),
    "okb_rand_reset": (_int, [_vp]),
    "okb_set_streams": (_int, [_vp, _vp, _i64]),
    "okb_get_streams": (_int, [_vp, _vp, _i64]),
    "okb_sample": (_int, [_vp, _i64, _i64, _i64, _i64, _i64, _i64, _vp]),
    "okb_batch_ptrs": (_int, [_vp, _i64, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp)]),
    "okb_batch_to_host": (_int, [_vp, _i64, _vp, _vp, _vp, _vp, _vp]),
    "okb_sample_to_host": (_int, [_vp, _i64, _i64, _i64, _i64, _i64, _vp, _vp, _vp, _vp]),
    "okb_batch_from_host": (_int, [_vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp]),
    "okb_batch_check": (_int, [_vp, _vp]),
    "okb_grad_sizes": (_int, [_vp, C.POINTER(okb_model), _i64, _i64, _i64] + [C.POINTER(_i64)] * 4),
    "okb_plan": (_int, [_vp, _i64, _vp]),
    "okb_plan_steps": (_int, [_vp, _i64, _i64, _vp]),
    "okb_grad": (_int, [_vp, C.POINTER(okb_model), C.POINTER(okb_hyper), _i64, _i64, _i64, _vp, _vp, _vp, _vp]),
    "okb_update": (_int, [_vp, C.POINTER(okb_model), C.POINTER(okb_hyper), _i64, _vp, _vp, _vp, _vp, _vp]),
    "okb_train_step": (_int, [_vp, C.POINTER(okb_model), C.POINTER(okb_hyper), _i64, _vp, _vp]),
    "okb_train_steps": (_int, [_vp, C.POINTER(okb_model), C.POINTER(okb_hyper), _i64, _i64, _vp, _vp]),
    "okb_wait_word": (_int, [_vp, _vp, C.c_uint, _vp]),
    "okb_train_step_host": (_int, [_vp, C.POINTER(okb_model), C.POINTER(okb_hyper), _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp]),
    "okb_peer_alloc": (_int, [_vp, _i64, C.POINTER(_vp), _vp]),
    "okb_peer_open": (_int, [_vp, _vp, C.POINTER(_vp)]),
    "okb_peer_close": (_int, [_vp, _vp]),
    "okb_peer_free": (_int, [_vp, _vp]),
    "okb_dp_layout": (_int, [_vp, C.POINTER(okb_model), _i64, C.POINTER(okb_dp)]),
    "okb_dp_attach": (_int, [_vp, C.POINTER(okb_dp)]),
    "okb_dp_detach": (_int, [_vp]),
    "okb_dp_train_steps": (_int, [_vp, C.POINTER(okb_model), C.POINTER(okb_hyper), _i64, _i64, _vp, _vp]),
    "okb_dp_quiesce": (_int, [_vp, _vp]),
    "okb_transr_set_shard": (_int, [_vp, _i64, _i64]),
    "okb_chunk_begin": (_int, [_vp, _i64, _i64, _i64, _i64, _vp]),
    "okb_chunk_prefetch": (_int, [_vp, _i64, _i64, _i64, _i64, _vp]),
    "okb_dp_chunk_begin": (_int, [_vp, _i64, _i64, _i64, _i64, _i64, _i64, _vp]),
    "okb_dp_chunk_prefetch": (_int, [_vp, _i64, _i64, _i64, _i64, _i64, _i64, _vp]),
    "okb_predict": (_int, [_vp, C.POINTER(okb_model), _vp, _vp, _vp, _i64, _vp, _vp]),
    "okb_rank": (_int, [_vp, C.POINTER(okb_model), _i64, _i64, _int, _i64, _i64, _vp, _vp, _vp]),
    "okb_rank_finalize": (_int, [_vp, _i64, _i64, _vp, _vp, _vp, _vp]),
    "okb_rank_scores": (_int, [_vp, _i64, _int, _vp, _vp, _vp]),
    "okb_tc_batch": (_int, [_vp, _int] + [_vp] * 6),
    "okb_best_threshold": (_int, [_vp, _vp, _vp, _vp]),
    "okb_tc_eval": (_int, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "okb_tc_eval_valid": (_int, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "okb_tc_thresholds_dev": (_int, [_vp, _vp, _vp, _vp, _vp]),
    "okb_tc_counts_dev": (_int, [_vp, _vp, _vp, _vp, _int, _vp, _vp, _vp]),
    "okb_test_list": (_int, [_vp, _int, _vp, _vp, _vp]),
    "okb_n_interval": (_i64, [_vp, _i64, _vp, _vp]),
    "okb_tpfp": (_vp, [_vp, _i64, _vp, _vp, _vp, _vp]),
    "okb_set_flag": (_int, [_vp, _int, _i64]),
    "okb_prof_enable": (_int, [_vp, _int]),
    "okb_prof_read": (_int, [_vp, _int, C.POINTER(C.c_double), C.POINTER(_i64)]),
    "okb_debug_cuda_error": (C.c_char_p, []),
    "okb_debug_dp_trace": (_int, [_vp, _vp]),
    "okb_launch_count": (_i64, []),
}

# the reference's Base.so symbols (Config.py:30-51 binds these by name)
LEGACY_SYMBOLS = [
    "setInPath", "setOutPath", "setWorkThreads", "getWorkThreads", "setBern", "getEntityTotal", "getRelationTotal",
    "getTripleTotal", "getTrainTotal", "getTrainTotal_", "getBatchTotal", "getTestTotal", "getValidTotal", "randReset",
    "importTrainFiles", "importTestFiles", "importTypeFiles", "importOntologyFiles", "sampling", "getHeadBatch",
    "getTailBatch", "testHead", "testTail", "getNegTest", "getNegValid", "getTestBatch", "getValidBatch",
    "getBestThreshold", "test_triple_classification", "get_n_interval", "get_TPFP",
]
NATIVE_SYMBOLS = list(_SIGS)

_lib = None


def load(path=None):
    """Load libokb200.so (building is NOT attempted here; see openkeonspark_b200.build)."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = os.path.abspath(path or LIB_PATH)
    if not os.path.exists(p):
        raise OkbError("%s not found: build it with `python -m openkeonspark_b200.build` "
                       "(there is no CPU fallback)" % p)
    lib = C.CDLL(p)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    if path is None:
        _lib = lib
    return lib


class Ctx:
    """Thin object wrapper over an okb_ctx*; raises OkbError on any non-zero status."""

    def __init__(self, lib=None, default=False):
        self.lib = lib or load()
        if default:
            self.h = _vp(self.lib.okb_default_ctx())
            self._own = False
        else:
            h = _vp()
            if self.lib.okb_create(C.byref(h)) != 0:
                raise OkbError("okb_create failed")
            self.h = h
            self._own = True

    def call(self, name, *args):
        rc = getattr(self.lib, name)(self.h, *args)
        if rc != 0:
            msg = self.lib.okb_last_error(self.h)
            raise OkbError("%s -> %d: %s" % (name, rc, msg.decode() if msg else "?"))

    def total(self, what):
        return int(self.lib.okb_total(self.h, what))

    def close(self):
        if self._own and self.h:
            self.lib.okb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
