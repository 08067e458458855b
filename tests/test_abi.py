"""The C-ABI library loads on a CPU-only box and exports every symbol include/okb200.h declares.
No compute entry point is called here (there is no GPU and no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "okb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", src)
    skip = {"defined", "cudaSetDevice"}
    return sorted({n for n in names if n not in skip})


def test_header_symbols_are_exported(built):
    from openkeonspark_b200 import _native
    lib = _native.load()
    syms = header_symbols()
    assert len(syms) > 60
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing
    # and the binding tables cover the header
    bound = set(_native.LEGACY_SYMBOLS) | set(_native.NATIVE_SYMBOLS)
    assert set(syms) <= bound, sorted(set(syms) - bound)


def test_reference_symbol_set_is_covered(built):
    """Every symbol the reference's Config.py / distribute_training.py call through ctypes exists."""
    from openkeonspark_b200 import _native
    lib = _native.load()
    used = ["sampling", "getTailBatch", "testTail", "getHeadBatch", "testHead", "getTestBatch", "getValidBatch",
            "getBestThreshold", "test_triple_classification", "get_n_interval", "get_TPFP", "importTestFiles",
            "importTypeFiles", "importOntologyFiles", "getTestTotal", "getValidTotal", "getRelationTotal", "setInPath",
            "setBern", "setWorkThreads", "randReset", "importTrainFiles", "getEntityTotal", "getTrainTotal_", "getBatchTotal"]
    for s in used:
        assert hasattr(lib, s), s


def test_host_only_calls_work_without_gpu(built, tmp_path):
    from openkeonspark_b200 import _native
    c = _native.Ctx()
    assert c.lib.okb_version() >= 100
    c.call("okb_set_work_threads", 8)
    c.call("okb_set_bern", 1)
    with pytest.raises(_native.OkbError):
        c.call("okb_set_work_threads", 0)
    c.call("okb_set_in_path", str(tmp_path).encode())
    with pytest.raises(_native.OkbError):          # missing files -> OKB_ERR_IO, not a crash
        c.call("okb_import_train_files")
    # gradient-buffer geometry (pure host arithmetic): TransR's relation rows are per RELATION without relation negatives
    # and per (positive, relation slot) with them (TransR.py:57-65); the other models always use the latter
    m = _native.okb_model()
    m.ent_dim, m.rel_dim = 20, 12
    er, ec, rr, rc = (ctypes.c_int64() for _ in range(4))
    for model, kr, want_rows, want_cols in ((2, 0, 0, 12 + 20 * 12), (2, 2, 100 * 3, 12 + 20 * 12), (1, 2, 100 * 3, 2 * 12), (0, 0, 100, 12)):
        m.model = model
        c.call("okb_grad_sizes", ctypes.byref(m), 100, 3, kr, ctypes.byref(er), ctypes.byref(ec), ctypes.byref(rr), ctypes.byref(rc))
        assert (er.value, ec.value) == (100 * 5, 20), model
        assert (rr.value, rc.value) == (want_rows, want_cols), (model, kr)      # TransR, kr = 0: R rows = 0 before any import
    c.close()


def test_missing_library_fails_loudly(tmp_path):
    from openkeonspark_b200 import _native
    with pytest.raises(_native.OkbError):
        _native.load(str(tmp_path / "nope.so"))


def test_product_never_imports_oracle():
    """A product path that routes through the oracle voids every parity claim."""
    pkg = os.path.join(ROOT, "openkeonspark_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh")):
                txt = open(os.path.join(dirpath, f)).read()
                for needle in ("from oracle", "import oracle", "libkge_oracle", "oracle.harness", "_ref/Base.so", "orc_"):
                    assert needle not in txt, (f, needle)


def test_ctypes_signatures_match_the_header(built):
    """Every okb_* prototype of include/okb200.h has a ctypes signature with the same number of parameters, pointers bound
    as pointers and 64-bit INTs as c_int64 (a drift here corrupts arguments silently and only shows on the GPU box)."""
    from openkeonspark_b200 import _native
    src = open(os.path.join(ROOT, "include", "okb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = re.findall(r"\b(okb_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S)
    assert len(protos) > 50
    checked = 0
    for name, args in protos:
        if name not in _native._SIGS:
            continue
        params = [a.strip() for a in args.replace("\n", " ").split(",") if a.strip() and a.strip() != "void"]
        res, sig = _native._SIGS[name]
        assert len(sig) == len(params), (name, len(sig), params)
        for p, ct in zip(params, sig):
            is_ptr = "*" in p
            ct_ptr = ct in (ctypes.c_void_p, ctypes.c_char_p) or hasattr(ct, "contents")
            assert is_ptr == ct_ptr, (name, p, ct)
            if not is_ptr and re.match(r"(const\s+)?INT\b", p):
                assert ct is ctypes.c_int64, (name, p, ct)
        checked += 1
    assert checked > 50


def test_wait_word_host_side(built):
    """okb_wait_word is host code: a word that already differs from the sentinel returns at once without touching CUDA; a
    word that never changes ends in an error once the stream query fails (no device here) instead of spinning for ever."""
    import numpy as np
    from openkeonspark_b200 import _native
    c = _native.Ctx()
    w = np.array([0x3f800000], dtype=np.uint32)
    c.call("okb_wait_word", ctypes.c_void_p(w.ctypes.data), 0xFFC0DEAD, None)
    w[0] = 0xFFC0DEAD
    with pytest.raises(_native.OkbError):
        c.call("okb_wait_word", ctypes.c_void_p(w.ctypes.data), 0xFFC0DEAD, None)
    c.close()


def test_ctypes_struct_layouts_match_the_header(tmp_path):
    """okb_model / okb_hyper / okb_dp cross the C ABI by pointer: every field of the ctypes mirrors must sit at the offset
    the C compiler gives it in include/okb200.h (a field added on one side only would silently shift everything after it)."""
    import subprocess
    from openkeonspark_b200 import _native
    structs = {"okb_model": _native.okb_model, "okb_hyper": _native.okb_hyper, "okb_dp": _native.okb_dp}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "okb200.h"', 'int main(void) {']
    for name, cls in structs.items():
        lines.append('  printf("%s.sizeof %%zu\\n", sizeof(%s));' % (name, name))
        for field, _ in cls._fields_:
            lines.append('  printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (name, field, name, field))
    lines += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    inc = os.path.join(ROOT, "include")
    r = subprocess.run(["gcc", "-std=c99", "-I", inc, str(src), "-o", str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]            # a field the header does not have fails here
    out = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True).stdout.splitlines())
    for name, cls in structs.items():
        assert int(out[name + ".sizeof"]) == ctypes.sizeof(cls), name
        for field, _ in cls._fields_:
            assert int(out["%s.%s" % (name, field)]) == getattr(cls, field).offset, (name, field)
    # ... and the header has no field the mirror lacks (sizes equal + every mirrored field at its place covers it)
