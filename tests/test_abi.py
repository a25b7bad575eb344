"""CPU-side checks of the drop-in boundary: the library loads and exports every symbol that
include/kge_b200.h declares; the ctypes structs match the header; no compute is called."""
import os
import re
import ctypes as C

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "kge_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(kge_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from dglke_b200 import _lib
    lib = _lib.load_library()
    names = _header_functions()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), "libkge_b200.so does not export %s" % n
    assert sorted(names) == sorted(_lib.EXPORTS)
    assert lib.kge_abi_version() == 4


def test_struct_layouts_match_header():
    from dglke_b200 import _lib
    assert C.sizeof(_lib.Shard) == 40
    assert C.sizeof(_lib.Table) == 32
    assert C.sizeof(_lib.StepCfg) == 80
    assert C.sizeof(_lib.Batch) == 80
    assert _lib.StepCfg.batch.offset == 48 and _lib.StepCfg.neg_sample_size.offset == 60
    assert _lib.StepCfg.loss_genre.offset == 64 and _lib.StepCfg.margin.offset == 68 and _lib.StepCfg.pairwise.offset == 72


def test_fails_loudly_without_gpu():
    import torch
    from dglke_b200 import _lib
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.KgeError):
        _lib.Handle(0)
    h = C.c_void_p()
    rc = _lib.load_library().kge_create(0, C.byref(h))
    assert rc == -5 and b"no CPU path" in _lib.load_library().kge_last_error()
