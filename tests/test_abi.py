"""CPU-side checks of the drop-in boundary: libdmt.so loads, exports every symbol include/dmt.h declares, the Python
binding covers all of them, and — without a GPU — compute entry points fail loudly instead of falling back."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import dmt_b200
from dmt_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "dmt.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dmt_[a-z_A-Z0-9]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound():
    syms = declared_symbols()
    assert len(syms) >= 40
    lib = C.CDLL(_lib.LIB_PATH)
    for s in syms:
        assert hasattr(lib, s), "libdmt.so does not export %s" % s
    assert sorted(_lib.SIGNATURES) == syms, set(syms) ^ set(_lib.SIGNATURES)


def test_every_entry_point_cites_the_reference():
    src = open(os.path.join(ROOT, "include", "dmt.h")).read()
    assert src.count("src/") >= 25 and "src/biblock.jl:121-127" in src and "src/block.jl:104-110" in src


def test_model_dims_without_gpu():
    assert _lib.model_dims(_lib.FHN) == (2, 1, 5, True)
    assert _lib.model_dims(_lib.LORENZ) == (3, 3, 4, True)
    assert _lib.model_dims(_lib.PROK) == (4, 4, 9, False)
    assert _lib.model_dims(_lib.JR) == (6, 1, 10, True)
    with pytest.raises(ValueError):
        _lib.model_dims(17)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    n = np.array([11, 11], dtype=np.int32)
    tt = np.concatenate([np.linspace(0, 0.1, 11), np.linspace(0.1, 0.2, 11)])
    with pytest.raises(dmt_b200.DmtError) as e:
        dmt_b200.Ctx(_lib.LORENZ, n, tt, 8, obs_dim=2)
    assert e.value.code == 2 and "no CPU fallback" in str(e.value)


def test_product_never_touches_the_oracle():
    """only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference legs may use oracle/"""
    pk = dmt_b200.PKG_DIR
    for dirpath, _, files in os.walk(pk):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".jl")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "dmt_oracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, os.path.join(dirpath, f)


def test_synthetic_configs_shapes():
    from dmt_b200 import configs
    p = configs.named_config("c1")
    assert p.K == 10 and p.steps_per_chain == 1000 and p.M == 1 and p.v.shape == (10, 1, 1)
    p = configs.make_problem("lorenz", 8, K=40, layouts=configs.blocking_layouts(40, 20, 0.9))
    assert p.layouts[0][0] == [(0, 19), (20, 39)] and p.layouts[1][0] == [(0, 9), (10, 29), (30, 39)]
    a = configs.make_problem("lv", 6, K=3, seed=3)
    b = configs.make_problem("lv", 4, K=3, seed=3, chain_offset=2)
    assert np.array_equal(a.v[:, :, 2:], b.v) and np.array_equal(a.x0[:, 2:], b.x0)   # sharding sees the same data
