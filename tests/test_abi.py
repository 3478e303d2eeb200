"""CPU-side checks of the C-ABI boundary: the shared library builds/loads, exports every symbol declared in
include/cmgan_b200.h, the ctypes mirror of the argument block matches, and the product path refuses to run on CPU."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT


def test_library_exports_every_declared_symbol():
    from cmgan_b200 import _lib
    from cmgan_b200.build import build
    build()
    L = _lib.lib()
    decl = set(L.protos)
    with open(os.path.join(ROOT, "include", "cmgan_b200.h")) as fh:
        names = set(re.findall(r"\b(cmgan_[a-z0-9_]+)\(", fh.read()))
    assert decl == names and len(decl) > 40
    for n in decl:
        assert hasattr(L.cdll, n), n
    assert L.cdll.cmgan_abi_version() == 1
    assert L.cdll.cmgan_gemm_args_size() == ctypes.sizeof(_lib.GemmArgs)


def test_error_channel_reports_bad_arguments():
    from cmgan_b200 import _lib
    L = _lib.lib()
    rc = L.cdll.cmgan_ln_stats(None, 64, 10, None, None)
    assert rc == -1 and b"cmgan_ln_stats" in L.cdll.cmgan_last_error()
    a = _lib.GemmArgs()
    rc = L.cdll.cmgan_gemm_rows_f32(ctypes.byref(a), None)
    assert rc == -1 and b"gemm_rows" in L.cdll.cmgan_last_error()


def test_state_dict_contract(g_weights):
    import cmgan_b200
    m = cmgan_b200.TSCNet(num_channel=64, num_features=201)
    assert list(m.state_dict().keys()) == list(g_weights.keys())
    for k, v in m.state_dict().items():
        assert tuple(v.shape) == tuple(g_weights[k].shape), k
    m.load_state_dict(g_weights, strict=True)
    assert sum(p.numel() for p in m.parameters()) == 1834833


def test_no_cpu_fallback():
    import cmgan_b200
    m = cmgan_b200.TSCNet()
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 2, 5, 201))
    with pytest.raises(RuntimeError):
        cmgan_b200.power_compress(torch.zeros(1, 201, 5, 2))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "cmgan_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            src = open(os.path.join(pkg, f)).read()
            assert "oracle" not in src.replace("no CPU or", ""), f
