"""The C-ABI library loads and exports every symbol include/pdes_b200.h declares (no GPU)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built_lib():
    from pde_surrogate_b200 import build, _lib
    build.build()
    return _lib.lib()


def test_header_symbols_exported(built_lib):
    from pde_surrogate_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "pdes_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(pdes_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 25
    for name in sorted(declared):
        assert hasattr(built_lib, name), "library does not export %s" % name
    assert declared == set(_lib.SIGNATURES), "ctypes table and header disagree: %s" % (
        declared ^ set(_lib.SIGNATURES))


def test_host_only_calls(built_lib):
    from ctypes import byref, c_void_p
    from pde_surrogate_b200 import _lib
    assert built_lib.pdes_abi_version() == 1
    cfg = _lib.DensenetConfig()
    cfg.in_channels, cfg.out_channels, cfg.imsize, cfg.n_blocks = 1, 3, 64, 3
    for i, b in enumerate([6, 8, 6]):
        cfg.blocks[i] = b
    cfg.growth_rate, cfg.init_features, cfg.max_batch = 16, 48, 32
    h = c_void_p()
    assert built_lib.pdes_densenet_create(byref(cfg), byref(h)) == 0
    assert built_lib.pdes_densenet_num_params(h) == 82
    assert built_lib.pdes_densenet_num_bn(h) == 27
    # 1.4616 GFLOP/sample forward, 4.380 GFLOP/sample training
    assert abs(built_lib.pdes_densenet_flops(h, 1, 0) / 1e9 - 1.4616) < 1e-3
    assert abs(built_lib.pdes_densenet_flops(h, 1, 1) / 1e9 - 4.380) < 2e-3
    built_lib.pdes_densenet_destroy(h)
    cfg.n_blocks = 2  # even number of blocks is rejected like the reference (codec.py:231-233)
    assert built_lib.pdes_densenet_create(byref(cfg), byref(h)) != 0
    assert b"odd" in built_lib.pdes_last_error()
