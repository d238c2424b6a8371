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
    assert built_lib.pdes_abi_version() == 3
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


def test_every_densenet_layer_has_a_tensor_core_plan_within_the_sm_limits(built_lib):
    """Host-only check of the tiling logic: for every convolution of DenseED (blocks [6,8,6] and a few other
    layouts) but the first, forward (K = Cin, N = Cout) and dgrad (K = Cout, N = Cin) get a plan that fits an
    sm_100a SM: <= 227 KB of dynamic shared memory, <= 512 TMEM columns, rings at least 2 deep."""
    from ctypes import c_int64
    from oracle import pdes_oracle as orc

    def rup(v, m):
        return (v + m - 1) // m * m

    n_checked = 0
    for blocks, growth, init in (([6, 8, 6], 16, 48), ([3, 4, 3], 16, 48), ([2, 2, 2], 8, 16), ([4, 4, 4, 4, 4], 24, 64)):
        plan = orc.densenet_plan(1, 3, 64, blocks, growth_rate=growth, init_features=init)
        convs = [(L["cin"], L["cout"], L["k"]) for L in plan]  # every stage holds exactly one convolution
        for cin, cout, k in convs[1:]:
            for ck, n in ((cin, rup(cout, 16)), (cout, rup(cin, 16))):
                if n > 256:
                    continue  # served by the CUDA-core kernels
                out = (c_int64 * 11)()
                assert built_lib.pdes_conv_tc_plan(k, ck, n, out) == 0
                sup, KC, nchunks, ngroups, S, TS, AST, NB, TPB, smem, tmem = [int(v) for v in out]
                assert sup == 1, (blocks, cin, cout, k, ck, n, list(out))
                assert KC in (16, 32) and nchunks * KC >= ck and ngroups == 2
                assert smem <= 227 * 1024 and tmem <= 512 and AST >= 2 and NB >= 2
                assert (k * k) % TPB == 0 and 1 <= S <= 4 and TS in (1, 2)
                n_checked += 1
    assert n_checked > 100
