"""Host-side logic of the reference-facing modules on CPU (oracle-backed executor, see
tests/_cpu_backend.py): module tree / state_dict parity, flat storage, autograd wiring,
loss memoisation, error behaviour, and the UNMODIFIED reference training script end to end."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import pdes_oracle as orc
from tests._cpu_backend import cpu_backend

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("PDES_REFERENCE", "/root/reference")


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def test_state_dict_matches_reference_layout(golden_dir):
    from models.codec import DenseED
    g = np.load(os.path.join(golden_dir, "densenet_full64.npz"))
    torch.manual_seed(1)
    m = DenseED(1, 3, 64, [6, 8, 6])
    assert [n for n, _ in m.named_parameters()] == [str(s) for s in g["param_names"]]
    assert tuple(m.model_size) == (740091, 28)
    plan = orc.densenet_plan(1, 3, 64, [6, 8, 6])
    assert [(k, tuple(v.shape)) for k, v in m.state_dict().items()] == [(k, tuple(s)) for k, s in orc.state_layout(plan)]
    flat, gflat = m.flat_parameters()
    assert all(p.data_ptr() >= flat.data_ptr() and p.data_ptr() < flat.data_ptr() + flat.numel() * 4
               for p in m.parameters())
    m2 = m.double()  # _apply re-flattens and keeps values
    assert m2._flat.dtype == torch.float64 and len(m2.state_dict()) == 163
    with pytest.raises(ValueError):
        DenseED(1, 3, 64, [6, 8])
    # bottleneck dense layers (codec.py:56-64): conv1 (1x1) / conv2 (3x3) above bn_size * growth input channels
    mk = DenseED(1, 3, 64, [6, 8, 6], bottleneck=True, bn_size=4)
    plan_k = orc.densenet_plan(1, 3, 64, [6, 8, 6], bottleneck=4)
    assert [(k, tuple(v.shape)) for k, v in mk.state_dict().items()] == [(k, tuple(s)) for k, s in orc.state_layout(plan_k)]
    assert tuple(mk.state_dict()["features.EncBlock1.denselayer2.conv1.weight"].shape) == (16, 64, 3, 3)   # 64 <= 4 * 16
    assert tuple(mk.state_dict()["features.EncBlock1.denselayer3.conv1.weight"].shape) == (64, 80, 1, 1)
    assert tuple(mk.state_dict()["features.EncBlock1.denselayer3.conv2.weight"].shape) == (16, 64, 3, 3)
    with pytest.raises(NotImplementedError):
        DenseED(1, 3, 64, [6, 8, 6], upsample='bicubic')
    # upsample=None: ConvTranspose2d transitions named convT2, same state_dict layout as the reference
    mt = DenseED(1, 3, 64, [6, 8, 6], upsample=None, out_activation='softplus')
    plan_t = orc.densenet_plan(1, 3, 64, [6, 8, 6], upsample=None)
    assert [(k, tuple(v.shape)) for k, v in mt.state_dict().items()] == [(k, tuple(s)) for k, s in orc.state_layout(plan_t)]
    assert "features.TransUp1.convT2.weight" in mt.state_dict() and isinstance(mt.features.softplus, torch.nn.Softplus)
    mb = DenseED(1, 3, 64, [6, 8, 6], upsample='bilinear', drop_rate=0.1)   # script-reachable options
    assert len(mb.state_dict()) == 163


def test_no_cpu_fallback_in_product():
    from models.codec import DenseED
    from models.darcy import conv_boundary_condition
    from utils.image_gradient import SobelFilter
    m = DenseED(1, 3, 16, [1, 1, 1], growth_rate=4, init_features=8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(2, 1, 16, 16))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        conv_boundary_condition(torch.zeros(2, 3, 16, 16))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        SobelFilter(16).grad_h(torch.zeros(1, 1, 16, 16))


def test_autograd_wiring_and_memo(golden_dir):
    g = np.load(os.path.join(golden_dir, "densenet_small16.npz"))
    cfg = dict(in_channels=1, out_channels=3, imsize=16, blocks=[1, 2, 1], growth_rate=4, init_features=8)
    with cpu_backend():
        from models.codec import DenseED
        from models import darcy
        from pde_surrogate_b200 import darcy as pd
        from utils.image_gradient import SobelFilter
        model = DenseED(1, 3, 16, [1, 2, 1], growth_rate=4, init_features=8)
        model.load_state_dict(orc.make_state(orc.densenet_plan(**cfg), int(g["seed"])))
        K = orc.make_input(int(g["B"]), 16, int(g["seed"]))
        sob = SobelFilter(16, correct=True, device="cpu")
        model.train()
        model.zero_grad()
        out = model(K)
        calls = []
        real = pd._DarcyLossFn.apply
        pd._DarcyLossFn.apply = staticmethod(lambda *a: (calls.append(1), real(*a))[1])
        l_c = darcy.conv_constitutive_constraint(K, out, sob)
        l_d = darcy.conv_continuity_constraint(out, sob)
        l_dir, l_neu = darcy.conv_boundary_condition(out)
        pd._DarcyLossFn.apply = staticmethod(real)
        assert len(calls) == 1, "the three loss calls must share one fused evaluation"
        loss = (l_c + l_d) + (l_dir + l_neu) * 10.0
        loss.backward()
        assert abs(float(loss) - float(g["loss"])) <= 2e-5 * float(g["loss"])
        flat = np.concatenate([p.grad.numpy().ravel() for p in model.parameters()])
        assert rel(flat, g["grads32"]) < 1e-4
        assert all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(model.parameters(), model._grad_views))
        g1 = model.flat_parameters()[1].clone()
        # a new tensor object (even at a recycled address) must not hit the memo
        out2 = model(K)
        assert darcy.conv_boundary_condition(out2)[0] is not l_dir
        model.zero_grad()
        assert all(p.grad is None for p in model.parameters())
        out3 = model(K)
        d, n = darcy.conv_boundary_condition(out3)
        (d + n).backward()
        assert model._params[0].grad is not None and float(model.flat_parameters()[1].abs().sum()) > 0
        assert rel(model.flat_parameters()[1].numpy(), g1.numpy()) > 1e-3  # different loss, fresh buffer
        model.eval()
        with pytest.raises(NotImplementedError):
            model(K).sum().backward()
        with torch.no_grad():
            assert model(K).requires_grad is False


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "train_codec_mixed_residual.py")),
                    reason="reference checkout not present (only in the build container)")
def test_unmodified_training_script_runs(tmp_path):
    """BASELINE config 0 plumbing: the reference's own train_codec_mixed_residual.py, byte for byte,
    against this repo's models/ and utils/ packages (oracle-backed executor on CPU)."""
    from pde_surrogate_b200 import data
    d = tmp_path / "datasets" / "32x32"
    d.mkdir(parents=True)
    rs = np.random.RandomState(0)
    x = data.grf_kle(24, 32, 64, 0.2, seed=1, device="cpu").numpy()
    data.write_hdf5(str(d / "kle512_lhs10000_train.hdf5"), x[:16])
    data.write_hdf5(str(d / "kle512_lhs1000_val.hdf5"), x[16:], rs.standard_normal((8, 3, 32, 32)))
    import run_reference_script
    argv = ["--", "--data-dir", str(tmp_path / "datasets"), "--exp-dir", str(tmp_path / "exp"), "--imsize", "32",
            "--ntrain", "16", "--ntest", "8", "--batch-size", "8", "--test-batch-size", "8", "--epochs", "1",
            "--cuda", "0", "--plot-freq", "1", "--ckpt-freq", "1"]
    old_argv, old_path = list(sys.argv), list(sys.path)
    try:
        with cpu_backend():
            run_reference_script.main(argv)
    finally:
        sys.argv, sys.path[:] = old_argv, old_path
    run_dirs = list((tmp_path / "exp").rglob("args.txt"))
    assert len(run_dirs) == 1
    run = run_dirs[0].parent
    assert (run / "checkpoints" / "model_epoch1.pth").exists()
    assert (run / "training" / "loss_train.txt").exists() and (run / "training" / "r2_test.txt").exists()
    sd = torch.load(str(run / "checkpoints" / "model_epoch1.pth"))
    assert len(sd) == 163 and "features.LastTransUp.conv3.weight" in sd


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "train_codec_max_likelihood.py")),
                    reason="reference checkout not present (only in the build container)")
def test_unmodified_max_likelihood_script_runs(tmp_path):
    """the data-driven script train_codec_max_likelihood.py (same DenseED,
    F.mse_loss on labelled data, eval loop) runs byte for byte on this repo's modules — the network's
    autograd Function takes an arbitrary dL/d(output), not only the Darcy loss gradient."""
    from pde_surrogate_b200 import data
    d = tmp_path / "datasets" / "32x32"
    d.mkdir(parents=True)
    rs = np.random.RandomState(1)
    x = data.grf_kle(24, 32, 64, 0.2, seed=2, device="cpu").numpy()
    data.write_hdf5(str(d / "kle512_lhs10000_train.hdf5"), x[:16], rs.standard_normal((16, 3, 32, 32)))
    data.write_hdf5(str(d / "kle512_lhs1000_val.hdf5"), x[16:], rs.standard_normal((8, 3, 32, 32)))
    import run_reference_script
    argv = ["--script", os.path.join(REF, "train_codec_max_likelihood.py"), "--", "--data-dir",
            str(tmp_path / "datasets"), "--exp-dir", str(tmp_path / "exp"), "--imsize", "32", "--ntrain", "16",
            "--ntest", "8", "--batch-size", "8", "--test-batch-size", "8", "--epochs", "2", "--cuda", "0",
            "--plot-freq", "1", "--ckpt-freq", "1"]
    old_argv, old_path = list(sys.argv), list(sys.path)
    try:
        with cpu_backend():
            run_reference_script.main(argv)
    finally:
        sys.argv, sys.path[:] = old_argv, old_path
    run_dirs = list((tmp_path / "exp").rglob("args.txt"))
    assert len(run_dirs) == 1
    run = run_dirs[0].parent
    assert (run / "checkpoints" / "model_epoch2.pth").exists()
    losses = np.loadtxt(str(run / "training" / "loss_train.txt"))
    assert losses.shape == (2,) and np.all(np.isfinite(losses)) and losses[1] < losses[0]


def test_decoder_module_surface_and_lbfgs_cpu():
    """`Decoder` (models/codec.py:321-370 upstream) on the repo's modules: reference state_dict layout, any
    latent size, and torch.optim.LBFGS closures (several forward/backward pairs per step) drive it."""
    from models.codec import Decoder
    from models.darcy import (conv_boundary_condition, conv_constitutive_constraint_nonlinear,
                              conv_continuity_constraint)
    from utils.image_gradient import SobelFilter
    with cpu_backend():
        model = Decoder(2, 3, [3, 2], growth_rate=8, init_features=16)
        plan = orc.decoder_plan(2, 3, [3, 2], 8, 16)
        assert list(model.state_dict().keys()) == [n for n, _ in orc.state_layout(plan)]
        model.load_state_dict(orc.make_state(plan, 23))
        z = 0.5 * torch.randn(1, 2, 8, 8, generator=torch.Generator().manual_seed(0))
        K = orc.make_input(1, 32, 23)
        sob = SobelFilter(32, correct=True, device="cpu")
        model.train()
        opt = torch.optim.LBFGS(model.parameters(), lr=0.5, max_iter=4, history_size=50)
        hist = []

        def closure():
            opt.zero_grad()
            out = model(z)
            assert tuple(out.shape) == (1, 3, 32, 32)
            e = conv_constitutive_constraint_nonlinear(K, out, sob, 1.0, 0.7) + conv_continuity_constraint(out, sob)
            d, n = conv_boundary_condition(out)
            loss = e + (d + n) * 10.0
            loss.backward()
            hist.append(float(loss))
            return loss

        for _ in range(2):
            opt.step(closure)
        assert np.all(np.isfinite(hist)) and hist[-1] < hist[0]
        out4 = model(0.5 * torch.randn(1, 2, 4, 4))      # another latent size: 4 -> 16
        assert tuple(out4.shape) == (1, 3, 16, 16)


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "solve_conv_mixed_residual.py")),
                    reason="reference checkout not present (only in the build container)")
def test_unmodified_solver_script_runs(tmp_path):
    """solve_conv_mixed_residual.py (Decoder + L-BFGS on the conv losses), byte for
    byte, against this repo's modules (oracle-backed executor on CPU), linear law."""
    from pde_surrogate_b200 import data
    d = tmp_path / "datasets" / "64x64"
    d.mkdir(parents=True)
    x = data.grf_kle(10, 64, 64, 0.2, seed=3, device="cpu").numpy()
    data.write_hdf5(str(d / "kle512_lhs1000_test.hdf5"), x, data.darcy_fv_dataset(x))
    import run_reference_script
    argv = ["--script", os.path.join(REF, "solve_conv_mixed_residual.py"), "--", "--data-dir", str(tmp_path / "datasets"),
            "--exp-dir", str(tmp_path / "exp"), "--idx", "3", "--epochs", "2", "--test-freq", "1", "--ckpt-freq", "2",
            "--cuda", "0"]
    old_argv, old_path = list(sys.argv), list(sys.path)
    try:
        with cpu_backend():
            run_reference_script.main(argv)
    finally:
        sys.argv, sys.path[:] = old_argv, old_path
    losses = list((tmp_path / "exp").rglob("loss.txt"))
    assert len(losses) == 1
    vals = np.loadtxt(str(losses[0]))
    assert vals.shape == (2,) and np.all(np.isfinite(vals)) and vals[1] < vals[0]
    assert list((tmp_path / "exp").rglob("model_epoch2.pth"))


def test_dropout_option_host_logic():
    """DenseED(drop_rate > 0) builds, trains and evaluates through the module API (oracle-backed executor)."""
    from models.codec import DenseED
    from models.darcy import conv_boundary_condition
    with cpu_backend():
        model = DenseED(1, 3, 16, [1, 2, 1], growth_rate=4, init_features=8, drop_rate=0.3)
        K = orc.make_input(3, 16, 3)
        model.train()
        out = model(K)
        d, n = conv_boundary_condition(out)
        (d + n).backward()
        assert all(p.grad is not None and bool(torch.isfinite(p.grad).all()) for p in model.parameters())
        model.eval()
        with torch.no_grad():
            e1, e2 = model(K), model(K)
        assert torch.equal(e1, e2)          # no dropout in evaluation mode


def test_resident_loader_semantics():
    """utils.load.ResidentLoader = DataLoader(shuffle=True, drop_last=True) over device-resident tensors
    (CPU device here): every sample at most once per epoch, ragged tail dropped, fresh order every epoch."""
    from utils.load import ResidentLoader
    x = torch.arange(50, dtype=torch.float32).view(50, 1, 1, 1)
    y = -x.clone()
    ld = ResidentLoader([x, y], 8, "cpu")
    assert len(ld) == 6 and ld.dataset[3][1].numel() == 1
    torch.manual_seed(0)
    e1 = [b for b in ld]
    e2 = [b for b in ld]
    seen = torch.cat([b[0].flatten() for b in e1])
    assert seen.numel() == 48 and seen.unique().numel() == 48
    assert all(torch.equal(b[0], -b[1]) and b[0].shape == (8, 1, 1, 1) for b in e1)
    assert not torch.equal(seen, torch.cat([b[0].flatten() for b in e2]))


def test_coupling_layer_host_logic():
    """`AffineCouplingLayer` / `_DenseCoupling` module surface (models/glow_msc.py:276-344 upstream) with the
    oracle-backed executor: reference state_dict layout, gradients w.r.t. the flow variable and the conditioning."""
    from models.glow_msc import AffineCouplingLayer
    with cpu_backend():
        layer = AffineCouplingLayer(6, 9)
        plan = orc.coupling_plan(12, 6)
        assert list(layer.coupling_nn.state_dict().keys()) == [n for n, _ in orc.state_layout(plan)]
        layer.coupling_nn.load_state_dict(orc.make_state(plan, 55))
        layer.train()
        x = torch.randn(2, 6, 8, 8, requires_grad=True)
        cond = torch.randn(2, 9, 8, 8, requires_grad=True)
        y, logdet = layer.reverse(x, cond)
        (y.sum() + logdet.sum()).backward()
        assert x.grad is not None and cond.grad is not None and float(cond.grad.abs().sum()) > 0
        sd64 = orc.to_dtype(orc.make_state(plan, 55), torch.float64)
        x64, c64 = x.detach().double().requires_grad_(True), cond.detach().double().requires_grad_(True)
        y64, ld64 = orc.affine_coupling(plan, sd64, x64, c64, reverse=True)
        (y64.sum() + ld64.sum()).backward()
        assert rel(y.detach().numpy(), y64.detach().numpy()) < 1e-5
        assert rel(cond.grad.numpy(), c64.grad.numpy()) < 1e-4
    with pytest.raises(ImportError):
        from models.glow_msc import _CouplingNN  # noqa: F401  (the 'wide' coupling network is not built)
    from models.glow_msc import MultiScaleCondGlow  # noqa: F401


def test_fused_adam_class_falls_back_to_stock_step():
    """pde_surrogate_b200.optim.Adam on parameters that are not an executor network's (here: a CPU nn.Linear) is
    torch.optim.Adam, bit for bit; install() / uninstall() swap the name the scripts resolve."""
    from pde_surrogate_b200 import optim as pdes_optim
    stock = pdes_optim._StockAdam
    torch.manual_seed(0)
    a, b = torch.nn.Linear(5, 3), torch.nn.Linear(5, 3)
    b.load_state_dict(a.state_dict())
    oa, ob = pdes_optim.Adam(a.parameters(), lr=1e-2, weight_decay=1e-2), stock(b.parameters(), lr=1e-2, weight_decay=1e-2)
    x = torch.randn(7, 5)
    for _ in range(3):
        for m, o in ((a, oa), (b, ob)):
            o.zero_grad()
            m(x).pow(2).sum().backward()
            o.step()
    assert oa.fused_steps == 0
    assert all(torch.equal(p, q) for p, q in zip(a.parameters(), b.parameters()))
    try:
        assert pdes_optim.install() and torch.optim.Adam is pdes_optim.Adam and issubclass(torch.optim.Adam, stock)
    finally:
        pdes_optim.uninstall()
    assert torch.optim.Adam is stock


@pytest.mark.skipif(not os.path.isdir("/root/reference/models"), reason="needs the reference checkout (build container only)")
@pytest.mark.parametrize("kw", [dict(), dict(upsample=None), dict(bottleneck=True, bn_size=2, out_activation="tanh")])
def test_forward_test_prints_the_reference_trace(kw, capsys):
    """DenseED.forward_test / Decoder.forward_test (models/codec.py:298-304, 365-370): the printed shape trace is the
    reference's own, line for line (the reference is imported from /root/reference for this comparison only)."""
    import importlib.util
    import sys
    shim = os.path.join(ROOT, "pde_surrogate_b200", "_shims")
    added = importlib.util.find_spec("matplotlib") is None
    if added:
        sys.path.insert(0, shim)
    saved = {k: sys.modules.pop(k) for k in [m for m in sys.modules if m.split(".")[0] in ("models", "utils")]}
    sys.path.insert(0, "/root/reference")
    try:
        ref_codec = importlib.import_module("models.codec")
        assert ref_codec.__file__.startswith("/root/reference")
        ref = ref_codec.DenseED(1, 3, 32, [2, 3, 2], growth_rate=4, init_features=8, **kw)
        ref_dec = ref_codec.Decoder(2, 3, [2, 2], growth_rate=4, init_features=8)
        capsys.readouterr()
        ref.eval(), ref_dec.eval()
        with torch.no_grad():
            ref.forward_test(torch.zeros(2, 1, 32, 32))
            ref_dec.forward_test(torch.zeros(2, 2, 8, 8))
        want = capsys.readouterr().out
    finally:
        sys.path.remove("/root/reference")
        if added:
            sys.path.remove(shim)
        for k in [m for m in sys.modules if m.split(".")[0] in ("models", "utils")]:
            del sys.modules[k]
        sys.modules.update(saved)
    from pde_surrogate_b200.codec import DenseED, Decoder
    with cpu_backend():
        mine = DenseED(1, 3, 32, [2, 3, 2], growth_rate=4, init_features=8, **kw)
        mine_dec = Decoder(2, 3, [2, 2], growth_rate=4, init_features=8)
        capsys.readouterr()
        mine.eval(), mine_dec.eval()
        with torch.no_grad():
            mine.forward_test(torch.zeros(2, 1, 32, 32))
            mine_dec.forward_test(torch.zeros(2, 2, 8, 8))
        got = capsys.readouterr().out
    assert got == want


@pytest.mark.skipif(not os.path.isdir("/root/reference/models"), reason="needs the reference checkout (build container only)")
def test_exponential_law_losses_match_reference():
    """conv_constitutive_constraint_nonlinear_exp / energy_functional_exp (models/darcy.py:151-159, 193-207; unused by
    the scripts): values and gradients equal the reference's functions evaluated with the reference's SobelFilter."""
    import importlib.util
    import sys
    shim = os.path.join(ROOT, "pde_surrogate_b200", "_shims")
    added = importlib.util.find_spec("matplotlib") is None
    if added:
        sys.path.insert(0, shim)
    saved = {k: sys.modules.pop(k) for k in [m for m in sys.modules if m.split(".")[0] in ("models", "utils")]}
    sys.path.insert(0, "/root/reference")
    torch.manual_seed(3)
    K = torch.exp(0.3 * torch.randn(2, 1, 16, 16))
    out = (0.3 * torch.randn(2, 3, 16, 16)).requires_grad_(True)
    u = (0.3 * torch.randn(2, 1, 16, 16)).requires_grad_(True)
    try:
        ref_darcy = importlib.import_module("models.darcy")
        ref_sob = importlib.import_module("utils.image_gradient").SobelFilter(16, correct=True, device="cpu")
        assert ref_darcy.__file__.startswith("/root/reference")
        want = []
        for fn, arg in ((ref_darcy.conv_constitutive_constraint_nonlinear_exp, out), (ref_darcy.energy_functional_exp, u)):
            v = fn(K, arg, ref_sob)
            g, = torch.autograd.grad(v, arg)
            want.append((v.detach(), g))
    finally:
        sys.path.remove("/root/reference")
        if added:
            sys.path.remove(shim)
        for k in [m for m in sys.modules if m.split(".")[0] in ("models", "utils")]:
            del sys.modules[k]
        sys.modules.update(saved)
    from pde_surrogate_b200 import darcy as my_darcy
    from pde_surrogate_b200.image_gradient import SobelFilter
    with cpu_backend():
        sob = SobelFilter(16, correct=True, device="cpu")
        for (fn, arg), (v_ref, g_ref) in zip(((my_darcy.conv_constitutive_constraint_nonlinear_exp, out),
                                              (my_darcy.energy_functional_exp, u)), want):
            v = fn(K, arg, sob)
            g, = torch.autograd.grad(v, arg)
            assert torch.allclose(v, v_ref, rtol=1e-5, atol=1e-7) and torch.allclose(g, g_ref, rtol=1e-4, atol=1e-7)


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "train_cglow_reverse_kl.py")),
                    reason="reference checkout not present (only in the build container)")
def test_unmodified_cglow_reverse_kl_script_runs(tmp_path):
    """SURVEY section 8(f) row 1 / BASELINE config 5: train_cglow_reverse_kl.py, byte for byte, on this repo's
    models.glow_msc (coupling networks on the executor - here its oracle-backed stand-in - flow plumbing in PyTorch),
    the fused Darcy losses and the harness modules: two epochs of reverse-KL training, the test pass with
    `model.predict` / `model.sample`, checkpoint and statistics."""
    from pde_surrogate_b200 import data
    d = tmp_path / "datasets" / "32x32"
    d.mkdir(parents=True)
    rs = np.random.RandomState(2)
    x = data.grf_kle(16, 32, 64, 0.2, seed=3, device="cpu").numpy()
    data.write_hdf5(str(d / "kle100_lhs10000_train.hdf5"), x[:8])
    data.write_hdf5(str(d / "kle100_lhs1000_val.hdf5"), x[8:], rs.standard_normal((8, 3, 32, 32)))
    import run_reference_script
    argv = ["--script", os.path.join(REF, "train_cglow_reverse_kl.py"), "--", "--data-dir", str(tmp_path / "datasets"),
            "--exp-dir", str(tmp_path / "exp"), "--imsize", "32", "--ntrain", "8", "--ntest", "8", "--batch-size", "4",
            "--test-batch-size", "8", "--epochs", "2", "--cuda", "0", "--plot-freq", "1000", "--ckpt-freq", "2"]
    # (two epochs: the script divides by (epochs - 1) * steps_per_epoch, train_cglow_reverse_kl.py:233, 268; a test
    # batch of >= 6: its last epoch plots six members of the first test batch, 198-202)
    old_argv, old_path = list(sys.argv), list(sys.path)
    try:
        with cpu_backend():
            run_reference_script.main(argv)
    finally:
        sys.argv, sys.path[:] = old_argv, old_path
    run = list((tmp_path / "exp").rglob("args.txt"))[0].parent
    ck = torch.load(str(run / "checkpoints" / "model_epoch2.pth"), weights_only=False)
    assert "flow.revblock3.revlayers.revlayer6.coupling.coupling_nn.reduce.conv_zero.conv.weight" in ck["model_state_dict"]
    assert (run / "training" / "loss_train.txt").exists() and (run / "training" / "entropy_test.txt").exists()
    assert np.isfinite(np.loadtxt(str(run / "training" / "loss_train.txt"))).all()
