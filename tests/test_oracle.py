"""Pins the CPU oracle (oracle/pdes_oracle.py) against fixtures produced by the unmodified
reference (tests/golden/make_golden.py).  Runs without a GPU."""
import os

import numpy as np
import pytest
import torch

from oracle import pdes_oracle as orc

CASES = ["densenet_small16", "densenet_fiveblk16", "densenet_full32", "densenet_full64", "densenet_full64_channel",
         "densenet_full32_b32", "densenet_full64_b32",   # *_b32: the batch bench.py times (fields stored as fp32)
         "densenet_bilinear16", "densenet_bilinear32",   # DenseED(upsample='bilinear')
         "densenet_convt16", "densenet_convt32",         # DenseED(upsample=None): nn.ConvTranspose2d transitions
         "densenet_bottleneck16", "densenet_bottleneck32", "densenet_bottleneck32b"]   # DenseED(bottleneck=True, bn_size=...)


def _ups(name):
    return "bilinear" if "bilinear" in name else (None if "convt" in name else "nearest")


def _bn(g):
    return int(g["bn_size"]) if "bn_size" in g.files else 0


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"))


def _cfg(g):
    return dict(in_channels=int(g["cfg_in_channels"]), out_channels=int(g["cfg_out_channels"]),
                imsize=int(g["cfg_imsize"]), blocks=[int(b) for b in g["cfg_blocks"]],
                growth_rate=int(g["cfg_growth_rate"]), init_features=int(g["cfg_init_features"]))


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


@pytest.mark.parametrize("name", CASES)
def test_structure_matches_reference(golden_dir, name):
    g = _load(golden_dir, name)
    plan = orc.densenet_plan(**_cfg(g), upsample=_ups(name), bottleneck=_bn(g))
    assert orc.param_names(plan) == [str(s) for s in g["param_names"]]
    n_params = sum(int(np.prod(s)) for n, s in orc.state_layout(plan)
                   if n in set(orc.param_names(plan)))
    n_conv = sum(1 for n in orc.param_names(plan) if "conv" in n)
    assert (n_params, n_conv) == tuple(int(v) for v in g["model_size"])


@pytest.mark.parametrize("name", CASES)
def test_train_step_fp64_matches_reference(golden_dir, name):
    torch.set_num_threads(4)
    g = _load(golden_dir, name)
    cfg = _cfg(g)
    plan = orc.densenet_plan(**cfg, upsample=_ups(name), bottleneck=_bn(g))
    sd = orc.to_dtype(orc.make_state(plan, int(g["seed"])), torch.float64)
    K = orc.make_input(int(g["B"]), cfg["imsize"], int(g["seed"]), kind=str(g["input_kind"]) if "input_kind" in g.files else "lognormal").double()
    sd_eval = orc.to_dtype(sd, torch.float64)
    with torch.no_grad():
        out_eval = orc.densenet_forward(plan, sd_eval, K, training=False, upsample=_ups(name))
    ftol = 1e-7 if "compact" in g.files else 1.0   # compact fixtures hold the fp64 fields rounded to fp32
    assert rel(out_eval.numpy(), g["out_eval64"]) < (ftol if ftol < 1.0 else 1e-12)
    out, l4, loss, dout, grads = orc.train_step(plan, sd, K, upsample=_ups(name))
    assert rel(out.numpy(), g["out64"]) < (ftol if ftol < 1.0 else 1e-11)
    assert rel(l4.numpy(), g["l4_64"]) < 1e-11
    assert abs(float(loss) - float(g["loss64"])) / float(g["loss64"]) < 1e-11
    assert rel(dout.numpy(), g["dout64"]) < (ftol if ftol < 1.0 else 1e-10)
    names = [str(s) for s in g["param_names"]]
    norms = np.array([float(grads[n].norm()) for n in names])
    assert np.allclose(norms, g["grad_norm64"], rtol=1e-8, atol=1e-300)
    if "grads64" in g.files:
        flat = np.concatenate([grads[n].numpy().ravel() for n in names])
        assert rel(flat, g["grads64"]) < 1e-9
    else:
        head = np.concatenate([grads[n].numpy().ravel()[:16] for n in names])
        assert rel(head, g["grads64_head"]) < 1e-8
    run = np.concatenate([sd[str(n)].detach().numpy().ravel() for n in g["running_names"]])
    assert rel(run, g["running64"]) < 1e-12
    nbt = [int(v) for k, v in sd.items() if k.endswith("num_batches_tracked")]
    assert nbt == [int(v) for v in g["num_batches_tracked"]]


@pytest.mark.parametrize("name", CASES[:3])
def test_train_step_fp32_within_noise_floor(golden_dir, name):
    torch.set_num_threads(1)
    g = _load(golden_dir, name)
    cfg = _cfg(g)
    plan = orc.densenet_plan(**cfg)
    sd = orc.make_state(plan, int(g["seed"]))
    K = orc.make_input(int(g["B"]), cfg["imsize"], int(g["seed"]), kind=str(g["input_kind"]) if "input_kind" in g.files else "lognormal")
    out, l4, loss, dout, grads = orc.train_step(plan, sd, K)
    assert rel(out.numpy(), g["out"]) < 2e-5
    assert rel(l4.numpy(), g["l4"]) < 1e-5
    assert rel(dout.numpy(), g["dout"]) < 2e-5


def test_sobel_and_losses_match_reference(golden_dir):
    g = _load(golden_dir, "sobel_darcy")
    for tag in "abc":
        img = torch.tensor(g[f"{tag}_img"], requires_grad=True)
        for c in (1, 0):
            gh, gv = orc.sobel_grad_h(img, bool(c)), orc.sobel_grad_v(img, bool(c))
            assert rel(gh.detach().numpy(), g[f"{tag}{c}_gh"]) < 1e-13
            assert rel(gv.detach().numpy(), g[f"{tag}{c}_gv"]) < 1e-13
            w = torch.tensor(g[f"{tag}{c}_w"])
            ah, = torch.autograd.grad((gh * w).sum(), img, retain_graph=True)
            av, = torch.autograd.grad((gv * w).sum(), img)
            assert rel(ah.numpy(), g[f"{tag}{c}_ah"]) < 1e-13
            assert rel(av.numpy(), g[f"{tag}{c}_av"]) < 1e-13
    for tag in "pqr":
        K = torch.tensor(g[f"{tag}_K"])
        out = torch.tensor(g[f"{tag}_out"], requires_grad=True)
        gw = torch.tensor(g[f"{tag}_gw"])
        for tb in (1, 0):
            l4 = orc.darcy_losses(K, out, use_tb=bool(tb))
            assert rel(l4.detach().numpy(), g[f"{tag}{tb}_l4"]) < 1e-13
            d, = torch.autograd.grad((gw * l4).sum(), out)
            assert rel(d.numpy(), g[f"{tag}{tb}_dout"]) < 1e-12


def test_adam_reference_matches_torch_optim():
    """oracle.adam_reference (the checker of the fused Adam kernel) against torch.optim.Adam, with and
    without weight decay (train_codec_mixed_residual.py:151, 239)."""
    gen = torch.Generator().manual_seed(0)
    for wd in (0.0, 1e-2):
        p = torch.randn(257, generator=gen, dtype=torch.float64).requires_grad_(True)
        opt = torch.optim.Adam([p], lr=1e-3, weight_decay=wd)
        q, m, v = p.detach().clone(), torch.zeros(257, dtype=torch.float64), torch.zeros(257, dtype=torch.float64)
        for step in (1, 2, 3, 4, 5):
            g = torch.randn(257, generator=gen, dtype=torch.float64)
            lr = 1e-3 * step
            for grp in opt.param_groups:
                grp["lr"] = lr
            p.grad = g.clone()
            opt.step()
            q, m, v = orc.adam_reference(q, g, m, v, lr, step, wd=wd)
            assert rel(q.numpy(), p.detach().numpy()) < 1e-13


def test_finite_volume_reference_solver():
    """pde_surrogate_b200.data.darcy_fv_solve (stand-in for the FEniCS labels, utils/fenics.py upstream):
    exact for constant permeability, discretely conservative, satisfies the boundary conditions, and its
    fields make the reference's mixed-residual loss small compared with a perturbed field."""
    from pde_surrogate_b200 import data
    u, s1, s2 = data.darcy_fv_solve(np.full((16, 16), 2.5))
    assert np.allclose(u[3], np.linspace(1.0, 0.0, 16)) and np.allclose(s2, 0.0)
    assert np.allclose(s1, 2.5 * 16.0 / 15.0)
    rs = np.random.RandomState(0)
    K = np.exp(0.5 * rs.standard_normal((32, 32)))
    o = data.darcy_fv_solve(K)
    assert np.allclose(o[0][:, 0], 1.0) and np.allclose(o[0][:, -1], 0.0)
    assert o[0].min() >= -1e-12 and o[0].max() <= 1.0 + 1e-12      # discrete maximum principle
    col_flux = o[1].sum(0)[1:-1]
    assert np.ptp(col_flux) < 1e-10 * abs(col_flux.mean())          # the same total flux crosses every column
    Kt = torch.tensor(K[None, None])
    good = orc.total_loss(Kt, torch.tensor(o[None]))[0]
    bad = orc.total_loss(Kt, torch.tensor(o[None] + 0.05 * rs.standard_normal(o[None].shape)))[0]
    assert float(good) < 0.2 * float(bad)


@pytest.mark.parametrize("tag", ["a", "b"])
def test_decoder_and_nonlinear_law_match_reference(golden_dir, tag):
    """the oracle's `Decoder` plan (models/codec.py:321-370) and nonlinear
    constitutive law (models/darcy.py:179-191) against the reference's own outputs (decoder_solver.npz)."""
    torch.set_num_threads(4)
    g = _load(golden_dir, "decoder_solver")
    nz, hz, gr, feat, seed, imsize = [int(v) for v in g[f"{tag}_cfg"]]
    blocks = [int(b) for b in g[f"{tag}_blocks"]]
    plan = orc.decoder_plan(nz, 3, blocks, gr, feat)
    assert orc.param_names(plan) == [str(s) for s in g[f"{tag}_param_names"]]
    z = torch.tensor(g[f"{tag}_z"])
    K = orc.make_input(1, imsize, seed).double()
    a1, a2 = [float(v) for v in g[f"{tag}_alphas"]]
    for law in ("lin", "nl"):
        sd = orc.to_dtype(orc.make_state(plan, seed), torch.float64)
        names = orc.param_names(plan)
        for n in names:
            sd[n].requires_grad_(True)
        out = orc.densenet_forward(plan, sd, z, training=True)
        out.retain_grad()
        e = orc.constitutive_nonlinear(K, out, a1, a2) if law == "nl" else orc.constitutive(K, out)
        d_, n_ = orc.boundary(out)
        l4 = torch.stack([e, orc.continuity(out), d_, n_])
        loss = (l4[0] + l4[1]) + (l4[2] + l4[3]) * 10.0
        loss.backward()
        assert rel(out.detach().numpy(), g[f"{tag}_{law}_out64"]) < 1e-11
        assert rel(l4.detach().numpy(), g[f"{tag}_{law}_l4_64"]) < 1e-11
        assert rel(out.grad.numpy(), g[f"{tag}_{law}_dout64"]) < 1e-10
        norms = np.array([float(sd[n].grad.norm()) for n in names])
        assert np.allclose(norms, g[f"{tag}_{law}_grad_norm64"], rtol=1e-8, atol=1e-300)
        head = np.concatenate([sd[n].grad.numpy().ravel()[:16] for n in names])
        assert rel(head, g[f"{tag}_{law}_grads64_head"]) < 1e-8


def test_dropout_step_matches_reference(golden_dir):
    """DenseED(drop_rate=0.2): nn.Dropout2d behind the convolutions (models/codec.py:70-71, 110-149, 171-172).
    With the CPU generator seeded right before the forward pass the oracle draws the reference's masks."""
    torch.set_num_threads(1)
    g = _load(golden_dir, "densenet_dropout16")
    cfg = dict(in_channels=1, out_channels=3, imsize=int(g["cfg_imsize"]), blocks=[int(b) for b in g["cfg_blocks"]],
               growth_rate=int(g["cfg_growth_rate"]), init_features=int(g["cfg_init_features"]))
    plan = orc.densenet_plan(**cfg)
    assert sum(1 for s in plan if orc._drop_site(s)) == int(g["n_dropout_modules"])
    sd = orc.make_state(plan, int(g["seed"]))
    names = orc.param_names(plan)
    for n in names:
        sd[n].requires_grad_(True)
    K = orc.make_input(int(g["B"]), cfg["imsize"], int(g["seed"]))
    torch.manual_seed(int(g["mask_seed"]))
    out = orc.densenet_forward(plan, sd, K, training=True, drop_rate=float(g["drop_rate"]))
    out.retain_grad()
    loss, l4 = orc.total_loss(K, out)
    loss.backward()
    assert rel(out.detach().numpy(), g["out"]) < 1e-5
    assert rel(l4.detach().numpy(), g["l4"]) < 1e-5
    assert rel(out.grad.numpy(), g["dout"]) < 1e-5
    flat = np.concatenate([sd[n].grad.numpy().ravel() for n in names])
    assert rel(flat, g["grads"]) < 1e-3
    with torch.no_grad():
        ev = orc.densenet_forward(plan, sd, K, training=False, drop_rate=float(g["drop_rate"]))
    assert rel(ev.numpy(), g["out_eval"]) < 1e-5


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_cglow_coupling_matches_reference(golden_dir, tag):
    """the oracle's `_DenseCoupling` / Conv2dZeros / AffineCouplingLayer restatement
    (models/glow_msc.py:240-255, 276-294, 326-344) against the reference's own fp64 outputs and gradients
    (w.r.t. parameters, the flow variable and the conditioning)."""
    torch.set_num_threads(4)
    g = _load(golden_dir, "cglow_coupling")
    fin, fcond, H, B, seed, cin, cout = [int(v) for v in g[f"{tag}_cfg"]]
    plan = orc.coupling_plan(cin, cout)
    names = orc.param_names(plan)
    assert names == [str(s) for s in g[f"{tag}_param_names"]]
    rs = np.random.RandomState(900 + seed)
    x0 = rs.standard_normal((B, fin, H, H))
    c0 = rs.standard_normal((B, fcond, H, H))
    wy = torch.tensor(rs.standard_normal((B, fin, H, H)))
    for mode in ("fwd", "rev"):
        sd = orc.to_dtype(orc.make_state(plan, seed), torch.float64)
        for n in names:
            sd[n].requires_grad_(True)
        x = torch.tensor(x0, requires_grad=True)
        cond = torch.tensor(c0, requires_grad=True)
        y, logdet = orc.affine_coupling(plan, sd, x, cond, reverse=(mode == "rev"))
        ((y * wy).sum() + 0.3 * logdet.sum()).backward()
        assert rel(y.detach().numpy(), g[f"{tag}_{mode}_y"]) < 1e-6        # (stored as fp32)
        assert rel(logdet.detach().numpy(), g[f"{tag}_{mode}_logdet"]) < 1e-11
        assert rel(x.grad.numpy(), g[f"{tag}_{mode}_dx"]) < 1e-6
        assert rel(cond.grad.numpy(), g[f"{tag}_{mode}_dcond"]) < 1e-6
        flat = np.concatenate([sd[n].grad.numpy().ravel() for n in names])
        assert rel(flat, g[f"{tag}_{mode}_grads"]) < 1e-9
