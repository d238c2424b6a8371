"""DenseED executor parity: training/eval forward, losses, dL/d(output), parameter gradients,
running statistics — against the reference-generated fixtures and the fp64 oracle.

Bars: fields and the four partial losses within 1e-4 relative of the
reference; gradients within 3x the reference's own fp32-vs-fp64 error (noise-floor criterion),
measured against the fp64 reference."""
import os

import numpy as np
import pytest
import torch

from oracle import pdes_oracle as orc

pytestmark = pytest.mark.gpu
CASES = ["densenet_small16", "densenet_fiveblk16", "densenet_full32", "densenet_full64", "densenet_full64_channel",
         # batch 32 = the shape bench.py times: 256-1024 pixel tiles per layer, so every persistent CTA of the
         # tensor-core kernels walks several tiles (accumulator-stage ring, operand ring wrap, multi-tile split-K)
         "densenet_full32_b32", "densenet_full64_b32",
         # DenseED(upsample='bilinear') (train_codec_mixed_residual.py --upsample bilinear)
         "densenet_bilinear16", "densenet_bilinear32",
         # DenseED(upsample=None): nn.ConvTranspose2d transitions, output imsize/2 wide (models/codec.py:139-142, 176-179)
         "densenet_convt16", "densenet_convt32",
         # DenseED(bottleneck=True, bn_size=...): 1x1 -> 3x3 dense layers above bn_size * growth channels (codec.py:56-64)
         # (the deeper `densenet_bottleneck32` fixture - 46 convolutions - stays a CPU oracle pin: on the GPU it is a
         # ReLU-mask-flip lottery; tools/diag_fixture.py shows the CUDA-core path flipping in
         # DecBlock2.denselayer3 and the tensor-core path not at all)
         "densenet_bottleneck16", "densenet_bottleneck32b"]


FLIP_PRONE = {"densenet_bottleneck32b"}   # see test_train_step_matches_reference


def _ups(name):
    return "bilinear" if "bilinear" in name else (None if "convt" in name else "nearest")


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def _cfg(g):
    return dict(in_channels=int(g["cfg_in_channels"]), out_channels=int(g["cfg_out_channels"]),
                imsize=int(g["cfg_imsize"]), blocks=[int(b) for b in g["cfg_blocks"]],
                growth_rate=int(g["cfg_growth_rate"]), init_features=int(g["cfg_init_features"]))


def _model(g, upsample="nearest"):
    from models.codec import DenseED
    cfg = _cfg(g)
    bn_size = int(g["bn_size"]) if "bn_size" in g.files else 0
    plan = orc.densenet_plan(**cfg, upsample=upsample, bottleneck=bn_size)
    sd = orc.make_state(plan, int(g["seed"]))
    model = DenseED(cfg["in_channels"], cfg["out_channels"], cfg["imsize"], cfg["blocks"],
                    growth_rate=cfg["growth_rate"], init_features=cfg["init_features"], upsample=upsample,
                    **(dict(bottleneck=True, bn_size=bn_size) if bn_size else {}))
    assert list(model.state_dict().keys()) == list(sd.keys())
    model.load_state_dict(sd)
    model = model.to("cuda")
    K = orc.make_input(int(g["B"]), cfg["imsize"], int(g["seed"]), kind=str(g["input_kind"]) if "input_kind" in g.files else "lognormal").to("cuda")
    return model, K, cfg


@pytest.mark.parametrize("impl", [0, 1])
@pytest.mark.parametrize("name", CASES)
def test_train_step_matches_reference(golden_dir, name, impl):
    from models.darcy import conv_boundary_condition, conv_constitutive_constraint, conv_continuity_constraint
    from utils.image_gradient import SobelFilter
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    model, K, cfg = _model(g, _ups(name))
    model.conv_impl = impl  # 0: tcgen05 (two-piece fp16 operands) where supported, 1: CUDA-core fp32 everywhere
    assert tuple(model.model_size) == tuple(int(v) for v in g["model_size"])
    osz = cfg["imsize"] // 2 if _ups(name) is None else cfg["imsize"]
    Kl = K[:, :, ::2, ::2].contiguous() if _ups(name) is None else K   # (as the fixture generator does)
    sob = SobelFilter(osz, correct=True, device="cuda")
    # eval forward with the given running statistics
    model.eval()
    with torch.no_grad():
        out_eval = model(K)
    assert rel(out_eval.cpu().numpy(), g["out_eval64"]) < 1e-4
    # the step body of train_codec_mixed_residual.py:226-233
    model.train()
    model.zero_grad()
    out = model(K)
    assert out.shape[-1] == osz
    out.retain_grad()
    l_c = conv_constitutive_constraint(Kl, out, sob)
    l_d = conv_continuity_constraint(out, sob)
    l_dir, l_neu = conv_boundary_condition(out)
    loss = (l_c + l_d) + (l_dir + l_neu) * 10.0
    loss.backward()
    torch.cuda.synchronize()
    assert rel(out.detach().cpu().numpy(), g["out64"]) < 1e-4
    l4 = torch.stack([l_c, l_d, l_dir, l_neu]).detach().cpu().numpy()
    assert np.all(np.abs(l4 - g["l4_64"]) <= 1e-4 * np.abs(g["l4_64"])), (l4, g["l4_64"])
    assert abs(float(loss) - float(g["loss64"])) <= 1e-4 * float(g["loss64"])
    assert rel(out.grad.cpu().numpy(), g["dout64"]) < 1e-4
    # parameter gradients vs the fp64 reference, noise-floor relative
    names = [str(s) for s in g["param_names"]]
    params = dict(model.named_parameters())
    assert list(params.keys()) == names
    # Noise-floor criterion with ReLU-mask flips allowed: an element whose fp64 pre-activation lies
    # within the fp32 forward error of zero flips its mask in ANY fp32 implementation (the reference's
    # own fp32-vs-fp64 parameter gradients differ by 1.4e-3 at 64x64 / batch 32 for that reason,
    # SURVEY.md section 7.7), and one flip moves the few BatchNorm gradients it feeds by up to ~1e-2.
    # So: the typical tensor must sit at the reference's own fp32 noise floor, nearly all tensors
    # within 3x of it, and the flip-affected remainder stays bounded.
    ratios, rels = [], []
    tot_err, tot_norm = 0.0, 0.0
    pos = 0
    for i, n in enumerate(names):
        gr = params[n].grad.detach().double().cpu().numpy().ravel()
        if "grads64" in g.files:
            ref = g["grads64"][pos:pos + gr.size]
            pos += gr.size
            err = np.linalg.norm(gr - ref)
        else:
            k = int(g["grads64_head_len"][i])
            ref = g["grads64_head"][pos:pos + k]
            pos += k
            err = np.linalg.norm(gr[:k] - ref) * np.sqrt(gr.size / k)
            assert abs(np.linalg.norm(gr) - g["grad_norm64"][i]) <= 3 * g["grad_err32"][i] + 2e-2 * g["grad_norm64"][i]
        norm = float(g["grad_norm64"][i])
        ratios.append(err / max(3 * float(g["grad_err32"][i]), 1e-5 * norm))
        rels.append(err / max(norm, 1e-300))
        tot_err += err ** 2
        tot_norm += norm ** 2
    ratios, rels = np.array(ratios), np.array(rels)
    if name in FLIP_PRONE:
        # Bottleneck networks (a 1x1 convolution + BatchNorm + ReLU in front of every 3x3 one) flip a ReLU mask in most
        # fp32 runs, and WHICH element flips depends on the rounding of that run (tools/diag_fixture.py: the same
        # fixture sits at 0.2-0.3 x the bar in one run and flips in DecBlock2 in the next).  Everything downstream of
        # the flip must still be at the floor, the rest bounded; the kernels themselves are checked flip-free by
        # test_tensor_core_backward_strict_per_tensor on this fixture.
        assert np.mean(ratios <= 1.0) >= 0.15, "only %.0f%% of the tensors at the noise floor" % (100 * np.mean(ratios <= 1.0))
        assert rels.max() <= 0.3 and np.sqrt(tot_err / tot_norm) <= 3e-2, (rels.max(), np.sqrt(tot_err / tot_norm))
    else:
        assert np.median(ratios) <= 1.0, "typical tensor %.2f x the 3x-noise-floor bar" % np.median(ratios)
        # (a flip deep in the decoder perturbs every gradient upstream of it, i.e. up to the whole encoder)
        assert np.mean(ratios <= 1.0) >= 0.5, "only %.0f%% of the tensors at the noise floor" % (100 * np.mean(ratios <= 1.0))
        assert rels.max() <= 0.3 and np.sqrt(tot_err / tot_norm) <= 5e-3, (rels.max(), np.sqrt(tot_err / tot_norm))
    # running statistics and step counters
    sd = model.state_dict()
    run = np.concatenate([sd[str(n)].double().cpu().numpy().ravel() for n in g["running_names"]])
    assert rel(run, g["running64"]) < 1e-5
    nbt = [int(v) for k, v in sd.items() if k.endswith("num_batches_tracked")]
    assert nbt == [int(v) for v in g["num_batches_tracked"]]


@pytest.mark.parametrize("name", ["densenet_fiveblk16", "densenet_full32_b32"])
def test_train_step_legacy_thin_layer_path(golden_dir, name, monkeypatch):
    """The round-1 launch structure of the thin layers (operand split + conv_tc2 forward, dY split + conv_tc2
    dgrad) stays selectable (PDES_DENSE_FWD=0 / PDES_DENSE_BWD=0) and stays correct."""
    monkeypatch.setenv("PDES_DENSE_FWD", "0")
    monkeypatch.setenv("PDES_DENSE_BWD", "0")
    test_train_step_matches_reference(golden_dir, name, 0)


def test_grad_accumulation_and_zero_grad(golden_dir):
    from models.darcy import conv_boundary_condition
    g = np.load(os.path.join(golden_dir, "densenet_small16.npz"))
    model, K, cfg = _model(g)
    model.train()

    def run():
        out = model(K)
        d, n = conv_boundary_condition(out)
        (d + n).backward()

    model.zero_grad()
    run()
    g1 = model.flat_parameters()[1].clone()
    run()  # no zero_grad: accumulates (BN running stats moved, batch stats identical)
    g2 = model.flat_parameters()[1].clone()
    assert rel(g2.cpu().numpy(), 2 * g1.cpu().numpy()) < 1e-4
    model.zero_grad()
    assert all(p.grad is None for p in model.parameters())
    run()
    assert rel(model.flat_parameters()[1].cpu().numpy(), g1.cpu().numpy()) < 1e-4


def test_adam_step_matches_torch(golden_dir):
    """The unmodified script's torch.optim.Adam works on the flat-view parameters, and the fused
    flat Adam kernel reproduces it."""
    from pde_surrogate_b200 import _lib
    g = np.load(os.path.join(golden_dir, "densenet_small16.npz"))
    model, K, cfg = _model(g)
    from models.darcy import conv_boundary_condition
    model.train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    flat, gflat = model.flat_parameters()
    p0 = flat.clone()
    m = torch.zeros_like(flat)
    v = torch.zeros_like(flat)
    mine = flat.clone()
    for step in (1, 2, 3):
        model.zero_grad()
        d, n = conv_boundary_condition(model(K))
        (d + n).backward()
        _lib.check(_lib.lib().pdes_adam_step(_lib.ptr(mine), _lib.ptr(gflat), _lib.ptr(m), _lib.ptr(v),
                                             mine.numel(), 1e-3, 0.9, 0.999, 1e-8, 0.0, 1.0, step,
                                             _lib.stream_ptr()))
        opt.step()
        assert rel(flat.cpu().numpy() - p0.cpu().numpy(), mine.cpu().numpy() - p0.cpu().numpy()) < 1e-5


def test_errors_are_loud():
    from models.codec import DenseED
    with pytest.raises(ValueError):
        DenseED(1, 3, 64, [6, 8])
    with pytest.raises(NotImplementedError):
        DenseED(1, 3, 64, [6, 8, 6], upsample='bicubic')
    with pytest.raises(ValueError):
        DenseED(1, 3, 64, [6, 8, 6], drop_rate=1.5)
    m = DenseED(1, 3, 16, [1, 1, 1], growth_rate=4, init_features=8)
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 1, 16, 16))  # CPU: no fallback
    m = m.cuda()
    with pytest.raises(ValueError):
        m(torch.zeros(1, 1, 32, 32, device="cuda"))


def test_engine_graph_replay_matches_eager(golden_dir):
    """TrainStep.step (eager launches) and TrainStep.step_graph (CUDA-graph replay with device-side
    Adam scalars) walk the same trajectory: capture() snapshots and restores everything its warm-up
    mutates, so replay i is optimisation step i."""
    from pde_surrogate_b200.engine import TrainStep
    g = np.load(os.path.join(golden_dir, "densenet_fiveblk16.npz"))
    losses, finals = {}, {}
    for mode in ("eager", "graph"):
        model, K, cfg = _model(g)
        ts = TrainStep(model, lr=2e-3)
        out = []
        for i in range(5):
            lr = 2e-3 * (1.0 + 0.1 * i)   # a per-step schedule: every replay must see ITS learning rate
            loss = ts.step_graph(K, lr=lr) if mode == "graph" else ts.step(K, lr=lr)
            out.append(float(loss))
        losses[mode] = out
        sd = model.state_dict()
        finals[mode] = (model.flat_parameters()[0].detach().cpu().numpy().copy(),
                        np.concatenate([v.double().cpu().numpy().ravel() for k, v in sd.items()
                                        if k.endswith(("running_mean", "running_var"))]),
                        [int(v) for k, v in sd.items() if k.endswith("num_batches_tracked")])
    assert abs(losses["eager"][0] - float(g["loss"])) <= 1e-4 * float(g["loss"])
    for i in range(5):
        assert abs(losses["graph"][i] - losses["eager"][i]) <= 2e-4 * abs(losses["eager"][i]), (i, losses)
    assert rel(finals["graph"][0], finals["eager"][0]) < 1e-4
    assert rel(finals["graph"][1], finals["eager"][1]) < 1e-5
    assert finals["graph"][2] == finals["eager"][2] == [5] * len(finals["eager"][2])


@pytest.mark.parametrize("imsize", [32, 64])
def test_trajectory_batch32_matches_reference(golden_dir, imsize):
    """Three optimisation steps at the timed shape (batch 32, DenseED[6,8,6]; 64x64 = BASELINE config 2)
    through the CUDA-graph engine against the trajectory of the reference itself (its DenseED, losses and
    torch.optim.Adam, train_codec_mixed_residual.py:226-240; tests/golden/trajectory_b32.npz holds its fp32
    and fp64 runs).  Step 1 must match to 1e-4.  Later steps are chaotic (ReLU flips, sign-like first Adam
    steps on noise-level gradients): the reference's own fp32 run leaves its fp64 run by 1e-5..5e-4 in the
    loss and by 7 % in the parameter displacement; the bar is 4x / 2x that self-noise, measured against
    the fp64 run.  The oracle's CPU loop (same torch kernels) is stepped alongside as a second witness."""
    from oracle.cpu_train import CpuTrainer
    from pde_surrogate_b200.engine import TrainStep
    from models.codec import DenseED
    t = np.load(os.path.join(golden_dir, "trajectory_b32.npz"))
    l32, l64 = t["loss32_%d" % imsize], t["loss64_%d" % imsize]
    torch.set_num_threads(max(1, min(16, os.cpu_count() or 1)))
    cpu = CpuTrainer(imsize, lr=1e-3, seed=1)
    plan = cpu.plan
    sd = orc.make_state(plan, 1)
    model = DenseED(1, 3, imsize, [6, 8, 6])
    model.load_state_dict(sd)
    model = model.to("cuda")
    ts = TrainStep(model, lr=1e-3)
    ref, got = [], []
    for i, lr in enumerate(t["lrs"]):
        K = orc.make_input(32, imsize, 100 + i)
        ref.append(cpu.step(K, lr=float(lr)))
        got.append(float(ts.step_graph(K.cuda(), lr=float(lr))))
    assert abs(ref[0] - l32[0]) <= 1e-5 * abs(l32[0])            # the oracle loop IS the reference loop
    assert abs(got[0] - l64[0]) <= 1e-4 * abs(l64[0]), (got, l64)
    for i in range(3):
        bar = max(1e-4 * abs(l64[i]), 4.0 * abs(l32[i] - l64[i]))
        assert abs(got[i] - l64[i]) <= bar, (i, got, list(l64), list(l32))
    names = orc.param_names(plan)
    params = dict(model.named_parameters())
    st = int(t["stride"])
    p_got = np.concatenate([params[n].detach().double().cpu().numpy().ravel() for n in names])[::st]
    p_0 = np.concatenate([sd[n].double().numpy().ravel() for n in names])[::st]
    p32, p64 = t["params32_%d" % imsize], t["params64_%d" % imsize]
    self_noise = rel(p32 - p_0, p64 - p_0)
    assert rel(p_got - p_0, p64 - p_0) <= 2.0 * self_noise + 1e-3, (rel(p_got - p_0, p64 - p_0), self_noise)
    assert rel(p_got, p64) <= 2.0 * rel(p32, p64) + 1e-5


def test_fused_adam_weight_decay_and_grad_scale():
    """pdes_adam_step / pdes_adam_step_dev with weight_decay != 0 and grad_scale != 1 against the oracle's
    restatement of torch.optim.Adam (train_codec_mixed_residual.py:151, 239)."""
    from pde_surrogate_b200 import _lib
    L = _lib.lib()
    gen = torch.Generator().manual_seed(3)
    n = 10007
    p0 = torch.randn(n, generator=gen)
    pad = (-n) % 4
    mk = lambda t: torch.cat([t, torch.zeros(pad)]).cuda()
    for wd, gs in ((0.0, 1.0), (1e-2, 1.0), (5e-4, 0.125)):
        p_ref, m_ref, v_ref = p0.double(), torch.zeros(n, dtype=torch.float64), torch.zeros(n, dtype=torch.float64)
        p, m, v = mk(p0), mk(torch.zeros(n)), mk(torch.zeros(n))
        p2, m2, v2 = p.clone(), m.clone(), v.clone()
        hyper_h = torch.zeros(8)
        for step in (1, 2, 3, 4):
            graw = torch.randn(n, generator=gen)
            lr = 1e-3 * step
            p_ref, m_ref, v_ref = orc.adam_reference(p_ref, graw.double() * gs, m_ref, v_ref, lr, step, wd=wd)
            gd = mk(graw)
            _lib.check(L.pdes_adam_step(_lib.ptr(p), _lib.ptr(gd), _lib.ptr(m), _lib.ptr(v), n, lr, 0.9, 0.999,
                                        1e-8, wd, gs, step, _lib.stream_ptr()))
            _lib.check(L.pdes_adam_hyper(hyper_h.data_ptr(), lr, 0.9, 0.999, 1e-8, wd, gs, step))
            hd = hyper_h.cuda()
            _lib.check(L.pdes_adam_step_dev(_lib.ptr(p2), _lib.ptr(gd), _lib.ptr(m2), _lib.ptr(v2), n, _lib.ptr(hd),
                                            _lib.stream_ptr()))
            torch.cuda.synchronize()
        d_ref = (p_ref - p0.double()).numpy()
        # fp32 state and arithmetic against an fp64 restatement: a few 1e-6 of the displacement
        assert rel(p[:n].cpu().double().numpy() - p0.double().numpy(), d_ref) < 5e-5, (wd, gs)
        assert rel(p2[:n].cpu().double().numpy() - p0.double().numpy(), d_ref) < 5e-5, (wd, gs)


def test_backward_after_another_forward_is_loud(golden_dir):
    """The executor keeps the activations of the LAST forward only (the reference nn.Module keeps one
    autograd graph per output): a backward through an output whose activations were overwritten raises."""
    from models.darcy import conv_boundary_condition
    g = np.load(os.path.join(golden_dir, "densenet_small16.npz"))
    model, K, cfg = _model(g)
    model.train()
    out = model(K)
    model.eval()
    with torch.no_grad():
        model(K)            # a validation batch in between
    model.train()
    d, n = conv_boundary_condition(out)
    with pytest.raises(RuntimeError):
        (d + n).backward()
    out1 = model(K)
    out2 = model(K)         # a second training forward
    with pytest.raises(RuntimeError):
        out1.sum().backward()
    out2.sum().backward()   # the last one is fine
    torch.cuda.synchronize()


def test_arbitrary_output_gradient_mse(golden_dir):
    """SURVEY section 8(f) row 2 (train_codec_max_likelihood.py:201-213): F.mse_loss on labelled data
    drives the same network; the executor's backward takes any dL/d(output).  Checked against the
    fp64 oracle evaluated here (small config, seconds on CPU)."""
    import torch.nn.functional as F
    from oracle.pdes_oracle import param_names
    g = np.load(os.path.join(golden_dir, "densenet_fiveblk16.npz"))
    model, K, cfg = _model(g)
    plan = orc.densenet_plan(**cfg)
    tg = torch.Generator().manual_seed(5)
    target = torch.randn(K.shape[0], cfg["out_channels"], cfg["imsize"], cfg["imsize"], generator=tg)
    model.train()
    model.zero_grad()
    out = model(K)
    loss = F.mse_loss(out, target.cuda())
    loss.backward()
    torch.cuda.synchronize()
    sd64 = orc.to_dtype(orc.make_state(plan, int(g["seed"])), torch.float64)
    names = param_names(plan)
    for n in names:
        sd64[n].requires_grad_(True)
    o64 = orc.densenet_forward(plan, sd64, K.cpu().double(), training=True)
    l64 = F.mse_loss(o64, target.double())
    l64.backward()
    assert rel(out.detach().cpu().numpy(), o64.detach().numpy()) < 1e-4
    assert abs(float(loss) - float(l64)) <= 1e-4 * abs(float(l64))
    params = dict(model.named_parameters())
    ref = np.concatenate([sd64[n].grad.numpy().ravel() for n in names])
    got = np.concatenate([params[n].grad.detach().double().cpu().numpy().ravel() for n in names])
    assert rel(got, ref) < 5e-3, rel(got, ref)   # flip-tolerant aggregate (see test_train_step_matches_reference)


def test_batch_size_changes_rebind_executor(golden_dir):
    """The script alternates training batches (32) with larger evaluation batches (64,
    train_codec_mixed_residual.py:166-206): a larger batch re-creates the executor (workspace, side
    stream, tables) and the smaller one must keep working on it afterwards."""
    from models.darcy import conv_boundary_condition
    g = np.load(os.path.join(golden_dir, "densenet_small16.npz"))
    model, K, cfg = _model(g)
    B = K.shape[0]

    def train_out():
        model.train()
        model.zero_grad()
        out = model(K)
        d, n = conv_boundary_condition(out)
        (d + n).backward()
        return out.detach().clone(), model.flat_parameters()[1].clone()

    o1, g1 = train_out()
    model.eval()
    with torch.no_grad():
        big = model(torch.cat([K, K, K], 0))           # 3x the batch: new executor
        small = model(K[:1])                           # and a smaller one on the same executor
    assert big.shape[0] == 3 * B and small.shape[0] == 1
    assert rel(big[:B].cpu().numpy(), big[B:2 * B].cpu().numpy()) < 1e-6
    assert rel(small.cpu().numpy(), big[:1].cpu().numpy()) < 1e-6
    o2, g2 = train_out()
    assert rel(o2.cpu().numpy(), o1.cpu().numpy()) < 1e-6
    assert rel(g2.cpu().numpy(), g1.cpu().numpy()) < 1e-5


@pytest.mark.parametrize("name", ["densenet_full32", "densenet_full32_b32", "densenet_bottleneck32b"])
def test_tensor_core_backward_strict_per_tensor(golden_dir, name):
    """Strict per-tensor check of the tensor-core dgrad / wgrad kernels with ReLU-mask flips excluded by
    construction: conv_impl 4 / 5 keep the exact-fp32 CUDA-core FORWARD (bitwise the forward, hence the
    masks, of conv_impl 6) and put only the dgrad / only the wgrad on tcgen05.  Every one of the 82
    gradient tensors must then agree with the all-CUDA-core run (6) to 2e-4 (two-piece fp16 operands:
    ~1e-6 per product; a saturated dY piece or a wrong layer would show as O(1))."""
    from models.darcy import conv_boundary_condition, conv_constitutive_constraint, conv_continuity_constraint
    from utils.image_gradient import SobelFilter
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    grads = {}
    outs = {}
    for impl in (6, 4, 5, 7):
        model, K, cfg = _model(g)
        model.conv_impl = impl
        sob = SobelFilter(cfg["imsize"], correct=True, device="cuda")
        model.train()
        model.zero_grad()
        out = model(K)
        loss = (conv_constitutive_constraint(K, out, sob) + conv_continuity_constraint(out, sob))
        d, n = conv_boundary_condition(out)
        (loss + 10.0 * (d + n)).backward()
        torch.cuda.synchronize()
        grads[impl] = {k: p.grad.detach().double().cpu().numpy().copy() for k, p in model.named_parameters()}
        outs[impl] = out.detach().cpu().numpy().copy()
    assert np.array_equal(outs[4], outs[6]) and np.array_equal(outs[5], outs[6])   # the same forward, bit for bit
    assert np.array_equal(outs[7], outs[6])
    worst = {}
    for impl in (4, 5, 7):   # 7: dgrad AND wgrad on tensor cores = the fused thin-layer dgrad (conv_dense_bwd.cu)
        errs = {k: rel(grads[impl][k], grads[6][k]) for k in grads[6]}
        worst[impl] = max(errs.items(), key=lambda kv: kv[1])
        assert worst[impl][1] < 2e-4, (impl, worst[impl])


@pytest.mark.parametrize("dtype", ["bf16", "fp16"])
@pytest.mark.parametrize("name", ["densenet_fiveblk16", "densenet_full64_channel"])
def test_one_piece_precision_modes(golden_dir, name, dtype, monkeypatch):
    """BASELINE config 3 ("bf16 tensor-core conv path"; PDES_CONV_DTYPE = bf16 | fp16): every convolution but the
    first runs ONE tensor-core product on one-piece operands.  Parity gate of SURVEY.md section 8d: the training
    forward must match a CPU emulation with rounded conv operands (oracle.densenet_forward(operand_round=...))
    far more closely than either matches fp32, the loss likewise, and the gradients must stay within the
    format's error of the fp32 reference."""
    from models.darcy import conv_boundary_condition, conv_constitutive_constraint, conv_continuity_constraint
    from utils.image_gradient import SobelFilter
    monkeypatch.setenv("PDES_CONV_DTYPE", dtype)
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    model, K, cfg = _model(g)
    plan = orc.densenet_plan(**cfg)
    sob = SobelFilter(cfg["imsize"], correct=True, device="cuda")
    model.train()
    model.zero_grad()
    out = model(K)
    loss = conv_constitutive_constraint(K, out, sob) + conv_continuity_constraint(out, sob)
    d, n = conv_boundary_condition(out)
    loss = loss + 10.0 * (d + n)
    loss.backward()
    torch.cuda.synchronize()
    sd = orc.make_state(plan, int(g["seed"]))
    with torch.no_grad():
        emu = orc.densenet_forward(plan, sd, K.cpu(), training=True, operand_round=dtype)
    e_emu = rel(out.detach().cpu().numpy(), emu.numpy())
    e_f32 = rel(out.detach().cpu().numpy(), g["out64"])
    # the format's own error against fp32: 3e-2 (bf16) on the full network (SURVEY.md headline facts), up to
    # 9e-2 / 1.1e-2 (bf16 / fp16) on the small five-block test network (CPU emulation)
    bar = 0.2 if dtype == "bf16" else 0.03
    assert e_f32 < bar, e_f32
    # and it IS the rounded-operand computation: far closer to the emulation than to fp32 on the small network;
    # on the 28-convolution network values that straddle a rounding boundary (GPU fmaf vs the CPU's two-step
    # BatchNorm) decorrelate the two bf16 computations to about half the format error
    assert e_emu < max(0.25 * e_f32 + 2e-3, 0.6 * e_f32), (e_emu, e_f32)
    loss_emu = float(orc.total_loss(K.cpu(), emu)[0])
    assert abs(float(loss) - loss_emu) <= 2e-2 * abs(loss_emu), (float(loss), loss_emu)
    names = [str(s) for s in g["param_names"]]
    params = dict(model.named_parameters())
    num = den = 0.0
    for i, nme in enumerate(names):
        gn = float(params[nme].grad.double().norm())
        num += (gn - float(g["grad_norm64"][i])) ** 2
        den += float(g["grad_norm64"][i]) ** 2
    assert np.sqrt(num / den) < (0.5 if dtype == "bf16" else 0.1), np.sqrt(num / den)
    assert all(bool(torch.isfinite(p.grad).all()) for p in model.parameters())


@pytest.mark.parametrize("tag", ["a", "b"])
def test_decoder_solver_step_matches_reference(golden_dir, tag):
    """`Decoder` (models/codec.py:321-370) at batch 1 driven by the solver's
    closure (solve_conv_mixed_residual.py:131-145) with the linear and the NONLINEAR constitutive law
    (models/darcy.py:179-191) against the reference's fp64 run; then torch.optim.LBFGS (the solver's
    optimiser, line 124) over the executor's parameters."""
    from models.codec import Decoder
    from models.darcy import (conv_boundary_condition, conv_constitutive_constraint,
                              conv_constitutive_constraint_nonlinear, conv_continuity_constraint)
    from utils.image_gradient import SobelFilter
    g = np.load(os.path.join(golden_dir, "decoder_solver.npz"))
    nz, hz, gr, feat, seed, imsize = [int(v) for v in g[f"{tag}_cfg"]]
    blocks = [int(b) for b in g[f"{tag}_blocks"]]
    plan = orc.decoder_plan(nz, 3, blocks, gr, feat)
    a1, a2 = [float(v) for v in g[f"{tag}_alphas"]]
    z = torch.tensor(g[f"{tag}_z"]).float().cuda()
    K = orc.make_input(1, imsize, seed).cuda()
    sob = SobelFilter(imsize, correct=True, device="cuda")
    for law in ("lin", "nl"):
        model = Decoder(nz, 3, blocks, growth_rate=gr, init_features=feat)
        sd = orc.make_state(plan, seed)
        assert list(model.state_dict().keys()) == list(sd.keys())
        assert tuple(model.model_size) == tuple(int(v) for v in g[f"{tag}_model_size"])
        model.load_state_dict(sd)
        model = model.cuda()
        model.train()
        model.zero_grad()
        out = model(z)
        assert tuple(out.shape) == (1, 3, imsize, imsize)
        out.retain_grad()
        if law == "nl":
            e = conv_constitutive_constraint_nonlinear(K, out, sob, a1, a2)
        else:
            e = conv_constitutive_constraint(K, out, sob)
        c = conv_continuity_constraint(out, sob)
        l_dir, l_neu = conv_boundary_condition(out)
        loss = e + c + (l_dir + l_neu) * 10.0
        loss.backward()
        torch.cuda.synchronize()
        assert rel(out.detach().cpu().numpy(), g[f"{tag}_{law}_out64"]) < 1e-4
        l4 = torch.stack([e, c, l_dir, l_neu]).detach().cpu().numpy()
        assert np.all(np.abs(l4 - g[f"{tag}_{law}_l4_64"]) <= 1e-4 * np.abs(g[f"{tag}_{law}_l4_64"])), (l4, g[f"{tag}_{law}_l4_64"])
        assert rel(out.grad.cpu().numpy(), g[f"{tag}_{law}_dout64"]) < 1e-4
        norms = np.array([float(p.grad.double().norm()) for p in model.parameters()])
        ref = g[f"{tag}_{law}_grad_norm64"]
        assert np.sqrt(((norms - ref) ** 2).sum() / (ref ** 2).sum()) < 2e-2
    # L-BFGS over the executor (closure = several forward/backward pairs per step)
    opt = torch.optim.LBFGS(model.parameters(), lr=0.5, max_iter=5, history_size=50)
    hist = []

    def closure():
        opt.zero_grad()
        o = model(z)
        en = conv_constitutive_constraint_nonlinear(K, o, sob, a1, a2) + conv_continuity_constraint(o, sob)
        d_, n_ = conv_boundary_condition(o)
        ls = en + (d_ + n_) * 10.0
        ls.backward()
        hist.append(float(ls))
        return ls

    for _ in range(3):
        opt.step(closure)
    assert np.all(np.isfinite(hist)) and hist[-1] < 0.5 * hist[0], hist


def test_dropout_step_matches_reference(golden_dir):
    """DenseED(drop_rate=0.2) (train_codec_mixed_residual.py --drop-rate): nn.Dropout2d behind the convolutions
    (models/codec.py:70-71, 110-149, 171-172) as a per-(sample, channel) mask applied by the executor.  The masks
    are drawn like torch.feature_dropout draws them; drawn on the CPU generator with the fixture's seed they are
    the reference's masks, and the whole training step must match the reference's fp32 run."""
    from models.codec import DenseED
    from models.darcy import conv_boundary_condition, conv_constitutive_constraint, conv_continuity_constraint
    from utils.image_gradient import SobelFilter
    g = np.load(os.path.join(golden_dir, "densenet_dropout16.npz"))
    cfg = dict(in_channels=1, out_channels=3, imsize=int(g["cfg_imsize"]), blocks=[int(b) for b in g["cfg_blocks"]],
               growth_rate=int(g["cfg_growth_rate"]), init_features=int(g["cfg_init_features"]))
    plan = orc.densenet_plan(**cfg)
    for impl in (0, 1):
        model = DenseED(1, 3, cfg["imsize"], cfg["blocks"], growth_rate=cfg["growth_rate"],
                        init_features=cfg["init_features"], drop_rate=float(g["drop_rate"]))
        model.load_state_dict(orc.make_state(plan, int(g["seed"])))
        model = model.cuda()
        model.conv_impl = impl
        model._mask_device = "cpu"       # draw the masks on the CPU generator, like the CPU reference did
        K = orc.make_input(int(g["B"]), cfg["imsize"], int(g["seed"])).cuda()
        sob = SobelFilter(cfg["imsize"], correct=True, device="cuda")
        model.train()
        model.zero_grad()
        torch.manual_seed(int(g["mask_seed"]))
        out = model(K)
        out.retain_grad()
        l_c = conv_constitutive_constraint(K, out, sob)
        l_d = conv_continuity_constraint(out, sob)
        l_dir, l_neu = conv_boundary_condition(out)
        loss = (l_c + l_d) + (l_dir + l_neu) * 10.0
        loss.backward()
        torch.cuda.synchronize()
        assert rel(out.detach().cpu().numpy(), g["out"]) < 1e-4, (impl, rel(out.detach().cpu().numpy(), g["out"]))
        l4 = torch.stack([l_c, l_d, l_dir, l_neu]).detach().cpu().numpy()
        assert np.all(np.abs(l4 - g["l4"]) <= 1e-4 * np.abs(g["l4"]))
        assert rel(out.grad.cpu().numpy(), g["dout"]) < 1e-4
        names = [str(s) for s in g["param_names"]]
        params = dict(model.named_parameters())
        flat = np.concatenate([params[n].grad.detach().cpu().numpy().ravel() for n in names])
        assert rel(flat, g["grads"]) < 5e-3, (impl, rel(flat, g["grads"]))
        model.eval()
        with torch.no_grad():
            ev = model(K)
        assert rel(ev.cpu().numpy(), g["out_eval"]) < 1e-4
        # device-drawn masks (the default): whole channels are dropped and the survivors scaled by 1/(1-p)
        model._mask_device = None
        model.train()
        o2 = model(K)
        assert bool(torch.isfinite(o2).all()) and rel(o2.detach().cpu().numpy(), g["out"]) > 1e-3


@pytest.mark.parametrize("act", ["tanh", "softplus", "sigmoid", "lrelu"])
def test_out_activation(golden_dir, act):
    """out_activation (models/codec.py:190-204, 288-289): the named module sits behind the last convolution in
    `features`; forward and the parameter gradients equal the chain rule through the plain network."""
    from models.codec import DenseED
    g = np.load(os.path.join(golden_dir, "densenet_small16.npz"))
    base, K, cfg = _model(g)
    model = DenseED(cfg["in_channels"], cfg["out_channels"], cfg["imsize"], cfg["blocks"], growth_rate=cfg["growth_rate"],
                    init_features=cfg["init_features"], out_activation=act).to("cuda")
    model.load_state_dict(base.state_dict())
    assert list(model.features._modules)[-1] == act
    ref_act = {"tanh": torch.tanh, "sigmoid": torch.sigmoid, "lrelu": torch.nn.functional.leaky_relu,
               "softplus": lambda t: torch.nn.functional.softplus(t, beta=4)}[act]
    base.train(), model.train()
    base.zero_grad(), model.zero_grad()
    z = base(K)
    zl = z.detach().clone().requires_grad_(True)
    w = torch.linspace(-1.0, 1.0, z.numel(), device="cuda").view_as(z)
    (ref_act(zl) * w).sum().backward()
    z.backward(zl.grad)
    out = model(K)
    assert torch.allclose(out, ref_act(z.detach()), rtol=1e-6, atol=1e-7)
    (out * w).sum().backward()
    for (n, p), (_, q) in zip(model.named_parameters(), base.named_parameters()):
        # (two executions of the same backward: equal up to the order of the atomic accumulations)
        assert rel(p.grad.cpu().numpy(), q.grad.cpu().numpy()) < 1e-4, n


def test_fused_adam_optimizer_matches_torch(golden_dir):
    """pde_surrogate_b200.optim.Adam (what run_reference_script.py installs as torch.optim.Adam): eight training
    steps with OneCycle-style lr / beta changes and weight decay give the parameters torch's own Adam gives; the
    optimizer state keeps torch's layout, survives state_dict round trips and a switch to the stock step."""
    from models.darcy import conv_boundary_condition
    from pde_surrogate_b200 import optim as pdes_optim
    g = np.load(os.path.join(golden_dir, "densenet_small16.npz"))
    ma, K, cfg = _model(g)
    mb, _, _ = _model(g)
    stock = pdes_optim._StockAdam
    assert not getattr(stock, "_pdes_fused", False)
    oa = pdes_optim.Adam(ma.parameters(), lr=1e-3, weight_decay=1e-3)
    ob = stock(mb.parameters(), lr=1e-3, weight_decay=1e-3)

    def one(model, opt, i):
        for grp in opt.param_groups:
            grp["lr"] = 1e-3 * (1.0 + 0.3 * i)
            grp["betas"] = (0.95 - 0.01 * i, 0.999)
        model.train()
        model.zero_grad()
        out = model(K)
        d, n = conv_boundary_condition(out)
        ((out ** 2).mean() + d + n).backward()
        opt.step()

    for i in range(4):
        one(ma, oa, i), one(mb, ob, i)
    assert oa.fused_steps == 4
    sd = oa.state_dict()
    assert len(sd["state"]) == len(list(ma.parameters())) and float(sd["state"][0]["step"]) == 4.0
    oa.load_state_dict(sd)   # fresh state tensors: re-bound (and copied into the flat moments) at the next step
    for i in range(4, 6):
        one(ma, oa, i), one(mb, ob, i)
    assert oa.fused_steps == 6
    os.environ["PDES_FUSED_ADAM"] = "0"   # the stock step takes over mid-run on the same state
    try:
        for i in range(6, 8):
            one(ma, oa, i), one(mb, ob, i)
    finally:
        os.environ.pop("PDES_FUSED_ADAM")
    assert oa.fused_steps == 6
    for (n, p), (_, q) in zip(ma.named_parameters(), mb.named_parameters()):
        assert rel(p.detach().cpu().numpy(), q.detach().cpu().numpy()) < 2e-5, n
    st = oa.state[next(iter(ma.parameters()))]
    assert float(st["step"]) == 8.0
