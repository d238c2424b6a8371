"""Generate golden fixtures by running the UNMODIFIED reference (/root/reference) on CPU.

Run in the build container only (the reference does not travel to the GPU box):

    python tests/golden/make_golden.py

The reference imports matplotlib at module top (models/codec.py:10-11, models/darcy.py:9-10);
matplotlib is not installed here, so an empty stub package is put on sys.path for the import.
Weights / inputs come from oracle.pdes_oracle.make_state / make_input (numpy MT19937 streams,
machine independent), so fixtures store only inputs' seeds and the reference's OUTPUTS.
"""
import os
import sys
import tempfile
from collections import OrderedDict

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("PDES_REFERENCE", "/root/reference")


def import_reference():
    stub = tempfile.mkdtemp(prefix="mplstub_")
    os.makedirs(os.path.join(stub, "matplotlib"))
    open(os.path.join(stub, "matplotlib", "__init__.py"), "w").close()
    with open(os.path.join(stub, "matplotlib", "pyplot.py"), "w") as f:
        f.write("def switch_backend(*a, **k):\n    return None\n")
    sys.path.insert(0, stub)
    sys.path.insert(0, REF)
    for m in [k for k in sys.modules if k.split(".")[0] in ("models", "utils")]:
        del sys.modules[m]
    from models.codec import DenseED
    from models import darcy
    from utils.image_gradient import SobelFilter
    import models.codec as _codec
    import models.glow_msc  # noqa: F401  (cGlow coupling layers; needs scipy)
    sys.path.remove(REF)
    import_reference.Decoder = _codec.Decoder
    return DenseED, darcy, SobelFilter


def main():
    torch.set_num_threads(1)  # bit-stable fixtures
    sys.path.insert(0, ROOT)
    from oracle import pdes_oracle as orc

    DenseED, darcy, SobelFilter = import_reference()

    def ref_step(cfg, B, seed, dtype, kind="lognormal", upsample="nearest", bn_size=0):
        # bn_size > 0: DenseED(bottleneck=True, bn_size=bn_size) (models/codec.py:56-64)
        plan = orc.densenet_plan(**cfg, upsample=upsample, bottleneck=bn_size)
        sd = orc.to_dtype(orc.make_state(plan, seed), dtype)
        K = orc.make_input(B, cfg["imsize"], seed, kind=kind).to(dtype)
        model = DenseED(cfg["in_channels"], cfg["out_channels"], cfg["imsize"], cfg["blocks"],
                        growth_rate=cfg["growth_rate"], init_features=cfg["init_features"], upsample=upsample,
                        **(dict(bottleneck=True, bn_size=bn_size) if bn_size else {}))
        model = model.to(dtype)
        missing = model.load_state_dict(sd, strict=True)
        assert not missing.missing_keys and not missing.unexpected_keys
        assert [k for k in model.state_dict().keys()] == list(sd.keys()), "state_dict key order"
        # upsample=None: the reference's last decoding does not upsample (codec.py:176-179), the output is
        # imsize/2 wide; the residual loss of these fixtures is taken on the 2x subsampled permeability
        osz = cfg["imsize"] if upsample is not None else cfg["imsize"] // 2
        Kl = K if upsample is not None else K[:, :, ::2, ::2].contiguous()
        sob = SobelFilter(osz, correct=True, device="cpu")
        if dtype == torch.float64:
            for a in ("HSOBEL_WEIGHTS_3x3", "VSOBEL_WEIGHTS_3x3", "modifier"):
                setattr(sob, a, getattr(sob, a).double())
        # eval-mode forward first (running stats as given)
        model.eval()
        with torch.no_grad():
            out_eval = model(K).clone()
        # one training step body: train_codec_mixed_residual.py:226-233
        model.train()
        model.zero_grad()
        out = model(K)
        assert out.shape[-1] == osz
        out.retain_grad()
        l_c = darcy.conv_constitutive_constraint(Kl, out, sob)
        l_d = darcy.conv_continuity_constraint(out, sob)
        l_dir, l_neu = darcy.conv_boundary_condition(out)
        loss = (l_c + l_d) + (l_dir + l_neu) * 10.0
        loss.backward()
        l_d_notb = darcy.conv_continuity_constraint(out.detach(), sob, use_tb=False)
        grads = OrderedDict((n, p.grad.detach().clone()) for n, p in model.named_parameters())
        running = OrderedDict((n, b.detach().clone()) for n, b in model.named_buffers())
        return dict(K=K, out=out.detach(), out_eval=out_eval, l4=torch.stack([l_c, l_d, l_dir, l_neu]).detach(),
                    loss=loss.detach(), dout=out.grad.detach().clone(), grads=grads, running=running,
                    l_d_notb=l_d_notb.detach(), names=[n for n, _ in model.named_parameters()],
                    model_size=model.model_size)

    def save_case(fname, cfg, B, seed, full_grads, kind="lognormal", compact=False, upsample="nearest", bn_size=0):
        r32 = ref_step(cfg, B, seed, torch.float32, kind, upsample, bn_size)
        r64 = ref_step(cfg, B, seed, torch.float64, kind, upsample, bn_size)
        d = dict(input_kind=kind, cfg_in_channels=cfg["in_channels"], cfg_out_channels=cfg["out_channels"],
                 cfg_imsize=cfg["imsize"], cfg_blocks=np.array(cfg["blocks"]),
                 cfg_growth_rate=cfg["growth_rate"], cfg_init_features=cfg["init_features"], B=B,
                 seed=seed, bn_size=bn_size, model_size=np.array(r32["model_size"]),
                 out=r32["out"].numpy(), out_eval=r32["out_eval"].numpy(), l4=r32["l4"].numpy(),
                 loss=r32["loss"].numpy(), dout=r32["dout"].numpy(), l_d_notb=r32["l_d_notb"].numpy(),
                 out64=r64["out"].numpy().astype(np.float64), l4_64=r64["l4"].numpy(),
                 loss64=r64["loss"].numpy(), dout64=r64["dout"].numpy(),
                 out_eval64=r64["out_eval"].numpy())
        if compact:
            # large batches: the fp64 fields are stored rounded to fp32 (6e-8 relative, far below the
            # 1e-4 bar) and the reference's own fp32 fields are dropped (their scalars are kept)
            d["compact"] = 1
            for k in ("out", "out_eval", "dout"):
                del d[k]
            for k in ("out64", "dout64", "out_eval64"):
                d[k] = d[k].astype(np.float32)
            d["out_err32"] = float((r32["out"].double() - r64["out"]).norm() / r64["out"].norm())
            d["dout_err32"] = float((r32["dout"].double() - r64["dout"]).norm() / r64["dout"].norm())
        names = r32["names"]
        d["param_names"] = np.array(names)
        d["grad_norm32"] = np.array([float(r32["grads"][n].double().norm()) for n in names])
        d["grad_norm64"] = np.array([float(r64["grads"][n].norm()) for n in names])
        # fp32-vs-fp64 error of the reference itself = the gradient noise floor
        d["grad_err32"] = np.array([float((r32["grads"][n].double() - r64["grads"][n]).norm())
                                    for n in names])
        if full_grads:
            d["grads32"] = np.concatenate([r32["grads"][n].numpy().ravel() for n in names])
            d["grads64"] = np.concatenate([r64["grads"][n].numpy().ravel() for n in names])
        else:
            d["grads64_head"] = np.concatenate([r64["grads"][n].numpy().ravel()[:16] for n in names])
            d["grads64_head_len"] = np.array([min(16, r64["grads"][n].numel()) for n in names])
        bn_names = [n for n in r32["running"] if n.endswith(("running_mean", "running_var"))]
        d["running_names"] = np.array(bn_names)
        d["running64"] = np.concatenate([r64["running"][n].numpy().ravel() for n in bn_names])
        nbt = [int(v) for n, v in r32["running"].items() if n.endswith("num_batches_tracked")]
        d["num_batches_tracked"] = np.array(nbt)
        np.savez_compressed(os.path.join(HERE, fname), **d)
        print(fname, "loss", float(r32["loss"]), "l4", r32["l4"].tolist(), "model_size", r32["model_size"])

    full = dict(in_channels=1, out_channels=3, blocks=[6, 8, 6], growth_rate=16, init_features=48)

    def channel_case():
        # BASELINE config 3's data: two-valued channelized permeability, 64x64, a batch of 4
        save_case("densenet_full64_channel.npz", dict(full, imsize=64), B=4, seed=13, full_grads=False,
                  kind="channel")

    def timed_shape_cases():
        # the shapes bench.py times (BASELINE configs 1/2): batch 32 at 64x64 and at 32x32 - every
        # persistent CTA of the tensor-core kernels walks several tiles here
        save_case("densenet_full64_b32.npz", dict(full, imsize=64), B=32, seed=17, full_grads=False, compact=True)
        save_case("densenet_full32_b32.npz", dict(full, imsize=32), B=32, seed=19, full_grads=False, compact=True)

    def trajectory_case():
        # Three optimisation steps of the reference itself (its DenseED + losses + torch.optim.Adam, the
        # loop body of train_codec_mixed_residual.py:226-240 with a per-step learning rate) at batch 32, in
        # fp32 and in fp64: the losses, and a strided sample of the final parameters.  The fp32-vs-fp64
        # gap IS the reference's own trajectory noise (chaotic: ReLU flips + sign-like first Adam steps).
        d = {}
        lrs = [5e-4, 7e-4, 1e-3]
        for imsize in (32, 64):
            cfg = dict(full, imsize=imsize)
            plan = orc.densenet_plan(**cfg)
            res = {}
            for dtype in (torch.float32, torch.float64):
                sd = orc.to_dtype(orc.make_state(plan, 1), dtype)
                model = DenseED(1, 3, imsize, cfg["blocks"]).to(dtype)
                model.load_state_dict(sd, strict=True)
                sob = SobelFilter(imsize, correct=True, device="cpu")
                if dtype == torch.float64:
                    for a in ("HSOBEL_WEIGHTS_3x3", "VSOBEL_WEIGHTS_3x3", "modifier"):
                        setattr(sob, a, getattr(sob, a).double())
                opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=0.0)
                model.train()
                losses = []
                for i, lr in enumerate(lrs):
                    K = orc.make_input(32, imsize, 100 + i).to(dtype)
                    model.zero_grad()
                    out = model(K)
                    loss = (darcy.conv_constitutive_constraint(K, out, sob) + darcy.conv_continuity_constraint(out, sob))
                    l_dir, l_neu = darcy.conv_boundary_condition(out)
                    loss = loss + (l_dir + l_neu) * 10.0
                    loss.backward()
                    for grp in opt.param_groups:
                        grp["lr"] = lr
                    opt.step()
                    losses.append(float(loss))
                flat = np.concatenate([p.detach().double().numpy().ravel() for p in model.parameters()])
                res[dtype] = (np.array(losses), flat)
            d["loss32_%d" % imsize], d["loss64_%d" % imsize] = res[torch.float32][0], res[torch.float64][0]
            d["params32_%d" % imsize] = res[torch.float32][1][::97].copy()
            d["params64_%d" % imsize] = res[torch.float64][1][::97].copy()
            print("trajectory", imsize, res[torch.float32][0], res[torch.float64][0])
        d["lrs"] = np.array(lrs)
        d["stride"] = 97
        np.savez_compressed(os.path.join(HERE, "trajectory_b32.npz"), **d)

    def decoder_cases():
        # `Decoder` (models/codec.py:321-370) driven by the solver's closure
        # (solve_conv_mixed_residual.py:131-145) with the linear and the nonlinear constitutive law
        # (models/darcy.py:179-191), batch 1, in fp64 and fp32.
        Decoder = import_reference.Decoder
        d = {}
        for tag, (nz, hz, blocks, gr, feat, seed) in dict(a=(2, 8, [3, 2], 8, 16, 23), b=(1, 16, [8, 6], 16, 48, 29)).items():
            plan = orc.decoder_plan(nz, 3, blocks, gr, feat)
            imsize = hz * 2 ** len(blocks)
            rs = np.random.RandomState(500 + seed)
            z = torch.tensor(0.5 * rs.standard_normal((1, nz, hz, hz)))
            K = orc.make_input(1, imsize, seed).double()
            for dtype, sfx in ((torch.float64, "64"), (torch.float32, "32")):
                sd = orc.to_dtype(orc.make_state(plan, seed), dtype)
                model = Decoder(nz, 3, blocks, growth_rate=gr, init_features=feat).to(dtype)
                res = model.load_state_dict(sd, strict=True)
                assert not res.missing_keys and not res.unexpected_keys
                assert list(model.state_dict().keys()) == list(sd.keys())
                sob = SobelFilter(imsize, correct=True, device="cpu")
                if dtype == torch.float64:
                    for a in ("HSOBEL_WEIGHTS_3x3", "VSOBEL_WEIGHTS_3x3", "modifier"):
                        setattr(sob, a, getattr(sob, a).double())
                model.train()
                for law, (a1, a2) in dict(lin=(0.0, 0.0), nl=(1.0, 0.7)).items():
                    model.zero_grad()
                    for m_ in model.modules():   # same BatchNorm running statistics before every pass
                        if isinstance(m_, torch.nn.BatchNorm2d):
                            m_.reset_running_stats()
                    out = model(z.to(dtype))
                    out.retain_grad()
                    Kd = K.to(dtype)
                    if law == "nl":
                        e = darcy.conv_constitutive_constraint_nonlinear(Kd, out, sob, a1, a2)
                    else:
                        e = darcy.conv_constitutive_constraint(Kd, out, sob)
                    c = darcy.conv_continuity_constraint(out, sob)
                    l_dir, l_neu = darcy.conv_boundary_condition(out)
                    loss = e + c + (l_dir + l_neu) * 10.0
                    loss.backward()
                    names = [n for n, _ in model.named_parameters()]
                    d[f"{tag}_{law}_l4_{sfx}"] = torch.stack([e, c, l_dir, l_neu]).detach().numpy()
                    d[f"{tag}_{law}_loss_{sfx}"] = loss.detach().numpy()
                    if sfx == "64":
                        d[f"{tag}_{law}_out64"] = out.detach().numpy()
                        d[f"{tag}_{law}_dout64"] = out.grad.detach().numpy()
                        d[f"{tag}_{law}_grad_norm64"] = np.array([float(p.grad.norm()) for _, p in model.named_parameters()])
                        d[f"{tag}_{law}_grads64_head"] = np.concatenate([p.grad.numpy().ravel()[:16] for _, p in model.named_parameters()])
                d[f"{tag}_param_names"] = np.array(names)
                d[f"{tag}_model_size"] = np.array(model.model_size)
            d[f"{tag}_cfg"] = np.array([nz, hz, gr, feat, seed, imsize])
            d[f"{tag}_blocks"] = np.array(blocks)
            d[f"{tag}_z"] = z.numpy()
            d[f"{tag}_alphas"] = np.array([1.0, 0.7])
            print("decoder", tag, "loss lin/nl", float(d[f"{tag}_lin_loss_64"]), float(d[f"{tag}_nl_loss_64"]))
        np.savez_compressed(os.path.join(HERE, "decoder_solver.npz"), **d)

    def bilinear_cases():
        # --upsample bilinear (train_codec_mixed_residual.py:47; models/codec.py:33-40, 143-146, 177-178)
        small5 = dict(in_channels=1, out_channels=3, imsize=16, blocks=[2, 1, 2, 1, 2], growth_rate=8, init_features=16)
        save_case("densenet_bilinear16.npz", small5, B=2, seed=41, full_grads=True, upsample="bilinear")
        save_case("densenet_bilinear32.npz", dict(full, imsize=32), B=3, seed=43, full_grads=False, upsample="bilinear")

    def convt_cases():
        # upsample=None: nn.ConvTranspose2d transitions (models/codec.py:139-142), no upsampling in last_decoding
        small5 = dict(in_channels=1, out_channels=3, imsize=16, blocks=[2, 1, 2, 1, 2], growth_rate=8, init_features=16)
        save_case("densenet_convt16.npz", small5, B=2, seed=47, full_grads=True, upsample=None)
        save_case("densenet_convt32.npz", dict(full, imsize=32, blocks=[3, 4, 3, 4, 3]), B=3, seed=53, full_grads=False,
                  upsample=None)

    def bottleneck_cases():
        # DenseED(bottleneck=True): dense layers wider than bn_size * growth take the 1x1 -> 3x3 form (codec.py:56-64)
        small = dict(in_channels=1, out_channels=3, imsize=16, blocks=[2, 3, 2], growth_rate=4, init_features=8)
        save_case("densenet_bottleneck16.npz", small, B=3, seed=59, full_grads=True, bn_size=2)
        save_case("densenet_bottleneck32.npz", dict(full, imsize=32), B=3, seed=61, full_grads=False, bn_size=4)
        save_case("densenet_bottleneck32b.npz", dict(full, imsize=32, blocks=[3, 4, 3]), B=2, seed=67, full_grads=False,
                  bn_size=4)

    def coupling_cases():
        # SURVEY.md section 8(f) row 1 / BASELINE config 5: the cGlow coupling network `_DenseCoupling`
        # (models/glow_msc.py:276-294, Conv2dZeros 240-255) and `AffineCouplingLayer` forward / reverse (297-344),
        # imported from the untouched reference (these classes do not touch the in-place clamp of line 438),
        # with gradients w.r.t. parameters AND inputs, in fp64 and fp32.
        from models import glow_msc as ref_glow   # resolved from /root/reference (sys.modules entry of the import above)
        d = {}
        # (tag, in_features, cond_features, H, B, seed): Appendix B shapes 82->16..130->2 @32^2 and 158->12 @16^2
        for tag, (fin, fcond, H, B, seed) in dict(a=(3, 80, 32, 2, 51), b=(12, 104, 16, 3, 53), c=(6, 9, 8, 2, 55)).items():
            for dtype, sfx in ((torch.float64, "64"), (torch.float32, "32")):
                layer = ref_glow.AffineCouplingLayer(fin, fcond, coupling_net="dense").to(dtype)
                net = layer.coupling_nn
                cin = (fin // 2 + (fin % 2)) + fcond
                cout = fin if fin % 2 == 0 else fin - 1
                plan = orc.coupling_plan(cin, cout)
                sd = orc.to_dtype(orc.make_state(plan, seed), dtype)
                res = net.load_state_dict(sd, strict=True)
                assert not res.missing_keys and not res.unexpected_keys
                assert list(net.state_dict().keys()) == list(sd.keys()), "coupling state_dict order"
                rs = np.random.RandomState(900 + seed)
                x = torch.tensor(rs.standard_normal((B, fin, H, H)), dtype=dtype, requires_grad=True)
                cond = torch.tensor(rs.standard_normal((B, fcond, H, H)), dtype=dtype, requires_grad=True)
                wy = torch.tensor(rs.standard_normal((B, fin, H, H)), dtype=dtype)
                layer.train()
                for mode in ("fwd", "rev"):
                    for m_ in layer.modules():
                        if isinstance(m_, torch.nn.BatchNorm2d):
                            m_.reset_running_stats()
                    layer.zero_grad()
                    x.grad = cond.grad = None
                    y, logdet = (layer.forward if mode == "fwd" else layer.reverse)(x, cond)
                    obj = (y * wy).sum() + 0.3 * logdet.sum()
                    obj.backward()
                    names = [n for n, _ in net.named_parameters()]
                    if sfx == "64":
                        d[f"{tag}_{mode}_y"], d[f"{tag}_{mode}_logdet"] = y.detach().numpy(), logdet.detach().numpy()
                        d[f"{tag}_{mode}_dx"], d[f"{tag}_{mode}_dcond"] = x.grad.numpy().copy(), cond.grad.numpy().copy()
                        d[f"{tag}_{mode}_grads"] = np.concatenate([p.grad.numpy().ravel() for _, p in net.named_parameters()])
                        d[f"{tag}_{mode}_grad_norm"] = np.array([float(p.grad.norm()) for _, p in net.named_parameters()])
                    else:
                        g64 = d[f"{tag}_{mode}_grads"]
                        g32 = np.concatenate([p.grad.double().numpy().ravel() for _, p in net.named_parameters()])
                        d[f"{tag}_{mode}_grad_err32"] = float(np.linalg.norm(g32 - g64) / np.linalg.norm(g64))
                        d[f"{tag}_{mode}_y_err32"] = float((y.detach().double() - torch.tensor(d[f"{tag}_{mode}_y"])).norm() /
                                                          torch.tensor(d[f"{tag}_{mode}_y"]).norm())
                d[f"{tag}_param_names"] = np.array(names)
            d[f"{tag}_cfg"] = np.array([fin, fcond, H, B, seed, cin, cout])
            print("coupling", tag, "cin", cin, "cout", cout, "grad_err32", d[f"{tag}_fwd_grad_err32"], d[f"{tag}_rev_grad_err32"])
        for k in list(d):   # fields rounded to fp32 for storage (6e-8 relative, far below the 1e-4 bars)
            if k.endswith(("_y", "_dx", "_dcond")):
                d[k] = d[k].astype(np.float32)
        np.savez_compressed(os.path.join(HERE, "cglow_coupling.npz"), **d)

    def cglow_model_case():
        # The whole MultiScaleCondGlow (models/glow_msc.py:672-828) through ONE reverse-KL step body
        # (train_cglow_reverse_kl.py:250-262): generate -> Darcy residuals -> entropy term -> backward, in fp64 and fp32.
        # The reference runs with ONE patch: GaussianDiag's in-place clamp (glow_msc.py:438), which current PyTorch
        # refuses to differentiate, is made out of place (same values and gradients).
        import math
        from models import glow_msc as ref_glow
        from models import darcy as ref_darcy

        def patched_init(self, mean, log_stddev):
            self.mean = mean
            self.log_stddev = log_stddev.clamp(min=-10., max=math.log(5.))
        ref_glow.GaussianDiag.__init__ = patched_init
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from test_glow_flow import randomise
        cfg = dict(img_size=16, x_channels=1, y_channels=3, enc_blocks=[2, 2, 2], flow_blocks=[2, 3, 2], LUdecompose=True)
        np.random.seed(21)
        torch.manual_seed(21)
        base = ref_glow.MultiScaleCondGlow(**cfg)
        sd = randomise(base, 23)
        B = 4
        x = torch.exp(0.3 * torch.randn(B, 1, 16, 16))
        eps = [0.7 * torch.randn(B, *s_) for s_ in base._z_shapes()]
        d = dict(x=x.numpy(), B=B)
        for i, e in enumerate(eps):
            d["eps%d" % i] = e.numpy()
        names = list(sd.keys())
        d["state_names"] = np.array(names)
        for i, k in enumerate(names):
            d["state%d" % i] = sd[k].numpy()
        for dtype, sfx in ((torch.float64, "64"), (torch.float32, "32")):
            np.random.seed(21)
            torch.manual_seed(21)
            model = ref_glow.MultiScaleCondGlow(**cfg)
            model.load_state_dict(sd)
            model = model.to(dtype)
            sob = SobelFilter(16, correct=True, device="cpu")
            if dtype == torch.float64:
                for a in ("HSOBEL_WEIGHTS_3x3", "VSOBEL_WEIGHTS_3x3", "modifier"):
                    setattr(sob, a, getattr(sob, a).double())
            model.train()
            model.zero_grad()
            xi = x.to(dtype)
            y, logp = model.generate(xi, eps_list=[e.to(dtype) for e in eps])
            res = ref_darcy.conv_constitutive_constraint(xi, y, sob) + ref_darcy.conv_continuity_constraint(y, sob)
            l_dir, l_neu = ref_darcy.conv_boundary_condition(y)
            neg_entropy = logp.mean() / math.log(2.) / (3 * 16 * 16)
            loss = (res + (l_dir + l_neu) * 50.0) * 150.0 + neg_entropy
            loss.backward()
            pn = [n for n, p in model.named_parameters() if p.grad is not None]
            g = np.concatenate([p.grad.double().numpy().ravel() for n, p in model.named_parameters() if p.grad is not None])
            if sfx == "64":
                d["y64"], d["logp64"], d["loss64"] = y.detach().numpy().astype(np.float32), logp.detach().numpy(), float(loss)
                d["grads64"] = g.astype(np.float32)
                d["grad_names"] = np.array(pn)
                d["grad_sizes"] = np.array([p.grad.numel() for n, p in model.named_parameters() if p.grad is not None])
                g64 = g
            else:
                d["y_err32"] = float((y.detach().double().numpy() - d["y64"]).ravel().__abs__().max())
                d["grad_err32"] = float(np.linalg.norm(g - g64) / np.linalg.norm(g64))
                d["loss32"] = float(loss)
        d["cfg_enc"], d["cfg_flow"] = np.array(cfg["enc_blocks"]), np.array(cfg["flow_blocks"])
        np.savez_compressed(os.path.join(HERE, "cglow_model.npz"), **d)
        print("cglow model: loss64", d["loss64"], "loss32", d["loss32"], "grad_err32", d["grad_err32"], "n state", len(names))

    def dropout_case():
        # DenseED(drop_rate=0.2) (train_codec_mixed_residual.py --drop-rate; nn.Dropout2d behind the convolutions,
        # models/codec.py:70-71, 110-149, 171-172): one fp32 training step of the reference with the CPU
        # generator seeded right before the forward pass - the masks are then a function of the seed alone.
        cfg = dict(in_channels=1, out_channels=3, imsize=16, blocks=[2, 1, 2, 1, 2], growth_rate=8, init_features=16)
        B, seed, p_drop, mask_seed = 4, 31, 0.2, 77
        plan = orc.densenet_plan(**cfg)
        sd = orc.make_state(plan, seed)
        K = orc.make_input(B, cfg["imsize"], seed)
        model = DenseED(1, 3, cfg["imsize"], cfg["blocks"], growth_rate=8, init_features=16, drop_rate=p_drop)
        model.load_state_dict(sd, strict=True)
        sob = SobelFilter(cfg["imsize"], correct=True, device="cpu")
        model.train()
        model.zero_grad()
        torch.manual_seed(mask_seed)
        out = model(K)
        out.retain_grad()
        l_c = darcy.conv_constitutive_constraint(K, out, sob)
        l_d = darcy.conv_continuity_constraint(out, sob)
        l_dir, l_neu = darcy.conv_boundary_condition(out)
        loss = (l_c + l_d) + (l_dir + l_neu) * 10.0
        loss.backward()
        model.eval()
        with torch.no_grad():
            out_eval = model(K)
        names = [n for n, _ in model.named_parameters()]
        d = dict(cfg_imsize=16, cfg_blocks=np.array(cfg["blocks"]), cfg_growth_rate=8, cfg_init_features=16, B=B,
                 seed=seed, drop_rate=p_drop, mask_seed=mask_seed, out=out.detach().numpy(),
                 out_eval=out_eval.numpy(), l4=torch.stack([l_c, l_d, l_dir, l_neu]).detach().numpy(),
                 loss=loss.detach().numpy(), dout=out.grad.numpy(), param_names=np.array(names),
                 grads=np.concatenate([p.grad.numpy().ravel() for _, p in model.named_parameters()]),
                 n_dropout_modules=sum(1 for m_ in model.modules() if isinstance(m_, torch.nn.Dropout2d)))
        np.savez_compressed(os.path.join(HERE, "densenet_dropout16.npz"), **d)
        print("dropout case: loss", float(loss), "dropout modules", d["n_dropout_modules"])

    if "--only-channel" in sys.argv:
        channel_case()
        return
    if "--only-dropout" in sys.argv:
        dropout_case()
        return
    if "--only-cglow-model" in sys.argv:
        cglow_model_case()
        return
    if "--only-coupling" in sys.argv:
        coupling_cases()
        return
    if "--only-bottleneck" in sys.argv:
        bottleneck_cases()
        return
    if "--only-convt" in sys.argv:
        convt_cases()
        return
    if "--only-bilinear" in sys.argv:
        bilinear_cases()
        return
    if "--only-decoder" in sys.argv:
        decoder_cases()
        return
    if "--only-trajectory" in sys.argv:
        trajectory_case()
        return
    if "--only-b32" in sys.argv:
        timed_shape_cases()
        return
    small = dict(in_channels=1, out_channels=3, imsize=16, blocks=[1, 2, 1], growth_rate=4,
                 init_features=8)
    save_case("densenet_small16.npz", small, B=3, seed=3, full_grads=True)
    small5 = dict(in_channels=1, out_channels=3, imsize=16, blocks=[2, 1, 2, 1, 2], growth_rate=8,
                  init_features=16)
    save_case("densenet_fiveblk16.npz", small5, B=2, seed=5, full_grads=True)
    full = dict(in_channels=1, out_channels=3, blocks=[6, 8, 6], growth_rate=16, init_features=48)
    save_case("densenet_full32.npz", dict(full, imsize=32), B=2, seed=7, full_grads=False)
    save_case("densenet_full64.npz", dict(full, imsize=64), B=2, seed=11, full_grads=False)
    channel_case()
    timed_shape_cases()
    trajectory_case()
    decoder_cases()
    dropout_case()
    bilinear_cases()
    convt_cases()
    bottleneck_cases()
    coupling_cases()
    cglow_model_case()

    # ---- Sobel operators and loss terms on their own, incl. odd size and autograd adjoint ----
    rs = np.random.RandomState(42)
    d = {}
    for tag, (n, H) in dict(a=(2, 16), b=(1, 65), c=(1, 32)).items():
        img = torch.tensor(rs.standard_normal((n, 1, H, H)), dtype=torch.float64, requires_grad=True)
        for correct in (True, False):
            sob = SobelFilter(H, correct=correct, device="cpu")
            for a in ("HSOBEL_WEIGHTS_3x3", "VSOBEL_WEIGHTS_3x3", "modifier"):
                setattr(sob, a, getattr(sob, a).double())
            gh, gv = sob.grad_h(img), sob.grad_v(img)
            w = torch.tensor(rs.standard_normal(gh.shape))
            ah, = torch.autograd.grad((gh * w).sum(), img, retain_graph=True)
            av, = torch.autograd.grad((gv * w).sum(), img)
            c = int(correct)
            d[f"{tag}{c}_gh"], d[f"{tag}{c}_gv"] = gh.detach().numpy(), gv.detach().numpy()
            d[f"{tag}{c}_w"], d[f"{tag}{c}_ah"], d[f"{tag}{c}_av"] = w.numpy(), ah.numpy(), av.numpy()
        d[f"{tag}_img"] = img.detach().numpy()
    for tag, (B, H) in dict(p=(3, 16), q=(1, 65), r=(2, 64)).items():
        K = torch.tensor(np.exp(0.5 * rs.standard_normal((B, 1, H, H))))
        out = torch.tensor(rs.standard_normal((B, 3, H, H)), requires_grad=True)
        sob = SobelFilter(H, correct=True, device="cpu")
        for a in ("HSOBEL_WEIGHTS_3x3", "VSOBEL_WEIGHTS_3x3", "modifier"):
            setattr(sob, a, getattr(sob, a).double())
        for tb in (1, 0):
            l_c = darcy.conv_constitutive_constraint(K, out, sob)
            l_d = darcy.conv_continuity_constraint(out, sob, use_tb=bool(tb))
            l_dir, l_neu = darcy.conv_boundary_condition(out)
            gw = torch.tensor([0.7, 1.3, 10.0, 4.0], dtype=torch.float64)
            tot = gw[0] * l_c + gw[1] * l_d + gw[2] * l_dir + gw[3] * l_neu
            g, = torch.autograd.grad(tot, out)
            d[f"{tag}{tb}_l4"] = torch.stack([l_c, l_d, l_dir, l_neu]).detach().numpy()
            d[f"{tag}{tb}_dout"] = g.numpy()
        d[f"{tag}_K"], d[f"{tag}_out"], d[f"{tag}_gw"] = K.numpy(), out.detach().numpy(), gw.numpy()
    np.savez_compressed(os.path.join(HERE, "sobel_darcy.npz"), **d)
    print("sobel_darcy.npz written")


if __name__ == "__main__":
    main()
