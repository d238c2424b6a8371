"""CPU oracle for the physics-constrained DenseED training hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package (`pde_surrogate_b200/`, `models/`,
`utils/`) may import this module; only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` do, and only as the checker or the
CPU baseline — never as the thing shipped.

It restates, with plain `torch` CPU ops and explicit slicing arithmetic, what the reference
computes on this path.  Each function cites the reference file:line it follows
(paths relative to the cics-nd/pde-surrogate root).  The arithmetic of the reference lives in
PyTorch (un-vendored dependency, requirements.txt:1 `pytorch>=1.0.0`; 2.11.0 here); its
documented semantics are used: conv2d = cross-correlation, BatchNorm2d normalises with the
biased batch variance and updates running_var with the unbiased one, nearest upsampling maps
dst -> floor(dst/2).

Pinning: the reference ships no tests or golden vectors, so the oracle is
pinned against outputs of the reference itself, imported from /root/reference in the build
container by tests/golden/make_golden.py; tests/test_oracle.py replays those fixtures.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

# ---------------------------------------------------------------------------------------
# DenseED structure (models/codec.py:211-293)
# ---------------------------------------------------------------------------------------


def _dense_stages(name, cin, growth, bottleneck):
    """One _DenseLayer (codec.py:43-75).  `bottleneck` = bn_size when DenseED(bottleneck=True), else 0: a layer whose
    input is wider than bn_size * growth takes the bottleneck form norm1-relu-conv1 (1x1, cin -> bn_size * growth)
    - norm2-relu-conv2 (3x3 -> growth) (lines 56-64), expressed here as two 'bnconv' stages: the first keeps the
    layer input aside, the second concatenates it with its output (line 75)."""
    if bottleneck and cin > bottleneck * growth:
        mid = bottleneck * growth
        return [dict(kind="bnconv", name=name, bn="norm1", conv="conv1", cin=cin, cout=mid, k=1, stride=1, pad=0,
                     up=False, keep=True),
                dict(kind="bnconv", name=name, bn="norm2", conv="conv2", cin=mid, cout=growth, k=3, stride=1, pad=1,
                     up=False, cat=True)]
    return [dict(kind="dense", name=name, cin=cin, cout=growth, k=3, stride=1, pad=1, up=False)]


def _up_stage(t, c, upsample):
    """conv2 of a decoding transition (codec.py:136-150): Conv2d behind a x2 upsampling, or - upsample=None - the
    transposed convolution convT2 = ConvTranspose2d(c, c, 3, stride 2, padding 1, output_padding 1) (139-142)."""
    if upsample is None:
        return dict(kind="bnconv", name=t, bn="norm2", conv="convT2", cin=c, cout=c, k=3, stride=1, pad=1, up=False,
                    convT=True)
    return dict(kind="bnconv", name=t, bn="norm2", conv="conv2", cin=c, cout=c, k=3, stride=1, pad=1, up=True)


def decoder_plan(dim_latent, out_channels, blocks, growth_rate=16, init_features=48, upsample="nearest"):
    """Stage list of `Decoder` (models/codec.py:321-370): conv0 = Conv2d(dim_latent, init_features, 3, 1, 1)
    (line 331), dense decoding blocks (333-338), nearest-upsampling transitions between them (341-348),
    last decoding (351-353).  Same stage kinds as densenet_plan."""
    st = [dict(kind="conv", name="features.conv0", cin=dim_latent, cout=init_features, k=3, stride=1, pad=1,
               up=False)]
    c = init_features
    blocks = list(blocks)
    for i, n in enumerate(blocks):
        for j in range(n):
            st.append(dict(kind="dense", name=f"features.DecBlock{i + 1}.denselayer{j + 1}",
                           cin=c + j * growth_rate, cout=growth_rate, k=3, stride=1, pad=1, up=False))
        c += n * growth_rate
        if i < len(blocks) - 1:
            t = f"features.TransUp{i + 1}"
            st.append(dict(kind="bnconv", name=t, bn="norm1", conv="conv1", cin=c, cout=c // 2, k=1,
                           stride=1, pad=0, up=False))
            st.append(_up_stage(t, c // 2, upsample))
            c //= 2
    t = "features.LastTransUp"
    st.append(dict(kind="bnconv", name=t, bn="norm1", conv="conv1", cin=c, cout=c // 2, k=3, stride=1,
                   pad=1, up=False))
    # last_decoding adds an upsampling module only for 'nearest' / 'bilinear' (codec.py:176-179)
    st.append(dict(kind="bnconv", name=t, bn="norm2", conv="conv2", cin=c // 2, cout=c // 4, k=3,
                   stride=1, pad=1, up=upsample is not None))
    st.append(dict(kind="bnconv", name=t, bn="norm3", conv="conv3", cin=c // 4, cout=out_channels, k=5,
                   stride=1, pad=2, up=False))
    return st


def coupling_plan(in_features, out_features, num_layers=3, growth_rate=16):
    """Stage list of the cGlow coupling network `_DenseCoupling` (models/glow_msc.py:276-294): `num_layers` dense
    layers on the input (281-284), then reduce = BatchNorm -> ReLU -> Conv2dZeros (287-293; Conv2dZeros 240-255:
    3x3 convolution WITH bias, times exp(3 * scale))."""
    st = []
    for j in range(num_layers):
        st.append(dict(kind="dense", name=f"denselayer{j + 1}", cin=in_features + j * growth_rate, cout=growth_rate,
                       k=3, stride=1, pad=1, up=False))
    st.append(dict(kind="zeros", name="reduce", bn="norm1", conv="conv_zero.conv", cin=in_features + num_layers * growth_rate,
                   cout=out_features, k=3, stride=1, pad=1, up=False))
    return st


def densenet_plan(in_channels=1, out_channels=3, imsize=64, blocks=(6, 8, 6), growth_rate=16,
                  init_features=48, arch=0, upsample="nearest", bottleneck=0):
    """Ordered list of stages describing DenseED with the defaults the training script uses
    (bottleneck=False in dense layers, bottleneck=True transitions, upsample='nearest',
    drop_rate=0, out_activation=None).

    Each stage is a dict.  kind:
      'conv'   : plain convolution (In_conv)                         codec.py:242-243
      'dense'  : BN -> ReLU -> conv3x3(cin->growth) -> cat([x, y])   codec.py:65-69, 73-75
      'bnconv' : BN -> ReLU -> [nearest x2] -> conv                  codec.py:103-150, 163-188
    """
    if arch == 1:
        return decoder_plan(in_channels, out_channels, blocks, growth_rate, init_features, upsample)
    if arch == 2:
        return coupling_plan(in_channels, out_channels, list(blocks)[0], growth_rate)
    blocks = list(blocks)
    if len(blocks) > 1 and len(blocks) % 2 == 0:
        raise ValueError("length of blocks must be odd")  # codec.py:231-233
    enc = blocks[: len(blocks) // 2]  # codec.py:234
    dec = blocks[len(blocks) // 2:]  # codec.py:235
    pad = 3 if imsize % 2 == 0 else 2  # codec.py:238
    st = []
    st.append(dict(kind="conv", name="features.In_conv", cin=in_channels, cout=init_features, k=7,
                   stride=2, pad=pad, up=False))
    c = init_features
    for i, n in enumerate(enc):  # codec.py:247-262
        for j in range(n):
            st += _dense_stages(f"features.EncBlock{i + 1}.denselayer{j + 1}", c + j * growth_rate, growth_rate, bottleneck)
        c += n * growth_rate
        t = f"features.TransDown{i + 1}"
        st.append(dict(kind="bnconv", name=t, bn="norm1", conv="conv1", cin=c, cout=c // 2, k=1,
                       stride=1, pad=0, up=False))
        st.append(dict(kind="bnconv", name=t, bn="norm2", conv="conv2", cin=c // 2, cout=c // 2, k=3,
                       stride=2, pad=1, up=False))
        c //= 2
    for i, n in enumerate(dec):  # codec.py:265-282
        for j in range(n):
            st += _dense_stages(f"features.DecBlock{i + 1}.denselayer{j + 1}", c + j * growth_rate, growth_rate, bottleneck)
        c += n * growth_rate
        if i < len(dec) - 1:
            t = f"features.TransUp{i + 1}"
            st.append(dict(kind="bnconv", name=t, bn="norm1", conv="conv1", cin=c, cout=c // 2, k=1,
                           stride=1, pad=0, up=False))
            st.append(_up_stage(t, c // 2, upsample))
            c //= 2
    t = "features.LastTransUp"  # codec.py:163-188
    st.append(dict(kind="bnconv", name=t, bn="norm1", conv="conv1", cin=c, cout=c // 2, k=3, stride=1,
                   pad=1, up=False))
    # last_decoding adds an upsampling module only for 'nearest' / 'bilinear' (codec.py:176-179)
    st.append(dict(kind="bnconv", name=t, bn="norm2", conv="conv2", cin=c // 2, cout=c // 4, k=3,
                   stride=1, pad=1, up=upsample is not None))
    st.append(dict(kind="bnconv", name=t, bn="norm3", conv="conv3", cin=c // 4, cout=out_channels, k=5,
                   stride=1, pad=2, up=False))
    return st


def _bn_name(s):
    return s["name"] + "." + (s["bn"] if s["kind"] in ("bnconv", "zeros") else "norm1")


def _conv_name(s):
    if s["kind"] == "conv":
        return s["name"]
    return s["name"] + "." + (s["conv"] if s["kind"] in ("bnconv", "zeros") else "conv1")


def state_layout(plan):
    """(name, shape) of every state_dict entry in the reference's order
    (BatchNorm: weight, bias, running_mean, running_var, num_batches_tracked; then conv)."""
    out = []
    for s in plan:
        if s["kind"] != "conv":
            b = _bn_name(s)
            c = s["cin"]
            out += [(b + ".weight", (c,)), (b + ".bias", (c,)), (b + ".running_mean", (c,)),
                    (b + ".running_var", (c,)), (b + ".num_batches_tracked", ())]
        if s["kind"] == "zeros":   # Conv2dZeros: the module's own `scale` comes before its child convolution
            out.append((s["name"] + ".conv_zero.scale", (1, s["cout"], 1, 1)))
        out.append((_conv_name(s) + ".weight", (s["cout"], s["cin"], s["k"], s["k"])))
        if s["kind"] == "zeros":
            out.append((_conv_name(s) + ".bias", (s["cout"],)))
    return out


def param_names(plan):
    return [n for n, _ in state_layout(plan) if not n.endswith(("running_mean", "running_var",
                                                                "num_batches_tracked"))]


def make_state(plan, seed=0, dtype=torch.float32):
    """Deterministic, well-scaled synthetic weights (numpy legacy MT19937 stream: identical
    on every machine, so fixtures need not store 740k weights)."""
    rs = np.random.RandomState(seed)
    sd = OrderedDict()
    for name, shape in state_layout(plan):
        if name.endswith("num_batches_tracked"):
            sd[name] = torch.zeros((), dtype=torch.int64)
        elif name.endswith("running_mean"):
            sd[name] = torch.tensor(0.1 * rs.standard_normal(shape), dtype=dtype)
        elif name.endswith("running_var"):
            sd[name] = torch.tensor(rs.uniform(0.5, 1.5, shape), dtype=dtype)
        elif name.endswith(".scale"):
            sd[name] = torch.tensor(rs.uniform(-0.2, 0.2, shape), dtype=dtype)
        elif len(shape) == 4:
            bound = 1.0 / math.sqrt(shape[1] * shape[2] * shape[3])
            sd[name] = torch.tensor(rs.uniform(-bound, bound, shape), dtype=dtype)
        elif name.endswith(".weight"):
            sd[name] = torch.tensor(rs.uniform(0.5, 1.5, shape), dtype=dtype)
        else:
            sd[name] = torch.tensor(rs.uniform(-0.3, 0.3, shape), dtype=dtype)
    return sd


def make_input(B, imsize, seed=0, dtype=torch.float32, kind="lognormal"):
    """Positive permeability-like field.  kind "lognormal": K = exp(0.5 * noise);  kind "channel":
    a two-valued channelized field K in {1, e^2.5} (thresholded smooth noise) like the reference's
    `channelized` dataset (train_codec_mixed_residual.py:55, BASELINE config 3)."""
    rs = np.random.RandomState(1000 + seed)
    g = rs.standard_normal((B, 1, imsize, imsize))
    if kind == "channel":
        k = max(3, imsize // 8) | 1   # odd box width; anisotropic smoothing gives channel-like bands
        pad = k // 2
        gp = np.pad(g, ((0, 0), (0, 0), (0, 0), (pad, pad)), mode="wrap")
        sm = sum(gp[..., i:i + imsize] for i in range(k)) / k
        gp = np.pad(sm, ((0, 0), (0, 0), (1, 1), (0, 0)), mode="wrap")
        sm = (gp[:, :, :-2] + gp[:, :, 1:-1] + gp[:, :, 2:]) / 3.0
        return torch.tensor(np.where(sm > 0.0, np.exp(2.5), 1.0), dtype=dtype)
    return torch.tensor(np.exp(0.5 * g), dtype=dtype)


def round_operand(t, mode, scale_log2=0):
    """Value of a convolution operand as the reduced-precision tensor-core modes see it (BASELINE config 3,
    SURVEY.md section 8d "CPU emulation with bf16-rounded conv operands"): one bf16 piece, or one fp16 piece of
    the power-of-two scaled value.  Products of the rounded operands are exact in fp32; accumulation is fp32."""
    if mode is None:
        return t
    if mode == "bf16":
        return t.detach().to(torch.bfloat16).to(t.dtype) + (t - t.detach())   # straight-through for autograd
    if mode == "fp16":
        s = 2.0 ** scale_log2
        return ((t.detach() * s).to(torch.float16).to(t.dtype) / s) + (t - t.detach())
    raise ValueError(mode)


def _drop_site(s):
    """nn.Dropout2d follows every convolution but In_conv / conv0 and the last two of the last decoding
    (codec.py:70-71, 110-149, 171-172)."""
    return (s["kind"] != "conv" and not s.get("keep") and
            not (s["name"] == "features.LastTransUp" and s.get("conv") in ("conv2", "conv3")))


def densenet_forward(plan, sd, x, training=True, momentum=0.1, eps=1e-5, update_running=True, operand_round=None,
                     drop_rate=0.0, upsample="nearest"):
    """DenseED.forward (codec.py:295-296) as a flat functional program.

    operand_round = "bf16" | "fp16": emulate the one-piece tensor-core modes - every convolution but the first
    (which runs on exact-fp32 CUDA-core kernels) sees rounded activations and filters.

    sd maps reference state_dict keys to tensors (parameters may require grad).  In training
    mode running_mean / running_var / num_batches_tracked entries of `sd` are updated in place
    exactly as nn.BatchNorm2d does.
    """
    h = x
    kept = None
    for s in plan:
        if s.get("keep"):
            kept = h   # input of a bottleneck dense layer (concatenated with the layer's output below)
        if s["kind"] == "conv":
            h = F.conv2d(h, sd[_conv_name(s) + ".weight"], None, s["stride"], s["pad"])
            continue
        b = _bn_name(s)
        rm, rv = sd[b + ".running_mean"], sd[b + ".running_var"]
        if training:
            # nn.BatchNorm2d training semantics (codec.py:57-66,103,113,135,167,173,184)
            a = F.batch_norm(h, rm if update_running else None, rv if update_running else None,
                             sd[b + ".weight"], sd[b + ".bias"], True, momentum, eps)
            if update_running:
                sd[b + ".num_batches_tracked"] += 1
        else:
            a = F.batch_norm(h, rm, rv, sd[b + ".weight"], sd[b + ".bias"], False, momentum, eps)
        a = F.relu(a)
        if s["up"] and upsample == "bilinear":
            a = F.interpolate(a, scale_factor=2.0, mode="bilinear", align_corners=True)  # codec.py:33-40
        elif s["up"]:
            a = F.interpolate(a, scale_factor=2.0, mode="nearest")  # codec.py:24-30
        w = sd[_conv_name(s) + ".weight"]
        if operand_round is not None:
            a, w = round_operand(a, operand_round, 4), round_operand(w, operand_round, 8)
        if s["kind"] == "zeros":   # Conv2dZeros.forward (glow_msc.py:253-255)
            y = F.conv2d(a, w, sd[_conv_name(s) + ".bias"], s["stride"], s["pad"])
            h = y * torch.exp(sd[s["name"] + ".conv_zero.scale"] * 3)
            continue
        if s.get("convT"):   # nn.ConvTranspose2d(k3, s2, p1, op1), weight (cin, cout, 3, 3)  (codec.py:139-142)
            y = F.conv_transpose2d(a, w, None, stride=2, padding=1, output_padding=1)
        else:
            y = F.conv2d(a, w, None, s["stride"], s["pad"])
        if drop_rate > 0 and _drop_site(s):
            y = F.dropout2d(y, drop_rate, training)   # nn.Dropout2d: whole channels, scaled by 1/(1-p)
        if s.get("cat"):
            h = torch.cat([kept, y], 1)
        else:
            h = torch.cat([h, y], 1) if s["kind"] == "dense" else y  # codec.py:73-75
    return h


# ---------------------------------------------------------------------------------------
# Sobel stencils (utils/image_gradient.py:24-92), filter_size = 3
# ---------------------------------------------------------------------------------------


def _pad_rep(img):
    return F.pad(img, (1, 1, 1, 1), mode="replicate")  # image_gradient.py:68, 85


def sobel_grad_h(img, correct=True):
    """SobelFilter.grad_h: d/dx.  Cross-correlation with VSOBEL = [[-1,0,1],[-2,0,2],[-1,0,1]]/8
    (image_gradient.py:28-33, 62-69), times image width (line 69), then right-multiplication by
    the modifier matrix (lines 43-46, 72-73): column 0 := 4 g0 - g1, column W-1 := 4 g[W-1] - g[W-2]."""
    W = img.shape[-1]
    p = _pad_rep(img)
    g = ((p[..., 0:-2, 2:] - p[..., 0:-2, 0:-2]) + 2.0 * (p[..., 1:-1, 2:] - p[..., 1:-1, 0:-2]) +
         (p[..., 2:, 2:] - p[..., 2:, 0:-2])) / 8.0 * W
    if not correct:
        return g
    first = 4.0 * g[..., :, 0:1] - g[..., :, 1:2]
    last = 4.0 * g[..., :, -1:] - g[..., :, -2:-1]
    return torch.cat([first, g[..., :, 1:-1], last], dim=-1)


def sobel_grad_v(img, correct=True):
    """SobelFilter.grad_v: d/dy with HSOBEL (image_gradient.py:28-31, 77-92); rows 0 and H-1
    corrected by left-multiplication with modifier^T."""
    H = img.shape[-2]
    p = _pad_rep(img)
    g = ((p[..., 2:, 0:-2] - p[..., 0:-2, 0:-2]) + 2.0 * (p[..., 2:, 1:-1] - p[..., 0:-2, 1:-1]) +
         (p[..., 2:, 2:] - p[..., 0:-2, 2:])) / 8.0 * H
    if not correct:
        return g
    first = 4.0 * g[..., 0:1, :] - g[..., 1:2, :]
    last = 4.0 * g[..., -1:, :] - g[..., -2:-1, :]
    return torch.cat([first, g[..., 1:-1, :], last], dim=-2)


# ---------------------------------------------------------------------------------------
# Darcy mixed-residual loss (models/darcy.py:162-176, 210-224, 226-233)
# ---------------------------------------------------------------------------------------


def constitutive(K, out):
    """conv_constitutive_constraint (darcy.py:170-176)."""
    u = out[:, 0:1]
    r1 = out[:, 1:2] + K * sobel_grad_h(u)
    r2 = out[:, 2:3] + K * sobel_grad_v(u)
    return (r1 ** 2 + r2 ** 2).mean()


def constitutive_nonlinear_exp(K, out):
    """conv_constitutive_constraint_nonlinear_exp (darcy.py:193-207): sigma = -exp(K u) grad(u)."""
    u = out[:, 0:1]
    ke = torch.exp(K * u)
    return ((out[:, 1:2] + ke * sobel_grad_h(u)) ** 2 + (out[:, 2:3] + ke * sobel_grad_v(u)) ** 2).mean()


def energy_functional_exp(K, u):
    """energy_functional_exp (darcy.py:151-159): mean[0.5 exp(K u) |grad u|^2]."""
    return (0.5 * torch.exp(K * u) * (sobel_grad_h(u) ** 2 + sobel_grad_v(u) ** 2)).mean()


def constitutive_nonlinear(K, out, beta1, beta2):
    """conv_constitutive_constraint_nonlinear (darcy.py:179-191):
    -K grad(u) = sigma + beta1 sqrt(K) sigma^2 + beta2 K sigma^3, squared residual mean."""
    ku_h = -K * sobel_grad_h(out[:, 0:1])
    ku_v = -K * sobel_grad_v(out[:, 0:1])
    sigma = out[:, 1:3]
    rhs = sigma + beta1 * torch.sqrt(K) * sigma ** 2 + beta2 * K * sigma ** 3
    return ((ku_h - rhs[:, 0:1]) ** 2 + (ku_v - rhs[:, 1:2]) ** 2).mean()


def affine_coupling(plan, sd, x, cond, reverse=False, training=True):
    """AffineCouplingLayer.forward / .reverse (models/glow_msc.py:326-344) around a coupling_plan network."""
    x1, x2 = x.chunk(2, 1)
    h = densenet_forward(plan, sd, torch.cat((x1, cond), 1), training=training)
    shift = h[:, 0::2]
    scale = torch.sigmoid(h[:, 1::2] + 2.0)
    x2 = (x2 / scale - shift) if reverse else ((x2 + shift) * scale)
    logdet = scale.log().view(x.shape[0], -1).sum(1)
    return torch.cat((x1, x2), 1), logdet


def continuity(out, use_tb=True):
    """conv_continuity_constraint (darcy.py:217-224)."""
    r3 = sobel_grad_h(out[:, 1:2]) + sobel_grad_v(out[:, 2:3])
    if use_tb:
        return (r3 ** 2).mean()
    return (r3 ** 2)[:, :, 1:-1, :].mean()


def boundary(out):
    """conv_boundary_condition (darcy.py:227-233) -> (dirichlet, neumann)."""
    left, right = out[:, 0, :, 0], out[:, 0, :, -1]
    tb = out[:, 2, [0, -1], :]
    return ((left - 1.0) ** 2).mean() + (right ** 2).mean(), (tb ** 2).mean()


def darcy_losses(K, out, use_tb=True):
    d, n = boundary(out)
    return torch.stack([constitutive(K, out), continuity(out, use_tb), d, n])


def total_loss(K, out, weight_bound=10.0):
    """train_codec_mixed_residual.py:228-232."""
    l4 = darcy_losses(K, out)
    return (l4[0] + l4[1]) + (l4[2] + l4[3]) * weight_bound, l4


def train_step(plan, sd, K, weight_bound=10.0, upsample="nearest"):
    """One step body of train_codec_mixed_residual.py:226-233 (zero_grad, forward, loss,
    backward).  Returns output, 4 partial losses, loss, dL/d(output), {param name: grad}."""
    names = param_names(plan)
    for n in names:
        sd[n].requires_grad_(True)
        sd[n].grad = None
    out = densenet_forward(plan, sd, K, training=True, upsample=upsample)
    out.retain_grad()
    if out.shape[-1] * 2 == K.shape[-1]:
        # upsample=None: the output is imsize/2 wide (codec.py:176-179); the golden fixtures of that option take the
        # residual loss on the 2x subsampled permeability (tests/golden/make_golden.py ref_step)
        K = K[:, :, ::2, ::2]
    loss, l4 = total_loss(K, out, weight_bound)
    loss.backward()
    grads = OrderedDict((n, sd[n].grad.detach().clone()) for n in names)
    return out.detach(), l4.detach(), loss.detach(), out.grad.detach().clone(), grads


def adam_reference(p, g, m, v, lr, step, b1=0.9, b2=0.999, eps=1e-8, wd=0.0):
    """torch.optim.Adam single-tensor update (train_codec_mixed_residual.py:151, 239)."""
    if wd != 0.0:
        g = g + wd * p
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
    p = p - (lr / bc1) * m / (v.sqrt() / math.sqrt(bc2) + eps)
    return p, m, v


def to_dtype(sd, dtype):
    return OrderedDict((k, v.clone() if v.dtype == torch.int64 else v.detach().to(dtype).clone())
                       for k, v in sd.items())
