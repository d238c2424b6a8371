"""Darcy mixed-residual loss on the fused sm_100a stencil kernel.

Host-side mirror of models/darcy.py:162-176 (conv_constitutive_constraint), 210-224
(conv_continuity_constraint) and 226-233 (conv_boundary_condition).  The training script calls
the three functions one after another on the same `output` tensor
(train_codec_mixed_residual.py:228-230); the first call launches ONE fused kernel that produces
all four partial losses, the other two are served from a one-entry memo keyed on the identity
and version of the tensors, and the backward of all three is ONE fused kernel.
"""
import weakref

import torch

from . import _lib
from .image_gradient import SobelFilter

_ws = {}      # device index -> zero-initialised workspace of the loss kernel
_memo = {}    # device index -> dict(out_ref, out_ver, K_ref, K_ver, grad, l4, parts)


def _workspace(device):
    idx = device.index if device.index is not None else torch.cuda.current_device()
    w = _ws.get(idx)
    if w is None:
        w = torch.zeros(max(64, int(_lib.lib().pdes_darcy_loss_workspace_bytes())), dtype=torch.uint8,
                        device=device)
        _ws[idx] = w
    return w


def _check(t, name, channels):
    if not t.is_cuda:
        raise RuntimeError("pde_surrogate_b200.darcy: CUDA tensors only (%s is on %s); there is no CPU "
                           "fallback in this backend" % (name, t.device))
    if t.dtype != torch.float32:
        raise TypeError("pde_surrogate_b200.darcy: %s must be float32, got %s" % (name, t.dtype))
    if t.dim() != 4 or t.shape[1] != channels:
        raise ValueError("pde_surrogate_b200.darcy: %s must be (B, %d, H, W), got %s"
                         % (name, channels, tuple(t.shape)))


class _DarcyLossFn(torch.autograd.Function):
    """(K, out) -> tensor[4] = [constitutive, continuity, dirichlet, neumann]."""

    @staticmethod
    def forward(ctx, K, out, use_tb, beta1=0.0, beta2=0.0):
        outc = out.contiguous()
        Kc = K.contiguous() if K is not None else None
        B, _, H, W = outc.shape
        l4 = torch.empty(4, dtype=torch.float32, device=outc.device)
        ctx.betas = (float(beta1), float(beta2))
        with torch.cuda.device(outc.device):
            if ctx.betas != (0.0, 0.0):
                rc = _lib.lib().pdes_darcy_loss_nl_fwd(_lib.ptr(Kc), _lib.ptr(outc), B, H, W, int(use_tb),
                                                       ctx.betas[0], ctx.betas[1], _lib.ptr(l4),
                                                       _lib.ptr(_workspace(outc.device)), _lib.stream_ptr())
            else:
                rc = _lib.lib().pdes_darcy_loss_fwd(_lib.ptr(Kc), _lib.ptr(outc), B, H, W, int(use_tb),
                                                    _lib.ptr(l4), _lib.ptr(_workspace(outc.device)),
                                                    _lib.stream_ptr())
        _lib.check(rc, "pdes_darcy_loss_fwd")
        ctx.use_tb = use_tb
        ctx.has_K = Kc is not None
        if ctx.has_K:
            ctx.save_for_backward(Kc, outc)
        else:
            ctx.save_for_backward(outc)
        return l4

    @staticmethod
    def backward(ctx, g4):
        if ctx.has_K:
            Kc, outc = ctx.saved_tensors
        else:
            Kc, (outc,) = None, ctx.saved_tensors
        if ctx.has_K and ctx.needs_input_grad[0]:
            raise NotImplementedError("pde_surrogate_b200.darcy: gradient w.r.t. the permeability input "
                                      "is not implemented (the training path never needs it)")
        B, _, H, W = outc.shape
        g4 = g4.contiguous().float()
        dout = torch.empty_like(outc)
        with torch.cuda.device(outc.device):
            if ctx.betas != (0.0, 0.0):
                rc = _lib.lib().pdes_darcy_loss_nl_bwd(_lib.ptr(Kc), _lib.ptr(outc), _lib.ptr(g4), B, H, W,
                                                       int(ctx.use_tb), ctx.betas[0], ctx.betas[1],
                                                       _lib.ptr(dout), _lib.stream_ptr())
            else:
                rc = _lib.lib().pdes_darcy_loss_bwd(_lib.ptr(Kc), _lib.ptr(outc), _lib.ptr(g4), B, H, W,
                                                    int(ctx.use_tb), _lib.ptr(dout), _lib.stream_ptr())
        _lib.check(rc, "pdes_darcy_loss_bwd")
        return None, dout, None, None, None


def _alive(ref, t, ver):
    return ref is not None and ref() is t and t._version == ver


def _fused_parts(input, output, use_tb=True, betas=None):
    """The four partial losses of (input, output) as 0-d tensors, computed once per distinct pair.
    betas = (beta1, beta2) selects the nonlinear constitutive law for parts[0]; calls that do not name a
    law (continuity / boundary terms: betas None) are served from whatever evaluation is memoised."""
    idx = output.device.index
    grad_mode = torch.is_grad_enabled() and output.requires_grad
    m = _memo.get(idx)
    if m is not None and _alive(m["out_ref"], output, m["out_ver"]) and m["grad"] == grad_mode \
            and m["use_tb"] == use_tb and (betas is None or m.get("betas", (0.0, 0.0)) == betas):
        if input is None or (m["K_ref"] is not None and _alive(m["K_ref"], input, m["K_ver"])):
            return m["parts"]
    _check(output, "output", 3)
    if input is not None:
        _check(input, "input", 1)
        if input.shape[0] != output.shape[0] or input.shape[2:] != output.shape[2:]:
            raise ValueError("input %s and output %s do not match" % (tuple(input.shape), tuple(output.shape)))
    b = betas if betas is not None else (0.0, 0.0)
    if b != (0.0, 0.0):
        l4 = _DarcyLossFn.apply(input, output, bool(use_tb), b[0], b[1])
    else:
        l4 = _DarcyLossFn.apply(input, output, bool(use_tb))
    parts = l4.unbind(0)
    _memo[idx] = dict(betas=b, out_ref=weakref.ref(output), out_ver=output._version,
                      K_ref=weakref.ref(input) if input is not None else None,
                      K_ver=input._version if input is not None else None, grad=grad_mode,
                      use_tb=use_tb, parts=parts)
    return parts


def _needs_composite(sobel_filter):
    return isinstance(sobel_filter, SobelFilter) and not sobel_filter.correct


def conv_constitutive_constraint(input, output, sobel_filter):
    """sigma = -K grad(u): mean[(sigma1 + K du/dx)^2 + (sigma2 + K du/dy)^2]  (darcy.py:162-176)."""
    if _needs_composite(sobel_filter):
        gh = sobel_filter.grad_h(output[:, [0]])
        gv = sobel_filter.grad_v(output[:, [0]])
        return ((output[:, [1]] + input * gh) ** 2 + (output[:, [2]] + input * gv) ** 2).mean()
    return _fused_parts(input, output, betas=(0.0, 0.0))[0]


def conv_constitutive_constraint_nonlinear(input, output, sobel_filter, beta1, beta2):
    """Nonlinear extension of Darcy's law, -K grad(u) = sigma + beta1 sqrt(K) sigma^2 + beta2 K sigma^3
    (darcy.py:179-191; solve_conv_mixed_residual.py --nonlinear).  Same fused kernel pair as the linear law."""
    if _needs_composite(sobel_filter):
        ku_h = -input * sobel_filter.grad_h(output[:, [0]])
        ku_v = -input * sobel_filter.grad_v(output[:, [0]])
        sigma = output[:, [1, 2]]
        rhs = sigma + beta1 * torch.sqrt(input) * sigma ** 2 + beta2 * input * sigma ** 3
        return ((ku_h - rhs[:, [0]]) ** 2 + (ku_v - rhs[:, [1]]) ** 2).mean()
    return _fused_parts(input, output, betas=(float(beta1), float(beta2)))[0]


def conv_constitutive_constraint_nonlinear_exp(input, output, sobel_filter):
    """Exponential nonlinear law sigma = -exp(K u) grad(u) (darcy.py:193-207; no script uses it): residual of the two
    flux channels, on the Sobel kernels (`SobelFilter.grad_h / grad_v` are differentiable) and elementwise torch ops."""
    u = output[:, [0]]
    k_eff = torch.exp(input * u)
    r_h = output[:, [1]] + k_eff * sobel_filter.grad_h(u)
    r_v = output[:, [2]] + k_eff * sobel_filter.grad_v(u)
    return (r_h ** 2 + r_v ** 2).mean()


def energy_functional_exp(input, output, sobel_filter):
    """V(u, K) = mean[ 0.5 exp(K u) |grad u|^2 ] (darcy.py:151-159; no script uses it), on the Sobel kernels."""
    g2 = sobel_filter.grad_h(output) ** 2 + sobel_filter.grad_v(output) ** 2
    return (0.5 * torch.exp(input * output) * g2).mean()


def conv_continuity_constraint(output, sobel_filter, use_tb=True):
    """div(sigma) = 0: mean[(d sigma1/dx + d sigma2/dy)^2]  (darcy.py:210-224)."""
    if _needs_composite(sobel_filter):
        r = sobel_filter.grad_h(output[:, [1]]) + sobel_filter.grad_v(output[:, [2]])
        return (r ** 2).mean() if use_tb else (r ** 2)[:, :, 1:-1, :].mean()
    return _fused_parts(None, output, use_tb=bool(use_tb))[1]


def conv_boundary_condition(output):
    """(dirichlet, neumann) boundary losses  (darcy.py:226-233)."""
    parts = _fused_parts(None, output)
    return parts[2], parts[3]
