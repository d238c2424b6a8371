"""SobelFilter on the sm_100a stencil kernels.

Host-side mirror of utils/image_gradient.py:24-92 of the reference (same constructor, same
`grad_h` / `grad_v` methods, autograd-capable).  The arithmetic runs in
csrc/stencil.cu (pdes_sobel_grad); the backward pass applies the exact transpose operator.
"""
import torch

from . import _lib


class _SobelFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image, direction, correct):
        ctx.direction, ctx.correct = direction, correct
        return _sobel_call(image, direction, correct, adjoint=False)

    @staticmethod
    def backward(ctx, grad_out):
        return _sobel_call(grad_out, ctx.direction, ctx.correct, adjoint=True), None, None


def _sobel_call(image, direction, correct, adjoint):
    if not image.is_cuda:
        raise RuntimeError("pde_surrogate_b200.SobelFilter: CUDA tensors only (got device %s); there is "
                           "no CPU fallback in this backend" % image.device)
    if image.dtype != torch.float32:
        raise TypeError("pde_surrogate_b200.SobelFilter: float32 only, got %s" % image.dtype)
    if image.dim() < 2:
        raise ValueError("SobelFilter expects (..., H, W)")
    img = image.contiguous()
    H, W = img.shape[-2], img.shape[-1]
    out = torch.empty_like(img)
    n_img = img.numel() // (H * W) if img.numel() else 0
    with torch.cuda.device(img.device):
        rc = _lib.lib().pdes_sobel_grad(_lib.ptr(img), _lib.ptr(out), n_img, H, W, int(direction),
                                        int(bool(correct)), int(bool(adjoint)), _lib.stream_ptr())
    _lib.check(rc, "pdes_sobel_grad")
    return out


class SobelFilter(object):
    """3x3 Sobel finite-difference gradients with replicate padding and 3-point one-sided
    boundary correction (reference: utils/image_gradient.py:26-47)."""

    def __init__(self, imsize, correct=True, device='cpu'):
        self.imsize = imsize
        self.correct = correct
        self.device = torch.device(device) if not isinstance(device, torch.device) else device
        # the reference exposes these tensors as attributes; kept for introspection only
        h = torch.tensor([[-1., -2., -1.], [0., 0., 0.], [1., 2., 1.]]) / 8.0
        self.HSOBEL_WEIGHTS_3x3 = h.view(1, 1, 3, 3).to(self.device)
        self.VSOBEL_WEIGHTS_3x3 = self.HSOBEL_WEIGHTS_3x3.transpose(-1, -2)
        mod = torch.eye(imsize)
        mod[0, 0], mod[1, 0] = 4., -1.
        mod[-2, -1], mod[-1, -1] = -1., 4.
        self.modifier = mod.to(self.device)

    def _check(self, filter_size):
        if filter_size != 3:
            raise NotImplementedError("pde_surrogate_b200.SobelFilter: only filter_size=3 is implemented "
                                      "(the 5x5 kernels are never used by the training scripts)")

    def grad_h(self, image, filter_size=3):
        """d/dx (last dim); reference image_gradient.py:50-75."""
        self._check(filter_size)
        return _SobelFn.apply(image, 0, self.correct)

    def grad_v(self, image, filter_size=3):
        """d/dy (dim -2); reference image_gradient.py:77-92."""
        self._check(filter_size)
        return _SobelFn.apply(image, 1, self.correct)
