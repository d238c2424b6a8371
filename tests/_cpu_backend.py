"""Oracle-backed stand-ins for the CUDA executor and the fused loss, used ONLY by the CPU tests to
exercise the host-side logic (module tree, flat storage, autograd wiring, memoisation, the
unmodified training script) on a box without a GPU.  The product never imports this."""
import contextlib

import torch

from oracle import pdes_oracle as orc


class OracleExecutor(object):
    def __init__(self, module):
        self.m = module
        self.ctx = None

    def _plan_state(self, requires_grad):
        m = self.m
        cfg = dict(m._cfg)
        self.drop_rate = float(cfg.pop("drop_rate", 0.0))
        self.upsample = cfg.pop("upsample", "nearest")
        plan = orc.densenet_plan(**cfg, upsample=self.upsample)   # (cfg carries `bottleneck` = bn_size or 0)
        sd = {}
        for k, v in m.state_dict().items():
            sd[k] = v.detach().clone() if k.endswith("num_batches_tracked") else v.detach()
        leaves = {}
        for name, p in m.named_parameters():
            t = p.detach().clone().requires_grad_(requires_grad)
            sd[name] = t
            leaves[name] = t
        return plan, sd, leaves

    def forward(self, x, training):
        need = bool(training)  # autograd.Function.forward runs with grad mode off: always keep a graph
        plan, sd, leaves = self._plan_state(need)
        xin = x.detach().clone().requires_grad_(need)
        with torch.enable_grad() if need else torch.no_grad():
            out = orc.densenet_forward(plan, sd, xin, training=training, drop_rate=self.drop_rate, upsample=self.upsample)
        self.ctx = (out, leaves, xin) if need else None
        self.fwd_gen = getattr(self, "fwd_gen", 0) + 1
        return out.detach()

    def backward(self, dout, want_dx=False):
        out, leaves, xin = self.ctx
        out.backward(dout)
        for (name, p), v in zip(self.m.named_parameters(), self.m._grad_views):
            v.add_(leaves[name].grad)
        self.ctx = None
        return xin.grad if want_dx else None

    def flops(self, B, training):
        return 0.0


class _OracleDarcyFn(object):
    @staticmethod
    def apply(K, out, use_tb, beta1=0.0, beta2=0.0):
        if K is None:
            c = out.new_zeros(())
        elif beta1 != 0.0 or beta2 != 0.0:
            c = orc.constitutive_nonlinear(K, out, beta1, beta2)
        else:
            c = orc.constitutive(K, out)
        d, n = orc.boundary(out)
        return torch.stack([c, orc.continuity(out, use_tb), d, n])


@contextlib.contextmanager
def cpu_backend():
    """Patch the product modules to run on the oracle (CPU).  Test-only."""
    from pde_surrogate_b200 import codec, darcy, image_gradient
    saved = (codec._executor_factory, darcy._DarcyLossFn, darcy._check, image_gradient._sobel_call)
    codec._executor_factory = OracleExecutor
    darcy._DarcyLossFn = _OracleDarcyFn
    darcy._check = lambda t, name, channels: None
    darcy._memo.clear()

    def sobel_call(image, direction, correct, adjoint):
        if adjoint:
            with torch.enable_grad():   # (called from an autograd.Function.backward: grad mode is off there)
                x = torch.zeros_like(image, requires_grad=True)
                y = orc.sobel_grad_h(x, correct) if direction == 0 else orc.sobel_grad_v(x, correct)
                g, = torch.autograd.grad(y, x, image)
            return g
        with torch.no_grad():
            return orc.sobel_grad_h(image, correct) if direction == 0 else orc.sobel_grad_v(image, correct)

    image_gradient._sobel_call = sobel_call
    try:
        yield
    finally:
        codec._executor_factory, darcy._DarcyLossFn, darcy._check, image_gradient._sobel_call = saved
        darcy._memo.clear()
