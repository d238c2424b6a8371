"""`Adam`: torch.optim.Adam whose step() is ONE launch of the fused flat-buffer kernel (pdes_adam_step, the
kernel the CUDA-graph engine uses) whenever a parameter group is exactly the parameter list of one executor
network (DenseED / Decoder / coupling network), i.e. the call train_codec_mixed_residual.py:151, 239 makes:

    optimizer = optim.Adam(model.parameters(), lr=args.lr, weight_decay=args.weight_decay)
    ...
    optimizer.step()

torch's own foreach implementation walks the 82 parameter views with ~10 multi-tensor launches (194 us of GPU
time per step at the benchmark shape, measured with tools/e2e_breakdown.py); the parameters, their gradients
and - here - both moment buffers are views into flat buffers, so the same update is one elementwise pass.
Everything else (other parameter sets, amsgrad / maximize / capturable, missing gradients, CPU tensors) goes
to the stock torch step: that is torch's implementation, not a fallback of this repo's kernels.

`install()` makes `torch.optim.Adam` resolve to this class, which is how run_reference_script.py runs the
UNMODIFIED script with it (PDES_FUSED_ADAM=0 keeps the stock class).  The state keeps torch's layout
(per-parameter 'step', 'exp_avg', 'exp_avg_sq'; the moments are views of the flat buffers, 'step' is one
shared tensor per group), so state_dict() / load_state_dict() / lr schedulers / OneCycle beta schedules work
unchanged.
"""
import os

import torch

from . import _lib
from . import codec as _codec

_StockAdam = torch.optim.Adam
if getattr(_StockAdam, "_pdes_fused", False):   # re-import after install(): keep the true base class
    _StockAdam = _StockAdam.__mro__[1]


class Adam(_StockAdam):
    _pdes_fused = True

    def __init__(self, params, *args, **kwargs):
        super(Adam, self).__init__(params, *args, **kwargs)
        self._pdes_bound = {}   # group index -> dict(owner, flat_ptr, m, v, step_t)
        self._pdes_checked = {}  # id(group's parameter list) -> owner whose parameter list it was verified to be
        self.fused_steps = 0    # number of step() calls served by the fused kernel (tests / bench read it)

    # ------------------------------------------------------------------ eligibility
    @staticmethod
    def _plain(group):
        return not (group.get("amsgrad") or group.get("maximize") or group.get("capturable") or
                    group.get("differentiable") or group.get("fused") or group.get("foreach") is False or
                    isinstance(group["lr"], torch.Tensor) or group.get("decoupled_weight_decay"))

    def _owner(self, group):
        params = group["params"]
        if not params:
            return None
        owner = _codec.owner_of(params[0])
        if owner is None or owner._flat is None or not owner._flat.is_cuda or owner._flat.dtype != torch.float32:
            return None
        mine, views = owner._params, owner._grad_views
        if len(mine) != len(params):
            return None
        known = self._pdes_checked.get(id(params)) is owner
        if not known:
            for a, b in zip(params, mine):
                if a is not b:
                    return None
            self._pdes_checked[id(params)] = owner
        # the gradients must be the views of the flat gradient buffer: every backward pass of the executor leaves
        # them so (codec._prepare_grads), and zero_grad() sets all of them to None together - the two ends tell
        for i in (0, len(params) - 1):
            g = params[i].grad
            if g is None or g.data_ptr() != views[i].data_ptr():
                return None
        return owner

    def _bind(self, gi, group, owner):
        b = self._pdes_bound.get(gi)
        if b is not None and b["owner"] is owner and b["flat_ptr"] == owner._flat.data_ptr():
            return b
        flat = owner._flat
        m, v = torch.zeros_like(flat), torch.zeros_like(flat)
        step_t = None
        for p, (_name, off, shape, _kind) in zip(owner._params, owner._param_table):
            n = p.numel()
            mv, vv = m[off:off + n].view(shape), v[off:off + n].view(shape)
            st = self.state[p]
            if "exp_avg" in st:   # earlier stock steps / a loaded state_dict: carry the moments over
                mv.copy_(st["exp_avg"])
                vv.copy_(st["exp_avg_sq"])
                if step_t is None:
                    step_t = torch.as_tensor(st["step"], dtype=torch.float32).detach().clone().cpu()
            st["exp_avg"], st["exp_avg_sq"] = mv, vv
        if step_t is None:
            step_t = torch.tensor(0.0, dtype=torch.float32)
        for p in owner._params:
            self.state[p]["step"] = step_t
        b = dict(owner=owner, flat_ptr=flat.data_ptr(), m=m, v=v, step_t=step_t)
        self._pdes_bound[gi] = b
        return b

    def _unbind(self):
        """Back to torch's own layout (one 'step' tensor per parameter) before the stock step touches the state."""
        for b in self._pdes_bound.values():
            for p in b["owner"]._params:
                st = self.state.get(p)
                if st is not None and "step" in st:
                    st["step"] = b["step_t"].clone()
        self._pdes_bound = {}

    def load_state_dict(self, state_dict):
        self._pdes_bound = {}   # the loaded moments are fresh tensors: re-bind (and copy them) at the next step
        return super(Adam, self).load_state_dict(state_dict)

    # ------------------------------------------------------------------ step
    @torch.no_grad()
    def step(self, closure=None):
        owners = []
        if os.environ.get("PDES_FUSED_ADAM", "1") != "0":
            for group in self.param_groups:
                owner = self._owner(group) if self._plain(group) else None
                if owner is None:
                    owners = None
                    break
                owners.append(owner)
        if not owners:
            self._unbind()
            return super(Adam, self).step(closure)
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        L = _lib.lib()
        for gi, (group, owner) in enumerate(zip(self.param_groups, owners)):
            b = self._bind(gi, group, owner)
            b["step_t"] += 1
            beta1, beta2 = group["betas"]
            flat = owner._flat
            with torch.cuda.device(flat.device):
                _lib.check(L.pdes_adam_step(_lib.ptr(flat), _lib.ptr(owner._flat_grad), _lib.ptr(b["m"]), _lib.ptr(b["v"]),
                                            flat.numel(), float(group["lr"]), float(beta1), float(beta2),
                                            float(group["eps"]), float(group["weight_decay"]), 1.0,
                                            int(b["step_t"].item()), _lib.stream_ptr()), "pdes_adam_step")
        self.fused_steps += 1
        return loss


def install():
    """Make `torch.optim.Adam` (what the unmodified scripts construct) resolve to the fused subclass."""
    if os.environ.get("PDES_FUSED_ADAM", "1") == "0":
        return False
    torch.optim.Adam = Adam
    if hasattr(torch.optim, "adam"):
        torch.optim.adam.Adam = Adam
    return True


def uninstall():
    torch.optim.Adam = _StockAdam
    if hasattr(torch.optim, "adam"):
        torch.optim.adam.Adam = _StockAdam
