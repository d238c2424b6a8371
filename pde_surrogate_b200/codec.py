"""DenseED on the sm_100a executor.

Host-side mirror of models/codec.py:210-318 of the reference: same constructor signature,
same state_dict keys / shapes / order, same default initialisation stream, `model_size`,
`reset_parameters`, train()/eval() semantics of nn.BatchNorm2d — but the module tree only HOLDS
parameters.  All arithmetic (28 convolutions, 27 BatchNorm+ReLU, 2 nearest upsamplings, the
dense-block concatenations and the whole backward) runs in libpdes_b200.so
(csrc/net.cu, csrc/conv_*.cu) through the C-ABI in include/pdes_b200.h.

Storage: all parameters are views into ONE flat fp32 buffer (and their .grad into one flat
gradient buffer) so that the kernels see stable pointers, wgrad writes straight into the bucket
that data-parallel training all-reduces, and a fused Adam can walk one array.
"""
import math
import os
import weakref
from collections import OrderedDict
from ctypes import byref, c_int32, c_int64, c_void_p, create_string_buffer

import torch
import torch.nn as nn

from . import _lib


_OWNERS = weakref.WeakValueDictionary()   # id(parameter) -> the executor network whose flat buffer holds it


def owner_of(param):
    """The executor network (DenseED / Decoder / coupling network) `param` belongs to, or None."""
    m = _OWNERS.get(id(param))
    if m is None:
        return None
    for q in m._params:
        if q is param:
            return m
    return None


def module_size(module):
    """(n_params, n_conv_layers) — reference models/codec.py:14-21."""
    n_params, n_conv = 0, 0
    for name, p in module.named_parameters():
        n_conv += int("conv" in name)
        n_params += p.numel()
    return n_params, n_conv


class _Group(nn.Module):
    """Name-only container (a dense block / transition of the reference tree)."""

    def extra_repr(self):
        return getattr(self, "_pdes_repr", "")


class _ConvParams(nn.Module):
    """Holds `weight` (Cout, Cin, K, K); bias-free like every conv of DenseED."""

    def __init__(self, cout, cin, k, desc=""):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(cout, cin, k, k))
        self._pdes_repr = desc
        self.reset_parameters()

    def reset_parameters(self):
        # identical to nn.Conv2d.reset_parameters with bias=False (same RNG consumption)
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))

    def extra_repr(self):
        return self._pdes_repr


class _BNParams(nn.Module):
    """Holds the parameters / buffers of an nn.BatchNorm2d (eps 1e-5, momentum 0.1)."""

    def __init__(self, c):
        super().__init__()
        self.num_features = c
        self.weight = nn.Parameter(torch.ones(c))
        self.bias = nn.Parameter(torch.zeros(c))
        self.register_buffer("running_mean", torch.zeros(c))
        self.register_buffer("running_var", torch.ones(c))
        self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))

    def reset_parameters(self):
        with torch.no_grad():
            self.running_mean.zero_()
            self.running_var.fill_(1)
            self.num_batches_tracked.zero_()
            self.weight.fill_(1)
            self.bias.zero_()

    def extra_repr(self):
        return "%d, eps=1e-05, momentum=0.1 (fused into the consuming conv)" % self.num_features


class _NetHandle(object):
    """Owns a pdes_net_t* (host-side object of the C library)."""

    def __init__(self, cfg, max_batch, imsize=None):
        L = _lib.lib()
        c = _lib.DensenetConfig()
        self.imsize = int(cfg["imsize"] if imsize is None else imsize)
        c.in_channels, c.out_channels, c.imsize = cfg["in_channels"], cfg["out_channels"], self.imsize
        c.arch = int(cfg.get("arch", 0))
        c.dropout = 1 if cfg.get("drop_rate", 0.0) > 0 else 0
        c.upsample = {"nearest": 0, "bilinear": 1, None: 2}[cfg.get("upsample", "nearest")]
        c.bottleneck = int(cfg.get("bottleneck", 0))
        c.n_blocks = len(cfg["blocks"])
        for i, b in enumerate(cfg["blocks"]):
            c.blocks[i] = int(b)
        c.growth_rate, c.init_features, c.max_batch = cfg["growth_rate"], cfg["init_features"], max_batch
        h = c_void_p()
        _lib.check(L.pdes_densenet_create(byref(c), byref(h)), "pdes_densenet_create")
        self.h, self.max_batch = h, max_batch
        self.out_size = int(L.pdes_densenet_output_size(h))
        ns = int(L.pdes_densenet_dropout_sites(h, None, 0))
        ch = (c_int32 * max(ns, 1))()
        L.pdes_densenet_dropout_sites(h, ch, ns)
        self.drop_channels = [int(ch[i]) for i in range(ns)]

    def __del__(self):
        try:
            if self.h:
                _lib.lib().pdes_densenet_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def params(self):
        L, out = _lib.lib(), []
        for i in range(L.pdes_densenet_num_params(self.h)):
            name = create_string_buffer(256)
            off, nd, kind = c_int64(), c_int32(), c_int32()
            shape = (c_int64 * 4)()
            _lib.check(L.pdes_densenet_param_info(self.h, i, name, 256, byref(off), byref(nd), shape,
                                                  byref(kind)), "pdes_densenet_param_info")
            out.append((name.value.decode(), off.value, tuple(shape[:nd.value]), kind.value))
        return out, L.pdes_densenet_param_floats(self.h)

    def bns(self):
        L, out = _lib.lib(), []
        for i in range(L.pdes_densenet_num_bn(self.h)):
            name = create_string_buffer(256)
            mo, vo, ch = c_int64(), c_int64(), c_int32()
            _lib.check(L.pdes_densenet_bn_info(self.h, i, name, 256, byref(mo), byref(vo), byref(ch)),
                       "pdes_densenet_bn_info")
            out.append((name.value.decode(), mo.value, vo.value, ch.value))
        return out, L.pdes_densenet_running_floats(self.h)


class CudaExecutor(object):
    """Binds the module's flat buffers to a pdes_net_t and runs forward/backward on the current
    CUDA stream.  (Tests substitute this class to exercise the host logic without a GPU.)"""

    def __init__(self, module):
        self.m = module
        self.handle = None
        self.ws = None
        self.bound = None
        self.applied_impl = None
        self.fwd_gen = 0
        self.masks = None

    def _ensure(self, x):
        m = self.m
        if not x.is_cuda:
            raise RuntimeError("pde_surrogate_b200.DenseED: CUDA tensors only (input is on %s); there is no "
                               "CPU fallback in this backend" % x.device)
        if x.dtype != torch.float32 or m._flat.dtype != torch.float32:
            raise TypeError("pde_surrogate_b200.DenseED: float32 only (input %s, parameters %s)"
                            % (x.dtype, m._flat.dtype))
        if m._flat.device != x.device:
            raise RuntimeError("DenseED parameters are on %s but the input is on %s" % (m._flat.device, x.device))
        c = m._cfg
        if c.get("arch", 0) >= 1:
            # Decoder / coupling network: the input's spatial size is free (no imsize argument upstream)
            if x.dim() != 4 or x.shape[1] != c["in_channels"] or x.shape[2] != x.shape[3]:
                raise ValueError("expected a square input (B, %d, h, h), got %s" % (c["in_channels"], tuple(x.shape)))
            imsize = int(x.shape[2])
        else:
            if x.dim() != 4 or tuple(x.shape[1:]) != (c["in_channels"], c["imsize"], c["imsize"]):
                raise ValueError("DenseED expects input (B, %d, %d, %d), got %s"
                                 % (c["in_channels"], c["imsize"], c["imsize"], tuple(x.shape)))
            imsize = c["imsize"]
        B = x.shape[0]
        if self.handle is None or B > self.handle.max_batch or self.handle.imsize != imsize:
            with torch.cuda.device(x.device):
                self.handle = _NetHandle(c, max(B, 1), imsize)
            self.ws, self.bound = None, None
            self.applied_impl = None
        if self.applied_impl != m.conv_impl:
            _lib.check(_lib.lib().pdes_densenet_set_conv_impl(self.handle.h, int(m.conv_impl)),
                       "pdes_densenet_set_conv_impl")
            self.applied_impl = m.conv_impl
        key = (m._flat.data_ptr(), m._flat_grad.data_ptr(), m._flat_running.data_ptr(), str(x.device))
        if self.bound != key:
            L = _lib.lib()
            with torch.cuda.device(x.device):
                nbytes = int(L.pdes_densenet_workspace_bytes(self.handle.h))
                if self.ws is None or self.ws.numel() < nbytes or self.ws.device != x.device:
                    self.ws = torch.zeros(nbytes, dtype=torch.uint8, device=x.device)
                _lib.check(L.pdes_densenet_bind(self.handle.h, _lib.ptr(m._flat), _lib.ptr(m._flat_grad),
                                                _lib.ptr(m._flat_running), _lib.ptr(self.ws), nbytes),
                           "pdes_densenet_bind")
            self.bound = key

    def _dropout_masks(self, x, training):
        """nn.Dropout2d masks of this pass (models/codec.py:70-71, 110-149, 171-172): one (B, C, 1, 1) block per
        site, drawn exactly like torch.feature_dropout draws them (empty(B,C,1,1).bernoulli_(1-p).div_(1-p), in
        module order) so that a seeded run consumes the generator like the reference does."""
        p = float(self.m._cfg.get("drop_rate", 0.0))
        L = _lib.lib()
        if not (training and p > 0 and self.handle.drop_channels):
            if self.handle.drop_channels:
                _lib.check(L.pdes_densenet_set_dropout(self.handle.h, None), "pdes_densenet_set_dropout")
            self.masks = None
            return
        B = x.shape[0]
        dev = getattr(self.m, "_mask_device", None) or x.device   # test hook: draw on the CPU generator
        flat = torch.empty(B * sum(self.handle.drop_channels), dtype=torch.float32, device=dev)
        o = 0
        for C in self.handle.drop_channels:
            flat[o:o + B * C].view(B, C, 1, 1).bernoulli_(1.0 - p).div_(1.0 - p)
            o += B * C
        self.masks = flat.to(x.device)
        _lib.check(L.pdes_densenet_set_dropout(self.handle.h, _lib.ptr(self.masks)), "pdes_densenet_set_dropout")

    def forward(self, x, training):
        self._ensure(x)
        self.fwd_gen += 1   # the executor keeps ONE set of saved activations: see _DenseEDTrainFn.backward
        self._dropout_masks(x, training)
        x = x.contiguous()
        c = self.m._cfg
        hw = self.handle.out_size
        out = torch.empty(x.shape[0], c["out_channels"], hw, hw, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            rc = _lib.lib().pdes_densenet_forward(self.handle.h, _lib.ptr(x), _lib.ptr(out), x.shape[0],
                                                  int(bool(training)), _lib.stream_ptr())
        _lib.check(rc, "pdes_densenet_forward")
        return out

    def backward(self, dout, want_dx=False):
        dout = dout.contiguous()
        dx = None
        with torch.cuda.device(dout.device):
            if want_dx:
                c = self.m._cfg
                dx = torch.empty(dout.shape[0], c["in_channels"], self.handle.imsize, self.handle.imsize,
                                 dtype=torch.float32, device=dout.device)
                rc = _lib.lib().pdes_densenet_backward_dx(self.handle.h, _lib.ptr(dout), _lib.ptr(dx), _lib.stream_ptr())
            else:
                rc = _lib.lib().pdes_densenet_backward(self.handle.h, _lib.ptr(dout), _lib.stream_ptr())
        _lib.check(rc, "pdes_densenet_backward")
        return dx

    def flops(self, B, training):
        if self.handle is None:
            self.handle = _NetHandle(self.m._cfg, max(B, 1))
            self.ws, self.bound = None, None
        return float(_lib.lib().pdes_densenet_flops(self.handle.h, B, int(training)))


_executor_factory = CudaExecutor


class _DenseEDTrainFn(torch.autograd.Function):
    """Training-mode forward with the executor's own backward.  Parameter gradients are
    accumulated by the kernels directly into the module's flat gradient buffer (of which every
    p.grad is a view), so the function returns no tensors for them."""

    @staticmethod
    def forward(ctx, x, anchor, module):
        ctx.module = module
        out = module._ex.forward(x, True)
        ctx.fwd_gen = getattr(module._ex, "fwd_gen", None)
        return out

    @staticmethod
    def backward(ctx, dout):
        m = ctx.module
        want_dx = bool(ctx.needs_input_grad[0])
        if want_dx and m._cfg.get("arch", 0) != 2:
            raise NotImplementedError("pde_surrogate_b200.DenseED: gradient w.r.t. the network input is "
                                      "not implemented (the training path never needs it)")
        if ctx.fwd_gen != getattr(m._ex, "fwd_gen", None):
            raise RuntimeError("pde_surrogate_b200.DenseED: another forward pass (training or evaluation) ran "
                               "between this output's forward and its backward; the executor keeps the saved "
                               "activations of the LAST forward only - call backward() before the next forward")
        m._prepare_grads()
        dx = m._ex.backward(dout, want_dx) if want_dx else m._ex.backward(dout)
        return dx, None, None


class _EvalGuardFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, out, anchor):
        return out.view_as(out)

    @staticmethod
    def backward(ctx, g):
        raise NotImplementedError("pde_surrogate_b200.DenseED: backward through eval-mode BatchNorm is not "
                                  "implemented; call model.train() or use torch.no_grad()")


def activation(name):
    """models/codec.py:190-204."""
    if name in ['tanh', 'Tanh']:
        return nn.Tanh()
    elif name in ['relu', 'ReLU']:
        return nn.ReLU(inplace=True)
    elif name in ['lrelu', 'LReLU']:
        return nn.LeakyReLU(inplace=True)
    elif name in ['sigmoid', 'Sigmoid']:
        return nn.Sigmoid()
    elif name in ['softplus', 'Softplus']:
        return nn.Softplus(beta=4)
    else:
        raise ValueError('Unknown activation function')


def _reject_options(cls, drop_rate=0, bottleneck=False, upsample='nearest', out_activation=None):
    unsupported = []
    if drop_rate and not (0.0 <= drop_rate < 1.0):
        raise ValueError("dropout probability has to be between 0 and 1, but got {}".format(drop_rate))
    if upsample not in ('nearest', 'bilinear', None):
        unsupported.append("upsample=%r" % (upsample,))
    if unsupported:
        raise NotImplementedError("pde_surrogate_b200.%s does not implement: %s" % (cls, ", ".join(unsupported)))


class _ExecutorNet(nn.Module):
    """Parameter holder + autograd wiring shared by DenseED and Decoder: the module tree with the
    reference's names, ONE flat parameter / gradient / running-statistics buffer, the executor."""

    def _build(self, cfg):
        self._cfg = cfg
        self._out_act = None
        layout = _NetHandle(self._cfg, 1)
        self._param_table, self._n_flat = layout.params()
        self._bn_table, self._n_running = layout.bns()
        del layout
        # module tree with the reference's names; creation order == reference order so that the
        # default initialisation consumes the torch RNG identically
        leaves = OrderedDict()
        for name, _off, shape, kind in self._param_table:
            path = name.split(".")[:-1]
            key = ".".join(path)
            parent = self
            for part in path[:-1]:
                if part not in parent._modules:
                    parent.add_module(part, _Group())
                parent = parent._modules[part]
            if kind == 4:
                # Conv2dZeros.scale (glow_msc.py:252): a parameter of the GROUP `path[-1]`, not of a leaf
                if path[-1] not in parent._modules:
                    parent.add_module(path[-1], _Group())
                parent._modules[path[-1]].register_parameter("scale", nn.Parameter(torch.zeros(shape)))
                continue
            if kind == 3:
                # Conv2dZeros.conv.bias (glow_msc.py:247-251): zero-initialised, like the weight
                torch.empty(shape).uniform_(-1.0, 1.0)   # nn.Conv2d draws its bias before Conv2dZeros zeroes it
                leaves[key].register_parameter("bias", nn.Parameter(torch.zeros(shape)))
                with torch.no_grad():
                    leaves[key].weight.zero_()
                continue
            if key in leaves:
                continue
            if kind == 0:
                leaf = _ConvParams(shape[0], shape[1], shape[2])
            else:
                leaf = _BNParams(shape[0])
            parent.add_module(path[-1], leaf)
            leaves[key] = leaf
        self._flat = self._flat_grad = self._flat_running = self._flat_nbt = None
        # convolution implementation: 0 = tcgen05 on two-piece fp16 operands where supported (default),
        # 1 = CUDA-core fp32 everywhere, 3 / 4 / 5 = tensor cores for the forward / dgrad / wgrad only, 6 = none, 7 = dgrad + wgrad
        self.conv_impl = int(os.environ.get("PDES_CONV_IMPL", "0"))
        self._ex = _executor_factory(self)
        self._flatten()

    # ------------------------------------------------------------------ flat storage
    def _named_param_list(self):
        d = dict(self.named_parameters())
        return [(d[name], off, shape) for name, off, shape, _k in self._param_table]

    def _flatten(self):
        plist = self._named_param_list()
        ref = plist[0][0]
        flat = torch.zeros(self._n_flat, dtype=ref.dtype, device=ref.device)
        gflat = torch.zeros(self._n_flat, dtype=ref.dtype, device=ref.device)
        self._params, self._grad_views = [], []
        with torch.no_grad():
            for p, off, shape in plist:
                n = p.numel()
                flat[off:off + n].copy_(p.detach().reshape(-1))
                p.data = flat[off:off + n].view(shape)
                gv = gflat[off:off + n].view(shape)
                if p.grad is not None:
                    gv.copy_(p.grad)
                    p.grad = gv
                self._params.append(p)
                self._grad_views.append(gv)
                _OWNERS[id(p)] = self
            mods = dict(self.named_modules())
            run = torch.zeros(self._n_running, dtype=ref.dtype, device=ref.device)
            nbt = torch.zeros(len(self._bn_table), dtype=torch.long, device=ref.device)
            for i, (name, mo, vo, ch) in enumerate(self._bn_table):
                bn = mods[name]
                run[mo:mo + ch].copy_(bn.running_mean)
                run[vo:vo + ch].copy_(bn.running_var)
                nbt[i] = bn.num_batches_tracked.to(nbt.device)
                bn._buffers["running_mean"] = run[mo:mo + ch]
                bn._buffers["running_var"] = run[vo:vo + ch]
                bn._buffers["num_batches_tracked"] = nbt[i]
        self._flat, self._flat_grad, self._flat_running, self._flat_nbt = flat, gflat, run, nbt

    def _apply(self, fn, recurse=True):
        r = super(_ExecutorNet, self)._apply(fn, recurse)
        if self._flat is not None:
            self._flatten()
        return r

    def _prepare_grads(self):
        """Make every p.grad the matching view of the flat gradient buffer before the kernels
        accumulate into it (p.grad is None after zero_grad(set_to_none=True))."""
        params, views = self._params, self._grad_views
        n_none = 0
        for p in params:
            if p.grad is None:
                n_none += 1
        if n_none == len(params):
            self._flat_grad.zero_()
            for p, v in zip(params, views):
                p.grad = v
            return
        for p, v in zip(params, views):
            g = p.grad
            if g is None:
                v.zero_()
                p.grad = v
            elif g.data_ptr() != v.data_ptr():
                v.copy_(g)
                p.grad = v

    # ------------------------------------------------------------------ nn.Module surface
    def zero_grad(self, set_to_none=True):
        """nn.Module.zero_grad semantics (train_codec_mixed_residual.py:226) without the walk over the
        module tree: the script calls it once per step right after a host synchronisation, i.e. with the
        GPU idle, so its host time (126 us for 82 parameters through named_parameters) is on the
        critical path of every step."""
        if set_to_none:
            for p in self._params:
                p.grad = None
        else:
            self._flat_grad.zero_()
            for p, v in zip(self._params, self._grad_views):
                if p.grad is not None and p.grad.data_ptr() != v.data_ptr():
                    p.grad.detach_()
                    p.grad.zero_()

    def forward(self, x):
        out = self._features(x)
        # out_activation (models/codec.py:190-204, 288-289): an elementwise module behind the last convolution;
        # not on the training path of any script (all use None) - applied with the stock torch op
        return out if self._out_act is None else self._out_act(out)

    def _features(self, x):
        anchor = self._params[0]
        grad_on = torch.is_grad_enabled() and anchor.requires_grad
        if self.training:
            if grad_on:
                out = _DenseEDTrainFn.apply(x, anchor, self)
            else:
                out = self._ex.forward(x, True)
            self._flat_nbt.add_(1)
            return out
        out = self._ex.forward(x, False)
        if grad_on:
            out = _EvalGuardFn.apply(out, anchor)
        return out

    def forward_test(self, x):
        """The reference's shape trace (models/codec.py:298-304, 365-370): prints the tensor size behind every
        top-level module of `features`, then returns the output.  The executor does not materialise per-module
        tensors on the host, so the sizes are derived from the layer table (they are what the reference prints)."""
        print('input: {}'.format(x.data.size()))
        out = self.forward(x)
        c = self._cfg
        B, h = int(x.shape[0]), int(x.shape[-1])
        arch, blocks, g = c.get("arch", 0), list(c["blocks"]), c["growth_rate"]
        ch = c["init_features"]
        lines = []
        if arch == 0:
            pad = 3 if h % 2 == 0 else 2
            h = (h + 2 * pad - 7) // 2 + 1
            lines.append(("In_conv", ch, h))
            n_enc = len(blocks) // 2
        else:
            lines.append(("conv0", ch, h))
            n_enc = 0
        for i, nl in enumerate(blocks):
            enc = i < n_enc
            ch += nl * g
            lines.append(("%sBlock%d" % ("Enc" if enc else "Dec", i + 1 if enc else i - n_enc + 1), ch, h))
            if i < len(blocks) - 1:
                ch //= 2
                h = (h + 2 - 3) // 2 + 1 if enc else 2 * h
                lines.append(("Trans%s%d" % ("Down" if enc else "Up", i + 1 if enc else i - n_enc + 1), ch, h))
        lines.append(("LastTransUp", c["out_channels"], int(out.shape[-1])))
        for name, cc, hh in lines:
            print('{}: {}'.format(name, torch.Size([B, cc, hh, hh])))
        for name, module in self.features._modules.items():
            if module is self._out_act:
                print('{}: {}'.format(name, out.data.size()))
        return out

    @property
    def model_size(self):
        return module_size(self)

    def reset_parameters(self, verbose=False):
        for module in self.modules():
            if isinstance(module, (_ConvParams, _BNParams)):
                module.reset_parameters()
                if verbose:
                    print("Reset parameters in {}".format(module))

    def _set_out_activation(self, name):
        if name is not None:
            act = activation(name)
            self.features.add_module(name, act)   # same place in the tree as the reference (codec.py:288-289)
            self._out_act = act

    def flat_parameters(self):
        """(params, grads) flat fp32 buffers; every parameter / .grad is a view into them."""
        return self._flat, self._flat_grad

    def flops(self, batch, training=True):
        """Useful 2*MAC FLOPs of one forward (or forward+backward) at this batch size."""
        return self._ex.flops(batch, training)


class DenseED(_ExecutorNet):
    """Dense convolutional encoder-decoder (reference models/codec.py:210-318).

    Args as in the reference: drop_rate, upsample in ('nearest', 'bilinear', None = transposed convolutions,
    whose last decoding does not upsample: the output is imsize/2 wide, as in the reference) and
    out_activation and bottleneck dense layers (bottleneck=True, bn_size) are implemented.
    """

    def __init__(self, in_channels, out_channels, imsize, blocks, growth_rate=16, init_features=48,
                 drop_rate=0, bn_size=8, bottleneck=False, out_activation=None, upsample='nearest'):
        super(DenseED, self).__init__()
        blocks = [int(b) for b in blocks]
        if len(blocks) > 1 and len(blocks) % 2 == 0:
            raise ValueError('length of blocks must be an odd number, but got {}'.format(len(blocks)))
        _reject_options("DenseED", drop_rate, bottleneck, upsample, out_activation)
        self._build(dict(in_channels=int(in_channels), out_channels=int(out_channels), imsize=int(imsize),
                         blocks=blocks, growth_rate=int(growth_rate), init_features=int(init_features), arch=0,
                         drop_rate=float(drop_rate or 0.0), upsample=upsample,
                         bottleneck=int(bn_size) if bottleneck else 0))
        self._set_out_activation(out_activation)
        print('# params {}, # conv layers {}'.format(*self.model_size))


class Decoder(_ExecutorNet):
    """Decoder to solve one PDE (reference models/codec.py:321-370): conv0 (3x3) on a latent
    (B, dim_latent, h, h), dense decoding blocks with nearest-upsampling transitions, last decoding;
    the output is h * 2^len(blocks) wide.  Same kernels as DenseED (solve_conv_mixed_residual.py:122)."""

    def __init__(self, dim_latent, out_channels, blocks, growth_rate=16, init_features=48, drop_rate=0.,
                 upsample='nearest', out_activation=None):
        super(Decoder, self).__init__()
        blocks = [int(b) for b in blocks]
        _reject_options("Decoder", drop_rate, False, upsample, out_activation)
        # imsize here only sizes the layout query: the latent's spatial size is taken from the input
        self._build(dict(in_channels=int(dim_latent), out_channels=int(out_channels), imsize=16, blocks=blocks,
                         growth_rate=int(growth_rate), init_features=int(init_features), arch=1,
                         drop_rate=float(drop_rate or 0.0), upsample=upsample))
        self._set_out_activation(out_activation)
