"""Multiscale conditional Glow around the executor-backed coupling networks (BASELINE
config 5; upstream models/glow_msc.py).

What runs where: the 3 x (flow layers) coupling networks - `_DenseCoupling` incl. their `Conv2dZeros` heads, i.e.
the BatchNorm -> ReLU -> conv3x3 dense layers that dominate the step - run on the sm_100a executor (glow.py,
architecture 2, with the gradient w.r.t. their input).  Everything in THIS file is the flow plumbing the survey
leaves in PyTorch: ActNorm, the invertible 1x1 convolutions (plain and LU-parameterised), squeeze / split, the
diagonal-Gaussian priors with their zero-initialised 3x3 heads, and the input encoder (dense blocks and
down-transitions as stock torch modules).  Module names, constructor arguments, state_dict keys, default
initialisation (numpy QR draws included) and the forward / reverse / log-determinant conventions are upstream's,
so checkpoints and `train_cglow_reverse_kl.py` see the same object.

One deliberate difference: upstream clamps the log-standard-deviation of every Gaussian IN PLACE on a view of a
convolution output (glow_msc.py:438), which current PyTorch refuses to differentiate; here the clamp is
out of place - same values, same gradient (identity inside [-10, log 5], zero outside).
"""
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .codec import module_size
from .glow import AffineCouplingLayer

_LOG_STD_MIN, _LOG_STD_MAX = -10.0, math.log(5.0)


def _exact_torch_convolutions():
    """The PyTorch parts of the flow (input encoder, 1x1 convolutions, prior heads) run cuDNN convolutions, which
    PyTorch lets use TF32 by default: measured on the B200, that alone puts 6e-4 on the conditioning features and 7e-4
    on the generated fields (tools/diag_cglow.py) - seven times the 1e-4 parity bar, with the executor's own
    convolutions exact.  The model therefore switches cuDNN TF32 off for the process (forward AND the autograd
    backward, which runs outside any scoped flag); PDES_ALLOW_TF32=1 leaves PyTorch's default."""
    import os
    if os.environ.get("PDES_ALLOW_TF32", "0") != "1":
        torch.backends.cudnn.allow_tf32 = False


# ------------------------------------------------------------------------------------------------
# elementwise / 1x1 flow steps
# ------------------------------------------------------------------------------------------------
class ActNorm(nn.Module):
    """Per-channel affine y = w * x + b (glow_msc.py:51-95): identity at construction, or initialised from the first
    minibatch when `data_init` is set.  Both directions return sum(log|w|) * H * W as the log-determinant term."""

    def __init__(self, in_features, return_logdet=True, data_init=False):
        super(ActNorm, self).__init__()
        self.weight = nn.Parameter(torch.ones(in_features, 1, 1))
        self.bias = nn.Parameter(torch.zeros(in_features, 1, 1))
        self.data_init = data_init
        self.data_initialized = False
        self.return_logdet = return_logdet

    def _init_parameters(self, input):
        flat = input.transpose(0, 1).contiguous().view(input.shape[1], -1)   # (C, B*H*W)
        mean, std = flat.mean(1), flat.std(1) + 1e-6
        self.bias.data = -(mean / std).unsqueeze(-1).unsqueeze(-1)
        self.weight.data = 1. / std.unsqueeze(-1).unsqueeze(-1)

    def _logdet(self, t):
        return self.weight.abs().log().sum() * t.shape[-1] * t.shape[-2]

    def forward(self, x):
        if self.data_init and not self.data_initialized:
            self._init_parameters(x)
            self.data_initialized = True
        y = self.weight * x + self.bias
        return (y, self._logdet(x)) if self.return_logdet else y

    def reverse(self, y):
        x = (y - self.bias) / self.weight
        return (x, self._logdet(y)) if self.return_logdet else x


def _random_rotation(n):
    """Orthogonal initialisation of the 1x1 convolutions: Q of a QR factorisation of a numpy normal draw (the same
    numpy global-RNG call as upstream, glow_msc.py:120, 179)."""
    return np.linalg.qr(np.random.randn(n, n))[0].astype(np.float32)


class InvertibleConv1x1(nn.Module):
    """Learned channel mixing (glow_msc.py:98-156).  ONE matrix serves both directions; the direction used for
    training (`train_sampling`: z -> x) applies it as it is, the other one its fp64 inverse."""

    def __init__(self, in_channels, train_sampling=True):
        super(InvertibleConv1x1, self).__init__()
        self.w_shape = (in_channels, in_channels)
        self.train_sampling = train_sampling
        self.weight = nn.Parameter(torch.Tensor(_random_rotation(in_channels)))

    def _inverse(self):
        return torch.inverse(self.weight.double()).float()

    def log_determinant(self, x, W):
        det = torch.det(W.to(torch.float64)).to(torch.float32)
        if det.item() == 0:
            det += 1e-6
        return x.shape[2] * x.shape[3] * det.abs().log()

    def _apply_matrix(self, t, W):
        return F.conv2d(t, W.view(*self.w_shape, 1, 1)), self.log_determinant(t, W)

    def forward(self, x):
        return self._apply_matrix(x, self._inverse() if self.train_sampling else self.weight)

    def reverse(self, z):
        out, logdet = self._apply_matrix(z, self.weight if self.train_sampling else self._inverse())
        return out, -logdet   # (upstream's convention: the reverse pass reports minus the determinant of what it applied)


class InvertibleConv1x1LU(nn.Module):
    """The same mixing with W = P L (U + diag(sign_s * exp(log_s))) (glow_msc.py:159-236): the log-determinant is
    sum(log_s) * H * W without any factorisation at run time."""

    def __init__(self, in_channels, train_sampling=True):
        super(InvertibleConv1x1LU, self).__init__()
        import scipy.linalg
        self.w_shape = (in_channels, in_channels)
        self.train_sampling = train_sampling
        w0 = _random_rotation(in_channels)
        p, lower, upper = scipy.linalg.lu(w0)
        s = np.diag(upper)
        f32 = lambda a: torch.Tensor(np.asarray(a, dtype=np.float32))   # noqa: E731
        self.register_buffer('p', f32(p))
        self.l = nn.Parameter(f32(lower))
        self.u = nn.Parameter(f32(np.triu(upper, k=1)))
        self.log_s = nn.Parameter(f32(np.log(np.abs(s))))
        self.register_buffer('sign_s', f32(np.sign(s)))
        self.register_buffer('l_mask', f32(np.tril(np.ones_like(w0), -1)))
        self.register_buffer('u_mask', f32(np.triu(np.ones_like(w0), k=1)))
        self.register_buffer('eye', f32(np.eye(in_channels)))

    def _factors(self):
        lower = self.l * self.l_mask + self.eye
        upper = self.u * self.u_mask + torch.diag(self.log_s.exp() * self.sign_s)
        return lower, upper

    def weight(self):
        lower, upper = self._factors()
        return torch.matmul(self.p, torch.matmul(lower, upper))

    def inv_weight(self):
        lower, upper = self._factors()
        return torch.matmul(upper.inverse(), torch.matmul(lower.inverse(), self.p.inverse()))

    def _run(self, t, use_inverse):
        logdet = self.log_s.sum() * t.shape[2] * t.shape[3]
        w = self.inv_weight() if use_inverse else self.weight()
        # the sign follows the matrix that is applied relative to the training direction (upstream 213-236)
        return F.conv2d(t, w.view(*self.w_shape, 1, 1)), (-logdet if self.train_sampling else logdet)

    def forward(self, x):
        return self._run(x, use_inverse=self.train_sampling)

    def reverse(self, x):
        return self._run(x, use_inverse=not self.train_sampling)


class Conv2dZeros(nn.Module):
    """Zero-initialised 3x3 convolution with bias and a learned per-channel gain exp(3 * scale) (glow_msc.py:240-255);
    the stand-alone torch form used by the priors.  (Inside the coupling networks the same head runs on the executor.)"""

    def __init__(self, in_channels, out_channels):
        super(Conv2dZeros, self).__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size=3, stride=1, padding=1, bias=True)
        self.conv.weight.data.zero_()
        self.conv.bias.data.zero_()
        self.scale = nn.Parameter(torch.zeros(1, out_channels, 1, 1))

    def forward(self, x):
        return self.conv(x) * torch.exp(self.scale * 3)


class Squeeze(nn.Module):
    """(B, C, H, W) <-> (B, C f^2, H/f, W/f) with upstream's element order (glow_msc.py:400-429)."""

    def __init__(self, factor=2):
        super(Squeeze, self).__init__()
        assert factor >= 1
        self.factor = factor

    def forward(self, x):
        f = self.factor
        if f == 1:
            return x
        C, H, W = x.shape[1:]
        assert H % f == 0 and W % f == 0
        x = x.reshape(-1, C, f, H // f, f, W // f).transpose(3, 4)
        return x.reshape(-1, C * f * f, H // f, W // f)

    def reverse(self, x):
        f = self.factor
        if f == 1:
            return x
        C, H, W = x.shape[1:]
        assert C >= f * f and C % (f * f) == 0
        x = x.reshape(-1, C // (f * f), f, f, H, W).transpose(3, 4)
        return x.reshape(-1, C // (f * f), H * f, W * f)


class GaussianDiag(object):
    """Diagonal Gaussian with a clamped log-standard-deviation (glow_msc.py:432-456; out-of-place clamp, see the
    module docstring)."""
    Log2PI = float(np.log(2 * np.pi))

    def __init__(self, mean, log_stddev):
        self.mean = mean
        self.log_stddev = log_stddev.clamp(min=_LOG_STD_MIN, max=_LOG_STD_MAX)

    def likelihood(self, x):
        return -0.5 * (GaussianDiag.Log2PI + self.log_stddev * 2. + (x - self.mean) ** 2 / (self.log_stddev * 2.).exp())

    def log_prob(self, x):
        return self.likelihood(x).view(x.shape[0], -1).sum(1)

    def sample(self, eps=None):
        if eps is None:
            eps = torch.randn_like(self.log_stddev)
        return self.mean + self.log_stddev.exp() * eps


class LatentEncoder(nn.Module):
    """Prior of a factored-out half given the half that stays: (mean, log_stddev) = Conv2dZeros(z1) (459-471)."""

    def __init__(self, in_channels):
        super(LatentEncoder, self).__init__()
        self.conv2d = Conv2dZeros(in_channels, in_channels * 2)

    def forward(self, x):
        mean, log_stddev = self.conv2d(x).chunk(2, 1)
        return GaussianDiag(mean, log_stddev)


# ------------------------------------------------------------------------------------------------
# input encoder: dense blocks / down-transitions as stock torch modules (upstream imports them from models.codec)
# ------------------------------------------------------------------------------------------------
def _bn_relu_conv(seq, idx, cin, cout, k, stride, pad):
    seq.add_module('norm%d' % idx, nn.BatchNorm2d(cin))
    seq.add_module('relu%d' % idx, nn.ReLU(inplace=True))
    seq.add_module('conv%d' % idx, nn.Conv2d(cin, cout, kernel_size=k, stride=stride, padding=pad, bias=False))


class _EncDenseLayer(nn.Sequential):
    """codec._DenseLayer without bottleneck / dropout (the encoder's blocks pass neither; codec.py:65-75)."""

    def __init__(self, in_features, growth_rate):
        super(_EncDenseLayer, self).__init__()
        _bn_relu_conv(self, 1, in_features, growth_rate, 3, 1, 1)

    def forward(self, x):
        return torch.cat([x, super(_EncDenseLayer, self).forward(x)], 1)


class _EncDenseBlock(nn.Sequential):
    def __init__(self, num_layers, in_features, growth_rate):
        super(_EncDenseBlock, self).__init__()
        for i in range(num_layers):
            self.add_module('denselayer%d' % (i + 1), _EncDenseLayer(in_features + i * growth_rate, growth_rate))


class _DenseBlockInput(nn.Sequential):
    """First encoder block (glow_msc.py:28-48): a biased 3x3 `in_conv` to init_features - 1 channels concatenated
    with the input itself, then num_layers - 1 dense layers."""

    def __init__(self, num_layers, in_features, init_features, growth_rate, drop_rate=0., bn_size=4, bottleneck=False):
        super(_DenseBlockInput, self).__init__()
        self.num_layers = num_layers
        self.add_module('in_conv', nn.Conv2d(in_features, init_features - 1, kernel_size=3, stride=1, padding=1))
        for i in range(num_layers - 1):
            self.add_module('denselayer%d' % (i + 1), _EncDenseLayer(init_features + i * growth_rate, growth_rate))

    def forward(self, x):
        out = torch.cat((x, self.in_conv(x)), 1)
        for i in range(self.num_layers - 1):
            out = self[i + 1](out)
        return out


class _EncTransitionDown(nn.Sequential):
    """codec._Transition(down=True) (codec.py:103-125): BN-ReLU-conv1x1 + BN-ReLU-conv3x3/s2 with `bottleneck`,
    one BN-ReLU-conv3x3/s2 without."""

    def __init__(self, in_features, out_features, bottleneck):
        super(_EncTransitionDown, self).__init__()
        if bottleneck:
            _bn_relu_conv(self, 1, in_features, out_features, 1, 1, 0)
            _bn_relu_conv(self, 2, out_features, out_features, 3, 2, 1)
        else:
            _bn_relu_conv(self, 1, in_features, out_features, 3, 2, 1)


class InputEncoder(nn.Sequential):
    """x -> multiscale conditioning features + the conditional prior of the top latent (glow_msc.py:474-552)."""

    def __init__(self, in_channels, latent_features, blocks, growth_rate=16, init_features=48, drop_rate=0.):
        super(InputEncoder, self).__init__()
        if drop_rate and drop_rate > 0:
            raise NotImplementedError("pde_surrogate_b200.InputEncoder: drop_rate > 0 (upstream passes 0, glow_msc.py:707)")
        self.num_blocks = len(blocks)
        num_features = in_channels
        for i, num_layers in enumerate(blocks):
            if i == 0:
                block = _DenseBlockInput(num_layers, in_channels, init_features, growth_rate)
                num_features = init_features + (num_layers - 1) * growth_rate
            else:
                block = _EncDenseBlock(num_layers, num_features, growth_rate)
                num_features = num_features + num_layers * growth_rate
            self.add_module('dense_block%d' % (i + 1), block)
            if i < len(blocks) - 1:
                self.add_module('trans_down%d' % (i + 1), _EncTransitionDown(num_features, num_features // 2, bottleneck=i > 0))
                num_features = num_features // 2
        self.add_module('top_latent', Conv2dZeros(num_features, latent_features * 2))

    def forward(self, x):
        conditions = []
        for i in range(self.num_blocks):
            x = self[2 * i](x)
            conditions.append(x)
            x = self[2 * i + 1](x)   # down-transition, or the top-latent head behind the last block
        mean, log_stddev = x.chunk(2, 1)
        # upstream clamps `log_stddev.data`: the prior's spread carries no gradient back into the encoder (line 533)
        return conditions, GaussianDiag(mean, log_stddev.detach())

    def feature_sizes(self, x):
        sizes = []
        for i in range(self.num_blocks):
            x = self[2 * i](x)
            sizes.append(x.shape[1:])
            if i < self.num_blocks - 1:
                x = self[2 * i + 1](x)
        return sizes


# ------------------------------------------------------------------------------------------------
# reversible layers and blocks
# ------------------------------------------------------------------------------------------------
class RevLayer(nn.Module):
    """ActNorm -> invertible 1x1 convolution -> affine coupling (glow_msc.py:348-377)."""

    def __init__(self, in_features, cond_features, LUdecompose=False, train_sampling=True, coupling_net='dense'):
        super(RevLayer, self).__init__()
        self.norm = ActNorm(in_features)
        mixer = InvertibleConv1x1LU if LUdecompose else InvertibleConv1x1
        self.conv1x1 = mixer(in_features, train_sampling=train_sampling)
        self.coupling = AffineCouplingLayer(in_features, cond_features, coupling_net=coupling_net)

    def forward(self, x, cond):
        x, a = self.norm(x)
        x, b = self.conv1x1(x)
        x, c = self.coupling(x, cond)
        return x, a + b + c

    def reverse(self, y, cond):
        y, a = self.coupling.reverse(y, cond)
        y, b = self.conv1x1.reverse(y)
        y, c = self.norm.reverse(y)
        return y, a + b + c


class FirstRevLayer(nn.Module):
    """The layer next to the data: the coupling alone (glow_msc.py:380-397)."""

    def __init__(self, in_features, cond_features, coupling_net='dense'):
        super(FirstRevLayer, self).__init__()
        self.coupling = AffineCouplingLayer(in_features, cond_features, coupling_net=coupling_net)

    def forward(self, x, cond):
        return self.coupling(x, cond)

    def reverse(self, y, cond):
        return self.coupling.reverse(y, cond)


class Split(nn.Module):
    """Factor out half of the channels behind a block; their prior is conditioned on the half that stays (554-582)."""

    def __init__(self, in_features):
        super(Split, self).__init__()
        self.latent_encoder = LatentEncoder(in_features // 2)

    def forward(self, z, return_eps=False):
        z, z2 = z.chunk(2, 1)
        prior = self.latent_encoder(z)
        eps = (z2 - prior.mean) / prior.log_stddev.exp() if return_eps else None
        return z, prior.log_prob(z2), eps

    def reverse(self, z1, eps=None):
        prior = self.latent_encoder(z1)
        z2 = prior.sample(eps)
        return torch.cat((z1, z2), 1), prior.log_prob(z2)


def _rev_stack(first, in_features, cond_features, n_layers, coupling_net, LUdecompose, train_sampling):
    layers = nn.Sequential()
    for i in range(n_layers):
        if first and i == 0:
            layer = FirstRevLayer(in_features, cond_features)
        else:
            layer = RevLayer(in_features, cond_features, LUdecompose=LUdecompose, train_sampling=train_sampling,
                             coupling_net=coupling_net)
        layers.add_module('revlayer%d' % (i + 1), layer)
    return layers


class RevBlock(nn.Module):
    """Squeeze -> RevLayers -> Split (no split in front of the top latent) (glow_msc.py:585-633)."""

    def __init__(self, in_features, cond_features, n_layers, coupling_net='dense', factor=2, LUdecompose=False,
                 train_sampling=True, do_split=True):
        super(RevBlock, self).__init__()
        self.do_split = do_split
        self.squeeze = Squeeze(factor)
        in_features = in_features * factor ** 2
        self.revlayers = _rev_stack(False, in_features, cond_features, n_layers, coupling_net, LUdecompose, train_sampling)
        if do_split:
            self.split = Split(in_features)

    def forward(self, x, cond, return_eps=False):
        x = self.squeeze(x)
        logdet = 0.
        for layer in self.revlayers._modules.values():
            x, d = layer(x, cond)
            logdet = logdet + d
        if not self.do_split:
            return x, logdet, None
        x, log_prob_prior, eps = self.split(x, return_eps=return_eps)
        return x, logdet + log_prob_prior, eps

    def reverse(self, y, cond, eps):
        logdet = 0.
        if self.do_split:
            y, log_prob_prior = self.split.reverse(y, eps)
            logdet = logdet + log_prob_prior
        for layer in reversed(self.revlayers._modules.values()):
            y, d = layer.reverse(y, cond)
            logdet = logdet + d
        return self.squeeze.reverse(y), logdet


class FirstRevBlock(nn.Module):
    """The block at the data resolution: no squeeze, no split, a bare coupling first (glow_msc.py:636-669)."""

    def __init__(self, in_features, cond_features, n_layers, coupling_net='dense', LUdecompose=False, train_sampling=True):
        super(FirstRevBlock, self).__init__()
        self.revlayers = _rev_stack(True, in_features, cond_features, n_layers, coupling_net, LUdecompose, train_sampling)

    def forward(self, x, cond):
        logdet = 0.
        for layer in self.revlayers._modules.values():
            x, d = layer(x, cond)
            logdet = logdet + d
        return x, logdet

    def reverse(self, y, cond):
        logdet = 0.
        for layer in reversed(self.revlayers._modules.values()):
            y, d = layer.reverse(y, cond)
            logdet = logdet + d
        return y, logdet


# ------------------------------------------------------------------------------------------------
# the model
# ------------------------------------------------------------------------------------------------
class MultiScaleCondGlow(nn.Module):
    """p(y | x) as a multiscale conditional flow (glow_msc.py:672-966): `forward` encodes y -> z and evaluates
    log p(y | x); `generate` draws one y per x through the reverse pass - the path `train_cglow_reverse_kl.py`
    trains (250-273); `sample` / `predict` / `propagate` are the Monte-Carlo summaries built on it."""

    def __init__(self, img_size, x_channels, y_channels, enc_blocks, flow_blocks, flow_coupling='dense', squeeze_factor=2,
                 LUdecompose=False, train_sampling=True, data_init=False):
        super(MultiScaleCondGlow, self).__init__()
        _exact_torch_convolutions()
        if isinstance(img_size, int):
            self.img_size = [img_size, img_size]
        else:
            assert isinstance(img_size, (list, tuple)) and len(img_size) == 2, 'Images, 2D!'
            self.img_size = list(img_size)
        self.data_init = data_init
        self.data_initialized = False
        self.y_channels = y_channels
        self.flow_blocks = flow_blocks
        self.factor = squeeze_factor
        top_features = self._z_shapes()[-1][0]
        self.encoder = InputEncoder(x_channels, top_features, enc_blocks, growth_rate=16, init_features=48, drop_rate=0.)
        with torch.no_grad():
            cond_sizes = self.encoder.feature_sizes(torch.randn(1, x_channels, *self.img_size))
        self.flow = nn.Sequential()
        n_features = y_channels
        for i, n_layers in enumerate(flow_blocks):
            if i == 0:
                block = FirstRevBlock(n_features, cond_sizes[i][0], n_layers, coupling_net=flow_coupling,
                                      LUdecompose=LUdecompose, train_sampling=train_sampling)
            else:
                block = RevBlock(n_features, cond_sizes[i][0], n_layers, coupling_net=flow_coupling, factor=squeeze_factor,
                                 LUdecompose=LUdecompose, train_sampling=train_sampling,
                                 do_split=i < len(flow_blocks) - 1)
                n_features = n_features * (squeeze_factor ** 2) // 2
            self.flow.add_module('revblock%d' % (i + 1), block)
        if self.data_init:
            for module in self.modules():
                if isinstance(module, ActNorm):
                    module.data_init = True

    @property
    def device(self):
        return next(self.parameters()).device

    @property
    def model_size(self):
        return module_size(self)

    # ---- y -> z -----------------------------------------------------------------------------------
    def forward(self, y, x, return_eps=False):
        conditions, cond_prior = self.encoder(x)
        logdet = 0.
        eps_list = []
        last = len(self.flow_blocks) - 1
        for i, block in enumerate(self.flow._modules.values()):
            if i == 0:
                y, d = block(y, conditions[i])
            elif i == last:
                y, d, _ = block(y, conditions[i])
                logdet = logdet + cond_prior.log_prob(y)
                if return_eps:
                    eps_list.append((y - cond_prior.mean) / cond_prior.log_stddev.exp())
            else:
                y, d, eps = block(y, conditions[i], return_eps=return_eps)
                if return_eps:
                    eps_list.append(eps)
            logdet = logdet + d
        return y, logdet, (eps_list if return_eps else None)

    # ---- z -> y -----------------------------------------------------------------------------------
    def generate(self, x, eps_list=None):
        n_latent = len(self.flow_blocks) - 1
        if eps_list is not None:
            assert len(eps_list) == n_latent, 'The specified noise must have the same size as the latent variables'
        else:
            eps_list = [None] * n_latent
        eps_list = [None] + list(eps_list)   # the block at the data resolution has no latent of its own
        conditions, cond_prior = self.encoder(x)
        z = cond_prior.sample(eps_list[-1])
        logp = cond_prior.log_prob(z)
        blocks = list(self.flow._modules.values())
        for i in range(len(blocks) - 1, -1, -1):
            if i == 0:
                z, d = blocks[i].reverse(z, conditions[i])
            else:
                z, d = blocks[i].reverse(z, conditions[i], eps_list[i])
            logp = logp + d
        return z, logp

    def approx_pred_mean(self, x):
        return self.generate(x, eps_list=self.create_zero_noise(batch_size=x.shape[0]))

    def sample(self, x, n_samples, eps_list=None, temperature=None):
        if temperature is None:
            temperature = 0.7
        if eps_list is not None:
            assert n_samples == eps_list[-1].shape[0] and x.shape[0] == eps_list[-1].shape[1]
        else:
            eps_list = self.create_fixed_noise(n_samples, batch_size=x.shape[0])
        eps_list = [None] + list(eps_list)
        conditions, cond_prior = self.encoder(x)
        blocks = list(self.flow._modules.values())
        ys = []
        for s in range(n_samples):
            z = cond_prior.sample(eps_list[-1][s])
            for i in range(len(blocks) - 1, -1, -1):
                if eps_list[i] is None:
                    z, _ = blocks[i].reverse(z, conditions[i])
                else:
                    z, _ = blocks[i].reverse(z, conditions[i], eps_list[i][s] * temperature)
            ys.append(z)
        return torch.stack(ys, 0)

    def _z_shapes(self):
        size = list(self.img_size)
        n = self.y_channels
        shapes = []
        for _ in range(len(self.flow_blocks) - 2):
            size = [v // 2 for v in size]
            n = n * self.factor ** 2 // 2
            shapes.append((n, *size))
        size = [v // 2 for v in size]
        shapes.append((n * self.factor ** 2, *size))   # the top latent is not factored out
        return shapes

    def create_fixed_noise(self, n_samples, batch_size=1):
        return [torch.randn(n_samples, batch_size, *s).to(self.device) for s in self._z_shapes()]

    def create_zero_noise(self, batch_size):
        return [torch.zeros(batch_size, *s).to(self.device) for s in self._z_shapes()]

    def init_actnorm(self):
        for module in self.modules():
            if isinstance(module, ActNorm):
                module.data_initialized = True
        self.data_initialized = True

    def predict(self, x_test, n_samples=20, temperature=1.0):
        pred = self.sample(x_test, n_samples, temperature=temperature)
        return pred.mean(0), pred.var(0)

    def propagate(self, mc_loader, n_samples=20, temperature=1.0, var_samples=10):
        """E[Y] = E_X E[Y|X], Var[Y] = E_X Var(Y|X) + Var_X E[Y|X], each estimated `var_samples` times (939-966)."""
        out_shape = mc_loader.dataset[0][1].shape
        Ey = torch.zeros(var_samples, *out_shape, device=self.device)
        Eyy = torch.zeros_like(Ey)
        for i in range(var_samples):
            print(f'propagating for the {i}-th time...')
            for x_mc, _ in mc_loader:
                y = self.sample(x_mc.to(self.device), n_samples=n_samples, temperature=temperature)
                Ey[i] += y.mean(0).mean(0)
                Eyy[i] += y.pow(2).mean(0).mean(0)
        Ey /= len(mc_loader)
        Eyy /= len(mc_loader)
        Vy = Eyy - Ey ** 2
        return Ey.mean(0), Ey.var(0), Vy.mean(0), Vy.var(0)
