"""cGlow coupling networks on the sm_100a executor (BASELINE config 5).

Host-side mirror of the conv-heavy parts of models/glow_msc.py upstream:
  * `_DenseCoupling` (276-294): three BatchNorm -> ReLU -> conv3x3(->16) dense layers + the `reduce` head
    BatchNorm -> ReLU -> `Conv2dZeros` (240-255: zero-initialised 3x3 convolution WITH bias, times
    exp(3 * scale)) - executor architecture 2: the same fused thin-layer tcgen05 kernels as DenseED's dense
    blocks, plus the gradient w.r.t. the network input (a coupling network sits inside a flow);
  * `AffineCouplingLayer` (297-344): forward / reverse of the sigmoid-affine coupling around it (elementwise
    PyTorch ops: the flow plumbing stays PyTorch).
The rest of MultiScaleCondGlow (ActNorm, LU 1x1 convolutions, squeeze / split, Gaussian priors, the input
encoder) is NOT rebuilt here.
"""
import torch
import torch.nn as nn

from .codec import _ExecutorNet, _DenseEDTrainFn


class _DenseCoupling(_ExecutorNet):
    """out = reduce(dense block(x)); same constructor and state_dict layout as glow_msc.py:276-294."""

    def __init__(self, in_features, out_features, num_layers=3, growth_rate=16, drop_rate=0.):
        super(_DenseCoupling, self).__init__()
        if drop_rate and drop_rate > 0:
            raise NotImplementedError("pde_surrogate_b200._DenseCoupling does not implement drop_rate > 0 "
                                      "(the reference's coupling layers pass drop_rate=0, glow_msc.py:320-321)")
        self._build(dict(in_channels=int(in_features), out_channels=int(out_features), imsize=16,
                         blocks=[int(num_layers)], growth_rate=int(growth_rate), init_features=1, arch=2))

    def forward(self, x):
        anchor = self._params[0]
        if self.training and torch.is_grad_enabled() and (anchor.requires_grad or x.requires_grad):
            out = _DenseEDTrainFn.apply(x, anchor, self)
            self._flat_nbt.add_(1)
            return out
        out = self._ex.forward(x, self.training)
        if self.training:
            self._flat_nbt.add_(1)
        return out


class AffineCouplingLayer(nn.Module):
    """Affine coupling layer (glow_msc.py:297-344), coupling_net='dense'."""

    def __init__(self, in_features, cond_features, coupling_net='dense'):
        super(AffineCouplingLayer, self).__init__()
        if coupling_net != 'dense':
            raise NotImplementedError("pde_surrogate_b200.AffineCouplingLayer: coupling_net=%r (only 'dense')" % (coupling_net,))
        if in_features % 2 == 0:
            in_channels = in_features // 2 + cond_features
            out_channels = in_features
        else:
            # chunk is (2, 1) if in_features == 3
            in_channels = in_features // 2 + 1 + cond_features
            out_channels = in_features - 1
        self.coupling_nn = _DenseCoupling(in_channels, out_channels, num_layers=3, growth_rate=16, drop_rate=0.)

    def _shift_scale(self, a, cond):
        h = self.coupling_nn(torch.cat((a, cond), 1))
        return h[:, 0::2], torch.sigmoid(h[:, 1::2] + 2.)

    def forward(self, x, cond):
        x1, x2 = x.chunk(2, 1)
        shift, scale = self._shift_scale(x1, cond)
        x2 = (x2 + shift) * scale
        logdet = scale.log().view(x.shape[0], -1).sum(1)
        return torch.cat((x1, x2), 1), logdet

    def reverse(self, y, cond):
        y1, y2 = y.chunk(2, 1)
        shift, scale = self._shift_scale(y1, cond)
        y2 = y2 / scale - shift
        logdet = scale.log().view(y.shape[0], -1).sum(1)
        return torch.cat((y1, y2), 1), logdet
