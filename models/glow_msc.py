"""`from models.glow_msc import MultiScaleCondGlow` (train_cglow_reverse_kl.py:19): the multiscale conditional Glow
with its coupling networks (`_DenseCoupling` incl. `Conv2dZeros`, `AffineCouplingLayer`; glow_msc.py:240-344) on the
sm_100a executor and the flow plumbing in PyTorch (pde_surrogate_b200/glow_flow.py).  The 'wide' coupling network
(`_CouplingNN`) is not built: asking for it raises."""
from pde_surrogate_b200.glow import AffineCouplingLayer, _DenseCoupling  # noqa: F401
from pde_surrogate_b200.glow_flow import (ActNorm, Conv2dZeros, FirstRevBlock, FirstRevLayer, GaussianDiag,  # noqa: F401
                                          InputEncoder, InvertibleConv1x1, InvertibleConv1x1LU, LatentEncoder,
                                          MultiScaleCondGlow, RevBlock, RevLayer, Split, Squeeze, _DenseBlockInput)


def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)
    raise AttributeError("models.glow_msc.%s is not part of this backend (the 'wide' coupling network _CouplingNN "
                         "is not built)" % name)
