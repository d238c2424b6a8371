"""The convolution-heavy pieces of models/glow_msc.py upstream on the sm_100a executor: `_DenseCoupling`
(276-294, incl. `Conv2dZeros` 240-255) and `AffineCouplingLayer` (297-344).  The rest of MultiScaleCondGlow is
not rebuilt (SURVEY.md section 8f row 1 is partial): importing it from here raises ImportError."""
from pde_surrogate_b200.glow import AffineCouplingLayer, _DenseCoupling  # noqa: F401


def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)
    raise AttributeError("models.glow_msc.%s is not part of this backend: only _DenseCoupling and "
                         "AffineCouplingLayer (the coupling networks of the cGlow reverse-KL step) are built" % name)
