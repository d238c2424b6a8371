"""pde_surrogate_b200 — B200-native backend for the physics-constrained DenseED training hot path
of cics-nd/pde-surrogate (DenseED forward -> Darcy mixed-residual loss via Sobel stencils ->
backward), behind the reference's own Python surface.

Public surface (same names as the reference modules they replace):
    pde_surrogate_b200.codec.DenseED                      <- models/codec.py:210-318
    pde_surrogate_b200.darcy.conv_constitutive_constraint <- models/darcy.py:162-176
    pde_surrogate_b200.darcy.conv_continuity_constraint   <- models/darcy.py:210-224
    pde_surrogate_b200.darcy.conv_boundary_condition      <- models/darcy.py:226-233
    pde_surrogate_b200.image_gradient.SobelFilter         <- utils/image_gradient.py:24-92
The top-level `models/` and `utils/` packages of this repo re-export them under the reference's
import paths so that train_codec_mixed_residual.py runs unmodified.
"""
__version__ = "0.1.0"
