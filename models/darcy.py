"""`from models.darcy import conv_*` (train_codec_mixed_residual.py:19-21) -> fused stencil kernel."""
from pde_surrogate_b200.darcy import (conv_boundary_condition, conv_constitutive_constraint,  # noqa: F401
                                      conv_continuity_constraint)
