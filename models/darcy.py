"""`from models.darcy import conv_*` (train_codec_mixed_residual.py:19-21, solve_conv_mixed_residual.py:21-22,
75, 79) -> fused stencil kernels."""
from pde_surrogate_b200.darcy import (conv_boundary_condition, conv_constitutive_constraint,  # noqa: F401
                                      conv_constitutive_constraint_nonlinear,
                                      conv_constitutive_constraint_nonlinear_exp, conv_continuity_constraint,
                                      energy_functional_exp)
