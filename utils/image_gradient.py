"""`from utils.image_gradient import SobelFilter` (train_codec_mixed_residual.py:22)."""
from pde_surrogate_b200.image_gradient import SobelFilter  # noqa: F401
