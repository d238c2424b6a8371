"""`from models.codec import DenseED` (train_codec_mixed_residual.py:18) -> sm_100a executor."""
from pde_surrogate_b200.codec import DenseED, module_size  # noqa: F401
