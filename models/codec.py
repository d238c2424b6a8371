"""`from models.codec import DenseED` (train_codec_mixed_residual.py:18) and `from models.codec import Decoder`
(solve_conv_mixed_residual.py:19) -> sm_100a executor."""
from pde_surrogate_b200.codec import DenseED, Decoder, activation, module_size  # noqa: F401
