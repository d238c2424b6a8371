"""Stand-in for matplotlib on boxes where it is not installed: the reference training script
only calls `matplotlib.pyplot.switch_backend('agg')` at import time
(train_codec_mixed_residual.py:33-34).  Put on sys.path by run_reference_script.py ONLY when
the real package is missing."""
__pdes_shim__ = True
