#!/usr/bin/env python
"""Run an UNMODIFIED reference training script (default: train_codec_mixed_residual.py) against
this repo's backend:

    python run_reference_script.py [--script /path/to/train_codec_mixed_residual.py] -- \
        --data-dir ./datasets --imsize 64 --ntrain 4096 --batch-size 32 --epochs 1 --cuda 0

The repo root is put FIRST on sys.path so that `models.codec`, `models.darcy`,
`utils.image_gradient`, `utils.load`, ... resolve to this repo (the script's own directory would
otherwise shadow them), stand-ins for matplotlib / h5py are added only if the real packages are
missing, `torch.optim.Adam` is pointed at the fused subclass (pde_surrogate_b200/optim.py), and the script is
executed with runpy as __main__.
"""
import importlib.util
import os
import runpy
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))


def main(argv):
    script = os.path.join(os.environ.get("PDES_REFERENCE", "/root/reference"), "train_codec_mixed_residual.py")
    args = list(argv)
    if args and args[0] == "--script":
        script = args[1]
        args = args[2:]
    if args and args[0] == "--":
        args = args[1:]
    if not os.path.exists(script):
        raise SystemExit("reference script not found: %s (pass --script PATH)" % script)
    shims = os.path.join(ROOT, "pde_surrogate_b200", "_shims")
    for mod in ("matplotlib", "h5py"):
        if importlib.util.find_spec(mod) is None and shims not in sys.path:
            sys.path.insert(0, shims)
    # runpy puts the script's directory at sys.path[0]; the repo must win over it
    for m in [k for k in sys.modules if k.split(".")[0] in ("models", "utils")]:
        del sys.modules[m]
    sys.argv = [script] + args
    code_dir = os.path.dirname(os.path.abspath(script))

    class _Front(list):
        pass
    sys.path.insert(0, ROOT)
    import models.codec  # noqa: F401  (bind the repo's packages before the script's directory is added)
    import models.darcy  # noqa: F401
    import utils.image_gradient, utils.load, utils.misc, utils.plot, utils.practices  # noqa: F401,E401
    if code_dir in sys.path:
        sys.path.remove(code_dir)
    # optim.Adam(model.parameters(), ...) in the script resolves to the fused subclass (one launch per step when the
    # parameters are an executor network's; the stock torch step otherwise; PDES_FUSED_ADAM=0 keeps torch's class)
    from pde_surrogate_b200 import optim as _optim
    _optim.install()
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main(sys.argv[1:])
