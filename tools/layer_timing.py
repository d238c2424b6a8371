"""In-situ per-launch CUDA-event times of one eager training step (BASELINE config): warm caches,
real data dependencies — unlike an ncu launch list, which is cold-cache and serialised.

    python tools/layer_timing.py [--steps 4] 2> gpurun_out/layer_timing.txt
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from models.codec import DenseED  # noqa: E402
from pde_surrogate_b200 import _lib  # noqa: E402
from pde_surrogate_b200.engine import TrainStep  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=4)
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--imsize", type=int, default=64)
args = ap.parse_args()
torch.manual_seed(1)
dev = torch.device("cuda:0")
model = DenseED(1, 3, args.imsize, [6, 8, 6]).to(dev)
ts = TrainStep(model)
g = torch.Generator(device="cpu").manual_seed(1)
K = torch.exp(0.5 * torch.randn(args.batch, 1, args.imsize, args.imsize, generator=g)).to(dev)
L = _lib.lib()
for i in range(args.steps):
    if i == args.steps - 1:
        h = model._ex.handle.h
        _lib.check(L.pdes_densenet_set_timing(h, 1))
    loss = ts.step(K)
_lib.check(L.pdes_densenet_timing_report(h))
_lib.check(L.pdes_densenet_set_timing(h, 0))
print("loss", float(loss), "launches/step", ts.kernel_launches)
