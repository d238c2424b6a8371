"""Runs a few device-resident training steps of the BASELINE config (DenseED [6,8,6], 64x64,
batch 32) for profiling under ncu:

  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:pdes -c 400 --csv \
      --log-file gpurun_out/launches.csv python tools/profile_step.py --steps 3
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from models.codec import DenseED  # noqa: E402
from pde_surrogate_b200.engine import TrainStep  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--imsize", type=int, default=64)
ap.add_argument("--impl", type=int, default=0)
args = ap.parse_args()
torch.manual_seed(1)
dev = torch.device("cuda:0")
model = DenseED(1, 3, args.imsize, [6, 8, 6]).to(dev)
model.conv_impl = args.impl
ts = TrainStep(model)
g = torch.Generator(device="cpu").manual_seed(1)
K = torch.exp(0.5 * torch.randn(args.batch, 1, args.imsize, args.imsize, generator=g)).to(dev)
for i in range(args.steps):
    loss = ts.step(K)
torch.cuda.synchronize()
print("loss", float(loss), "launches/step", ts.kernel_launches)
