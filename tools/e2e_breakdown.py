"""Host/device timeline of the e2e step (the unmodified script's loop body through the reference-facing
modules): host time stamps and CUDA events at every stage boundary, medians over the timed steps."""
import os
import statistics
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from models.codec import DenseED  # noqa: E402
from models.darcy import conv_boundary_condition, conv_constitutive_constraint, conv_continuity_constraint  # noqa: E402
from utils.image_gradient import SobelFilter  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(1)
model = DenseED(1, 3, 64, [6, 8, 6]).to(dev)
from pde_surrogate_b200 import optim as pdes_optim  # noqa: E402
pdes_optim.install()   # what run_reference_script.py does (PDES_FUSED_ADAM=0: torch's own step)
opt = torch.optim.Adam(model.parameters(), lr=1e-3)
sob = SobelFilter(64, correct=True, device=dev)
host = torch.exp(0.5 * torch.randn(4096, 1, 64, 64)).pin_memory()
names = ["h2d", "zero_grad", "forward", "losses", "backward", "adam", "item"]
H = {n: [] for n in names}
G = {n: [] for n in names}
tot_h, tot_g = [], []
_side = torch.cuda.Stream() if os.environ.get("E2E_STREAM") else None   # E2E_STREAM=1: the loop on a non-default stream
if _side is not None:
    torch.cuda.set_stream(_side)
for it in range(60):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
    t = [time.perf_counter()]
    ev[0].record()
    inp = host[(it % 128) * 32:(it % 128 + 1) * 32].to(dev, non_blocking=True)
    t.append(time.perf_counter()); ev[1].record()
    model.zero_grad()
    t.append(time.perf_counter()); ev[2].record()
    out = model(inp)
    t.append(time.perf_counter()); ev[3].record()
    loss_pde = conv_constitutive_constraint(inp, out, sob) + conv_continuity_constraint(out, sob)
    l_dir, l_neu = conv_boundary_condition(out)
    loss = loss_pde + (l_dir + l_neu) * 10.0
    t.append(time.perf_counter()); ev[4].record()
    loss.backward()
    t.append(time.perf_counter()); ev[5].record()
    opt.step()
    t.append(time.perf_counter()); ev[6].record()
    v = loss.item()
    t.append(time.perf_counter()); ev[7].record()
    torch.cuda.synchronize()
    if it >= 10:
        for i, n in enumerate(names):
            H[n].append((t[i + 1] - t[i]) * 1e6)
            G[n].append(ev[i].elapsed_time(ev[i + 1]) * 1e3)
        tot_h.append((t[-1] - t[0]) * 1e6)
        tot_g.append(ev[0].elapsed_time(ev[7]) * 1e3)
print("%-10s %10s %10s" % ("stage", "host us", "event us"))
for n in names:
    print("%-10s %10.1f %10.1f" % (n, statistics.median(H[n]), statistics.median(G[n])))
print("%-10s %10.1f %10.1f" % ("total", statistics.median(tot_h), statistics.median(tot_g)))
