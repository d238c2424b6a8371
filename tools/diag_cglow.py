"""Where does the GPU cGlow step leave the CPU stand-in?  Per coupling network (in call order of `generate`):
input shape and rel-L2 of its output, GPU executor (conv_impl 0 and 1) vs the oracle-backed CPU executor."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from tests._cpu_backend import cpu_backend
from tests.test_glow_flow import load_cglow_fixture
from pde_surrogate_b200.glow import _DenseCoupling

gd = os.path.join(ROOT, "tests", "golden")

def run(device, impl=None):
    model, x, eps, g = load_cglow_fixture(gd, device=device)
    rec = []
    for name, m in model.named_modules():
        if isinstance(m, _DenseCoupling):
            if impl is not None:
                m.conv_impl = impl
            m.register_forward_hook(lambda mod, inp, out, name=name: rec.append((name, tuple(inp[0].shape), inp[0].detach().cpu().double(), out.detach().cpu().double())))
    model.train()
    with torch.no_grad():
        y, logp = model.generate(x, eps_list=eps)
    return rec, y.detach().cpu().double(), g

with cpu_backend():
    ref, y_ref, g = run("cpu")
print("cpu stand-in vs fixture y: %.2e" % float((y_ref - torch.tensor(g["y64"]).double()).norm() / torch.tensor(g["y64"]).double().norm()))
for impl in (0, 1):
    got, y, _ = run("cuda", impl)
    print("== conv_impl", impl, "y rel %.2e" % float((y - y_ref).norm() / y_ref.norm()))
    for (n, shp, i_r, o_r), (_, _, i_g, o_g) in zip(ref, got):
        print("  %-55s in %-18s in-err %.1e out-err %.1e |out| %.2e" % (n, shp, float((i_g - i_r).norm() / i_r.norm()), float((o_g - o_r).norm() / max(float(o_r.norm()), 1e-30)), float(o_r.norm())))
