"""Cold, larger-than-L2 batch through the fused Darcy loss kernels (for ncu):
   ncu --set full -k regex:darcy_ -c 4 -o gpurun_out/prof_stencil python tools/profile_stencil.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pde_surrogate_b200 import _lib, darcy
L = _lib.lib()
nb, H = 8192, 64
K = torch.exp(0.5 * torch.randn(nb, 1, H, H, device="cuda"))
out = torch.randn(nb, 3, H, H, device="cuda")
dout = torch.empty_like(out)
l4 = torch.zeros(4, device="cuda")
gw = torch.tensor([1., 1., 10., 10.], device="cuda")
ws = darcy._workspace(K.device)
st = _lib.stream_ptr()
for _ in range(2):
    _lib.check(L.pdes_darcy_loss_fwd(_lib.ptr(K), _lib.ptr(out), nb, H, H, 1, _lib.ptr(l4), _lib.ptr(ws), st))
    _lib.check(L.pdes_darcy_loss_bwd(_lib.ptr(K), _lib.ptr(out), _lib.ptr(gw), nb, H, H, 1, _lib.ptr(dout), st))
torch.cuda.synchronize()
print(l4.tolist())
