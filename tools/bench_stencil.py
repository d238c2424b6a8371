"""A/B timing of the fused Darcy-loss tile kernels on a cold (larger than L2) batch.
   python tools/bench_stencil.py [impl ...]   (impl numbers of pdes_darcy_loss_set_impl)"""
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
impls = [int(a) for a in sys.argv[1:]] or [0, 2, 3, 4, 5]
ref = None
for impl in impls:
    _lib.check(L.pdes_darcy_loss_set_impl(impl))
    def fwd():
        _lib.check(L.pdes_darcy_loss_fwd(_lib.ptr(K), _lib.ptr(out), nb, H, H, 1, _lib.ptr(l4), _lib.ptr(ws), st))
    def bwd():
        _lib.check(L.pdes_darcy_loss_bwd(_lib.ptr(K), _lib.ptr(out), _lib.ptr(gw), nb, H, H, 1, _lib.ptr(dout), st))
    res = {}
    for name, fn, nbytes in (("fwd", fwd, nb * 4 * H * H * 4), ("bwd", bwd, nb * 7 * H * H * 4)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = 10
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        res[name] = (ms, nbytes / ms / 1e6)
    chk = (l4.tolist(), float(dout.double().abs().sum()))
    if ref is None:
        ref = chk
    dl = max(abs(a - b) / abs(b) for a, b in zip(chk[0], ref[0]))
    print(f"impl {impl}: fwd {res['fwd'][0]:.3f} ms {res['fwd'][1]:.0f} GB/s | bwd {res['bwd'][0]:.3f} ms "
          f"{res['bwd'][1]:.0f} GB/s | loss dev {dl:.1e} |dout| dev {abs(chk[1]-ref[1])/ref[1]:.1e}", flush=True)
_lib.check(L.pdes_darcy_loss_set_impl(0))
