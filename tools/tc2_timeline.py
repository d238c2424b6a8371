"""Per-phase SM-clock timestamps of conv_tc2 CTAs (PDES_TC2_DBG=1) for a few layer shapes."""
import os, sys
os.environ["PDES_TC2_DBG"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ctypes import byref
from pde_surrogate_b200 import _lib
L = _lib.lib()
for (B, H, Cin, Cout, K, up) in [(32, 32, 196, 98, 3, 0), (32, 32, 98, 49, 3, 1), (32, 32, 100, 100, 3, 0)]:
    d = _lib.ConvDesc()
    Hv = 2 * H if up else H
    d.B, d.Hin, d.Win, d.Cin, d.ld_in = B, H, H, Cin, (Cin + 3) // 4 * 4
    d.Hout, d.Wout, d.Cout, d.ld_out, d.c_off_out = Hv, Hv, Cout, (Cout + 3) // 4 * 4, 0
    d.KH = d.KW = K; d.stride = 1; d.pad = K // 2; d.upsample = up; d.bn_relu = 1; d.out_nchw = 0
    x = torch.randn(B, H, H, d.ld_in, device="cuda"); w = torch.randn(Cout, Cin, K, K, device="cuda") * 0.05
    sc = torch.rand(Cin, device="cuda") + 0.5; sh = torch.randn(Cin, device="cuda") * 0.1
    y = torch.zeros(B, Hv, Hv, d.ld_out, device="cuda")
    cs = torch.zeros(Cout, dtype=torch.float64, device="cuda"); cq = torch.zeros_like(cs)
    print("fwd", (B, H, Cin, Cout, K, up), flush=True)
    _lib.check(L.pdes_conv2d_fwd(byref(d), _lib.ptr(x), _lib.ptr(w), _lib.ptr(sc), _lib.ptr(sh), _lib.ptr(y), _lib.ptr(cs), _lib.ptr(cq), 2, _lib.stream_ptr()))
    torch.cuda.synchronize()
    dy = torch.randn(B, Hv, Hv, d.ld_out, device="cuda"); da = torch.zeros(B, H, H, Cin, device="cuda")
    print("dgrad", flush=True)
    _lib.check(L.pdes_conv2d_dgrad(byref(d), _lib.ptr(dy), _lib.ptr(w), _lib.ptr(da), 2, _lib.stream_ptr()))
    torch.cuda.synchronize()
