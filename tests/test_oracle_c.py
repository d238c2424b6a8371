"""Pins the C restatement of the stencil path (oracle/darcy_oracle.c) against the
reference-generated fixture tests/golden/sobel_darcy.npz.  CPU only."""
import os

import numpy as np

from oracle import darcy_c


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def test_c_sobel_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "sobel_darcy.npz"))
    for tag in "abc":
        img = g[f"{tag}_img"]
        for c in (1, 0):
            assert rel(darcy_c.sobel(img, 0, c), g[f"{tag}{c}_gh"]) < 1e-13
            assert rel(darcy_c.sobel(img, 1, c), g[f"{tag}{c}_gv"]) < 1e-13
            w = g[f"{tag}{c}_w"]
            assert rel(darcy_c.sobel(w, 0, c, adjoint=True), g[f"{tag}{c}_ah"]) < 1e-13
            assert rel(darcy_c.sobel(w, 1, c, adjoint=True), g[f"{tag}{c}_av"]) < 1e-13


def test_c_darcy_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "sobel_darcy.npz"))
    for tag in "pqr":
        # fixture is float64; the C oracle takes float32 fields -> compare at fp32 rounding level
        K, out, gw = g[f"{tag}_K"], g[f"{tag}_out"], g[f"{tag}_gw"]
        for tb in (1, 0):
            l4, dout = darcy_c.darcy(K, out, gw, use_tb=bool(tb))
            assert rel(l4, g[f"{tag}{tb}_l4"]) < 5e-7
            assert rel(dout, g[f"{tag}{tb}_dout"]) < 5e-7
