#!/usr/bin/env python
"""Run the UNMODIFIED reference solver solve_conv_mixed_residual.py (Decoder + L-BFGS on the conv mixed-residual
loss, linear and --nonlinear law) on a GPU against this repo's backend.  Stage the script first (build container):

    mkdir -p baseline/_ref && cp /root/reference/solve_conv_mixed_residual.py baseline/_ref/
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--script", default=os.path.join(ROOT, "baseline", "_ref", "solve_conv_mixed_residual.py"))
    ap.add_argument("--epochs", type=int, default=30)
    ap.add_argument("--work", default="/tmp/pdes_solver_run")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "solver_run"))
    a = ap.parse_args()
    from pde_surrogate_b200 import data
    x = data.grf_kle(16, 64, 512, 0.1, seed=5).numpy()
    data.write_hdf5(os.path.join(a.work, "datasets", "64x64", "kle512_lhs1000_test.hdf5"), x, data.darcy_fv_dataset(x))
    res = {}
    for mode, extra in (("linear", []), ("nonlinear", ["--nonlinear", "--alpha1", "1.0", "--alpha2", "1.0"])):
        cmd = [sys.executable, os.path.join(ROOT, "run_reference_script.py"), "--script", a.script, "--",
               "--data-dir", os.path.join(a.work, "datasets"), "--exp-dir", os.path.join(a.work, "exp_" + mode),
               "--idx", "8", "--epochs", str(a.epochs), "--test-freq", str(a.epochs), "--ckpt-freq", str(a.epochs),
               "--cuda", "0"] + extra
        t0 = time.time()
        r = subprocess.run(cmd, capture_output=True, text=True)
        wall = time.time() - t0
        lines = [l for l in r.stdout.splitlines() if l.startswith("epoch ")]
        losses = [float(l.split("loss")[1]) for l in lines if "loss" in l]
        os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
        with open(a.out + "_" + mode + ".log", "w") as f:
            f.write("$ " + " ".join(cmd) + "\n" + r.stdout[-6000:] + "\n--- stderr ---\n" + r.stderr[-4000:])
        res[mode] = dict(returncode=r.returncode, wall_s=round(wall, 2), epochs=a.epochs,
                         lbfgs_iterations=a.epochs * 20, first_loss=losses[0] if losses else None,
                         last_loss=losses[-1] if losses else None, tail=r.stdout.strip().splitlines()[-3:])
    json.dump(res, open(a.out + ".json", "w"), indent=1)
    print(json.dumps(res))
    return max(v["returncode"] for v in res.values())


if __name__ == "__main__":
    sys.exit(main())
