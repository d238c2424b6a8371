#!/usr/bin/env python
"""Run the UNMODIFIED reference script train_cglow_reverse_kl.py (BASELINE config 5: cGlow, 32x32, batch 32, encoder
[3,4,4], flow [6,6,6], LU 1x1 convolutions) on a GPU against this repo's backend; log + the script's own clock.

Stage the script first in the build container (git-ignored, travels with the gpurun snapshot, never committed):

    mkdir -p baseline/_ref && cp /root/reference/train_cglow_reverse_kl.py baseline/_ref/

Datasets are synthetic (pde_surrogate_b200.data: GRF KLE100 inputs at 32x32, finite-volume reference outputs for the
validation file).  Note: the script wraps every training step in autograd.detect_anomaly() (line 254), which is a
host-side cost of the reference's own loop."""
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
    ap.add_argument("--script", default=os.path.join(ROOT, "baseline", "_ref", "train_cglow_reverse_kl.py"))
    ap.add_argument("--ntrain", type=int, default=512)
    ap.add_argument("--ntest", type=int, default=64)
    ap.add_argument("--batch-size", type=int, default=32)
    ap.add_argument("--epochs", type=int, default=3)
    ap.add_argument("--work", default="/tmp/pdes_cglow_run")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "cglow_run"))
    a = ap.parse_args()
    from pde_surrogate_b200 import data
    data.write_script_datasets(os.path.join(a.work, "datasets"), 32, a.ntrain, a.ntest, kind="grf_kle100", seed=1)
    cmd = [sys.executable, os.path.join(ROOT, "run_reference_script.py"), "--script", a.script, "--",
           "--data-dir", os.path.join(a.work, "datasets"), "--exp-dir", os.path.join(a.work, "exp"), "--imsize", "32",
           "--ntrain", str(a.ntrain), "--ntest", str(a.ntest), "--batch-size", str(a.batch_size), "--test-batch-size", "64",
           "--epochs", str(a.epochs), "--cuda", "0", "--plot-freq", "1000", "--ckpt-freq", str(a.epochs)]
    os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
    t0 = time.time()
    r = subprocess.run(cmd, capture_output=True, text=True)
    wall = time.time() - t0
    with open(a.out + ".log", "w") as f:
        f.write("$ " + " ".join(cmd) + "\n" + r.stdout[-20000:] + "\n--- stderr ---\n" + r.stderr[-8000:])
    tt = None
    for root, _d, files in os.walk(os.path.join(a.work, "exp")):
        if "args.txt" in files:
            tt = json.load(open(os.path.join(root, "args.txt"))).get("training_time")
    summary = dict(returncode=r.returncode, wall_s=round(wall, 2), script_training_time_s=tt, epochs=a.epochs,
                   steps=a.epochs * (a.ntrain // a.batch_size),
                   samples_per_s_script_clock=(a.epochs * a.ntrain / tt) if tt else None,
                   tail=[l for l in r.stdout.strip().splitlines() if l.startswith(("Epoch", "Finished", "("))][-8:],
                   note="script clock = the reference's own time.time() around its epoch loop: reverse-KL training steps under "
                        "autograd.detect_anomaly() + test() every epoch (the last one with 6 x (20 + 15) sampling passes) + checkpoint")
    with open(a.out + ".json", "w") as f:
        json.dump(summary, f, indent=1)
    print(json.dumps(summary))
    return r.returncode


if __name__ == "__main__":
    sys.exit(main())
