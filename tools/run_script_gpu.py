#!/usr/bin/env python
"""Run the UNMODIFIED reference script train_codec_mixed_residual.py on a GPU against this repo's
backend and record log + wall-clock throughput.

The reference checkout does not exist on the GPU box.  Stage the one script first, in the build
container (git-ignored, travels with the gpurun snapshot, never committed):

    mkdir -p baseline/_ref && cp /root/reference/train_codec_mixed_residual.py baseline/_ref/

then on the box:

    python tools/run_script_gpu.py --epochs 3 --out gpurun_out/script_run

Datasets are synthetic (pde_surrogate_b200.data: GRF KLE512 inputs, finite-volume reference outputs
for the validation file)."""
import argparse
import json
import os
import re
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--script", default=os.path.join(ROOT, "baseline", "_ref", "train_codec_mixed_residual.py"))
    ap.add_argument("--imsize", type=int, default=64)
    ap.add_argument("--ntrain", type=int, default=4096)
    ap.add_argument("--ntest", type=int, default=512)
    ap.add_argument("--batch-size", type=int, default=32)
    ap.add_argument("--epochs", type=int, default=3)
    ap.add_argument("--data", default="grf_kle512")
    ap.add_argument("--work", default="/tmp/pdes_script_run")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "script_run"))
    ap.add_argument("--extra", nargs="*", default=[])
    a = ap.parse_args()
    from pde_surrogate_b200 import data
    t0 = time.time()
    data.write_script_datasets(os.path.join(a.work, "datasets"), a.imsize, a.ntrain, a.ntest, kind=a.data, seed=1)
    t_data = time.time() - t0
    cmd = [sys.executable, os.path.join(ROOT, "run_reference_script.py"), "--script", a.script, "--",
           "--data-dir", os.path.join(a.work, "datasets"), "--exp-dir", os.path.join(a.work, "exp"),
           "--data", a.data, "--imsize", str(a.imsize), "--ntrain", str(a.ntrain), "--ntest", str(a.ntest),
           "--batch-size", str(a.batch_size), "--epochs", str(a.epochs), "--cuda", "0", "--plot-freq", "1000",
           "--ckpt-freq", str(a.epochs)] + list(a.extra)
    os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
    summary, rc_all = {}, 0
    # twice: the reference's CPU DataLoader (a host->device copy per step), then the GPU-resident loader
    for tag, dev in (("cpu_dataloader", ""), ("resident_loader", "cuda:0")):
        env = dict(os.environ)
        env["PDES_DATA_DEVICE"] = dev
        exp = os.path.join(a.work, "exp_" + tag)
        c2 = [x if x != os.path.join(a.work, "exp") else exp for x in cmd]
        t0 = time.time()
        r = subprocess.run(c2, capture_output=True, text=True, env=env)
        wall = time.time() - t0
        rc_all = max(rc_all, r.returncode)
        with open(a.out + "_" + tag + ".log", "w") as f:
            f.write("$ " + " ".join(c2) + "\n" + r.stdout + "\n--- stderr ---\n" + r.stderr[-8000:])
        # the script stores its own wall time of the epoch loop in args.txt (train_codec_mixed_residual.py:255-260)
        tt = None
        for root, _d, files in os.walk(exp):
            if "args.txt" in files:
                tt = json.load(open(os.path.join(root, "args.txt"))).get("training_time")
        epochs = re.findall(r"[Ee]poch[: ]+(\d+).*", r.stdout)
        summary[tag] = dict(returncode=r.returncode, wall_s=round(wall, 2), script_training_time_s=tt, epochs=a.epochs,
                            steps=a.epochs * (a.ntrain // a.batch_size),
                            samples_per_s_script_clock=(a.epochs * a.ntrain / tt) if tt else None,
                            n_epoch_lines=len(epochs), tail=r.stdout.strip().splitlines()[-6:])
    summary["dataset_s"] = round(t_data, 2)
    summary["note"] = ("script clock = the reference's own time.time() around its epoch loop: training steps (H2D, "
                       "forward, 3 loss calls, backward, Adam, loss.item()) + test() every epoch + checkpoint")
    with open(a.out + ".json", "w") as f:
        json.dump(summary, f, indent=1)
    print(json.dumps(summary))
    return rc_all


if __name__ == "__main__":
    sys.exit(main())
