#!/usr/bin/env python
"""Benchmark of the physics-constrained DenseED training step (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port)

A step = one optimisation step on one batch of 32 synthetic GRF-KLE512 64x64 fields per GPU:
H2D copy (e2e only) -> zero_grad -> DenseED forward -> fused Darcy loss -> backward ->
[NCCL all-reduce of the flat gradient bucket] -> Adam -> loss read-back (e2e only).
Prints ONE JSON line on rank 0.  Timing: CUDA events on the launching stream, barrier +
synchronize on both sides, max over ranks.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "training-samples/sec GRF-KLE512 64x64 codec mixed-residual"
IMSIZE, BATCH, NTRAIN = 64, 32, 4096
WORKLOAD = "DenseED[6,8,6] codec mixed-residual, GRF KLE512 64x64, ntrain=4096, batch 32/GPU, fp32"
WORKLOAD_LOWP = "DenseED[6,8,6] codec mixed-residual, Channelized 64x64, ntrain=4096, batch 32/GPU, %s tensor-core conv path"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ntrain", type=int, default=NTRAIN)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-steps", type=int, default=6, help="bounded CPU-baseline sample (steps of 32)")
    ap.add_argument("--dtype", default="f32", choices=["f32", "bf16", "fp16"],
                    help="f32: fp32-accurate tensor-core path (headline, BASELINE config 2); bf16 / fp16: one-piece "
                         "tensor-core conv path on channelized 64x64 data (BASELINE config 3)")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], bf16_tflops=d["bf16_tflops"],
                    bf16_tflops_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]), src="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, src="fallback")


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=max(smax) if smax else None,
                    samples=len(sm), reasons=sorted(reasons))


def host_threads():
    """All host threads the CPU arm can use: the physical cores visible to this process (what torch picks
    by default) — set explicitly because torchrun exports OMP_NUM_THREADS=1 to every rank."""
    try:
        avail = len(os.sched_getaffinity(0))
    except AttributeError:
        avail = os.cpu_count() or 1
    cores = set()
    try:
        phys = core = None
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("physical id"):
                phys = ln.split(":")[1].strip()
            elif ln.startswith("core id"):
                core = ln.split(":")[1].strip()
            elif not ln.strip():
                if phys is not None and core is not None:
                    cores.add((phys, core))
                phys = core = None
    except OSError:
        pass
    n = len(cores) if cores else avail
    return max(1, min(n, avail))


def cpu_baseline(steps, threads=None):
    """Oracle port (same PyTorch CPU kernels as the reference) on the host cores, bounded sample."""
    import torch
    from oracle.cpu_train import CpuTrainer
    tr = CpuTrainer(IMSIZE, threads=threads or host_threads())
    g = torch.Generator().manual_seed(1)
    batches = [torch.exp(0.5 * torch.randn(BATCH, 1, IMSIZE, IMSIZE, generator=g)) for _ in range(steps)]
    sps, dt = tr.timed(batches, warmup=1)
    return dict(value=round(sps, 2), unit="samples/s", cores=tr.threads, kind="port",
                sample="%d steps of batch %d at 64x64 after 1 warm-up (%.1f s), oracle/cpu_train.py on "
                       "torch %s CPU kernels" % (steps, BATCH, dt, torch.__version__))


def gpu_library_baseline(dev, host, n_batches, steps=30, warmup=5):
    """The reference's own program (oracle port = the same torch ops) on PyTorch's eager CUDA kernels
    (cuDNN convolutions, native batch norm, torch.optim.Adam) on THIS GPU, end to end like the e2e leg
    (pinned host batch -> H2D -> step -> loss.item()): the library path this repo's kernels replace
    (SURVEY.md section 8d "also record PyTorch-eager CUDA (cuDNN) on one B200").  Not part of any timed
    region of the product arm."""
    import torch
    from oracle.cpu_train import CpuTrainer
    res = {}
    prev = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark)
    try:
        for name, tf32 in (("fp32", False), ("tf32", True)):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            torch.backends.cudnn.benchmark = True
            tr = CpuTrainer(IMSIZE, device=dev)

            def one(i):
                K = host[(i % n_batches) * BATCH:(i % n_batches + 1) * BATCH].to(dev, non_blocking=True)
                return tr.step(K)

            for i in range(warmup):
                one(i)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(steps):
                one(warmup + i)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            res[name] = dict(value=round(BATCH / (ms * 1e-3), 1), unit="samples/s", ms_per_step=round(ms, 3))
            del tr
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark = prev
    res["what"] = ("PyTorch %s eager on this GPU: cuDNN convolutions + native batch norm + torch.optim.Adam, same "
                   "step body and host-buffer protocol as e2e, batch %d, %d steps after %d warm-up; tf32 = cuDNN "
                   "TF32 convolutions allowed (fails the 1e-4 parity bar, SURVEY.md headline facts)"
                   % (torch.__version__, BATCH, steps, warmup))
    return res


# first training-mode loss of DenseED(1,3,64,[6,8,6]) default-initialised under torch.manual_seed(1) on
# K = exp(0.5*randn(32,1,64,64)) drawn right after: probed on the REFERENCE itself (SURVEY.md section 8c:
# 649.2476 in fp32) and reproduced by the fp64 oracle (649.24758541)
PINNED_LOSS = 649.24758541


def shutdown_process_group(ts=None, grace_s=20.0):
    """Tear the NCCL group down without ever hanging the bench: the JSON line is already printed; if the
    communicator does not go away within `grace_s` the process exits anyway."""
    import torch
    import torch.distributed as dist
    sys.stdout.flush()
    sys.stderr.flush()

    def _down():
        try:
            if ts is not None:
                ts.release()   # the step graph holds the captured NCCL all-reduce
            torch.cuda.synchronize()
            dist.barrier()
            dist.destroy_process_group()
        except Exception:
            pass

    th = threading.Thread(target=_down, daemon=True)
    th.start()
    th.join(grace_s)
    if th.is_alive():
        os._exit(0)


def run_reference(args, rank):
    """--impl reference: the reference's CPU implementation of the path (oracle port; the reference is
    pure Python/PyTorch and absent from the GPU box).  Rank 0 only."""
    if rank != 0:
        return
    import torch
    from oracle.cpu_train import CpuTrainer
    tr = CpuTrainer(IMSIZE, threads=host_threads())
    g = torch.Generator().manual_seed(1)
    n = max(1, min(args.steps, 40))
    batches = [torch.exp(0.5 * torch.randn(BATCH, 1, IMSIZE, IMSIZE, generator=g)) for _ in range(min(n, 8))]
    for i in range(max(1, min(args.warmup, 3))):
        tr.step(batches[i % len(batches)])
    t0 = time.perf_counter()
    for i in range(n):
        tr.step(batches[i % len(batches)])
    dt = time.perf_counter() - t0
    sps = n * BATCH / dt
    line = dict(metric=METRIC, value=round(sps, 2), unit="samples/s", n_gpus=args.gpus, steps=n,
                warmup=args.warmup, ms_per_step=round(1e3 * dt / n, 3), higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="f32", data="synthetic", impl="reference",
                config=dict(workload=WORKLOAD + " (CPU, one process, all host threads)"),
                cpu_baseline=dict(value=round(sps, 2), unit="samples/s", cores=tr.threads, kind="port",
                                  sample="%d timed steps of batch %d (bounded from --steps %d)" % (n, BATCH, args.steps)),
                e2e=dict(value=round(sps, 2), unit="samples/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.dtype != "f32":
        os.environ["PDES_CONV_DTYPE"] = args.dtype   # read by the executor when a network is created
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the sm_100a path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_P2P_LEVEL", "NVL")
        dist.init_process_group("nccl", device_id=dev)
    from models.codec import DenseED
    from models.darcy import conv_boundary_condition, conv_constitutive_constraint, conv_continuity_constraint
    from pde_surrogate_b200 import _lib, data
    from pde_surrogate_b200.engine import TrainStep
    from utils.image_gradient import SobelFilter
    from utils.practices import OneCycleScheduler, adjust_learning_rate

    pk = peaks()
    # ---- data: synthetic GRF KLE512 64x64, rank-sharded ------------------------------------
    n_local = args.ntrain // world
    lowp = args.dtype != "f32"
    if lowp:   # BASELINE config 3: two-valued channelized permeability fields
        host = data.channelized(n_local, IMSIZE, seed=1 + rank).pin_memory()
    else:
        host = data.grf_kle(n_local, IMSIZE, 512, 0.1, seed=1 + rank, device=dev).pin_memory()
    dset = host.to(dev)
    n_batches = n_local // BATCH

    def make_model():
        import contextlib
        torch.manual_seed(1)
        with contextlib.redirect_stdout(sys.stderr):  # the constructor prints '# params ...' like the reference
            return DenseED(1, 3, IMSIZE, [6, 8, 6]).to(dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for i in range(warmup):
            fn(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(warmup + i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    sched = OneCycleScheduler(lr_max=1e-3, div_factor=2.0, pct_start=0.3)
    total_steps = 300 * n_batches

    # ---- value: inputs resident in HBM, whole step on the device ---------------------------
    model = make_model()
    pg = dist.group.WORLD if world > 1 else None
    ts = TrainStep(model, weight_bound=10.0, lr=1e-3, process_group=pg, world_size=world)
    if world > 1:
        ts.broadcast_parameters()
    last = {}
    use_graph = os.environ.get("PDES_BENCH_GRAPH", "1") != "0"

    def dev_step(i):
        K = dset[(i % n_batches) * BATCH:(i % n_batches + 1) * BATCH]
        lr = sched.step((i + 1) / total_steps)
        last["loss"] = ts.step_graph(K, lr=lr) if use_graph else ts.step(K, lr=lr)

    # ---- parity preflight on the timed code path (graph replay, batch 32, 64x64) --------------
    parity = None
    if rank == 0:
        import contextlib
        torch.manual_seed(1)
        with contextlib.redirect_stdout(sys.stderr):
            pm = DenseED(1, 3, IMSIZE, [6, 8, 6])
        Kp = torch.exp(0.5 * torch.randn(BATCH, 1, IMSIZE, IMSIZE))
        pts = TrainStep(pm.to(dev), weight_bound=10.0, lr=1e-3)
        lp = float((pts.step_graph(Kp.to(dev), lr=1e-3) if use_graph else pts.step(Kp.to(dev), lr=1e-3)).item())
        pbar = {"f32": 1e-4, "fp16": 2e-2, "bf16": 1e-1}[args.dtype]   # one-piece modes: the format's own error
        parity = dict(loss=round(lp, 5), expected=PINNED_LOSS, rel_err=abs(lp - PINNED_LOSS) / PINNED_LOSS,
                      bar=pbar, what="first training-mode loss, seed-1 default init, K = exp(0.5 randn) "
                                     "(reference-probed scalar, SURVEY.md section 8c), through the timed "
                                     "engine path")
        if not parity["rel_err"] <= pbar:
            raise SystemExit("bench.py: parity preflight failed: loss %.6f vs pinned %.6f" % (lp, PINNED_LOSS))
        del pts, pm
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms = timed(dev_step, args.steps, max(args.warmup, 3))
    clocks = sampler.stop() if rank == 0 else None
    value = world * BATCH * args.steps / (ms * 1e-3)
    launches = ts.kernel_launches * args.steps
    final_loss = float(last["loss"].item())

    # ---- roofline of the conv path: useful FLOPs of fwd+bwd / device time (whole executor) ------
    L = _lib.lib()
    ex = model._ex
    flops_step = model.flops(BATCH, True)
    Kb = dset[:BATCH]
    dummy = torch.randn(BATCH, 3, IMSIZE, IMSIZE, device=dev) * 1e-3

    def conv_only(i):
        ex.forward(Kb, True)
        ex.backward(dummy)

    ms_conv = timed(conv_only, max(5, args.steps // 4), 3)
    conv_ms_step = ms_conv / max(5, args.steps // 4)
    ach_tf = flops_step / (conv_ms_step * 1e-3) / 1e12
    roofline_step = dict(bound="tensor", kernel="DenseED conv path (all forward / dgrad / wgrad launches of a step)",
                         achieved=round(ach_tf, 3), peak=pk["bf16_tflops_sustained"], unit="TFLOP/s",
                         frac=round(ach_tf / pk["bf16_tflops_sustained"], 5), traffic=None,
                         note="useful fp32 FLOPs (2*MAC, %.1f GFLOP/step) / CUDA-event time of the executor's "
                              "forward+backward" % (flops_step / 1e9))

    # ---- roofline of the dominant kernel: per-launch CUDA events on the launching stream, in situ ----
    # (a second executor with the wgrad side stream off so that every launch is timed on one stream)
    roofline, families = None, {}
    if rank == 0:
        import ctypes
        os.environ["PDES_WGRAD_STREAMS"] = "0"
        model3 = make_model()
        ts3 = TrainStep(model3, weight_bound=10.0, lr=1e-3)
        for i in range(3):
            ts3.step(Kb)
        h3 = model3._ex.handle.h
        cap = 512
        us = (ctypes.c_float * cap)()
        fl = (ctypes.c_double * cap)()
        lab = ctypes.create_string_buffer(64 * cap)
        per = {}
        reps = 7
        for r in range(reps):
            _lib.check(L.pdes_densenet_set_timing(h3, 1))
            ts3.step(Kb)
            k = L.pdes_densenet_timing_read(h3, us, fl, cap, lab, len(lab))
            _lib.check(L.pdes_densenet_set_timing(h3, 0))
            names = lab.value.decode().split("\n")
            for i in range(min(k, cap)):
                per.setdefault(names[i], ([], fl[i]))[0].append(us[i])
        os.environ.pop("PDES_WGRAD_STREAMS", None)
        del ts3, model3
        med = {n: (statistics.median(v), f) for n, (v, f) in per.items()}
        kname = {"conv.f": "conv_tc2_kernel (forward)", "dgrad": "conv_tc2_kernel (dgrad + BatchNorm-backward epilogue)",
                 "wgrad": "wgrad_tc_kernel"}
        for n, (t_us, f) in med.items():
            fam = n.split()[0]
            if f > 0:
                a = families.setdefault(fam, [0.0, 0.0, 0])
                a[0] += t_us
                a[1] += f
                a[2] += 1
        top = max(((t_us, f, n) for n, (t_us, f) in med.items() if f > 0 and n.split()[0] in kname), default=None)
        traffic = {}
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath))
        if top is not None:
            t_us, f, n = top
            ach = f / (t_us * 1e-6) / 1e12
            roofline = dict(bound="tensor", kernel="%s, launch '%s'" % (kname[n.split()[0]], n),
                            achieved=round(ach, 2), peak=pk["bf16_tflops_sustained"], unit="TFLOP/s",
                            frac=round(ach / pk["bf16_tflops_sustained"], 4), traffic=traffic.get(n),
                            us_per_launch=round(t_us, 2), gflop_per_launch=round(f / 1e9, 3),
                            tensor_issue_frac=round((1 if lowp else 3) * ach / pk["bf16_tflops_sustained"], 4),
                            note="dominant launch of the step; achieved = useful fp32 FLOPs (2*MAC) of the launch / "
                                 "median CUDA-event time between consecutive launches of an eager step (includes "
                                 "the ~2 us launch gap); peak = %s dense bf16 (sustained); %s" % (pk["src"],
                                 "one tensor product per useful product (one-piece operands)" if lowp else
                                 "fp32 accuracy costs 3 fp16 tensor products per useful product, so 1/3 is the "
                                 "ceiling and tensor_issue_frac = 3*frac is what the tensor pipe executes"))
        families = {k: dict(us_per_step=round(v[0], 1), launches=v[2], useful_tflops=round(v[1] / (v[0] * 1e-6) / 1e12, 2))
                    for k, v in families.items()}

    # ---- roofline of the fused stencil kernel on a cold, larger-than-L2 batch ----------------
    nb = 8192
    Kbig = dset[:min(nb, dset.shape[0])].repeat((nb + dset.shape[0] - 1) // dset.shape[0], 1, 1, 1)[:nb].contiguous()
    obig = torch.randn(nb, 3, IMSIZE, IMSIZE, device=dev)
    dbig = torch.empty_like(obig)
    l4 = torch.zeros(4, device=dev)
    gw = torch.tensor([1., 1., 10., 10.], device=dev)
    from pde_surrogate_b200 import darcy as _d
    ws = _d._workspace(dev)
    st = _lib.stream_ptr()

    def sten_f(i):
        _lib.check(L.pdes_darcy_loss_fwd(_lib.ptr(Kbig), _lib.ptr(obig), nb, IMSIZE, IMSIZE, 1, _lib.ptr(l4),
                                         _lib.ptr(ws), st))

    def sten_b(i):
        _lib.check(L.pdes_darcy_loss_bwd(_lib.ptr(Kbig), _lib.ptr(obig), _lib.ptr(gw), nb, IMSIZE, IMSIZE, 1,
                                         _lib.ptr(dbig), st))

    msf = timed(sten_f, 10, 3) / 10
    msb = timed(sten_b, 10, 3) / 10
    gbs_f = nb * 65536 / (msf * 1e-3) / 1e9
    gbs_b = nb * 114688 / (msb * 1e-3) / 1e9
    _tr = {}
    if os.path.exists(os.path.join(ROOT, "profiles", "traffic.json")):
        _tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    roofline_stencil = dict(bound="hbm", kernel="darcy_fwd_tile_kernel / darcy_bwd_tile_kernel",
                            achieved=round(gbs_b, 1), peak=pk["hbm_gbs"], unit="GB/s",
                            frac=round(gbs_b / pk["hbm_gbs"], 4), traffic=_tr.get("darcy_bwd_tile_kernel"),
                            fwd_traffic=_tr.get("darcy_fwd_tile_kernel"), algorithmic_bytes=nb * 114688,
                            fwd_algorithmic_bytes=nb * 65536,
                            fwd_achieved=round(gbs_f, 1), fwd_frac=round(gbs_f / pk["hbm_gbs"], 4),
                            note="cold %d-sample batch (%.1f GB, > L2); algorithmic bytes 65536 (fwd) / "
                                 "114688 (bwd) per sample; peak = %s copy bandwidth" % (nb, nb * 114688 / 1e9, pk["src"]))
    del Kbig, obig, dbig

    # ---- e2e: the unmodified script's loop body through the reference-facing modules ---------
    model2 = make_model()
    # the optimizer the unmodified script gets under run_reference_script.py: `optim.Adam` resolves to the fused
    # subclass of torch.optim.Adam (pde_surrogate_b200/optim.py)
    from pde_surrogate_b200 import optim as pdes_optim
    pdes_optim.install()
    opt = torch.optim.Adam(model2.parameters(), lr=1e-3, weight_decay=0.0)
    sob = SobelFilter(IMSIZE, correct=True, device=dev)
    if world > 1:
        flat2, gflat2 = model2.flat_parameters()
        dist.broadcast(flat2, 0)

    def e2e_step(i):
        inp = host[(i % n_batches) * BATCH:(i % n_batches + 1) * BATCH].to(dev, non_blocking=True)
        model2.zero_grad()
        out = model2(inp)
        loss_pde = conv_constitutive_constraint(inp, out, sob) + conv_continuity_constraint(out, sob)
        l_dir, l_neu = conv_boundary_condition(out)
        loss = loss_pde + (l_dir + l_neu) * 10.0
        loss.backward()
        if world > 1:
            dist.all_reduce(gflat2, op=dist.ReduceOp.AVG)   # NCCL averages in the collective: no extra launch
        adjust_learning_rate(opt, sched.step((i + 1) / total_steps))
        opt.step()
        last["e2e_loss"] = loss.item()

    ms_e = timed(e2e_step, args.steps, max(args.warmup, 3))
    e2e_val = world * BATCH * args.steps / (ms_e * 1e-3)
    fused_adam_steps = int(getattr(opt, "fused_steps", 0))
    # the same loop with torch's own Adam step (PDES_FUSED_ADAM=0), reported beside the headline
    os.environ["PDES_FUSED_ADAM"] = "0"
    ms_e_stock = timed(e2e_step, args.steps, 3)
    os.environ.pop("PDES_FUSED_ADAM")
    e2e_stock_val = world * BATCH * args.steps / (ms_e_stock * 1e-3)

    cpu = None
    lib_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not lowp:
        lib_base = gpu_library_baseline(dev, host, n_batches)
        cpu = cpu_baseline(args.cpu_steps)

    if rank == 0:
        line = dict(metric=METRIC, value=round(value, 1), unit="samples/s", n_gpus=world, steps=args.steps,
                    warmup=max(args.warmup, 3), ms_per_step=round(ms / args.steps, 4), higher_is_better=True,
                    scaling="weak", vs_baseline=None, dtype=args.dtype, data="synthetic",
                    config=dict(workload=WORKLOAD if not lowp else WORKLOAD_LOWP % args.dtype, global_batch=BATCH * world,
                                parallelism="dp%d" % world if world > 1 else "single",
                                cuda_graph=bool(use_graph),
                                l2="each step streams ~210 MB of activations+gradients (> 126 MB L2) and a "
                                   "different batch of the HBM-resident dataset; no explicit flush",
                                grf=("exp covariance, l=0.1, 512 KLE modes, seed 1" if not lowp else
                                     "channelized: two-valued fields {1, e^2.5} from thresholded low-pass noise"),
                                conv_dtype=("two fp16 pieces per operand, 3 tensor-core products (fp32-accurate)"
                                            if not lowp else "one %s piece per operand, 1 tensor-core product" % args.dtype)),
                    e2e=dict(value=round(e2e_val, 1), unit="samples/s", h2d_bytes_per_step=BATCH * IMSIZE * IMSIZE * 4,
                             d2h_bytes_per_step=4, ms_per_step=round(ms_e / args.steps, 4),
                             api="models.codec.DenseED + models.darcy.conv_* + optim.Adam as run_reference_script.py "
                                 "installs it (fused subclass of torch.optim.Adam), loss.item() per step",
                             fused_adam_steps=fused_adam_steps,
                             stock_adam=dict(value=round(e2e_stock_val, 1), ms_per_step=round(ms_e_stock / args.steps, 4),
                                             note="same loop, torch's own foreach Adam step (PDES_FUSED_ADAM=0)")),
                    gpu_launches=launches, launches_per_step=ts.kernel_launches, clocks=clocks,
                    roofline=roofline, roofline_step=roofline_step, roofline_families=families,
                    roofline_stencil=roofline_stencil, cpu_baseline=cpu, gpu_library_baseline=lib_base,
                    parity_check=parity,
                    final_loss=round(final_loss, 5), conv_path_ms_per_step=round(conv_ms_step, 4),
                    useful_gflop_per_step=round(flops_step / 1e9, 2))
        print(json.dumps(line), flush=True)
    if world > 1:
        shutdown_process_group(ts)


if __name__ == "__main__":
    main()
