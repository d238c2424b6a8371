"""CPU training loop of the oracle port: the step body of train_codec_mixed_residual.py:225-240
(host batch -> zero_grad -> DenseED forward -> 3 Darcy losses -> backward -> LR set -> Adam ->
loss.item()) restated on oracle/pdes_oracle.py, i.e. on the same PyTorch CPU kernels the
reference itself runs (MKL-DNN convolutions, native batch norm, Adam).

TEST / BASELINE INFRASTRUCTURE ONLY: used by bench.py's `cpu_baseline` leg and `--impl reference`
arm.  (The reference is pure Python and is not present on the GPU box)
"""
import time

import torch

from . import pdes_oracle as orc


class CpuTrainer(object):
    def __init__(self, imsize=64, blocks=(6, 8, 6), growth_rate=16, init_features=48, lr=1e-3, seed=1,
                 threads=None, device="cpu"):
        """device="cpu": the reference's CPU path.  device="cuda": the SAME program on PyTorch's eager CUDA
        kernels (cuDNN convolutions, native batch norm) - the library baseline of bench.py, i.e. what the
        unmodified reference would run on the same GPU."""
        if threads:
            torch.set_num_threads(int(threads))
        self.threads = torch.get_num_threads()
        self.device = torch.device(device)
        self.plan = orc.densenet_plan(1, 3, imsize, blocks, growth_rate, init_features)
        self.sd = orc.make_state(self.plan, seed)
        if self.device.type != "cpu":
            for k in self.sd:
                self.sd[k] = self.sd[k].to(self.device)
        self.names = orc.param_names(self.plan)
        for n in self.names:
            self.sd[n].requires_grad_(True)
        self.opt = torch.optim.Adam([self.sd[n] for n in self.names], lr=lr)

    def step(self, K, lr=None):
        for n in self.names:
            self.sd[n].grad = None
        out = orc.densenet_forward(self.plan, self.sd, K, training=True)
        loss, _ = orc.total_loss(K, out, 10.0)
        loss.backward()
        if lr is not None:
            for g in self.opt.param_groups:
                g['lr'] = lr
        self.opt.step()
        return loss.item()

    def timed(self, batches, warmup=1):
        """samples/s over `batches` (list of (B,1,H,W) CPU tensors) after `warmup` untimed steps."""
        for i in range(warmup):
            self.step(batches[i % len(batches)])
        t0 = time.perf_counter()
        n = 0
        for K in batches:
            self.step(K)
            n += K.shape[0]
        dt = time.perf_counter() - t0
        return n / dt, dt
