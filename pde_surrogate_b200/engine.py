"""Whole training step on device-resident data (the loop body of
train_codec_mixed_residual.py:224-240 without the host round trips): zero the flat gradient
bucket, DenseED forward, fused Darcy loss forward+backward, DenseED backward (wgrad writes into
the flat bucket), optional NCCL all-reduce of the bucket over NVLink, fused flat Adam.
One process per GPU; data parallel = shard the minibatch, average gradients.
"""
import torch

from . import _lib
from . import darcy as _darcy
from . import ddp


class TrainStep(object):
    def __init__(self, model, weight_bound=10.0, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0,
                 process_group=None, world_size=1):
        self.model = model
        self.flat, self.gflat = model.flat_parameters()
        if not self.flat.is_cuda:
            raise RuntimeError("TrainStep needs the model on a CUDA device")
        if model._cfg.get("drop_rate", 0.0) > 0:
            raise NotImplementedError("TrainStep (CUDA-graph engine) does not redraw dropout masks inside the graph; "
                                      "train a drop_rate > 0 network through the module API")
        self.m = torch.zeros_like(self.flat)
        self.v = torch.zeros_like(self.flat)
        self.gw = torch.tensor([1.0, 1.0, weight_bound, weight_bound], device=self.flat.device)
        self.lr, self.betas, self.eps, self.wd = lr, betas, eps, weight_decay
        self.pg, self.world = process_group, world_size
        # data parallel: the gradient all-reduce and Adam are captured INSIDE the step graph (falls back to
        # eager launches behind the replay if the collective cannot be captured)
        self.collective_in_graph = world_size > 1
        self.steps = 0
        self.kernel_launches = 0
        self.l4 = torch.zeros(4, device=self.flat.device)
        # Adam scalars of step t travel host -> device through a RING of pinned slots, each guarded by a
        # CUDA event: the host may run many graph replays ahead of the device, so a single pinned
        # buffer would be overwritten before its asynchronous copy has executed
        self._hyper_slots = 32
        self.hyper_host = torch.zeros(self._hyper_slots, 8, dtype=torch.float32).pin_memory()
        self._hyper_ev = [None] * self._hyper_slots
        self._hyper_i = 0
        self.hyper_dev = torch.zeros(8, dtype=torch.float32, device=self.flat.device)
        self.graph = None
        self.static_K = None
        self.static_loss = None
        model.train()
        for p, v in zip(model._params, model._grad_views):
            p.grad = v

    def release(self):
        """Drop the captured step graph (it holds the NCCL all-reduce when data parallel): call before
        torch.distributed.destroy_process_group(), which otherwise waits on a communicator that a live
        graph still references."""
        import gc
        self.graph = None
        self.static_loss = None
        gc.collect()
        torch.cuda.synchronize()

    def broadcast_parameters(self):
        """Rank 0's parameters and BatchNorm buffers to every rank (start of DDP training)."""
        ddp.broadcast_state_([self.flat, self.model._flat_running], 0, group=self.pg)

    # ------------------------------------------------------------------ CUDA-graph replay
    def _device_work(self, K):
        """Everything of a step that is pure device work with step-independent launch arguments
        (the Adam scalars are read from self.hyper_dev): capturable into a CUDA graph."""
        L = _lib.lib()
        ex = self.model._ex
        st = _lib.stream_ptr()
        B, _, H, W = K.shape
        self.gflat.zero_()
        out = ex.forward(K, True)
        n = L.pdes_densenet_last_launches(ex.handle.h)
        _lib.check(L.pdes_darcy_loss_fwd(_lib.ptr(K), _lib.ptr(out), B, H, W, 1, _lib.ptr(self.l4),
                                         _lib.ptr(_darcy._workspace(K.device)), st), "pdes_darcy_loss_fwd")
        dout = torch.empty_like(out)
        _lib.check(L.pdes_darcy_loss_bwd(_lib.ptr(K), _lib.ptr(out), _lib.ptr(self.gw), B, H, W, 1,
                                         _lib.ptr(dout), st), "pdes_darcy_loss_bwd")
        ex.backward(dout)
        n += L.pdes_densenet_last_launches(ex.handle.h) + 2
        if self.world == 1 or self.collective_in_graph:
            if self.world > 1:
                # ONE sum all-reduce of the flat gradient bucket (wgrad wrote straight into it); the 1/world
                # of the mean is folded into the fused Adam (grad_scale).  NCCL collectives are capturable:
                # inside the step graph the all-reduce and Adam follow the backward without host launches.
                ddp.allreduce_sum_(self.gflat, group=self.pg)
            _lib.check(L.pdes_adam_step_dev(_lib.ptr(self.flat), _lib.ptr(self.gflat), _lib.ptr(self.m),
                                            _lib.ptr(self.v), self.flat.numel(), _lib.ptr(self.hyper_dev), st),
                       "pdes_adam_step_dev")
            n += 1
        self.model._flat_nbt.add_(1)
        self.kernel_launches = n
        return torch.dot(self.l4, self.gw)

    def capture(self, K_example):
        """Record the whole step into one CUDA graph (single-GPU: including Adam; data-parallel: up to
        the gradient bucket, the NCCL all-reduce and Adam follow eagerly)."""
        self.static_K = torch.empty_like(K_example)
        self.static_K.copy_(K_example)
        # The warm-up (lazily-set kernel attributes, workspaces, tensor maps) runs the real device work,
        # Adam and BatchNorm running-statistics updates included: snapshot everything a step mutates and
        # restore it afterwards, so that replay i IS optimisation step i of the eager / reference trajectory.
        m = self.model
        snap = [t.clone() for t in (self.flat, self.m, self.v, m._flat_running, m._flat_nbt)]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                self._set_hyper(0.0)
                self._device_work(self.static_K)
        torch.cuda.current_stream().wait_stream(side)
        for t, s0 in zip((self.flat, self.m, self.v, m._flat_running, m._flat_nbt), snap):
            t.copy_(s0)
        torch.cuda.synchronize()
        try:
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.static_loss = self._device_work(self.static_K)
        except Exception:
            if not (self.world > 1 and self.collective_in_graph):
                raise
            # the collective refused capture: keep it (and Adam) eager behind the replayed backward
            self.collective_in_graph = False
            torch.cuda.synchronize()
            for t, s0 in zip((self.flat, self.m, self.v, m._flat_running, m._flat_nbt), snap):
                t.copy_(s0)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.static_loss = self._device_work(self.static_K)

    def _set_hyper(self, lr):
        L = _lib.lib()
        i = self._hyper_i % self._hyper_slots
        self._hyper_i += 1
        if self._hyper_ev[i] is not None:
            self._hyper_ev[i].synchronize()   # the copy that last read this slot has executed
        slot = self.hyper_host[i]
        _lib.check(L.pdes_adam_hyper(slot.data_ptr(), float(lr), self.betas[0], self.betas[1], self.eps,
                                     self.wd, 1.0 / self.world, self.steps + 1), "pdes_adam_hyper")
        self.hyper_dev.copy_(slot, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._hyper_ev[i] = ev

    def step_graph(self, K, lr=None):
        """Replay the captured step on batch K (copied into the graph's static input)."""
        if self.graph is None:
            self.capture(K)
        self.static_K.copy_(K, non_blocking=True)
        self._set_hyper(self.lr if lr is None else lr)
        self.graph.replay()
        self.steps += 1
        if self.world > 1 and not self.collective_in_graph:
            L = _lib.lib()
            ddp.allreduce_sum_(self.gflat, group=self.pg)
            _lib.check(L.pdes_adam_step_dev(_lib.ptr(self.flat), _lib.ptr(self.gflat), _lib.ptr(self.m),
                                            _lib.ptr(self.v), self.flat.numel(), _lib.ptr(self.hyper_dev),
                                            _lib.stream_ptr()), "pdes_adam_step_dev")
        return self.static_loss

    def step(self, K, lr=None):
        """One optimisation step on the (B,1,H,W) device batch K; returns the 0-d device loss."""
        L = _lib.lib()
        ex = self.model._ex
        st = _lib.stream_ptr()
        B, _, H, W = K.shape
        self.gflat.zero_()
        out = ex.forward(K, True)
        n = L.pdes_densenet_last_launches(ex.handle.h)
        _lib.check(L.pdes_darcy_loss_fwd(_lib.ptr(K), _lib.ptr(out), B, H, W, 1, _lib.ptr(self.l4),
                                         _lib.ptr(_darcy._workspace(K.device)), st), "pdes_darcy_loss_fwd")
        dout = torch.empty_like(out)
        _lib.check(L.pdes_darcy_loss_bwd(_lib.ptr(K), _lib.ptr(out), _lib.ptr(self.gw), B, H, W, 1,
                                         _lib.ptr(dout), st), "pdes_darcy_loss_bwd")
        ex.backward(dout)
        n += L.pdes_densenet_last_launches(ex.handle.h) + 2
        if self.world > 1:
            ddp.allreduce_sum_(self.gflat, group=self.pg)
        self.steps += 1
        _lib.check(L.pdes_adam_step(_lib.ptr(self.flat), _lib.ptr(self.gflat), _lib.ptr(self.m),
                                    _lib.ptr(self.v), self.flat.numel(), float(self.lr if lr is None else lr),
                                    self.betas[0], self.betas[1], self.eps, self.wd, 1.0 / self.world,
                                    self.steps, st), "pdes_adam_step")
        self.model._flat_nbt.add_(1)
        self.kernel_launches = n + 1
        return torch.dot(self.l4, self.gw)
