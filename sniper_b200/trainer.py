"""Training-step driver: the `Module.forward_backward` / `update` / `update_metric` inner loop of
BaseModule.fit (SNIPER-mxnet/python/mxnet/module/base_module.py:505-535) for the SNIPER graph.

One process per GPU.  Per step: pinned host batch -> H2D (on a copy stream, one step ahead when the caller names the
next batch: `step(batch, prefetch=next_batch)`, as MNIteratorE2E's prefetch threads do), one CUDA-graph replay of
forward+backward, ONE NCCL all-reduce over the flat gradient bucket (replaces the per-key kvstore
push/pull, model.py:126-136 / comm.h:433-553; sum without 1/N as rescale_grad=1.0, utils.py:30,37), one
CUDA-graph replay of the fused SGD-momentum update, D2H of the four loss scalars (what the reference's
metrics read with asnumpy(), lib/train_utils/metric.py:109-125).
"""
import torch

from . import model, ops


class Trainer:
    def __init__(self, cfg=None, device="cuda:0", world_size=1, use_graph=True, seed=5, deform_offset_std=0.0):
        self.cfg = cfg or model.Cfg()
        self.device = torch.device(device)
        torch.cuda.set_device(self.device)
        self.world_size = world_size
        self.net = model.SniperResNet101(self.cfg, device=self.device, seed=seed, deform_offset_std=deform_offset_std)
        self.use_graph = use_graph
        self.static = None
        self.g_fb = None
        self.g_up = None
        self.out = None
        self.loss_host = torch.zeros(8).pin_memory()
        self.launches_per_step = 0
        # input pipeline: two device staging sets filled by a copy stream while the previous step computes
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.stage = [None, None]
        self.stage_ready = [None, None]      # event: H2D into stage[i] finished
        self.stage_free = [None, None]       # event: the step that consumed stage[i] has copied it out
        self.stage_owner = [None, None]      # the host batch object sitting in stage[i] (kept alive: identity match)
        self.next_stage = 0

    # ---- device-resident step (inputs already in HBM)
    def _alloc_static(self, batch):
        self.static = {k: torch.empty_like(v, device=self.device) for k, v in batch.items()}

    def load(self, host_batch):
        """H2D of one chip batch (pinned host tensors) into the static device buffers."""
        if self.static is None:
            self._alloc_static(host_batch)
        for k, v in host_batch.items():
            self.static[k].copy_(v, non_blocking=True)

    def prefetch(self, host_batch):
        """Starts the H2D of a FUTURE batch on the copy stream (overlaps the step in flight)."""
        i = self.next_stage
        self.next_stage ^= 1
        if self.stage[i] is None:
            self.stage[i] = {k: torch.empty_like(v, device=self.device) for k, v in host_batch.items()}
        if self.stage_free[i] is not None:
            self.copy_stream.wait_event(self.stage_free[i])
        with torch.cuda.stream(self.copy_stream):
            for k, v in host_batch.items():
                self.stage[i][k].copy_(v, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        self.stage_ready[i] = ev
        self.stage_owner[i] = host_batch

    def _load_or_take(self, host_batch):
        """Brings `host_batch` into the static buffers: from its staging set if it was prefetched, else by H2D now."""
        for i in (0, 1):
            if self.stage_owner[i] is host_batch and self.stage_ready[i] is not None:
                if self.static is None:
                    self._alloc_static(host_batch)
                cur = torch.cuda.current_stream()
                cur.wait_event(self.stage_ready[i])
                for k in self.static:
                    self.static[k].copy_(self.stage[i][k], non_blocking=True)     # D2D, ~25 us
                ev = torch.cuda.Event()
                ev.record(cur)
                self.stage_free[i] = ev
                self.stage_owner[i] = None
                return
        self.load(host_batch)

    def _allreduce(self):
        if self.world_size > 1:
            torch.distributed.all_reduce(self.net.P.g, op=torch.distributed.ReduceOp.SUM)

    def capture(self):
        """Warm up eagerly, then capture forward+backward and the optimizer update as two CUDA graphs."""
        s = torch.cuda.Stream(device=self.device)
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):
                ops.reset_launch_count()
                self.out = self.net.forward_backward(self.static)
                self._allreduce()
                self.net.update()
                self.launches_per_step = ops.launch_count()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        if not self.use_graph:
            return
        self.g_fb = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.g_fb):
            self.out = self.net.forward_backward(self.static)
        self.g_up = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.g_up):
            self.net.update()
        torch.cuda.synchronize()

    def step_device(self):
        """One training step on the batch currently in the static buffers."""
        if self.g_fb is not None:
            self.g_fb.replay()
            self._allreduce()
            self.g_up.replay()
        else:
            self.out = self.net.forward_backward(self.static)
            self._allreduce()
            self.net.update()
        return self.out

    # ---- public end-to-end step: host batch in, host losses out
    def step(self, host_batch, prefetch=None):
        """One training step on `host_batch` (pinned host tensors) -> host loss scalars.  `prefetch`: the batch of the
        NEXT call; its H2D copy is issued now on the copy stream and overlaps this step's compute."""
        self._load_or_take(host_batch)
        if prefetch is not None:
            self.prefetch(prefetch)
        if self.g_fb is None and self.use_graph:
            self.capture()
        out = self.step_device()
        self.loss_host.copy_(out["losses"], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return {"rpn_cls_loss": float(self.loss_host[0]), "rpn_bbox_loss": float(self.loss_host[1]),
                "rcnn_cls_loss": float(self.loss_host[2]), "rcnn_bbox_loss": float(self.loss_host[3])}
