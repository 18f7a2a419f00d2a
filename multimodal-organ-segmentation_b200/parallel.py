"""Data-parallel gradient exchange for training (one process per GPU, torch.distributed: NCCL over NVLink on the GPUs,
gloo in the CPU tests).

InstanceNorm statistics are per sample and both DiceCE terms are means over equal-sized samples, so the mean over ranks
of the per-rank gradients IS the global-batch gradient: the only collective of a training step is one
all-reduce(sum)/world of the gradients.  `GradBucketReducer` packs them into flat fp32 buckets in backward-execution
order (out_conv, decoder, ..., encoder).  When armed, the kernel backward (train_engine) hands every weight gradient to
`grad_ready` the moment its wgrad/reduce kernels have been enqueued; a bucket whose last gradient arrived is all-reduced
on a side stream behind a CUDA event, so the exchange overlaps the dgrad / wgrad kernels still to come.  `finish()`
joins the streams and writes the averaged gradients into `.grad`.  With gradient accumulation only the stepping
micro-batch is armed (earlier micro-batches accumulate locally, exactly like the reference's loop).
"""
from typing import Dict, List, Optional

import torch
import torch.distributed as dist


class GradBucketReducer:
    def __init__(self, model: torch.nn.Module, bucket_bytes: int = 32 << 20, group=None):
        self.model, self.group = model, group
        self.world = dist.get_world_size(group)
        params = [p for p in model.parameters() if p.requires_grad][::-1]
        self.buckets: List[List[torch.nn.Parameter]] = []
        cur, size = [], 0
        for p in params:
            cur.append(p)
            size += p.numel() * 4
            if size >= bucket_bytes:
                self.buckets.append(cur)
                cur, size = [], 0
        if cur:
            self.buckets.append(cur)
        self.slot: Dict[torch.nn.Parameter, tuple] = {}
        for i, b in enumerate(self.buckets):
            off = 0
            for p in b:
                self.slot[p] = (i, off)
                off += p.numel()
        self._flat: List[Optional[torch.Tensor]] = [None] * len(self.buckets)
        self._ready = [0] * len(self.buckets)
        self._works: list = []
        self._stream = None
        self._producers: List[set] = [set() for _ in self.buckets]   # CUDA streams that packed into bucket i
        self.armed = False

    # ------------------------------------------------------------------ helpers
    def _flat_for(self, i: int, device) -> torch.Tensor:
        n = sum(p.numel() for p in self.buckets[i])
        flat = self._flat[i]
        if flat is None or flat.device != device:
            flat = self._flat[i] = torch.zeros(n, dtype=torch.float32, device=device)
        return flat

    def _launch(self, i: int) -> None:
        flat = self._flat[i]
        if flat.is_cuda:
            if self._stream is None:
                self._stream = torch.cuda.Stream(device=flat.device)
            # only this bucket's producers (the backward's main stream and its wgrad side stream), not the rest of
            # the backward
            for st in self._producers[i] | {torch.cuda.current_stream(flat.device)}:
                ev = torch.cuda.Event()
                ev.record(st)
                self._stream.wait_event(ev)
            self._producers[i] = set()
            with torch.cuda.stream(self._stream):
                self._works.append(dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        else:
            self._works.append(dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    # ------------------------------------------------------------------ overlapped path (called by train_engine)
    def arm(self) -> None:
        self.armed = True
        self._ready = [0] * len(self.buckets)
        self._producers = [set() for _ in self.buckets]
        self._works = []

    @torch.no_grad()
    def grad_ready(self, p: torch.nn.Parameter, g: torch.Tensor) -> None:
        """g: this micro-batch's gradient of p (fp32).  Adds the locally accumulated p.grad, packs, maybe launches."""
        i, off = self.slot[p]
        flat = self._flat_for(i, g.device)
        dst = flat[off:off + p.numel()]
        dst.copy_(g.reshape(-1))
        if p.grad is not None:
            dst.add_(p.grad.reshape(-1).float())
        if g.is_cuda:
            self._producers[i].add(torch.cuda.current_stream(g.device))
        self._ready[i] += 1
        if self._ready[i] == len(self.buckets[i]):
            self._launch(i)

    @torch.no_grad()
    def finish(self) -> None:
        """Join: flush buckets that never filled (parameters without a gradient), wait, write averaged .grad."""
        for i, b in enumerate(self.buckets):
            if self._ready[i] != len(b):
                if self._flat[i] is None:
                    self._flat_for(i, next(self.model.parameters()).device)
                self._launch(i)
        for w in self._works:
            w.wait()
        if self._stream is not None:
            torch.cuda.current_stream().wait_stream(self._stream)
        inv = 1.0 / self.world
        for i, b in enumerate(self.buckets):
            flat = self._flat[i]
            for p in b:
                _, off = self.slot[p]
                g = flat[off:off + p.numel()].view_as(p)
                if p.grad is None:
                    p.grad = (g * inv).to(p.dtype)
                else:
                    p.grad.copy_(g * inv)
        self.armed = False
        self._works = []

    # ------------------------------------------------------------------ non-overlapped path (.grad already populated)
    @torch.no_grad()
    def reduce_gradients(self) -> None:
        self.arm()
        for b in self.buckets:
            for p in b:
                g = p.grad if p.grad is not None else torch.zeros_like(p)
                pg, p.grad = p.grad, None          # grad_ready adds p.grad: hand the total over exactly once
                self.grad_ready(p, g.float())
                p.grad = pg
        self.finish()
