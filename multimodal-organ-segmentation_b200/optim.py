"""Fused multi-tensor AdamW on the sm_100a kernel `mmseg_adamw_multi` (csrc/optim.cu) — SURVEY.md §8(f) N1.

Drop-in for the `torch.optim.AdamW` the reference builds (src/trainer/trainer.py:115-117) and steps (:245-248): same
constructor arguments, same update rule (decoupled weight decay, bias correction, eps outside the sqrt), same
state_dict layout (per parameter `step`, `exp_avg`, `exp_avg_sq`), so reference checkpoints' optimizer states load and
ours load into torch.  One launch updates every parameter of a group; hyper-parameters and the step counter live on the
device, so the step is CUDA-graph capturable and follows learning-rate schedulers (`sync_hyper`).
"""
import ctypes as C
from typing import Dict, List, Optional

import torch

from . import _lib
from . import kernels as K


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2,
                 capturable: bool = True):
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1) or not (0 <= betas[1] < 1) or weight_decay < 0:
            raise ValueError("invalid AdamW hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay))
        self._plans: Dict[int, dict] = {}          # per param group: device tables, keyed by the gradient addresses
        self._hyper: Dict[int, dict] = {}          # per param group: device hyper-parameter array + the values it holds
        self._captured: List[dict] = []
        self.grad_scale = 1.0

    # ------------------------------------------------------------------ state
    def _init_state(self, p: torch.Tensor) -> dict:
        st = self.state[p]
        if len(st) == 0:
            # like torch.optim.AdamW(capturable=True): a device-side fp32 step counter PER parameter (a parameter that
            # receives no gradient in some step does not advance)
            st["step"] = torch.zeros((), dtype=torch.float32, device=p.device)
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        return st

    def _group_plan(self, gi: int, group: dict) -> Optional[dict]:
        params = [p for p in group["params"] if p.grad is not None]
        if not params:
            return None
        for p in params:
            if p.dtype != torch.float32 or not p.is_contiguous() or not p.is_cuda:
                raise RuntimeError("FusedAdamW updates contiguous fp32 CUDA parameters (there is no CPU fallback)")
            if p.grad.dtype != torch.float32 or not p.grad.is_contiguous():
                p.grad = p.grad.float().contiguous()
        states = [self._init_state(p) for p in params]
        key = tuple((p.data_ptr(), p.grad.data_ptr(), st["step"].data_ptr()) for p, st in zip(params, states))
        capturing = torch.cuda.is_current_stream_capturing()
        plan = self._plans.get(gi)
        if plan is not None and plan["key"] == key and plan["capturing"] == capturing:
            return plan
        dev = params[0].device
        tens = (_lib.AdamwTensor * len(params))()
        chunks: List[int] = []
        for i, (p, st) in enumerate(zip(params, states)):
            if st["step"].dtype != torch.float32 or st["step"].device != dev:   # e.g. a checkpoint written on the CPU
                st["step"] = st["step"].to(dev, torch.float32)
            tens[i].p, tens[i].g = p.data_ptr(), p.grad.data_ptr()
            tens[i].m, tens[i].v, tens[i].step = st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), st["step"].data_ptr()
            tens[i].n = p.numel()
            for c in range((p.numel() + _lib.ADAMW_CHUNK - 1) // _lib.ADAMW_CHUNK):
                chunks += [i, c]
        # pinned staging: the host->device copies of the tables are stream-ordered (and capturable); a capture gets its
        # own staging buffers so that a later eager step cannot change what a replay uploads
        t_host = torch.frombuffer(bytearray(bytes(tens)), dtype=torch.uint8).pin_memory()
        c_host = torch.tensor(chunks, dtype=torch.int32).pin_memory()
        hyper = self._hyper.get(gi)
        if hyper is None:
            hyper = self._hyper[gi] = {"dev": torch.zeros(6, dtype=torch.float32, device=dev), "vals": None}
        plan = {"key": key, "capturing": capturing, "n_chunks": len(chunks) // 2, "n_tensors": len(params),
                "tensors": t_host.to(dev, non_blocking=True), "chunks": c_host.to(dev, non_blocking=True),
                "_keep": (t_host, c_host)}
        if capturing:
            self._captured.append(plan)     # the graph replays these uploads: keep the staging buffers alive
        self._plans[gi] = plan
        return plan

    def load_state_dict(self, state_dict) -> None:
        super().load_state_dict(state_dict)
        self._plans.clear()          # the loaded state tensors are new objects

    def sync_hyper(self) -> None:
        """Push the groups' current hyper-parameters (lr under a scheduler, grad_scale) to the device arrays the kernel
        reads.  Called by step(); call it yourself before replaying a captured graph after the values changed."""
        for gi, group in enumerate(self.param_groups):
            hyper = self._hyper.get(gi)
            if hyper is None:
                continue
            vals = (float(group["lr"]), float(group["betas"][0]), float(group["betas"][1]), float(group["eps"]),
                    float(group["weight_decay"]), float(self.grad_scale))
            if vals != hyper["vals"]:
                if torch.cuda.is_current_stream_capturing():
                    raise RuntimeError("hyper-parameters changed inside a CUDA-graph capture; call sync_hyper() before")
                hyper["dev"].copy_(torch.tensor(vals, dtype=torch.float32))
                hyper["vals"] = vals

    # ------------------------------------------------------------------ step
    @torch.no_grad()
    def step(self, closure=None, zero_grad: bool = False):
        """One AdamW update of every parameter that has a gradient; zero_grad=True also zeroes the gradients in the same
        pass (optimizer.zero_grad(set_to_none=False) for free)."""
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        _lib.require_device()
        for gi, group in enumerate(self.param_groups):
            plan = self._group_plan(gi, group)
            if plan is None:
                continue
            self.sync_hyper()
            K._call("mmseg_adamw_multi", C.c_void_p(plan["tensors"].data_ptr()), plan["n_tensors"],
                    C.c_void_p(plan["chunks"].data_ptr()), plan["n_chunks"], C.c_void_p(self._hyper[gi]["dev"].data_ptr()),
                    1 if zero_grad else 0, K._stream())
            # parameters changed in place behind autograd's back: bump the version counters that the packed-weight
            # caches and the captured inference graphs key on
            torch.autograd.graph.increment_version([p for p in group["params"] if p.grad is not None])
        return loss
