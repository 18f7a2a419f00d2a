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
        self._hyper_host: Dict[int, list] = {}
        self.grad_scale = 1.0

    # ------------------------------------------------------------------ state
    def _init_state(self, p: torch.Tensor) -> dict:
        st = self.state[p]
        if len(st) == 0:
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
        key = tuple((p.data_ptr(), p.grad.data_ptr()) for p in params)
        capturing = torch.cuda.is_current_stream_capturing()
        plan = self._plans.get(gi)
        if plan is not None and plan["key"] == key and plan["capturing"] == capturing:
            return plan
        dev = params[0].device
        states = [self._init_state(p) for p in params]
        # ONE device step counter per group: every parameter's `step` is a view of it (they advance together)
        shared = plan["step"] if plan is not None else None
        if shared is None or any(st["step"].data_ptr() != shared.data_ptr() for st in states):
            # first build, or load_state_dict replaced the per-parameter counters (equal values): unify them
            shared = torch.full((1,), float(states[0]["step"]), dtype=torch.float32, device=dev)
            for st in states:
                st["step"] = shared.view(())
        step = shared
        tens = (_lib.AdamwTensor * len(params))()
        chunks: List[int] = []
        for i, (p, st) in enumerate(zip(params, states)):
            tens[i].p, tens[i].g = p.data_ptr(), p.grad.data_ptr()
            tens[i].m, tens[i].v, tens[i].n = st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), p.numel()
            for c in range((p.numel() + _lib.ADAMW_CHUNK - 1) // _lib.ADAMW_CHUNK):
                chunks += [i, c]
        raw = bytes(tens)
        # pinned staging: the host->device copies of the tables are stream-ordered (and capturable)
        t_host = torch.frombuffer(bytearray(raw), dtype=torch.uint8).pin_memory()
        c_host = torch.tensor(chunks, dtype=torch.int32).pin_memory()
        plan = {"key": key, "capturing": capturing, "n_chunks": len(chunks) // 2, "step": step,
                "tensors": t_host.to(dev, non_blocking=True), "chunks": c_host.to(dev, non_blocking=True),
                "hyper": torch.zeros(6, dtype=torch.float32, device=dev), "hyper_vals": None,
                "_keep": (t_host, c_host)}
        self._plans[gi] = plan
        return plan

    def load_state_dict(self, state_dict) -> None:
        super().load_state_dict(state_dict)
        self._plans.clear()          # the loaded state tensors are new objects

    def sync_hyper(self) -> None:
        """Push the groups' current hyper-parameters (lr under a scheduler, grad_scale) to the device arrays the kernel
        reads.  Called by step(); call it yourself before replaying a captured graph after the values changed."""
        for gi, group in enumerate(self.param_groups):
            plan = self._plans.get(gi)
            if plan is None:
                continue
            vals = (float(group["lr"]), float(group["betas"][0]), float(group["betas"][1]), float(group["eps"]),
                    float(group["weight_decay"]), float(self.grad_scale))
            if vals != plan["hyper_vals"]:
                if torch.cuda.is_current_stream_capturing():
                    raise RuntimeError("hyper-parameters changed inside a CUDA-graph capture; call sync_hyper() before")
                plan["hyper"].copy_(torch.tensor(vals, dtype=torch.float32))
                plan["hyper_vals"] = vals

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
            K._call("mmseg_adamw_multi", C.c_void_p(plan["tensors"].data_ptr()), C.c_void_p(plan["chunks"].data_ptr()),
                    plan["n_chunks"], C.c_void_p(plan["step"].data_ptr()), C.c_void_p(plan["hyper"].data_ptr()),
                    1 if zero_grad else 0, K._stream())
            # parameters changed in place behind autograd's back: bump the version counters that the packed-weight
            # caches and the captured inference graphs key on
            torch.autograd.graph.increment_version([p for p in group["params"] if p.grad is not None])
        return loss
