"""Fused flat-buffer AdamW + flat gradient all-reduce.

Drop-in for the optimiser `utils/training.py:48-50` builds (`torch.optim.AdamW(params, lr=..., weight_decay=...)`):
same constructor arguments, `param_groups` (so `_anneal_lr`, training.py:84-90, keeps working), `zero_grad()`, `step()`,
`state_dict()` / `load_state_dict()` in torch's own format (training.py:70-82,105-106 save and resume `opt.pt`).

What is different underneath: every parameter, its gradient and both Adam moments are views into four flat fp32 buffers,
so one `am_adamw_flat` launch replaces the ~100 foreach kernels per step, `zero_grad()` is one memset, and data-
parallel training all-reduces ONE flat gradient buffer (`all_reduce_grads`) instead of DDP's bucketed graph traversal
(train_ddp.py:63-65, `find_unused_parameters=True`).  CUDA only: there is no CPU fallback."""
from typing import Iterable, Optional

import torch
import torch.distributed as dist

from . import lib as _l


def _align(n: int, a: int = 4) -> int:
    return (n + a - 1) // a * a


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params: Iterable, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self._flat = []  # per param group: dict(p, g, m, v, n, offsets)
        for group in self.param_groups:
            ps = [p for p in group["params"] if p.requires_grad]
            if not ps:
                self._flat.append(None)
                continue
            dev = ps[0].device
            if dev.type != "cuda":
                raise _l.AmbError("FusedAdamW is CUDA-only (sm_100a): move the model to the GPU before building the optimiser")
            for p in ps:
                if p.dtype != torch.float32 or p.device != dev:
                    raise _l.AmbError("FusedAdamW: all parameters of a group must be fp32 on one device")
            offs, n = [], 0
            for p in ps:
                offs.append(n)
                n += _align(p.numel())  # every parameter starts 16-byte aligned
            fp = torch.zeros(n, device=dev)
            fg, fm, fv = torch.zeros_like(fp), torch.zeros_like(fp), torch.zeros_like(fp)
            for p, o in zip(ps, offs):
                k = p.numel()
                fp[o:o + k].copy_(p.data.reshape(-1))
                had_grad = p.grad is not None
                if had_grad:
                    fg[o:o + k].copy_(p.grad.reshape(-1))
                p.data = fp[o:o + k].view(p.shape)  # parameters become views of the flat buffer (same values, same Parameter objects)
                p.grad = fg[o:o + k].view(p.shape)  # autograd accumulates in place into the flat gradient buffer
                self.state[p] = {"step": torch.tensor(0.0), "exp_avg": fm[o:o + k].view(p.shape), "exp_avg_sq": fv[o:o + k].view(p.shape)}
            self._flat.append(dict(p=fp, g=fg, m=fm, v=fv, n=n, params=ps, offs=offs, step=0))

    # ------------------------------------------------------------------ torch.optim.Optimizer surface
    def zero_grad(self, set_to_none: bool = True):
        """One memset of the flat gradient buffer.  Gradients stay views of it (never None), whatever `set_to_none` says, so
        the next backward accumulates straight into the buffer the fused step and the all-reduce read."""
        for f in self._flat:
            if f is None:
                continue
            f["g"].zero_()
            for p, o in zip(f["params"], f["offs"]):  # re-attach if user code dropped or replaced a view
                if p.grad is None or p.grad.data_ptr() != f["g"].data_ptr() + 4 * o:
                    p.grad = f["g"][o:o + p.numel()].view(p.shape)

    @torch.no_grad()
    def step(self, closure=None, grad_scale: Optional[float] = None):
        if grad_scale is None:  # 1/world after all_reduce_grads(), else 1
            grad_scale = getattr(self, "_pending_scale", 1.0)
        self._pending_scale = 1.0
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _l.load()
        st = torch.cuda.current_stream().cuda_stream
        for group, f in zip(self.param_groups, self._flat):
            if f is None:
                continue
            # every p.grad must still be a view of the flat gradient buffer (zero_grad(set_to_none=True) from user code or DDP's
            # gradient_as_bucket_view detach it): copy stragglers in, treat a missing gradient as zero like the flat kernel does
            for p, o in zip(f["params"], f["offs"]):
                gview = f["g"][o:o + p.numel()]
                if p.grad is None:
                    gview.zero_()
                elif p.grad.data_ptr() != f["g"].data_ptr() + 4 * o:
                    gview.copy_(p.grad.reshape(-1))
                    p.grad = gview.view(p.shape)
            f["step"] += 1
            b1, b2 = group["betas"]
            _l.check(lib.am_adamw_flat(f["p"].data_ptr(), f["g"].data_ptr(), f["m"].data_ptr(), f["v"].data_ptr(), f["n"],
                                       float(group["lr"]), float(b1), float(b2), float(group["eps"]), float(group["weight_decay"]),
                                       f["step"], float(grad_scale), 0, st), "am_adamw_flat")
            # the kernel wrote the parameters behind torch's back: bump their version counters so the sampling engines
            # (amb200.pack.params_version) re-pack their folded / bf16-split weight copies
            torch.autograd.graph.increment_version(f["params"])
            step_t = torch.tensor(float(f["step"]))
            for p in f["params"]:
                self.state[p]["step"] = step_t
        return loss

    def load_state_dict(self, state_dict):
        """Accepts a torch.optim.AdamW state_dict (opt.pt written by utils/training.py:105-106): moments are copied INTO the
        flat views, the step counter is restored."""
        groups = state_dict["param_groups"]
        for group, saved in zip(self.param_groups, groups):
            for k, v in saved.items():
                if k != "params":
                    group[k] = v
        flat_ids = [pid for g in groups for pid in g["params"]]
        mine = [p for g in self.param_groups for p in g["params"]]
        for pid, p in zip(flat_ids, mine):
            s = state_dict["state"].get(pid)
            if s is None or p not in self.state:
                continue
            self.state[p]["exp_avg"].copy_(s["exp_avg"])
            self.state[p]["exp_avg_sq"].copy_(s["exp_avg_sq"])
            self.state[p]["step"] = torch.tensor(float(s["step"]))
        for f in self._flat:
            if f is not None:
                steps = {int(self.state[p]["step"]) for p in f["params"]}
                f["step"] = max(steps) if steps else 0

    # ------------------------------------------------------------------ data parallel (one exchange step)
    def flat_grads(self):
        return [f["g"] for f in self._flat if f is not None]

    def all_reduce_grads(self, group: Optional[dist.ProcessGroup] = None):
        """Data-parallel gradient exchange: all-reduce(sum) of the flat buffer (48.8 MB for CMDM); the division by the
        world size is folded into the next fused step (`grad_scale`), so the mean costs no extra pass over the buffer.
        After `begin_overlap()` the pieces whose gradients were complete during backward are already in flight: this call
        launches the rest (pieces holding parameters that received no gradient) and waits for all of them."""
        ov = getattr(self, "_overlap", None)
        if ov is not None and ov.armed:
            world = ov.finish()
        else:
            from .dist import allreduce_flat_
            world = allreduce_flat_(self.flat_grads(), group)
        self._pending_scale = 1.0 / world
        return world

    # ---- exchange overlapped with backward (train_ddp.py:63-65 gets this from DDP's buckets)
    def enable_overlap(self, chunks: int = 3, group: Optional[dist.ProcessGroup] = None) -> int:
        """Hook the parameters so that pieces of the flat gradient buffer are all-reduced while backward is still running
        (`amb200.dist.ChunkedGradExchange`).  Call `begin_overlap()` before the final backward of a step and
        `all_reduce_grads()` after it, as before.  Returns the number of pieces."""
        from .dist import ChunkedGradExchange
        self._overlap = ChunkedGradExchange([(f["g"], f["params"], f["offs"], f["n"]) for f in self._flat if f is not None], chunks, group)
        return len(self._overlap.chunks)

    def begin_overlap(self):
        self._overlap.begin()
