"""AdamW for the training step: bf16 parameters, fp32 master weights and moments, one multi-tensor kernel per step.

The reference's optimizer is DeepSpeed FusedAdam in adam_w_mode (VisualRWKV-v7/v7.00/src/model.py:376-410) under
`--precision bf16`; this is the same update (torch.optim.AdamW's formula) through csrc/optim.cu: one pass over the
gradients and the state instead of stock PyTorch's cast / fused step / cast-back."""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib


class FusedAdamW:
    """step() updates `params` (bf16, in place) from their `.grad` (bf16).  CUDA-graph capturable: the step counter lives
    on the device; the table of gradient pointers is re-uploaded (pinned -> device copy on the current stream) whenever a
    `.grad` tensor moved."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, grad_scale=1.0):
        self.params = [p for p in params if p.requires_grad]
        assert self.params and all(p.dtype == torch.bfloat16 and p.is_cuda and p.is_contiguous() for p in self.params)
        self.lr, self.betas, self.eps, self.weight_decay, self.grad_scale = lr, betas, eps, weight_decay, grad_scale
        dev = self.params[0].device
        self.chunk = _lib.lib().vrwkv_adamw_chunk()
        offs, total = [], 0
        for p in self.params:
            offs.append(total)
            total += (p.numel() + 3) // 4 * 4          # 16-byte aligned fp32 slices
        self.offsets = offs
        self.master = torch.zeros(total, dtype=torch.float32, device=dev)
        for p, o in zip(self.params, offs):
            self.master[o:o + p.numel()].copy_(p.detach().reshape(-1).float())
        self.exp_avg = torch.zeros_like(self.master)
        self.exp_avg_sq = torch.zeros_like(self.master)
        self.step_count = torch.zeros(1, dtype=torch.int32, device=dev)
        chunks = [(ti, ci) for ti, p in enumerate(self.params) for ci in range((p.numel() + self.chunk - 1) // self.chunk)]
        self.nchunks = len(chunks)
        self.chunks = torch.tensor(chunks, dtype=torch.int32, device=dev)
        self._table_host = torch.zeros(len(self.params), 4, dtype=torch.int64).pin_memory()
        self._table = torch.zeros(len(self.params), 4, dtype=torch.int64, device=dev)
        self._grad_ptrs = None

    def master_of(self, i):
        p = self.params[i]
        return self.master[self.offsets[i]:self.offsets[i] + p.numel()].view_as(p)

    def _refresh_table(self):
        ptrs = []
        for p in self.params:
            g = p.grad
            if g is None:
                raise RuntimeError("FusedAdamW.step: a parameter has no gradient")
            if g.dtype != torch.bfloat16 or not g.is_contiguous():
                raise RuntimeError("FusedAdamW.step: gradients must be contiguous bf16")
            ptrs.append(g.data_ptr())
        if ptrs != self._grad_ptrs:
            t = self._table_host.numpy()
            t[:, 0] = np.array([p.data_ptr() for p in self.params], dtype=np.int64)
            t[:, 1] = np.array(ptrs, dtype=np.int64)
            t[:, 2] = np.array(self.offsets, dtype=np.int64)
            t[:, 3] = np.array([p.numel() for p in self.params], dtype=np.int64)
            # Outside a capture the upload blocks the host: the pinned staging table is rewritten by the next step that sees
            # moved gradients, and an asynchronous copy still queued behind a slow GPU would then read the NEXT step's
            # pointers.  Inside a capture the copy becomes a graph node (addresses are static from then on).
            self._table.copy_(self._table_host, non_blocking=torch.cuda.is_current_stream_capturing())
            self._grad_ptrs = ptrs

    @torch.no_grad()
    def step(self):
        self._refresh_table()
        L = _lib.lib()
        lr_dev = self.lr if torch.is_tensor(self.lr) else None
        rc = L.vrwkv_adamw_step(ctypes.c_int(self.nchunks), ctypes.c_void_p(self._table.data_ptr()), ctypes.c_void_p(self.chunks.data_ptr()),
                                ctypes.c_void_p(self.master.data_ptr()), ctypes.c_void_p(self.exp_avg.data_ptr()),
                                ctypes.c_void_p(self.exp_avg_sq.data_ptr()), ctypes.c_void_p(self.step_count.data_ptr()),
                                ctypes.c_void_p(lr_dev.data_ptr()) if lr_dev is not None else None,
                                ctypes.c_float(0.0 if lr_dev is not None else float(self.lr)), ctypes.c_float(self.betas[0]),
                                ctypes.c_float(self.betas[1]), ctypes.c_float(self.eps), ctypes.c_float(self.weight_decay),
                                ctypes.c_float(self.grad_scale), _lib.cur_stream())
        if rc != 0:
            raise RuntimeError(f"vrwkv_adamw_step failed ({rc}): {L.vrwkv_last_error().decode()}")
