"""Data-parallel gradient exchange: one process per GPU, a bucketed all-reduce of the
gradients over NCCL (NVLink 5 / NVSwitch) launched from autograd hooks so it overlaps the backward of
earlier layers.  Replaces what DeepSpeed ZeRO-1 does implicitly for the reference
(v7.00/train.py:55,214-216: reduce-scatter + all-gather in 200 MB buckets).

Gradients live as views into flat per-bucket buffers (no gather/scatter copies); a bucket is reduced
as soon as the last of its parameters has accumulated its gradient.  The path shards over the batch
only; there is no other collective on the data path (loss logging gathers one scalar per rank,
v7.00/src/model.py:436-440).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def _al(n: int) -> int:
    return (n + 7) // 8 * 8


class GradBucketReducer:
    def __init__(self, params, bucket_bytes: int = 32 << 20, process_group=None):
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.params = [p for p in params if p.requires_grad]
        # backward produces gradients roughly in reverse registration order
        order = list(reversed(self.params))
        self.buckets = []  # (flat, [params])
        cur, cur_bytes = [], 0
        for p in order:
            nb = p.numel() * p.element_size()
            if cur and (cur_bytes + nb > bucket_bytes or p.dtype != cur[0].dtype or p.device != cur[0].device):
                self._close(cur)
                cur, cur_bytes = [], 0
            cur.append(p)
            cur_bytes += nb
        if cur:
            self._close(cur)
        self._pending = [0] * len(self.buckets)
        self._handles = []
        self.overlap = True   # False: hooks do nothing, the caller runs allreduce_all() after backward (CUDA-graphed backward)
        self._bucket_of = {}
        for bi, (_, ps) in enumerate(self.buckets):
            for p in ps:
                self._bucket_of[p] = bi
                p.register_post_accumulate_grad_hook(self._hook)
        self.reset()

    def _close(self, ps):
        n = sum(_al(p.numel()) for p in ps)   # every gradient view starts on a 16-byte boundary (vector loads in the optimizer kernel)
        flat = torch.zeros(n, dtype=ps[0].dtype, device=ps[0].device)
        off = 0
        for p in ps:
            p.grad = flat[off:off + p.numel()].view_as(p)
            off += _al(p.numel())
        self.buckets.append((flat, ps))

    def _rebind(self):
        """p.grad must stay a view into its bucket: `optimizer.zero_grad(set_to_none=True)` or `p.grad = None`
        would silently detach it (the all-reduce would then average stale zeros).  Re-attach, keeping any
        gradient the detached tensor already holds."""
        for flat, ps in self.buckets:
            off = 0
            for p in ps:
                view = flat[off:off + p.numel()].view_as(p)
                if p.grad is None or p.grad.data_ptr() != view.data_ptr():
                    if p.grad is not None:
                        view.copy_(p.grad)
                    p.grad = view
                off += _al(p.numel())

    def reset(self):
        """Call once per step before backward: zero the flat gradient buffers (grads are views into them)."""
        for bi, (flat, ps) in enumerate(self.buckets):
            flat.zero_()
            self._pending[bi] = len(ps)
        self._rebind()
        self._handles = []
        self._next = 0  # buckets are launched strictly in index order on every rank

    def _launch(self, bi):
        flat = self.buckets[bi][0]
        if dist.get_backend(self.pg) == "nccl":
            h = dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=self.pg, async_op=True)
            self._handles.append((h, None))
        else:  # gloo (CPU tests): SUM then scale
            h = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.pg, async_op=True)
            self._handles.append((h, flat))

    def zero_buckets(self):
        """Zero the flat gradient buffers (capturable: plain memsets)."""
        for flat, _ in self.buckets:
            flat.zero_()

    def allreduce_all(self):
        """Average every bucket over the ranks, in bucket order, without the autograd hooks: for a backward pass that was
        replayed from a CUDA graph (collectives stay outside the graph; the gradients already sit in the flat buffers)."""
        if self.world <= 1:
            return
        self._handles = []
        for bi in range(len(self.buckets)):
            self._launch(bi)
        for h, flat in self._handles:
            h.wait()
            if flat is not None:
                flat.div_(self.world)
        self._handles = []

    def _hook(self, p):
        if not self.overlap:
            return
        bi = self._bucket_of[p]
        if p.grad is None or p.grad.data_ptr() < self.buckets[bi][0].data_ptr() or \
                p.grad.data_ptr() >= self.buckets[bi][0].data_ptr() + self.buckets[bi][0].numel() * p.element_size():
            raise RuntimeError("GradBucketReducer: a parameter's .grad no longer points into its bucket "
                               "(zero_grad(set_to_none=True)?); call reducer.reset() before every backward")
        self._pending[bi] -= 1
        # launch every complete bucket at the head of the queue: the collective order is the bucket order on every rank,
        # whatever order the hooks fire in (a bucket whose parameters got no gradient on this rank is launched in finish())
        while self.world > 1 and self._next < len(self.buckets) and self._pending[self._next] == 0:
            self._launch(self._next)
            self._next += 1

    def finish(self):
        """Call after backward, before the optimizer: launches what is left (same order on every rank) and waits."""
        if self.world > 1:
            while self._next < len(self.buckets):
                self._pending[self._next] = 0
                self._launch(self._next)
                self._next += 1
        for h, flat in self._handles:
            h.wait()
            if flat is not None:
                flat.div_(self.world)
        self._handles = []

    def grad_bytes(self) -> int:
        return sum(f.numel() * f.element_size() for f, _ in self.buckets)
