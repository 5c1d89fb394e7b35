"""CUDA-graph replay of one training step (forward + backward + whatever the caller does with the gradients).

The step is ~1000 kernel launches of a few tens of microseconds each; replayed as one graph the launch gaps between
dependent kernels (~2 us each from the CPU-driven stream) disappear.  Shapes are static (fixed batch, ctx, image
count), inputs live in static device buffers that `__call__` refills (H2D from pinned host tensors or D2D), gradients and
activations come out of the graph's private memory pool at the same addresses on every replay.  The native kernels take
part unchanged: their stream-ordered workspaces (`cudaMallocAsync` / `cudaFreeAsync`) become allocation nodes, TMA
descriptors are passed by value and stay valid because every buffer keeps its address.

Not capturable, and therefore skipped while capturing (ops.embed_and_scatter): the deferred host-side check of the
image-token count — call `ops.flush_checks()` on an eager step if the data can be malformed."""
from __future__ import annotations

import torch


class GraphedTrainStep:
    """step = GraphedTrainStep(model, example_batch, after_backward=fn); loss = step(batch)

    `after_backward()` runs inside the captured region after `loss.backward()` (gradient all-reduce, master-weight copy,
    optimizer step with a capturable optimizer, ...).  `before_forward()` likewise runs first (e.g. a bucket reset)."""

    def __init__(self, model, example_batch, after_backward=None, before_forward=None, warmup: int = 3, set_grads_to_none: bool = True):
        self.model = model
        self.static = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in example_batch.items()}
        self._params = [p for p in model.parameters() if p.requires_grad]

        def run():
            if before_forward is not None:
                before_forward()
            if set_grads_to_none:
                for p in self._params:
                    p.grad = None
            loss = model.training_step(self.static)
            loss.backward()
            if after_backward is not None:
                after_backward()
            return loss

        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):   # sizes every lazily allocated workspace, creates communicators, warms the caches
                run()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        from . import ops
        ops.flush_checks()   # host-side checks left over from the eager warm-up steps (they cannot run while capturing)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = run()

    def __call__(self, batch):
        for k, v in batch.items():
            if torch.is_tensor(v):
                self.static[k].copy_(v, non_blocking=True)
        self.graph.replay()
        return self.loss
