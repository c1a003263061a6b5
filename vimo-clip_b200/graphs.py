"""CUDA-graph replay of an inference forward for the launch-bound regime.

The reference's ``inference.py:129`` / ``inference_frame_diff.py`` push ONE clip (15 frames) at a time through the student, and
``TFAM`` evaluation runs small batches: a ViT-B/32 forward is ~110 kernel launches of a few microseconds each, so the step
is bound by launch latency, not by any GPU pipe.  ``graphed(fn, *example_inputs)`` captures one call of ``fn`` (any of the
drop-in modules in ``.eval()`` mode, or a function chaining them) into a ``torch.cuda.CUDAGraph`` and replays it: one
``cudaGraphLaunch`` per forward instead of one launch per kernel.  Everything our kernels need is capture-safe: they are
enqueued on the stream torch captures, weights are packed (and the tower workspace allocated) during the warm-up calls,
activations come from torch's graph-private memory pool, and TMA descriptors bake in pointers that stay fixed for the
life of the graph.

    student = FlowStudentModel("ViT-B/32").eval()
    g = graphed(student, clip_u8)            # clip_u8: [1, 15, 3, 224, 224] uint8 on the GPU
    emb, emb_distill, logits = g(next_clip)  # same shapes / dtypes as the example; outputs are overwritten by the next call

Only tensors of the example's shape, dtype and device may be passed afterwards; parameters must not be modified (re-capture
after ``load_state_dict``).  Non-tensor arguments (``None`` masks, flags) are baked into the graph.
"""
from __future__ import annotations

import torch

from . import _lib


class GraphedForward:
    def __init__(self, fn, *example_inputs, warmup: int = 2):
        tensors = [x for x in example_inputs if isinstance(x, torch.Tensor)]
        if not tensors or any(not x.is_cuda for x in tensors):
            raise _lib.VmcError("graphed() needs CUDA example inputs (no CPU fallback)")
        self._fn = fn
        self._args = [x.clone() if isinstance(x, torch.Tensor) else x for x in example_inputs]
        self._slots = [i for i, x in enumerate(self._args) if isinstance(x, torch.Tensor)]
        dev = tensors[0].device
        with torch.cuda.device(dev), torch.no_grad():
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):  # weight packing, workspace allocation, lazy library state
                for _ in range(max(1, warmup)):
                    fn(*self._args)
            torch.cuda.current_stream().wait_stream(side)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self._out = fn(*self._args)

    def __call__(self, *inputs):
        if len(inputs) != len(self._args):
            raise ValueError(f"expected {len(self._args)} arguments like the example call")
        for i in self._slots:
            src, dst = inputs[i], self._args[i]
            if not isinstance(src, torch.Tensor) or src.shape != dst.shape or src.dtype != dst.dtype:
                raise ValueError(f"argument {i}: expected a {tuple(dst.shape)} {dst.dtype} tensor like the example")
            dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self._out


def graphed(fn, *example_inputs, warmup: int = 2) -> GraphedForward:
    """Capture ``fn(*example_inputs)`` and return a callable that replays it (see the module docstring)."""
    return GraphedForward(fn, *example_inputs, warmup=warmup)
