"""Whole training step (forward + loss + backward of the 10-unit model) replayed as ONE CUDA graph.

The reference launches every ATen kernel of a step from Python (torch_src/session/procedures/step.py:39-43).  The C ABI
of this package never allocates, never synchronises and takes the stream from the caller, so a step is capturable: the
~630 launches of the NTU model become one ``cudaGraphLaunch`` and the host leaves the critical path (the eager step is
launch-bound below ~8 sequences per GPU, profiles/r1q).  Inputs, logits, loss and the parameter gradients live at fixed
addresses in the graph's memory pool; ``__call__`` copies a new batch into the static input and replays.
"""
from typing import Callable, Optional

import torch

from . import capi


def loss_and_logits(model: torch.nn.Module, loss_fn: Callable, x: torch.Tensor, y: torch.Tensor):
    """loss_fn(model(x), y) -- through the fused classifier + cross-entropy head (``Model.loss``) when ``loss_fn`` is a plain
    ``nn.CrossEntropyLoss()`` (mean reduction, no class weights, no label smoothing: what the reference builds at
    torch_src/session/session.py:53) and the model has one; any other loss runs as given."""
    fused = (isinstance(loss_fn, torch.nn.CrossEntropyLoss) and loss_fn.weight is None and loss_fn.reduction == "mean"
             and loss_fn.label_smoothing == 0.0 and getattr(model, "fc", None) is not None and hasattr(model, "loss")
             and y.dtype == torch.int64 and y.dim() == 1)
    if fused:
        return model.loss(x, y)
    logits = model(x)
    return loss_fn(logits, y), logits


class GraphedStep:
    """``step = GraphedStep(model, loss_fn, x_example, y_example)``; ``loss = step(x, y)`` runs zero-grad + forward + loss +
    backward and leaves the gradients in ``p.grad`` (static tensors, overwritten by every replay).  ``after_backward`` (for
    instance a gradient all-reduce) is captured into the same graph when given."""

    def __init__(self, model: torch.nn.Module, loss_fn: Callable, x_example: torch.Tensor, y_example: torch.Tensor,
                 warmup: int = 3, after_backward: Optional[Callable[[], None]] = None, preserve_buffers: bool = False):
        """``preserve_buffers``: the warm-up steps run real forward passes (BatchNorm running statistics move); with this flag the
        module buffers are put back afterwards, so that the first replay is the first step the model sees (training sessions that
        build the graph lazily on their first batch)."""
        if not x_example.is_cuda:
            raise RuntimeError("GraphedStep needs CUDA tensors (there is no CPU path)")
        self.model, self.loss_fn = model, loss_fn
        self.x = x_example.detach().clone()
        self.y = y_example.detach().clone()
        saved_buffers = [b.detach().clone() for b in model.buffers()] if preserve_buffers else None
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):          # warm-up off the default stream: lazy initialisation (function attributes,
            for _ in range(max(1, warmup)):    # driver entry points, allocator) must not happen during capture
                model.zero_grad(set_to_none=True)
                loss_and_logits(model, loss_fn, self.x, self.y)[0].backward()
                if after_backward is not None:
                    after_backward()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        model.zero_grad(set_to_none=True)      # gradients are (re)allocated inside the capture, from the graph's pool
        self.graph = torch.cuda.CUDAGraph()
        before = capi.lib().agcn_launch_count()
        with torch.cuda.graph(self.graph):
            self.loss, self.logits = loss_and_logits(model, loss_fn, self.x, self.y)
            self.loss.backward()
            if after_backward is not None:
                after_backward()
        self.launches_per_replay = int(capi.lib().agcn_launch_count() - before)   # kernels of libagcn_b200.so in the graph
        # the static gradient tensors: a caller whose loop sets p.grad to None (optimizer.zero_grad()) re-attaches them after a replay
        self.grads = [(p, p.grad) for p in model.parameters() if p.grad is not None]
        if saved_buffers is not None:
            with torch.no_grad():
                for b, old in zip(model.buffers(), saved_buffers):
                    b.copy_(old)

    def attach_grads(self) -> None:
        """p.grad <- the graph's static gradient tensor, for every parameter the captured backward reaches."""
        for p, g in self.grads:
            p.grad = g

    def __call__(self, x: Optional[torch.Tensor] = None, y: Optional[torch.Tensor] = None) -> torch.Tensor:
        if x is not None:
            self.x.copy_(x, non_blocking=True)
        if y is not None:
            self.y.copy_(y, non_blocking=True)
        self.graph.replay()
        return self.loss
