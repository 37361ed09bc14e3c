"""Whole-step CUDA-graph capture for training loops that can afford static shapes.

The reference's `train.py` launches ~1 500 kernels per step from Python; on 64x64 patches the B200 finishes them faster than
the host can issue them (cfg4: 48 ms eager vs 43 ms of GPU work).  Every launch of libvsrb200.so is stream-ordered and
allocation-free, so forward + backward + gradient clipping + optimizer step capture into ONE graph with torch's standard
whole-network recipe; `GraphedTrainStep` packages that recipe.  Single GPU (or one graph per DDP rank with
`torch.distributed` collectives outside the graph is NOT handled here).

    step = GraphedTrainStep(model, optimizer, loss_fn, example_inputs=(lr, hr), clip_grad_norm=1.0)
    for lr, hr in loader:
        loss = step(lr, hr)          # copies the batch into the graph's static buffers and replays

`optimizer` must have been built with `capturable=True` (torch.optim.Adam / AdamW).  `loss_fn(outputs, *targets)` gets the
model's outputs and the remaining inputs.
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence

import torch


class GraphedTrainStep:
    def __init__(self, model: torch.nn.Module, optimizer: torch.optim.Optimizer, loss_fn: Callable,
                 example_inputs: Sequence[torch.Tensor], clip_grad_norm: Optional[float] = None,
                 autocast_dtype: Optional[torch.dtype] = torch.bfloat16, warmup: int = 3):
        if not all(g.get("capturable", False) for g in optimizer.param_groups):
            raise ValueError("GraphedTrainStep needs an optimizer created with capturable=True")
        self.model, self.opt, self.loss_fn, self.clip, self.dtype = model, optimizer, loss_fn, clip_grad_norm, autocast_dtype
        self.static_in = [t.clone() for t in example_inputs]
        self.stream = torch.cuda.Stream(device=self.static_in[0].device)
        self.stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.stream):              # warm-up on the capture stream (also where the grad accumulators live)
            for _ in range(warmup):
                self.opt.zero_grad(set_to_none=True)
                self._body()
        torch.cuda.current_stream().wait_stream(self.stream)
        torch.cuda.synchronize()
        self.opt.zero_grad(set_to_none=True)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=self.stream):
            self.static_loss = self._body()

    def _body(self) -> torch.Tensor:
        x = self.static_in[0].clone()                     # the model refines its input in place (reference contract)
        if self.dtype is not None:
            with torch.autocast("cuda", dtype=self.dtype):
                out = self.model(x)
        else:
            out = self.model(x)
        loss = self.loss_fn(out, *self.static_in[1:])
        loss.backward()
        if self.clip is not None:
            torch.nn.utils.clip_grad_norm_(self.model.parameters(), self.clip)
        self.opt.step()
        return loss.detach()

    def __call__(self, *inputs: torch.Tensor) -> torch.Tensor:
        for dst, src in zip(self.static_in, inputs):
            dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.static_loss
