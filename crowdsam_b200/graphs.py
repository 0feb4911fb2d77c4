"""CUDA-graph capture of the two launch-bound regions of the hot path.

`set_image` (~380 kernel launches: both ViT encoders + the prompt-independent decoder work) and `decode`
(~110 launches per prompt batch, many of them a few microseconds long on the 7-token side of the two-way
transformer) have static shapes for a given image size / prompt count, no host synchronisation and no
data-dependent control flow.  Launched one by one from Python through ctypes the GPU idles ~6 % of a step
between kernels (round 1: sum of kernel times 47.3 ms of a 50.6 ms step); replayed as one graph the launch
gaps disappear and the host is free for the previous image's RLE encoding.

A region is captured the SECOND time its key (shapes) is seen: the first call runs eagerly (it also fills the
host-side caches that must not be touched during capture: tensor-map cache, gather maps, scratch buffers),
one-off shapes (a ragged last EPS batch) never pay for a capture.  Captured regions live in a small LRU; every
graph owns the private memory pool of its intermediates, so an evicted graph returns its memory.

Measured (round 2, B200, ViT-L + DINOv2-L, 1024 prompts): replaying both regions leaves the step time unchanged
(49.9 ms against 50.1 ms eager): launched from Python the host already runs several milliseconds ahead of the GPU,
and the gaps between dependent kernels are the same ~2 us inside a graph.  Cluster-launched kernels (the CTA-pair
GEMM) replayed from a graph were 30 % SLOWER than launched eagerly (65 ms step).  The mechanism therefore stays
opt-in (`CSAM_GRAPHS=1`): it frees ~7 ms of host time per image for callers that overlap their own CPU work.

Replays do not pass through the C-ABI launch counter, so the number of kernels a replay launches is recorded at
capture time and accumulated here (`replayed_launches`, reported by bench.py inside `gpu_launches`).
"""
from __future__ import annotations

import os
from collections import OrderedDict
from typing import Callable, Sequence

import torch

from . import lib as L

ENABLED = os.environ.get("CSAM_GRAPHS", "0") == "1"
replayed_launches = 0          # kernels launched through graph replays (the C-ABI counter only sees captures)
captures = 0


def usable() -> bool:
    """Graphs are used on the plain hot path only: not under the per-launch event profiler (events cannot be
    timed inside a capture) and not while another capture is running."""
    from . import ops

    return ENABLED and ops.PROFILER is None and not torch.cuda.is_current_stream_capturing()


class Graphed:
    """fn(*static_inputs) captured once; __call__ copies the inputs into the static buffers and replays.
    The returned object is whatever fn returned during capture (tensors in the graph's private pool: they are
    overwritten by the next replay, callers clone what must outlive it)."""

    def __init__(self, fn: Callable, inputs: Sequence[torch.Tensor]):
        global captures
        self.static_in = [torch.empty_like(t) for t in inputs]
        for s, t in zip(self.static_in, inputs):
            s.copy_(t)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        l0 = L.launch_count()
        with torch.cuda.graph(self.graph):
            self.out = fn(*self.static_in)
        self.n_launches = L.launch_count() - l0
        captures += 1

    def __call__(self, *inputs: torch.Tensor):
        global replayed_launches
        for s, t in zip(self.static_in, inputs):
            s.copy_(t, non_blocking=True)
        self.graph.replay()
        replayed_launches += self.n_launches
        return self.out


class GraphCache:
    """key -> Graphed with capture-on-second-sight and LRU eviction."""

    def __init__(self, capacity: int):
        self.capacity = capacity
        self.graphs: "OrderedDict[object, Graphed]" = OrderedDict()
        self.seen = set()

    def get(self, key):
        g = self.graphs.get(key)
        if g is not None:
            self.graphs.move_to_end(key)
        return g

    def second_sight(self, key) -> bool:
        """True when `key` was seen before (now worth capturing)."""
        if key in self.seen:
            return True
        if len(self.seen) > 4096:
            self.seen.clear()
        self.seen.add(key)
        return False

    def put(self, key, g: Graphed) -> Graphed:
        self.graphs[key] = g
        while len(self.graphs) > self.capacity:
            self.graphs.popitem(last=False)         # drops the graph and its memory pool
        return g

    def clear(self):
        self.graphs.clear()
        self.seen.clear()
