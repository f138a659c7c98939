"""GPU: lrpx._graph.capture — a CUDA-graph capture that Python's cyclic garbage collector cannot invalidate.

Regression for a failure seen on a B200 box (test_gpu_beam, "operation not permitted when stream is capturing (function
reset)"): graphs owned by reference cycles of EARLIER explainers were destroyed by a collection that happened to start
inside a later capture; destroying a graph is not a stream operation, so the capture was invalidated."""
import gc

import pytest
import torch

pytestmark = pytest.mark.gpu


class _Owner:
    """A reference cycle that owns a captured graph: only the cyclic collector can free it."""

    def __init__(self):
        self.me = self
        self.x = torch.zeros(1024, device="cuda")
        self.graph = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            self.x.add_(1.0)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        with torch.cuda.graph(self.graph):
            self.x.add_(1.0)


def test_capture_flushes_and_pauses_the_collector():
    from lrpx import _graph
    gc.collect()
    was = gc.isenabled()
    gc.disable()                      # keep the cyclic garbage below alive until the helper collects it
    try:
        for _ in range(3):
            _Owner()                  # garbage at once, but in a cycle: its CUDAGraph is still alive
        y = torch.zeros(256, device="cuda")
        g = torch.cuda.CUDAGraph()
        gc.enable()
        with _graph.capture(g):
            assert not gc.isenabled()
            # allocation pressure that would trigger a generation-0 collection if the collector were running
            junk = [[i] for i in range(20000)]
            del junk
            y.add_(2.0)
        assert gc.isenabled()
        g.replay()
        g.replay()
        torch.cuda.synchronize()
        assert float(y[0]) == 4.0
    finally:
        if was:
            gc.enable()
        else:
            gc.disable()
