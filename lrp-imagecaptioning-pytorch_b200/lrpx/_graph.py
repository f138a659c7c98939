"""CUDA-graph capture that cannot be invalidated by Python's garbage collector.

torch captures in the global capture mode: while a stream is capturing, a CUDA call that is not a stream operation —
from anywhere in the process — invalidates the capture.  Destroying a ``torch.cuda.CUDAGraph`` is such a call
(cudaGraphExecDestroy), and graphs owned by objects in reference cycles (an explainer, its search plans and their
closures) are destroyed whenever the cyclic collector happens to run: a collection that starts inside somebody else's
capture fails it with cudaErrorStreamCaptureInvalidated ("operation not permitted when stream is capturing (function
reset)").  torch >= 2.9 no longer collects before a capture by default, so: collect first, keep the collector off until
the capture has ended."""
import contextlib
import gc

import torch


@contextlib.contextmanager
def capture(graph, **kw):
    """``with capture(g): ...`` == ``with torch.cuda.graph(g): ...`` with the collector flushed before and paused during."""
    gc.collect()
    was_enabled = gc.isenabled()
    gc.disable()
    try:
        with torch.cuda.graph(graph, **kw):
            yield
    finally:
        if was_enabled:
            gc.enable()
