"""GPU probe: lrpx_lstm_step_f32 vs (library addmm + lrpx_lstm_cell_f32) for the explainer's two LSTM shapes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lrp-imagecaptioning-pytorch_b200"))
import torch
from lrpx import ops
torch.backends.cuda.matmul.allow_tf32 = False
dev = "cuda"
B, H = 64, 512
def timeit(fn, n=20):
    """time per call when n calls are replayed from one CUDA graph (no host launch overhead)"""
    fn(); torch.cuda.synchronize()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (5 * n) * 1e3
for K, G in ((1024, 5), (1536, 4)):
    x = torch.randn(B, K, device=dev); W = torch.randn(K, G * H, device=dev) * 0.05; add = torch.randn(B, G * H, device=dev)
    c_prev = torch.randn(B, H, device=dev)
    outs = [torch.empty(B, H, device=dev) for _ in range(6)]
    wp = ops.lstm_prep_weights(W, G)
    def fused():
        ops.lstm_step(x, wp, add, G, c_prev, outs[0], outs[1], outs[2], outs[3], outs[4], s=outs[5] if G == 5 else None)
    ref = [torch.empty(B, H, device=dev) for _ in range(6)]
    def lib():
        z = torch.addmm(add, x, W)
        ops.lstm_cell(z, c_prev, ref[0], ref[1], ref[2], ref[3], ref[4], gate_pre=z[:, 4 * H:] if G == 5 else None, s=ref[5] if G == 5 else None)
    fused(); lib(); torch.cuda.synchronize()
    err = max(float((a - b).abs().max()) for a, b in zip(outs[:5], ref[:5]))
    print(f"K={K} G={G}: fused {timeit(fused):.1f} us, addmm+cell {timeit(lib):.1f} us, addmm alone {timeit(lambda: torch.addmm(add, x, W)):.1f} us, max diff {err:.2e}")
