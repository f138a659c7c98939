"""GPU probe: how does tcgen05 round when it adds a K=16 product block into its fp32 accumulator?  bf16 operands in [1,2)
(exact 16-bit products), K up to 4608: signed error of the tensor-core sum vs fp64 in units of ulp(result)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("lrp-imagecaptioning-pytorch_b200", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import torch
from lrpx import tc, _lib
import ctypes as C
dev = "cuda"
g = torch.Generator().manual_seed(0)
for K in (64, 576, 4608, 13824):
    m, n = 1024, 64
    a = (1 + torch.rand(m, K, generator=g)).to(torch.bfloat16).to(dev)
    w = (1 + torch.rand(n, K, generator=g)).to(torch.bfloat16).to(dev)
    out = torch.empty(m, n, device=dev)
    _lib.check(_lib.lib().lrpx_tc_gemm_bf16_f32(a.data_ptr(), w.data_ptr(), out.data_ptr(), m, n, K, torch.cuda.current_stream().cuda_stream), "gemm")
    ref = a.double() @ w.double().t()
    ulp = torch.pow(2.0, torch.floor(torch.log2(ref)) - 23)
    e = (out.double() - ref) / ulp
    t = torch.matmul(a.float(), w.float().t())       # fp32 CUDA-core / cublas result for comparison
    e2 = (t.double() - ref) / ulp
    print(f"K={K}: MMA steps {K // 16}: tcgen05 error mean {float(e.mean()):+.2f} ulp, std {float(e.std()):.2f}, "
          f"rel mean {float(((out.double() - ref) / ref).mean()):+.3e}; torch fp32 matmul mean {float(e2.mean()):+.2f} std {float(e2.std()):.2f}")
