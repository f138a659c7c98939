#!/usr/bin/env python
"""bench.py — LRP explanations/s (image heat-map + linguistic relevance) on B200.

Workload (BASELINE.json configs[1]): gridTDmodel (VGG16 encoder + adaptive-attention LSTM decoder), random-init
weights, B synthetic 224x224 images x T caption words, alpha=1/beta=0 image relevance + linguistic relevance for
every word.  One "step" = one pass of the hot path over the batch:
  encoder forward + per-layer gains (tcgen05)  ->  explainer teacher-forced forward (torch tensor ops)
  ->  decoder relevance kernels for all B*T requests  ->  encoder relevance chain (tcgen05) -> B*T heat-maps.

  python bench.py [--gpus N] [--steps K] [--warmup W]           one JSON line (rank 0)
  python bench.py --impl reference ...                           the reference's algorithm on the host CPU cores
                                                                 (oracle port; the Python reference cannot travel)
Under torchrun (N > 1) each rank explains its own B images (requests are independent: no collective on the
data path), timing is CUDA events, max over ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "lrp-imagecaptioning-pytorch_b200")
for p in (PKG, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

METRIC = "lrp_explanations_per_s"
UNIT = "explanations/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--images", type=int, default=64, help="images per GPU per step (configs[1]: 64)")
    ap.add_argument("--words", type=int, default=19, help="caption words per image (random-init captions run to max length)")
    ap.add_argument("--vocab", type=int, default=10000)
    ap.add_argument("--chunk", type=int, default=128, help="explanations per relevance-chain launch group")
    ap.add_argument("--no-graph", action="store_true", help="launch eagerly instead of replaying a CUDA graph of the step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-step", action="store_true",
                    help="run 1 warm-up + 1 step and exit (the command line captured under ncu for profiles/)")
    ap.add_argument("--cpu-words", type=int, default=19, help="words per image of the bounded CPU sample")
    ap.add_argument("--cpu-images", type=int, default=4, help="images of the bounded CPU sample (about 10-30 s of host work)")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(index), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); smax = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------
def build_problem(args, device, seed):
    """Random-init gridTD/VGG16 model + synthetic images and captions (seeded)."""
    import argparse as ap
    import synth
    from models import gridTDmodel as G
    V, H, E = args.vocab, 512, 512
    torch.manual_seed(seed)
    model = G.GridTDModel(E, H, V, "vgg16")
    model.load_state_dict(synth.gridtd_decoder_state(1000 + seed, V, H, E), strict=False)
    model.img_encoder.encoder.load_state_dict(synth.vgg_state(2000 + seed))
    if device is not None:
        model.to(device)
    model.eval()
    ns = ap.Namespace(embed_dim=E, hidden_dim=H, num_head=8, encoder="vgg16", height=224, width=224,
                      save_path="/tmp/lrpx_bench", dataset="syn", weight="")
    ex = G.ExplainGridTDAttention(ns, synth.word_map(V), model=model, precision="bf16")
    B, T = args.images, args.words
    g = torch.Generator().manual_seed(3000 + seed)
    imgs = torch.randn(B, 3, 224, 224, generator=g)
    toks = torch.randint(1, V - 4, (B, T + 1), generator=g)
    toks[:, 0] = V - 2
    return model, ex, imgs, toks


def run_ours(args):
    from lrpx import ops, _lib
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: lrpx has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa_node(local)      # before any pinned allocation: host buffers land next to the GPU
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False

    model, ex, imgs_h, toks_h = build_problem(args, dev, seed=rank)
    B, T = args.images, args.words
    Q = B * T
    eng = ex.engine()
    W = ex._lrp_weights()
    imgs_h = imgs_h.pin_memory()
    toks_h = toks_h.pin_memory()
    req_img = torch.arange(B, dtype=torch.int32).repeat_interleave(T).to(dev)
    req_t = torch.arange(T, dtype=torch.int32).repeat(B).to(dev)
    heat = torch.empty(Q, 3, 224, 224, device=dev, dtype=torch.float32)
    heat_h = torch.empty(Q, 3, 224, 224, dtype=torch.float32).pin_memory()
    words_h = torch.empty(Q, T, dtype=torch.float32).pin_memory()
    imgs_d, toks_d = imgs_h.to(dev), toks_h.to(dev)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    from lrpx.pipeline import BatchExplainer
    pipe = BatchExplainer(ex, chunk=args.chunk, use_graph=not args.no_graph)

    def step(imgs, toks):
        """The public batched call (lrpx.pipeline.BatchExplainer.explain)."""
        return pipe.explain(imgs, toks, out=heat)

    def step_e2e():
        """Same call with HOST buffers: pinned images/captions in, pinned heat-maps / word relevances out."""
        pipe.explain(imgs_h, toks_h, out=heat, host_out=(heat_h, words_h))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = ev(), ev()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    if args.profile_step:
        step(imgs_d, toks_d)
        torch.cuda.synchronize()
        step(imgs_d, toks_d)
        torch.cuda.synchronize()
        print(json.dumps({"profile_step": "done", "explanations": Q}))
        return

    def breakdown():
        """One instrumented step (outside the timed region): CUDA-event time of each phase."""
        marks = [ev() for _ in range(6)]
        calls0 = dict(_lib.CALLS)
        marks[0].record()
        est = eng.forward(imgs_d); marks[1].record()
        feat = eng.features(est, "pixel")
        st = ex.explainer_forward(feat, toks_d); marks[2].record()
        r_feat, r_words = ops.gridtd_decoder_lrp(st, W, req_img, req_t, toks_d[:, 1:].reshape(-1).to(torch.int32),
                                                 tc_gemm=True)
        marks[3].record()
        eng.relevance(est, r_feat, req_img, chunk=args.chunk, out=heat); marks[4].record()
        torch.cuda.synchronize()
        names = ["encoder_forward_gains", "explainer_forward", "decoder_relevance", "encoder_relevance_chain"]
        calls = {k: _lib.CALLS[k] - calls0.get(k, 0) for k in _lib.CALLS}
        return {n: round(marks[i].elapsed_time(marks[i + 1]), 3) for i, n in enumerate(names)}, calls

    # eager warm-up + the instrumented step BEFORE any graph exists (a captured graph pins a private memory pool,
    # which would distort eager timings taken after it)
    eager = BatchExplainer(ex, chunk=args.chunk, use_graph=False)
    for _ in range(2):
        eager.explain(imgs_d, toks_d, out=heat)
    phase_ms, calls = breakdown()
    for _ in range(2):                       # per-phase minimum of three instrumented steps
        again, _ = breakdown()
        phase_ms = {k: min(v, again[k]) for k, v in phase_ms.items()}
    del eager
    for _ in range(max(args.warmup, 3)):
        step(imgs_d, toks_d)
    sampler = ClockSampler(local) if rank == 0 else None
    ms = timed(lambda: step(imgs_d, toks_d), args.steps)
    clocks = sampler.stop() if sampler else None
    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    # the chain's time inside a step: measured in the instrumented (eager) step, per step
    chain_total = phase_ms["encoder_relevance_chain"] * args.steps
    step_eager_ms = sum(phase_ms.values())

    # launches of OUR kernels inside the timed region: one per C-ABI call, except the decoder call which
    # enqueues 4 weight splits + init + per step (3 element-wise kernels that also write the split GEMM operand + 2
    # tensor-core GEMMs) + 7 tail kernels
    # (glob, split + GEMM, avg, attention rows, projector GEMM with fused epilogue, word norm; csrc/decoder.cu,
    # LRPX_DEC_TC_GEMM path)
    launches = sum(v for k, v in calls.items() if k != "lrpx_gridtd_decoder_lrp_f32")
    launches += calls.get("lrpx_gridtd_decoder_lrp_f32", 0) * (5 + 5 * T + 7)
    launches *= args.steps          # the same kernels per step whether launched eagerly or replayed from the graph

    out = None
    if rank == 0:
        pk = peaks()
        total = world * Q * args.steps
        chain_flops = eng.flops_per_explanation() * Q * args.steps
        achieved = chain_flops / (chain_total * 1e-3) / 1e12 if chain_total > 0 else 0.0
        out = {
            "metric": METRIC, "value": total / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "gridTD VGG16 LRP alpha1beta0 image+linguistic explanations, "
                                   f"{B} images x {T} words per GPU per step, 224x224, V={args.vocab}, H=E=512",
                       "explanations_per_step_per_gpu": Q, "chunk": args.chunk,
                       "host_affinity": (f"{len(numa)} cores nearest to the GPU (NVML)" if isinstance(numa, list) else numa), "cuda_graph": not args.no_graph, "parallelism": f"request-sharded x{world}",
                       "l2": "working set (gains 1.9 GB + chain buffers) far larger than the 126 MB L2; no flush needed",
                       "decoder_relevance_dtype": "f32 element-wise, GEMMs as error-compensated bf16x3 on tensor cores (f32 accumulate)", "encoder_relevance_dtype": "bf16 operands, f32 accumulate"},
            "e2e": {"value": total / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": imgs_h.numel() * 4 + toks_h.numel() * 8,
                    "d2h_bytes_per_step": heat_h.numel() * 4 + words_h.numel() * 4},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": "tc_conv_slab_kernel<MUL|MUL_UNPOOL|INPUT3> (encoder relevance chain)",
                         "achieved": achieved, "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                         "frac": achieved / pk["tf_sustained"], "traffic": chain_traffic(args.chunk),
                         "traffic_note": "dram read+write bytes of the 13 chain layers per chunk of explanations (ncu --set "
                                         "full of the 8 low-resolution launches over 1216 requests + the 5 launches "
                                         "of one 128-request chunk, profiles/r1_chain_full.md); algorithmic FLOPs per "
                                         "chunk = chunk x algorithmic_gflop_per_explanation",
                         "peak_source": pk["src"] + " sustained bf16",
                         "share_of_step": phase_ms["encoder_relevance_chain"] / step_eager_ms,
                         "algorithmic_gflop_per_explanation": eng.flops_per_explanation() / 1e9,
                         # SURVEY 8(d) counts the reference's formulation (z+ recomputed per word: two contractions per
                         # layer = 2 x the figure above); `achieved` counts only the FLOPs this chain executes
                         "reference_formulation_gflop_per_explanation": 2.0 * eng.flops_per_explanation() / 1e9},
        }
        out["breakdown_ms"] = phase_ms
        try:
            out["roofline"]["layers"] = layer_table(eng, args.chunk, dev, pk, requests=args.images * args.words)
        except Exception as e:  # never lose the bench line to the per-layer microbenchmark
            out["roofline"]["layers_error"] = repr(e)
        if not args.no_cpu_baseline and world == 1:
            out["cpu_baseline"] = cpu_baseline(args, steps=1)
        print(json.dumps(out), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def bind_to_gpu_numa_node(index):
    """Pins this rank's process to the CPU cores nearest to its GPU (NVML's ideal affinity) so that the pinned
    host buffers of the end-to-end path are first-touched on the GPU's own NUMA node.  With 8 ranks each moving
    ~0.7 GB of heat-maps per step, buffers on the far socket put the whole result stream on the socket link."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return sorted(os.sched_getaffinity(0))
    except Exception as e:      # affinity is an optimisation only
        return f"unavailable: {e!r}"


def chain_traffic(chunk):
    """dram__bytes_read.sum + dram__bytes_write.sum of the relevance-chain launches, per `chunk` explanations, from the
    committed ncu --set full capture of all chain launches of one step (profiles/r1_chain_full.json, written by
    scripts/ncu_summary.py); None if absent."""
    path = os.path.join(ROOT, "profiles", "r1_chain_full.json")
    if not os.path.exists(path):
        return None
    d = json.load(open(path))
    if not d.get("dram_bytes_per_explanation"):
        return None
    return d["dram_bytes_per_explanation"] * chunk


def layer_table(eng, chunk, dev, pk, requests=None):
    """Per-layer achieved TFLOP/s of the relevance-chain launches (CUDA events, 5 launches each) at the launch sizes the
    engine uses: the low-resolution layers over all `requests` of a step in one launch (TcVggEngine.relevance_head),
    the others per chunk; `ms` is per chunk of explanations either way."""
    from lrpx import tc
    rows = []
    x = torch.randn(1, 3, eng.convs[0].h, eng.convs[0].w, device=dev)
    st = eng.forward(x[:1])
    L = len(eng.convs)
    Q = requests or chunk
    n_wide = 0                                  # the engine's own rule (relevance_head): 2 GB per stage-1 buffer
    if Q > chunk:
        for li in range(L - 1, 0, -1):
            c, below = eng.convs[li], eng.convs[li - 1]
            oh, ow = (2 * c.h, 2 * c.w) if below.pool_after else (c.h, c.w)
            if max(tc.pf_rows(Q, c.h, c.w) * c.cout, tc.pf_rows(Q, oh, ow) * c.cin) * 2 > (2 << 30):
                break
            n_wide += 1
    for li in range(L - 1, -1, -1):
        c = eng.convs[li]
        n = Q if li >= L - n_wide else chunk
        rimg = torch.zeros(n, dtype=torch.int32, device=dev)
        a = torch.randn(tc.pf_rows(n, c.h, c.w), c.cout, device=dev).to(torch.bfloat16)
        if li == 0:
            out = torch.empty(n, 3, c.h, c.w, device=dev)
            if c.w_rel3 is not None:          # the engine's own choice: filter columns folded into N
                fn = lambda: tc.tc_conv(a, c.w_rel3, n, c.h, c.w, c.cout, 24, 3, tc.EPI_INPUT3, out, row_img=rimg, x=st.x)
            else:
                fn = lambda: tc.tc_conv(a, c.w_rel, n, c.h, c.w, c.cout, 16, 3, tc.EPI_INPUT, out, row_img=rimg, x=st.x)
            flops = 2.0 * n * c.h * c.w * c.cin * c.cout * 9
        else:
            below = eng.convs[li - 1]
            if below.pool_after:
                out = torch.empty(tc.pf_rows(n, 2 * c.h, 2 * c.w), c.cin, device=dev, dtype=torch.bfloat16)
                fn = lambda: tc.tc_conv(a, c.w_rel, n, c.h, c.w, c.cout, c.cin, 3, tc.EPI_MUL_UNPOOL, out,
                                        gain=st.gain[li - 1], row_img=rimg, pool_idx=st.idx[li - 1])
            else:
                out = torch.empty(tc.pf_rows(n, c.h, c.w), c.cin, device=dev, dtype=torch.bfloat16)
                fn = lambda: tc.tc_conv(a, c.w_rel, n, c.h, c.w, c.cout, c.cin, 3, tc.EPI_MUL, out,
                                        gain=st.gain[li - 1], row_img=rimg)
            flops = 2.0 * n * c.h * c.w * c.cin * c.cout * 9
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        rows.append({"layer": li, "hw": c.h, "k": 9 * c.cout, "n": c.cin if li else 6, "launch_requests": n,
                     "ms": round(ms * chunk / n, 4), "tflops": round(flops / (ms * 1e-3) / 1e12, 1)})
        del a, out
    return rows


# ---------------------------------------------------------------------------------------------------
def cpu_baseline(args, steps):
    """The reference's algorithm on the host cores (oracle port: torch-CPU restatement of the reference's rules,
    pinned against the reference's own outputs in tests/golden).  Bounded sample: 1 image, `cpu_words` words."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import lrp_oracle as O
    import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    V, H, E = args.vocab, 512, 512
    p = synth.gridtd_decoder_state(1000, V, H, E)
    vsd = synth.vgg_state(2000)
    layers = O.vgg_layers_from_state(vsd)
    nimg = max(1, args.cpu_images)
    imgs = synth.images(3000, nimg)
    toks = synth.tokens(3001, args.words, V)
    nwords = max(1, min(args.cpu_words, args.words))
    best = None
    for _ in range(max(1, steps)):
        t0 = time.perf_counter()
        for b in range(nimg):
            img = imgs[b:b + 1]
            feats = O.sequential_forward(layers, img)[-1]
            st = O.gridtd_explainer_forward(p, feats[0], toks)
            for t in range(args.words - nwords, args.words):
                rf, rw, _ = O.gridtd_explain_wordt(p, st, t)
                O.sequential_lrp(layers, img, rf.t().reshape(1, 512, 14, 14))
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return {"value": nimg * nwords / best, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{nimg} images (forward once each) x {nwords} words (decoder + VGG16 encoder relevance, the "
                      f"reference's per-word formulation), fp32 torch-CPU, {best:.2f} s"}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    # keep the whole run within a few minutes whatever --steps / --warmup are: time one image, then size the
    # per-step sample (whole images of `cpu_words` words) for a ~180 s total
    probe = argparse.Namespace(**vars(args))
    probe.cpu_images = 1
    t0 = time.perf_counter()
    cpu_baseline(probe, steps=1)
    per_image = time.perf_counter() - t0
    n_steps = max(1, args.warmup + args.steps)
    args.cpu_images = max(1, min(args.cpu_images, int(180.0 / n_steps / max(per_image, 1e-3))))
    t_all = []
    for i in range(args.warmup + args.steps):
        r = cpu_baseline(args, steps=1)
        if i >= args.warmup:
            t_all.append(r)
    v = sum(x["value"] for x in t_all) / len(t_all)
    nwords = max(1, min(args.cpu_words, args.words)) * max(1, args.cpu_images)
    out = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": 1e3 * nwords / v, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": "gridTD VGG16 LRP alpha1beta0 image+linguistic explanations, 224x224, "
                                  f"V={args.vocab}, H=E=512; bounded sample per step: {max(1, args.cpu_images)} images x "
                                  f"{max(1, min(args.cpu_words, args.words))} words"},
           "cpu_baseline": {"value": v, "unit": UNIT, "cores": t_all[0]["cores"], "kind": "port",
                            "sample": t_all[0]["sample"]},
           "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
