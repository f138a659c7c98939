"""Builds liblrpx.so in-tree with nvcc for sm_100a (no torch headers, plain C ABI)."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(os.path.dirname(HERE), "csrc")
LIB = os.path.join(HERE, "liblrpx.so")
SOURCES = ["elementwise.cu", "conv_simt.cu", "decoder.cu", "decoder_grad.cu", "conv_tc.cu", "tc_aux.cu", "explainer_fwd.cu", "beam.cu", "ablation.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "--use_fast_math=false"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(os.path.dirname(os.path.dirname(HERE)), "include", "lrpx.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    nvcc = _nvcc()
    objs = []
    flags = [f for f in NVCC_FLAGS if not f.startswith("--use_fast_math")]
    os.makedirs(os.path.join(HERE, "_obj"), exist_ok=True)
    procs = []
    header = os.path.join(os.path.dirname(os.path.dirname(HERE)), "include", "lrpx.h")
    shared_deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))] + [header]
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        if not os.path.exists(src):
            continue
        obj = os.path.join(HERE, "_obj", s.replace(".cu", ".o"))
        objs.append(obj)
        # every object depends on its source, the shared csrc headers AND the public ABI header (struct layouts)
        if (not force) and os.path.exists(obj) and os.path.getmtime(obj) > max(
                os.path.getmtime(d) for d in shared_deps + [src]):
            continue
        cmd = [nvcc] + flags + ["-c", src, "-o", obj]
        if verbose:
            print(" ".join(cmd))
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for s, p in procs:
        out = p.communicate()[0].decode()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed for {s}:\n{out}")
        if verbose and out.strip():
            print(out)
    tmp = LIB + f".tmp{os.getpid()}"
    cmd = [nvcc, "-shared", "-o", tmp] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    if r.returncode != 0:
        if os.path.exists(tmp):
            os.remove(tmp)
        raise RuntimeError("link failed:\n" + r.stdout.decode())
    os.replace(tmp, LIB)          # a reader never sees a half-written library
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
