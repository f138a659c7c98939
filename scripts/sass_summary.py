#!/usr/bin/env python
"""profiles/rNN_sass_summary.md: per kernel of liblrpx.so the SASS evidence for the tcgen05 / TMEM / TMA path and the
resource usage ptxas reports — counts of UTCHMMA (tcgen05.mma), LDTM (tcgen05.ld), UTMALDG / UTMASTG (TMA tile
loads / stores, .MULTICAST variants), UTCBAR (tcgen05.commit -> mbarrier), SYNCS (mbarrier ops), registers, stack
bytes (spills) and static shared memory.  Runs in the build container (cuobjdump needs no GPU).
  python scripts/sass_summary.py profiles/r2_sass_summary.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "lrp-imagecaptioning-pytorch_b200", "lrpx", "liblrpx.so")
OPS = ["UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "HMMA", "FFMA"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main(dst):
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    counts, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m:
            op = m.group(1)
            base = op.split(".")[0]
            if base in OPS:
                counts[cur][base] += 1
                if ".MULTICAST" in op:
                    counts[cur][base + ".MULTICAST"] += 1
    usage = {}
    fn = None
    for line in res.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            fn = m.group(1)
            continue
        if fn and "REG:" in line:
            usage[fn] = dict(re.findall(r"(REG|STACK|SHARED|LOCAL):(\d+)", line))
            fn = None
    names = demangle(list(counts))
    arch = re.findall(r"arch = (sm_\w+)", sass)
    with open(dst, "w") as f:
        f.write("# SASS / resource summary of liblrpx.so\n\n"
                f"`cuobjdump -sass` / `-res-usage` of `lrpx/liblrpx.so` built by `lrpx/build.py` (nvcc -gencode "
                f"arch=compute_100a,code=sm_100a -O3 -lineinfo); architectures in the fatbin: {sorted(set(arch))}.\n"
                "UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld (TMEM -> registers), UTMALDG / UTMASTG = cp.async.bulk.tensor "
                "(TMA) loads / stores, UTCBAR = tcgen05.commit, SYNCS = mbarrier operations; HMMA (mma.sync) must be 0.\n\n"
                "| kernel | UTCHMMA | LDTM | UTMALDG (multicast) | UTMASTG | UTCBAR | SYNCS | HMMA | FFMA | regs | stack B | static smem B |\n"
                "|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|\n")
        tot = collections.Counter()
        for k, c in counts.items():
            u = usage.get(k, {})
            n = re.sub(r"\(.*", "", names.get(k, k)).replace("void lrpx::", "").replace("lrpx::", "")
            f.write(f"| `{n[:70]}` | {c['UTCHMMA']} | {c['LDTM']} | {c['UTMALDG']} ({c['UTMALDG.MULTICAST']}) | {c['UTMASTG']} | "
                    f"{c['UTCBAR']} | {c['SYNCS']} | {c['HMMA']} | {c['FFMA']} | {u.get('REG', '?')} | {u.get('STACK', '?')} | "
                    f"{u.get('SHARED', '?')} |\n")
            tot.update(c)
        f.write(f"\ntotals: UTCHMMA {tot['UTCHMMA']}, LDTM {tot['LDTM']}, UTMALDG {tot['UTMALDG']} "
                f"({tot['UTMALDG.MULTICAST']} multicast), UTMASTG {tot['UTMASTG']}, UTCBAR {tot['UTCBAR']}, HMMA {tot['HMMA']} "
                f"over {len(counts)} kernels.\n\nEpilogue codes of `tc_conv_kernel<E>` / `tc_conv_slab_kernel<E>` (include/lrpx.h): "
                "1 FWD_GAIN, 2 MUL, 3 MUL_UNPOOL, 4 INPUT, 5 STORE_F32, 6 FEAT, 7 FEAT_DIV, 8 INPUT3, 9 MULX, 10 MULX_UNPOOL, 11 FWDX; "
                "internal instantiations: 12 FWDX of a VGG-style layer, 13 / 14 MULX / MULX_UNPOOL with one gain group, 15 MUL with folded filter columns; "
                "`<E, true>` = the pair-mode instantiation (tcgen05 cta_group::2).\n")
        # whole-library mnemonic census with the modifiers kept: .2CTA = cta_group::2 (pair mode), .MULTICAST = cluster multicast
        census = collections.Counter(re.findall(r"\b(UTCHMMA[.\w]*|UTMALDG[.\w]*|UTCBAR[.\w]*|UTCATOMSWS[.\w]*)", sass))
        f.write("\nMnemonic census with modifiers (`.2CTA` = `cta_group::2`, the pair mode of DESIGN.md §4.1c): "
                + ", ".join(f"{k} {v}" for k, v in sorted(census.items())) + ".\n")
    print(open(dst).read()[:3000])


if __name__ == "__main__":
    main(sys.argv[1])
