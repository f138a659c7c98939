#!/bin/bash
# Layer-0 (224^2, 64 -> 3, INPUT3) attribution: row walk on/off x timing switches (LRPX_TC_DEBUG bit 0: epilogue without
# global loads/stores, bit 1: no MMAs, bit 2: no A traffic), then an ncu launch list of one decoder call.
mkdir -p gpurun_out
for walk in 1 0; do for dbg in 0 1 2 4 5; do
  echo "walk=$walk debug=$dbg $(LRPX_TC_WALK=$walk LRPX_TC_DEBUG=$dbg LAYERS=0 REPS=9 timeout 100 python scripts/one_layer.py 2>&1 | grep 'layer\|rror' | head -3)"
done; done 2>&1 | tee gpurun_out/l0_exp.log
timeout 200 python scripts/dec_probe.py > gpurun_out/dec_plain.log 2>&1
tail -1 gpurun_out/dec_plain.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/dec_launches.csv python scripts/dec_probe.py > gpurun_out/dec_ncu.log 2>&1
echo "ncu rc=$?"
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(l for l in open("gpurun_out/dec_launches.csv") if l.startswith('"'))]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    try: v = float(r[vi].replace(",", ""))
    except ValueError: continue
    k = r[ki][:70]
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f"{k:72s} {n:5d} {t/1e3:10.1f} us total {t/n/1e3:8.2f} us each")
PY
