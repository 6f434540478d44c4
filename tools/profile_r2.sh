#!/bin/bash
# Round-2 measurement pass (run under gpurun): the launch list of a short bench (every kernel's device time) and the bench line itself.
set -x
mkdir -p gpurun_out
python bench.py --steps 30 --warmup 3 --cpu-repeats 1 2>gpurun_out/bench_err.txt | tee gpurun_out/bench_c2.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 2 > gpurun_out/ncu_bench.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open('gpurun_out/r2_launches.csv', errors='ignore')))
hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
H = rows[hdr]
ki, vi = H.index('Kernel Name'), H.index('Metric Value')
ui = H.index('Metric Unit')
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= vi: continue
    name = r[ki].split('(')[0][:60]
    v = float(r[vi].replace(',', ''))
    if r[ui] == 'ns': v /= 1e3
    elif r[ui] == 'ms': v *= 1e3
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1; a[1] += v
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f"{t/n:10.1f} us x {n:4d}  {k}")
PY
