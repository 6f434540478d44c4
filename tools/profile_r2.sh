#!/bin/bash
# Round-2 measurement pass (run under gpurun, one GPU): bench lines of the five configs + the reference arm, the launch list of a short bench (every
# kernel's device time, cold-cache and serialised), and one ncu --set full capture of the hot kernel per benched config (details + raw metrics as text;
# the .ncu-rep files stay on the box: two of them exceed what gpurun_out carries back).
mkdir -p gpurun_out/r2
O=gpurun_out/r2
python bench.py --steps 100 --warmup 10 2>$O/bench_c2.err | tail -1 > $O/r2_bench_c2.json
python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > $O/r2_bench_c2_reference.json
for c in c3 c4 c5; do python bench.py --config $c --steps 20 --warmup 3 --cpu-repeats 1 2>$O/bench_$c.err | tail -1 > $O/r2_bench_$c.json; done
python bench.py --config c1 --steps 20 --warmup 3 2>$O/bench_c1.err | tail -1 > $O/r2_bench_c1.json
for c in c2 c3; do
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"pvert|pileup|score_|cand|reads_|sink|unpack|apply_seq|amplicon|dirs_from" -c 600 --csv --log-file $O/r2_launches_$c.csv \
      python bench.py --config $c --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --e2e-jobs 1 > /dev/null 2>&1
  python tools/launch_summary.py $O/r2_launches_$c.csv 40 > $O/r2_launches_${c}_summary.txt
  ncu --set full --clock-control none --import-source on -k regex:pileup_pvert_score_kernel -s 3 -c 1 -o /tmp/hot_$c -f \
      python bench.py --config $c --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
  ncu -i /tmp/hot_$c.ncu-rep --page details > $O/r2_pvert_hot_${c}_v2_ncu_details.txt 2>&1
  ncu -i /tmp/hot_$c.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv, sys
rows = list(csv.reader(sys.stdin))
H = rows[0]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'lts__t_sector_hit_rate.pct',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__issue_active.avg.pct_of_peak_sustained_active']
for r in rows[2:]:
    for w in want:
        if w in H: print(w, r[H.index(w)], rows[1][H.index(w)])
" > $O/r2_pvert_hot_${c}_v2_ncu_raw.txt
done
ls -la $O
