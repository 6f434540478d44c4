# Round profile: bench lines, ncu launch list of the bench command, full captures of the hot kernel and the candidate scorer.
set -x
python bench.py --steps 10 --warmup 3 > gpurun_out/r1_bench_default.json 2> gpurun_out/r1_bench_default.err
python bench.py --steps 10 --warmup 3 --gvcf 1 --no-e2e --no-cpu-baseline > gpurun_out/r1_bench_gvcf.json 2>> gpurun_out/r1_bench_default.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1_bench_reference.json 2>> gpurun_out/r1_bench_default.err
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"pileup|score_|tile_|reads_|gather_|prune_|DeviceScan" -c 300 --csv --log-file gpurun_out/r1_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:pileup_vcount -s 4 -c 1 -o gpurun_out/r1_vcount -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:score_candidates -s 4 -c 1 -o gpurun_out/r1_score_candidates -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:pileup_vcount -s 4 -c 1 -o gpurun_out/r1_vcount_gvcf -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --gvcf 1 > /dev/null 2>&1
cat gpurun_out/r1_bench_default.json gpurun_out/r1_bench_gvcf.json gpurun_out/r1_bench_reference.json
