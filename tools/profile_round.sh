# Round profile, part 1 (each gpurun call may bring back 64 MiB): bench lines, ncu launch list of the bench command, full captures of the hot kernel
# (PNIB16) and the candidate scorer. Part 2 (tools/profile_round2.sh): gVCF mode and the PNIB16 staging kernel.
set -x
python bench.py --steps 20 --warmup 3 > gpurun_out/r1_bench_default.json 2> gpurun_out/r1_bench_default.err
python bench.py --steps 20 --warmup 3 --gvcf 1 --no-e2e --no-cpu-baseline > gpurun_out/r1_bench_gvcf.json 2>> gpurun_out/r1_bench_default.err
python bench.py --steps 20 --warmup 3 --tune-prefetch 9 --no-e2e --no-cpu-baseline > gpurun_out/r1_bench_ptile32.json 2>> gpurun_out/r1_bench_default.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1_bench_reference.json 2>> gpurun_out/r1_bench_default.err
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"pileup|score_|tile_|nib_|unpack_|apply_|reads_|gather_|prune_|DeviceScan" -c 400 --csv --log-file gpurun_out/r1_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:pileup_nib -s 4 -c 1 -o gpurun_out/r1_nib -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
ncu --set full --clock-control none -k regex:score_candidates -s 4 -c 1 -o gpurun_out/r1_score_candidates -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
ls -la gpurun_out
