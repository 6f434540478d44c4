# Round profile, part 2: the hot kernel in gVCF mode and the PNIB16 staging kernel.
set -x
ncu --set full --clock-control none -k regex:pileup_nib -s 4 -c 1 -o gpurun_out/r1_nib_gvcf -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --gvcf 1 > /dev/null 2>&1
ncu --set full --clock-control none -k regex:nib_scatter -s 2 -c 1 -o gpurun_out/r1_nib_scatter -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out
