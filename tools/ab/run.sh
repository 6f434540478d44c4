# A/B of prebuilt library variants (tools/ab/lib_<name>.so) on one box: VCF-mode and gVCF-mode hot-kernel times
cp pisces_b200/libpisces_b200.so /tmp/lib_current.so
P='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], round(d["ms_per_step"],4), round(d["roofline"]["kernel_ms"],4), round(d["roofline"]["frac"],3))'
for rep in 1 2; do
for v in "$@"; do
  cp tools/ab/lib_$v.so pisces_b200/libpisces_b200.so
  python bench.py --no-e2e --no-cpu-baseline 2>/dev/null | python -c "$P" "$v vcf"
  python bench.py --gvcf 1 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "$P" "$v gvcf"
done; done
cp /tmp/lib_current.so pisces_b200/libpisces_b200.so
