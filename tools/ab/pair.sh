P='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], round(d["ms_per_step"],4), round(d["roofline"]["kernel_ms"],4))'
for rep in 1 2; do
for v in 0 8 5; do python bench.py --no-e2e --no-cpu-baseline --tune-prefetch $v 2>/dev/null | python -c "$P" "vcf tune $v"; done; done
