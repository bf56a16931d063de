run() { name=$1; sys=$2; shift; shift; env "$@" python bench.py --steps 5 --warmup 3 --systems $sys --no-c5 --no-cpu-baseline --no-library-baseline 2>/dev/null | python -c "
import sys,json
l=[x for x in sys.stdin.read().splitlines() if x.startswith('{')][-1]
d=json.loads(l); print('$name', round(d['value'],1), round(d['e2e']['value'],1))"; }
run k32_b2x16 32 X=1
run k32_b4x8 32 BHS_SWEEP_BATCH=4
run k32_b1x32 32 BHS_SWEEP_BATCH=1
run k32_b8x4 32 BHS_SWEEP_BATCH=8
run k64_b2x32 64 X=1
run k64_b4x16 64 BHS_SWEEP_BATCH=4
run k128_b4x32 128 X=1
