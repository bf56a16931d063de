run() { name=$1; sys=$2; shift; shift; env "$@" python bench.py --steps 5 --warmup 3 --systems $sys --no-c5 --no-cpu-baseline --no-library-baseline 2>/dev/null | python -c "
import sys,json
l=[x for x in sys.stdin.read().splitlines() if x.startswith('{')][-1]
d=json.loads(l); print('$name', round(d['value'],1), round(d['e2e']['value'],1))"; }
run k32_tail1024 32 X=1
run k32_tail0 32 BHS_LU_TAIL=0
run k32_tail2048 32 BHS_LU_TAIL=2048
run k32_tail512 32 BHS_LU_TAIL=512
