B="python bench.py --steps 1 --warmup 3 --no-c5 --no-cpu-baseline --no-library-baseline"
run() { name=$1; shift; env "$@" $B 2>/dev/null | python -c "
import sys,json
l=[x for x in sys.stdin.read().splitlines() if x.startswith('{')][-1]
d=json.loads(l); print('$name', round(d['value'],1), round(d['e2e']['value'],1))"; }
run default X=1
run b1x64_tourn BHS_SWEEP_BATCH=1 BHS_SWEEP_SLOTS=64 BHS_LU_CLUSTER=0 BHS_LU_LOOKAHEAD=0
run b1x64_cluster BHS_SWEEP_BATCH=1 BHS_SWEEP_SLOTS=64 BHS_LU_LOOKAHEAD=0
run b1x32_tourn BHS_SWEEP_BATCH=1 BHS_SWEEP_SLOTS=32 BHS_LU_CLUSTER=0 BHS_LU_LOOKAHEAD=0
run b4x16 BHS_SWEEP_BATCH=4 BHS_SWEEP_SLOTS=16
run b2x48 BHS_SWEEP_BATCH=2 BHS_SWEEP_SLOTS=48
run b4x24 BHS_SWEEP_BATCH=4 BHS_SWEEP_SLOTS=24
run b2x32_minb3 BHS_GEMM_MINB=3 BHS_GEMM_BARRIER=1
