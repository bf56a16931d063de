# Scan of the sweep engine's (systems per launch) x (groups in flight) shape on the C3 sweep: bash tools/sweep_shapes.sh
B="python bench.py --steps 1 --warmup 3 --no-c5 --no-cpu-baseline --no-library-baseline"
run() { name=$1; shift; env "$@" $B 2>/dev/null | python -c "
import sys,json
l=[x for x in sys.stdin.read().splitlines() if x.startswith('{')][-1]
d=json.loads(l); print('$name', round(d['value'],1), round(d['e2e']['value'],1))"; }
run default X=1
run b2x48 BHS_SWEEP_BATCH=2 BHS_SWEEP_SLOTS=48
run b2x64 BHS_SWEEP_BATCH=2 BHS_SWEEP_SLOTS=64
run b4x24 BHS_SWEEP_BATCH=4 BHS_SWEEP_SLOTS=24
run b4x32 BHS_SWEEP_BATCH=4 BHS_SWEEP_SLOTS=32
run b1x64_cluster BHS_SWEEP_BATCH=1 BHS_SWEEP_SLOTS=64 BHS_LU_LOOKAHEAD=0
run b2x32_cluster BHS_LU_CLUSTER=2
run gemm4m BHS_GEMM_4M=1
