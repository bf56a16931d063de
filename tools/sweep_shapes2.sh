# Scan of (outer block width) x (systems per launch x groups in flight) on the C3 sweep:  bash tools/sweep_shapes2.sh
run() { name=$1; shift; env "$@" python bench.py --steps 2 --warmup 3 --no-c5 --no-cpu-baseline --no-library-baseline 2>gpurun_out/sweep_shapes2.err | python -c "
import sys,json
l=[x for x in sys.stdin.read().splitlines() if x.startswith('{')][-1]
d=json.loads(l); print('$name', round(d['value'],1), round(d['e2e']['value'],1))"; }
run nbo256_b4x32 X=1
run nbo256_b4x24 BHS_SWEEP_SLOTS=24
run nbo256_b4x48 BHS_SWEEP_SLOTS=48
run nbo256_b2x48 BHS_SWEEP_BATCH=2 BHS_SWEEP_SLOTS=48
run nbo256_b8x16 BHS_SWEEP_BATCH=8 BHS_SWEEP_SLOTS=16
run nbo384_b4x32 BHS_LU_NBO=384
run nbo384_b8x16 BHS_LU_NBO=384 BHS_SWEEP_BATCH=8 BHS_SWEEP_SLOTS=16
run nbo256_b4x32_cluster BHS_LU_CLUSTER=2
