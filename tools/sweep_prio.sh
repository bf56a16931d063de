# Scan: stream priorities of the sweep groups (BHS_SWEEP_PRIO = number of priority levels cycled over the groups) and
# group shapes, for a short sweep (32 systems = one rank's share at 8 GPUs) and the full 256:  bash tools/sweep_prio.sh
run() { name=$1; sys=$2; shift; shift; env "$@" python bench.py --steps 3 --warmup 3 --systems $sys --no-c5 --no-cpu-baseline --no-library-baseline 2>gpurun_out/sweep_prio.err | python -c "
import sys,json
l=[x for x in sys.stdin.read().splitlines() if x.startswith('{')][-1]
d=json.loads(l); print('$name', round(d['value'],1), round(d['e2e']['value'],1))"; }
run k32_default 32 X=1
run k32_prio6 32 BHS_SWEEP_PRIO=6
run k32_b1x32 32 BHS_SWEEP_BATCH=1
run k32_b1x32_prio6 32 BHS_SWEEP_BATCH=1 BHS_SWEEP_PRIO=6
run k32_b4x8 32 BHS_SWEEP_BATCH=4
run k64_default 64 X=1
run k64_prio6 64 BHS_SWEEP_PRIO=6
run k256_default 256 X=1
run k256_prio6 256 BHS_SWEEP_PRIO=6
run k256_prio3 256 BHS_SWEEP_PRIO=3
