# Is the non-update part of the sweep bound by the kernel launch rate?  Same work with 2 / 4 / 8 / 16 systems per launch (BHS_LU_SKIP=128:
# no trailing updates).  Answer: no (profiles/r02c_launches_sweep_group.txt).  bash tools/sweep_launch_bound.sh
run() { name=$1; shift; env "$@" python bench.py --steps 2 --warmup 3 --no-c5 --no-cpu-baseline --no-library-baseline 2>gpurun_out/sweep_skip.err | python -c "
import sys,json
l=[x for x in sys.stdin.read().splitlines() if x.startswith('{')][-1]
d=json.loads(l); print('$name', round(1e3/d['value'],3), 'ms/system', d['gpu_launches'])"; }
run rest_b4x32 BHS_LU_SKIP=128
run rest_b8x16 BHS_LU_SKIP=128 BHS_SWEEP_BATCH=8 BHS_SWEEP_SLOTS=16
run rest_b16x8 BHS_LU_SKIP=128 BHS_SWEEP_BATCH=16 BHS_SWEEP_SLOTS=8
run rest_b2x64 BHS_LU_SKIP=128 BHS_SWEEP_BATCH=2 BHS_SWEEP_SLOTS=64
run full_b16x8 BHS_SWEEP_BATCH=16 BHS_SWEEP_SLOTS=8
