# The kernel families of the sweep WITHOUT the trailing updates (BHS_LU_SKIP bit 128): what they cost when nothing competes.
run() { name=$1; shift; env "$@" python bench.py --steps 2 --warmup 3 --no-c5 --no-cpu-baseline --no-library-baseline 2>gpurun_out/sweep_skip.err | python -c "
import sys,json
l=[x for x in sys.stdin.read().splitlines() if x.startswith('{')][-1]
d=json.loads(l); print('$name', round(1e3/d['value'],3), 'ms/system')"; }
run everything_but_trailing BHS_LU_SKIP=128
run panels_only BHS_LU_SKIP=252
run all_but_trailing_and_panels BHS_LU_SKIP=129
run assembly_and_probe_only BHS_LU_SKIP=255
