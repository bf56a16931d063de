run() { name=$1; shift; env "$@" python bench.py --steps 2 --warmup 3 --no-c5 --no-cpu-baseline --no-library-baseline 2>gpurun_out/sweep_skip.err | python -c "
import sys,json
l=[x for x in sys.stdin.read().splitlines() if x.startswith('{')][-1]
d=json.loads(l); print('$name', round(d['value'],1), 'systems/s', round(1e3/d['value'],3), 'ms/system', d['clocks'])"; }
run trailing_plus_rhs BHS_LU_SKIP=111
run gemm_only_no_backward BHS_LU_GEMM_ONLY=1 BHS_LU_SKIP=16
run gemm_only BHS_LU_GEMM_ONLY=1
run trailing_only BHS_LU_SKIP=127
