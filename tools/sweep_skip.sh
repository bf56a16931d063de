# What each kernel family costs the C3 sweep: the sweep with that family NOT launched (BHS_LU_SKIP, wrong results by design).
# bash tools/sweep_skip.sh
run() { name=$1; shift; env "$@" python bench.py --steps 2 --warmup 3 --no-c5 --no-cpu-baseline --no-library-baseline 2>gpurun_out/sweep_skip.err | python -c "
import sys,json
l=[x for x in sys.stdin.read().splitlines() if x.startswith('{')][-1]
d=json.loads(l); print('$name', round(d['value'],1), 'systems/s', round(1e3/d['value'],3), 'ms/system')"; }
run full X=1
run no_panel BHS_LU_SKIP=1
run no_permute BHS_LU_SKIP=2
run no_trsm32 BHS_LU_SKIP=4
run no_inner_gemm BHS_LU_SKIP=8
run no_rhs BHS_LU_SKIP=16
run only_updates_K128plus BHS_LU_SKIP=31
run gemm_only BHS_LU_GEMM_ONLY=1
