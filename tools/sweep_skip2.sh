run() { name=$1; shift; env "$@" python bench.py --steps 2 --warmup 3 --no-c5 --no-cpu-baseline --no-library-baseline 2>gpurun_out/sweep_skip.err | python -c "
import sys,json
l=[x for x in sys.stdin.read().splitlines() if x.startswith('{')][-1]
d=json.loads(l); print('$name', round(d['value'],1), 'systems/s', round(1e3/d['value'],3), 'ms/system')"; }
run skip_inblock128 BHS_LU_SKIP=32
run skip_trsm128 BHS_LU_SKIP=64
run skip_both128 BHS_LU_SKIP=96
run skip_all_but_trailing BHS_LU_SKIP=127
run nbo128_skip_all_but_trailing BHS_LU_SKIP=127 BHS_LU_NBO=128
run nbo128_gemm_only BHS_LU_GEMM_ONLY=1 BHS_LU_NBO=128
