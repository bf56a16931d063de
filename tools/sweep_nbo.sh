# Scan of the outer block width of the LU (BHS_LU_NBO) on the C3 sweep and on the C5 factorisation:  bash tools/sweep_nbo.sh
run() { name=$1; shift; env "$@" python bench.py --steps 2 --warmup 3 --no-c5 --no-cpu-baseline --no-library-baseline 2>gpurun_out/sweep_nbo.err | python -c "
import sys,json
l=[x for x in sys.stdin.read().splitlines() if x.startswith('{')][-1]
d=json.loads(l); print('$name', round(d['value'],1), round(d['e2e']['value'],1))"; }
run nbo128 BHS_LU_NBO=128
run nbo256 BHS_LU_NBO=256
run nbo384 BHS_LU_NBO=384
run nbo512 BHS_LU_NBO=512
