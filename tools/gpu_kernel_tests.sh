#!/bin/bash
# Run each kernel test group in its own process so a CUDA fault in one does not poison the rest.
mkdir -p gpurun_out
for g in fp64 bessel harmonics plan_tables rhs_expand assemble zgemm zgesv zgetrf uscat; do
  echo "=== $g" >> gpurun_out/kernels.log
  CUDA_LAUNCH_BLOCKING=${BLOCKING:-0} timeout 600 python -m pytest tests/test_gpu_kernels.py -q -s -m gpu -k "$g" 2>&1 \
    | grep -E "rel err|abs err|scaled|FP64|CUDA error|Error|assert|passed|failed|FAILED" >> gpurun_out/kernels.log
done
grep -E "===|passed|failed" gpurun_out/kernels.log
