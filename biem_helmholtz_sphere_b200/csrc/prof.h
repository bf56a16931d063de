// Optional per-category device timing (CUDA events around groups of launches) used by bench.py to split a
// step into its kernels.  Off by default: bhs_prof_begin/end are no-ops unless bhs_profile(1) was called.
// Not usable while a stream is being captured into a CUDA graph.
#pragma once
#include <cuda_runtime.h>

void bhs_prof_begin(int cat, cudaStream_t st);
void bhs_prof_end(int cat, double work, cudaStream_t st);
