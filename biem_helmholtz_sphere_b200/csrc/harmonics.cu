// K2: orthonormal harmonics on S^{d-1} for chain coordinate trees, ultrasphere (Phase(0)) ordering.
//
//   Y_h = prod_i f^{(desc_i)}_{n_i, n_{i+1}}(theta_i) * e^{i m phi} / sqrt(2 pi)        
//
// One warp per direction: lanes build the node-function tables column-wise in shared memory
// (three-term recurrences, coefficients from the plan), then stream the flattened harmonics out with
// coalesced 16-byte stores.  Replaces ush.harmonics (_biem.py:922) and the harmonics inside
// ush.expand / ush.harmonics_translation_coef (_biem.py:627,697).
#include "harmonics.cuh"

__global__ void __launch_bounds__(128) harmonics_kernel(HarmTables tb, int Lb, const int32_t* __restrict__ idx,
                                                        int Hb, const double* __restrict__ xyz, int64_t npts,
                                                        const double* __restrict__ scale, int conj_out,
                                                        cplx* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t per_warp = harm_smem_bytes_per_warp(tb.d, Lb);
    double* F;
    cplx* E;
    harm_smem_carve(smem_raw + per_warp * warp, Lb, F, E);
    const int s = tb.d - 1;
    for (int64_t p = (int64_t)blockIdx.x * warps + warp; p < npts; p += (int64_t)gridDim.x * warps) {
        double x[BHS_MAX_NODES + 2];
        for (int i = 0; i < tb.d; ++i) x[i] = xyz[(int64_t)i * npts + p];
        warp_harmonic_tables(tb, Lb, x, F, E, lane);
        __syncwarp();
        double sc = scale ? scale[p] : 1.0;
        for (int h = lane; h < Hb; h += 32) {
            cplx y = harmonic_from_tables(tb, Lb, idx + (int64_t)h * s, F, E);
            if (conj_out) y.y = -y.y;
            out[p * Hb + h] = cscale(y, sc);
        }
        __syncwarp();
    }
}

static int launch_harmonics(const bhs_plan* plan, int Lb, const int32_t* d_idx, int Hb, const double* d_xyz,
                            int64_t npts, const double* d_scale, int conj_out, cplx* d_out, cudaStream_t st) {
    if (npts <= 0) return BHS_OK;
    HarmTables tb = harm_tables_of(plan);
    int warps = 4;
    while (warps > 1 && harm_smem_bytes_per_warp(plan->d, Lb) * warps > 200 * 1024) warps >>= 1;
    size_t smem = harm_smem_bytes_per_warp(plan->d, Lb) * warps;
    if (smem > 200 * 1024) return BHS_ERR_UNSUPPORTED;
    cudaFuncSetAttribute(harmonics_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int64_t blocks = (npts + warps - 1) / warps;
    if (blocks > bhs_sm_count() * 16) blocks = bhs_sm_count() * 16;
    harmonics_kernel<<<(unsigned)blocks, warps * 32, smem, st>>>(tb, Lb, d_idx, Hb, d_xyz, npts, d_scale, conj_out,
                                                                 d_out);
    BHS_CHECK_LAUNCH();
    return BHS_OK;
}

extern "C" int bhs_harmonics(const bhs_plan_t* plan, int use_double_band, const double* d_xyz, int64_t npts,
                             double* d_out, void* stream) {
    if (!plan || !d_xyz || !d_out || npts < 0) return BHS_ERR_INVALID;
    if (use_double_band)
        return launch_harmonics(plan, plan->L2, plan->d_idx2, plan->H2, d_xyz, npts, nullptr, 0, (cplx*)d_out,
                                (cudaStream_t)stream);
    return launch_harmonics(plan, plan->n_end, plan->d_idx, plan->H, d_xyz, npts, nullptr, 0, (cplx*)d_out,
                            (cudaStream_t)stream);
}

// WY[q][h] = w_q conj(Y_h(y_q)) on the RHS quadrature nodes (called once from bhs_plan_create)
int bhs_fill_WY(bhs_plan* p) {
    return launch_harmonics(p, p->n_end, p->d_idx, p->H, p->d_qdirs, p->Q, p->d_qw, 1, p->d_WY, 0);
}
