// K5: dense complex128 solve  A x = f  (row-major), replacing batch_tensorsolve.btensorsolve -> zgesv
// (_biem.py:797).
//
// Right-looking recursive blocked LU with TOURNAMENT partial pivoting (CALU): per 32-column panel the
// pivot rows are chosen by a reduction tree of register-resident Gaussian eliminations (128 rows per
// CTA, one row per thread), so no step of the factorisation needs a grid-wide sync.  Outer blocks are
// 128 wide; all Schur-complement work (inside and outside the outer block) is one kernel:
//
//   zgemm_sub_kernel :  C -= A * B  on FP64 tensor cores (mma.sync m8n8k4.f64 -> SASS DMMA.8x8x4), complex
//   arithmetic through the real embedding  [ar ai] x [[br bi],[-bi br]], operands pre-packed by the
//   triangular-solve kernels into the exact shared-memory image (padded, conflict-free) so that every
//   pipeline stage is two 1-D TMA bulk copies (cp.async.bulk + mbarrier), 3 stages deep.
#include <cstdlib>

#include <cuda.h>
#include <type_traits>  // CUtensorMap types only; the encoder is fetched with cudaGetDriverEntryPoint

#include "common.cuh"
#include "prof.h"

#define LU_NB 32     // panel width
#define LU_NBO 128   // outer block width
#define LU_R 128     // rows per tournament CTA
#define LU_DBLK (LU_NB * LU_NB)  // complex slots per system: factored diagonal block of the current panel
#define LU_PMAP (2 * LU_NB)     // ints per panel: its net row map (kept for every panel: look-ahead applies them late)
#define G_TM 64
#define G_TN 64
#define G_KC 8
#define G_LDS 20
#define G_A_STAGE (G_TM * G_LDS)
#define G_B_STAGE (2 * G_TN * G_LDS)
#define G_STAGES 3
#define G_THREADS 128

static inline int64_t al256(int64_t v) { return (v + 255) & ~(int64_t)255; }
static inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Batched factorisation: every kernel of this file handles `nbatch` independent systems of one size in lock step, system
// z = blockIdx.z, so that a k-sweep issues one launch per step for a whole group of systems.  Per-system strides
// (in elements of the respective array):
struct LuBatch {
    int64_t sA;     // matrix (complex)
    int64_t sRhs;   // right-hand sides (complex)
    int64_t sIpiv;  // pivots (int32)
    int64_t sCand;  // tournament candidate lists (int32)
    int64_t sDblk;  // factored diagonal block scratch (complex)
    int64_t sPmap;  // net row maps of all panels (int32)
    int64_t sLp;    // packed L image (double)
    int64_t sUp;    // packed U image (double)
};

// =====================================================================================================
// GEMM:  C[rows, cols] -= Lp * Up
// =====================================================================================================
struct GemmArgs {
    const double* Lp;
    const double* Up;
    int nks_total;  // packed stages per tile (pitch of both packed buffers)
    int ks0, nks;   // stage window used by this product
    cplx* C;
    int64_t ldc;
    int64_t row_base, col_base;  // global row / col of packed tile (0, 0)
    int rt0, ct0;                // first row / col tile of this launch
    int64_t row_lo, row_hi, col_lo, col_hi;  // output window (global indices, half open)
    int64_t sC, sLp, sUp;                    // per-system strides (blockIdx.z)
};

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(G_THREADS, 2) zgemm_sub_kernel(GemmArgs g) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full[G_STAGES];
    double* smA = reinterpret_cast<double*>(smem_raw);
    double* smB = smA + G_STAGES * G_A_STAGE;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 1, wn = warp & 1;
    const int rt = g.rt0 + blockIdx.y, ct = g.ct0 + blockIdx.x;
    g.C += (int64_t)blockIdx.z * g.sC;
    const double* Lt = g.Lp + (int64_t)blockIdx.z * g.sLp + ((int64_t)rt * g.nks_total + g.ks0) * G_A_STAGE;
    const double* Ut = g.Up + (int64_t)blockIdx.z * g.sUp + ((int64_t)ct * g.nks_total + g.ks0) * G_B_STAGE;
    if (tid == 0) {
        for (int s = 0; s < G_STAGES; ++s) mbar_init(&full[s], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0) {
        for (int s = 0; s < G_STAGES && s < g.nks; ++s) {
            mbar_expect_tx(&full[s], (uint32_t)((G_A_STAGE + G_B_STAGE) * sizeof(double)));
            tma_load_1d(smA + s * G_A_STAGE, Lt + (int64_t)s * G_A_STAGE, G_A_STAGE * sizeof(double), &full[s]);
            tma_load_1d(smB + s * G_B_STAGE, Ut + (int64_t)s * G_B_STAGE, G_B_STAGE * sizeof(double), &full[s]);
        }
    }
    // accumulators <- C
    double acc[4][8][2];
    const int64_t row0 = g.row_base + (int64_t)rt * G_TM + wm * 32 + (lane >> 2);
    const int64_t col0 = g.col_base + (int64_t)ct * G_TN + wn * 32 + (lane & 3);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t r = row0 + 8 * i;
        const bool rv = r >= g.row_lo && r < g.row_hi;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int64_t c = col0 + 4 * j;
            cplx v = cmake(0.0, 0.0);
            if (rv && c >= g.col_lo && c < g.col_hi) v = g.C[r * g.ldc + c];
            acc[i][j][0] = v.x;
            acc[i][j][1] = v.y;
        }
    }
    const int a_off = (wm * 32 + (lane >> 2)) * G_LDS + (lane & 3);
    const int b_off = (wn * 64 + (lane >> 2)) * G_LDS + (lane & 3);
    for (int it = 0; it < g.nks; ++it) {
        const int s = it % G_STAGES;
        mbar_wait(&full[s], (it / G_STAGES) & 1);
        const double* As = smA + s * G_A_STAGE + a_off;
        const double* Bs = smB + s * G_B_STAGE + b_off;
#pragma unroll
        for (int k4 = 0; k4 < (2 * G_KC) / 4; ++k4) {
            double a[4], b[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[i * 8 * G_LDS + 4 * k4];
#pragma unroll
            for (int j = 0; j < 8; ++j) b[j] = Bs[j * 8 * G_LDS + 4 * k4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
        __syncthreads();
        if (tid == 0 && it + G_STAGES < g.nks) {
            fence_proxy_async();
            mbar_expect_tx(&full[s], (uint32_t)((G_A_STAGE + G_B_STAGE) * sizeof(double)));
            tma_load_1d(smA + s * G_A_STAGE, Lt + (int64_t)(it + G_STAGES) * G_A_STAGE, G_A_STAGE * sizeof(double), &full[s]);
            tma_load_1d(smB + s * G_B_STAGE, Ut + (int64_t)(it + G_STAGES) * G_B_STAGE, G_B_STAGE * sizeof(double), &full[s]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t r = row0 + 8 * i;
        const bool rv = r >= g.row_lo && r < g.row_hi;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int64_t c = col0 + 4 * j;
            if (rv && c >= g.col_lo && c < g.col_hi) g.C[r * g.ldc + c] = cmake(acc[i][j][0], acc[i][j][1]);
        }
    }
}

static const size_t G_SMEM = (size_t)G_STAGES * (G_A_STAGE + G_B_STAGE) * sizeof(double);

// =====================================================================================================
// GEMM, operands straight from the row-major matrices:  C[rows, cols] -= L[rows, k0 .. k0+K) * U[k0 .. k0+K, cols]
// =====================================================================================================
// No packed operand images: both operands are fetched by 2-D tensor-map TMA (cp.async.bulk.tensor, SASS UTMALDG) from the
// row-major complex matrices, 8 complex k per pipeline stage, 128-byte swizzle:
//   L tile  : box {16 doubles, 64 rows}            -> smem [64 rows][128 B]           (8 KB)
//   U tile  : 8 boxes {16 doubles, 8 k-rows}, one per group of 8 columns -> smem [8 groups][8 k][128 B]   (8 KB)
// The real embedding of the complex product is formed in REGISTERS.  One mma.m8n8k4 step contracts the four reals
// (re, im) of the complex k-indices q and q + 4 of the stage (the pairing that keeps every fragment load free of bank
// conflicts under the 128-byte swizzle).  Output columns are split by part: one 8x8 tile holds the REAL parts of 8
// consecutive complex columns, its twin the IMAGINARY parts,
//   re-tile:  [ar, -ai] . [br, bi]      im-tile:  [ar, ai] . [bi, br]
// so the B fragments of both are the two halves of ONE 16-byte shared-memory load, the A fragment of the re-tile is the
// im-tile's with the sign of the odd lanes flipped, and every thread ends up with complete complex numbers (two adjacent
// columns per 8x8 pair): 32-byte contiguous C accesses.  Accumulators hold -C (negated on load and store), so that the
// tensor cores only ever add.
#define T_STAGES 4
#define T_STAGE_BYTES 8192  // per operand
struct TmaGemmArgs {
    cplx* C;
    int64_t ldc;
    int64_t row_base, col_base;  // global row / col of tile (0, 0): C indices and TMA coordinates (rows of L, columns of U)
    int rt0, ct0;                // first row / col tile of this launch
    int64_t row_lo, row_hi, col_lo, col_hi;  // output window (half open)
    int64_t sC;                  // per-system stride of C (blockIdx.z)
    int k0, nks;                 // first k (column of L = row of U) and number of 8-wide stages
};

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const void* tmap, int x, int y, int z, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
            smem_u32(smem_dst)),
        "l"(tmap), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar))
        : "memory");
}

__device__ __forceinline__ double flip_sign_if(double v, unsigned mask_hi) {
    return __hiloint2double(__double2hiint(v) ^ (int)mask_hi, __double2loint(v));
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// PROD = true: no block barrier inside the k loop.  Every warp hands a finished stage back through an `empty` mbarrier (one
// arrival per warp); one iteration LATER -- when the other warps have almost certainly arrived too -- the warps' first lanes
// refill that stage (each its share of the nine box loads).  Prefetch distance T_STAGES - 1.
// PROD = false: one block barrier per stage, refill right behind it.
template <int MINB, bool PROD>
__global__ void __launch_bounds__(G_THREADS, MINB)
    zgemm_tma_kernel(const __grid_constant__ CUtensorMap mapL, const __grid_constant__ CUtensorMap mapU, TmaGemmArgs g) {
    extern __shared__ unsigned char smem_dyn[];
    __shared__ __align__(8) uint64_t full[T_STAGES];
    __shared__ __align__(8) uint64_t empty[T_STAGES];
    // 1024-byte alignment: the 128-byte swizzle pattern is a function of the shared-memory address bits 7..9
    unsigned char* smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    unsigned char* smA = smem;
    unsigned char* smB = smem + T_STAGES * T_STAGE_BYTES;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 1, wn = warp & 1;
    const int rt = g.rt0 + blockIdx.y, ct = g.ct0 + blockIdx.x;
    const int z = blockIdx.z;
    g.C += (int64_t)z * g.sC;
    const int trow = (int)(g.row_base + (int64_t)rt * G_TM);  // TMA coordinates of the tile
    const int tcol = (int)(g.col_base + (int64_t)ct * G_TN);
    // The nine box loads of a stage are issued by the four warps' first lanes (warp w: operand boxes 3w .. 3w+2 of
    // {L, U group 0, .., U group 7}), so that no single warp carries the whole issue cost of a stage; the barrier's
    // expected byte count is posted by warp 0 (complete_tx of the other warps' loads may arrive first: the phase cannot
    // complete before its one pending arrival).
    auto issue = [&](int it, int s, int part) {
        const int kk = g.k0 + it * G_KC;
        if (part <= 0) mbar_expect_tx(&full[s], 2 * T_STAGE_BYTES);
        const int b0 = part < 0 ? 0 : (part == 0 ? 0 : 3 * part - 1), b1 = part < 0 ? 9 : (part == 3 ? 9 : 3 * part + 2);
        for (int b = b0; b < b1; ++b) {
            if (b == 0)
                tma_load_3d(smA + s * T_STAGE_BYTES, &mapL, 2 * kk, trow, z, &full[s]);
            else
                tma_load_3d(smB + s * T_STAGE_BYTES + (b - 1) * 1024, &mapU, 2 * (tcol + 8 * (b - 1)), kk, z, &full[s]);
        }
    };
    if (tid == 0) {
        for (int s = 0; s < T_STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 4);
        }
        mbar_fence_init();
    }
    __syncthreads();
    if (lane == 0)
        for (int s = 0; s < T_STAGES && s < g.nks; ++s) issue(s, s, warp);
    // accumulators <- -C : acc_re / acc_im [row block i][column block j][2 adjacent columns]
    double are[4][4][2], aim[4][4][2];
    const int64_t row0 = g.row_base + (int64_t)rt * G_TM + wm * 32 + (lane >> 2);
    const int64_t col0 = g.col_base + (int64_t)ct * G_TN + wn * 32 + 2 * (lane & 3);
    // interior tiles (the great majority) skip the per-element window tests
    const bool interior = (int64_t)trow >= g.row_lo && (int64_t)trow + G_TM <= g.row_hi && (int64_t)tcol >= g.col_lo &&
                          (int64_t)tcol + G_TN <= g.col_hi;
    cplx* const Cw = g.C + row0 * g.ldc + col0;
    if (interior) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const cplx v = Cw[(int64_t)(8 * i) * g.ldc + 8 * j + e];
                    are[i][j][e] = -v.x;
                    aim[i][j][e] = -v.y;
                }
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int64_t r = row0 + 8 * i;
            const bool rv = r >= g.row_lo && r < g.row_hi;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int64_t c = col0 + 8 * j + e;
                    cplx v = cmake(0.0, 0.0);
                    if (rv && c >= g.col_lo && c < g.col_hi) v = g.C[r * g.ldc + c];
                    are[i][j][e] = -v.x;
                    aim[i][j][e] = -v.y;
                }
            }
        }
    }
    const unsigned odd_sign = (lane & 1) ? 0x80000000u : 0u;
    const int kx = (lane & 2) ? 4 : 0, rx = lane >> 2, odd8 = (lane & 1) * 8;
    const int a_base = (wm * 32 + rx) * 128 + odd8;
    const int b_base = wn * 4 * 1024;
    // byte offsets of this lane's 16-byte chunk in the four k-steps of a stage (128-byte swizzle: chunk ^ row)
    int a_sw[4], b_sw[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        a_sw[q] = ((q + kx) ^ rx) << 4;
        b_sw[q] = (q + kx) * 128 + ((rx ^ (q + kx)) << 4);
    }
    for (int it = 0; it < g.nks; ++it) {
        const int s = it % T_STAGES;
        mbar_wait(&full[s], (it / T_STAGES) & 1);
        const unsigned char* As = smA + s * T_STAGE_BYTES + a_base;
        const unsigned char* Bs = smB + s * T_STAGE_BYTES + b_base;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            double ap[4], ac[4], bre[4], bim[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                ap[i] = *reinterpret_cast<const double*>(As + i * 1024 + a_sw[q]);
                ac[i] = flip_sign_if(ap[i], odd_sign);
            }
            // B fragments: even lanes (real k') take (re, im) of U[k][n] for the (re-tile, im-tile), odd lanes (im, re)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                bre[j] = *reinterpret_cast<const double*>(Bs + j * 1024 + b_sw[q] + odd8);
                bim[j] = *reinterpret_cast<const double*>(Bs + j * 1024 + b_sw[q] + (8 - odd8));
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    dmma884(are[i][j][0], are[i][j][1], ac[i], bre[j]);
                    dmma884(aim[i][j][0], aim[i][j][1], ap[i], bim[j]);
                }
        }
        if (PROD) {
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&empty[s]);
                if (it >= 1 && it - 1 + T_STAGES < g.nks) {
                    const int sp = (it - 1) % T_STAGES;
                    mbar_wait(&empty[sp], ((it - 1) / T_STAGES) & 1);
                    issue(it - 1 + T_STAGES, sp, warp);
                }
            }
        } else {
            __syncthreads();
            if (lane == 0 && it + T_STAGES < g.nks) issue(it + T_STAGES, s, warp);
        }
    }
    if (interior) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int e = 0; e < 2; ++e) Cw[(int64_t)(8 * i) * g.ldc + 8 * j + e] = cmake(-are[i][j][e], -aim[i][j][e]);
        return;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t r = row0 + 8 * i;
        const bool rv = r >= g.row_lo && r < g.row_hi;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int64_t c = col0 + 8 * j + e;
                if (rv && c >= g.col_lo && c < g.col_hi) g.C[r * g.ldc + c] = cmake(-are[i][j][e], -aim[i][j][e]);
            }
        }
    }
}
static const size_t T_SMEM = (size_t)T_STAGES * 2 * T_STAGE_BYTES + 1024;

// =====================================================================================================
// The same update with THREE real products per complex one ("3M", the ZGEMM3M scheme of the BLAS):
//     T1 = Ar Br,  T2 = Ai Bi,  T3 = (Ar + Ai)(Br + Bi);     Re(AB) = T1 - T2,   Im(AB) = T3 - T1 - T2
// i.e. 6 instead of 8 real flops per complex multiply-add on the tensor cores: a quarter of the DMMA work of the real
// embedding above is gone.  The price is the usual one of 3M: the imaginary part is formed by cancellation, so its error is
// bounded normwise (by eps (|Ar| + |Ai|)(|Br| + |Bi|)), not componentwise; LU with partial pivoting only needs the normwise
// bound (measured: residuals and densities unchanged at the 1e-15 level, tests/test_gpu_kernels.py).
// Operands exactly as above (tensor-map TMA straight from the matrices, 8 complex k per stage, 128-byte swizzle), CTA tile
// 64 x 32, four warps of 32 x 16, three resident CTAs per SM.  A k-step contracts the four complex k of one parity of the
// stage (k = 2 (lane & 3) + p: the pairing that keeps the 16-byte fragment loads conflict-free under the swizzle); every
// fragment load brings (re, im) of one element and the third operand is one DADD away.
// C folds into the accumulators at load time: T2 <- Cr, T3 <- Cr - Ci, T1 <- 0 give  C - AB = (T2 - T1, T1 + T2 - T3).
#define M3_TN 32
#define M3_STAGES 5
#define M3_A_BYTES 8192
#define M3_B_BYTES 4096
__global__ void __launch_bounds__(G_THREADS, 3)
    zgemm3m_tma_kernel(const __grid_constant__ CUtensorMap mapL, const __grid_constant__ CUtensorMap mapU, TmaGemmArgs g) {
    extern __shared__ unsigned char smem_dyn[];
    __shared__ __align__(8) uint64_t full[M3_STAGES];
    __shared__ __align__(8) uint64_t empty[M3_STAGES];
    unsigned char* smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    unsigned char* smA = smem;
    unsigned char* smB = smem + M3_STAGES * M3_A_BYTES;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 1, wn = warp & 1;
    const int rt = g.rt0 + blockIdx.y, ct = g.ct0 + blockIdx.x;
    const int z = blockIdx.z;
    g.C += (int64_t)z * g.sC;
    const int trow = (int)(g.row_base + (int64_t)rt * G_TM);
    const int tcol = (int)(g.col_base + (int64_t)ct * M3_TN);
    // five box loads per stage (L, four groups of 8 columns of U), shared by the warps' first lanes
    auto issue = [&](int it, int s, int part) {
        const int kk = g.k0 + it * G_KC;
        if (part == 0) {
            mbar_expect_tx(&full[s], M3_A_BYTES + M3_B_BYTES);
            tma_load_3d(smA + s * M3_A_BYTES, &mapL, 2 * kk, trow, z, &full[s]);
            tma_load_3d(smB + s * M3_B_BYTES, &mapU, 2 * tcol, kk, z, &full[s]);
        } else {
            tma_load_3d(smB + s * M3_B_BYTES + part * 1024, &mapU, 2 * (tcol + 8 * part), kk, z, &full[s]);
        }
    };
    if (tid == 0) {
        for (int s = 0; s < M3_STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 4);
        }
        mbar_fence_init();
    }
    __syncthreads();
    if (lane == 0)
        for (int s = 0; s < M3_STAGES && s < g.nks; ++s) issue(s, s, warp);
    // accumulators [row block i][column block j][2 adjacent columns]
    double t1[4][2][2], t2[4][2][2], t3[4][2][2];
    const int64_t row0 = (int64_t)trow + wm * 32 + (lane >> 2);
    const int64_t col0 = (int64_t)tcol + wn * 16 + 2 * (lane & 3);
    const bool interior = (int64_t)trow >= g.row_lo && (int64_t)trow + G_TM <= g.row_hi && (int64_t)tcol >= g.col_lo &&
                          (int64_t)tcol + M3_TN <= g.col_hi;
    cplx* const Cw = g.C + row0 * g.ldc + col0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t r = row0 + 8 * i;
        const bool rv = interior || (r >= g.row_lo && r < g.row_hi);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int64_t c = col0 + 8 * j + e;
                cplx v = cmake(0.0, 0.0);
                if (rv && (interior || (c >= g.col_lo && c < g.col_hi))) v = Cw[(int64_t)(8 * i) * g.ldc + 8 * j + e];
                t1[i][j][e] = 0.0;
                t2[i][j][e] = v.x;
                t3[i][j][e] = v.x - v.y;
            }
        }
    }
    const int rx = lane >> 2, kq = 2 * (lane & 3);
    const int a_base = (wm * 32 + rx) * 128;
    const int b_base = wn * 2 * 1024;
    for (int it = 0; it < g.nks; ++it) {
        const int s = it % M3_STAGES;
        mbar_wait(&full[s], (it / M3_STAGES) & 1);
        const unsigned char* As = smA + s * M3_A_BYTES + a_base;
        const unsigned char* Bs = smB + s * M3_B_BYTES + b_base;
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            const int k = kq + p;  // this lane's complex k inside the stage
            const int asw = (k ^ rx) << 4, bsw = k * 128 + ((rx ^ k) << 4);
            double ar[4], ai[4], as[4], br[2], bi[2], bs[2];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const double2 v = *reinterpret_cast<const double2*>(As + i * 1024 + asw);
                ar[i] = v.x; ai[i] = v.y; as[i] = v.x + v.y;
            }
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const double2 v = *reinterpret_cast<const double2*>(Bs + j * 1024 + bsw);
                br[j] = v.x; bi[j] = v.y; bs[j] = v.x + v.y;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    dmma884(t1[i][j][0], t1[i][j][1], ar[i], br[j]);
                    dmma884(t2[i][j][0], t2[i][j][1], ai[i], bi[j]);
                    dmma884(t3[i][j][0], t3[i][j][1], as[i], bs[j]);
                }
        }
        __syncwarp();
        if (lane == 0) {
            mbar_arrive(&empty[s]);
            // refill the stage that everybody left one iteration ago
            if (it >= 1 && it - 1 + M3_STAGES < g.nks) {
                const int sp = (it - 1) % M3_STAGES;
                mbar_wait(&empty[sp], ((it - 1) / M3_STAGES) & 1);
                issue(it - 1 + M3_STAGES, sp, warp);
            }
        }
    }
    // C - AB = (Cr + T2 - T1) + i (T1 + T2 - (T3 - Cr + Ci) ...) with the load-time folding: (t2 - t1, t1 + t2 - t3)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t r = row0 + 8 * i;
        const bool rv = interior || (r >= g.row_lo && r < g.row_hi);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int64_t c = col0 + 8 * j + e;
                if (rv && (interior || (c >= g.col_lo && c < g.col_hi)))
                    Cw[(int64_t)(8 * i) * g.ldc + 8 * j + e] =
                        cmake(t2[i][j][e] - t1[i][j][e], (t1[i][j][e] + t2[i][j][e]) - t3[i][j][e]);
            }
        }
    }
}
static const size_t M3_SMEM = (size_t)M3_STAGES * (M3_A_BYTES + M3_B_BYTES) + 1024;


static int gemm_minb() {  // resident CTAs per SM the kernel is compiled for (measurement switch)
    static const int v = [] { const char* e = getenv("BHS_GEMM_MINB"); return e ? atoi(e) : 2; }();
    return v;
}
static void launch_zgemm_tma(dim3 grid, cudaStream_t st, const CUtensorMap& mL, const CUtensorMap& mU, const TmaGemmArgs& t) {
    static bool attr_set = false;
    static const bool prod = getenv("BHS_GEMM_BARRIER") == nullptr;  // default: barrier-free k loop (measured 3 % faster)
    if (!attr_set) {
        cudaFuncSetAttribute(zgemm_tma_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T_SMEM);
        cudaFuncSetAttribute(zgemm_tma_kernel<3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T_SMEM);
        cudaFuncSetAttribute(zgemm_tma_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T_SMEM);
        attr_set = true;
    }
    if (prod) zgemm_tma_kernel<2, true><<<grid, G_THREADS, T_SMEM, st>>>(mL, mU, t);
    else if (gemm_minb() == 3) zgemm_tma_kernel<3, false><<<grid, G_THREADS, T_SMEM, st>>>(mL, mU, t);
    else zgemm_tma_kernel<2, false><<<grid, G_THREADS, T_SMEM, st>>>(mL, mU, t);
}

// C[r_lo..r_hi, c_lo..c_hi) -= L[rows, k0 .. k0 + 8 nks) U[.., cols] on the tile grid anchored at (row_base, column 0): the 3M
// kernel (default) or the real-embedding ("4M") kernel (BHS_GEMM_4M=1, the A/B reference)
static void launch_gemm_window(cudaStream_t st, const CUtensorMap& mL, const CUtensorMap& mU, cplx* C, int64_t ldc, int64_t sC,
                               int64_t row_base, int64_t r_lo, int64_t r_hi, int64_t c_lo, int64_t c_hi, int k0, int nks,
                               int nbatch) {
    static const bool use4m = getenv("BHS_GEMM_4M") != nullptr;
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(zgemm3m_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)M3_SMEM);
        attr_set = true;
    }
    const int tn = use4m ? G_TN : M3_TN;
    TmaGemmArgs t;
    t.C = C; t.ldc = ldc; t.sC = sC;
    t.row_base = row_base; t.col_base = 0;
    t.rt0 = (int)((r_lo - row_base) / G_TM);
    t.ct0 = (int)(c_lo / tn);
    const int rt1 = (int)((r_hi - 1 - row_base) / G_TM), ct1 = (int)((c_hi - 1) / tn);
    t.row_lo = r_lo; t.row_hi = r_hi; t.col_lo = c_lo; t.col_hi = c_hi;
    t.k0 = k0; t.nks = nks;
    dim3 grid(ct1 - t.ct0 + 1, rt1 - t.rt0 + 1, nbatch);
    if (use4m) launch_zgemm_tma(grid, st, mL, mU, t);
    else zgemm3m_tma_kernel<<<grid, G_THREADS, M3_SMEM, st>>>(mL, mU, t);
}

// ---- tensor maps (driver entry point fetched through the runtime: no link-time dependency on libcuda) -----------------
typedef CUresult (*bhs_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static bhs_encode_tiled_fn get_encode_tiled() {
    static bhs_encode_tiled_fn fn = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            f = nullptr;
        return (bhs_encode_tiled_fn)f;
    }();
    return fn;
}
// View of `nbatch` row-major complex matrices [nrows, ncols] (leading dimension ld, `stride` complex elements apart) as a
// rank-3 tensor of doubles {2 ncols, nrows, nbatch}; box = {16 doubles, box_rows, 1}, 128-byte swizzle, zero fill outside.
static int make_operand_map(CUtensorMap* m, const void* base, int64_t nrows, int64_t ncols, int64_t ld, int nbatch,
                            int64_t stride, int box_rows) {
    bhs_encode_tiled_fn enc = get_encode_tiled();
    if (!enc) return BHS_ERR_UNSUPPORTED;
    if (((uintptr_t)base & 15) != 0) return BHS_ERR_INVALID;
    cuuint64_t dims[3] = {(cuuint64_t)(2 * ncols), (cuuint64_t)nrows, (cuuint64_t)nbatch};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 16, (cuuint64_t)(nbatch > 1 ? stride : ld * nrows) * 16};
    cuuint32_t box[3] = {16, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? BHS_OK : BHS_ERR_INVALID;
}


// ---- packing -------------------------------------------------------------------------------------------
// Lp tile (rt, stage): [64 rows][20] doubles, row r = interleaved complex A[row_base + 64 rt + r][k0 + 8 stage ..]
__global__ void pack_l_kernel(const cplx* __restrict__ A, int64_t ld, int64_t N, int64_t row_base, int64_t r_begin,
                              int64_t r_end_pad, int64_t k0, int K, int Kpad, double* __restrict__ Lp, int nks_total,
                              int ks_off, int64_t sA, int64_t sLp) {
    A += (int64_t)blockIdx.z * sA;
    Lp += (int64_t)blockIdx.z * sLp;
    // thread -> (row, kc); kc fastest
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t nrows = r_end_pad - r_begin;
    if (idx >= nrows * Kpad) return;
    int kc = (int)(idx % Kpad);
    int64_t r = r_begin + idx / Kpad;
    cplx v = cmake(0.0, 0.0);
    if (r < N && kc < K) v = A[r * ld + k0 + kc];
    int64_t rel = r - row_base;
    int64_t rt = rel / G_TM;
    int rr = (int)(rel % G_TM);
    double* dst = Lp + (((int64_t)rt * nks_total + ks_off + kc / G_KC) * G_TM + rr) * G_LDS + 2 * (kc % G_KC);
    *reinterpret_cast<double2*>(dst) = v;
}
// Up tile (ct, stage): [128 real cols][20] doubles: row 2j: (-br, +bi) pairs, row 2j+1: (-bi, -br)
__global__ void pack_u_kernel(const cplx* __restrict__ A, int64_t ld, int64_t ncols_total, int64_t k_row0, int K,
                              int Kpad, int64_t c_begin, int64_t c_end_pad, double* __restrict__ Up, int nks_total,
                              int ks_off, int64_t sA, int64_t sUp) {
    A += (int64_t)blockIdx.z * sA;
    Up += (int64_t)blockIdx.z * sUp;
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t ncols = c_end_pad - c_begin;
    if (idx >= ncols * Kpad) return;
    int64_t c = c_begin + idx % ncols;
    int p = (int)(idx / ncols);
    cplx v = cmake(0.0, 0.0);
    if (c < ncols_total && p < K) v = A[(k_row0 + p) * ld + c];
    int64_t ct = c / G_TN;
    int jj = (int)(c % G_TN);
    double* base = Up + (((int64_t)ct * nks_total + ks_off + p / G_KC) * (2 * G_TN)) * G_LDS + 2 * (p % G_KC);
    *reinterpret_cast<double2*>(base + (int64_t)(2 * jj) * G_LDS) = make_double2(-v.x, v.y);
    *reinterpret_cast<double2*>(base + (int64_t)(2 * jj + 1) * G_LDS) = make_double2(-v.y, -v.x);
}

// =====================================================================================================
// Tournament pivot selection
// =====================================================================================================
// One CTA eliminates up to LU_R candidate rows (one per thread, LU_NB complex in registers) with partial
// pivoting and emits the rows it pivoted on, in order.  rows_in == nullptr: contiguous rows
// [row_begin + 128*blockIdx.x, ...) ; otherwise rows_in[128*blockIdx.x + t] (-1 = empty slot).
//
// The LAST round of a panel (one CTA, fin.ipiv != nullptr) also finishes the panel's bookkeeping, so that no
// single-thread / single-CTA kernels sit on the critical path:
//   * the eliminated pivot rows ARE the factored diagonal block (multipliers left of the diagonal, U on and right of
//     it): they are written to fin.dblk [w][LU_NB] and copied into A by lu_permute_kernel;
//   * the ordered pivot list is converted to LAPACK-style sequential swaps ipiv[j + c];
//   * a zero pivot sets info = j + c + 1.
struct SelectFinal {
    int64_t j;
    int32_t* ipiv;
    int32_t* info;
    cplx* dblk;
    int32_t* pmap;  // this panel's slot of the row-map array
};

// 1 / p without the two divisions of Smith's algorithm when |p|^2 can neither overflow nor underflow
__device__ __forceinline__ cplx crecip_fast(cplx p) {
    const double s = fmax(fabs(p.x), fabs(p.y));
    if (s > 1e-140 && s < 1e140) {
        const double d = 1.0 / fma(p.x, p.x, p.y * p.y);
        return cmake(p.x * d, -p.y * d);
    }
    return crecip(p);
}

// Warp-wide maximum of a 64-bit key and the lowest lane that holds it.  One reduction on the high words settles it unless
// several lanes share the maximal high word (for magnitude keys: the top 20 mantissa bits agree -- rare), in which case the
// low words decide.
__device__ __forceinline__ unsigned long long warp_argmax_u64(unsigned long long key, int& lane_out) {
    const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
    const unsigned mhi = __reduce_max_sync(0xffffffffu, hi);
    const bool cand = hi == mhi;
    unsigned vote = __ballot_sync(0xffffffffu, cand);
    unsigned mlo;
    if (__popc(vote) > 1) {
        mlo = __reduce_max_sync(0xffffffffu, cand ? lo : 0u);
        vote = __ballot_sync(0xffffffffu, cand && lo == mlo);
        lane_out = __ffs(vote) - 1;
    } else {
        lane_out = __ffs(vote) - 1;
        mlo = __shfl_sync(0xffffffffu, lo, lane_out);
    }
    return ((unsigned long long)mhi << 32) | mlo;
}

// Bookkeeping of the last tournament round (one CTA, all threads call it after their last step): LAPACK-style
// sequential swaps ipiv[j + c] and the panel's net row map for lu_permute_kernel.
__device__ __forceinline__ void select_finish(const SelectFinal& fin, int w, const int32_t* s_win, int tid) {
    // Ordered pivot rows -> LAPACK-style sequential swaps ipiv[j + c], in O(w) steps: rows from below the diagonal
    // block are still at home when their turn comes; rows of the diagonal block may have been displaced by an earlier
    // swap, so their current position (where_top) and the row held by each diagonal-block position (cont_top) are
    // tracked.
    __shared__ int32_t where_top[LU_NB], cont_top[LU_NB];
    const int32_t j0 = (int32_t)fin.j;
    if (tid < LU_NB) { where_top[tid] = j0 + tid; cont_top[tid] = j0 + tid; }
    __syncthreads();
    if (tid == 0) {
        for (int c = 0; c < w; ++c) {
            const int32_t r = s_win[c];
            int32_t loc = j0 + c;  // r < 0 cannot happen for a square matrix; keep the row in place
            if (r >= 0) {
                loc = (r < j0 + w) ? where_top[r - j0] : r;
                const int32_t d = cont_top[c];  // always a row of the diagonal block
                where_top[d - j0] = loc;
                if (loc < j0 + w) cont_top[loc - j0] = d;
                if (r < j0 + w) where_top[r - j0] = j0 + c;
            }
            fin.ipiv[fin.j + c] = loc;
        }
    }
    __syncthreads();
    // net effect of those swaps, for lu_permute_kernel: position j + q receives row s_win[q]; a row j + t of the
    // diagonal block that was pushed out ends at where_top[t] >= j + w
    int32_t* pmap = fin.pmap;
    if (tid < LU_NB) {
        pmap[tid] = (tid < w && s_win[tid] >= 0) ? s_win[tid] : j0 + tid;
        pmap[LU_NB + tid] = (tid < w && where_top[tid] >= j0 + w) ? where_top[tid] : -1;
    }
}

// One elimination step costs ONE block barrier: every warp finds its own best row with warp-wide reductions (the
// magnitude's bit pattern is a monotone 64-bit key), that row's owners publish the row and its reciprocal pivot in the
// warp's slot of a double-buffered shared array, and after the barrier every thread picks the winning warp's slot.
//
// Layout: SEL_Q lanes per candidate row (lane g of the group holds the columns = g mod SEL_Q).  The step loop is ROLLED
// over groups of SEL_Q columns and the registers ROTATE by one local column per group, so that every register index is
// static: in group cb, x[i] holds the column SEL_Q * ((cb + i) mod (32 / SEL_Q)) + g, the pivot columns of the group are
// the x[0] of lanes g = 0 .. SEL_Q - 1 in turn, and finished columns (multipliers) travel round the back.  (The first
// version was fully unrolled: 257 KB of straight-line SASS run once per CTA, 54 % of its stall samples in "no
// instruction" under ncu.)  Candidate order, tie-breaking and every FMA are the same for every SEL_Q, so the pivots are too.
//
// SEL_Q = 4 (8 rows per warp, 16 warps) serves a lone system: a quarter of the dependent work per lane and four warps
// per scheduler; together with lu_permute_kernel it takes an N = 4096 factorisation from 43 ms to 34 ms (34 us instead of
// 41 us per tournament round; the chain of reductions, barrier and multiplier is what is left).  Systems factorised in
// lock step inside a sweep use lu_select_unrolled_kernel below.
template <int SEL_Q>
__global__ void __launch_bounds__(LU_R * SEL_Q)
    lu_select_kernel(const cplx* __restrict__ A, int64_t ld, int64_t col0, int w, const int32_t* __restrict__ rows_in,
                     int64_t n_in, int64_t row_begin, int64_t row_end, int32_t* __restrict__ rows_out, SelectFinal fin,
                     LuBatch bs) {
    A += (int64_t)blockIdx.z * bs.sA;
    if (rows_in) rows_in += (int64_t)blockIdx.z * bs.sCand;
    rows_out += (int64_t)blockIdx.z * bs.sCand;
    if (fin.ipiv) {
        fin.ipiv += (int64_t)blockIdx.z * bs.sIpiv;
        fin.info += blockIdx.z;
        fin.dblk += (int64_t)blockIdx.z * bs.sDblk;
        fin.pmap += (int64_t)blockIdx.z * bs.sPmap;
    }
    constexpr int SEL_LC = LU_NB / SEL_Q, SEL_THREADS = LU_R * SEL_Q;
    constexpr int NW = SEL_THREADS / 32;
    __shared__ cplx prow[2][NW][LU_NB];
    __shared__ cplx prinv[2][NW];
    __shared__ unsigned long long wkey[2][NW];
    __shared__ int wquad[2][NW];
    __shared__ int32_t s_win[LU_NB];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane & (SEL_Q - 1), quad = lane / SEL_Q;
    const bool is_final = fin.ipiv != nullptr;
    // which global row does this quad own?
    int64_t slot = (int64_t)blockIdx.x * LU_R + warp * (32 / SEL_Q) + quad;
    int32_t myrow = -1;
    if (rows_in) {
        if (slot < n_in) myrow = rows_in[slot];
    } else {
        int64_t r = row_begin + slot;
        if (r < row_end) myrow = (int32_t)r;
    }
    // The candidate row goes straight into registers (a quad reads 64 contiguous bytes per instruction).  No
    // shared-memory staging: a 67 KB tile would keep a second zgemm CTA of a concurrent system off this SM.
    cplx x[SEL_LC];
    {
        const cplx* src = A + (int64_t)(myrow >= 0 ? myrow : 0) * ld + col0 + g;
#pragma unroll
        for (int i = 0; i < SEL_LC; ++i) x[i] = (myrow >= 0 && SEL_Q * i + g < w) ? __ldg(src + SEL_Q * i) : cmake(0.0, 0.0);
    }
    bool active = myrow >= 0;
    if (tid < LU_NB && tid >= w) {  // columns beyond a narrow (tail) panel select nothing
        rows_out[(int64_t)blockIdx.x * LU_NB + tid] = -1;
        s_win[tid] = -1;
    }
    const int ncb = (w + SEL_Q - 1) / SEL_Q;
#pragma unroll 1
    for (int cb = 0; cb < ncb; ++cb) {
#pragma unroll
        for (int gg = 0; gg < SEL_Q; ++gg) {
            const int c = SEL_Q * cb + gg;
            if (c < w) {
                const int buf = c & 1;
                // key: 0 = no candidate, otherwise 1 + bits(|re| + |im|)  (monotone in the magnitude)
                const double mag = fabs(x[0].x) + fabs(x[0].y);
                const unsigned long long key =
                    (active && g == gg) ? (unsigned long long)__double_as_longlong(mag) + 1ULL : 0ULL;
                int bl;
                const unsigned long long wbest = warp_argmax_u64(key, bl);
                // this row's entry in the pivot column, for the whole quad (the multiplier needs it after the barrier)
                cplx xc;
                xc.x = __shfl_sync(0xffffffffu, x[0].x, (lane & ~(SEL_Q - 1)) | gg);
                xc.y = __shfl_sync(0xffffffffu, x[0].y, (lane & ~(SEL_Q - 1)) | gg);
                if (quad == bl / SEL_Q) {
                    if (lane == bl) {
                        wkey[buf][warp] = wbest;
                        wquad[buf][warp] = quad;
                        if (wbest) prinv[buf][warp] = wbest > 1ULL ? crecip_fast(x[0]) : cmake(0.0, 0.0);
                    }
                    if (wbest) {
#pragma unroll
                        for (int i = 0; i < SEL_LC; ++i) prow[buf][warp][SEL_Q * i + g] = x[i];  // rotated like x
                    }
                }
                __syncthreads();
                // best warp: the same three reductions over the per-warp keys (lanes q and q + NW hold warp q's key; the
                // lowest warp wins a tie, as the lowest lane does inside a warp)
                int bw;
                const unsigned long long best = warp_argmax_u64(wkey[buf][lane & (NW - 1)], bw);
                const bool any = best != 0ULL;       // at least one active row left
                const bool nonzero = best > 1ULL;    // its pivot entry is not exactly zero
                if (any && warp == bw && quad == wquad[buf][bw]) {
                    if (g == gg) {
                        rows_out[(int64_t)blockIdx.x * LU_NB + c] = myrow;
                        s_win[c] = myrow;
                        if (is_final && !nonzero) atomicCAS(fin.info, 0, (int)(fin.j + c + 1));
                    }
                    if (is_final) {
#pragma unroll
                        for (int i = 0; i < SEL_LC; ++i)
                            fin.dblk[c * LU_NB + SEL_Q * ((cb + i) & (SEL_LC - 1)) + g] = x[i];
                    }
                    active = false;
                }
                if (!any && tid == 0) { rows_out[(int64_t)blockIdx.x * LU_NB + c] = -1; s_win[c] = -1; }
                if (any && active && nonzero) {
                    const cplx l = cmul(xc, prinv[buf][bw]);
                    const cplx ml = cmake(-l.x, -l.y);
                    const cplx* pr = prow[buf][bw];
                    // the group's own columns: the pivot column receives the multiplier (part of the factored diagonal
                    // block if this row pivots later), the columns right of it are updated, those left of it are done
                    if (g == gg) x[0] = l;
                    else if (g > gg) x[0] = cfma(ml, pr[g], x[0]);
#pragma unroll
                    for (int i = 1; i < SEL_LC; ++i)
                        if (cb + i < SEL_LC) x[i] = cfma(ml, pr[SEL_Q * i + g], x[i]);
                }
            }
        }
        // rotate: the group's columns are finished and go to the back
        const cplx t = x[0];
#pragma unroll
        for (int i = 1; i < SEL_LC; ++i) x[i - 1] = x[i];
        x[SEL_LC - 1] = t;
    }
    if (is_final) select_finish(fin, w, s_win, tid);
}

// One lane per candidate row, every step unrolled with static register indices (no rotation, no predicated-off work:
// the fewest issued instructions, at the price of 257 KB of straight-line SASS).  Used for systems factorised in lock
// step inside a sweep, where the panels of many systems share the SMs with the tensor-core updates: measured on the
// same box, C3 sweep, 146.2 systems/s with this form against 145.2 with the rolled one-lane form and 140 with quads.
__global__ void __launch_bounds__(LU_R) lu_select_unrolled_kernel(const cplx* __restrict__ A, int64_t ld, int64_t col0, int w,
                                                         const int32_t* __restrict__ rows_in, int64_t n_in,
                                                         int64_t row_begin, int64_t row_end,
                                                         int32_t* __restrict__ rows_out, SelectFinal fin, LuBatch bs) {
    A += (int64_t)blockIdx.z * bs.sA;
    if (rows_in) rows_in += (int64_t)blockIdx.z * bs.sCand;
    rows_out += (int64_t)blockIdx.z * bs.sCand;
    if (fin.ipiv) {
        fin.ipiv += (int64_t)blockIdx.z * bs.sIpiv;
        fin.info += blockIdx.z;
        fin.dblk += (int64_t)blockIdx.z * bs.sDblk;
        fin.pmap += (int64_t)blockIdx.z * bs.sPmap;
    }
    constexpr int NW = LU_R / 32;
    __shared__ cplx prow[2][NW][LU_NB];
    __shared__ cplx prinv[2][NW];
    __shared__ unsigned long long wkey[2][NW];
    __shared__ int wlane[2][NW];
    __shared__ int32_t s_win[LU_NB];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool is_final = fin.ipiv != nullptr;
    // which global row does this thread own?
    int64_t slot = (int64_t)blockIdx.x * LU_R + tid;
    int32_t myrow = -1;
    if (rows_in) {
        if (slot < n_in) myrow = rows_in[slot];
    } else {
        int64_t r = row_begin + slot;
        if (r < row_end) myrow = (int32_t)r;
    }
    // Each thread reads its own candidate row straight into registers (512 contiguous bytes per thread).  No
    // shared-memory staging: a 67 KB tile would keep a second zgemm CTA of a concurrent system off this SM.
    cplx x[LU_NB];
    {
        const cplx* src = A + (int64_t)(myrow >= 0 ? myrow : 0) * ld + col0;
#pragma unroll
        for (int c = 0; c < LU_NB; ++c) x[c] = (myrow >= 0 && c < w) ? __ldg(src + c) : cmake(0.0, 0.0);
    }
    bool active = myrow >= 0;
#pragma unroll
    for (int c = 0; c < LU_NB; ++c) {
        if (c < w) {
            const int buf = c & 1;
            // key: 0 = no candidate, otherwise 1 + bits(|re| + |im|)  (monotone in the magnitude)
            const double mag = fabs(x[c].x) + fabs(x[c].y);
            const unsigned long long key = active ? (unsigned long long)__double_as_longlong(mag) + 1ULL : 0ULL;
            const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
            const unsigned mhi = __reduce_max_sync(0xffffffffu, hi);
            const bool cand = hi == mhi;
            const unsigned mlo = __reduce_max_sync(0xffffffffu, cand ? lo : 0u);
            const unsigned vote = __ballot_sync(0xffffffffu, cand && lo == mlo);
            const int bl = __ffs(vote) - 1;
            if (lane == bl) {
                wkey[buf][warp] = key;
                wlane[buf][warp] = bl;
                if (key) {
#pragma unroll
                    for (int j = c; j < LU_NB; ++j) prow[buf][warp][j] = x[j];
                    prinv[buf][warp] = key > 1ULL ? crecip_fast(x[c]) : cmake(0.0, 0.0);
                }
            }
            __syncthreads();
            int bw = 0;
            unsigned long long best = wkey[buf][0];
#pragma unroll
            for (int q = 1; q < NW; ++q) {
                const unsigned long long kq = wkey[buf][q];
                if (kq > best) { best = kq; bw = q; }
            }
            const bool any = best != 0ULL;       // at least one active row left
            const bool nonzero = best > 1ULL;    // its pivot entry is not exactly zero
            const int winner = bw * 32 + wlane[buf][bw];
            if (any && tid == winner) {
                rows_out[(int64_t)blockIdx.x * LU_NB + c] = myrow;
                s_win[c] = myrow;
                if (is_final) {
#pragma unroll
                    for (int j = 0; j < LU_NB; ++j) fin.dblk[c * LU_NB + j] = x[j];
                    if (!nonzero) atomicCAS(fin.info, 0, (int)(fin.j + c + 1));
                }
                active = false;
            }
            if (!any && tid == 0) { rows_out[(int64_t)blockIdx.x * LU_NB + c] = -1; s_win[c] = -1; }
            if (any && active && nonzero) {
                const cplx l = cmul(x[c], prinv[buf][bw]);
                x[c] = l;  // multiplier: becomes part of the factored diagonal block if this row pivots later
#pragma unroll
                for (int j = c + 1; j < LU_NB; ++j) x[j] = cfma(cmake(-l.x, -l.y), prow[buf][bw][j], x[j]);
            }
        } else {
            if (tid == 0) { rows_out[(int64_t)blockIdx.x * LU_NB + c] = -1; s_win[c] = -1; }
        }
    }
    if (is_final) select_finish(fin, w, s_win, tid);
}

// =====================================================================================================
// Cluster-resident panel: partial pivoting over up to 4096 rows in ONE launch
// =====================================================================================================
// A thread-block cluster of up to 16 CTAs holds 256 candidate rows each (two lanes per row, 16 columns per lane, in
// registers) and eliminates them with ordinary partial pivoting: per pivot column every CTA finds its best row (warp
// reductions + one block barrier, as in lu_select_kernel), pushes {key, 1/pivot, pivot row} into a slot of EVERY CTA of the
// cluster with st.async (distributed shared memory; the bytes complete a transaction barrier in the receiving CTA, so there is
// no cluster-wide barrier in the loop), and after its own barrier has collected the cluster's candidates each CTA picks the
// same winner and updates its rows.  Slots and barriers are double buffered by step parity: a CTA can only send step c + 2
// after it has received every CTA's step c + 1, which they send after their last read of step c.
// Measured cost of one exchange (tools/cluster_probe.cu, B200): 2 / 4 / 8 / 16 CTAs: ~1.0 / 1.2 / 1.5 / 2.4 k cycles.
//
// When the cluster holds ALL rows of the panel (M <= 256 x cluster size) this is the whole panel factorisation: the rows that
// were never chosen end up holding their multipliers, i.e. L21, which they write back themselves (no lu_l21_kernel, no
// tournament rounds: 1 launch instead of 5).  Larger panels (C5: 36 864 rows) first run tournament rounds of
// lu_select_kernel until at most 4096 candidates are left; the cluster then plays the final.
#define CP_ROWS 256
#define CP_Q 2
#define CP_THREADS (CP_ROWS * CP_Q)
#define CP_NW (CP_THREADS / 32)
#define CP_LC (LU_NB / CP_Q)
#define CP_MAXC 16
#define CP_PERM_COLS 64  // columns per tile of the fused row interchange (as lu_permute_kernel)
#define CP_SLOT_CHUNKS (2 + LU_NB)  // 16-byte chunks of a slot: {key, row}, 1/pivot, 32 row entries
struct __align__(16) CpSlot {
    unsigned long long key;
    int32_t row;
    int32_t pad;
    cplx rinv;
    cplx prow[LU_NB];
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void st_async_16(uint32_t dst, unsigned long long v0, unsigned long long v1, uint32_t remote_bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b64 [%0], {%1, %2}, [%3];" ::"r"(dst), "l"(v0),
                 "l"(v1), "r"(remote_bar)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

__global__ void __launch_bounds__(CP_THREADS, 1)
    lu_panel_cluster_kernel(cplx* __restrict__ A, int64_t ld, int64_t col0, int w, const int32_t* __restrict__ rows_in, int64_t n_in,
                            int64_t row_begin, int64_t row_end, SelectFinal fin, int write_l21, LuBatch bs, int64_t perm_c_lo,
                            int64_t perm_c_hi) {
    A += (int64_t)blockIdx.z * bs.sA;
    if (rows_in) rows_in += (int64_t)blockIdx.z * bs.sCand;
    fin.ipiv += (int64_t)blockIdx.z * bs.sIpiv;
    fin.info += blockIdx.z;
    fin.dblk += (int64_t)blockIdx.z * bs.sDblk;
    fin.pmap += (int64_t)blockIdx.z * bs.sPmap;
    // the published pivot rows and the exchange slots share one buffer with the row tile of the fused interchange (below)
    constexpr size_t CP_PROW_BYTES = sizeof(cplx) * 2 * CP_NW * LU_NB, CP_SLOT_BYTES = sizeof(CpSlot) * 2 * CP_MAXC;
    static_assert(CP_PROW_BYTES + CP_SLOT_BYTES >= sizeof(cplx) * LU_NB * CP_PERM_COLS, "row tile of the fused interchange");
    __shared__ __align__(16) unsigned char s_raw[CP_PROW_BYTES + CP_SLOT_BYTES];
    cplx(*prow)[CP_NW][LU_NB] = reinterpret_cast<cplx(*)[CP_NW][LU_NB]>(s_raw);
    CpSlot(*cslot)[CP_MAXC] = reinterpret_cast<CpSlot(*)[CP_MAXC]>(s_raw + CP_PROW_BYTES);
    __shared__ cplx prinv[2][CP_NW];
    __shared__ unsigned long long wkey[2][CP_NW];
    __shared__ int32_t wrow[2][CP_NW];
    __shared__ __align__(8) uint64_t cbar[2];
    __shared__ int32_t s_win[LU_NB];
    const uint32_t rank = cluster_ctarank(), csz = cluster_nctarank();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane & (CP_Q - 1), pair = lane / CP_Q;
    if (tid == 0) {
        mbar_init(&cbar[0], 1);
        mbar_init(&cbar[1], 1);
        mbar_fence_init();
    }
    if (tid < LU_NB) s_win[tid] = -1;
    // which global row does this pair of lanes own?
    const int64_t slot = (int64_t)rank * CP_ROWS + warp * (32 / CP_Q) + pair;
    int32_t myrow = -1;
    if (rows_in) {
        if (slot < n_in) myrow = rows_in[slot];
    } else {
        const int64_t r = row_begin + slot;
        if (r < row_end) myrow = (int32_t)r;
    }
    cplx x[CP_LC];
    {
        const cplx* src = A + (int64_t)(myrow >= 0 ? myrow : 0) * ld + col0 + g;
#pragma unroll
        for (int i = 0; i < CP_LC; ++i) x[i] = (myrow >= 0 && CP_Q * i + g < w) ? __ldg(src + CP_Q * i) : cmake(0.0, 0.0);
    }
    bool active = myrow >= 0;
    __syncthreads();
    if (csz > 1) cluster_sync_all();  // every CTA's barriers are initialised before anybody sends
    const int ncb = (w + CP_Q - 1) / CP_Q;
    // The step loop runs in four phases of four column groups.  In the phase that starts at group 4 p only the first
    // LIVE = CP_LC - 4 p register slots still hold unfinished columns (finished ones travel round the back of the rotation), so
    // that phase publishes, exchanges and updates LIVE slots instead of CP_LC: no predicated-off column work, and the
    // cluster exchange shrinks from 34 to 26 / 18 / 10 sixteen-byte chunks per peer.
    auto phase = [&](auto live_c, int cb_begin, int cb_end) {
        constexpr int LIVE = decltype(live_c)::value;
        constexpr int NCH = 2 + CP_Q * LIVE;  // chunks of a slot that are sent: {key, row}, 1/pivot, the live row entries
#pragma unroll 1
        for (int cb = cb_begin; cb < cb_end; ++cb) {
#pragma unroll
            for (int gg = 0; gg < CP_Q; ++gg) {
                const int c = CP_Q * cb + gg;
                if (c < w) {
                    const int buf = c & 1;
                    const double mag = fabs(x[0].x) + fabs(x[0].y);
                    const unsigned long long key =
                        (active && g == gg) ? (unsigned long long)__double_as_longlong(mag) + 1ULL : 0ULL;
                    int bl;
                    const unsigned long long wbest = warp_argmax_u64(key, bl);
                    cplx xc;  // this row's entry in the pivot column, for both lanes of the pair
                    xc.x = __shfl_sync(0xffffffffu, x[0].x, (lane & ~(CP_Q - 1)) | gg);
                    xc.y = __shfl_sync(0xffffffffu, x[0].y, (lane & ~(CP_Q - 1)) | gg);
                    if (pair == bl / CP_Q) {
                        if (lane == bl) {
                            wkey[buf][warp] = wbest;
                            wrow[buf][warp] = myrow;
                            if (wbest) prinv[buf][warp] = wbest > 1ULL ? crecip_fast(x[0]) : cmake(0.0, 0.0);
                        }
                        if (wbest) {
#pragma unroll
                            for (int i = 0; i < LIVE; ++i) prow[buf][warp][CP_Q * i + g] = x[i];  // rotated like x
                        }
                    }
                    __syncthreads();
                    int bw;
                    const unsigned long long ctabest = warp_argmax_u64(wkey[buf][lane & (CP_NW - 1)], bw);
                    unsigned long long best = ctabest;
                    const cplx* pr = prow[buf][bw];
                    cplx rinv = prinv[buf][bw];
                    int32_t winrow = wrow[buf][bw];
                    if (csz > 1) {
                        if (tid == 0) mbar_expect_tx(&cbar[buf], csz * (uint32_t)(NCH * 16));
                        const uint32_t my_slot = smem_u32(&cslot[buf][rank]), my_bar = smem_u32(&cbar[buf]);
                        for (int t = tid; t < NCH * (int)csz; t += CP_THREADS) {
                            const uint32_t peer = t / NCH, ch = t % NCH;
                            unsigned long long v0, v1;
                            if (ch == 0) {
                                v0 = ctabest;
                                v1 = (unsigned long long)(uint32_t)winrow;
                            } else {
                                const cplx e = ch == 1 ? rinv : pr[ch - 2];
                                v0 = (unsigned long long)__double_as_longlong(e.x);
                                v1 = (unsigned long long)__double_as_longlong(e.y);
                            }
                            st_async_16(mapa_shared(my_slot + ch * 16, peer), v0, v1, mapa_shared(my_bar, peer));
                        }
                        while (!mbar_try_wait(&cbar[buf], (uint32_t)(c >> 1) & 1u)) {
                        }
                        int bc;
                        best = warp_argmax_u64(lane < (int)csz ? cslot[buf][lane].key : 0ULL, bc);  // lowest rank wins a tie
                        pr = cslot[buf][bc].prow;
                        rinv = cslot[buf][bc].rinv;
                        winrow = cslot[buf][bc].row;
                    }
                    const bool any = best != 0ULL;     // at least one active row left
                    const bool nonzero = best > 1ULL;  // its pivot entry is not exactly zero
                    if (tid == 0) {
                        s_win[c] = any ? winrow : -1;
                        if (rank == 0 && any && !nonzero) atomicCAS(fin.info, 0, (int)(fin.j + c + 1));
                    }
                    if (any && active && myrow == winrow) {  // the pivot row IS row c of the factored diagonal block
#pragma unroll
                        for (int i = 0; i < CP_LC; ++i) fin.dblk[c * LU_NB + CP_Q * ((cb + i) & (CP_LC - 1)) + g] = x[i];
                        active = false;
                    }
                    if (any && active && nonzero) {
                        const cplx l = cmul(xc, rinv);
                        const cplx ml = cmake(-l.x, -l.y);
                        if (g == gg) x[0] = l;
                        else if (g > gg) x[0] = cfma(ml, pr[g], x[0]);
#pragma unroll
                        for (int i = 1; i < LIVE; ++i)
                            if (cb + i < CP_LC) x[i] = cfma(ml, pr[CP_Q * i + g], x[i]);
                    }
                }
            }
            const cplx t = x[0];
#pragma unroll
            for (int i = 1; i < CP_LC; ++i) x[i - 1] = x[i];
            x[CP_LC - 1] = t;
        }
    };
    static_assert(CP_LC == 16, "the four phases below assume 16 register slots per lane");
    phase(std::integral_constant<int, 16>{}, 0, ncb < 4 ? ncb : 4);
    phase(std::integral_constant<int, 12>{}, 4, ncb < 8 ? ncb : 8);
    phase(std::integral_constant<int, 8>{}, 8, ncb < 12 ? ncb : 12);
    phase(std::integral_constant<int, 4>{}, 12, ncb);
    // rows that never pivoted hold their multipliers: that is L21 (only when every row of the panel was a candidate)
    if (write_l21 && active) {
        cplx* dst = A + (int64_t)myrow * ld + col0;
#pragma unroll
        for (int i = 0; i < CP_LC; ++i) {
            const int col = CP_Q * ((ncb + i) & (CP_LC - 1)) + g;
            if (col < w) dst[col] = x[i];
        }
    }
    if (csz > 1) cluster_sync_all();  // nobody leaves while a peer could still be writing into its slots
    if (rank == 0) {
        __syncthreads();
        select_finish(fin, w, s_win, tid);
    }
    // Fused row interchange (look-ahead path: the columns [perm_c_lo, perm_c_hi) of the current outer block): what
    // lu_permute_kernel would do in a launch of its own right after this one.  The panel's net row map (rank 0, select_finish)
    // and every CTA's L21 / diagonal-block stores are made visible by one more cluster barrier; CTA r then moves the rows of the
    // column tiles r, r + csz, ...  All global loads of a tile are independent (the 32 rows of the diagonal block go through
    // shared memory).
    if (perm_c_hi > perm_c_lo) {
        __threadfence();
        if (csz > 1) cluster_sync_all();
        else __syncthreads();
        cplx(*tile)[CP_PERM_COLS] = reinterpret_cast<cplx(*)[CP_PERM_COLS]>(s_raw);
        __shared__ int32_t s_map[2 * LU_NB];
        const int ntile = (int)((perm_c_hi - perm_c_lo + CP_PERM_COLS - 1) / CP_PERM_COLS);
        const int64_t j = fin.j;
        const int col = tid & (CP_PERM_COLS - 1), q0 = (tid / CP_PERM_COLS) * (LU_NB / (CP_THREADS / CP_PERM_COLS));
        constexpr int RPT = LU_NB / (CP_THREADS / CP_PERM_COLS);  // rows per thread: 512 threads = 64 columns x 8 row groups
        for (int tl = (int)rank; tl < ntile; tl += (int)csz) {
            const int64_t c = perm_c_lo + (int64_t)tl * CP_PERM_COLS + col;
            const bool valid = c < perm_c_hi;
            __syncthreads();  // the previous tile (or the exchange buffers) are no longer read
            if (tid < 2 * LU_NB) s_map[tid] = __ldcg(fin.pmap + tid);
            if (valid) {
#pragma unroll
                for (int i = 0; i < RPT; ++i)
                    if (q0 + i < w) tile[q0 + i][col] = __ldcg(A + (j + q0 + i) * ld + c);
            }
            __syncthreads();
            cplx v[RPT];
            if (valid) {
#pragma unroll
                for (int i = 0; i < RPT; ++i) {
                    const int q = q0 + i;
                    if (q < w) {
                        const int64_t src = s_map[q];
                        v[i] = (src < j + w) ? tile[src - j][col] : __ldcg(A + src * ld + c);
                    }
                }
            }
            __syncthreads();  // the rows read above are overwritten below, possibly by another thread of the column
            if (valid) {
                const bool in_panel = c >= j && c < j + w;
#pragma unroll
                for (int i = 0; i < RPT; ++i) {
                    const int q = q0 + i;
                    if (q < w) {
                        A[(j + q) * ld + c] = in_panel ? __ldcg(fin.dblk + q * LU_NB + (c - j)) : v[i];
                        const int64_t dst = s_map[LU_NB + q];
                        if (dst >= 0) A[dst * ld + c] = tile[q][col];
                    }
                }
            }
        }
    }
}

// Apply the row interchanges of `npan` consecutive panels (the first at column j0, 32 wide except possibly the last) to the
// matrix columns [c_lo, c_lo + ncols) (and, in the same launch, to the right-hand sides) from the panels' NET row maps: all
// the rows of a diagonal block go through a shared-memory tile, so that every global load is independent (the LAPACK-style
// sequence of 32 swaps is a chain of 32 dependent round trips per column).  64 columns per CTA, 4 threads per column.
// A panel's own columns [j, j+w) of rows [j, j+w) receive the factored diagonal block produced by the pivot-selection kernel
// (dblk != nullptr: only when one panel is applied right after its selection).  With look-ahead the panels of a block column
// are applied to the columns left and right of it later, several at once.
#define PERM_COLS 64
__global__ void __launch_bounds__(256) lu_permute_kernel(cplx* __restrict__ M, int64_t ld, int64_t c_lo, int64_t ncols, int nct,
                                                         cplx* __restrict__ rhs, int nrhs, int64_t j0, int npan, int wlast,
                                                         const cplx* __restrict__ dblk, const int32_t* __restrict__ pmap0,
                                                         int64_t sM, int64_t sRhs, int64_t sDblk, int64_t sPmap) {
    __shared__ cplx tile[LU_NB][PERM_COLS];
    __shared__ int32_t s_map[2 * LU_NB];
    if (dblk) dblk += (int64_t)blockIdx.z * sDblk;
    pmap0 += (int64_t)blockIdx.z * sPmap;
    int64_t c0 = c_lo + (int64_t)blockIdx.x * PERM_COLS;
    int64_t c_end = c_lo + ncols;
    if ((int)blockIdx.x >= nct) {  // right-hand-side columns
        M = rhs + (int64_t)blockIdx.z * sRhs;
        ld = nrhs;
        c_end = nrhs;
        c0 = (int64_t)((int)blockIdx.x - nct) * PERM_COLS;
        dblk = nullptr;
    } else {
        M += (int64_t)blockIdx.z * sM;
    }
    const int tid = threadIdx.x, col = tid & (PERM_COLS - 1), q0 = (tid / PERM_COLS) * (LU_NB / 4);
    const int64_t c = c0 + col;
    const bool valid = c < c_end;
    for (int p = 0; p < npan; ++p) {
        const int64_t j = j0 + (int64_t)p * LU_NB;
        const int w = (p == npan - 1) ? wlast : LU_NB;
        if (p) __syncthreads();  // the previous panel's maps and tile have been consumed
        if (tid < 2 * LU_NB) s_map[tid] = pmap0[p * LU_PMAP + tid];
        if (valid) {
#pragma unroll
            for (int i = 0; i < LU_NB / 4; ++i)
                if (q0 + i < w) tile[q0 + i][col] = M[(j + q0 + i) * ld + c];
        }
        __syncthreads();
        cplx v[LU_NB / 4];
        if (valid) {
#pragma unroll
            for (int i = 0; i < LU_NB / 4; ++i) {
                const int q = q0 + i;
                if (q < w) {
                    const int64_t src = s_map[q];
                    v[i] = (src < j + w) ? tile[src - j][col] : M[src * ld + c];
                }
            }
        }
        __syncthreads();  // the rows read above are overwritten below, possibly by another thread of the column
        if (valid) {
            const bool in_panel = dblk != nullptr && c >= j && c < j + w;
#pragma unroll
            for (int i = 0; i < LU_NB / 4; ++i) {
                const int q = q0 + i;
                if (q < w) {
                    M[(j + q) * ld + c] = in_panel ? dblk[q * LU_NB + (c - j)] : v[i];
                    const int64_t dst = s_map[LU_NB + q];
                    if (dst >= 0) M[dst * ld + c] = tile[q][col];
                }
            }
        }
    }
}

// L21 = A21 U11^{-1}: one row per thread (registers), U11 in shared memory.  Also emits the packed
// GEMM image of the new L columns (rows below the diagonal block).
__global__ void __launch_bounds__(LU_R) lu_l21_kernel(cplx* __restrict__ A, int64_t ld, int64_t N, int64_t j, int w,
                                                      int64_t sA) {
    A += (int64_t)blockIdx.z * sA;
    __shared__ cplx U[LU_NB][LU_NB + 1];
    __shared__ cplx rdiag[LU_NB];
    const int tid = threadIdx.x;
    const int64_t r0 = j + w + (int64_t)blockIdx.x * LU_R;
    for (int e = tid; e < LU_NB * LU_NB; e += LU_R) {
        int r = e / LU_NB, c = e % LU_NB;
        U[r][c] = (r < w && c < w && c >= r) ? A[(j + r) * ld + j + c] : cmake(0.0, 0.0);
    }
    const int64_t myr = r0 + tid;
    cplx* rowp = A + (myr < N ? myr : 0) * ld + j;
    cplx x[LU_NB];
#pragma unroll
    for (int c = 0; c < LU_NB; ++c) x[c] = (myr < N && c < w) ? rowp[c] : cmake(0.0, 0.0);
    __syncthreads();
    if (tid < LU_NB) {
        cplx d = U[tid][tid];
        rdiag[tid] = (tid < w && (d.x != 0.0 || d.y != 0.0)) ? crecip(d) : cmake(0.0, 0.0);
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < LU_NB; ++c) {
        // x[c] = (a[c] - sum_{q<c} x[q] U[q][c]) / U[c][c]
        cplx s = x[c];
#pragma unroll
        for (int q = 0; q < c; ++q) s = cfma(cmake(-x[q].x, -x[q].y), U[q][c], s);
        x[c] = cmul(s, rdiag[c]);
    }
    if (myr < N) {
#pragma unroll
        for (int c = 0; c < LU_NB; ++c)
            if (c < w) rowp[c] = x[c];
    }
}

// U12 = L11^{-1} A12 for a 32-row block: one column per thread (registers), L11 (unit lower) in smem.
__global__ void __launch_bounds__(128) lu_trsm32_kernel(cplx* __restrict__ A, int64_t ld, int64_t j, int w,
                                                        cplx* __restrict__ X, int64_t ldx, int64_t c_begin,
                                                        int64_t c_end, int64_t sA) {
    A += (int64_t)blockIdx.z * sA;
    X += (int64_t)blockIdx.z * sA;
    __shared__ cplx Ls[LU_NB][LU_NB + 1];
    const int tid = threadIdx.x;
    for (int e = tid; e < LU_NB * LU_NB; e += 128) {
        int r = e / LU_NB, c = e % LU_NB;
        Ls[r][c] = (r < w && c < r) ? A[(j + r) * ld + j + c] : cmake(0.0, 0.0);
    }
    __syncthreads();
    int64_t c = c_begin + (int64_t)blockIdx.x * 128 + tid;
    if (c >= c_end) return;
    cplx x[LU_NB];
#pragma unroll
    for (int r = 0; r < LU_NB; ++r) x[r] = (r < w) ? X[(j + r) * ldx + c] : cmake(0.0, 0.0);
#pragma unroll
    for (int r = 1; r < LU_NB; ++r) {
        cplx s = x[r];
#pragma unroll
        for (int q = 0; q < r; ++q) s = cfma(cmake(-Ls[r][q].x, -Ls[r][q].y), x[q], s);
        x[r] = s;
    }
#pragma unroll
    for (int r = 0; r < LU_NB; ++r)
        if (r < w) X[(j + r) * ldx + c] = x[r];
}

// =====================================================================================================
// right-hand-side kernels (few columns): y -= L * x style updates and triangular block solves
// =====================================================================================================
// rhs[r, :] -= sum_{q<K} M[r, k0+q] * rhs[k0+q, :]   for r in [r_begin, r_end): one warp per row
__global__ void __launch_bounds__(256) rhs_gemv_sub_kernel(const cplx* __restrict__ M, int64_t ld, int64_t r_begin,
                                                           int64_t r_end, int64_t k0, int K, cplx* __restrict__ rhs,
                                                           int nrhs, int64_t sA, int64_t sRhs) {
    M += (int64_t)blockIdx.z * sA;
    rhs += (int64_t)blockIdx.z * sRhs;
    const int lane = threadIdx.x & 31;
    int64_t r = r_begin + (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= r_end) return;
    for (int c = 0; c < nrhs; ++c) {
        double sr = 0.0, si = 0.0;
        for (int q = lane; q < K; q += 32) {
            cplx m = M[r * ld + k0 + q], v = rhs[(k0 + q) * nrhs + c];
            sr += m.x * v.x - m.y * v.y;
            si += m.x * v.y + m.y * v.x;
        }
        for (int o = 16; o > 0; o >>= 1) {
            sr += __shfl_xor_sync(0xffffffffu, sr, o);
            si += __shfl_xor_sync(0xffffffffu, si, o);
        }
        if (lane == 0) {
            cplx v = rhs[r * nrhs + c];
            rhs[r * nrhs + c] = cmake(v.x - sr, v.y - si);
        }
    }
}
// Solve the T x T (T <= 128) diagonal block at k0 against rhs rows [k0, k0+T): lower-unit (forward) or upper (backward).
// The block is consumed in 32-column strips (rows at / below the strip's diagonal sub-block for the lower solve, at / above
// it for the upper one).  Each strip is staged in shared memory row by row (512 contiguous bytes per row, pitch 33 so that
// both access patterns below are conflict-free); the loads of the NEXT strip are issued into registers before the current
// one is used, so that their latency hides behind the arithmetic.  Per strip: the warp that owns the 32 x 32 diagonal
// sub-block solves it with one row per lane (the solved entries travel by warp shuffle, no block barrier inside the 32
// steps), then every other row subtracts its 32-term dot product.  Two block barriers per strip and column instead of two
// per matrix column (54 us -> ~8 us for a 128 x 128 block).
#define RS_W 32
#define RS_PITCH (RS_W + 1)
#define RS_NC 4  // right-hand-side columns handled per pass over the strips
__global__ void __launch_bounds__(128) rhs_block_solve_kernel(const cplx* __restrict__ M, int64_t ld, int64_t k0, int T, int upper,
                                                              cplx* __restrict__ rhs, int nrhs, int64_t sA, int64_t sRhs) {
    M += (int64_t)blockIdx.z * sA;
    rhs += (int64_t)blockIdx.z * sRhs;
    extern __shared__ __align__(16) cplx s_strip[];  // [LU_NBO][RS_PITCH]
    __shared__ cplx v[RS_NC][LU_NBO];
    __shared__ cplx rdiag[LU_NBO];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nsub = (T + RS_W - 1) / RS_W;
    M += k0 * ld + k0;
    if (upper) {
        cplx dg = cmake(1.0, 0.0);
        if (tid < T) dg = M[(int64_t)tid * ld + tid];
        rdiag[tid] = (dg.x != 0.0 || dg.y != 0.0) ? crecip(dg) : cmake(1.0, 0.0);
    }
    // strip `si` (in processing order): sub-block index sb, rows [r_lo, r_hi) of the block, columns [32 sb, 32 sb + 32)
    auto strip_rows = [&](int si, int& sb, int& r_lo, int& r_hi) {
        sb = upper ? nsub - 1 - si : si;
        r_lo = upper ? 0 : RS_W * sb;
        r_hi = upper ? min(T, RS_W * sb + RS_W) : T;
    };
    cplx pre[RS_W];  // this thread's share of the next strip: elements e = tid + 128 i  ->  (row r_lo + e / 32, column e % 32)
    auto prefetch = [&](int si) {
        int sb, r_lo, r_hi;
        strip_rows(si, sb, r_lo, r_hi);
        const int n = (r_hi - r_lo) * RS_W;
#pragma unroll
        for (int i = 0; i < RS_W; ++i) {
            const int e = tid + 128 * i;
            const int r = r_lo + e / RS_W, cc = RS_W * sb + (e % RS_W);
            pre[i] = (e < n && cc < T) ? M[(int64_t)r * ld + cc] : cmake(0.0, 0.0);
        }
    };
    for (int c0 = 0; c0 < nrhs; c0 += RS_NC) {
        const int nc = min(RS_NC, nrhs - c0);
        __syncthreads();
        for (int c = 0; c < nc; ++c) v[c][tid] = tid < T ? rhs[(k0 + tid) * nrhs + c0 + c] : cmake(0.0, 0.0);
        prefetch(0);
        for (int si = 0; si < nsub; ++si) {
            int sb, r_lo, r_hi;
            strip_rows(si, sb, r_lo, r_hi);
            __syncthreads();  // the previous strip has been consumed
            {
                const int n = (r_hi - r_lo) * RS_W;
#pragma unroll
                for (int i = 0; i < RS_W; ++i) {
                    const int e = tid + 128 * i;
                    if (e < n) s_strip[(e / RS_W) * RS_PITCH + (e % RS_W)] = pre[i];
                }
            }
            if (si + 1 < nsub) prefetch(si + 1);
            __syncthreads();
            const int d0 = RS_W * sb;  // first row / column of the diagonal sub-block
            for (int c = 0; c < nc; ++c) {
                if (warp == sb) {
                    // 32 x 32 triangular solve, row d0 + lane per lane
                    // (rows beyond T of a tail block hold x = 0 and never touch the -- unstaged -- strip rows)
                    const bool valid = d0 + lane < T;
                    const cplx* row = s_strip + (valid ? d0 + lane - r_lo : 0) * RS_PITCH;
                    cplx x = v[c][d0 + lane];
                    if (upper) {
#pragma unroll
                        for (int q = RS_W - 1; q >= 0; --q) {
                            if (lane == q) x = cmul(x, rdiag[d0 + q]);
                            cplx xq;
                            xq.x = __shfl_sync(0xffffffffu, x.x, q);
                            xq.y = __shfl_sync(0xffffffffu, x.y, q);
                            const cplx m = row[q];
                            if (lane < q && valid) x = cfma(cmake(-m.x, -m.y), xq, x);
                        }
                    } else {
#pragma unroll
                        for (int q = 0; q < RS_W; ++q) {
                            cplx xq;
                            xq.x = __shfl_sync(0xffffffffu, x.x, q);
                            xq.y = __shfl_sync(0xffffffffu, x.y, q);
                            const cplx m = row[q];
                            if (lane > q && valid) x = cfma(cmake(-m.x, -m.y), xq, x);
                        }
                    }
                    v[c][d0 + lane] = x;
                }
                __syncthreads();
                // the other rows of the strip: v[r] -= sum_q strip[r][q] x[q]
                const bool mine = upper ? tid < d0 : (tid >= d0 + RS_W && tid < T);
                if (mine) {
                    const cplx* row = s_strip + (tid - r_lo) * RS_PITCH;
                    cplx a0 = v[c][tid], a1 = cmake(0.0, 0.0);
#pragma unroll
                    for (int q = 0; q < RS_W; q += 2) {
                        const cplx m0 = row[q], m1 = row[q + 1];
                        a0 = cfma(cmake(-m0.x, -m0.y), v[c][d0 + q], a0);
                        a1 = cfma(cmake(-m1.x, -m1.y), v[c][d0 + q + 1], a1);
                    }
                    v[c][tid] = cadd(a0, a1);
                }
                // (rows written here are read by the next strip's diagonal solve only after the barrier at its top)
            }
        }
        __syncthreads();
        for (int c = 0; c < nc; ++c)
            if (tid < T) rhs[(k0 + tid) * nrhs + c0 + c] = v[c][tid];
    }
}
static const size_t RS_SMEM = (size_t)LU_NBO * RS_PITCH * sizeof(cplx);
// permutation from sequential swaps, applied to a few columns (used by the stand-alone zgetrs)
__global__ void rhs_apply_ipiv_kernel(const int32_t* __restrict__ ipiv, int64_t N, cplx* __restrict__ rhs, int nrhs) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;  // one thread per right-hand-side column, any number of columns
    if (c >= nrhs) return;
    for (int64_t i = 0; i < N; ++i) {
        int64_t p = ipiv[i];
        if (p != i) {
            cplx a = rhs[i * nrhs + c], b = rhs[p * nrhs + c];
            rhs[i * nrhs + c] = b;
            rhs[p * nrhs + c] = a;
        }
    }
}

// =====================================================================================================
// host driver
// =====================================================================================================
struct LuCtx {
    cplx* A;
    int64_t ld, N;
    cplx* rhs;  // may be null
    int nrhs;
    int32_t* ipiv;
    int32_t* info;
    int32_t* cand[2];
    cplx* dblk;  // [LU_NB][LU_NB] factored diagonal block handed from the pivot-selection kernel to lu_permute_kernel
    int32_t* pmaps;  // [ceil(N / 32)][LU_PMAP] net row maps of every panel
    bool lookahead;  // panels of block J+1 on the caller's stream while the trailing update of block J runs on a side stream
    double* Lp;
    double* Up;
    int nks_total;      // LU_NBO / G_KC
    int64_t J;          // current outer block start (anchor of Lp rows and of the stage index)
    int nbo;            // outer block width of this factorisation: LU_NBO, or a multiple of it (see lu_factor)
    bool defer_perm;    // panels interchange rows inside their outer block only; the other columns follow once per block
    int nbatch;         // systems factorised in lock step (grid z)
    LuBatch bs;
    cudaStream_t st;
    int err;
    bool tma;           // operands by tensor-map TMA straight from A (default); false = packed operand images (BHS_GEMM_PACKED=1)
    CUtensorMap mapL, mapU;  // A viewed with the L-operand box {8 complex, 64 rows} and the U-operand box {8 complex, 8 rows}
};

static bool gemm_packed_requested() {
    static const bool v = getenv("BHS_GEMM_PACKED") != nullptr;
    return v;
}

struct LuWork {
    int32_t* cand0;
    int32_t* cand1;
    cplx* dblk;
    int32_t* pmaps;
    double* Lp;
    double* Up;
    int64_t ncand, sLp, sUp, sPmap;
    int64_t bytes;
};
static LuWork lu_carve(int64_t N, int nbatch, void* base) {
    LuWork w;
    unsigned char* c = (unsigned char*)base;
    int64_t off = 0;
    auto take = [&](int64_t b) { unsigned char* r = c + off; off += al256(b); return r; };
    w.ncand = cdiv64(N, LU_R) * LU_NB + LU_NB;
    w.cand0 = (int32_t*)take(w.ncand * 4 * nbatch);
    w.cand1 = (int32_t*)take(w.ncand * 4 * nbatch);
    w.dblk = (cplx*)take((int64_t)LU_DBLK * sizeof(cplx) * nbatch);
    w.sPmap = cdiv64(N, LU_NB) * LU_PMAP;
    w.pmaps = (int32_t*)take(w.sPmap * 4 * nbatch);
    int64_t rtiles = cdiv64(N, G_TM) + 1, ctiles = cdiv64(N, G_TN) + 1;
    int nks = LU_NBO / G_KC;
    w.sLp = rtiles * nks * G_A_STAGE;
    w.sUp = ctiles * nks * G_B_STAGE;
    w.Lp = (double*)take(w.sLp * 8 * nbatch);
    w.Up = (double*)take(w.sUp * 8 * nbatch);
    w.bytes = off;
    return w;
}

#define LU_LAUNCH_CHECK(ctx)                                  \
    do {                                                      \
        BHS_COUNT_LAUNCH();                                   \
        cudaError_t e__ = cudaGetLastError();                 \
        if (e__ != cudaSuccess && !(ctx).err) (ctx).err = (int)e__; \
    } while (0)

// BHS_LU_SKIP (measurement aid, WRONG results): bit mask of kernel families that are not launched, to see what each costs a
// sweep -- 1: pivot selection + interchanges + L21 (the whole panel), 2: interchanges only, 4: 32-row triangular solves,
// 8: updates with K < 128, 16: right-hand-side kernels, 32: the K = 128 update inside a 256-wide block, 64: the K = 128 update
// inside its U12 solve, 128: the trailing updates (K = outer block width).
static int lu_skip_mask() {
    static const int m = [] { const char* e = getenv("BHS_LU_SKIP"); return e ? atoi(e) : 0; }();
    return m;
}

// C[rows r_lo.., cols c_lo..c_hi) -= L[rows, k0..k0+K) * U[k0..k0+K, cols]   (operands already packed)
static void lu_pack_l(LuCtx& x, int64_t r_lo, int64_t r_hi, int64_t k0, int K);
static void lu_gemm(LuCtx& x, int64_t r_lo, int64_t r_hi, int64_t c_lo, int64_t c_hi, int64_t k0, int K) {
    if (r_lo >= r_hi || c_lo >= c_hi || K <= 0) return;
    if ((lu_skip_mask() & 8) && K < LU_NBO) return;
    if ((lu_skip_mask() & 128) && K > LU_NBO) return;
    const int pcat = (K >= LU_NBO || (k0 == x.J && r_lo >= x.J + LU_NBO)) ? BHS_PROF_LU_GEMM : BHS_PROF_LU_GEMM_IN;
    if (x.tma) {
        bhs_prof_begin(pcat, x.st);
        launch_gemm_window(x.st, x.mapL, x.mapU, x.A, x.ld, x.bs.sA, x.J, r_lo, r_hi, c_lo, c_hi, (int)k0, (K + G_KC - 1) / G_KC,
                           x.nbatch);
        bhs_prof_end(pcat, 8.0 * (double)(r_hi - r_lo) * (double)(c_hi - c_lo) * (double)K * x.nbatch, x.st);
        LU_LAUNCH_CHECK(x);
        return;
    }
    // the L operand is packed here, from the current A: pivoting of later panels of the same outer block
    // permutes rows of earlier L columns, so an image packed at panel time would be stale
    lu_pack_l(x, r_lo, r_hi, k0, K);
    GemmArgs g;
    g.Lp = x.Lp; g.Up = x.Up; g.nks_total = x.nks_total;
    g.ks0 = (int)((k0 - x.J) / G_KC);
    g.nks = (K + G_KC - 1) / G_KC;
    g.C = x.A; g.ldc = x.ld;
    g.row_base = x.J; g.col_base = 0;
    g.rt0 = (int)((r_lo - x.J) / G_TM);
    g.ct0 = (int)(c_lo / G_TN);
    int rt1 = (int)((r_hi - 1 - x.J) / G_TM), ct1 = (int)((c_hi - 1) / G_TN);
    g.row_lo = r_lo; g.row_hi = r_hi; g.col_lo = c_lo; g.col_hi = c_hi;
    g.sC = x.bs.sA; g.sLp = x.bs.sLp; g.sUp = x.bs.sUp;
    dim3 grid(ct1 - g.ct0 + 1, rt1 - g.rt0 + 1, x.nbatch);
    bhs_prof_begin(pcat, x.st);
    zgemm_sub_kernel<<<grid, G_THREADS, G_SMEM, x.st>>>(g);
    bhs_prof_end(pcat, 8.0 * (double)(r_hi - r_lo) * (double)(c_hi - c_lo) * (double)K * x.nbatch, x.st);
    LU_LAUNCH_CHECK(x);
}
static void lu_pack_l(LuCtx& x, int64_t r_lo, int64_t r_hi, int64_t k0, int K) {
    // rows [r_lo, r_hi) padded up to the tile grid anchored at J (rows >= N are zero filled)
    if (r_lo >= r_hi) return;
    int64_t r_end_pad = x.J + cdiv64(r_hi - x.J, G_TM) * G_TM;
    int Kpad = ((K + G_KC - 1) / G_KC) * G_KC;
    int64_t tot = (r_end_pad - r_lo) * Kpad;
    bhs_prof_begin(BHS_PROF_LU_PACK, x.st);
    pack_l_kernel<<<dim3((unsigned)cdiv64(tot, 256), 1, x.nbatch), 256, 0, x.st>>>(
        x.A, x.ld, x.N, x.J, r_lo, r_end_pad, k0, K, Kpad, x.Lp, x.nks_total, (int)((k0 - x.J) / G_KC), x.bs.sA, x.bs.sLp);
    bhs_prof_end(BHS_PROF_LU_PACK, 0.0, x.st);
    LU_LAUNCH_CHECK(x);
}
static void lu_pack_u(LuCtx& x, int64_t k0, int K, int64_t c_lo, int64_t c_hi) {
    if (c_lo >= c_hi || x.tma) return;
    int64_t c_begin = (c_lo / G_TN) * G_TN, c_end_pad = cdiv64(c_hi, G_TN) * G_TN;
    int Kpad = ((K + G_KC - 1) / G_KC) * G_KC;
    int64_t tot = (c_end_pad - c_begin) * Kpad;
    bhs_prof_begin(BHS_PROF_LU_PACK, x.st);
    pack_u_kernel<<<dim3((unsigned)cdiv64(tot, 256), 1, x.nbatch), 256, 0, x.st>>>(
        x.A, x.ld, x.N, k0, K, Kpad, c_begin, c_end_pad, x.Up, x.nks_total, (int)((k0 - x.J) / G_KC), x.bs.sA, x.bs.sUp);
    bhs_prof_end(BHS_PROF_LU_PACK, 0.0, x.st);
    LU_LAUNCH_CHECK(x);
}

// Largest cluster the device can schedule for lu_panel_cluster_kernel (16 needs the non-portable attribute), 0 = none.
// BHS_LU_CLUSTER=0 disables the cluster panel, BHS_LU_CLUSTER=2 also uses it for systems factorised in lock step.
static int g_cluster_mode = -1, g_cluster_max = 0;
static void cluster_init() {
    if (g_cluster_mode >= 0) return;
    const char* e = getenv("BHS_LU_CLUSTER");
    g_cluster_mode = e ? atoi(e) : 1;
    g_cluster_max = 0;
    if (g_cluster_mode == 0) return;
    cudaFuncSetAttribute(lu_panel_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    for (int csz = CP_MAXC; csz >= 2 && !g_cluster_max; csz >>= 1) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(csz, 1, 1);
        cfg.blockDim = dim3(CP_THREADS, 1, 1);
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = csz; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, lu_panel_cluster_kernel, &cfg) == cudaSuccess && n > 0) g_cluster_max = csz;
    }
    cudaGetLastError();
    if (!g_cluster_max) g_cluster_max = 1;  // a "cluster" of one CTA still factors up to 256 rows in one launch
}

// 32-wide (or narrower) panel at column j: pivot selection, swap, diagonal LU, L21
static void lu_panel(LuCtx& x, int64_t j, int w) {
    if (lu_skip_mask() & 1) return;
    const int64_t M = x.N - j;
    int64_t nsets = cdiv64(M, LU_R);
    int cur = 0;
    bhs_prof_begin(BHS_PROF_LU_PANEL, x.st);
    int32_t* const pmap_j = x.pmaps + (j / LU_NB) * LU_PMAP;
    const SelectFinal fin{j, x.ipiv, x.info, x.dblk, pmap_j}, nofin{0, nullptr, nullptr, nullptr, nullptr};
    const bool quad = x.nbatch == 1;  // 4 lanes per candidate row for a lone system (latency); unrolled one-lane form in a sweep
    auto select_round = [&](const int32_t* rows_in, int64_t n_in, int64_t ns, int32_t* rows_out, const SelectFinal& f) {
        if (quad)
            lu_select_kernel<4><<<dim3((unsigned)ns, 1, x.nbatch), LU_R * 4, 0, x.st>>>(x.A, x.ld, j, w, rows_in, n_in, j, x.N, rows_out, f, x.bs);
        else
            lu_select_unrolled_kernel<<<dim3((unsigned)ns, 1, x.nbatch), LU_R, 0, x.st>>>(x.A, x.ld, j, w, rows_in, n_in, j, x.N, rows_out, f, x.bs);
        LU_LAUNCH_CHECK(x);
    };
    cluster_init();
    bool l21_done = false, done = false, perm_fused = false;
    // Small groups (the sweep engine forms groups of 2 for short sweeps, e.g. one rank's share at 8 GPUs) also take the cluster
    // panel once at most 1024 rows are left: in a short sweep every group reaches its latency-bound last panels at the same
    // time, with nothing left to overlap them, and the one-launch panel shortens exactly that tail (BHS_LU_TAIL=0: off).
    static const int tail_rows = [] { const char* e = getenv("BHS_LU_TAIL"); return e ? atoi(e) : 1024; }();
    const bool tail = x.nbatch <= 2 && M <= tail_rows;
    const bool use_cluster = g_cluster_mode && (x.nbatch == 1 || g_cluster_mode >= 2 || tail) && M > LU_R;
    if (use_cluster) {
        // tournament rounds until one cluster can hold the candidates, then partial pivoting inside the cluster
        const int64_t cap = (int64_t)g_cluster_max * CP_ROWS;
        const int32_t* rows_in = nullptr;
        int64_t n_in = M;
        while (n_in > cap) {
            const int64_t ns = cdiv64(n_in, LU_R);
            select_round(rows_in, n_in, ns, x.cand[cur], nofin);
            rows_in = x.cand[cur];
            cur ^= 1;
            n_in = ns * LU_NB;
        }
        const int csz = (int)cdiv64(n_in, CP_ROWS);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(csz, 1, x.nbatch);
        cfg.blockDim = dim3(CP_THREADS, 1, 1);
        cfg.stream = x.st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = csz; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        const int direct = rows_in == nullptr;
        // with look-ahead the interchange of the block's own columns is fused into the panel kernel (one launch less per panel
        // on the chain that bounds a lone system; BHS_LU_FUSE_PERM=0 keeps the separate launch)
        static const bool fuse_off = [] { const char* e = getenv("BHS_LU_FUSE_PERM"); return e && atoi(e) == 0; }();
        int64_t pc_lo = 0, pc_hi = 0;
        if (x.lookahead && !fuse_off && !(lu_skip_mask() & 2)) {
            pc_lo = x.J;
            pc_hi = x.J + x.nbo < x.N ? x.J + x.nbo : x.N;
        }
        cudaError_t e = cudaLaunchKernelEx(&cfg, lu_panel_cluster_kernel, x.A, x.ld, j, w, rows_in, n_in, j, x.N, fin, direct, x.bs,
                                           pc_lo, pc_hi);
        if (e == cudaSuccess) {
            BHS_COUNT_LAUNCH();
            done = true;
            perm_fused = pc_hi > pc_lo;
            l21_done = direct;
        } else {
            cudaGetLastError();   // this device cannot place the cluster after all: tournament from here on
            g_cluster_mode = 0;
            nsets = cdiv64(M, LU_R);
            cur = 0;
        }
    }
    if (!done) {
        select_round(nullptr, 0, nsets, x.cand[cur], nsets == 1 ? fin : nofin);
        while (nsets > 1) {
            int64_t n_in = nsets * LU_NB;
            int64_t nsets2 = cdiv64(n_in, LU_R);
            select_round(x.cand[cur], n_in, nsets2, x.cand[cur ^ 1], nsets2 == 1 ? fin : nofin);
            cur ^= 1;
            nsets = nsets2;
        }
    }
    if (!(lu_skip_mask() & 2) && !perm_fused) {
        // row interchanges: every column (and the right-hand sides) now -- or, with look-ahead, only the columns of the outer
        // block being factorised; the columns left and right of it follow on the side stream (lu_permute_deferred)
        const bool local = x.lookahead || x.defer_perm;
        const int64_t c_lo = local ? x.J : 0;
        const int64_t c_hi = local ? (x.J + x.nbo < x.N ? x.J + x.nbo : x.N) : x.N;
        // (with look-ahead the right-hand sides are interchanged and solved on the update stream, off the panel chain)
        const int nct = (int)cdiv64(c_hi - c_lo, PERM_COLS), nrt = (x.rhs && !x.lookahead) ? (int)cdiv64(x.nrhs, PERM_COLS) : 0;
        lu_permute_kernel<<<dim3((unsigned)(nct + nrt), 1, x.nbatch), 256, 0, x.st>>>(
            x.A, x.ld, c_lo, c_hi - c_lo, nct, x.rhs, x.nrhs, j, 1, w, x.dblk, pmap_j, x.bs.sA, x.bs.sRhs, x.bs.sDblk, x.bs.sPmap);
        LU_LAUNCH_CHECK(x);
    }
    if (j + w < x.N && !l21_done) {
        lu_l21_kernel<<<dim3((unsigned)cdiv64(x.N - j - w, LU_R), 1, x.nbatch), LU_R, 0, x.st>>>(x.A, x.ld, x.N, j, w, x.bs.sA);
        LU_LAUNCH_CHECK(x);
    }
    bhs_prof_end(BHS_PROF_LU_PANEL, 0.0, x.st);
}

// U[j0..j0+T, cols) = L11^{-1} A[j0..j0+T, cols)   (T multiple of 32 except possibly the tail), packs U
static void lu_trsm(LuCtx& x, int64_t j0, int T, int64_t c_lo, int64_t c_hi) {
    if (c_lo >= c_hi || T <= 0) return;
    if (T <= LU_NB) {
        if (lu_skip_mask() & 4) return;
        bhs_prof_begin(BHS_PROF_LU_TRSM, x.st);
        lu_trsm32_kernel<<<dim3((unsigned)cdiv64(c_hi - c_lo, 128), 1, x.nbatch), 128, 0, x.st>>>(x.A, x.ld, j0, T, x.A, x.ld, c_lo,
                                                                                                  c_hi, x.bs.sA);
        bhs_prof_end(BHS_PROF_LU_TRSM, 0.0, x.st);
        LU_LAUNCH_CHECK(x);
        lu_pack_u(x, j0, T, c_lo, c_hi);
        return;
    }
    int h = (T > 128) ? 128 : (T > 64) ? 64 : 32;
    lu_trsm(x, j0, h, c_lo, c_hi);
    if (!((lu_skip_mask() & 64) && h >= LU_NBO)) lu_gemm(x, j0 + h, j0 + T, c_lo, c_hi, j0, h);
    lu_trsm(x, j0 + h, T - h, c_lo, c_hi);
}

// recursive LU of columns [j0, j0+w) (rows j0..N), w <= the outer block width
static void lu_rec(LuCtx& x, int64_t j0, int w) {
    if (w <= LU_NB) {
        lu_panel(x, j0, w);
        return;
    }
    int h = (w > 128) ? 128 : (w > 64) ? 64 : 32;
    lu_rec(x, j0, h);
    lu_trsm(x, j0, h, j0 + h, j0 + w);
    if (!((lu_skip_mask() & 32) && h >= LU_NBO)) lu_gemm(x, j0 + h, x.N, j0 + h, j0 + w, j0, h);
    lu_rec(x, j0 + h, w - h);
}

// The interchanges of the panels of the outer block at J, applied to the columns [c_lo, c_hi) that the panel-time launches
// skipped (look-ahead).
static void lu_permute_deferred(LuCtx& x, int64_t J, int w, int64_t c_lo, int64_t c_hi) {
    if (c_lo >= c_hi || (lu_skip_mask() & 3)) return;
    const int npan = (w + LU_NB - 1) / LU_NB, wlast = w - (npan - 1) * LU_NB;
    const int nct = (int)cdiv64(c_hi - c_lo, PERM_COLS);
    bhs_prof_begin(BHS_PROF_LU_PANEL, x.st);
    lu_permute_kernel<<<dim3((unsigned)nct, 1, x.nbatch), 256, 0, x.st>>>(x.A, x.ld, c_lo, c_hi - c_lo, nct, nullptr, 0, J, npan, wlast,
                                                                          nullptr, x.pmaps + (J / LU_NB) * LU_PMAP, x.bs.sA, 0,
                                                                          x.bs.sDblk, x.bs.sPmap);
    bhs_prof_end(BHS_PROF_LU_PANEL, 0.0, x.st);
    LU_LAUNCH_CHECK(x);
}

// The interchanges of the panels of the outer block at J applied to the right-hand sides only (look-ahead: the panel-time
// launches skip them).
static void lu_permute_rhs(LuCtx& x, int64_t J, int w) {
    if (!x.rhs || (lu_skip_mask() & 3)) return;
    const int npan = (w + LU_NB - 1) / LU_NB, wlast = w - (npan - 1) * LU_NB;
    const int nrt = (int)cdiv64(x.nrhs, PERM_COLS);
    bhs_prof_begin(BHS_PROF_LU_RHS, x.st);
    lu_permute_kernel<<<dim3((unsigned)nrt, 1, x.nbatch), 256, 0, x.st>>>(x.A, x.ld, 0, 0, 0, x.rhs, x.nrhs, J, npan, wlast, nullptr,
                                                                          x.pmaps + (J / LU_NB) * LU_PMAP, x.bs.sA, x.bs.sRhs,
                                                                          x.bs.sDblk, x.bs.sPmap);
    bhs_prof_end(BHS_PROF_LU_RHS, 0.0, x.st);
    LU_LAUNCH_CHECK(x);
}

static void lu_rhs_forward(LuCtx& x, int64_t J, int w) {
    if (lu_skip_mask() & 16) return;
    // forward substitution of this block row: y_J = L11^{-1} rhs_J ; rhs_below -= L21 y_J
    bhs_prof_begin(BHS_PROF_LU_RHS, x.st);
    rhs_block_solve_kernel<<<dim3(1, 1, x.nbatch), 128, RS_SMEM, x.st>>>(x.A, x.ld, J, w, 0, x.rhs, x.nrhs, x.bs.sA, x.bs.sRhs);
    LU_LAUNCH_CHECK(x);
    if (J + w < x.N) {
        rhs_gemv_sub_kernel<<<dim3((unsigned)cdiv64(x.N - J - w, 8), 1, x.nbatch), 256, 0, x.st>>>(
            x.A, x.ld, J + w, x.N, J, w, x.rhs, x.nrhs, x.bs.sA, x.bs.sRhs);
        LU_LAUNCH_CHECK(x);
    }
    bhs_prof_end(BHS_PROF_LU_RHS, 0.0, x.st);
}

// Panel stream of the look-ahead (one per host thread and device; forked from / joined to the caller's stream with events, so
// the whole factorisation can still be captured into a CUDA graph).  It has the HIGHEST priority: the trailing update on the
// caller's stream fills every SM with tensor-core CTAs, and the latency-bound panel kernels (a 16-CTA cluster that needs whole
// SMs) must get the slots those free first, or they would simply queue behind the update.
static cudaStream_t lu_panel_stream() {
    thread_local cudaStream_t pool[64] = {nullptr};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    if (!pool[dev]) {
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);  // hi = numerically smallest = greatest priority
        if (cudaStreamCreateWithPriority(&pool[dev], cudaStreamNonBlocking, hi) != cudaSuccess) pool[dev] = nullptr;
    }
    return pool[dev];
}

// Look-ahead (a lone system: nothing else keeps the device busy during the latency-bound panels).  The panel stream
// factorises the outer block columns one after the other -- panels, the updates inside the block, the forward substitution
// of the right-hand sides -- and touches nothing outside the current block column.  The caller's stream does everything else
// of block J once its panels are done: the deferred row interchanges left and right of the block column, the U12 solve, the
// update of the NEXT block column (after which the panel stream may start on it) and then the rest of the trailing matrix,
// which thus overlaps the panels of block J + 1.
static int lu_factor_lookahead(LuCtx& x, cudaStream_t pst) {
    const cudaStream_t ust = x.st;
    cudaEvent_t e_panel = nullptr, e_next = nullptr;
    if (cudaEventCreateWithFlags(&e_panel, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&e_next, cudaEventDisableTiming) != cudaSuccess)
        return BHS_ERR_ALLOC;
    cudaEventRecord(e_next, ust);  // fork: the panel stream starts behind the work queued ahead of this call
    cudaStreamWaitEvent(pst, e_next, 0);
    int64_t last_J = 0;
    int last_w = 0;
    for (int64_t J = 0; J < x.N; J += LU_NBO) {
        const int w = (int)((x.N - J < LU_NBO) ? (x.N - J) : LU_NBO);
        x.J = J;
        x.st = pst;
        lu_rec(x, J, w);
        last_J = J;
        last_w = w;
        cudaEventRecord(e_panel, pst);
        x.st = ust;
        cudaStreamWaitEvent(ust, e_panel, 0);  // (after the last block this is the join)
        if (J + w < x.N) {
            lu_permute_deferred(x, J, w, J + w, x.N);
            lu_trsm(x, J, w, J + w, x.N);
            const int64_t nc_hi = (J + w + LU_NBO < x.N) ? J + w + LU_NBO : x.N;
            lu_gemm(x, J + w, x.N, J + w, nc_hi, J, w);  // the next block column first
            cudaEventRecord(e_next, ust);
            cudaStreamWaitEvent(pst, e_next, 0);
            // everything below overlaps the panels of the next block: the interchanges on the columns LEFT of this block, the
            // right-hand sides of this block row (interchanges, forward substitution) and the rest of the trailing update
            lu_permute_deferred(x, J, w, 0, J);
            if (x.rhs) {
                lu_permute_rhs(x, J, w);
                lu_rhs_forward(x, J, w);
            }
            lu_gemm(x, J + w, x.N, nc_hi, x.N, J, w);
        } else if (x.rhs) {
            lu_permute_rhs(x, J, w);
            lu_rhs_forward(x, J, w);
        }
    }
    // the interchanges of the last block's panels on the columns left of it
    x.st = ust;
    lu_permute_deferred(x, last_J, last_w, 0, last_J);
    cudaEventDestroy(e_panel);
    cudaEventDestroy(e_next);
    return x.err;
}

static int lu_factor(LuCtx& x) {
    cudaMemsetAsync(x.info, 0, sizeof(int32_t) * x.nbatch, x.st);
    cudaFuncSetAttribute(zgemm_sub_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G_SMEM);
    cudaFuncSetAttribute(rhs_block_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RS_SMEM);
    // BHS_LU_GEMM_ONLY=1 (measurement aid, wrong results): issue only the trailing updates, to see how much of a
    // sweep's time the DMMA kernel accounts for on its own
    static const bool gemm_only = getenv("BHS_LU_GEMM_ONLY") != nullptr;
    // BHS_LU_LOOKAHEAD=0 disables the two-stream look-ahead of lone systems, =2 also uses it for systems in lock step and for
    // any size.  By default it serves 256 < N <= 12288: the panels cost O(N^2) and the update O(N^3), so beyond that there is
    // little left to hide, while the high-priority panel kernels (clusters that need whole SMs) disturb the update -- measured
    // at N = 36 864: 4.46 s without, 4.81 s with look-ahead; at N = 8192: 86 ms without, 69 ms with.
    static const int la_mode = [] { const char* e = getenv("BHS_LU_LOOKAHEAD"); return e ? atoi(e) : 1; }();
    x.lookahead = false;
    x.defer_perm = false;
    if (!gemm_only && x.tma && la_mode && (la_mode >= 2 || (x.nbatch == 1 && x.N <= 12288)) && x.N > 2 * LU_NBO) {
        cudaStream_t pst = lu_panel_stream();
        if (pst) {
            x.lookahead = true;
            x.nbo = LU_NBO;
            return lu_factor_lookahead(x, pst);
        }
    }
    // Outer block width without look-ahead: 256 with the tensor-map GEMM.  The trailing update then runs 256-deep k loops (half
    // as many passes over C, a longer steady state per tile), at the price of a 128-deep update inside the block; the panel chain
    // gets longer, which only matters for a lone system (those take the look-ahead path above, 128 wide).  BHS_LU_NBO overrides.
    const int nbo_env = [] { const char* e = getenv("BHS_LU_NBO"); return e ? atoi(e) : 0; }();  // (read per call: tests vary it)
    x.nbo = x.tma ? 256 : LU_NBO;
    const bool defer_off = getenv("BHS_LU_PERM_NOW") != nullptr;  // A/B: every panel interchanges all N columns at once
    x.defer_perm = !gemm_only && !defer_off;
    if (nbo_env >= LU_NBO && nbo_env % LU_NBO == 0 && x.tma) x.nbo = nbo_env;
    for (int64_t J = 0; J < x.N; J += x.nbo) {
        int w = (int)((x.N - J < x.nbo) ? (x.N - J) : x.nbo);
        x.J = J;
        if (gemm_only) {
            if (J + w < x.N) {
                lu_pack_u(x, J, w, J + w, x.N);
                lu_gemm(x, J + w, x.N, J + w, x.N, J, w);
            }
            continue;
        }
        lu_rec(x, J, w);
        if (x.defer_perm) {  // the interchanges of the block's panels on the columns left and right of it, all panels per launch
            lu_permute_deferred(x, J, w, 0, J);
            lu_permute_deferred(x, J, w, J + w, x.N);
        }
        if (x.rhs)  // (the block solve works on 128-wide blocks)
            for (int o = 0; o < w; o += LU_NBO) lu_rhs_forward(x, J + o, (w - o < LU_NBO) ? w - o : LU_NBO);
        if (J + w < x.N) {
            lu_trsm(x, J, w, J + w, x.N);
            lu_gemm(x, J + w, x.N, J + w, x.N, J, w);
        }
    }
    return x.err;
}

static int lu_backward(LuCtx& x) {
    if (lu_skip_mask() & 16) return x.err;
    // x = U^{-1} y, block rows from the bottom
    int64_t nblk = cdiv64(x.N, LU_NBO);
    bhs_prof_begin(BHS_PROF_LU_RHS, x.st);
    for (int64_t bi = nblk - 1; bi >= 0; --bi) {
        int64_t J = bi * LU_NBO;
        int w = (int)((x.N - J < LU_NBO) ? (x.N - J) : LU_NBO);
        rhs_block_solve_kernel<<<dim3(1, 1, x.nbatch), 128, RS_SMEM, x.st>>>(x.A, x.ld, J, w, 1, x.rhs, x.nrhs, x.bs.sA, x.bs.sRhs);
        LU_LAUNCH_CHECK(x);
        if (J > 0) {
            rhs_gemv_sub_kernel<<<dim3((unsigned)cdiv64(J, 8), 1, x.nbatch), 256, 0, x.st>>>(x.A, x.ld, 0, J, J, w, x.rhs, x.nrhs,
                                                                                         x.bs.sA, x.bs.sRhs);
            LU_LAUNCH_CHECK(x);
        }
    }
    bhs_prof_end(BHS_PROF_LU_RHS, 0.0, x.st);
    return x.err;
}

extern "C" int64_t bhs_zgesv_workspace(int64_t N, int nrhs) {
    if (N <= 0 || nrhs < 0) return BHS_ERR_INVALID;
    return lu_carve(N, 1, nullptr).bytes;
}
extern "C" int64_t bhs_zgesv_batched_workspace(int64_t N, int nrhs, int nbatch) {
    if (N <= 0 || nrhs < 0 || nbatch <= 0) return BHS_ERR_INVALID;
    return lu_carve(N, nbatch, nullptr).bytes;
}

static int lu_setup(LuCtx& x, int64_t N, double* d_A, int64_t ld, double* d_rhs, int nrhs, int32_t* d_ipiv,
                    int32_t* d_info, void* d_work, void* stream, int nbatch = 1, int64_t strideA = 0,
                    int64_t stride_rhs = 0) {
    if (N <= 0 || !d_A || ld < N || !d_ipiv || !d_info || !d_work || nbatch <= 0 || nbatch > 65535) return BHS_ERR_INVALID;
    if (N > 2000000000LL / LU_NB) return BHS_ERR_UNSUPPORTED;
    LuWork w = lu_carve(N, nbatch, d_work);
    x.nbatch = nbatch;
    x.bs.sA = strideA; x.bs.sRhs = stride_rhs; x.bs.sIpiv = N; x.bs.sCand = w.ncand;
    x.bs.sDblk = (int64_t)LU_DBLK; x.bs.sPmap = w.sPmap; x.bs.sLp = w.sLp; x.bs.sUp = w.sUp;
    x.pmaps = w.pmaps;
    x.A = (cplx*)d_A; x.ld = ld; x.N = N; x.rhs = (cplx*)d_rhs; x.nrhs = nrhs;
    x.ipiv = d_ipiv; x.info = d_info; x.cand[0] = w.cand0; x.cand[1] = w.cand1; x.dblk = w.dblk;
    x.Lp = w.Lp; x.Up = w.Up; x.nks_total = LU_NBO / G_KC; x.J = 0;
    x.st = (cudaStream_t)stream; x.err = 0;
    x.tma = !gemm_packed_requested();
    if (x.tma) {
        int rc = make_operand_map(&x.mapL, d_A, N, N, ld, nbatch, strideA, G_TM);
        if (rc == BHS_OK) rc = make_operand_map(&x.mapU, d_A, N, N, ld, nbatch, strideA, G_KC);
        if (rc != BHS_OK) return rc;
    }
    return BHS_OK;
}

extern "C" int bhs_zgetrf(int64_t N, double* d_A, int64_t ld, int32_t* d_ipiv, int32_t* d_info, void* d_work,
                          void* stream) {
    LuCtx x;
    int rc = lu_setup(x, N, d_A, ld, nullptr, 0, d_ipiv, d_info, d_work, stream);
    if (rc) return rc;
    return lu_factor(x);
}

extern "C" int bhs_zgesv(int64_t N, int nrhs, double* d_A, int64_t ld, double* d_rhs, int32_t* d_ipiv,
                         int32_t* d_info, void* d_work, void* stream) {
    if (nrhs <= 0 || !d_rhs) return BHS_ERR_INVALID;
    LuCtx x;
    int rc = lu_setup(x, N, d_A, ld, d_rhs, nrhs, d_ipiv, d_info, d_work, stream);
    if (rc) return rc;
    rc = lu_factor(x);
    if (rc) return rc;
    return lu_backward(x);
}

// nbatch systems of one size in lock step: A [nbatch][N, ld] (strideA complex elements apart), rhs [nbatch][N, nrhs]
// (stride_rhs apart), ipiv [nbatch][N], info [nbatch]; workspace bhs_zgesv_batched_workspace.
extern "C" int bhs_zgesv_batched(int64_t N, int nrhs, int nbatch, double* d_A, int64_t ld, int64_t strideA, double* d_rhs,
                                 int64_t stride_rhs, int32_t* d_ipiv, int32_t* d_info, void* d_work, void* stream) {
    if (nrhs <= 0 || !d_rhs || strideA < N * ld - (ld - N) || stride_rhs < N * nrhs) return BHS_ERR_INVALID;
    LuCtx x;
    int rc = lu_setup(x, N, d_A, ld, d_rhs, nrhs, d_ipiv, d_info, d_work, stream, nbatch, strideA, stride_rhs);
    if (rc) return rc;
    rc = lu_factor(x);
    if (rc) return rc;
    return lu_backward(x);
}

extern "C" int bhs_zgetrs(int64_t N, int nrhs, const double* d_LU, int64_t ld, const int32_t* d_ipiv, double* d_rhs,
                          void* d_work, void* stream) {
    if (N <= 0 || nrhs <= 0 || !d_LU || ld < N || !d_ipiv || !d_rhs) return BHS_ERR_INVALID;
    (void)d_work;
    LuCtx x;
    x.A = (cplx*)d_LU; x.ld = ld; x.N = N; x.rhs = (cplx*)d_rhs; x.nrhs = nrhs;
    x.nbatch = 1;
    x.bs = LuBatch{0, 0, 0, 0, 0, 0, 0};
    x.st = (cudaStream_t)stream; x.err = 0;
    cudaFuncSetAttribute(rhs_block_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RS_SMEM);
    rhs_apply_ipiv_kernel<<<(unsigned)cdiv64(nrhs, 32), 32, 0, x.st>>>(d_ipiv, N, x.rhs, nrhs);
    LU_LAUNCH_CHECK(x);
    for (int64_t J = 0; J < N; J += LU_NBO) {
        int w = (int)((N - J < LU_NBO) ? (N - J) : LU_NBO);
        rhs_block_solve_kernel<<<1, 128, RS_SMEM, x.st>>>(x.A, ld, J, w, 0, x.rhs, nrhs, 0, 0);
        LU_LAUNCH_CHECK(x);
        if (J + w < N) {
            rhs_gemv_sub_kernel<<<(unsigned)cdiv64(N - J - w, 8), 256, 0, x.st>>>(x.A, ld, J + w, N, J, w, x.rhs, nrhs, 0, 0);
            LU_LAUNCH_CHECK(x);
        }
    }
    return lu_backward(x);
}

// ---- stand-alone C -= A*B (row-major operands), for tests and the roofline measurement -------------------
extern "C" int64_t bhs_zgemm_workspace(int64_t M, int64_t N, int64_t K) {
    if (M <= 0 || N <= 0 || K <= 0) return BHS_ERR_INVALID;
    int64_t nks = cdiv64(K, G_KC);
    return al256((cdiv64(M, G_TM)) * nks * G_A_STAGE * 8) + al256((cdiv64(N, G_TN)) * nks * G_B_STAGE * 8);
}

extern "C" int bhs_zgemm_sub(int64_t M, int64_t N, int64_t K, const double* d_A, int64_t lda, const double* d_B,
                             int64_t ldb, double* d_C, int64_t ldc, void* d_work, void* stream) {
    if (M <= 0 || N <= 0 || K <= 0 || !d_A || !d_B || !d_C || !d_work || lda < K || ldb < N || ldc < N)
        return BHS_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    int64_t nks = cdiv64(K, G_KC);
    if (!gemm_packed_requested()) {
        // operands by TMA straight from the caller's matrices (columns of A / rows of B beyond K are zero filled)
        CUtensorMap mL, mU;
        int rc = make_operand_map(&mL, d_A, M, K, lda, 1, 0, G_TM);
        if (rc == BHS_OK) rc = make_operand_map(&mU, d_B, K, N, ldb, 1, 0, G_KC);
        if (rc != BHS_OK) return rc;
        launch_gemm_window(st, mL, mU, (cplx*)d_C, ldc, 0, 0, 0, M, 0, N, 0, (int)nks, 1);
        BHS_CHECK_LAUNCH();
        return BHS_OK;
    }
    int Kpad = (int)(nks * G_KC);
    double* Lp = (double*)d_work;
    double* Up = (double*)((unsigned char*)d_work + al256(cdiv64(M, G_TM) * nks * G_A_STAGE * 8));
    int64_t r_end_pad = cdiv64(M, G_TM) * G_TM, c_end_pad = cdiv64(N, G_TN) * G_TN;
    pack_l_kernel<<<(unsigned)cdiv64(r_end_pad * Kpad, 256), 256, 0, st>>>((const cplx*)d_A, lda, M, 0, 0, r_end_pad, 0, (int)K,
                                                                        Kpad, Lp, (int)nks, 0, 0, 0);
    BHS_CHECK_LAUNCH();
    pack_u_kernel<<<(unsigned)cdiv64(c_end_pad * Kpad, 256), 256, 0, st>>>((const cplx*)d_B, ldb, N, 0, (int)K, Kpad, 0,
                                                                        c_end_pad, Up, (int)nks, 0, 0, 0);
    BHS_CHECK_LAUNCH();
    cudaFuncSetAttribute(zgemm_sub_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G_SMEM);
    GemmArgs g;
    g.Lp = Lp; g.Up = Up; g.nks_total = (int)nks; g.ks0 = 0; g.nks = (int)nks;
    g.C = (cplx*)d_C; g.ldc = ldc; g.row_base = 0; g.col_base = 0; g.rt0 = 0; g.ct0 = 0;
    g.row_lo = 0; g.row_hi = M; g.col_lo = 0; g.col_hi = N;
    g.sC = 0; g.sLp = 0; g.sUp = 0;
    dim3 grid((unsigned)cdiv64(N, G_TN), (unsigned)cdiv64(M, G_TM));
    zgemm_sub_kernel<<<grid, G_THREADS, G_SMEM, st>>>(g);
    BHS_CHECK_LAUNCH();
    return BHS_OK;
}
