// FP64 tensor-core (DMMA.8x8x4) contraction core of the greedy design path.
//
//   T[i,j] = sum_k A[k*lda + i] * B[k*ldb + j]           both operands K-major (row k contiguous)
//
// with an optional *prologue* that evaluates the covariance k(x_i, y_j) on the tensor pipe as well
// (expanded form  k = f(alpha_i + beta_j + sum_q u_q(i) v_q(j)),  q < 16 extra K rows), and three epilogues:
//
//   EPI_IVAR   r[j] += sum_i (k(i,j) - T[i,j])^2          K5: IVAR scoring, experimentalDesign.py:105-117 restated
//   EPI_STORE  out[i,j] = k(i,j) - T[i,j]                  K1+K3: block row of the left-looking TRSM
//   EPI_SUB    C[i,j]  -= T[i,j]                           K2: Cholesky trailing update / materialised TRSM
//
// sm_100a has no f64 kind in tcgen05, so the FP64 tensor path is warp-level mma.sync.m8n8k4 (SASS DMMA.8x8x4)
// fed from shared memory.  CTA tile 128x128, 8 warps of 32x64, K chunks of 16 rows through a 4-stage
// cp.async (LDGSTS) ring, rows padded to 132 doubles so that the 16-byte fragment loads are conflict-free.
// The fragment <-> matrix index map is permuted so that every thread reads 2 adjacent doubles per LDS.128:
//   i_local = wm*32 + (t>>1)*16 + (lane>>2)*2 + (t&1)            t = 0..3  (A fragments)
//   j_local = wn*64 + (u>>1)*16 + (lane>>2)*2 + (u&1)            u = 0..7  (B fragments, load side)
//   accumulator (t,u,e) sits at column  wn*64 + (u>>1)*16 + ((lane&3)*2+e)*2 + (u&1)
#include <math.h>

#include "gpx_common.cuh"

namespace {

constexpr int BM = 128;
constexpr int BN = 128;
constexpr int BK = 16;
constexpr int STAGES = 4;
constexpr int NTHREADS = 256;
constexpr int LDSM = 132;                       // padded row stride (doubles)
constexpr int STAGE_DOUBLES = 2 * BK * LDSM;    // A rows then B rows
constexpr int SMEM_DOUBLES = STAGES * STAGE_DOUBLES + BM + BN + 4 * BN;
constexpr size_t SMEM_BYTES = (size_t)SMEM_DOUBLES * sizeof(double);

enum { EPI_IVAR = 0, EPI_STORE = 1, EPI_SUB = 2 };

struct CoreArgs {
    const double* A;   // K x I (main operand)
    const double* B;   // K x J
    const double* Ap;  // prologue rows (GPX_KROWS x lda), same column space as A
    const double* Bp;  // prologue rows (GPX_KROWS x ldb)
    const double* As;  // alpha[I]
    const double* Bs;  // beta[J]
    double* out;       // IVAR: partial sums [split][ldo] ; STORE / SUB: row-major I x J
    int64_t lda, ldb, ldo;
    int64_t I, J;
    int K;
    int dpad;          // prologue K extent, multiple of 4
    int tiles_per_cta; // IVAR: i-tiles each CTA walks
    int upper_only;
};

__device__ __forceinline__ void cp_async16(double* dst, const double* src, bool ok) {
    const unsigned int d = (unsigned int)__cvta_generic_to_shared(dst);
    const int sz = ok ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

template <int FAM, int EPI, bool PRO>
__global__ void __launch_bounds__(NTHREADS, 1)
    dmma_core_kernel(const __grid_constant__ CoreArgs a, const __grid_constant__ KParams kp) {
    extern __shared__ __align__(16) double smem[];
    double* s_alpha = smem + STAGES * STAGE_DOUBLES;
    double* s_beta = s_alpha + BM;
    double* s_red = s_beta + BN;

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int wm = warp & 3, wn = warp >> 2;
    const int g4 = lane >> 2, q4 = lane & 3;

    const int64_t j0 = (int64_t)blockIdx.x * BN;
    const int64_t itiles = (a.I + BM - 1) / BM;
    int64_t it_begin, it_end;
    if (EPI == EPI_IVAR) {
        it_begin = (int64_t)blockIdx.y * a.tiles_per_cta;
        it_end = it_begin + a.tiles_per_cta;
        if (it_end > itiles) it_end = itiles;
    } else {
        it_begin = blockIdx.y;
        it_end = it_begin + 1;
        // symmetric update: nothing to do for tiles strictly below the diagonal
        if (a.upper_only && it_begin * BM >= j0 + BN) return;
    }
    const int ntiles = it_end > it_begin ? (int)(it_end - it_begin) : 0;
    const int kch = (a.K + BK - 1) / BK;
    const int T = (PRO ? 1 : 0) + kch;   // chunks per tile
    const int G = ntiles * T;

    // ---- chunk loader: 2048 16-byte pieces per chunk, 8 per thread ---------------------------------
    auto issue = [&](int g) {
        if (g < G) {
            const int tl = g / T;
            const int ch = g - tl * T;
            const int64_t i0 = (it_begin + tl) * BM;
            const double *srcA, *srcB;
            int krows;
            if (PRO && ch == 0) {
                srcA = a.Ap;
                srcB = a.Bp;
                krows = a.dpad;
            } else {
                const int kc = ch - (PRO ? 1 : 0);
                srcA = a.A + (int64_t)kc * BK * a.lda;
                srcB = a.B + (int64_t)kc * BK * a.ldb;
                krows = a.K - kc * BK;
            }
            double* sA = smem + (g % STAGES) * STAGE_DOUBLES;
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const int p = tid + it * NTHREADS;
                const int mat = p >> 10;
                const int row = (p >> 6) & 15;
                const int c2 = (p & 63) * 2;
                const double* base = mat ? srcB : srcA;
                const int64_t ld = mat ? a.ldb : a.lda;
                const int64_t col = (mat ? j0 : i0) + c2;
                const bool ok = (row < krows) && (col < (mat ? a.J : a.I));
                cp_async16(sA + mat * (BK * LDSM) + row * LDSM + c2, ok ? base + (int64_t)row * ld + col : base, ok);
            }
        }
        cp_async_commit();
    };

    double acc[4][8][2];
    double rs[8][2];
#pragma unroll
    for (int u = 0; u < 8; ++u) rs[u][0] = rs[u][1] = 0.0;

    if (PRO) {
        if (tid < BN) s_beta[tid] = (j0 + tid < a.J) ? a.Bs[j0 + tid] : 0.0;
    }

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) issue(s);

    int tl = 0, ch = 0;
    for (int g = 0; g < G; ++g) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        issue(g + STAGES - 1);

        const int64_t i0 = (it_begin + tl) * BM;
        if (ch == 0) {
#pragma unroll
            for (int t = 0; t < 4; ++t)
#pragma unroll
                for (int u = 0; u < 8; ++u) acc[t][u][0] = acc[t][u][1] = 0.0;
            if (PRO && tid < BM) s_alpha[tid] = (i0 + tid < a.I) ? a.As[i0 + tid] : 0.0;
        }

        const double* sA = smem + (g % STAGES) * STAGE_DOUBLES;
        const double* sB = sA + BK * LDSM;
        int ksteps;
        if (PRO && ch == 0) {
            ksteps = a.dpad >> 2;
        } else {
            const int kc = ch - (PRO ? 1 : 0);
            const int rem = a.K - kc * BK;
            ksteps = rem >= BK ? 4 : ((rem + 3) >> 2);
        }
        const double* pa = sA + q4 * LDSM + wm * 32 + g4 * 2;
        const double* pb = sB + q4 * LDSM + wn * 64 + g4 * 2;
#pragma unroll 1
        for (int ks = 0; ks < ksteps; ++ks) {
            double af[4], bf[8];
#pragma unroll
            for (int t2 = 0; t2 < 2; ++t2) {
                const double2 v = *reinterpret_cast<const double2*>(pa + ks * 4 * LDSM + t2 * 16);
                af[t2 * 2] = v.x;
                af[t2 * 2 + 1] = v.y;
            }
#pragma unroll
            for (int u2 = 0; u2 < 4; ++u2) {
                const double2 v = *reinterpret_cast<const double2*>(pb + ks * 4 * LDSM + u2 * 16);
                bf[u2 * 2] = v.x;
                bf[u2 * 2 + 1] = v.y;
            }
#pragma unroll
            for (int t = 0; t < 4; ++t)
#pragma unroll
                for (int u = 0; u < 8; ++u) dmma(acc[t][u][0], acc[t][u][1], af[t], bf[u]);
        }

        if (PRO && ch == 0) {
            // covariance from the expanded form; accumulators become -k so that the main loop yields T - k
            __syncthreads();  // s_alpha (and, first time, s_beta) visible
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const double al = s_alpha[wm * 32 + (t >> 1) * 16 + g4 * 2 + (t & 1)];
#pragma unroll
                for (int u = 0; u < 8; ++u)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const double be = s_beta[wn * 64 + (u >> 1) * 16 + (q4 * 2 + e) * 2 + (u & 1)];
                        acc[t][u][e] = -kexpand<FAM>(acc[t][u][e] + al + be, kp);
                    }
            }
        }

        if (ch == T - 1) {
            // ---- tile epilogue ---------------------------------------------------------------------
            if (EPI == EPI_IVAR) {
                const bool full = (i0 + BM <= a.I);
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const int64_t i = i0 + wm * 32 + (t >> 1) * 16 + g4 * 2 + (t & 1);
                    const bool ok = full || (i < a.I);
#pragma unroll
                    for (int u = 0; u < 8; ++u)
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const double v = ok ? acc[t][u][e] : 0.0;
                            rs[u][e] = fma(v, v, rs[u][e]);
                        }
                }
            } else {
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const int64_t i = i0 + wm * 32 + (t >> 1) * 16 + g4 * 2 + (t & 1);
                    if (i >= a.I) continue;
#pragma unroll
                    for (int u2 = 0; u2 < 4; ++u2)
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int64_t j = j0 + wn * 64 + u2 * 16 + (q4 * 2 + e) * 2;
                            if (j >= a.J) continue;
                            double* dst = a.out + i * a.ldo + j;
                            const double v0 = acc[t][u2 * 2][e], v1 = acc[t][u2 * 2 + 1][e];
                            if (EPI == EPI_STORE) {
                                if (j + 1 < a.J) {
                                    *reinterpret_cast<double2*>(dst) = make_double2(-v0, -v1);
                                } else {
                                    dst[0] = -v0;
                                }
                            } else {
                                if (j + 1 < a.J) {
                                    double2 c = *reinterpret_cast<double2*>(dst);
                                    c.x -= v0;
                                    c.y -= v1;
                                    *reinterpret_cast<double2*>(dst) = c;
                                } else {
                                    dst[0] -= v0;
                                }
                            }
                        }
                }
            }
            ch = 0;
            ++tl;
        } else {
            ++ch;
        }
    }
    cp_async_wait<0>();

    if (EPI == EPI_IVAR) {
        // reduce the per-thread partial sums over the 8 lanes that share a column, then over the 4 i-warps
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                double v = rs[u][e];
                v += __shfl_xor_sync(0xffffffffu, v, 4);
                v += __shfl_xor_sync(0xffffffffu, v, 8);
                v += __shfl_xor_sync(0xffffffffu, v, 16);
                rs[u][e] = v;
            }
        __syncthreads();
        if (g4 == 0) {
#pragma unroll
            for (int u = 0; u < 8; ++u)
#pragma unroll
                for (int e = 0; e < 2; ++e)
                    s_red[wm * BN + wn * 64 + (u >> 1) * 16 + (q4 * 2 + e) * 2 + (u & 1)] = rs[u][e];
        }
        __syncthreads();
        if (tid < BN && j0 + tid < a.J) {
            const double r = ((s_red[tid] + s_red[BN + tid]) + s_red[2 * BN + tid]) + s_red[3 * BN + tid];
            a.out[(int64_t)blockIdx.y * a.ldo + j0 + tid] = r;
        }
    }
}

template <int FAM, int EPI, bool PRO>
int launch_core(const CoreArgs& a, const KParams& kp, dim3 grid, cudaStream_t st) {
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(dmma_core_kernel<FAM, EPI, PRO>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)SMEM_BYTES);
        if (e != cudaSuccess) {
            gpx_set_error("dmma core: cannot opt in to %zu bytes of shared memory: %s", SMEM_BYTES, cudaGetErrorString(e));
            return (int)e;
        }
        configured = true;
    }
    dmma_core_kernel<FAM, EPI, PRO><<<grid, NTHREADS, SMEM_BYTES, st>>>(a, kp);
    return gpx_check_launch("dmma core");
}

int check_operand(const double* p, int64_t ld, const char* name) {
    if (!gpx_aligned16(p) || (ld & 1)) {
        gpx_set_error("dmma core: operand %s must be 16-byte aligned with an even leading dimension", name);
        return GPX_EALIGN;
    }
    return GPX_OK;
}

}  // namespace

// number of i-splits for the IVAR grid: fill the machine in whole waves
int gpx_ivar_splits(gpx_handle h, int64_t M, int64_t C) {
    const int64_t jt = (C + BN - 1) / BN;
    const int64_t itl = (M + BM - 1) / BM;
    const int sms = h->sm_count > 0 ? h->sm_count : 148;
    int best = 1;
    double best_eff = -1.0;
    for (int s = 1; s <= 32 && s <= itl; ++s) {
        const int64_t tps = (itl + s - 1) / s;
        const int64_t seff = (itl + tps - 1) / tps;  // splits that actually have work
        if (seff != s) continue;
        const int64_t ctas = jt * s;
        const int64_t waves = (ctas + sms - 1) / sms;
        // time ~ waves * tiles-per-cta ; ideal = jt*itl / sms
        const double eff = (double)(jt * itl) / (double)(waves * sms * tps);
        if (eff > best_eff + 1e-9) {
            best_eff = eff;
            best = s;
        }
    }
    return best;
}

int gpx_launch_core_ivar(gpx_handle h, const double* Wm, int64_t ldm, const double* Ma_rows, const double* Ma_scal,
                         int64_t M, const double* Wc, int64_t ldc, const double* Cb_rows, const double* Cb_scal,
                         int64_t C, int64_t n, double* partial, int64_t ldp, int* nsplit_out, cudaStream_t st) {
    int rc;
    if ((rc = check_operand(Ma_rows, ldm, "Ma_rows"))) return rc;
    if ((rc = check_operand(Cb_rows, ldc, "Cb_rows"))) return rc;
    if (n > 0) {
        if ((rc = check_operand(Wm, ldm, "Wm"))) return rc;
        if ((rc = check_operand(Wc, ldc, "Wc"))) return rc;
    }
    const int splits = gpx_ivar_splits(h, M, C);
    const int64_t itl = (M + BM - 1) / BM;
    CoreArgs a;
    a.A = Wm;
    a.B = Wc;
    a.Ap = Ma_rows;
    a.Bp = Cb_rows;
    a.As = Ma_scal;
    a.Bs = Cb_scal;
    a.out = partial;
    a.lda = ldm;
    a.ldb = ldc;
    a.ldo = ldp;
    a.I = M;
    a.J = C;
    a.K = (int)n;
    a.dpad = (h->kp.d + 3) & ~3;
    a.tiles_per_cta = (int)((itl + splits - 1) / splits);
    a.upper_only = 0;
    dim3 grid((unsigned)((C + BN - 1) / BN), (unsigned)splits);
    *nsplit_out = splits;
    GPX_DISPATCH_FAMILY(h->kp.family, rc = (launch_core<FAM, EPI_IVAR, true>(a, h->kp, grid, st)));
    return rc;
}

int gpx_launch_core_store(gpx_handle h, const double* A, int64_t lda, const double* Ap, const double* As, int64_t I,
                          const double* B, int64_t ldb, const double* Bp, const double* Bs, int64_t J, int64_t K,
                          double* out, int64_t ldo, cudaStream_t st) {
    int rc;
    if ((rc = check_operand(Ap, lda, "Ap"))) return rc;
    if ((rc = check_operand(Bp, ldb, "Bp"))) return rc;
    if ((rc = check_operand(out, ldo, "out"))) return rc;
    if (K > 0) {
        if ((rc = check_operand(A, lda, "A"))) return rc;
        if ((rc = check_operand(B, ldb, "B"))) return rc;
    }
    CoreArgs a;
    a.A = A;
    a.B = B;
    a.Ap = Ap;
    a.Bp = Bp;
    a.As = As;
    a.Bs = Bs;
    a.out = out;
    a.lda = lda;
    a.ldb = ldb;
    a.ldo = ldo;
    a.I = I;
    a.J = J;
    a.K = (int)K;
    a.dpad = (h->kp.d + 3) & ~3;
    a.tiles_per_cta = 1;
    a.upper_only = 0;
    dim3 grid((unsigned)((J + BN - 1) / BN), (unsigned)((I + BM - 1) / BM));
    GPX_DISPATCH_FAMILY(h->kp.family, rc = (launch_core<FAM, EPI_STORE, true>(a, h->kp, grid, st)));
    return rc;
}

extern "C" int gpx_dgemm_tn_sub(gpx_handle h, const double* A, int64_t lda, const double* B, int64_t ldb, double* C,
                                int64_t ldc, int64_t I, int64_t J, int64_t K, int upper_only, void* stream) {
    GPX_REQUIRE(h != nullptr, GPX_EINVAL, "handle is NULL");
    GPX_REQUIRE(I >= 0 && J >= 0 && K >= 0, GPX_EINVAL, "negative size");
    if (I == 0 || J == 0 || K == 0) return GPX_OK;
    GPX_REQUIRE(A && B && C, GPX_EINVAL, "NULL pointer");
    GPX_REQUIRE((I + BM - 1) / BM <= 65535, GPX_ESIZE, "I too large for one launch");
    int rc;
    if ((rc = check_operand(A, lda, "A"))) return rc;
    if ((rc = check_operand(B, ldb, "B"))) return rc;
    if ((rc = check_operand(C, ldc, "C"))) return rc;
    CoreArgs a;
    a.A = A;
    a.B = B;
    a.Ap = a.Bp = a.As = a.Bs = nullptr;
    a.out = C;
    a.lda = lda;
    a.ldb = ldb;
    a.ldo = ldc;
    a.I = I;
    a.J = J;
    a.K = (int)K;
    a.dpad = 0;
    a.tiles_per_cta = 1;
    a.upper_only = upper_only;
    KParams kp = h->kp;
    dim3 grid((unsigned)((J + BN - 1) / BN), (unsigned)((I + BM - 1) / BM));
    return launch_core<GPX_SE, EPI_SUB, false>(a, kp, grid, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------
// K5 + K7: IVAR scores of every candidate and their arg-min
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ivar_finalize_kernel(const double* __restrict__ partial, int nsplit, int64_t ldp,
                                                             const double* __restrict__ varC, const double* __restrict__ sumVarM,
                                                             int64_t M, int64_t C, double noise, double zero_tol,
                                                             double* __restrict__ score) {
    const int64_t c = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (c >= C) return;
    double r = 0.0;
    for (int s = 0; s < nsplit; ++s) r += partial[(int64_t)s * ldp + c];
    const double base = sumVarM[0] / (double)M;            // (1/nMC) sum varMC   experimentalDesign.py:109
    const double den = varC[c] + noise;
    const double red = (den <= zero_tol) ? 0.0 : (r / den) / (double)M;
    score[c] = fabs(base - red);                           // np.abs(cost)        experimentalDesign.py:117
}

extern "C" int64_t gpx_score_ivar_workspace(gpx_handle h, int64_t M, int64_t C) {
    if (!h || M < 0 || C < 0) return 0;
    const int64_t ldp = (C + 1) & ~(int64_t)1;
    return (int64_t)gpx_ivar_splits(h, M, C) * ldp;
}

extern "C" int gpx_score_ivar(gpx_handle h, const double* Wm, int64_t ldm, const double* varM, const double* Ma_rows,
                              const double* Ma_scal, int64_t M, const double* Wc, int64_t ldc, const double* varC,
                              const double* Cb_rows, const double* Cb_scal, int64_t C, int64_t n, double noise,
                              double zero_tol, const uint8_t* mask, double* workspace, double* score_out, double* best,
                              int64_t* idx, void* stream) {
    GPX_NEED_KERNEL(h);
    GPX_REQUIRE(M >= 1 && C >= 1 && n >= 0, GPX_EINVAL, "bad sizes");
    GPX_REQUIRE(varM && Ma_rows && Ma_scal && varC && Cb_rows && Cb_scal && workspace && score_out && best && idx,
                GPX_EINVAL, "NULL pointer");
    GPX_REQUIRE(n == 0 || (Wm && Wc), GPX_EINVAL, "W is NULL");
    cudaStream_t st = (cudaStream_t)stream;
    int rc = gpx_sum_impl(h, varM, M, h->scal, st);
    if (rc) return rc;
    const int64_t ldp = (C + 1) & ~(int64_t)1;
    int nsplit = 1;
    rc = gpx_launch_core_ivar(h, Wm, ldm, Ma_rows, Ma_scal, M, Wc, ldc, Cb_rows, Cb_scal, C, n, workspace, ldp, &nsplit, st);
    if (rc) return rc;
    ivar_finalize_kernel<<<(unsigned)((C + 255) / 256), 256, 0, st>>>(workspace, nsplit, ldp, varC, h->scal, M, C, noise,
                                                                     zero_tol, score_out);
    rc = gpx_check_launch("gpx_score_ivar finalize");
    if (rc) return rc;
    return gpx_argreduce_impl(h, score_out, nullptr, mask, C, 1, best, idx, st);
}

// ---------------------------------------------------------------------------------------------
// Yard-sticks for bench.py: raw DMMA and DFMA issue rates (no memory traffic)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bench_dmma_kernel(int64_t iters, double* sink) {
    double c[16][2];
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i][0] = c[i][1] = 0.0;
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    for (int64_t it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) dmma(c[i][0], c[i][1], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) sink[0] = s;
}

__global__ void __launch_bounds__(256) bench_dfma_kernel(int64_t iters, double* sink) {
    double c[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = i;
    const double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9;
    for (int64_t it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) c[i] = fma(c[i], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i];
    if (s == 123.456) sink[0] = s;
}

// each launch: sm_count*4 CTAs of 8 warps; flops = ctas*8 warps*iters*16 DMMA*512  (DFMA: ctas*256 thr*iters*16*2)
extern "C" int gpx_bench_dmma(gpx_handle h, int64_t iters, double* sink, void* stream) {
    GPX_REQUIRE(h && sink && iters > 0, GPX_EINVAL, "bad arguments");
    bench_dmma_kernel<<<h->sm_count * 4, 256, 0, (cudaStream_t)stream>>>(iters, sink);
    return gpx_check_launch("gpx_bench_dmma");
}
extern "C" int gpx_bench_dfma(gpx_handle h, int64_t iters, double* sink, void* stream) {
    GPX_REQUIRE(h && sink && iters > 0, GPX_EINVAL, "bad arguments");
    bench_dfma_kernel<<<h->sm_count * 4, 256, 0, (cudaStream_t)stream>>>(iters, sink);
    return gpx_check_launch("gpx_bench_dfma");
}
