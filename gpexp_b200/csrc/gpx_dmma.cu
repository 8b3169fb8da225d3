// FP64 tensor-core (DMMA.8x8x4) contractions of the greedy design path.
//
//   T[i,j] = sum_k A[k*lda + i] * B[k*ldb + j]           both operands K-major (row k contiguous)
//
// with an optional *prologue* that evaluates the covariance k(x_i, y_j) on the tensor pipe as well
// (expanded form  k = f(alpha_i + beta_j + sum_q u_q(i) v_q(j)),  q < 16 extra K rows), and three epilogues:
//
//   EPI_IVAR   r[j] += sum_i (k(i,j) - T[i,j])^2          K5: IVAR scoring, experimentalDesign.py:105-117 restated
//   EPI_STORE  out[i,j] = k(i,j) - T[i,j]                  K1+K3: block row of the left-looking TRSM
//   EPI_SUB    C[i,j]  -= T[i,j]                           K2: Cholesky trailing update / materialised TRSM / MI set-up
//
// sm_100a has no f64 kind in tcgen05, so the FP64 tensor path is warp-level mma.sync.m8n8k4 (SASS DMMA.8x8x4)
// fed from shared memory.  Two kernels share the fragment layout:
//
//   ivar_ws_kernel     THE hot kernel (EPI_IVAR on fully padded operands): 128x128 CTA tile, 8 warps of 32x64,
//                      operand chunks by TMA bulk copies completing on mbarriers (3-stage ring of 32-row chunks),
//                      no CTA barrier in the main loop, table-driven exp prologue, butterfly column reduction.
//   dmma_core_kernel   the generic predicated kernel (cp.async ring + __syncthreads, zero-filled edges) used for
//                      EPI_STORE / EPI_SUB and as the fallback for unpadded IVAR operands.
//
// Rows are padded to 132 doubles and the fragment <-> matrix index map is permuted so that every thread reads
// 2 adjacent doubles per conflict-free LDS.128:
//   i_local = wm*32 + (t>>1)*16 + (lane>>2)*2 + (t&1)            t = 0..3  (A fragments)
//   j_local = wn*64 + (u>>1)*16 + (lane>>2)*2 + (u&1)            u = 0..7  (B fragments, load side)
//   accumulator (t,u,e) sits at column  wn*64 + (u>>1)*16 + ((lane&3)*2+e)*2 + (u&1)
#include <math.h>
#include <stdlib.h>

#include "gpx_common.cuh"

#ifndef GPX_DEFAULT_WM
#define GPX_DEFAULT_WM 2
#endif
#ifndef GPX_DEFAULT_IVAR_TN
#define GPX_DEFAULT_IVAR_TN 8
#endif
#ifndef GPX_DEFAULT_IVAR_GROUP
#define GPX_DEFAULT_IVAR_GROUP 1
#endif

namespace {

// expanded-form covariance with the table exp; tab already carries the signal variance
template <int FAM>
__device__ __forceinline__ double kexpand_tab(double e, const KParams& kp, const double* __restrict__ tab) {
    if (FAM == GPX_MATERN32) {
        const double t = kp.c0 * sqrt(fmax(e, 0.0));
        return (1.0 + t) * gpx_exp_tab(-t, tab);
    }
    return gpx_exp_tab(e, tab);
}

constexpr int BN = 128;
constexpr int BK = 16;
constexpr int STAGES = 4;

enum { EPI_IVAR = 0, EPI_STORE = 1, EPI_SUB = 2 };

// WM = warps along i.  WM=4: 256 threads, 128x128 tile, 1 CTA/SM.  WM=2: 128 threads, 64x128 tile, 2 CTAs/SM
// (two independent CTAs de-phase the barriers and the exp prologue against each other's DMMA stream).
template <int WM>
struct Cfg {
    static constexpr int NT = WM * 64;
    static constexpr int BM = WM * 32;
    static constexpr int LDA = BM + 4;  // row strides = 4 (mod 16) doubles: conflict-free LDS.128 fragment loads
    static constexpr int LDB = BN + 4;
    static constexpr int STAGE = BK * (LDA + LDB);
    static constexpr int PIECES_A = BK * BM / 2;  // 16-byte pieces per chunk
    static constexpr int PIECES = BK * (BM + BN) / 2;
    static constexpr int PER_THREAD = PIECES / NT;
    static constexpr int A_ITERS = PIECES_A / NT;
    static constexpr int SMEM_DOUBLES = STAGES * STAGE + 2 * BM + BN + WM * BN;
    static constexpr size_t SMEM_BYTES = (size_t)SMEM_DOUBLES * sizeof(double);
    static_assert(PIECES % NT == 0 && PIECES_A % NT == 0, "loader mapping");
};

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
    int upper_only;    // SUB: skip tiles below the diagonal ; ivar_ws_kernel: candidate tiles per wave group
    int ldo_splits;    // ivar_ws_kernel: number of M-splits
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

template <int LDA_, int LDB_>
__device__ __forceinline__ void load_frags(double (&af)[4], double (&bf)[8], const double* pa, const double* pb, int ks) {
#pragma unroll
    for (int t2 = 0; t2 < 2; ++t2) {
        const double2 v = *reinterpret_cast<const double2*>(pa + ks * 4 * LDA_ + t2 * 16);
        af[t2 * 2] = v.x;
        af[t2 * 2 + 1] = v.y;
    }
#pragma unroll
    for (int u2 = 0; u2 < 4; ++u2) {
        const double2 v = *reinterpret_cast<const double2*>(pb + ks * 4 * LDB_ + u2 * 16);
        bf[u2 * 2] = v.x;
        bf[u2 * 2 + 1] = v.y;
    }
}

__device__ __forceinline__ void mma_tile(double (&acc)[4][8][2], const double (&af)[4], const double (&bf)[8]) {
#pragma unroll
    for (int t = 0; t < 4; ++t)
#pragma unroll
        for (int u = 0; u < 8; ++u) dmma(acc[t][u][0], acc[t][u][1], af[t], bf[u]);
}

template <int FAM, int EPI, bool PRO, int WM>
__global__ void __launch_bounds__(WM * 64, WM == 4 ? 1 : 2)
    dmma_core_kernel(const __grid_constant__ CoreArgs a, const __grid_constant__ KParams kp) {
    using C = Cfg<WM>;
    constexpr int BM = C::BM, NT = C::NT, LDA = C::LDA, LDB = C::LDB;
    extern __shared__ __align__(16) double smem[];
    double* s_alpha = smem + STAGES * C::STAGE;  // [2][BM], by tile parity
    double* s_beta = s_alpha + 2 * BM;
    double* s_red = s_beta + BN;

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int wm = warp % WM, wn = warp / WM;
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
    const int T = (PRO ? 1 : 0) + kch;  // chunks per tile
    const int G = ntiles * T;

    // ---- chunk loader (cp.async ring, 3 chunks in flight) ------------------------------------------
    int is_g = 0, is_tl = 0, is_ch = 0;
    auto issue = [&]() {
        if (is_g < G) {
            const int64_t i0 = (it_begin + is_tl) * BM;
            const double *srcA, *srcB;
            int krows;
            if (PRO && is_ch == 0) {
                srcA = a.Ap;
                srcB = a.Bp;
                krows = a.dpad;
            } else {
                const int kc = is_ch - (PRO ? 1 : 0);
                srcA = a.A + (int64_t)kc * BK * a.lda;
                srcB = a.B + (int64_t)kc * BK * a.ldb;
                krows = a.K - kc * BK;
            }
            double* sA = smem + (is_g % STAGES) * C::STAGE;
            double* sB = sA + BK * LDA;
#pragma unroll
            for (int it = 0; it < C::PER_THREAD; ++it) {
                if (it < C::A_ITERS) {
                    const int p = tid + it * NT;
                    const int row = p / (BM / 2), c2 = (p % (BM / 2)) * 2;
                    const bool ok = (row < krows) && (i0 + c2 < a.I);
                    cp_async16(sA + row * LDA + c2, ok ? srcA + (int64_t)row * a.lda + i0 + c2 : srcA, ok);
                } else {
                    const int p = tid + it * NT - C::PIECES_A;
                    const int row = p / (BN / 2), c2 = (p % (BN / 2)) * 2;
                    const bool ok = (row < krows) && (j0 + c2 < a.J);
                    cp_async16(sB + row * LDB + c2, ok ? srcB + (int64_t)row * a.ldb + j0 + c2 : srcB, ok);
                }
            }
            if (++is_ch == T) {
                is_ch = 0;
                ++is_tl;
            }
        }
        ++is_g;
        cp_async_commit();
    };
    // chunk g+1 landed and visible to all warps; the stage of chunk g-1 is free -> refill it with chunk g+3
    auto midsync = [&]() {
        cp_async_wait<1>();
        __syncthreads();
        issue();
    };

    double acc[4][8][2];
    double rs[2] = {0.0, 0.0};
    double af0[4], bf0[8], af1[4], bf1[8];

    if (PRO) {
        if (tid < BN) s_beta[tid] = (j0 + tid < a.J) ? a.Bs[j0 + tid] : 0.0;
    }

    issue();
    issue();
    issue();
    cp_async_wait<2>();
    __syncthreads();

    const int fa = q4 * LDA + wm * 32 + g4 * 2;  // fragment offsets inside a stage
    const int fb = q4 * LDB + wn * 64 + g4 * 2;
    int tl = 0, ch = 0;
    bool pre = false;
    for (int g = 0; g < G; ++g) {
        const int64_t i0 = (it_begin + tl) * BM;
        if (ch == 0) {
#pragma unroll
            for (int t = 0; t < 4; ++t)
#pragma unroll
                for (int u = 0; u < 8; ++u) acc[t][u][0] = acc[t][u][1] = 0.0;
            if (PRO && tid < BM) s_alpha[(tl & 1) * BM + tid] = (i0 + tid < a.I) ? a.As[i0 + tid] : 0.0;
        }
        const double* pa = smem + (g % STAGES) * C::STAGE + fa;
        const double* pb = smem + (g % STAGES) * C::STAGE + BK * LDA + fb;
        int ksteps;
        bool next_full;
        {
            const int kc = ch - (PRO ? 1 : 0);
            if (PRO && ch == 0) {
                ksteps = a.dpad >> 2;
            } else {
                const int rem = a.K - kc * BK;
                ksteps = rem >= BK ? 4 : ((rem + 3) >> 2);
            }
            next_full = (a.K - (kc + 1) * BK) >= BK;
        }
        if (ksteps == 4) {
            // software-pipelined: fragments of k-step s+1 are fetched while the DMMAs of k-step s issue
            if (!pre) load_frags<LDA, LDB>(af0, bf0, pa, pb, 0);
            load_frags<LDA, LDB>(af1, bf1, pa, pb, 1);
            mma_tile(acc, af0, bf0);
            load_frags<LDA, LDB>(af0, bf0, pa, pb, 2);
            mma_tile(acc, af1, bf1);
            load_frags<LDA, LDB>(af1, bf1, pa, pb, 3);
            mma_tile(acc, af0, bf0);
            midsync();
            const bool np = (ch + 1 < T) && !(PRO && ch == 0) && next_full;
            if (np) {
                const double* na = smem + ((g + 1) % STAGES) * C::STAGE + fa;
                const double* nb = smem + ((g + 1) % STAGES) * C::STAGE + BK * LDA + fb;
                load_frags<LDA, LDB>(af0, bf0, na, nb, 0);
            }
            mma_tile(acc, af1, bf1);
            pre = np;
        } else {
#pragma unroll 1
            for (int ks = 0; ks < ksteps; ++ks) {
                if (ks == ksteps - 1) midsync();
                load_frags<LDA, LDB>(af0, bf0, pa, pb, ks);
                mma_tile(acc, af0, bf0);
            }
            pre = false;
        }

        if (PRO && ch == 0) {
            // covariance from the expanded form; accumulators become -k so that the main loop yields T - k
            const double* sal = s_alpha + (tl & 1) * BM;
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const double al = sal[wm * 32 + (t >> 1) * 16 + g4 * 2 + (t & 1)];
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
                double p[8][2];
#pragma unroll
                for (int u = 0; u < 8; ++u) p[u][0] = p[u][1] = 0.0;
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const int64_t i = i0 + wm * 32 + (t >> 1) * 16 + g4 * 2 + (t & 1);
                    const bool ok = full || (i < a.I);
#pragma unroll
                    for (int u = 0; u < 8; ++u)
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const double v = ok ? acc[t][u][e] : 0.0;
                            p[u][e] = fma(v, v, p[u][e]);
                        }
                }
                // butterfly reduce-scatter over the 8 lanes that share q4: lane g4 ends up owning u = g4
                const bool b4 = (lane & 16) != 0, b3 = (lane & 8) != 0, b2 = (lane & 4) != 0;
                double h[4][2], q[2][2];
#pragma unroll
                for (int u = 0; u < 4; ++u)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const double keep = b4 ? p[u + 4][e] : p[u][e];
                        const double send = b4 ? p[u][e] : p[u + 4][e];
                        h[u][e] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
                    }
#pragma unroll
                for (int u = 0; u < 2; ++u)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const double keep = b3 ? h[u + 2][e] : h[u][e];
                        const double send = b3 ? h[u][e] : h[u + 2][e];
                        q[u][e] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
                    }
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const double keep = b2 ? q[1][e] : q[0][e];
                    const double send = b2 ? q[0][e] : q[1][e];
                    rs[e] += keep + __shfl_xor_sync(0xffffffffu, send, 4);
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
        // lane (g4,q4) owns columns u = g4, e = 0,1 of its warp; add the WM i-warps in a fixed order
        __syncthreads();
#pragma unroll
        for (int e = 0; e < 2; ++e) s_red[wm * BN + wn * 64 + (g4 >> 1) * 16 + (q4 * 2 + e) * 2 + (g4 & 1)] = rs[e];
        __syncthreads();
        for (int c = tid; c < BN; c += NT) {
            if (j0 + c < a.J) {
                double r = s_red[c];
#pragma unroll
                for (int w = 1; w < WM; ++w) r += s_red[w * BN + c];
                a.out[(int64_t)blockIdx.y * a.ldo + j0 + c] = r;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// TMA + mbarrier IVAR contraction (the hot kernel): operand chunks arrive by TMA bulk copies
// (cp.async.bulk -> UBLKCP) that complete on mbarriers; the eight warps do LDS.128 + DMMA and take turns
// (warp g mod 8 for chunk g+WS_AHEAD) at issuing the bulk copies of a chunk -- a ninth, dedicated producer warp
// would put three warps on one sub-partition and cap everybody at 168 registers.  No CTA-wide barrier in
// the main loop, so warps drift apart and one warp's exp prologue / epilogue overlaps the other warp's
// DMMA stream on the same sub-partition.
// Requires fully padded operands: lda, ldb multiples of 128 covering whole tiles (the engines guarantee it).
// ---------------------------------------------------------------------------------------------
constexpr int WS_BM = 128;
constexpr int WS_LD = 132;
// Ring geometry.  Measured on B200 (n = 2047 / n = 255, C = M = 100k): 16 rows x 6 stages, 3 chunks ahead: 34.18 TFLOP/s /
// 165.5 ms; 32 rows x 3 stages, 1 ahead: 34.70 / 165.3 ms; 32 x 3 stages, 2 ahead: 23.7 / 240 ms -- the stage that is
// refilled must have been released at least one whole chunk ago, or the producing warp blocks on the slowest consumer.
#ifndef GPX_WS_BK
#define GPX_WS_BK 32
#define GPX_WS_STAGES 3
#define GPX_WS_AHEAD 1
#endif
constexpr int WS_BK = GPX_WS_BK;           // K rows per chunk
constexpr int WS_KSTEPS = WS_BK / 4;
constexpr int WS_STAGE = WS_BK * 2 * WS_LD;  // doubles
constexpr int WS_STAGES = GPX_WS_STAGES;   // ring depth
constexpr int WS_AHEAD = GPX_WS_AHEAD;     // chunks in flight ahead of the consumers
constexpr int WS_SMEM_DOUBLES = WS_STAGES * WS_STAGE + WS_STAGES * WS_BM + BN + 4 * BN + 2 * WS_STAGES + 256;
constexpr size_t WS_SMEM_BYTES = (size_t)WS_SMEM_DOUBLES * sizeof(double);

__device__ __forceinline__ unsigned int smem_u32(const void* p) { return (unsigned int)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, unsigned int bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned int parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(double* dst, const double* src, unsigned int bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// TN = B fragments per warp: 8 -> 8 warps of 32x64 (254 registers), 4 -> 16 warps of 32x32 (<= 128 registers, four
// warps per sub-partition: the DMMA pipe only idles when all four are outside their DMMA stream at once).
template <int TN>
struct WsCfg {
    static constexpr int WN = BN / (TN * 8);      // warps along the candidate dimension
    static constexpr int NW = 4 * WN;             // warps per CTA
    static constexpr int NT = NW * 32;
};

template <int TN>
__device__ __forceinline__ void load_frags_ws(double (&af)[4], double (&bf)[TN], const double* pa, const double* pb, int ks) {
#pragma unroll
    for (int t2 = 0; t2 < 2; ++t2) {
        const double2 v = *reinterpret_cast<const double2*>(pa + ks * 4 * WS_LD + t2 * 16);
        af[t2 * 2] = v.x;
        af[t2 * 2 + 1] = v.y;
    }
#pragma unroll
    for (int u2 = 0; u2 < TN / 2; ++u2) {
        const double2 v = *reinterpret_cast<const double2*>(pb + ks * 4 * WS_LD + u2 * 16);
        bf[u2 * 2] = v.x;
        bf[u2 * 2 + 1] = v.y;
    }
}

template <int FAM, int TN>
__global__ void __launch_bounds__(WsCfg<TN>::NT, 1)
    ivar_ws_kernel(const __grid_constant__ CoreArgs a, const __grid_constant__ KParams kp) {
    constexpr int BM = WS_BM, LD = WS_LD, NW = WsCfg<TN>::NW, NT = WsCfg<TN>::NT, WCOLS = TN * 8;
    extern __shared__ __align__(16) double smem[];
    double* s_alpha = smem + WS_STAGES * WS_STAGE;  // [WS_STAGES][BM], travels with the prologue chunk
    double* s_beta = s_alpha + WS_STAGES * BM;
    double* s_red = s_beta + BN;
    uint64_t* full = reinterpret_cast<uint64_t*>(s_red + 4 * BN);
    uint64_t* empty = full + WS_STAGES;
    double* s_tab = reinterpret_cast<double*>(empty + WS_STAGES);  // signal * 2^(j/256)

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    // Rasterisation for L2 reuse: the 1-D grid is walked as  j-group (one wave of candidate tiles) -> M-split -> tile,
    // so the waves that share a group's W_C tiles (re-read once per split) run back to back while every wave streams one
    // W_M slice; DRAM traffic ~ (#groups) x (|W_M| + |W_C group|) instead of (#waves) x (|W_M|/splits + |W_C group|).
    const int64_t jtiles = (a.J + BN - 1) / BN;
    const int64_t gw = a.upper_only;  // group width in tiles (reused field: CTAs per wave), >= 1
    const int64_t per_group = gw * a.ldo_splits;
    const int64_t jg = blockIdx.x / per_group;
    const int64_t rem = blockIdx.x - jg * per_group;
    const int64_t gsz = (jtiles - jg * gw) < gw ? (jtiles - jg * gw) : gw;  // tiles in this (possibly last, short) group
    const int64_t split = rem / gsz;
    const int64_t jt = jg * gw + rem % gsz;
    if (split >= a.ldo_splits) return;  // padding CTAs of a short last group
    const int64_t j0 = jt * BN;
    const int64_t itiles = (a.I + BM - 1) / BM;
    const int64_t it_begin = split * a.tiles_per_cta;
    int64_t it_end = it_begin + a.tiles_per_cta;
    if (it_end > itiles) it_end = itiles;
    const int ntiles = it_end > it_begin ? (int)(it_end - it_begin) : 0;
    const int kch = (a.K + WS_BK - 1) / WS_BK;
    const int T = 1 + kch;
    const int G = ntiles * T;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < WS_STAGES; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, NW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (tid < BN) s_beta[tid] = (j0 + tid < a.J) ? a.Bs[j0 + tid] : 0.0;
    if (tid < 256) s_tab[tid] = kp.signal * gpx_exp2_tab[tid];
    __syncthreads();

    double rs[2] = {0.0, 0.0};
    const int wm = warp & 3, wn = warp >> 2;
    const int g4 = lane >> 2, q4 = lane & 3;

    // ---- producer step, executed by one warp per chunk; all addresses are warp-uniform so that the bulk
    //      copies issue back to back from one lane (UBLKCP takes uniform registers) -----------------------
    auto produce = [&](int c) {
        if (c >= G) return;
        const int s = c % WS_STAGES;
        const unsigned int ph = (unsigned int)(c / WS_STAGES) & 1u;
        mbar_wait(empty + s, ph ^ 1u);
        const int ptl = c / T, pch = c - ptl * T;
        const int64_t i0 = (it_begin + ptl) * BM;
        const double *srcA, *srcB;
        int krows, ksteps;
        if (pch == 0) {
            srcA = a.Ap + i0;
            srcB = a.Bp + j0;
            krows = a.dpad;
            ksteps = a.dpad >> 2;
        } else {
            const int kc = pch - 1;
            srcA = a.A + (int64_t)kc * WS_BK * a.lda + i0;
            srcB = a.B + (int64_t)kc * WS_BK * a.ldb + j0;
            const int rem = a.K - kc * WS_BK;
            krows = rem < WS_BK ? rem : WS_BK;
            ksteps = rem >= WS_BK ? WS_KSTEPS : ((rem + 3) >> 2);
        }
        double* stA = smem + s * WS_STAGE;
        double* stB = stA + WS_BK * LD;
        if (krows < ksteps * 4) {
            // K tail: rows the DMMAs will read but the operand does not have -> explicit zeros (lane -> column pair)
            for (int r = krows; r < ksteps * 4; ++r) {
                for (int cc = lane * 2; cc < BM; cc += 64) {
                    *reinterpret_cast<double2*>(stA + r * LD + cc) = make_double2(0.0, 0.0);
                    *reinterpret_cast<double2*>(stB + r * LD + cc) = make_double2(0.0, 0.0);
                }
            }
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");  // generic writes before later TMA writes here
            __syncwarp();
        }
        if (lane == 0) {
            const unsigned int bytes = (unsigned int)krows * (BM + BN) * 8u + (pch == 0 ? BM * 8u : 0u);
            mbar_arrive_expect_tx(full + s, bytes);
#pragma unroll 4
            for (int r = 0; r < krows; ++r) {
                bulk_g2s(stA + r * LD, srcA + (int64_t)r * a.lda, BM * 8u, full + s);
                bulk_g2s(stB + r * LD, srcB + (int64_t)r * a.ldb, BN * 8u, full + s);
            }
            if (pch == 0) bulk_g2s(s_alpha + s * BM, a.As + i0, BM * 8u, full + s);
        }
        __syncwarp();
    };
    if (warp < WS_AHEAD) produce(warp);

    {
        double acc[4][TN][2];
        double af0[4], bf0[TN];
        const int fa = q4 * LD + wm * 32 + g4 * 2;
        const int fb = WS_BK * LD + q4 * LD + wn * WCOLS + g4 * 2;
        int tl = 0, ch = 0;
        for (int g = 0; g < G; ++g) {
            const int s = g % WS_STAGES;
            const unsigned int ph = (unsigned int)(g / WS_STAGES) & 1u;
            const int64_t i0 = (it_begin + tl) * BM;
            if (warp == (g & (NW - 1))) produce(g + WS_AHEAD);
            if (ch == 0) {
#pragma unroll
                for (int t = 0; t < 4; ++t)
#pragma unroll
                    for (int u = 0; u < TN; ++u) acc[t][u][0] = acc[t][u][1] = 0.0;
            }
            mbar_wait(full + s, ph);
            const double* pa = smem + s * WS_STAGE + fa;
            const double* pb = smem + s * WS_STAGE + fb;
            int ksteps;
            if (ch == 0) {
                ksteps = a.dpad >> 2;
            } else {
                const int rem = a.K - (ch - 1) * WS_BK;
                ksteps = rem >= WS_BK ? WS_KSTEPS : ((rem + 3) >> 2);
            }
            // single-buffered fragments: the other warps of the sub-partition cover the LDS latency (double
            // buffering measured identical: 34.44 vs 34.43 TFLOP/s at n = 4095)
#pragma unroll 1
            for (int ks = 0; ks < ksteps; ++ks) {
                load_frags_ws<TN>(af0, bf0, pa, pb, ks);
#pragma unroll
                for (int t = 0; t < 4; ++t)
#pragma unroll
                    for (int u = 0; u < TN; ++u) dmma(acc[t][u][0], acc[t][u][1], af0[t], bf0[u]);
            }
            if (ch == 0) {
                // covariance from the expanded form; accumulators become -k so that the main loop yields T - k
                const double* sal = s_alpha + s * BM;
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const double al = sal[wm * 32 + (t >> 1) * 16 + g4 * 2 + (t & 1)];
#pragma unroll
                    for (int u = 0; u < TN; ++u)
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const double be = s_beta[wn * WCOLS + (u >> 1) * 16 + (q4 * 2 + e) * 2 + (u & 1)];
                            acc[t][u][e] = -kexpand_tab<FAM>(acc[t][u][e] + al + be, kp, s_tab);
                        }
                }
            }
            // this warp is done with stage s (all its LDS results have been consumed by issued DMMAs)
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + s);

            if (ch == T - 1) {
                const bool fullt = (i0 + BM <= a.I);
                double p[TN][2];
#pragma unroll
                for (int u = 0; u < TN; ++u) p[u][0] = p[u][1] = 0.0;
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const int64_t i = i0 + wm * 32 + (t >> 1) * 16 + g4 * 2 + (t & 1);
                    const bool ok = fullt || (i < a.I);
#pragma unroll
                    for (int u = 0; u < TN; ++u)
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const double v = ok ? acc[t][u][e] : 0.0;
                            p[u][e] = fma(v, v, p[u][e]);
                        }
                }
                // butterfly reduce-scatter over the 8 lanes that share q4
                const bool b4 = (lane & 16) != 0, b3 = (lane & 8) != 0, b2 = (lane & 4) != 0;
                if (TN == 8) {
                    // lane g4 ends up owning u = g4 (both e)
                    double h[4][2], q[2][2];
#pragma unroll
                    for (int u = 0; u < 4; ++u)
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const double keep = b4 ? p[(u + 4) % TN][e] : p[u][e];
                            const double send = b4 ? p[u][e] : p[(u + 4) % TN][e];
                            h[u][e] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
                        }
#pragma unroll
                    for (int u = 0; u < 2; ++u)
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const double keep = b3 ? h[u + 2][e] : h[u][e];
                            const double send = b3 ? h[u][e] : h[u + 2][e];
                            q[u][e] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
                        }
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const double keep = b2 ? q[1][e] : q[0][e];
                        const double send = b2 ? q[0][e] : q[1][e];
                        rs[e] += keep + __shfl_xor_sync(0xffffffffu, send, 4);
                    }
                } else {
                    // TN == 4: lane g4 ends up owning u = g4 >> 1, e = g4 & 1 (one value)
                    double h[2][2], q[2];
#pragma unroll
                    for (int u = 0; u < 2; ++u)
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const double keep = b4 ? p[(u + 2) % TN][e] : p[u][e];
                            const double send = b4 ? p[u][e] : p[(u + 2) % TN][e];
                            h[u][e] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
                        }
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const double keep = b3 ? h[1][e] : h[0][e];
                        const double send = b3 ? h[0][e] : h[1][e];
                        q[e] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
                    }
                    {
                        const double keep = b2 ? q[1] : q[0];
                        const double send = b2 ? q[0] : q[1];
                        rs[0] += keep + __shfl_xor_sync(0xffffffffu, send, 4);
                    }
                }
                ch = 0;
                ++tl;
            } else {
                ++ch;
            }
        }
    }

    __syncthreads();
    if (TN == 8) {
#pragma unroll
        for (int e = 0; e < 2; ++e) s_red[wm * BN + wn * WCOLS + (g4 >> 1) * 16 + (q4 * 2 + e) * 2 + (g4 & 1)] = rs[e];
    } else {
        const int u = g4 >> 1, e = g4 & 1;
        s_red[wm * BN + wn * WCOLS + (u >> 1) * 16 + (q4 * 2 + e) * 2 + (u & 1)] = rs[0];
    }
    __syncthreads();
    if (tid < BN && j0 + tid < a.J) {
        const double r = ((s_red[tid] + s_red[BN + tid]) + s_red[2 * BN + tid]) + s_red[3 * BN + tid];
        a.out[split * a.ldo + j0 + tid] = r;
    }
}

bool ivar_use_tma() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("GPX_IVAR_TMA");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v == 1;
}
// ---------------------------------------------------------------------------------------------
// TMA + mbarrier variant of the plain update  C[i,j] -= sum_k A[k,i] B[k,j]  for callers that GUARANTEE fully padded
// operands (every 128-wide tile of A and B readable and finite; the distributed MI set-up does).  Same ring, fragment
// layout and rotating producer as ivar_ws_kernel, one 128x128 tile per CTA, no prologue.
// ---------------------------------------------------------------------------------------------
constexpr int SUBWS_SMEM_DOUBLES = WS_STAGES * WS_STAGE + 2 * WS_STAGES;
constexpr size_t SUBWS_SMEM_BYTES = (size_t)SUBWS_SMEM_DOUBLES * sizeof(double);

__global__ void __launch_bounds__(256, 1) sub_ws_kernel(const __grid_constant__ CoreArgs a) {
    constexpr int BM = WS_BM, LD = WS_LD, NW = 8;
    extern __shared__ __align__(16) double smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + WS_STAGES * WS_STAGE);
    uint64_t* empty = full + WS_STAGES;

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int64_t j0 = (int64_t)blockIdx.x * BN;
    const int64_t i0 = (int64_t)blockIdx.y * BM;
    if (a.upper_only && i0 >= j0 + BN) return;  // tile strictly below the diagonal of a symmetric update
    const int G = (a.K + WS_BK - 1) / WS_BK;     // chunks

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < WS_STAGES; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, NW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();

    const int wm = warp & 3, wn = warp >> 2;
    const int g4 = lane >> 2, q4 = lane & 3;

    auto produce = [&](int c) {
        if (c >= G) return;
        const int s = c % WS_STAGES;
        const unsigned int ph = (unsigned int)(c / WS_STAGES) & 1u;
        mbar_wait(empty + s, ph ^ 1u);
        const double* srcA = a.A + (int64_t)c * WS_BK * a.lda + i0;
        const double* srcB = a.B + (int64_t)c * WS_BK * a.ldb + j0;
        const int rem = a.K - c * WS_BK;
        const int krows = rem < WS_BK ? rem : WS_BK;
        const int ksteps = rem >= WS_BK ? WS_KSTEPS : ((rem + 3) >> 2);
        double* stA = smem + s * WS_STAGE;
        double* stB = stA + WS_BK * LD;
        if (krows < ksteps * 4) {
            for (int r = krows; r < ksteps * 4; ++r) {
                for (int cc = lane * 2; cc < BM; cc += 64) {
                    *reinterpret_cast<double2*>(stA + r * LD + cc) = make_double2(0.0, 0.0);
                    *reinterpret_cast<double2*>(stB + r * LD + cc) = make_double2(0.0, 0.0);
                }
            }
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
            __syncwarp();
        }
        if (lane == 0) {
            mbar_arrive_expect_tx(full + s, (unsigned int)krows * (BM + BN) * 8u);
#pragma unroll 4
            for (int r = 0; r < krows; ++r) {
                bulk_g2s(stA + r * LD, srcA + (int64_t)r * a.lda, BM * 8u, full + s);
                bulk_g2s(stB + r * LD, srcB + (int64_t)r * a.ldb, BN * 8u, full + s);
            }
        }
        __syncwarp();
    };
    if (warp < WS_AHEAD) produce(warp);

    double acc[4][8][2];
#pragma unroll
    for (int t = 0; t < 4; ++t)
#pragma unroll
        for (int u = 0; u < 8; ++u) acc[t][u][0] = acc[t][u][1] = 0.0;
    double af0[4], bf0[8];
    const int fa = q4 * LD + wm * 32 + g4 * 2;
    const int fb = WS_BK * LD + q4 * LD + wn * 64 + g4 * 2;
    for (int g = 0; g < G; ++g) {
        const int s = g % WS_STAGES;
        const unsigned int ph = (unsigned int)(g / WS_STAGES) & 1u;
        if (warp == (g & (NW - 1))) produce(g + WS_AHEAD);
        mbar_wait(full + s, ph);
        const double* pa = smem + s * WS_STAGE + fa;
        const double* pb = smem + s * WS_STAGE + fb;
        const int rem = a.K - g * WS_BK;
        const int ksteps = rem >= WS_BK ? WS_KSTEPS : ((rem + 3) >> 2);
#pragma unroll 1
        for (int ks = 0; ks < ksteps; ++ks) {
            load_frags_ws<8>(af0, bf0, pa, pb, ks);
#pragma unroll
            for (int t = 0; t < 4; ++t)
#pragma unroll
                for (int u = 0; u < 8; ++u) dmma(acc[t][u][0], acc[t][u][1], af0[t], bf0[u]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + s);
    }
    // C -= acc, two adjacent columns per 16-byte access
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

int ivar_tn() {
    static int v = 0;
    if (v == 0) {
        const char* e = getenv("GPX_IVAR_TN");
        v = (e && e[0] == '8') ? 8 : ((e && e[0] == '4') ? 4 : GPX_DEFAULT_IVAR_TN);
    }
    return v;
}

template <int FAM, int TN>
int launch_ivar_ws_tn(const CoreArgs& a, const KParams& kp, dim3 grid, cudaStream_t st) {
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(ivar_ws_kernel<FAM, TN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WS_SMEM_BYTES);
        if (e != cudaSuccess) {
            gpx_set_error("ivar_ws: cannot opt in to %zu bytes of shared memory: %s", WS_SMEM_BYTES, cudaGetErrorString(e));
            return (int)e;
        }
        configured = true;
    }
    ivar_ws_kernel<FAM, TN><<<grid, WsCfg<TN>::NT, WS_SMEM_BYTES, st>>>(a, kp);
    return gpx_check_launch("ivar_ws");
}
template <int FAM>
int launch_ivar_ws(const CoreArgs& a, const KParams& kp, dim3 grid, cudaStream_t st) {
    return ivar_tn() == 4 ? launch_ivar_ws_tn<FAM, 4>(a, kp, grid, st) : launch_ivar_ws_tn<FAM, 8>(a, kp, grid, st);
}

// tile shape used by every launch: GPX_WM=2 (64x128, 2 CTAs/SM) or 4 (128x128, 1 CTA/SM)
int core_wm() {
    static int wm = 0;
    if (wm == 0) {
        const char* e = getenv("GPX_WM");
        wm = (e && e[0] == '4') ? 4 : ((e && e[0] == '2') ? 2 : GPX_DEFAULT_WM);
    }
    return wm;
}

template <int FAM, int EPI, bool PRO, int WM>
int launch_core_wm(const CoreArgs& a, const KParams& kp, dim3 grid, cudaStream_t st) {
    using C = Cfg<WM>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(dmma_core_kernel<FAM, EPI, PRO, WM>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)C::SMEM_BYTES);
        if (e != cudaSuccess) {
            gpx_set_error("dmma core: cannot opt in to %zu bytes of shared memory: %s", C::SMEM_BYTES, cudaGetErrorString(e));
            return (int)e;
        }
        configured = true;
    }
    dmma_core_kernel<FAM, EPI, PRO, WM><<<grid, C::NT, C::SMEM_BYTES, st>>>(a, kp);
    return gpx_check_launch("dmma core");
}

template <int FAM, int EPI, bool PRO>
int launch_core(const CoreArgs& a, const KParams& kp, int64_t jt, int64_t it_or_splits, cudaStream_t st) {
    dim3 grid((unsigned)jt, (unsigned)it_or_splits);
    if (core_wm() == 4) return launch_core_wm<FAM, EPI, PRO, 4>(a, kp, grid, st);
    return launch_core_wm<FAM, EPI, PRO, 2>(a, kp, grid, st);
}

int core_bm() { return core_wm() * 32; }


int check_operand(const double* p, int64_t ld, const char* name) {
    if (!gpx_aligned16(p) || (ld & 1)) {
        gpx_set_error("dmma core: operand %s must be 16-byte aligned with an even leading dimension", name);
        return GPX_EALIGN;
    }
    return GPX_OK;
}

}  // namespace

// number of i-splits for the IVAR grid: fill the machine in whole waves
static int ivar_splits_for(gpx_handle h, int64_t M, int64_t C, int bm) {
    const int64_t jt = (C + BN - 1) / BN;
    const int64_t itl = (M + bm - 1) / bm;
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
int gpx_ivar_splits(gpx_handle, int64_t, int64_t) { return 32; }  // workspace bound: never more than 32 splits

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
    // the TMA path reads whole 128-wide tiles: operands must be padded to full tiles (the engines do that);
    // anything else goes through the predicated cp.async core
    const bool padded = (ldm % WS_BM) == 0 && (ldc % BN) == 0 && ldm >= (M + WS_BM - 1) / WS_BM * WS_BM &&
                        ldc >= (C + BN - 1) / BN * BN;
    const bool tma = ivar_use_tma() && padded;
    const int bm = tma ? WS_BM : core_bm();
    const int splits = ivar_splits_for(h, M, C, bm);
    const int64_t itl = (M + bm - 1) / bm;
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
    a.ldo_splits = 0;
    *nsplit_out = splits;
    if (tma) {
        const int64_t jt = (C + BN - 1) / BN;
        static int gmul = 0;
        if (gmul == 0) {
            const char* e = getenv("GPX_IVAR_GROUP");
            gmul = (e && e[0] >= '1' && e[0] <= '8') ? (e[0] - '0') : GPX_DEFAULT_IVAR_GROUP;
        }
        // one CTA per SM: a group is `gmul` waves of candidate tiles (their W_C tiles must stay L2-resident)
        const int64_t gw = (int64_t)gmul * (h->sm_count > 0 ? h->sm_count : 148);
        const int64_t groups = (jt + gw - 1) / gw;
        a.upper_only = (int)gw;
        a.ldo_splits = splits;
        // every group gets gw*splits CTA slots; the short last group leaves some idle (they exit at once)
        dim3 grid((unsigned)(groups * gw * splits), 1u);
        GPX_DISPATCH_FAMILY(h->kp.family, rc = (launch_ivar_ws<FAM>(a, h->kp, grid, st)));
        return rc;
    }
    GPX_DISPATCH_FAMILY(h->kp.family, rc = (launch_core<FAM, EPI_IVAR, true>(a, h->kp, (C + BN - 1) / BN, splits, st)));
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
    a.ldo_splits = 0;
    GPX_DISPATCH_FAMILY(h->kp.family,
                        rc = (launch_core<FAM, EPI_STORE, true>(a, h->kp, (J + BN - 1) / BN, (I + core_bm() - 1) / core_bm(), st)));
    return rc;
}

extern "C" int gpx_dgemm_tn_sub(gpx_handle h, const double* A, int64_t lda, const double* B, int64_t ldb, double* C,
                                int64_t ldc, int64_t I, int64_t J, int64_t K, int upper_only, void* stream) {
    GPX_REQUIRE(h != nullptr, GPX_EINVAL, "handle is NULL");
    GPX_REQUIRE(I >= 0 && J >= 0 && K >= 0, GPX_EINVAL, "negative size");
    if (I == 0 || J == 0 || K == 0) return GPX_OK;
    GPX_REQUIRE(A && B && C, GPX_EINVAL, "NULL pointer");
    GPX_REQUIRE((I + 63) / 64 <= 65535, GPX_ESIZE, "I too large for one launch");
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
    a.ldo_splits = 0;
    KParams kp = h->kp;
    return launch_core<GPX_SE, EPI_SUB, false>(a, kp, (J + BN - 1) / BN, (I + core_bm() - 1) / core_bm(), (cudaStream_t)stream);
}

// Same update for operands the CALLER guarantees to be fully padded (every 128-wide tile of A's I columns and B's J
// columns readable and finite, lda/ldb multiples of 128 or at least covering the rounded-up extents): TMA + mbarrier kernel.
extern "C" int gpx_dgemm_tn_sub_padded(gpx_handle h, const double* A, int64_t lda, const double* B, int64_t ldb, double* C,
                                       int64_t ldc, int64_t I, int64_t J, int64_t K, int upper_only, void* stream) {
    GPX_REQUIRE(h != nullptr, GPX_EINVAL, "handle is NULL");
    GPX_REQUIRE(I >= 0 && J >= 0 && K >= 0, GPX_EINVAL, "negative size");
    if (I == 0 || J == 0 || K == 0) return GPX_OK;
    GPX_REQUIRE(A && B && C, GPX_EINVAL, "NULL pointer");
    GPX_REQUIRE((I + WS_BM - 1) / WS_BM <= 65535, GPX_ESIZE, "I too large for one launch");
    int rc;
    if ((rc = check_operand(A, lda, "A"))) return rc;
    if ((rc = check_operand(B, ldb, "B"))) return rc;
    if ((rc = check_operand(C, ldc, "C"))) return rc;
    GPX_REQUIRE(lda >= (I + WS_BM - 1) / WS_BM * WS_BM && ldb >= (J + BN - 1) / BN * BN, GPX_EALIGN,
                "padded variant needs leading dimensions that cover whole 128-wide tiles");
    static const bool use_tma = []() {
        const char* e = getenv("GPX_SUB_TMA");
        return !(e && e[0] == '0');
    }();
    if (!use_tma) return gpx_dgemm_tn_sub(h, A, lda, B, ldb, C, ldc, I, J, K, upper_only, stream);
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
    a.ldo_splits = 0;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(sub_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SUBWS_SMEM_BYTES);
        if (e != cudaSuccess) {
            gpx_set_error("sub_ws: cannot opt in to %zu bytes of shared memory: %s", SUBWS_SMEM_BYTES, cudaGetErrorString(e));
            return (int)e;
        }
        configured = true;
    }
    dim3 grid((unsigned)((J + BN - 1) / BN), (unsigned)((I + WS_BM - 1) / WS_BM));
    sub_ws_kernel<<<grid, 256, SUBWS_SMEM_BYTES, (cudaStream_t)stream>>>(a);
    return gpx_check_launch("gpx_dgemm_tn_sub_padded");
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

// IVAR scores from per-segment column sums of squares (resident-covariance mode): same finalisation + arg-min
extern "C" int gpx_score_ivar_partials(gpx_handle h, const double* partial, int nseg, int64_t ldp, const double* varM, int64_t M,
                                       const double* varC, int64_t C, double noise, double zero_tol, const uint8_t* mask,
                                       double* score_out, double* best, int64_t* idx, void* stream) {
    GPX_REQUIRE(h && partial && varM && varC && score_out && best && idx && nseg >= 1 && M >= 1 && C >= 1, GPX_EINVAL,
                "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    int rc = gpx_sum_impl(h, varM, M, h->scal, st);
    if (rc) return rc;
    ivar_finalize_kernel<<<(unsigned)((C + 255) / 256), 256, 0, st>>>(partial, nseg, ldp, varC, h->scal, M, C, noise, zero_tol,
                                                                     score_out);
    rc = gpx_check_launch("gpx_score_ivar_partials finalize");
    if (rc) return rc;
    return gpx_argreduce_impl(h, score_out, nullptr, mask, C, 1, best, idx, st);
}

// cov[m,c] = k(m,c) - sum_{i<n} Wm[i,m] Wc[i,c] for a GIVEN design (DMMA contraction with the Gram prologue, stored)
extern "C" int gpx_cov_from_factors(gpx_handle h, const double* Wm, int64_t ldm, const double* Ma_rows, const double* Ma_scal,
                                    int64_t M, const double* Wc, int64_t ldc, const double* Cb_rows, const double* Cb_scal,
                                    int64_t C, int64_t n, double* cov, int64_t ldcov, void* stream) {
    GPX_NEED_KERNEL(h);
    GPX_REQUIRE(M >= 1 && C >= 1 && n >= 0 && cov && Ma_rows && Ma_scal && Cb_rows && Cb_scal, GPX_EINVAL, "bad arguments");
    GPX_REQUIRE((M + 63) / 64 <= 65535, GPX_ESIZE, "M too large for one launch");
    return gpx_launch_core_store(h, Wm, ldm, Ma_rows, Ma_scal, M, Wc, ldc, Cb_rows, Cb_scal, C, n, cov, ldcov,
                                 (cudaStream_t)stream);
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
