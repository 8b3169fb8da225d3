// the device pieces behind
//   f2  the heteroscedastic posterior-variance derivative            gp.py:282-341 (noiseFunc branch)
//   f3  the hyper-parameter gradient of the marginal log-likelihood  gp.py:447-468, kernels.py:125-144
//   f4  the matrix-free Gram x vector product of the Nystrom eigen-solver and the FITC sparse precision
//                                                                    gp_kernel_utilities.py:107-142, :70-104, gp.py:182-208
// All of them are thin, HBM- or FP64-issue-bound companions of K1 (gpx_gram): no new factorisation code.
#include <math.h>

#include "gpx_common.cuh"

// ---------------------------------------------------------------------------------------------
// f4  out[i] = sum_j k(x_i, y_j) b[j]      covTimesV (gp_kernel_utilities.py:107-142) without the n x n matrix.
// One thread per output row keeps its coordinates in registers; the y tile (coordinates + b) is staged in shared
// memory and read by broadcast.  n*m covariance evaluations, 8*(n+m)*d bytes: FP64-issue bound (table exp).
// The j range is split over blockIdx.y; partial sums are added in split order by a second pass (deterministic).
// ---------------------------------------------------------------------------------------------
#define MV_TILE 128
#define MV_SPLITS_MAX 64

template <int FAM, int D>
__global__ void __launch_bounds__(128, 1) gram_matvec_kernel(const __grid_constant__ KParams kp, const double* __restrict__ X,
                                                           int64_t n, int64_t ldx, const double* __restrict__ Y, int64_t m,
                                                           int64_t ldy, const double* __restrict__ b, int64_t cols_per_split,
                                                           double* __restrict__ part, int64_t ldp) {
    __shared__ double sy[D][MV_TILE];
    __shared__ double sb[MV_TILE];
    __shared__ double s_tab[256];
    for (int t = threadIdx.x; t < 256; t += 128) s_tab[t] = kp.signal * gpx_exp2_tab[t];
    const int64_t i = (int64_t)blockIdx.x * 128 + threadIdx.x;
    double x[D];
#pragma unroll
    for (int q = 0; q < D; ++q) x[q] = (i < n) ? X[q * ldx + i] : 0.0;
    const int64_t j_begin = (int64_t)blockIdx.y * cols_per_split;
    int64_t j_end = j_begin + cols_per_split;
    if (j_end > m) j_end = m;
    double acc = 0.0;
    for (int64_t j0 = j_begin; j0 < j_end; j0 += MV_TILE) {
        __syncthreads();
        const int64_t j = j0 + threadIdx.x;
        const bool live = j < j_end;
#pragma unroll
        for (int q = 0; q < D; ++q) sy[q][threadIdx.x] = live ? Y[q * ldy + j] : 0.0;
        sb[threadIdx.x] = live ? b[j] : 0.0;
        __syncthreads();
        const int cnt = (j_end - j0) < MV_TILE ? (int)(j_end - j0) : MV_TILE;
#pragma unroll 4
        for (int t = 0; t < cnt; ++t) {
            double a = 0.0;
#pragma unroll
            for (int q = 0; q < D; ++q) kacc_dim<FAM>(a, kp, q, x[q], sy[q][t]);
            acc = fma(kfinish_tab<FAM>(a, kp, s_tab), sb[t], acc);
        }
    }
    if (i < n) part[(int64_t)blockIdx.y * ldp + i] = acc;
}

__global__ void __launch_bounds__(256) sum_splits_kernel(const double* __restrict__ part, int nsplit, int64_t n, int64_t ldp,
                                                          double* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    double s = 0.0;
    for (int c = 0; c < nsplit; ++c) s += part[(int64_t)c * ldp + i];
    out[i] = s;
}

extern "C" int64_t gpx_gram_matvec_workspace(int64_t n) { return n <= 0 ? 0 : (int64_t)MV_SPLITS_MAX * ((n + 1) & ~(int64_t)1); }

extern "C" int gpx_gram_matvec(gpx_handle h, const double* X, int64_t n, int64_t ldx, const double* Y, int64_t m, int64_t ldy,
                               const double* b, double* workspace, double* out, void* stream) {
    GPX_NEED_KERNEL(h);
    GPX_REQUIRE(n >= 0 && m >= 0, GPX_EINVAL, "negative size");
    if (n == 0) return GPX_OK;
    GPX_REQUIRE(X && out && workspace && (m == 0 || (Y && b)), GPX_EINVAL, "NULL pointer");
    cudaStream_t st = (cudaStream_t)stream;
    // enough j-splits to fill the machine when there are few row blocks, whole tiles per split
    const int64_t row_blocks = (n + 127) / 128;
    const int sms = h->sm_count > 0 ? h->sm_count : 148;
    int64_t splits = (4 * sms + row_blocks - 1) / row_blocks;
    const int64_t tiles = (m + MV_TILE - 1) / MV_TILE;
    if (splits > tiles) splits = tiles;
    if (splits > MV_SPLITS_MAX) splits = MV_SPLITS_MAX;
    if (splits < 1) splits = 1;
    const int64_t cols_per_split = ((tiles + splits - 1) / splits) * MV_TILE;
    splits = m > 0 ? (m + cols_per_split - 1) / cols_per_split : 1;
    const int64_t ldp = (n + 1) & ~(int64_t)1;
    dim3 grid((unsigned)row_blocks, (unsigned)splits);
    GPX_DISPATCH_FAMILY(h->kp.family, GPX_DISPATCH_DIM(h->kp.d, (gram_matvec_kernel<FAM, D><<<grid, 128, 0, st>>>(
                                                                    h->kp, X, n, ldx, Y, m, ldy, b, cols_per_split, workspace, ldp))));
    int rc = gpx_check_launch("gpx_gram_matvec");
    if (rc) return rc;
    sum_splits_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(workspace, (int)splits, n, ldp, out);
    return gpx_check_launch("gpx_gram_matvec reduce");
}

// ---------------------------------------------------------------------------------------------
// small dense helpers (FITC Woodbury algebra, gradient assembly): one pass, coalesced
// ---------------------------------------------------------------------------------------------
// out[i,j] = A[i,j] * (r ? r[i] : 1) * (c ? c[j] : 1) * scale        (out may alias A)
__global__ void __launch_bounds__(256) scale_rc_kernel(const double* __restrict__ A, int64_t rows, int64_t cols, int64_t lda,
                                                        const double* __restrict__ r, const double* __restrict__ c, double scale,
                                                        double* __restrict__ out, int64_t ldo) {
    const int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const int64_t i = blockIdx.y;
    if (j >= cols || i >= rows) return;
    double v = A[i * lda + j] * scale;
    if (r) v *= r[i];
    if (c) v *= c[j];
    out[i * ldo + j] = v;
}

extern "C" int gpx_scale_rows_cols(gpx_handle h, const double* A, int64_t rows, int64_t cols, int64_t lda, const double* r,
                                   const double* c, double scale, double* out, int64_t ldo, void* stream) {
    GPX_REQUIRE(h && rows >= 0 && cols >= 0, GPX_EINVAL, "bad arguments");
    if (rows == 0 || cols == 0) return GPX_OK;
    GPX_REQUIRE(A && out && lda >= cols && ldo >= cols && rows <= 65535, GPX_EINVAL, "bad arguments");
    dim3 grid((unsigned)((cols + 255) / 256), (unsigned)rows);
    scale_rc_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(A, rows, cols, lda, r, c, scale, out, ldo);
    return gpx_check_launch("gpx_scale_rows_cols");
}

// A[i,i] = scale * A[i,i] + (d ? d[i] : shift)
__global__ void __launch_bounds__(256) diag_update_kernel(double* __restrict__ A, int64_t n, int64_t ld, double scale,
                                                           const double* __restrict__ d, double shift) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    A[i * ld + i] = scale * A[i * ld + i] + (d ? d[i] : shift);
}

extern "C" int gpx_diag_update(gpx_handle h, double* A, int64_t n, int64_t ld, double scale, const double* d, double shift,
                               void* stream) {
    GPX_REQUIRE(h && n >= 0, GPX_EINVAL, "bad arguments");
    if (n == 0) return GPX_OK;
    GPX_REQUIRE(A && ld >= n, GPX_EINVAL, "bad arguments");
    diag_update_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(A, n, ld, scale, d, shift);
    return gpx_check_launch("gpx_diag_update");
}

// y[i] = alpha * x[i] + beta * y[i]
__global__ void __launch_bounds__(256) axpby_kernel(int64_t n, double alpha, const double* __restrict__ x, double beta,
                                                     double* __restrict__ y) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i < n) y[i] = alpha * x[i] + (beta == 0.0 ? 0.0 : beta * y[i]);
}

extern "C" int gpx_axpby(gpx_handle h, int64_t n, double alpha, const double* x, double beta, double* y, void* stream) {
    GPX_REQUIRE(h && n >= 0, GPX_EINVAL, "bad arguments");
    if (n == 0) return GPX_OK;
    GPX_REQUIRE(x && y, GPX_EINVAL, "NULL pointer");
    axpby_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, alpha, x, beta, y);
    return gpx_check_launch("gpx_axpby");
}

// out[j] = (base ? base[j] : 0) - sum_i A[i,j] * B[i,j]       k^T P k with Z = P k materialised (FITC posterior variance)
__global__ void __launch_bounds__(128) coldot_kernel(const double* __restrict__ A, const double* __restrict__ B, int64_t n,
                                                      int64_t ncols, int64_t ld, const double* __restrict__ base,
                                                      double* __restrict__ out) {
    const int64_t j = (int64_t)blockIdx.x * 128 + threadIdx.x;
    if (j >= ncols) return;
    double a = 0.0;
    for (int64_t i = 0; i < n; ++i) a = fma(A[i * ld + j], B[i * ld + j], a);
    out[j] = (base ? base[j] : 0.0) - a;
}

extern "C" int gpx_coldot(gpx_handle h, const double* A, const double* B, int64_t n, int64_t ncols, int64_t ld,
                          const double* base, double* out, void* stream) {
    GPX_REQUIRE(h && out && n >= 0 && ncols >= 0, GPX_EINVAL, "bad arguments");
    if (ncols == 0) return GPX_OK;
    GPX_REQUIRE((A && B) || n == 0, GPX_EINVAL, "NULL pointer");
    coldot_kernel<<<(unsigned)((ncols + 127) / 128), 128, 0, (cudaStream_t)stream>>>(A, B, n, ncols, ld, base, out);
    return gpx_check_launch("gpx_coldot");
}

// ---------------------------------------------------------------------------------------------
// f3  gradient of the marginal log-likelihood with respect to the SE hyper-parameters (gp.py:447-468):
//         out[q] = 1/2 tr( (alpha alpha^T - P) dK/dtheta_q ),   theta = cl_0 .. cl_{d-1}, signalSize, noise
//     dK/dcl_q = k (x_q - y_q)^2 / cl_q^3 ,  dK/dsignalSize = k / signalSize (kernels.py:125-144),  dK/dnoise = I.
//     (The chain factor 2*noise the reference applies to the noise entry, gp.py:463-464, is left to the caller.)
// One thread per matrix element, block partials in a fixed layout, second pass adds them in order.
// ---------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(256) se_loglike_grad_kernel(const __grid_constant__ KParams kp, const double* __restrict__ X,
                                                               int64_t n, int64_t ldx, const double* __restrict__ P,
                                                               int64_t ldp, const double* __restrict__ alpha,
                                                               double* __restrict__ part) {
    __shared__ double sm[8][D + 2];
    const int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const int64_t i = blockIdx.y;
    double g[D + 2];
#pragma unroll
    for (int q = 0; q < D + 2; ++q) g[q] = 0.0;
    if (j < n) {
        const double t = alpha[i] * alpha[j] - P[i * ldp + j];
        double acc = 0.0, df2[D];
#pragma unroll
        for (int q = 0; q < D; ++q) {
            const double df = X[q * ldx + i] - X[q * ldx + j];
            df2[q] = df * df;
            acc = fma(df2[q], kp.a[q], acc);
        }
        const double e = exp(-0.5 * acc);
        const double kv = kp.signal * e;
#pragma unroll
        for (int q = 0; q < D; ++q) g[q] = t * kv * df2[q] * kp.a[q] * sqrt(kp.a[q]);  // a = cl^-2  ->  a^(3/2) = cl^-3
        g[D] = t * e;
        g[D + 1] = (i == j) ? t : 0.0;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < D + 2; ++q) {
        double v = g[q];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if (lane == 0) sm[warp][q] = v;
    }
    __syncthreads();
    if (threadIdx.x < D + 2) {
        double s = 0.0;
        for (int w = 0; w < 8; ++w) s += sm[w][threadIdx.x];
        part[((int64_t)blockIdx.y * gridDim.x + blockIdx.x) * (D + 2) + threadIdx.x] = s;
    }
}

__global__ void __launch_bounds__(256) grad_reduce_kernel(const double* __restrict__ part, int64_t nblocks, int nq,
                                                           double* __restrict__ out) {
    __shared__ double sm[8];
    const int q = blockIdx.x;
    double s = 0.0;
    for (int64_t b = threadIdx.x; b < nblocks; b += 256) s += part[b * nq + q];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += sm[w];
        out[q] = 0.5 * t;
    }
}

extern "C" int64_t gpx_se_loglike_grad_workspace(int64_t n, int d) {
    if (n <= 0 || d <= 0) return 0;
    return ((n + 255) / 256) * n * (int64_t)(d + 2);
}

extern "C" int gpx_se_loglike_grad(gpx_handle h, const double* X, int64_t n, int64_t ldx, const double* P, int64_t ldp,
                                   const double* alpha, double* workspace, double* out, void* stream) {
    GPX_NEED_KERNEL(h);
    GPX_REQUIRE(h->kp.family == GPX_SE, GPX_EINVAL, "derivativeWrtHypParams exists for the squared-exponential family only");
    GPX_REQUIRE(n >= 1 && n <= 65535 && X && P && alpha && workspace && out && ldp >= n, GPX_EINVAL, "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid((unsigned)((n + 255) / 256), (unsigned)n);
    GPX_DISPATCH_DIM(h->kp.d, (se_loglike_grad_kernel<D><<<grid, 256, 0, st>>>(h->kp, X, n, ldx, P, ldp, alpha, workspace)));
    int rc = gpx_check_launch("gpx_se_loglike_grad");
    if (rc) return rc;
    grad_reduce_kernel<<<(unsigned)(h->kp.d + 2), 256, 0, st>>>(workspace, (int64_t)grid.x * grid.y, h->kp.d + 2, out);
    return gpx_check_launch("gpx_se_loglike_grad reduce");
}
