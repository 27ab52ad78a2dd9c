// Device side of the BUILDER: the regularised least-squares decoder solves of nengo's LstsqL2
// (sspslam/networks/pathintegration.py:180-182, binding.py:316-317, slam.py:298-303 -> 509 solves per model, x models for
// trials with their own network seed), batched over ensembles of one shape:
//     sigma = reg * max(A),   G = A^T A + m sigma^2 I,   X = G^-1 A^T Y        (A: m x n activities, Y: m x k targets)
// in float64 like nengo.  This is build-time plumbing, not the stepped hot path: the Gram products are plain library GEMMs
// (cublasDgemmStridedBatched), the factorisation and triangular solves cuSOLVER / cuBLAS batched routines; the only kernels
// of our own are the row-maximum / diagonal-shift passes.  Built as a SEPARATE shared library (libssb_builder.so) so that
// the stepping library does not depend on cuBLAS / cuSOLVER.
#include "../../include/sspslam_b200_builder.h"

#include <cublas_v2.h>
#include <cuda_runtime.h>
#include <cusolverDn.h>

#include <string>
#include <vector>

namespace {

thread_local std::string g_builder_err;

int ssb_builder_fail(int code, const std::string& msg) {
    g_builder_err = msg;
    return code;
}

// one CTA per system: max over the m x n activity matrix, then G[i][i] += m * (reg * max)^2
__global__ void k_builder_shift(const double* __restrict__ A, double* __restrict__ G, long long a_stride, long long g_stride,
                                int m, int n, double reg) {
    __shared__ double red[256];
    const double* a = A + (size_t)blockIdx.x * a_stride;
    double mx = 0.0;
    for (long long i = threadIdx.x; i < (long long)m * n; i += blockDim.x) mx = fmax(mx, a[i]);
    red[threadIdx.x] = mx;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) red[threadIdx.x] = fmax(red[threadIdx.x], red[threadIdx.x + s]);
        __syncthreads();
    }
    const double sigma = reg * red[0];
    double* g = G + (size_t)blockIdx.x * g_stride;
    for (int i = threadIdx.x; i < n; i += blockDim.x) g[(size_t)i * n + i] += (double)m * sigma * sigma;
}

#define BLD_CUDA(expr)                                                                             \
    do {                                                                                           \
        cudaError_t e__ = (expr);                                                                  \
        if (e__ != cudaSuccess) return ssb_builder_fail(-2, std::string(#expr) + ": " + cudaGetErrorString(e__)); \
    } while (0)
#define BLD_LIB(expr, what)                                                                        \
    do {                                                                                           \
        if ((int)(expr) != 0) return ssb_builder_fail(-2, std::string(what) + " failed");          \
    } while (0)

}  // namespace

extern "C" const char* ssb_builder_last_error(void) { return g_builder_err.c_str(); }

extern "C" int ssb_solve_decoders(int device, int n_sys, int m, int n, int k, const double* A, const double* Y, double reg,
                                  double* X) {
    if (n_sys <= 0) return 0;
    if (!A || !Y || !X || m < n || n <= 0 || k <= 0) return ssb_builder_fail(-1, "ssb_solve_decoders: bad arguments (needs m >= n)");
    BLD_CUDA(cudaSetDevice(device));
    cublasHandle_t blas = nullptr;
    cusolverDnHandle_t sol = nullptr;
    BLD_LIB(cublasCreate(&blas), "cublasCreate");
    BLD_LIB(cusolverDnCreate(&sol), "cusolverDnCreate");
    const size_t a_sz = (size_t)m * n, y_sz = (size_t)m * k, g_sz = (size_t)n * n, x_sz = (size_t)n * k;
    // systems are processed in slabs that fit a few GB
    const size_t per_sys = (a_sz + y_sz + g_sz + x_sz) * sizeof(double);
    const int slab = (int)std::max<size_t>(1, std::min<size_t>((size_t)n_sys, (size_t)(4ull << 30) / per_sys));
    double *dA = nullptr, *dY = nullptr, *dG = nullptr, *dX = nullptr;
    double **dGp = nullptr, **dXp = nullptr;
    int* dinfo = nullptr;
    BLD_CUDA(cudaMalloc((void**)&dA, slab * a_sz * sizeof(double)));
    BLD_CUDA(cudaMalloc((void**)&dY, slab * y_sz * sizeof(double)));
    BLD_CUDA(cudaMalloc((void**)&dG, slab * g_sz * sizeof(double)));
    BLD_CUDA(cudaMalloc((void**)&dX, slab * x_sz * sizeof(double)));
    BLD_CUDA(cudaMalloc((void**)&dGp, slab * sizeof(double*)));
    BLD_CUDA(cudaMalloc((void**)&dXp, slab * sizeof(double*)));
    BLD_CUDA(cudaMalloc((void**)&dinfo, slab * sizeof(int)));
    std::vector<double*> hGp(slab), hXp(slab);
    for (int i = 0; i < slab; ++i) {
        hGp[i] = dG + (size_t)i * g_sz;
        hXp[i] = dX + (size_t)i * x_sz;
    }
    BLD_CUDA(cudaMemcpy(dGp, hGp.data(), slab * sizeof(double*), cudaMemcpyHostToDevice));
    BLD_CUDA(cudaMemcpy(dXp, hXp.data(), slab * sizeof(double*), cudaMemcpyHostToDevice));
    const double one = 1.0, zero = 0.0;
    int rc = 0;
    for (int s0 = 0; s0 < n_sys && rc == 0; s0 += slab) {
        const int ns = std::min(slab, n_sys - s0);
        BLD_CUDA(cudaMemcpy(dA, A + (size_t)s0 * a_sz, ns * a_sz * sizeof(double), cudaMemcpyHostToDevice));
        BLD_CUDA(cudaMemcpy(dY, Y + (size_t)s0 * y_sz, ns * y_sz * sizeof(double), cudaMemcpyHostToDevice));
        // Host arrays are row-major: A is (m x n) row-major = (n x m) column-major "A^T" for cuBLAS.  In column-major
        // terms  G (n x n) = At * At^T  and  B (k x n)^T ... we compute B^T: with Yt = Y^T (k x m) column-major,
        //   Xt0 (k x n) = Yt * At^T  is the row-major (n x k) matrix A^T Y.
        BLD_LIB(cublasDgemmStridedBatched(blas, CUBLAS_OP_N, CUBLAS_OP_T, n, n, m, &one, dA, n, (long long)a_sz, dA, n,
                                          (long long)a_sz, &zero, dG, n, (long long)g_sz, ns), "Gram GEMM");
        BLD_LIB(cublasDgemmStridedBatched(blas, CUBLAS_OP_N, CUBLAS_OP_T, k, n, m, &one, dY, k, (long long)y_sz, dA, n,
                                          (long long)a_sz, &zero, dX, k, (long long)x_sz, ns), "A^T Y GEMM");
        k_builder_shift<<<ns, 256>>>(dA, dG, (long long)a_sz, (long long)g_sz, m, n, reg);
        BLD_CUDA(cudaGetLastError());
        // G is symmetric, so row- / column-major agree.  Cholesky (lower, column-major), then solve G Z = (A^T Y) for the
        // k right-hand sides: in column-major storage dX holds (A^T Y)^T (k x n), i.e. we solve  Zt * G = Bt  from the right.
        BLD_LIB(cusolverDnDpotrfBatched(sol, CUBLAS_FILL_MODE_LOWER, n, dGp, n, dinfo, ns), "potrfBatched");
        std::vector<int> info(ns);
        BLD_CUDA(cudaMemcpy(info.data(), dinfo, ns * sizeof(int), cudaMemcpyDeviceToHost));
        for (int i = 0; i < ns; ++i)
            if (info[i] != 0) rc = ssb_builder_fail(-5, "ssb_solve_decoders: Gram matrix not positive definite");
        if (rc) break;
        // Zt * (L L^T) = Bt  ->  W * L^T = Bt (right, lower, transposed), then Zt * L = W (right, lower, not transposed)
        BLD_LIB(cublasDtrsmBatched(blas, CUBLAS_SIDE_RIGHT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_T, CUBLAS_DIAG_NON_UNIT, k, n, &one,
                                   (const double* const*)dGp, n, dXp, k, ns), "trsm 1");
        BLD_LIB(cublasDtrsmBatched(blas, CUBLAS_SIDE_RIGHT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, k, n, &one,
                                   (const double* const*)dGp, n, dXp, k, ns), "trsm 2");
        BLD_CUDA(cudaMemcpy(X + (size_t)s0 * x_sz, dX, ns * x_sz * sizeof(double), cudaMemcpyDeviceToHost));
    }
    cudaFree(dA);
    cudaFree(dY);
    cudaFree(dG);
    cudaFree(dX);
    cudaFree(dGp);
    cudaFree(dXp);
    cudaFree(dinfo);
    cusolverDnDestroy(sol);
    cublasDestroy(blas);
    return rc;
}
