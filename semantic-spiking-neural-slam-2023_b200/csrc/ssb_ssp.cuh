// sm_100a kernels of the SSP-SLAM step engine: stand-alone SSP encode / decode-prep kernels.
// Included by ssb_kernels.cuh (after ssb_common.cuh); see that file for the layout rules.
#pragma once
#include "ssb_common.cuh"

// --------------------------------------------------------------------------------------
// Stand-alone SSP encode: out[p][m] = (1/d) * sum_k cos(theta_k + 2 pi k m / d), theta = A_scaled x.
__global__ void k_ssp_encode(const double* __restrict__ A, const double* __restrict__ x, double* __restrict__ out,
                             long long n_points, int n, int d) {
    extern __shared__ double cs[];  // [2][d]
    const long long p = blockIdx.x;
    if (p >= n_points) return;
    for (int k = threadIdx.x; k < d; k += blockDim.x) {
        double th = 0.0;
        for (int j = 0; j < n; ++j) th += A[(size_t)k * n + j] * x[(size_t)p * n + j];
        double sn, cn;
        sincos(th, &sn, &cn);
        cs[k] = cn;
        cs[d + k] = sn;
    }
    __syncthreads();
    for (int m = threadIdx.x; m < d; m += blockDim.x) {
        double acc = 0.0;
        for (int k = 0; k < d; ++k) {
            // exp(i*theta_k) * exp(+2 pi i k m / d); reduce k*m mod d to keep the angle small
            const int km = (int)(((long long)k * m) % d);
            double sn, cn;
            sincospi(2.0 * (double)km / (double)d, &sn, &cn);
            acc += cs[k] * cn - cs[d + k] * sn;
        }
        out[(size_t)p * d + m] = acc / (double)d;
    }
}

// Normalise query rows (skip if norm < 1e-6) and write them group-tiled [g][k][32] in float for the scan.
__global__ void k_decode_prep(const double* __restrict__ q, float* __restrict__ cx, long long n_q, int B, int d, int dpad,
                              long long q0) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B) return;
    float* cxg = cx + ((size_t)(t >> 5) * dpad) * 32 + (t & 31);
    const long long row = q0 + t;
    if (row >= n_q) {
        for (int k = 0; k < dpad; ++k) cxg[(size_t)k * 32] = 0.f;
        return;
    }
    double nrm = 0.0;
    for (int k = 0; k < d; ++k) nrm += q[(size_t)row * d + k] * q[(size_t)row * d + k];
    nrm = sqrt(nrm);
    const double sc = nrm < 1e-6 ? 1.0 : 1.0 / nrm;
    for (int k = 0; k < dpad; ++k) cxg[(size_t)k * 32] = k < d ? (float)(q[(size_t)row * d + k] * sc) : 0.f;
}

