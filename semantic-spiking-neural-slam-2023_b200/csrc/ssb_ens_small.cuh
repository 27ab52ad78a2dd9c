// sm_100a kernels of the SSP-SLAM step engine: narrow ensembles (VCOs, product squares): k_ens_small.
// Included by ssb_kernels.cuh (after ssb_common.cuh); see that file for the layout rules.
#pragma once
#include "ssb_common.cuh"

// --------------------------------------------------------------------------------------
// Narrow ensembles (VCO 3-D x 500, product squares 1-D x 50): fused encode -> neuron -> decode.
// Each warp walks a contiguous neuron range of one (ensemble, trial group) in chunks of SSB_SM_CH
// neurons.  A chunk's packed weights [bias, enc[DIMS], dec[nout]] and its 128-byte state rows are
// staged in shared memory by TMA bulk copies (double-buffered per warp, mbarrier completion); the
// updated state goes back with a bulk store.  Input vector and decoded sums live in registers.
//   blocks [0, n_split*G)  "split":  a CTA of 4 warps owns one (ensemble, group); the neuron range is
//                                    quartered and the partial decodes are reduced in shared memory;
//   remaining blocks       "packed": each warp owns one (ensemble, group) of a small ensemble.
// desc: n, dims, nout, state0, w_off, in_row0, out_vec, ntype, stride
struct __align__(128) SsbSmallSmem {
    float st[4][2][SSB_SM_CH * 32];
    float w[4][2][SSB_SM_CH * SSB_SM_WMAX];
    float red[4][8][32];
    unsigned long long bar[4][2];
};

template <int DIMS, int S4, int MODE>
__device__ __forceinline__ void ssb_small_range(const SsbCtx& c, const int* __restrict__ d, const SsbNeuron& nt,
                                                const float* vg, int g, int i_begin, int i_end, float (&acc)[8],
                                                SsbSmallSmem& sm) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int state0 = d[3], w_off = d[4], in_vec = d[5];
    const bool stateful = nt.type == 0;
    constexpr int NCOL = (4 * S4 - 1 - DIMS) < 8 ? (4 * S4 - 1 - DIMS) : 8;   // decoder columns present (zero padded)
    const float* wsrc = c.W + w_off;
    float* sg = c.st + ((size_t)g * c.nn + state0) * 32;
    const int n_chunks = (i_end - i_begin + SSB_SM_CH - 1) / SSB_SM_CH;
    uint32_t phases = 0;
    auto issue = [&](int ck) {
        if (lane == 0) {
            const int b = ck & 1, i0 = i_begin + ck * SSB_SM_CH, cnt = min(SSB_SM_CH, i_end - i0);
            const uint32_t bw = (uint32_t)cnt * S4 * 16, bs = stateful ? (uint32_t)cnt * 128 : 0u;
            ssb_mbar_expect_tx(&sm.bar[warp][b], bw + bs);
            ssb_bulk_g2s(sm.w[warp][b], wsrc + (size_t)i0 * 4 * S4, bw, &sm.bar[warp][b]);
            if (stateful) ssb_bulk_g2s(sm.st[warp][b], sg + (size_t)i0 * 32, bs, &sm.bar[warp][b]);
        }
    };
    if (n_chunks > 0) issue(0);
    if (n_chunks > 1) issue(1);
    float x[DIMS];                       // the materialised input vector (written by k_lin for this level)
#pragma unroll
    for (int k = 0; k < DIMS; ++k) x[k] = vg[(size_t)(in_vec + k) * 32];
    for (int ck = 0; ck < n_chunks; ++ck) {
        const int b = ck & 1, i0 = i_begin + ck * SSB_SM_CH, cnt = min(SSB_SM_CH, i_end - i0);
        ssb_mbar_wait(&sm.bar[warp][b], (phases >> b) & 1u);
        phases ^= 1u << b;
        float* ss = sm.st[warp][b] + lane;
        const float4* ww = reinterpret_cast<const float4*>(sm.w[warp][b]);
        int k = 0;
        if (MODE == 0) {       // fast-LIF ensembles: two neurons per iteration on the packed fp32 pipe
#pragma unroll 2
            for (; k + 2 <= cnt; k += 2) {
                float wa[4 * S4], wb[4 * S4];
#pragma unroll
                for (int q = 0; q < S4; ++q) {
                    const float4 t = ww[k * S4 + q], u = ww[(k + 1) * S4 + q];
                    wa[4 * q + 0] = t.x;
                    wa[4 * q + 1] = t.y;
                    wa[4 * q + 2] = t.z;
                    wa[4 * q + 3] = t.w;
                    wb[4 * q + 0] = u.x;
                    wb[4 * q + 1] = u.y;
                    wb[4 * q + 2] = u.z;
                    wb[4 * q + 3] = u.w;
                }
                float2 J = make_float2(wa[0], wb[0]);
#pragma unroll
                for (int kk = 0; kk < DIMS; ++kk) J = ssb_fma2(make_float2(wa[1 + kk], wb[1 + kk]), ssb_splat(x[kk]), J);
                float2 sv = make_float2(ss[k * 32], ss[(k + 1) * 32]);
                const float2 out = ssb_lif_pair(nt, J, sv);
                ss[k * 32] = sv.x;
                ss[(k + 1) * 32] = sv.y;
#pragma unroll
                for (int j = 0; j < NCOL; ++j) acc[j] = fmaf(wb[1 + DIMS + j], out.y, fmaf(wa[1 + DIMS + j], out.x, acc[j]));
            }
        }
#pragma unroll 4
        for (; k < cnt; ++k) {
            float wl[4 * S4];
#pragma unroll
            for (int q = 0; q < S4; ++q) {
                const float4 t = ww[k * S4 + q];
                wl[4 * q + 0] = t.x;
                wl[4 * q + 1] = t.y;
                wl[4 * q + 2] = t.z;
                wl[4 * q + 3] = t.w;
            }
            float J = wl[0];
#pragma unroll
            for (int kk = 0; kk < DIMS; ++kk) J = fmaf(wl[1 + kk], x[kk], J);
            float sv = 0.f;
            if (MODE == 0 || stateful) sv = ss[k * 32];
            const float out = ssb_neuron_apply<MODE>(nt, J, sv);
            if (MODE == 0 || stateful) ss[k * 32] = sv;
#pragma unroll
            for (int j = 0; j < NCOL; ++j) acc[j] = fmaf(wl[1 + DIMS + j], out, acc[j]);
        }
        if (stateful) {
            ssb_fence_async();     // generic-proxy writes of this chunk -> visible to the bulk store
            __syncwarp();
            if (lane == 0) {
                ssb_bulk_s2g(sg + (size_t)i0 * 32, sm.st[warp][b], (uint32_t)cnt * 128);
                ssb_bulk_commit();
            }
        }
        if (ck + 2 < n_chunks) {
            if (stateful && lane == 0) ssb_bulk_wait_read0();   // the store has drained this buffer
            __syncwarp();
            issue(ck + 2);
        }
    }
}

template <int DIMS, int S4, int MODE>
__device__ __forceinline__ void ssb_small_item(const SsbCtx& c, const int* __restrict__ d, const SsbNeuron& nt, int g,
                                               bool split, SsbSmallSmem& sm) {
    const int n = d[0], nout = d[2], out_vec = d[6];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* vg = ssb_grp(c.vec, c.nv, g, lane);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    if (split) {
        const int q = (n + 3) >> 2;
        ssb_small_range<DIMS, S4, MODE>(c, d, nt, vg, g, min(n, warp * q), min(n, (warp + 1) * q), acc, sm);
#pragma unroll
        for (int j = 0; j < 8; ++j) sm.red[warp][j][lane] = acc[j];
        __syncthreads();
        for (int j = warp; j < nout; j += 4) {
            const float t = (sm.red[0][j][lane] + sm.red[1][j][lane]) + (sm.red[2][j][lane] + sm.red[3][j][lane]);
            vg[(size_t)(out_vec + j) * 32] = t;
        }
    } else {
        ssb_small_range<DIMS, S4, MODE>(c, d, nt, vg, g, 0, n, acc, sm);
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (j < nout) vg[(size_t)(out_vec + j) * 32] = acc[j];
    }
}

template <int MODE>
__device__ __forceinline__ void ssb_small_dispatch(const SsbCtx& c, const int* __restrict__ d, const SsbNeuron& nt, int g,
                                                   bool split, SsbSmallSmem& sm) {
    const int key = d[1] * 8 + (d[8] >> 2);
    switch (key) {
#define SSB_CASE(D, S) \
    case (D) * 8 + (S): ssb_small_item<D, S, MODE>(c, d, nt, g, split, sm); break;
        SSB_CASE(1, 1) SSB_CASE(1, 2) SSB_CASE(1, 3)
        SSB_CASE(2, 1) SSB_CASE(2, 2) SSB_CASE(2, 3)
        SSB_CASE(3, 1) SSB_CASE(3, 2) SSB_CASE(3, 3)
        SSB_CASE(4, 2) SSB_CASE(4, 3) SSB_CASE(4, 4)
#undef SSB_CASE
        default: break;  // excluded by the host-side lowering (dims <= 4, dims + nout <= 11)
    }
}

// desc: n, dims, nout, state0, w_off, in_vec, out_vec, ntype, stride
__global__ void __launch_bounds__(128, 6) k_ens_small(SsbCtx c, const int* __restrict__ desc, int n_items, int n_split) {
    __shared__ SsbSmallSmem sm;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
        ssb_mbar_init(&sm.bar[warp][0], 1);
        ssb_mbar_init(&sm.bar[warp][1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int G = c.G;
    const int split_blocks = n_split * G;
    int item, g;
    bool split, live = true;
    if ((int)blockIdx.x < split_blocks) {
        split = true;
        item = blockIdx.x / G;
        g = blockIdx.x - item * G;
    } else {
        split = false;
        const int w = (blockIdx.x - split_blocks) * 4 + warp;
        live = w < (n_items - n_split) * G;
        item = live ? n_split + w / G : 0;
        g = live ? w % G : 0;
    }
    if (live) {
        const int* d = desc + item * 9;
        const SsbNeuron nt = ssb_neuron(c, d[7]);
        if (nt.type == 0 && nt.fast) ssb_small_dispatch<0>(c, d, nt, g, split, sm);
        else ssb_small_dispatch<1>(c, d, nt, g, split, sm);
    }
    if (lane == 0) ssb_bulk_wait0();   // bulk stores complete before the CTA's shared memory is released
}

// --------------------------------------------------------------------------------------
// Narrow ensembles with PER-TRIAL static weights (every trial built from its own network seed, SURVEY.md 8b
// `share_weights=False`; experiments/run_slam.py:151 `nengo.Network(seed=args.seed)` batched over seeds).  The packed
// weights [bias | encoders | decoders] of a neuron are then 128-byte rows of 32 trials like the state, row
// (w_off + i * stride + q): nothing is shared between lanes, so there is nothing to stage or broadcast - every lane
// streams its own trial's words with coalesced loads (13 rows per VCO neuron: 12 weight rows + the state row) and the
// kernel is HBM-bound by construction (4 x (stride + 2) bytes per neuron and trial).  Same arithmetic, same block
// structure (split / packed items) as k_ens_small.
template <int DIMS, int S4, int MODE>
__device__ __forceinline__ void ssb_small_range_pt(const SsbCtx& c, const int* __restrict__ d, const SsbNeuron& nt,
                                                   const float* vg, int g, int i_begin, int i_end, float (&acc)[8]) {
    const int lane = threadIdx.x & 31;
    const int state0 = d[3], w_off = d[4], in_vec = d[5];
    const bool stateful = nt.type == 0;
    constexpr int ST = 4 * S4;
    constexpr int NCOL = (ST - 1 - DIMS) < 8 ? (ST - 1 - DIMS) : 8;
    constexpr int NW = 1 + DIMS + NCOL;                    // rows actually used (the rest of the stride is padding)
    const float* wg = c.wpt + ((size_t)g * c.n_wpt + w_off) * 32 + lane;
    float* sg = c.st + ((size_t)g * c.nn + state0) * 32 + lane;
    float x[DIMS];
#pragma unroll
    for (int k = 0; k < DIMS; ++k) x[k] = vg[(size_t)(in_vec + k) * 32];
    int i = i_begin;
    if (MODE == 0) {
        for (; i + 2 <= i_end; i += 2) {
            float wa[NW], wb[NW];
#pragma unroll
            for (int q = 0; q < NW; ++q) {
                wa[q] = __ldcs(wg + ((size_t)i * ST + q) * 32);
                wb[q] = __ldcs(wg + ((size_t)(i + 1) * ST + q) * 32);
            }
            float2 sv = make_float2(__ldcs(sg + (size_t)i * 32), __ldcs(sg + (size_t)(i + 1) * 32));
            float2 J = make_float2(wa[0], wb[0]);
#pragma unroll
            for (int kk = 0; kk < DIMS; ++kk) J = ssb_fma2(make_float2(wa[1 + kk], wb[1 + kk]), ssb_splat(x[kk]), J);
            const float2 out = ssb_lif_pair(nt, J, sv);
            __stcs(sg + (size_t)i * 32, sv.x);
            __stcs(sg + (size_t)(i + 1) * 32, sv.y);
#pragma unroll
            for (int j = 0; j < NCOL; ++j) acc[j] = fmaf(wb[1 + DIMS + j], out.y, fmaf(wa[1 + DIMS + j], out.x, acc[j]));
        }
    }
    for (; i < i_end; ++i) {
        float wl[NW];
#pragma unroll
        for (int q = 0; q < NW; ++q) wl[q] = __ldcs(wg + ((size_t)i * ST + q) * 32);
        float J = wl[0];
#pragma unroll
        for (int kk = 0; kk < DIMS; ++kk) J = fmaf(wl[1 + kk], x[kk], J);
        float sv = 0.f;
        if (MODE == 0 || stateful) sv = __ldcs(sg + (size_t)i * 32);
        const float out = ssb_neuron_apply<MODE>(nt, J, sv);
        if (MODE == 0 || stateful) __stcs(sg + (size_t)i * 32, sv);
#pragma unroll
        for (int j = 0; j < NCOL; ++j) acc[j] = fmaf(wl[1 + DIMS + j], out, acc[j]);
    }
}

template <int DIMS, int S4, int MODE>
__device__ __forceinline__ void ssb_small_item_pt(const SsbCtx& c, const int* __restrict__ d, const SsbNeuron& nt, int g,
                                                  bool split, float (*red)[8][32]) {
    const int n = d[0], nout = d[2], out_vec = d[6];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* vg = ssb_grp(c.vec, c.nv, g, lane);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    if (split) {
        const int q = (n + 3) >> 2;
        ssb_small_range_pt<DIMS, S4, MODE>(c, d, nt, vg, g, min(n, warp * q), min(n, (warp + 1) * q), acc);
#pragma unroll
        for (int j = 0; j < 8; ++j) red[warp][j][lane] = acc[j];
        __syncthreads();
        for (int j = warp; j < nout; j += 4) {
            const float t = (red[0][j][lane] + red[1][j][lane]) + (red[2][j][lane] + red[3][j][lane]);
            vg[(size_t)(out_vec + j) * 32] = t;
        }
    } else {
        ssb_small_range_pt<DIMS, S4, MODE>(c, d, nt, vg, g, 0, n, acc);
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (j < nout) vg[(size_t)(out_vec + j) * 32] = acc[j];
    }
}

template <int MODE>
__device__ __forceinline__ void ssb_small_dispatch_pt(const SsbCtx& c, const int* __restrict__ d, const SsbNeuron& nt, int g,
                                                      bool split, float (*red)[8][32]) {
    const int key = d[1] * 8 + (d[8] >> 2);
    switch (key) {
#define SSB_CASE(D, S) \
    case (D) * 8 + (S): ssb_small_item_pt<D, S, MODE>(c, d, nt, g, split, red); break;
        SSB_CASE(1, 1) SSB_CASE(1, 2) SSB_CASE(1, 3)
        SSB_CASE(2, 1) SSB_CASE(2, 2) SSB_CASE(2, 3)
        SSB_CASE(3, 1) SSB_CASE(3, 2) SSB_CASE(3, 3)
        SSB_CASE(4, 2) SSB_CASE(4, 3) SSB_CASE(4, 4)
#undef SSB_CASE
        default: break;
    }
}

__global__ void __launch_bounds__(128, 6) k_ens_small_pt(SsbCtx c, const int* __restrict__ desc, int n_items, int n_split) {
    __shared__ float red[4][8][32];
    const int warp = threadIdx.x >> 5;
    const int G = c.G;
    const int split_blocks = n_split * G;
    int item, g;
    bool split, live = true;
    if ((int)blockIdx.x < split_blocks) {
        split = true;
        item = blockIdx.x / G;
        g = blockIdx.x - item * G;
    } else {
        split = false;
        const int w = (blockIdx.x - split_blocks) * 4 + warp;
        live = w < (n_items - n_split) * G;
        item = live ? n_split + w / G : 0;
        g = live ? w % G : 0;
    }
    if (live) {
        const int* d = desc + item * 9;
        const SsbNeuron nt = ssb_neuron(c, d[7]);
        if (nt.type == 0 && nt.fast) ssb_small_dispatch_pt<0>(c, d, nt, g, split, red);
        else ssb_small_dispatch_pt<1>(c, d, nt, g, split, red);
    }
}
