// sm_100a kernels of the SSP-SLAM step engine: Voja-class ensembles with very long encoder rows (d = 649), CTA-cooperative form.
// Included by ssb_kernels.cuh after ssb_ens_wide.cuh.
#pragma once
#include "ssb_common.cuh"
#include "ssb_ens_wide.cuh"

// --------------------------------------------------------------------------------------
// k_wide_voja_stream (one warp per neuron range, read-only ring of sub-tiles, Voja update by per-lane sector accesses to
// L2) runs at 1.8 - 2.3 TB/s at d = 649: every neuron in which ANY of the 32 trials spiked costs 2 x 649 uncoalesced
// 32-byte accesses per spiking lane.  Here the whole CTA works on ONE neuron at a time instead:
//   * the neuron's encoder tile ([dims][32 trials] floats, contiguous: 83 KB at d = 649) arrives in one of two shared-memory
//     buffers by a TMA bulk copy issued by a producer warp;
//   * the 8 consumer warps split the dims rows (lane = trial); each keeps ITS slice of the input vector in registers
//     (RW rows), so the tile is the only shared-memory operand; partial dot products are combined in warp order;
//   * warp 0 does bias + neuron-current terms + LIF, publishes the activity; lanes that spiked then update their column of
//     the tile IN shared memory (same arithmetic as k_wide_voja) and the producer writes a dirty tile back with ONE bulk
//     store - full 128-byte rows in both directions instead of sector traffic.
// CTA = (neuron chunk, trial group, ensemble), 288 threads (8 consumer warps + producer), one CTA per SM.
// dynamic smem: 2 tiles [dims][32] | red [8][32] | out [32] | us [jn_m][32]
#define SSB_VC_NW 8
template <int RW>
__global__ void __launch_bounds__(32 * (SSB_VC_NW + 1), 1)
k_wide_voja_cta(SsbCtx c, const int* __restrict__ desc, SsbItemList items, int chunk, int i_rel) {
    extern __shared__ __align__(128) float sm[];
    __shared__ unsigned long long full[2], done[2];
    __shared__ int s_dirty[2];
    const int* d = desc + items.idx[blockIdx.z] * 16;
    const int n = d[0], dims = d[1], state0 = d[3], act0 = d[4], enc_off = d[5], bias_off = d[6];
    const int in_row0 = d[7], jn_row0 = d[10], jn_m = d[11], jn_w = d[12], voja_row = d[13], scale_off = d[14];
    const int n0 = blockIdx.x * chunk;
    if (n0 >= n) return;
    const int cnt = min(chunk, n - n0);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = blockIdx.y;
    const size_t tile_f = (size_t)dims * 32;
    float* tiles = sm;                                     // [2][dims][32]
    float* red = tiles + 2 * tile_f;                       // [8][32]
    float* s_out = red + SSB_VC_NW * 32;                   // [32]
    float* us = s_out + 32;                                // [jn_m][32]
    float* eg = c.lenc + ((size_t)g * c.n_lenc + enc_off + (size_t)n0 * dims) * 32;    // tile of neuron n0
    const uint32_t tile_bytes = (uint32_t)dims * 128u;
    if (threadIdx.x == 0) {
        for (int b = 0; b < 2; ++b) {
            ssb_mbar_init(&full[b], 1);
            ssb_mbar_init(&done[b], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp == SSB_VC_NW) {
        // ---------------- producer: tile t -> buffer t & 1; before a buffer is refilled, the tile it held is written back
        if (lane == 0) {
            for (int t = 0; t < cnt + 2; ++t) {
                const int b = t & 1;
                if (t >= 2) {
                    ssb_mbar_wait(&done[b], (uint32_t)((t - 2) >> 1) & 1u);
                    if (s_dirty[b]) {                      // (the consumers fenced their generic-proxy writes before arriving)
                        ssb_bulk_s2g(eg + (size_t)(t - 2) * tile_f, tiles + (size_t)b * tile_f, tile_bytes);
                        ssb_bulk_commit();
                        ssb_bulk_wait_read0();
                    }
                }
                if (t < cnt) {
                    ssb_mbar_expect_tx(&full[b], tile_bytes);
                    ssb_bulk_g2s(tiles + (size_t)b * tile_f, eg + (size_t)t * tile_f, tile_bytes, &full[b]);
                }
            }
            ssb_bulk_wait0();
        }
        return;
    }
    // ---------------- consumers
    const SsbNeuron nt = ssb_neuron(c, d[8]);
    const bool stateful = nt.type == 0;
    const int rw = (dims + SSB_VC_NW - 1) / SSB_VC_NW;     // rows of this warp: [k0, k0 + rows)
    const int k0 = warp * rw, rows = max(0, min(rw, dims - k0));
    const float* vg = ssb_grp(c.vec, c.nv, g, lane);
    float x[RW];
#pragma unroll
    for (int j = 0; j < RW; ++j) x[j] = j < rows ? vg[(size_t)(in_row0 + k0 + j) * 32] : 0.f;
    const float aL = __int_as_float(d[15]) * vg[(size_t)voja_row * 32];
    const bool pt = (d[9] & 4) != 0;                       // per-trial bias / scale / neuron-current weights (own network seed)
    const int pstride = pt ? 32 : 1;
    const float* __restrict__ bias_p = pt ? c.wpt + ((size_t)g * c.n_wpt + bias_off + n0) * 32 + lane : c.W + bias_off + n0;
    const float* __restrict__ scale_p = pt ? c.wpt + ((size_t)g * c.n_wpt + scale_off + n0) * 32 + lane : c.W + scale_off + n0;
    const float* __restrict__ jn_p = pt ? c.wpt + ((size_t)g * c.n_wpt + jn_w + (size_t)n0 * jn_m) * 32 + lane
                                        : c.W + jn_w + (size_t)n0 * jn_m;
    if (warp == 0)
        for (int m = 0; m < jn_m; ++m) us[m * 32 + lane] = vg[(size_t)(jn_row0 + m) * 32];
    float* sp = ssb_grp(c.st, c.nn, g, lane) + (size_t)(state0 + n0) * 32;
    float* ag = ssb_grp(c.act, c.n_act, g, lane) + (size_t)(act0 + n0) * 32;
    float sv_next = 0.f, bias_next = 0.f;
    if (warp == 0) {
        if (stateful) sv_next = __ldcs(sp);
        bias_next = __ldg(bias_p);
    }
    auto bar_consumers = [] { asm volatile("bar.sync 1, %0;" ::"n"(32 * SSB_VC_NW) : "memory"); };
    for (int t = 0; t < cnt; ++t) {
        const int b = t & 1;
        float* E = tiles + (size_t)b * tile_f + (size_t)k0 * 32 + lane;
        ssb_mbar_wait(&full[b], (uint32_t)(t >> 1) & 1u);
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
        for (int j = 0; j < RW; j += 4) {
            if (j + 0 < rows) a0 = fmaf(E[(j + 0) * 32], x[j + 0], a0);
            if (j + 1 < rows) a1 = fmaf(E[(j + 1) * 32], x[j + 1], a1);
            if (j + 2 < rows) a2 = fmaf(E[(j + 2) * 32], x[j + 2], a2);
            if (j + 3 < rows) a3 = fmaf(E[(j + 3) * 32], x[j + 3], a3);
        }
        red[warp * 32 + lane] = (a0 + a1) + (a2 + a3);
        bar_consumers();
        if (warp == 0) {
            float sv = sv_next, J = bias_next;
            if (t + 1 < cnt) {
                if (stateful) sv_next = __ldcs(sp + (size_t)(t + 1) * 32);
                bias_next = __ldg(bias_p + (size_t)(t + 1) * pstride);
            }
            for (int m = 0; m < jn_m; ++m) J = fmaf(__ldg(jn_p + (size_t)(t * jn_m + m) * pstride), us[m * 32 + lane], J);
            float dot = 0.f;
#pragma unroll
            for (int w = 0; w < SSB_VC_NW; ++w) dot += red[w * 32 + lane];      // warp order: a fixed summation order
            J += dot;
            float out;
            if (stateful) {
                out = nt.fast ? ssb_lif_packed<true>(nt, J, sv) : ssb_lif_packed<false>(nt, J, sv);
                __stcs(sp + (size_t)t * 32, sv);
            } else {
                out = ssb_rate(nt, J);
            }
            ag[(size_t)t * 32] = out;
            const bool fired = out != 0.f;
            const unsigned any_on = __ballot_sync(0xffffffffu, fired);
            const bool learn = fired && aL != 0.f;
            const unsigned any_learn = __ballot_sync(0xffffffffu, learn);
            s_out[lane] = learn ? out : 0.f;
            if (lane == 0) {
                c.aflag[(size_t)g * c.n_act + act0 + n0 + t] = (int)any_on;
                s_dirty[b] = any_learn != 0u;
            }
        }
        bar_consumers();
        const bool dirty = s_dirty[b] != 0;
        if (dirty) {
            const float out = s_out[lane];
            if (out != 0.f) {
                const float sc = __ldg(scale_p + (size_t)t * pstride);
#pragma unroll
                for (int j = 0; j < RW; ++j) {
                    if (j < rows) {
                        const float e = E[j * 32];
                        E[j * 32] = e + aL * (sc * (out * x[j]) - out * e);
                    }
                }
            }
            ssb_fence_async();                             // generic-proxy writes -> visible to the producer's bulk store
        }
        bar_consumers();                                   // every warp is done with buffer b (and with red / s_out / s_dirty[b])
        if (threadIdx.x == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(ssb_smem(&done[b])) : "memory");
    }
}
