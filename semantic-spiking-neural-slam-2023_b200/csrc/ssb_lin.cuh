// sm_100a kernels of the SSP-SLAM step engine: the row program k_lin (filters, probes, materialised sink rows) and k_advance.
// Included by ssb_kernels.cuh (after ssb_common.cuh); see that file for the layout rules.
#pragma once
#include "ssb_common.cuh"

// --------------------------------------------------------------------------------------
// End-of-step rows: Lowpass updates (y_new = a*y_old + b*u, written to the other half of
// the ping-pong buffer = nengo's update-after-read), probe samples, PES activity traces.
// The same kernel materialises the sink rows of a dependency level into vec scratch before the level's
// consumers run (kinds 3 / 4), so that no consumer evaluates CSR rows itself.
// kind 0 filter, 1 probe, 2 activity trace, 3 / 4 materialise (4: on the values the previous step read),
// 5 neuron-output probe (activity row -> probe block; CSR population only).
//
// Three CTA populations in one launch (the host sorts every segment's rows into them at finalize):
//  * dense items: rows that share one column list (the circular-convolution DFT matrices, to_Fourier / to_SSP,
//    decoder-to-filter fans) form a dense block T[R][Kpad].  A CTA owns 8 rows of a block for ONE trial group;
//    its four warps split the 32-column slabs (split-K), each warp gathers its slab's 32 source rows once into
//    registers and reads the 8 x 32 coefficient slab as shared-memory broadcasts; the four partial sums are
//    added in warp order.  A source row is fetched once per 8 sink rows instead of once per entry, and the
//    whole item is two dependent memory rounds (coefficients + column list, then the gathers);
//  * records: rows with at most 8 entries (most Lowpass inputs) and the PES activity traces are packed by the
//    host into one 128-byte record each: the warp reads it with one coalesced load and distributes the words
//    with shuffles, so a row is two dependent rounds (record, then gathers) and two rows per warp are in flight;
//  * CSR rows (the rest): one warp per (row, group), entries as warp-uniform 8-byte loads.
// record words: 0 kind | 1 dst | 2 a | 3 b | 4..11 columns on even steps | 12..19 columns on odd steps |
//               20..27 coefficients | 28 src (kind 2) | 29..31 unused
#define SSB_DENSE_RCH 8
#define SSB_DENSE_SLAB 32
#define SSB_REC_PER_WARP 4

// dense rows: kind dst a_bits b_bits
__device__ __forceinline__ void ssb_lin_store(const SsbCtx& c, const SsbStep& s, float* vg, int g, int lane, int kind, int dst,
                                              float a, float b, float u) {
    if (kind == 0) {
        const float y = vg[(size_t)(1 + dst + s.par_old) * 32];
        vg[(size_t)(1 + dst + s.par_new) * 32] = fmaf(b, u, a * y);
    } else if (kind >= 3) {
        vg[(size_t)dst * 32] = u;
    } else {
        float* pg = c.probe + (((size_t)g * c.probe_cap + (size_t)(s.step - c.dyn[2])) * c.n_probe + dst) * 32 + lane;
        __stcs(pg, u);
    }
}

struct SsbLinArgs {
    const int* rows;          // CSR rows [src kind dst lo hi]
    const float* ab;
    int n_rows;
    const int* items;         // dense items
    int n_items;
    const int* ddesc;
    const float* dT;
    const int* dcols;
    const int* drows;
    const int* recs;          // packed records, 32 words each
    int n_recs;
};

// MINB: resident CTAs per SM the register allocation aims at.  The launch is latency-bound and one to three waves long, so
// the host picks the variant with fewer waves for the launch's CTA count (6 per SM: 80 registers, 8: 64 with a few spills).
// (Walking the work units with a grid stride over one resident wave was measured slower: 18.4 vs 14.9 us per launch.)
template <int MINB>
__global__ void __launch_bounds__(128, MINB) k_lin(SsbCtx c, SsbLinArgs L, int i_rel) {
    __shared__ __align__(16) float s_t[4][SSB_DENSE_RCH][SSB_DENSE_SLAB];
    __shared__ float s_red[4][SSB_DENSE_RCH][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const SsbStep s = ssb_step(c, i_rel);
    const int n_dense_ctas = L.n_items * c.G;
    if ((int)blockIdx.x < n_dense_ctas) {
        const int item = blockIdx.x / c.G, g = blockIdx.x - item * c.G;
        // item: t_off (of its first row) | cols_off | kpad | rows_off (of its first row) | nr | previous-step view | - | -
        const int4 it = __ldg(reinterpret_cast<const int4*>(L.items + (size_t)item * 8));
        const int2 it2 = __ldg(reinterpret_cast<const int2*>(L.items + (size_t)item * 8 + 4));
        const int kpad = it.z, nr = it2.x;
        const float* __restrict__ T = L.dT + it.x;
        const int* __restrict__ dr = L.drows + (size_t)it.w * 4;
        const int* __restrict__ cols = L.dcols + it.y + ((s.odd ^ it2.y) ? kpad : 0);
        float* vg = ssb_grp(c.vec, c.nv, g, lane);
        // this warp's output rows (r = warp, warp + 4): descriptors requested now, used after the reduction
        int4 rd[2];
#pragma unroll
        for (int q = 0; q < 2; ++q)
            rd[q] = (warp + 4 * q < nr) ? __ldg(reinterpret_cast<const int4*>(dr) + warp + 4 * q) : make_int4(3, 0, 0, 0);
        float acc[SSB_DENSE_RCH];
#pragma unroll
        for (int r = 0; r < SSB_DENSE_RCH; ++r) acc[r] = 0.f;
        const int n_slabs = kpad / SSB_DENSE_SLAB;
        const int tr0 = lane >> 3, tq = lane & 7;                       // lane -> rows tr0, tr0 + 4, float4 tq of the slab
        for (int sl = warp; sl < n_slabs; sl += 4) {
            const int k0 = sl * SSB_DENSE_SLAB;
            const int col = __ldg(cols + k0 + lane);
            float4 t0 = make_float4(0.f, 0.f, 0.f, 0.f), t1 = t0;
            if (tr0 < nr) t0 = __ldg(reinterpret_cast<const float4*>(T + (size_t)tr0 * kpad + k0) + tq);
            if (tr0 + 4 < nr) t1 = __ldg(reinterpret_cast<const float4*>(T + (size_t)(tr0 + 4) * kpad + k0) + tq);
            float x[SSB_DENSE_SLAB];
#pragma unroll
            for (int e = 0; e < SSB_DENSE_SLAB; ++e) x[e] = ssb_ld_src(vg + (size_t)__shfl_sync(0xffffffffu, col, e) * 32);
            __syncwarp();                                               // previous slab's broadcasts are done
            *reinterpret_cast<float4*>(&s_t[warp][tr0][tq * 4]) = t0;
            *reinterpret_cast<float4*>(&s_t[warp][tr0 + 4][tq * 4]) = t1;
            __syncwarp();
#pragma unroll
            for (int r = 0; r < SSB_DENSE_RCH; ++r) {
#pragma unroll
                for (int q = 0; q < SSB_DENSE_SLAB / 4; ++q) {
                    const float4 t = *reinterpret_cast<const float4*>(&s_t[warp][r][q * 4]);
                    acc[r] = fmaf(t.x, x[4 * q + 0], acc[r]);
                    acc[r] = fmaf(t.y, x[4 * q + 1], acc[r]);
                    acc[r] = fmaf(t.z, x[4 * q + 2], acc[r]);
                    acc[r] = fmaf(t.w, x[4 * q + 3], acc[r]);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < SSB_DENSE_RCH; ++r) s_red[warp][r][lane] = acc[r];
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 2; ++q) {                                   // fixed order: ((w0 + w1) + w2) + w3
            const int r = warp + 4 * q;
            if (r < nr) {
                const float u = ((s_red[0][r][lane] + s_red[1][r][lane]) + s_red[2][r][lane]) + s_red[3][r][lane];
                ssb_lin_store(c, s, vg, g, lane, rd[q].x, rd[q].y, __int_as_float(rd[q].z), __int_as_float(rd[q].w), u);
            }
        }
        return;
    }
    const int rec_per_cta = 4 * SSB_REC_PER_WARP;
    const int n_rec_ctas = ((L.n_recs + rec_per_cta - 1) / rec_per_cta) * c.G;
    if ((int)blockIdx.x < n_dense_ctas + n_rec_ctas) {
        const int cb = blockIdx.x - n_dense_ctas;
        const int rblk = cb / c.G, g = cb - rblk * c.G;
        const int r0 = (rblk * 4 + warp) * SSB_REC_PER_WARP;
        float* vg = ssb_grp(c.vec, c.nv, g, lane);
        int w[SSB_REC_PER_WARP];
#pragma unroll
        for (int q = 0; q < SSB_REC_PER_WARP; ++q) w[q] = (r0 + q < L.n_recs) ? __ldg(L.recs + (size_t)(r0 + q) * 32 + lane) : 0;
        float x[SSB_REC_PER_WARP][8], y[SSB_REC_PER_WARP];
        int kind[SSB_REC_PER_WARP], dst[SSB_REC_PER_WARP];
        const int cbase = 4 + (s.odd ? 8 : 0);
#pragma unroll
        for (int q = 0; q < SSB_REC_PER_WARP; ++q) {
            kind[q] = __shfl_sync(0xffffffffu, w[q], 0);
            dst[q] = __shfl_sync(0xffffffffu, w[q], 1);
            y[q] = 0.f;
            if (r0 + q >= L.n_recs) {
                kind[q] = -1;
                continue;
            }
            if (kind[q] == 2) {
                const int src = __shfl_sync(0xffffffffu, w[q], 28);
                x[q][0] = ssb_grp(c.act, c.n_act, g, lane)[(size_t)src * 32];
                y[q] = ssb_grp(c.afilt, 2 * c.n_afilt, g, lane)[((size_t)s.odd * c.n_afilt + dst[q]) * 32];
            } else {
                const int cb4 = kind[q] == 4 ? 4 + (s.odd ? 0 : 8) : cbase;   // kind 4 reads the previous step's view
#pragma unroll
                for (int e = 0; e < 8; ++e) x[q][e] = ssb_ld_src(vg + (size_t)__shfl_sync(0xffffffffu, w[q], cb4 + e) * 32);
                if (kind[q] == 0) y[q] = vg[(size_t)(1 + dst[q] + s.par_old) * 32];
            }
        }
#pragma unroll
        for (int q = 0; q < SSB_REC_PER_WARP; ++q) {
            if (kind[q] < 0) continue;
            const float a = __int_as_float(__shfl_sync(0xffffffffu, w[q], 2)), b = __int_as_float(__shfl_sync(0xffffffffu, w[q], 3));
            if (kind[q] == 2) {
                ssb_grp(c.afilt, 2 * c.n_afilt, g, lane)[((size_t)(1 - s.odd) * c.n_afilt + dst[q]) * 32] = fmaf(b, x[q][0], a * y[q]);
                continue;
            }
            float u = 0.f;
#pragma unroll
            for (int e = 0; e < 8; ++e) u = fmaf(__int_as_float(__shfl_sync(0xffffffffu, w[q], 20 + e)), x[q][e], u);
            if (kind[q] == 0) vg[(size_t)(1 + dst[q] + s.par_new) * 32] = fmaf(b, u, a * y[q]);
            else if (kind[q] >= 3) vg[(size_t)dst[q] * 32] = u;
            else {
                float* pg = c.probe + (((size_t)g * c.probe_cap + (size_t)(s.step - c.dyn[2])) * c.n_probe + dst[q]) * 32 + lane;
                __stcs(pg, u);
            }
        }
        return;
    }
    // ---- CSR rows: flat index -> (row block, group)
    const int cb = blockIdx.x - n_dense_ctas - n_rec_ctas;
    const int rblk = cb / c.G, g = cb - rblk * c.G;
    const int r = rblk * 4 + warp;
    if (r >= L.n_rows) return;
    float* vg = ssb_grp(c.vec, c.nv, g, lane);
    const int* rp = L.rows + (size_t)r * 5;
    const int src = rp[0], kind = rp[1], dst = rp[2];
    const float a = L.ab[r * 2], b = L.ab[r * 2 + 1];
    if (kind == 2) {
        float* fg = ssb_grp(c.afilt, 2 * c.n_afilt, g, lane);
        const float y = fg[((size_t)s.odd * c.n_afilt + dst) * 32];
        const float u = ssb_grp(c.act, c.n_act, g, lane)[(size_t)src * 32];
        fg[((size_t)(1 - s.odd) * c.n_afilt + dst) * 32] = fmaf(b, u, a * y);
        return;
    }
    if (kind == 5) {                                          // neuron-output probe: this step's activity row, unfiltered
        ssb_lin_store(c, s, vg, g, lane, 1, dst, a, b, ssb_grp(c.act, c.n_act, g, lane)[(size_t)src * 32]);
        return;
    }
    const int2* __restrict__ ent = kind == 4 ? s.ent_new : s.ent_old;
    const int lo = rp[3], hi = rp[4];
    float u = 0.f;
    int p = lo;
    for (; p + 32 <= hi; p += 32) u = ssb_row_batch<32>(ent + p, vg, u);
    for (; p < hi; p += 8) u = ssb_row_batch<8>(ent + p, vg, u);
    ssb_lin_store(c, s, vg, g, lane, kind, dst, a, b, u);
}

__global__ void k_advance(long long* dyn, int n) { dyn[0] += n; }

