// sm_100a kernels of the SSP-SLAM step engine: PES-learned decoders — deferred PES (k_pes_hist / k_pes_defer / k_pes_fold / k_pes_clear).
// Included by ssb_kernels.cuh (after ssb_common.cuh); see that file for the layout rules.
#pragma once
#include "ssb_common.cuh"

// --------------------------------------------------------------------------------------
// PES-learned decoders (per trial).  nengo: SimPES computes delta = alpha * outer(err, a_filtered) as an update,
// Copy(delta -> weights, inc) applies it at the start of the next step, DotInc decodes: out(t) = D(t) . a(t),
// D(t) = D(t-1) + ae(t) (x) f(t) with ae = alpha * err and f the trace the previous step read.
//
// Deferred form.  The only per-step consumer of D is that decode, and a (the spikes) is sparse, so instead of rewriting D
// every step the last K rank-1 terms are kept as a history (ae_s: size_out rows, f_s: n rows per slot, slot = step mod K):
//     out(t) = D_base . a(t) + sum_s ae_s * (f_s . a(t)),
// which reads D_base only where a trial spiked and writes nothing; every K-th step (and before any read-back of the
// decoders) the K terms are folded into D_base in one streaming pass.  Same arithmetic up to fp32 summation order.
//
// Decoder layout in HBM: D_base of one trial group is [neuron i][trial lane][JP] floats (JP = size_out rounded up to 4),
// i.e. the size_out weights that ONE spike of ONE trial needs are contiguous (224 bytes at d = 55: seven 32-byte sectors,
// one DRAM page) instead of size_out sectors spread over size_out pages.  In units of 128-byte arena rows the block of a
// neuron is JP rows: float offset = (d_off + i * JP) * 32 + lane * JP + j.
//   k_pes_hist   appends this step's term (ae from the materialised error rows, f = the trace the previous step read) and
//                updates the activity traces; it runs AFTER the decode of its own step, which reads that term at its source
//   k_pes_defer  the sparse decode.  CTA = (neuron chunk, trial group, tile of 32 output columns); a warp walks the flag
//                words of its quarter of the chunk (bit t = trial t spiked, written by the ensemble kernel), and per spiking
//                (neuron, trial) the lane of that trial loads 8 x 16 bytes and accumulates 32 outputs in registers; the
//                column-tile-0 CTAs also accumulate the K history dot products f_s . a.  Lane = trial, so there is no
//                cross-lane reduction; warps are added through shared memory, chunks through the partial arena, and the
//                last CTA of a (decoder, group) adds the partials in chunk order and applies the history correction.
//   k_pes_fold   D_base += sum_s ae_s (x) f_s (when slot == K - 1, or when the host asks), then k_pes_clear zeroes the ae
//                rows, so an empty history always contributes exactly 0
// desc: n size_out d_off a_off act0 err_vec out_vec alpha_bits decay_bits onemdecay_bits n_chunks part_off counter0
// hdesc per decoder: e_row0 f_row0 part_row0 counter0 (rows of the hist_e / hist_f / pes_part arenas)
#define SSB_PES_JT 32          // output columns per k_pes_defer CTA
#define SSB_PES_FT 64          // output columns per k_pes_fold shared-memory tile

__host__ __device__ __forceinline__ int ssb_pes_jp(int size_out) { return (size_out + 3) & ~3; }

__global__ void __launch_bounds__(128) k_pes_hist(SsbCtx c, SsbPesDefer h, const int* __restrict__ desc,
                                                    const int* __restrict__ hdesc, int i_rel) {
    const int item = blockIdx.z;
    const int* d = desc + item * 13;
    const int* hd = hdesc + item * 4;
    const int n = d[0], size_out = d[1], a_off = d[3], err_vec = d[5];
    const float alpha = __int_as_float(d[7]);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = blockIdx.y;
    const SsbStep s = ssb_step(c, i_rel);
    const int slot = (int)(s.step % h.K);
    const int r = blockIdx.x * 4 + warp;
    float* he = ssb_grp(h.hist_e, h.rows_e, g, lane) + (size_t)hd[0] * 32;
    if (r < size_out) {
        const float e = ssb_grp(c.vec, c.nv, g, lane)[(size_t)(err_vec + r) * 32];
        he[(size_t)(slot * size_out + r) * 32] = s.step > 0 ? alpha * e : 0.f;
    } else if (r < size_out + n) {
        const int i = r - size_out;
        const int prev_buf = 1 - s.odd;      // afilt half that still holds what the previous step read
        float* fg = ssb_grp(c.afilt, 2 * c.n_afilt, g, lane);
        const float f = fg[((size_t)prev_buf * c.n_afilt + a_off + i) * 32];
        ssb_grp(h.hist_f, h.rows_f, g, lane)[(size_t)(hd[1] + slot * n + i) * 32] = f;
        // this kernel is the last reader of that half in the step, so it also performs the trace update the row
        // program would do (kind 2): new trace = decay * trace + (1 - decay) * activity, written over the old half
        const float y = fg[((size_t)s.odd * c.n_afilt + a_off + i) * 32];
        const float u = ssb_grp(c.act, c.n_act, g, lane)[(size_t)(d[4] + i) * 32];
        fg[((size_t)prev_buf * c.n_afilt + a_off + i) * 32] = fmaf(__int_as_float(d[9]), u, __int_as_float(d[8]) * y);
    }
}

// Sparse decode of up to U spiking neurons at a time: registers w[U][JT] are all in flight before the first FMA.
template <int K>
__global__ void __launch_bounds__(128) k_pes_defer(SsbCtx c, SsbPesDefer h, const int* __restrict__ desc,
                                                     const int* __restrict__ hdesc, int n_jt_max, int i_rel) {
    __shared__ float red[4][SSB_PES_JT + 8][32];
    __shared__ int flag;
    const int item = blockIdx.z / n_jt_max, jt = blockIdx.z - item * n_jt_max;
    const int* d = desc + item * 13;
    const int* hd = hdesc + item * 4;
    const int n = d[0], size_out = d[1], d_off = d[2], a_off = d[3], act0 = d[4], err_vec = d[5], out_vec = d[6];
    const int n_chunks = d[10], chunk = blockIdx.x;
    const int JP = ssb_pes_jp(size_out);
    const int n_jt = (JP + SSB_PES_JT - 1) / SSB_PES_JT;
    if (jt >= n_jt || chunk >= n_chunks) return;
    const int j0 = jt * SSB_PES_JT;
    const int jw = min(SSB_PES_JT, JP - j0);                 // columns of this tile (a multiple of 4)
    const bool dots = jt == 0;
    const int per = (n + n_chunks - 1) / n_chunks;
    const int i_lo = chunk * per, i_hi = min(n, i_lo + per);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = blockIdx.y;
    const SsbStep s = ssb_step(c, i_rel);
    const int slot = (int)(s.step % K);       // this step's term is not in the history yet: it is read at its source
    const float* __restrict__ ap = ssb_grp(c.act, c.n_act, g, lane) + (size_t)act0 * 32;
    const int* __restrict__ fl = c.aflag + (size_t)g * c.n_act + act0;
    // this lane's trial: decoder block of neuron i starts at dl + i * JP * 32
    const float* __restrict__ dl = c.ldec + ((size_t)g * c.n_ldec + d_off) * 32 + (size_t)lane * JP + j0;
    const float* __restrict__ hf = ssb_grp(h.hist_f, h.rows_f, g, lane) + (size_t)hd[1] * 32;
    const float* __restrict__ fcur = ssb_grp(c.afilt, 2 * c.n_afilt, g, lane) + ((size_t)(1 - s.odd) * c.n_afilt + a_off) * 32;
    float acc[SSB_PES_JT], dsum_l[K];
#pragma unroll
    for (int j = 0; j < SSB_PES_JT; ++j) acc[j] = 0.f;
#pragma unroll
    for (int q = 0; q < K; ++q) dsum_l[q] = 0.f;
    const int qn = (i_hi - i_lo + 3) >> 2;
    const int w_lo = i_lo + warp * qn, w_hi = min(i_hi, w_lo + qn);
    constexpr int U = 4;
    for (int base = w_lo; base < w_hi; base += 32) {
        const int myflag = (base + lane < w_hi) ? __ldg(fl + base + lane) : 0;
        unsigned m = __ballot_sync(0xffffffffu, myflag != 0);
        while (m) {
            int idx[U];
            unsigned bits[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                idx[u] = -1;
                int src = 0;
                if (m) {
                    src = __ffs(m) - 1;
                    idx[u] = base + src;
                    m &= m - 1;
                }
                bits[u] = (unsigned)__shfl_sync(0xffffffffu, myflag, src);
            }
            float a[U];
            float4 w[U][SSB_PES_JT / 4];
            float f[U][K];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const bool mine = idx[u] >= 0 && ((bits[u] >> lane) & 1u);
                const size_t ni = (size_t)(idx[u] >= 0 ? idx[u] : 0);
                a[u] = mine ? ap[ni * 32] : 0.f;
                const float4* src4 = reinterpret_cast<const float4*>(dl + ni * JP * 32);
#pragma unroll
                for (int q4 = 0; q4 < SSB_PES_JT / 4; ++q4)
                    w[u][q4] = (mine && 4 * q4 < jw) ? __ldcs(src4 + q4) : make_float4(0.f, 0.f, 0.f, 0.f);
                if (dots) {
#pragma unroll
                    for (int q = 0; q < K; ++q)
                        f[u][q] = mine ? ((q == slot) ? fcur[ni * 32] : hf[((size_t)q * n + ni) * 32]) : 0.f;
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
#pragma unroll
                for (int q4 = 0; q4 < SSB_PES_JT / 4; ++q4) {
                    acc[4 * q4 + 0] = fmaf(w[u][q4].x, a[u], acc[4 * q4 + 0]);
                    acc[4 * q4 + 1] = fmaf(w[u][q4].y, a[u], acc[4 * q4 + 1]);
                    acc[4 * q4 + 2] = fmaf(w[u][q4].z, a[u], acc[4 * q4 + 2]);
                    acc[4 * q4 + 3] = fmaf(w[u][q4].w, a[u], acc[4 * q4 + 3]);
                }
                if (dots) {
#pragma unroll
                    for (int q = 0; q < K; ++q) dsum_l[q] = fmaf(f[u][q], a[u], dsum_l[q]);
                }
            }
        }
    }
    // CTA partial: ((w0 + w1) + (w2 + w3)) per row, parked in the partial arena [chunk][size_out + K]
#pragma unroll
    for (int j = 0; j < SSB_PES_JT; ++j) red[warp][j][lane] = acc[j];
#pragma unroll
    for (int q = 0; q < K; ++q) red[warp][SSB_PES_JT + q][lane] = dsum_l[q];
    __syncthreads();
    const int prow = size_out + K;
    float* pg = ssb_grp(h.part, h.rows_p, g, lane) + (size_t)hd[2] * 32;
    for (int r = warp; r < SSB_PES_JT + (dots ? K : 0); r += 4) {
        const bool is_dot = r >= SSB_PES_JT;
        const int row = is_dot ? size_out + (r - SSB_PES_JT) : j0 + r;
        if (is_dot || row < size_out) {
            const float t = (red[0][r][lane] + red[1][r][lane]) + (red[2][r][lane] + red[3][r][lane]);
            pg[(size_t)(chunk * prow + row) * 32] = t;
        }
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        int* cnt_p = h.counters + hd[3] * c.G + g;
        const int old = atomicAdd(cnt_p, 1);
        const int last = old == n_jt * n_chunks - 1;
        if (last) *cnt_p = 0;
        flag = last;
    }
    __syncthreads();
    if (!flag) return;
    __threadfence();
    // the last CTA of this (decoder, group): history dot products, then every output row
    float dsum[K];
#pragma unroll
    for (int q = 0; q < K; ++q) dsum[q] = 0.f;
    for (int ck = 0; ck < n_chunks; ++ck) {              // K independent loads per chunk, added in chunk order
        float v[K];
#pragma unroll
        for (int q = 0; q < K; ++q) v[q] = __ldcg(pg + (size_t)(ck * prow + size_out + q) * 32);
#pragma unroll
        for (int q = 0; q < K; ++q) dsum[q] += v[q];
    }
    const float* __restrict__ he = ssb_grp(h.hist_e, h.rows_e, g, lane) + (size_t)hd[0] * 32;
    float* vg = ssb_grp(c.vec, c.nv, g, lane);
    const float alpha = s.step > 0 ? __int_as_float(d[7]) : 0.f;
    for (int jb = warp * 8; jb < size_out; jb += 32) {   // each warp takes 8 consecutive output rows at a time
        float t[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) t[u] = 0.f;
        for (int ck = 0; ck < n_chunks; ++ck) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = (jb + u < size_out) ? __ldcg(pg + (size_t)(ck * prow + jb + u) * 32) : 0.f;
#pragma unroll
            for (int u = 0; u < 8; ++u) t[u] += v[u];
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (jb + u < size_out) {
                float e[K];
#pragma unroll
                for (int q = 0; q < K; ++q)
                    e[q] = (q == slot) ? alpha * vg[(size_t)(err_vec + jb + u) * 32] : he[(size_t)(q * size_out + jb + u) * 32];
                float r = t[u];
#pragma unroll
                for (int q = 0; q < K; ++q) r = fmaf(e[q], dsum[q], r);
                vg[(size_t)(out_vec + jb + u) * 32] = r;
            }
        }
    }
}

// launched by the host after the step whose slot is K - 1, and before any read-back of the decoders.
// CTA = (neuron chunk, trial group, decoder); the ae rows of a tile of SSB_PES_FT output columns sit in shared memory
// ([K][FT][32], lane-interleaved: conflict-free), each warp takes a neuron, each lane its own trial's contiguous weights.
template <int K>
__global__ void __launch_bounds__(128) k_pes_fold(SsbCtx c, SsbPesDefer h, const int* __restrict__ desc,
                                                    const int* __restrict__ hdesc, int max_chunks, int i_rel, int force) {
    extern __shared__ __align__(16) float sm[];          // [K][SSB_PES_FT][32]
    const int item = blockIdx.z, chunk = blockIdx.x;
    const int* d = desc + item * 13;
    const int* hd = hdesc + item * 4;
    const int n = d[0], size_out = d[1], d_off = d[2];
    const int JP = ssb_pes_jp(size_out);
    const int per = (n + max_chunks - 1) / max_chunks;
    const int i_lo = chunk * per, i_hi = min(n, i_lo + per);
    if (i_lo >= i_hi) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = blockIdx.y;
    float* __restrict__ dl = c.ldec + ((size_t)g * c.n_ldec + d_off) * 32 + (size_t)lane * JP;
    const float* __restrict__ hf = ssb_grp(h.hist_f, h.rows_f, g, lane) + (size_t)hd[1] * 32;
    const float* __restrict__ he = ssb_grp(h.hist_e, h.rows_e, g, lane) + (size_t)hd[0] * 32;
    for (int j0 = 0; j0 < JP; j0 += SSB_PES_FT) {
        const int jw = min(SSB_PES_FT, JP - j0);
        __syncthreads();
        for (int r = warp; r < K * SSB_PES_FT; r += 4) {
            const int q = r / SSB_PES_FT, j = r - q * SSB_PES_FT;
            sm[r * 32 + lane] = (j0 + j < size_out) ? he[(size_t)(q * size_out + j0 + j) * 32] : 0.f;
        }
        __syncthreads();
        for (int i = i_lo + warp; i < i_hi; i += 4) {
            float fv[K];
#pragma unroll
            for (int q = 0; q < K; ++q) fv[q] = hf[((size_t)q * n + i) * 32];
            float4* w4 = reinterpret_cast<float4*>(dl + (size_t)i * JP * 32 + j0);
            float4 w[SSB_PES_FT / 4];
#pragma unroll
            for (int q4 = 0; q4 < SSB_PES_FT / 4; ++q4) w[q4] = (4 * q4 < jw) ? __ldcs(w4 + q4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int q = 0; q < K; ++q) {
                const float* aq = sm + (size_t)q * SSB_PES_FT * 32 + lane;
#pragma unroll
                for (int q4 = 0; q4 < SSB_PES_FT / 4; ++q4) {
                    w[q4].x = fmaf(aq[(4 * q4 + 0) * 32], fv[q], w[q4].x);
                    w[q4].y = fmaf(aq[(4 * q4 + 1) * 32], fv[q], w[q4].y);
                    w[q4].z = fmaf(aq[(4 * q4 + 2) * 32], fv[q], w[q4].z);
                    w[q4].w = fmaf(aq[(4 * q4 + 3) * 32], fv[q], w[q4].w);
                }
            }
#pragma unroll
            for (int q4 = 0; q4 < SSB_PES_FT / 4; ++q4)
                if (4 * q4 < jw) __stcs(w4 + q4, w[q4]);
        }
    }
}

__global__ void __launch_bounds__(128) k_pes_clear(SsbCtx c, SsbPesDefer h, int i_rel, int force) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int r = blockIdx.x * 4 + warp;
    if (r < h.rows_e) ssb_grp(h.hist_e, h.rows_e, blockIdx.y, lane)[(size_t)r * 32] = 0.f;
}

