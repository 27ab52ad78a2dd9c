// sm_100a kernels of the SSP-SLAM step engine: PES-learned decoders: k_pes, deferred PES (k_pes_hist / k_pes_defer / k_pes_fold / k_pes_clear).
// Included by ssb_kernels.cuh (after ssb_common.cuh); see that file for the layout rules.
#pragma once
#include "ssb_common.cuh"

// --------------------------------------------------------------------------------------
// PES-learned decoders (per trial): one streaming pass that applies the pending rank-1 delta,
// decodes with the updated weights and writes them back:
//   D <- D + outer(alpha*err_prev, a_prev)     (nengo: Copy(delta->weights, inc) at step start)
//   out = D . act                               (DotInc)
// err_prev / a_prev are the values the previous step read (the error rows are materialised from the
// not-yet-overwritten filter half, the trace comes from the other half of its ping-pong buffer), which
// is exactly SimPES' delta of the previous step.  For a fixed output row the weights of consecutive
// neurons are consecutive 128-byte lines.  A neuron whose trace and activity are zero in all 32 trials
// changes nothing and contributes nothing: its weights are neither read nor written (exact).
// desc: n size_out d_off a_off act0 err_vec out_vec alpha_bits decay_bits onemdecay_bits n_chunks part_off counter0
template <bool FULL>
__device__ __forceinline__ void ssb_pes_body(const float* __restrict__ ap, const float* __restrict__ fp, float* __restrict__ dp,
                                             int n, int jn, int i_lo, int i_hi, const float (&ae)[8], float (&acc)[8]) {
    const int warp = threadIdx.x >> 5;
    float* rowp[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) rowp[j] = dp + (size_t)((FULL || j < jn) ? j : 0) * n * 32;
    constexpr int U = 4;
    // activities / traces of the NEXT batch are requested before this batch's weights, so the two dependent
    // memory rounds of a batch (a, f -> vote -> weights) overlap across iterations
    float an[U], fn[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int ii = i_lo + warp + 4 * u;
        an[u] = 0.f;
        fn[u] = 0.f;
        if (ii < i_hi) {
            an[u] = ap[(size_t)ii * 32];
            fn[u] = fp[(size_t)ii * 32];
        }
    }
    for (int i = i_lo + warp; i < i_hi; i += 4 * U) {
        float a[U], f[U], w[U][8];
        bool on[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            a[u] = an[u];
            f[u] = fn[u];
            const int ii = i + 4 * U + 4 * u;
            an[u] = 0.f;
            fn[u] = 0.f;
            if (ii < i_hi) {
                an[u] = ap[(size_t)ii * 32];
                fn[u] = fp[(size_t)ii * 32];
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            on[u] = __any_sync(0xffffffffu, a[u] != 0.f || f[u] != 0.f);
            if (on[u]) {
                const size_t off = (size_t)(i + 4 * u) * 32;
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (FULL || j < jn) w[u][j] = __ldcs(rowp[j] + off);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (on[u]) {
                const size_t off = (size_t)(i + 4 * u) * 32;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (FULL || j < jn) {
                        const float wn = fmaf(ae[j], f[u], w[u][j]);
                        acc[j] = fmaf(wn, a[u], acc[j]);
                        __stcs(rowp[j] + off, wn);
                    }
                }
            }
        }
    }
}

__global__ void __launch_bounds__(128) k_pes(SsbCtx c, const int* __restrict__ desc, int max_chunks, int i_rel) {
    __shared__ float red[4][8][32];
    __shared__ int flag;
    const int item = blockIdx.z / max_chunks, chunk = blockIdx.z - item * max_chunks;
    const int* d = desc + item * 13;
    const int n = d[0], size_out = d[1], d_off = d[2], a_off = d[3], act0 = d[4], err_vec = d[5], out_vec = d[6];
    const int n_chunks = d[10];
    const float alpha = __int_as_float(d[7]);
    const int j0 = blockIdx.x * 8;
    if (j0 >= size_out || chunk >= n_chunks) return;
    const int per = (n + n_chunks - 1) / n_chunks;
    const int i_lo = chunk * per, i_hi = min(n, i_lo + per);
    const int lane = threadIdx.x & 31;
    const int g = blockIdx.y;
    const SsbStep s = ssb_step(c, i_rel);
    const int prev_buf = 1 - s.odd;  // afilt half that still holds what the previous step read
    const float* vg = ssb_grp(c.vec, c.nv, g, lane);
    const int jn = min(8, size_out - j0);
    float ae[8], acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        acc[j] = 0.f;
        float e = 0.f;
        if (j < jn) e = vg[(size_t)(err_vec + j0 + j) * 32];   // error of the previous step, materialised by k_lin
        ae[j] = s.step > 0 ? alpha * e : 0.f;
    }
    const float* __restrict__ ap = ssb_grp(c.act, c.n_act, g, lane) + (size_t)act0 * 32;
    const float* __restrict__ fp = ssb_grp(c.afilt, 2 * c.n_afilt, g, lane) + ((size_t)prev_buf * c.n_afilt + a_off) * 32;
    float* __restrict__ dp = ssb_grp(c.ldec, c.n_ldec, g, lane) + ((size_t)d_off + (size_t)j0 * n) * 32;
    if (jn == 8) ssb_pes_body<true>(ap, fp, dp, n, jn, i_lo, i_hi, ae, acc);
    else ssb_pes_body<false>(ap, fp, dp, n, jn, i_lo, i_hi, ae, acc);
    ssb_splitk_finish(c, red, &flag, acc, g, j0, size_out, out_vec, n_chunks, chunk, d[11],
                      (d[12] + (int)blockIdx.x) * c.G + g);
}

// --------------------------------------------------------------------------------------
// Deferred PES (default).  SimPES changes the decoders by one rank-1 term per step, D(t) = D(t-1) + ae(t) (x) f(t),
// and the only per-step consumer is out(t) = D(t) . a(t) with a sparse spike vector a.  Instead of rewriting D every
// step, the last K terms are kept as a history (ae_s: size_out rows, f_s: n rows per slot, slot = step mod K) and
//     out(t) = D_base . a(t) + sum_s ae_s * (f_s . a(t)),
// which reads D_base only where some trial of the group spiked and writes nothing; every K-th step (and before any
// read-back of the decoders) the K terms are folded into D_base in one streaming pass.  Same arithmetic up to fp32
// summation order; HBM traffic drops from 8 B to ~(active fraction * 4 + 8 / K) B per learned weight and step.
//   k_pes_hist   appends this step's term (ae from the materialised error rows, f = the trace the previous step read);
//                it runs AFTER the decode of its own step, which reads that term at its source
//   k_pes_defer  the sparse decode; CTA = (8-row tile, trial group, neuron chunk); tile-0 CTAs also accumulate the K
//                history dot products; the last CTA of a (decoder, group) adds partials in a fixed order and applies
//                the history correction
//   k_pes_fold   D_base += sum_s ae_s (x) f_s (runs when slot == K - 1, or when the host asks), then k_pes_clear zeroes
//                the ae rows, so an empty history always contributes exactly 0
// desc as k_pes; hdesc per decoder: e_row0 f_row0 part_row0 counter0 (rows of the hist_e / hist_f / pes_part arenas)

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

template <int K>
__global__ void __launch_bounds__(128) k_pes_defer(SsbCtx c, SsbPesDefer h, const int* __restrict__ desc,
                                                     const int* __restrict__ hdesc, int max_chunks, int i_rel) {
    __shared__ float red[4][8][32];
    __shared__ int flag;
    const int item = blockIdx.z / max_chunks, chunk = blockIdx.z - item * max_chunks;
    const int* d = desc + item * 13;
    const int* hd = hdesc + item * 4;
    const int n = d[0], size_out = d[1], d_off = d[2], a_off = d[3], act0 = d[4], err_vec = d[5], out_vec = d[6];
    const int n_chunks = d[10];
    const int n_jt = (size_out + 7) >> 3;
    // blockIdx.x < n_jt: an 8-row tile of D_base; blockIdx.x == n_jt: the K history rows (f_s . a), same loop
    const bool dots = (int)blockIdx.x == n_jt;
    const int j0 = blockIdx.x * 8;
    if ((int)blockIdx.x > n_jt || chunk >= n_chunks) return;
    const int per = (n + n_chunks - 1) / n_chunks;
    const int i_lo = chunk * per, i_hi = min(n, i_lo + per);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = blockIdx.y;
    const int jn = dots ? K : min(8, size_out - j0);
    const SsbStep s = ssb_step(c, i_rel);
    const int slot = (int)(s.step % K);       // this step's term is not in the history yet: it is read at its source
    const float* __restrict__ ap = ssb_grp(c.act, c.n_act, g, lane) + (size_t)act0 * 32;
    const int* __restrict__ fl = c.aflag + (size_t)g * c.n_act + act0;
    const float* rowp[8];
    {
        const float* dp = ssb_grp(c.ldec, c.n_ldec, g, lane) + ((size_t)d_off + (size_t)j0 * n) * 32;
        const float* hf = ssb_grp(h.hist_f, h.rows_f, g, lane) + (size_t)hd[1] * 32;
        const float* fcur = ssb_grp(c.afilt, 2 * c.n_afilt, g, lane) + ((size_t)(1 - s.odd) * c.n_afilt + a_off) * 32;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (dots) rowp[j] = (j == slot) ? fcur : hf + (size_t)(j < K ? j : 0) * n * 32;
            else rowp[j] = dp + (size_t)(j < jn ? j : 0) * n * 32;
        }
    }
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    // each warp owns a contiguous quarter of the chunk and walks only the neurons flagged active by their producer:
    // 32 flags per coalesced load -> ballot -> up to U active neurons per batch with all their loads in flight
    const int qn = (i_hi - i_lo + 3) >> 2;
    const int w_lo = i_lo + warp * qn, w_hi = min(i_hi, w_lo + qn);
    constexpr int U = 4;
    for (int base = w_lo; base < w_hi; base += 32) {
        // the flag word of a neuron is the ballot of its producer: bit t = trial t has a non-zero activity
        const int myflag = (base + lane < w_hi) ? __ldg(fl + base + lane) : 0;
        unsigned m = __ballot_sync(0xffffffffu, myflag != 0);
        while (m) {
            int idx[U];
            unsigned bits[U];
            float a[U], w[U][8];
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
#pragma unroll
            for (int u = 0; u < U; ++u) {
                a[u] = 0.f;
                if (idx[u] >= 0) {
                    const size_t off = (size_t)idx[u] * 32;
                    // a lane whose trial is inactive contributes w * 0: it does not load at all, so only the 32-byte
                    // sectors of the trials that spiked are fetched from DRAM (the flag ORs 32 trials); the predicate
                    // comes from the flag word, not from the activity load, so the two stay in flight together
                    const bool mine = (bits[u] >> lane) & 1u;
                    a[u] = mine ? ap[off] : 0.f;
#pragma unroll
                    for (int j = 0; j < 8; ++j) w[u][j] = (j < jn && mine) ? __ldcs(rowp[j] + off) : 0.f;
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (idx[u] >= 0) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[j] = fmaf(w[u][j], a[u], acc[j]);
                }
            }
        }
    }
    // CTA partial: ((w0 + w1) + (w2 + w3)) per row, parked in the partial arena [chunk][size_out + K]
#pragma unroll
    for (int j = 0; j < 8; ++j) red[warp][j][lane] = acc[j];
    __syncthreads();
    const int prow = size_out + K;
    float* pg = ssb_grp(h.part, h.rows_p, g, lane) + (size_t)hd[2] * 32;
    for (int j = warp; j < jn; j += 4) {
        const float t = (red[0][j][lane] + red[1][j][lane]) + (red[2][j][lane] + red[3][j][lane]);
        pg[(size_t)(chunk * prow + (dots ? size_out : j0) + j) * 32] = t;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        int* cnt_p = h.counters + hd[3] * c.G + g;
        const int old = atomicAdd(cnt_p, 1);
        const int last = old == (n_jt + 1) * n_chunks - 1;
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

// launched by the host after the step whose slot is K - 1, and before any read-back of the decoders
template <int K>
__global__ void __launch_bounds__(128) k_pes_fold(SsbCtx c, SsbPesDefer h, const int* __restrict__ desc,
                                                    const int* __restrict__ hdesc, int max_chunks, int i_rel, int force) {
    const int item = blockIdx.z / max_chunks, chunk = blockIdx.z - item * max_chunks;
    const int* d = desc + item * 13;
    const int* hd = hdesc + item * 4;
    const int n = d[0], size_out = d[1], d_off = d[2], n_chunks = d[10];
    const int j0 = blockIdx.x * 8;
    if (j0 >= size_out || chunk >= n_chunks) return;
    const int per = (n + n_chunks - 1) / n_chunks;
    const int i_lo = chunk * per, i_hi = min(n, i_lo + per);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = blockIdx.y;
    const int jn = min(8, size_out - j0);
    float* __restrict__ dp = ssb_grp(c.ldec, c.n_ldec, g, lane) + ((size_t)d_off + (size_t)j0 * n) * 32;
    const float* __restrict__ hf = ssb_grp(h.hist_f, h.rows_f, g, lane) + (size_t)hd[1] * 32;
    const float* __restrict__ he = ssb_grp(h.hist_e, h.rows_e, g, lane) + (size_t)hd[0] * 32;
    float ae[K][8];
#pragma unroll
    for (int q = 0; q < K; ++q)
#pragma unroll
        for (int j = 0; j < 8; ++j) ae[q][j] = (j < jn) ? he[(size_t)(q * size_out + j0 + j) * 32] : 0.f;
    constexpr int U = 2;                     // two neurons per iteration: 2 * (K + 8) loads in flight per warp
    for (int i = i_lo + warp; i < i_hi; i += 4 * U) {
        float fv[U][K], w[U][8];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int ii = i + 4 * u;
            const size_t off = (size_t)(ii < i_hi ? ii : i) * 32;
#pragma unroll
            for (int q = 0; q < K; ++q) fv[u][q] = hf[(size_t)q * n * 32 + off];
#pragma unroll
            for (int j = 0; j < 8; ++j) w[u][j] = (j < jn) ? __ldcs(dp + (size_t)j * n * 32 + off) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int ii = i + 4 * u;
            if (ii < i_hi) {
                const size_t off = (size_t)ii * 32;
#pragma unroll
                for (int q = 0; q < K; ++q)
#pragma unroll
                    for (int j = 0; j < 8; ++j) w[u][j] = fmaf(ae[q][j], fv[u][q], w[u][j]);
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (j < jn) __stcs(dp + (size_t)j * n * 32 + off, w[u][j]);
            }
        }
    }
}

__global__ void __launch_bounds__(128) k_pes_clear(SsbCtx c, SsbPesDefer h, int i_rel, int force) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int r = blockIdx.x * 4 + warp;
    if (r < h.rows_e) ssb_grp(h.hist_e, h.rows_e, blockIdx.y, lane)[(size_t)r * 32] = 0.f;
}

