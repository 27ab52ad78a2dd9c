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
//   k_pes_defer  the sparse decode.  CTA = (neuron chunk, 8 trials of a group, tile of 128 output columns); one warp per
//                trial walks the flag words of the chunk (bit t = trial t spiked, written by the ensemble kernel) and, per
//                spike of its trial, reads the contiguous weights with lanes = columns; the column-tile-0 CTAs also
//                accumulate the K history dot products f_s . a (lane q = history slot q).  Chunks are combined through the
//                partial arena; the last CTA of a (decoder, group) adds them in chunk order and applies the history terms.
//   k_pes_fold   D_base += sum_s ae_s (x) f_s (when slot == K - 1, or when the host asks), then k_pes_clear zeroes the ae
//                rows, so an empty history always contributes exactly 0
// desc: n size_out d_off a_off act0 err_vec out_vec alpha_bits decay_bits onemdecay_bits n_chunks part_off counter0
// hdesc per decoder: e_row0 f_row0 part_row0 counter0 (rows of the hist_e / hist_f / pes_part arenas)
#define SSB_PES_JT 128         // output columns per k_pes_defer CTA (4 per lane)
#define SSB_PES_FT 64          // output columns per k_pes_fold shared-memory tile
#define SSB_PES_FS 68          // floats per (slot, trial) row of that tile: 16-byte aligned, 4-way instead of 32-way bank conflicts on the fill

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

// Sparse decode.  One WARP per trial (8 trials = 8 warps per CTA), lanes = output columns: the JP weights one spike of one
// trial needs are contiguous, so a spike is one coalesced read per 32 columns and every lane of the warp works (with
// lane = trial only the ~3 lanes whose trial spiked would).  The CTA stages the flag words of its neuron chunk in shared
// memory; a warp enumerates the set bits of its trial (ballot over 32 neurons at a time) and keeps up to U spikes in flight.
// grid (n_chunks * 4 trial octets, G, decoders * column tiles of SSB_PES_JT) x 256
template <int K>
__global__ void __launch_bounds__(256) k_pes_defer(SsbCtx c, SsbPesDefer h, const int* __restrict__ desc,
                                                     const int* __restrict__ hdesc, int n_jt_max, int i_rel) {
    extern __shared__ int sflag[];                       // [per] flag words of the chunk | [8 warps][per] spike lists
    __shared__ float red[SSB_PES_JT + 8][8];             // [column | history slot][trial of the octet]
    __shared__ int flag;
    const int item = blockIdx.z / n_jt_max, jt = blockIdx.z - item * n_jt_max;
    const int* d = desc + item * 13;
    const int* hd = hdesc + item * 4;
    const int n = d[0], size_out = d[1], d_off = d[2], a_off = d[3], act0 = d[4], err_vec = d[5], out_vec = d[6];
    const int n_chunks = d[10], chunk = blockIdx.x >> 2, oct = blockIdx.x & 3;
    const int JP = ssb_pes_jp(size_out);
    const int n_jt = (JP + SSB_PES_JT - 1) / SSB_PES_JT;
    if (jt >= n_jt || chunk >= n_chunks) return;
    const int j0 = jt * SSB_PES_JT;
    const bool dots = jt == 0;
    const int per = (n + n_chunks - 1) / n_chunks;
    const int i_lo = chunk * per, i_hi = min(n, i_lo + per);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = blockIdx.y;
    const int t = oct * 8 + warp;                         // this warp's trial inside the group
    const SsbStep s = ssb_step(c, i_rel);
    const int slot = (int)(s.step % K);       // this step's term is not in the history yet: it is read at its source
    const int* __restrict__ fl = c.aflag + (size_t)g * c.n_act + act0;
    for (int i = i_lo + (int)threadIdx.x; i < i_hi; i += 256) sflag[i - i_lo] = __ldg(fl + i);
    __syncthreads();
    const float* __restrict__ ap = c.act + ((size_t)g * c.n_act + act0) * 32 + t;
    const float* __restrict__ dl = c.ldec + ((size_t)g * c.n_ldec + d_off) * 32 + (size_t)t * JP + j0 + lane;
    const float* __restrict__ hf = h.hist_f + ((size_t)g * h.rows_f + hd[1]) * 32 + t;
    const float* __restrict__ fcur = c.afilt + ((size_t)g * 2 * c.n_afilt + (size_t)(1 - s.odd) * c.n_afilt + a_off) * 32 + t;
    constexpr int NC = SSB_PES_JT / 32;                   // columns per lane
    float acc[NC], dot = 0.f;
#pragma unroll
    for (int q = 0; q < NC; ++q) acc[q] = 0.f;
    // phase 1: the neurons of the chunk where THIS trial spiked, compacted into a list (flags are in shared memory, so this
    // costs no memory round trip); phase 2: the list in batches of U with every load of a batch in flight - the number of
    // dependent round trips is spikes / U instead of one per 32 neurons
    int* list = sflag + per + warp * per;
    int cnt = 0;
    for (int base = i_lo; base < i_hi; base += 32) {
        const int fw = (base + lane < i_hi) ? sflag[base + lane - i_lo] : 0;
        const bool on = (fw >> t) & 1;
        const unsigned m = __ballot_sync(0xffffffffu, on);
        if (on) list[cnt + __popc(m & ((1u << lane) - 1u))] = base + lane;
        cnt += __popc(m);
    }
    __syncwarp();
    constexpr int U = 8;
    for (int s0 = 0; s0 < cnt; s0 += U) {
        float a[U], w[U][NC], f[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const bool on = s0 + u < cnt;
            const size_t ni = (size_t)(on ? list[s0 + u] : i_lo);
            a[u] = on ? __ldg(ap + ni * 32) : 0.f;
#pragma unroll
            for (int q = 0; q < NC; ++q)
                w[u][q] = (on && j0 + q * 32 + lane < JP) ? __ldcs(dl + ni * JP * 32 + q * 32) : 0.f;
            f[u] = 0.f;
            if (dots && on && lane < K) f[u] = (lane == slot) ? fcur[ni * 32] : hf[((size_t)lane * n + ni) * 32];
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
#pragma unroll
            for (int q = 0; q < NC; ++q) acc[q] = fmaf(w[u][q], a[u], acc[q]);
            dot = fmaf(f[u], a[u], dot);
        }
    }
    // CTA partial of the 8 trials -> partial arena [chunk][size_out + K] (32-byte runs: 8 consecutive trials per row)
#pragma unroll
    for (int q = 0; q < NC; ++q) red[q * 32 + lane][warp] = acc[q];
    if (lane < 8) red[SSB_PES_JT + lane][warp] = dot;
    __syncthreads();
    const int prow = size_out + K;
    float* pgo = h.part + ((size_t)g * h.rows_p + hd[2]) * 32 + oct * 8;
    for (int r = threadIdx.x >> 3; r < SSB_PES_JT + (dots ? K : 0); r += 32) {
        const bool is_dot = r >= SSB_PES_JT;
        const int row = is_dot ? size_out + (r - SSB_PES_JT) : j0 + r;
        if (is_dot || row < size_out) pgo[(size_t)(chunk * prow + row) * 32 + (threadIdx.x & 7)] = red[r][threadIdx.x & 7];
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        int* cnt_p = h.counters + hd[3] * c.G + g;
        const int old = atomicAdd(cnt_p, 1);
        const int last = old == n_jt * n_chunks * 4 - 1;
        if (last) *cnt_p = 0;
        flag = last;
    }
    __syncthreads();
    if (!flag) return;
    __threadfence();
    // the last CTA of this (decoder, group), lane = trial again: history dot products, then every output row
    float* pg = ssb_grp(h.part, h.rows_p, g, lane) + (size_t)hd[2] * 32;
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
    for (int jb = warp * 8; jb < size_out; jb += 64) {   // each warp takes 8 consecutive output rows at a time
        float tt[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) tt[u] = 0.f;
        for (int ck = 0; ck < n_chunks; ++ck) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = (jb + u < size_out) ? __ldcg(pg + (size_t)(ck * prow + jb + u) * 32) : 0.f;
#pragma unroll
            for (int u = 0; u < 8; ++u) tt[u] += v[u];
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (jb + u < size_out) {
                float e[K];
#pragma unroll
                for (int q = 0; q < K; ++q)
                    e[q] = (q == slot) ? alpha * vg[(size_t)(err_vec + jb + u) * 32] : he[(size_t)(q * size_out + jb + u) * 32];
                float r = tt[u];
#pragma unroll
                for (int q = 0; q < K; ++q) r = fmaf(e[q], dsum[q], r);
                vg[(size_t)(out_vec + jb + u) * 32] = r;
            }
        }
    }
}

// launched by the host after the step whose slot is K - 1, and before any read-back of the decoders:
//   D_base[i][t][j] += sum_q ae_q[j][t] * f_q[i][t].
// CTA = (neuron chunk, trial group, decoder), 8 warps.  The ae rows of the group are staged in shared memory as
// [q][trial][JP] (conflict-free when lanes walk j); a warp takes (neuron, pair of trials): its lanes are the 16-byte pieces
// of the two trials' contiguous weights, so the read-modify-write is fully coalesced.  Output widths above SSB_PES_FT columns
// are processed in column tiles.
template <int K>
__global__ void __launch_bounds__(256) k_pes_fold(SsbCtx c, SsbPesDefer h, const int* __restrict__ desc,
                                                    const int* __restrict__ hdesc, int max_chunks, int i_rel, int force) {
    extern __shared__ __align__(16) float sm[];          // [K][32 trials][SSB_PES_FS]
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
    float* __restrict__ dg = c.ldec + ((size_t)g * c.n_ldec + d_off) * 32;
    const float* __restrict__ hf = h.hist_f + ((size_t)g * h.rows_f + hd[1]) * 32;
    const float* __restrict__ he = h.hist_e + ((size_t)g * h.rows_e + hd[0]) * 32;
    for (int j0 = 0; j0 < JP; j0 += SSB_PES_FT) {
        const int jw = min(SSB_PES_FT, JP - j0);          // columns of this tile (multiple of 4)
        const int nq4 = jw >> 2;                          // 16-byte pieces per trial
        __syncthreads();
        for (int r0 = warp; r0 < K * jw; r0 += 64) {      // he row (q, j) -> sm[q][trial = lane][j], 8 rows in flight per warp
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int r = r0 + 8 * u, q = r / jw, j = r - q * jw;
                v[u] = (r < K * jw && j0 + j < size_out) ? he[(size_t)(q * size_out + j0 + j) * 32 + lane] : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int r = r0 + 8 * u, q = r / jw, j = r - q * jw;
                if (r < K * jw) sm[((size_t)q * 32 + lane) * SSB_PES_FS + j] = v[u];
            }
        }
        __syncthreads();
        const int tpw = 32 / nq4 > 0 ? 32 / nq4 : 1;      // trials per warp task (2 at d = 55)
        const int active = tpw * nq4;
        const int tasks_per_neuron = (32 + tpw - 1) / tpw;
        const int n_tasks = (i_hi - i_lo) * tasks_per_neuron;
        constexpr int TU = 4;                             // tasks in flight per warp (each is load -> 8 fma rounds -> store)
        for (int task0 = warp * TU; task0 < n_tasks; task0 += 8 * TU) {
            float4 w[TU];
            float fv[TU][K];
            float4* wp[TU];
            int tt[TU];
            bool ok[TU];
#pragma unroll
            for (int u = 0; u < TU; ++u) {
                const int task = task0 + u;
                const int i = i_lo + task / tasks_per_neuron;
                tt[u] = (task % tasks_per_neuron) * tpw + lane / nq4;
                ok[u] = task < n_tasks && lane < active && tt[u] < 32;
                const int ts = ok[u] ? tt[u] : 0;
                const int is = ok[u] ? i : i_lo;
                wp[u] = reinterpret_cast<float4*>(dg + (size_t)is * JP * 32 + (size_t)ts * JP + j0) + (lane % nq4);
                w[u] = ok[u] ? __ldcs(wp[u]) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int q = 0; q < K; ++q) fv[u][q] = ok[u] ? __ldg(hf + ((size_t)q * n + is) * 32 + ts) : 0.f;
                tt[u] = ts;
            }
#pragma unroll
            for (int u = 0; u < TU; ++u) {
#pragma unroll
                for (int q = 0; q < K; ++q) {
                    const float4 e = *reinterpret_cast<const float4*>(sm + ((size_t)q * 32 + tt[u]) * SSB_PES_FS + 4 * (lane % nq4));
                    w[u].x = fmaf(e.x, fv[u][q], w[u].x);
                    w[u].y = fmaf(e.y, fv[u][q], w[u].y);
                    w[u].z = fmaf(e.z, fv[u][q], w[u].z);
                    w[u].w = fmaf(e.w, fv[u][q], w[u].w);
                }
                if (ok[u]) __stcs(wp[u], w[u]);
            }
        }
    }
}

// CTA-cooperative fold for decoders up to 56 columns (JP * (K + SSB_PFC_NT) * 128 bytes of shared memory must fit): the
// warp-task form above keeps ~43 KB of loads in flight per SM and streams at 2.9 TB/s (154 us per fold on configs[1]).  Here a
// neuron's decoder block ([32 trials][JP] floats = JP 128-byte rows, contiguous) is a TILE of a ring of SSB_PFC_NT shared-memory
// buffers filled by TMA bulk copies (producer warp); the 8 consumer warps split the tile's 16-byte columns and keep THEIR
// slice of the K x [32][JP] history terms ae in registers, so the tile is the only shared-memory traffic; f_q[i][trial] is one
// coalesced load per (warp, slot) and reaches the lanes by shuffle.  A finished tile goes back with one bulk store; its
// buffer is refilled once that store has read it (the producer lags SSB_PFC_LAG stores behind).
// CTA = (neuron chunk, trial group, decoder), 288 threads, one CTA per SM.
#define SSB_PFC_NT 6           // ring buffers
#define SSB_PFC_NPT 4          // neurons per ring buffer (one bulk copy / store / fence / barrier round per 4 neurons)
#define SSB_PFC_LAG 1
template <int K>
__global__ void __launch_bounds__(288, 1) k_pes_fold_cta(SsbCtx c, SsbPesDefer h, const int* __restrict__ desc,
                                                           const int* __restrict__ hdesc, int max_chunks) {
    extern __shared__ __align__(128) float sm[];          // [SSB_PFC_NT][SSB_PFC_NPT][32 * JP] | f ring [SSB_PFC_NT][K][SSB_PFC_NPT][32]
    __shared__ unsigned long long full[SSB_PFC_NT], done[SSB_PFC_NT];
    const int item = blockIdx.z, chunk = blockIdx.x;
    const int* d = desc + item * 13;
    const int* hd = hdesc + item * 4;
    const int n = d[0], size_out = d[1], d_off = d[2];
    const int JP = ssb_pes_jp(size_out);
    const int per = (n + max_chunks - 1) / max_chunks;
    const int i_lo = chunk * per, i_hi = min(n, i_lo + per);
    if (i_lo >= i_hi) return;
    const int cnt = i_hi - i_lo;                          // neurons of this CTA
    const int n_tiles = (cnt + SSB_PFC_NPT - 1) / SSB_PFC_NPT;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = blockIdx.y;
    const int neuron_f = 32 * JP, tile_f = SSB_PFC_NPT * neuron_f;
    float* dg = c.ldec + ((size_t)g * c.n_ldec + d_off) * 32 + (size_t)i_lo * neuron_f;    // block of neuron i_lo
    auto tile_bytes = [&](int t) { return (uint32_t)min(SSB_PFC_NPT, cnt - t * SSB_PFC_NPT) * (uint32_t)neuron_f * 4u; };
    // the history factors f_q[i][trial] of a tile's neurons ride on the same barrier: for one slot q the rows of consecutive
    // neurons are contiguous, so they are K more (small) bulk copies per tile - as plain loads they were a dependent L2 round
    // trip per neuron and the pace of the whole CTA (231 / 194 us per fold without / with a two-neuron register prefetch)
    float* fr = sm + (size_t)SSB_PFC_NT * tile_f;
    const float* __restrict__ hfb = h.hist_f + ((size_t)g * h.rows_f + hd[1]) * 32;
    auto load_tile = [&](int t, int b) {
        const int nn = min(SSB_PFC_NPT, cnt - t * SSB_PFC_NPT);
        ssb_mbar_expect_tx(&full[b], tile_bytes(t) + (uint32_t)(K * nn * 128));
        ssb_bulk_g2s(sm + (size_t)b * tile_f, dg + (size_t)t * tile_f, tile_bytes(t), &full[b]);
        for (int q = 0; q < K; ++q)
            ssb_bulk_g2s(fr + ((size_t)(b * K + q) * SSB_PFC_NPT) * 32, hfb + ((size_t)q * n + i_lo + t * SSB_PFC_NPT) * 32,
                         (uint32_t)nn * 128u, &full[b]);
    };
    if (threadIdx.x == 0) {
        for (int b = 0; b < SSB_PFC_NT; ++b) {
            ssb_mbar_init(&full[b], 1);
            ssb_mbar_init(&done[b], 8);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp == 8) {
        if (lane == 0) {
            for (int t = 0; t < SSB_PFC_NT && t < n_tiles; ++t) load_tile(t, t);
            for (int t = 0; t < n_tiles; ++t) {
                const int b = t % SSB_PFC_NT;
                ssb_mbar_wait(&done[b], (uint32_t)(t / SSB_PFC_NT) & 1u);        // the 8 warps fenced their writes before arriving
                ssb_bulk_s2g(dg + (size_t)t * tile_f, sm + (size_t)b * tile_f, tile_bytes(t));
                ssb_bulk_commit();
                const int tr = t - SSB_PFC_LAG;                                  // the store of tile tr has read its buffer
                if (tr >= 0 && tr + SSB_PFC_NT < n_tiles) {
                    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(SSB_PFC_LAG) : "memory");
                    load_tile(tr + SSB_PFC_NT, tr % SSB_PFC_NT);
                }
            }
            ssb_bulk_wait0();
        }
        return;
    }
    // consumers: warp w owns the 16-byte columns m = w, w + 8 (< JP / 4) of every neuron block: floats p = 128 m + 4 lane .. + 3,
    // i.e. trial p / JP, columns p % JP .. + 3 of that trial
    const float* __restrict__ he = h.hist_e + ((size_t)g * h.rows_e + hd[0]) * 32;
    const int nq = JP >> 2;
    float4 e[2][K];
    int tr_of[2], valid[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int m = warp + 8 * u;
        valid[u] = m < nq;
        const int p = 128 * m + 4 * lane;
        const int trial = valid[u] ? p / JP : 0, j = valid[u] ? p - trial * JP : 0;
        tr_of[u] = trial;
#pragma unroll
        for (int q = 0; q < K; ++q) {
            float v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k)
                v[k] = (valid[u] && j + k < size_out) ? __ldg(he + (size_t)(q * size_out + j + k) * 32 + trial) : 0.f;
            e[u][q] = make_float4(v[0], v[1], v[2], v[3]);
        }
    }
    for (int t = 0; t < n_tiles; ++t) {
        const int b = t % SSB_PFC_NT;
        ssb_mbar_wait(&full[b], (uint32_t)(t / SSB_PFC_NT) & 1u);
        const int nn = min(SSB_PFC_NPT, cnt - t * SSB_PFC_NPT);
        for (int r = 0; r < nn; ++r) {
            float fq[K];
#pragma unroll
            for (int q = 0; q < K; ++q) fq[q] = fr[((size_t)(b * K + q) * SSB_PFC_NPT + r) * 32 + lane];      // lane = trial
            float* W = sm + (size_t)b * tile_f + (size_t)r * neuron_f;
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
                float4* wp = reinterpret_cast<float4*>(W + 128 * (warp + 8 * u) + 4 * lane);
                if (valid[u]) w = *wp;
#pragma unroll
                for (int q = 0; q < K; ++q) {
                    const float f = __shfl_sync(0xffffffffu, fq[q], tr_of[u]);
                    w.x = fmaf(e[u][q].x, f, w.x);
                    w.y = fmaf(e[u][q].y, f, w.y);
                    w.z = fmaf(e[u][q].z, f, w.z);
                    w.w = fmaf(e[u][q].w, f, w.w);
                }
                if (valid[u]) *wp = w;
            }
        }
        ssb_fence_async();
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(ssb_smem(&done[b])) : "memory");
    }
}

__global__ void __launch_bounds__(128) k_pes_clear(SsbCtx c, SsbPesDefer h, int i_rel, int force) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int r = blockIdx.x * 4 + warp;
    if (r < h.rows_e) ssb_grp(h.hist_e, h.rows_e, blockIdx.y, lane)[(size_t)r * 32] = 0.f;
}


// --------------------------------------------------------------------------------------
// Static decoders when every trial has its own network seed (per-trial static weights): no GEMM is left - every trial has
// its own Wd - so the decode takes the sparse form of k_pes_defer without the history: decoders live in the ldec arena in
// the same [neuron][trial][JP] layout, one WARP per trial, lanes = output columns, the trial's spikes compacted from the
// ensemble kernel's flag words (SSB_DPT_CH neurons at a time), 8 spikes in flight.  A warp walks ALL neurons of its
// ensemble, so there is no split-K and the result is written straight to the decoded rows.
// desc (the static-decoder descriptor): n size_out jpad act0 d_off out_vec - - -      (d_off = ldec arena row)
// grid (4 trial octets, G, decoders * column tiles of SSB_PES_JT) x 256
// SHARED = true: the same sparse walk over the SHARED static decoders ([neuron][jpad] rows of the weight array, d_off = float
// offset, L2-resident) - an alternative to the tensor-core decode for spiking runs (SSB_DECODE=sparse): ~36 KB of shared
// memory and no operand conversion, so it co-resides with the other kernels of the step instead of owning whole SMs.
#define SSB_DPT_CH 1024
template <bool SHARED>
__global__ void __launch_bounds__(256) k_decode_pt(SsbCtx c, const int* __restrict__ desc, int item0, int n_jt_max) {
    __shared__ int sflag[SSB_DPT_CH];
    __shared__ int slist[8][SSB_DPT_CH];
    __shared__ float red[SSB_PES_JT][8];
    const int item = blockIdx.z / n_jt_max, jt = blockIdx.z - item * n_jt_max;
    const int* d = desc + (item0 + item) * 9;
    const int n = d[0], size_out = d[1], act0 = d[3], d_off = d[4], out_vec = d[5];
    const int JP = SHARED ? d[2] : ssb_pes_jp(size_out);      // row length of one neuron's weights
    const int j0 = jt * SSB_PES_JT;
    if (j0 >= JP) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = blockIdx.y, oct = blockIdx.x;
    const int t = oct * 8 + warp;                         // this warp's trial inside the group
    const int* __restrict__ fl = c.aflag + (size_t)g * c.n_act + act0;
    const float* __restrict__ ap = c.act + ((size_t)g * c.n_act + act0) * 32 + t;
    const float* __restrict__ dl = SHARED ? c.W + d_off + j0 + lane
                                          : c.ldec + ((size_t)g * c.n_ldec + d_off) * 32 + (size_t)t * JP + j0 + lane;
    const size_t nstride = SHARED ? (size_t)JP : (size_t)JP * 32;      // floats between consecutive neurons
    constexpr int NC = SSB_PES_JT / 32;
    float acc[NC];
#pragma unroll
    for (int q = 0; q < NC; ++q) acc[q] = 0.f;
    int* list = slist[warp];
    for (int c0 = 0; c0 < n; c0 += SSB_DPT_CH) {
        const int c1 = min(n, c0 + SSB_DPT_CH);
        __syncthreads();                                  // the previous chunk's flag words are no longer read
        for (int i = c0 + (int)threadIdx.x; i < c1; i += 256) sflag[i - c0] = __ldg(fl + i);
        __syncthreads();
        int cnt = 0;
        for (int base = c0; base < c1; base += 32) {
            const int fw = (base + lane < c1) ? sflag[base + lane - c0] : 0;
            const bool on = (fw >> t) & 1;
            const unsigned m = __ballot_sync(0xffffffffu, on);
            if (on) list[cnt + __popc(m & ((1u << lane) - 1u))] = base + lane;
            cnt += __popc(m);
        }
        __syncwarp();
        constexpr int U = 8;
        for (int s0 = 0; s0 < cnt; s0 += U) {
            float a[U], w[U][NC];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const bool on = s0 + u < cnt;
                const size_t ni = (size_t)(on ? list[s0 + u] : c0);
                a[u] = on ? __ldg(ap + ni * 32) : 0.f;
#pragma unroll
                for (int q = 0; q < NC; ++q)
                    w[u][q] = (on && j0 + q * 32 + lane < JP) ? (SHARED ? __ldg(dl + ni * nstride + q * 32)
                                                                        : __ldcs(dl + ni * nstride + q * 32)) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
#pragma unroll
                for (int q = 0; q < NC; ++q) acc[q] = fmaf(w[u][q], a[u], acc[q]);
            }
        }
        __syncwarp();
    }
    // transpose through shared memory so that the decoded rows get 32-byte runs (8 consecutive trials per row)
#pragma unroll
    for (int q = 0; q < NC; ++q) red[q * 32 + lane][warp] = acc[q];
    __syncthreads();
    float* vo = c.vec + ((size_t)g * c.nv + out_vec) * 32 + oct * 8;
    for (int r = threadIdx.x >> 3; r < SSB_PES_JT; r += 32) {
        const int row = j0 + r;
        if (row < size_out) vo[(size_t)row * 32 + (threadIdx.x & 7)] = red[r][threadIdx.x & 7];
    }
}
