// sm_100a kernels of the SSP-SLAM step engine (device side).
//
// Layout rule: every per-trial array is [row][trial]; a warp's 32 lanes are 32
// consecutive trials, so state loads/stores are 128-byte coalesced and everything
// indexed by neuron / weight is warp-uniform (broadcast from L1/L2).  Static weights
// are shared by all trials; learned matrices (Voja encoders, PES decoders) are
// per-trial rows of the same [row][trial] form.
//
// Semantics restate nengo's operators, executed in
// dependency levels instead of one operator at a time; the CPU checker is
// oracle/nengo_ref_sim.py.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define SSB_TAB_BASE (1 << 24)
#define SSB_TOPK 4
#define SSB_SCAN_MAX_CHUNKS 256

struct SsbCtx {
    int B, nf, nt, n_probe, n_afilt;
    float dt;
    float* vec;          // [nv][B]   0: ones | 1..nf: filters A | nf+1..2nf: filters B | scratch
    const float* tab;    // [tab_cap][nt][B]
    float* v;            // [nn][B]
    float* ref;          // [nn][B]
    float* act;          // [n_act][B]
    float* lenc;         // [n_lenc][B]
    float* ldec;         // [n_ldec][B]
    float* afilt;        // [2][n_afilt][B]
    float* probe;        // [probe_cap][n_probe][B]
    const float* W;      // shared static weights
    const int* csr_ptr;
    const int* csr_idx;
    const float* csr_val;
    const float* ntypes; // [n][5] = type, tau_rc, tau_ref, min_voltage, amplitude
    const long long* dyn;  // [0] completed steps, [1] first step of resident tables, [2] first step of probe buffer
};

struct SsbStep {
    long long step;
    int par_old, par_new;     // row offset added to filter columns (0 or nf)
    const float* tabrow;      // table rows of this step
};

__device__ __forceinline__ SsbStep ssb_step(const SsbCtx& c) {
    SsbStep s;
    s.step = c.dyn[0];
    const int odd = (int)(s.step & 1);
    s.par_old = odd ? c.nf : 0;
    s.par_new = odd ? 0 : c.nf;
    s.tabrow = c.tab + (size_t)(s.step - c.dyn[1]) * (size_t)c.nt * (size_t)c.B;
    return s;
}

// One sink row: sparse linear combination of source columns for one trial.
// Warp-cooperative: all 32 lanes (= 32 trials) evaluate the same row, so the (column, coef)
// pairs are fetched 32 at a time with one coalesced load and broadcast by shuffle; the
// per-trial source loads of a batch are independent, so 8 of them are in flight at once.
// MUST be called by all 32 lanes of a warp with the same `row`.
__device__ __forceinline__ const float* ssb_src(const SsbCtx& c, const SsbStep& s, int idx, int par) {
    if (idx >= SSB_TAB_BASE) return s.tabrow + (size_t)(idx - SSB_TAB_BASE) * c.B;
    if (idx >= 1 && idx <= c.nf) idx += par;
    return c.vec + (size_t)idx * c.B;
}

__device__ __forceinline__ float ssb_row(const SsbCtx& c, const SsbStep& s, int row, int trial, int par) {
    const int lo = __ldg(c.csr_ptr + row), hi = __ldg(c.csr_ptr + row + 1);
    const int lane = threadIdx.x & 31;
    float acc = 0.f;
    for (int base = lo; base < hi; base += 32) {
        const int cnt = min(32, hi - base);
        int my_idx = 0;
        float my_val = 0.f;
        if (lane < cnt) {
            my_idx = __ldg(c.csr_idx + base + lane);
            my_val = __ldg(c.csr_val + base + lane);
        }
        int j = 0;
        for (; j + 8 <= cnt; j += 8) {
            float xv[8], cv[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int idx = __shfl_sync(0xffffffffu, my_idx, j + u);
                cv[u] = __shfl_sync(0xffffffffu, my_val, j + u);
                xv[u] = ssb_src(c, s, idx, par)[trial];
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) acc = fmaf(cv[u], xv[u], acc);
        }
        for (; j < cnt; ++j) {
            const int idx = __shfl_sync(0xffffffffu, my_idx, j);
            const float cv = __shfl_sync(0xffffffffu, my_val, j);
            acc = fmaf(cv, ssb_src(c, s, idx, par)[trial], acc);
        }
    }
    return acc;
}

struct SsbNeuron {
    int type;
    float tau_rc, tau_ref, min_v, amp_dt, amp, em1_full, dt;
};

__device__ __forceinline__ SsbNeuron ssb_neuron(const SsbCtx& c, int tid) {
    const float* p = c.ntypes + tid * 5;
    SsbNeuron n;
    n.type = (int)p[0];
    n.tau_rc = p[1];
    n.tau_ref = p[2];
    n.min_v = p[3];
    n.amp = p[4];
    n.dt = c.dt;
    n.amp_dt = p[4] / c.dt;
    n.em1_full = (n.type == 0) ? expm1f(-c.dt / p[1]) : 0.f;
    return n;
}

// nengo LIF.step / LIFRate.step / RectifiedLinear.step (App. A.4), fp32.
__device__ __forceinline__ float ssb_neuron_step(const SsbNeuron& n, float J, float& v, float& r) {
    if (n.type == 0) {
        r -= n.dt;
        const float delta = fminf(fmaxf(n.dt - r, 0.f), n.dt);
        const float em1 = (delta == n.dt) ? n.em1_full : expm1f(-delta / n.tau_rc);
        v = v - (J - v) * em1;
        float out = 0.f;
        if (v > 1.f) {
            const float t_spike = n.dt + n.tau_rc * log1pf(-(v - 1.f) / (J - 1.f));
            r = n.tau_ref + t_spike;
            v = 0.f;
            out = n.amp_dt;
        } else if (v < n.min_v) {
            v = n.min_v;
        }
        return out;
    } else if (n.type == 1) {
        const float j = J - 1.f;
        return j > 0.f ? n.amp / (n.tau_ref + n.tau_rc * log1pf(1.f / j)) : 0.f;
    }
    return n.amp * fmaxf(J, 0.f);
}

// --------------------------------------------------------------------------------------
// Narrow ensembles (VCO 3-D x 500, product squares 1-D x 50).  Input vector and decoded
// sums live in registers; packed per-neuron weights [bias, enc[DIMS], dec[nout]] are
// warp-uniform float4 loads.  One launch, two block ranges:
//   blocks [0, n_split*G)  "split":  a CTA of 4 warps owns one (ensemble, trial-group); warps take
//                                    interleaved neurons, partial decodes are reduced in shared memory
//                                    (a single warp streaming 500 neurons is a 150 us latency chain);
//   remaining blocks       "packed": each warp owns one (ensemble, trial-group) of a small ensemble.
// desc: n, dims, nout, state0, w_off, in_row0, out_vec, ntype, stride
template <int DIMS, int S4>
__device__ __forceinline__ void ssb_small_stream(const SsbCtx& c, const int* __restrict__ d, const SsbNeuron& nt,
                                                 const float (&x)[DIMS], int trial, int i_begin, int i_step,
                                                 float (&acc)[8]) {
    const int n = d[0], nout = d[2], state0 = d[3], w_off = d[4];
    const size_t B = c.B;
    const float4* __restrict__ w4 = reinterpret_cast<const float4*>(c.W + w_off);
    float* __restrict__ vp = c.v + (size_t)state0 * B + trial;
    float* __restrict__ rp = c.ref + (size_t)state0 * B + trial;
    const bool stateful = nt.type == 0;
    constexpr int U = 4;
    for (int i0 = i_begin; i0 < n; i0 += U * i_step) {
        float vv[U], rr[U], wl[U][4 * S4];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = i0 + u * i_step;
            vv[u] = 0.f;
            rr[u] = 0.f;
            if (i < n) {
                if (stateful) {
                    vv[u] = vp[(size_t)i * B];
                    rr[u] = rp[(size_t)i * B];
                }
#pragma unroll
                for (int q = 0; q < S4; ++q) {
                    const float4 t = __ldg(w4 + (size_t)i * S4 + q);
                    wl[u][4 * q + 0] = t.x;
                    wl[u][4 * q + 1] = t.y;
                    wl[u][4 * q + 2] = t.z;
                    wl[u][4 * q + 3] = t.w;
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = i0 + u * i_step;
            if (i < n) {
                float J = wl[u][0];
#pragma unroll
                for (int k = 0; k < DIMS; ++k) J = fmaf(wl[u][1 + k], x[k], J);
                const float out = ssb_neuron_step(nt, J, vv[u], rr[u]);
                if (stateful) {
                    vp[(size_t)i * B] = vv[u];
                    rp[(size_t)i * B] = rr[u];
                }
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (1 + DIMS + j < 4 * S4 && j < nout) acc[j] = fmaf(wl[u][1 + DIMS + j], out, acc[j]);
            }
        }
    }
}

template <int DIMS, int S4>
__device__ __forceinline__ void ssb_small_item(const SsbCtx& c, const SsbStep& s, const int* __restrict__ d, int trial,
                                               bool split, float* xs, float (*red)[8][32]) {
    const int nout = d[2], in_row0 = d[5], out_vec = d[6];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const SsbNeuron nt = ssb_neuron(c, d[7]);
    const size_t B = c.B;
    float x[DIMS], acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    if (split) {
        for (int k = warp; k < DIMS; k += 4) xs[k * 32 + lane] = ssb_row(c, s, in_row0 + k, trial, s.par_old);
        __syncthreads();
#pragma unroll
        for (int k = 0; k < DIMS; ++k) x[k] = xs[k * 32 + lane];
        ssb_small_stream<DIMS, S4>(c, d, nt, x, trial, warp, 4, acc);
#pragma unroll
        for (int j = 0; j < 8; ++j) red[warp][j][lane] = acc[j];
        __syncthreads();
        for (int j = warp; j < nout; j += 4) {
            const float t = (red[0][j][lane] + red[1][j][lane]) + (red[2][j][lane] + red[3][j][lane]);
            c.vec[(size_t)(out_vec + j) * B + trial] = t;
        }
    } else {
#pragma unroll
        for (int k = 0; k < DIMS; ++k) x[k] = ssb_row(c, s, in_row0 + k, trial, s.par_old);
        ssb_small_stream<DIMS, S4>(c, d, nt, x, trial, 0, 1, acc);
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (j < nout) c.vec[(size_t)(out_vec + j) * B + trial] = acc[j];
    }
}

__global__ void __launch_bounds__(128) k_ens_small(SsbCtx c, const int* __restrict__ desc, int n_items, int n_split,
                                                    int n_groups) {
    __shared__ float xs[4 * 32];
    __shared__ float red[4][8][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int split_blocks = n_split * n_groups;
    int item, group;
    bool split;
    if ((int)blockIdx.x < split_blocks) {
        split = true;
        item = blockIdx.x / n_groups;
        group = blockIdx.x - item * n_groups;
    } else {
        split = false;
        const int w = (blockIdx.x - split_blocks) * 4 + warp;
        if (w >= (n_items - n_split) * n_groups) return;
        item = n_split + w / n_groups;
        group = w % n_groups;
    }
    const int* d = desc + item * 9;
    const int trial = group * 32 + lane;
    const SsbStep s = ssb_step(c);
    const int key = d[1] * 8 + (d[8] >> 2);
    switch (key) {
#define SSB_CASE(D, S) \
    case (D) * 8 + (S): ssb_small_item<D, S>(c, s, d, trial, split, xs, red); break;
        SSB_CASE(1, 1) SSB_CASE(1, 2) SSB_CASE(1, 3)
        SSB_CASE(2, 1) SSB_CASE(2, 2) SSB_CASE(2, 3)
        SSB_CASE(3, 1) SSB_CASE(3, 2) SSB_CASE(3, 3)
        SSB_CASE(4, 2) SSB_CASE(4, 3) SSB_CASE(4, 4)
#undef SSB_CASE
        default: break;  // excluded by the host-side lowering (dims <= 4, dims + nout <= 11)
    }
}

// --------------------------------------------------------------------------------------
// Wide ensembles (OVC / memory / recall / error: 970 x 55).  A CTA owns (ensemble, trial-group,
// neuron chunk); the input vector is staged once in shared memory and (for the templated
// widths) copied to registers.  Each warp walks its neurons two at a time.  Output
// activities go to act[n][trial] for the decode / PES kernels.  Voja-learned encoders are
// per-trial rows (lenc), all loads of a row are issued before use, and rows that spiked are
// updated in place (post_synapse=None => the delta is row-sparse).
// desc: n dims dpad state0 act0 enc_off bias_off in_row0 ntype flags jn_row0 jn_m jn_w voja_row scale_off alpha_bits
template <int DP>
__device__ __forceinline__ void ssb_wide_neuron(const SsbCtx& c, const int* __restrict__ d, const SsbNeuron& nt, int i,
                                                int trial, const float* xs, const float* us, const float (&x)[DP > 0 ? DP : 1],
                                                float aL) {
    const int dims = d[1], dpad = d[2], state0 = d[3], act0 = d[4], enc_off = d[5], bias_off = d[6], flags = d[9];
    const int jn_m = d[11], jn_w = d[12], scale_off = d[14];
    const int lane = threadIdx.x & 31;
    const size_t B = c.B;
    const bool voja = flags & 1, stateful = nt.type == 0;
    const size_t so = (size_t)(state0 + i) * B + trial;
    float v = 0.f, r = 0.f;
    if (stateful) {
        v = c.v[so];
        r = c.ref[so];
    }
    float J = __ldg(c.W + bias_off + i);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    float* erow = nullptr;
    float ev[DP > 0 ? DP : 1];
    if (!voja) {
        const float4* __restrict__ e4 = reinterpret_cast<const float4*>(c.W + enc_off + (size_t)i * dpad);
        if (DP > 0) {
#pragma unroll
            for (int k4 = 0; k4 < DP / 4; ++k4) {
                const float4 e = __ldg(e4 + k4);
                a0 = fmaf(e.x, x[4 * k4 + 0], a0);
                a1 = fmaf(e.y, x[4 * k4 + 1], a1);
                a2 = fmaf(e.z, x[4 * k4 + 2], a2);
                a3 = fmaf(e.w, x[4 * k4 + 3], a3);
            }
        } else {
            for (int k4 = 0; k4 < (dpad >> 2); ++k4) {
                const float4 e = __ldg(e4 + k4);
                const float* xk = xs + (k4 * 4) * 32 + lane;
                a0 = fmaf(e.x, xk[0], a0);
                a1 = fmaf(e.y, xk[32], a1);
                a2 = fmaf(e.z, xk[64], a2);
                a3 = fmaf(e.w, xk[96], a3);
            }
        }
    } else {
        erow = c.lenc + ((size_t)enc_off + (size_t)i * dims) * B + trial;
        if (DP > 0) {
#pragma unroll
            for (int k = 0; k < DP; ++k) ev[k] = (k < dims) ? erow[(size_t)k * B] : 0.f;
#pragma unroll
            for (int k = 0; k < DP; k += 4) {
                a0 = fmaf(ev[k + 0], x[k + 0], a0);
                a1 = fmaf(ev[k + 1], x[k + 1], a1);
                a2 = fmaf(ev[k + 2], x[k + 2], a2);
                a3 = fmaf(ev[k + 3], x[k + 3], a3);
            }
        } else {
            int k = 0;
            for (; k + 4 <= dims; k += 4) {
                const float e0 = erow[(size_t)k * B], e1 = erow[(size_t)(k + 1) * B];
                const float e2 = erow[(size_t)(k + 2) * B], e3 = erow[(size_t)(k + 3) * B];
                a0 = fmaf(e0, xs[k * 32 + lane], a0);
                a1 = fmaf(e1, xs[(k + 1) * 32 + lane], a1);
                a2 = fmaf(e2, xs[(k + 2) * 32 + lane], a2);
                a3 = fmaf(e3, xs[(k + 3) * 32 + lane], a3);
            }
            for (; k < dims; ++k) a0 = fmaf(erow[(size_t)k * B], xs[k * 32 + lane], a0);
        }
    }
    J += (a0 + a1) + (a2 + a3);
    for (int m = 0; m < jn_m; ++m) J = fmaf(__ldg(c.W + jn_w + i * jn_m + m), us[m * 32 + lane], J);
    const float out = ssb_neuron_step(nt, J, v, r);
    if (stateful) {
        c.v[so] = v;
        c.ref[so] = r;
    }
    c.act[(size_t)(act0 + i) * B + trial] = out;
    if (voja && out != 0.f) {
        // SimVoja: delta = alpha*L*(scale*outer(post, x) - post[:,None]*E), applied to E for the next step
        const float sc = __ldg(c.W + scale_off + i);
        if (DP > 0) {
#pragma unroll
            for (int k = 0; k < DP; ++k)
                if (k < dims) erow[(size_t)k * B] = ev[k] + aL * (sc * (out * x[k]) - out * ev[k]);
        } else {
            for (int k = 0; k < dims; ++k) {
                const float e = erow[(size_t)k * B];
                erow[(size_t)k * B] = e + aL * (sc * (out * xs[k * 32 + lane]) - out * e);
            }
        }
    }
}

template <int DP>
__global__ void __launch_bounds__(128) k_ens_wide(SsbCtx c, const int* __restrict__ desc, int item0, int chunk) {
    extern __shared__ float sm[];
    const int* d = desc + (item0 + blockIdx.z) * 16;
    const int n = d[0], dims = d[1], dpad = d[2];
    const int in_row0 = d[7], flags = d[9], jn_row0 = d[10], jn_m = d[11], voja_row = d[13];
    const int n0 = blockIdx.x * chunk;
    if (n0 >= n) return;
    if (DP > 0 && dpad != DP) return;  // this instantiation only serves ensembles of its width
    if (DP == 0 && (dpad == 56 || dpad == 100)) return;
    const int n1 = min(n, n0 + chunk);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int trial = blockIdx.y * 32 + lane;
    const SsbStep s = ssb_step(c);
    const SsbNeuron nt = ssb_neuron(c, d[8]);
    float* xs = sm;                 // [dpad][32]
    float* us = sm + dpad * 32;     // [jn_m][32]
    for (int k = warp; k < dpad; k += nwarps)
        xs[k * 32 + lane] = (k < dims) ? ssb_row(c, s, in_row0 + k, trial, s.par_old) : 0.f;
    for (int m = warp; m < jn_m; m += nwarps) us[m * 32 + lane] = ssb_row(c, s, jn_row0 + m, trial, s.par_old);
    float aL = 0.f;
    if (flags & 1) aL = __int_as_float(d[15]) * ssb_row(c, s, voja_row, trial, s.par_old);
    __syncthreads();
    float x[DP > 0 ? DP : 1];
    if (DP > 0) {
#pragma unroll
        for (int k = 0; k < DP; ++k) x[k] = xs[k * 32 + lane];
    }
    for (int i = n0 + warp; i < n1; i += nwarps) ssb_wide_neuron<DP>(c, d, nt, i, trial, xs, us, x, aL);
}

// --------------------------------------------------------------------------------------
// Static decoders of wide ensembles: out[j][trial] = sum_n Wd[n][j] * act[n][trial].
// CTA = (decoder, trial-group, 8-row output tile, neuron chunk); warps split the chunk, shared-
// memory reduce; each neuron chunk writes its own partial slot (the consumers' CSR rows sum the
// partial slots, so there are no atomics and the result is deterministic).
// desc: n size_out jpad act0 w_off out_vec n_chunks
__global__ void __launch_bounds__(128) k_decode(SsbCtx c, const int* __restrict__ desc, int item0, int max_chunks) {
    __shared__ float red[4][8][32];
    const int item = blockIdx.z / max_chunks, chunk = blockIdx.z - item * max_chunks;
    const int* d = desc + (item0 + item) * 7;
    const int n = d[0], size_out = d[1], jpad = d[2], act0 = d[3], w_off = d[4], out_vec = d[5], n_chunks = d[6];
    const int j0 = blockIdx.x * 8;
    if (j0 >= size_out || chunk >= n_chunks) return;
    const int per = (n + n_chunks - 1) / n_chunks;
    const int i_lo = chunk * per, i_hi = min(n, i_lo + per);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int trial = blockIdx.y * 32 + lane;
    const size_t B = c.B;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    const float* __restrict__ ap = c.act + (size_t)act0 * B + trial;
    int i = i_lo + warp;
    for (; i + 4 < i_hi; i += 8) {
        const float a0 = ap[(size_t)i * B], a1 = ap[(size_t)(i + 4) * B];
        const float4* __restrict__ w0 = reinterpret_cast<const float4*>(c.W + w_off + (size_t)i * jpad + j0);
        const float4* __restrict__ w1 = reinterpret_cast<const float4*>(c.W + w_off + (size_t)(i + 4) * jpad + j0);
        const float4 wa = __ldg(w0), wb = __ldg(w0 + 1), wc = __ldg(w1), wd = __ldg(w1 + 1);
        acc[0] = fmaf(wa.x, a0, acc[0]);
        acc[1] = fmaf(wa.y, a0, acc[1]);
        acc[2] = fmaf(wa.z, a0, acc[2]);
        acc[3] = fmaf(wa.w, a0, acc[3]);
        acc[4] = fmaf(wb.x, a0, acc[4]);
        acc[5] = fmaf(wb.y, a0, acc[5]);
        acc[6] = fmaf(wb.z, a0, acc[6]);
        acc[7] = fmaf(wb.w, a0, acc[7]);
        acc[0] = fmaf(wc.x, a1, acc[0]);
        acc[1] = fmaf(wc.y, a1, acc[1]);
        acc[2] = fmaf(wc.z, a1, acc[2]);
        acc[3] = fmaf(wc.w, a1, acc[3]);
        acc[4] = fmaf(wd.x, a1, acc[4]);
        acc[5] = fmaf(wd.y, a1, acc[5]);
        acc[6] = fmaf(wd.z, a1, acc[6]);
        acc[7] = fmaf(wd.w, a1, acc[7]);
    }
    for (; i < i_hi; i += 4) {
        const float a0 = ap[(size_t)i * B];
        const float4* __restrict__ w0 = reinterpret_cast<const float4*>(c.W + w_off + (size_t)i * jpad + j0);
        const float4 wa = __ldg(w0), wb = __ldg(w0 + 1);
        acc[0] = fmaf(wa.x, a0, acc[0]);
        acc[1] = fmaf(wa.y, a0, acc[1]);
        acc[2] = fmaf(wa.z, a0, acc[2]);
        acc[3] = fmaf(wa.w, a0, acc[3]);
        acc[4] = fmaf(wb.x, a0, acc[4]);
        acc[5] = fmaf(wb.y, a0, acc[5]);
        acc[6] = fmaf(wb.z, a0, acc[6]);
        acc[7] = fmaf(wb.w, a0, acc[7]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) red[warp][j][lane] = acc[j];
    __syncthreads();
    for (int j = warp; j < 8; j += 4) {
        if (j0 + j < size_out) {
            const float t = (red[0][j][lane] + red[1][j][lane]) + (red[2][j][lane] + red[3][j][lane]);
            c.vec[(size_t)(out_vec + chunk * size_out + j0 + j) * B + trial] = t;
        }
    }
}

// --------------------------------------------------------------------------------------
// PES-learned decoders (per trial): one streaming pass that applies the pending rank-1
// delta, decodes with the updated weights and writes them back:
//   D <- D + outer(alpha*err_prev, a_prev)     (nengo: Copy(delta->weights, inc) at step start)
//   out = D . act                               (DotInc)
// err_prev / a_prev are the filter values the previous step read (the not-yet-overwritten
// half of the ping-pong buffers), which is exactly SimPES' delta from the previous step.
// This is the dominant HBM stream of the SLAM step (8 bytes per learned weight per trial-step).
// desc: n size_out d_off a_off act0 err_row0 out_vec alpha_bits decay_bits onemdecay_bits n_chunks
__global__ void __launch_bounds__(128) k_pes(SsbCtx c, const int* __restrict__ desc, int max_chunks) {
    __shared__ float red[4][8][32];
    const int item = blockIdx.z / max_chunks, chunk = blockIdx.z - item * max_chunks;
    const int* d = desc + item * 11;
    const int n = d[0], size_out = d[1], d_off = d[2], a_off = d[3], act0 = d[4], err_row0 = d[5], out_vec = d[6];
    const int n_chunks = d[10];
    const float alpha = __int_as_float(d[7]);
    const int j0 = blockIdx.x * 8;
    if (j0 >= size_out || chunk >= n_chunks) return;
    const int per = (n + n_chunks - 1) / n_chunks;
    const int i_lo = chunk * per, i_hi = min(n, i_lo + per);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int trial = blockIdx.y * 32 + lane;
    const size_t B = c.B;
    const SsbStep s = ssb_step(c);
    const int prev_buf = 1 - (int)(s.step & 1);  // afilt half that still holds what the previous step read
    float ae[8], acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        acc[j] = 0.f;
        float e = 0.f;
        if (j0 + j < size_out) e = ssb_row(c, s, err_row0 + j0 + j, trial, s.par_new);
        ae[j] = s.step > 0 ? alpha * e : 0.f;
    }
    const float* __restrict__ ap = c.act + (size_t)act0 * B + trial;
    const float* __restrict__ fp = c.afilt + ((size_t)prev_buf * c.n_afilt + a_off) * B + trial;
    float* __restrict__ dp = c.ldec + (size_t)d_off * B + trial;
    const int jn = min(8, size_out - j0);
    if (jn == 8) {
        int i = i_lo + warp;
        for (; i + 4 < i_hi; i += 8) {
            float w0[8], w1[8];
            const float a0 = ap[(size_t)i * B], f0 = fp[(size_t)i * B];
            const float a1 = ap[(size_t)(i + 4) * B], f1 = fp[(size_t)(i + 4) * B];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                w0[j] = dp[((size_t)(j0 + j) * n + i) * B];
                w1[j] = dp[((size_t)(j0 + j) * n + i + 4) * B];
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                w0[j] = fmaf(ae[j], f0, w0[j]);
                w1[j] = fmaf(ae[j], f1, w1[j]);
                acc[j] = fmaf(w0[j], a0, acc[j]);
                acc[j] = fmaf(w1[j], a1, acc[j]);
                dp[((size_t)(j0 + j) * n + i) * B] = w0[j];
                dp[((size_t)(j0 + j) * n + i + 4) * B] = w1[j];
            }
        }
        for (; i < i_hi; i += 4) {
            float w0[8];
            const float a0 = ap[(size_t)i * B], f0 = fp[(size_t)i * B];
#pragma unroll
            for (int j = 0; j < 8; ++j) w0[j] = dp[((size_t)(j0 + j) * n + i) * B];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                w0[j] = fmaf(ae[j], f0, w0[j]);
                acc[j] = fmaf(w0[j], a0, acc[j]);
                dp[((size_t)(j0 + j) * n + i) * B] = w0[j];
            }
        }
    } else {
        for (int i = i_lo + warp; i < i_hi; i += 4) {
            const float a0 = ap[(size_t)i * B], f0 = fp[(size_t)i * B];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (j < jn) {
                    float w = dp[((size_t)(j0 + j) * n + i) * B];
                    w = fmaf(ae[j], f0, w);
                    acc[j] = fmaf(w, a0, acc[j]);
                    dp[((size_t)(j0 + j) * n + i) * B] = w;
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) red[warp][j][lane] = acc[j];
    __syncthreads();
    for (int j = warp; j < 8; j += 4) {
        if (j0 + j < size_out) {
            const float t = (red[0][j][lane] + red[1][j][lane]) + (red[2][j][lane] + red[3][j][lane]);
            c.vec[(size_t)(out_vec + chunk * size_out + j0 + j) * B + trial] = t;
        }
    }
}

// --------------------------------------------------------------------------------------
// Grid clean-up / decode: argmax_g S[g].x with first-maximum-wins.  Scan in fp32 keeping the
// top-4 candidates per (grid chunk, trial); the pick kernel re-scores near-ties in fp64 so
// that the chosen index equals the float64 NumPy argmax on the same input.
// A CTA = 4 warps = 4 different trial groups scanning the SAME grid chunk: the chunk of S is
// staged once in shared memory and read back as warp-uniform (broadcast) float4s, the query
// vector sits in registers.
struct SsbTop {
    float v[SSB_TOPK];
    int g[SSB_TOPK];
};

__device__ __forceinline__ void ssb_top_init(SsbTop& t) {
#pragma unroll
    for (int i = 0; i < SSB_TOPK; ++i) {
        t.v[i] = -INFINITY;
        t.g[i] = 0x7fffffff;
    }
}

// keep sorted by (value desc, index asc); candidates arrive in ascending g
__device__ __forceinline__ void ssb_top_push(SsbTop& t, float val, int g) {
    if (val > t.v[SSB_TOPK - 1]) {
#pragma unroll
        for (int i = SSB_TOPK - 1; i >= 0; --i) {
            const bool shift = (i > 0) && (val > t.v[i - 1]);
            if (shift) {
                t.v[i] = t.v[i - 1];
                t.g[i] = t.g[i - 1];
            } else {
                t.v[i] = val;
                t.g[i] = g;
                break;
            }
        }
    }
}

// desc: G d dpad s_off in_row0 out_vec ; scratch rows: cx[dpad][B], pval/pidx[n_chunks*TOPK][B]
// A CTA owns grid rows [blockIdx.x*rows_per_chunk, +rows_per_chunk) and walks them in shared-memory
// tiles of tile_rows rows.  dynamic smem: tile_rows*dpad (S tile) + 4*dpad*32 (x staging, generic width only)
template <int DP, bool CSR_INPUT>
__global__ void __launch_bounds__(128)
k_cleanup_scan(SsbCtx c, const int* __restrict__ d, const float* __restrict__ S, float* __restrict__ cx,
               float* __restrict__ pval, int* __restrict__ pidx, int rows_per_chunk, int tile_rows, int n_groups) {
    extern __shared__ float sm[];
    const int G = d[0], dims = d[1], dpad = d[2], in_row0 = d[4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int group = blockIdx.y * 4 + warp;
    const bool live = group < n_groups;
    const int trial = (live ? group : 0) * 32 + lane;
    const size_t B = c.B;
    const int g_lo = blockIdx.x * rows_per_chunk;
    const int g_hi = min(G, g_lo + rows_per_chunk);
    float* tile = sm;                              // [tile_rows][dpad]
    float* xs = sm + (size_t)tile_rows * dpad;     // [4][dpad][32] (generic width only)
    float x[DP > 0 ? DP : 1];
    if (CSR_INPUT) {
        const SsbStep s = ssb_step(c);
        if (DP > 0) {
#pragma unroll
            for (int k = 0; k < DP; ++k) {
                x[k] = (k < dims) ? ssb_row(c, s, in_row0 + k, trial, s.par_old) : 0.f;
                if (blockIdx.x == 0 && live) cx[(size_t)k * B + trial] = x[k];
            }
        } else {
            for (int k = 0; k < dpad; ++k) {
                const float xv = (k < dims) ? ssb_row(c, s, in_row0 + k, trial, s.par_old) : 0.f;
                xs[(warp * dpad + k) * 32 + lane] = xv;
                if (blockIdx.x == 0 && live) cx[(size_t)k * B + trial] = xv;
            }
        }
    } else {
        if (DP > 0) {
#pragma unroll
            for (int k = 0; k < DP; ++k) x[k] = (k < dims) ? cx[(size_t)k * B + trial] : 0.f;
        } else {
            for (int k = 0; k < dpad; ++k) xs[(warp * dpad + k) * 32 + lane] = (k < dims) ? cx[(size_t)k * B + trial] : 0.f;
        }
    }
    SsbTop top;
    ssb_top_init(top);
    for (int g0 = g_lo; g0 < g_hi; g0 += tile_rows) {
        const int g1 = min(g_hi, g0 + tile_rows);
        __syncthreads();   // previous tile fully consumed (and xs visible on the first pass)
        {   // stage the grid tile (coalesced float4)
            const float4* __restrict__ src = reinterpret_cast<const float4*>(S + (size_t)g0 * dpad);
            float4* dst = reinterpret_cast<float4*>(tile);
            const int n4 = (g1 - g0) * (dpad >> 2);
            for (int i = threadIdx.x; i < n4; i += blockDim.x) dst[i] = __ldg(src + i);
        }
        __syncthreads();
        if (!live) continue;
        if (DP > 0) {
            for (int g = g0; g < g1; ++g) {
                const float4* s4 = reinterpret_cast<const float4*>(tile + (size_t)(g - g0) * DP);
                float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
                for (int k4 = 0; k4 < DP / 4; ++k4) {
                    const float4 e = s4[k4];
                    a0 = fmaf(e.x, x[4 * k4 + 0], a0);
                    a1 = fmaf(e.y, x[4 * k4 + 1], a1);
                    a2 = fmaf(e.z, x[4 * k4 + 2], a2);
                    a3 = fmaf(e.w, x[4 * k4 + 3], a3);
                }
                ssb_top_push(top, (a0 + a1) + (a2 + a3), g);
            }
        } else {
            const float* xw = xs + (size_t)warp * dpad * 32 + lane;
            for (int g = g0; g < g1; ++g) {
                const float4* s4 = reinterpret_cast<const float4*>(tile + (size_t)(g - g0) * dpad);
                float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
                for (int k4 = 0; k4 < (dpad >> 2); ++k4) {
                    const float4 e = s4[k4];
                    const float* xk = xw + (k4 * 4) * 32;
                    a0 = fmaf(e.x, xk[0], a0);
                    a1 = fmaf(e.y, xk[32], a1);
                    a2 = fmaf(e.z, xk[64], a2);
                    a3 = fmaf(e.w, xk[96], a3);
                }
                ssb_top_push(top, (a0 + a1) + (a2 + a3), g);
            }
        }
    }
    if (!live) return;
#pragma unroll
    for (int i = 0; i < SSB_TOPK; ++i) {
        pval[(size_t)(blockIdx.x * SSB_TOPK + i) * B + trial] = top.v[i];
        pidx[(size_t)(blockIdx.x * SSB_TOPK + i) * B + trial] = top.g[i];
    }
}

// CTA = 32 trials x 8 warps: warps split the candidate list, merge through shared memory, then
// candidates within eps of the fp32 maximum are re-scored in fp64 (S64 is the float64 grid) and
// the winning index / grid row are written.
__global__ void __launch_bounds__(256)
k_cleanup_pick(int B, int dims, int dpad, int ncand, const float* __restrict__ cx, const float* __restrict__ pval,
               const int* __restrict__ pidx, const double* __restrict__ S64, const float* __restrict__ S32,
               float* __restrict__ out_rows, int* __restrict__ out_idx, const double* __restrict__ q64, long long q0,
               long long n_q) {
    __shared__ float sv[8][32];
    __shared__ int sg[8][32];
    __shared__ float sn[8][32];
    __shared__ int sc[8][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int trial = blockIdx.x * 32 + lane;
    if (trial >= B) return;  // B is a multiple of 32: whole CTAs exit together
    float best = -INFINITY;
    int best_g = 0x7fffffff;
    for (int i = warp; i < ncand; i += 8) {
        const float v = pval[(size_t)i * B + trial];
        const int g = pidx[(size_t)i * B + trial];
        if (v > best || (v == best && g < best_g)) {
            best = v;
            best_g = g;
        }
    }
    float xn = 0.f;
    for (int k = warp; k < dims; k += 8) {
        const float xv = cx[(size_t)k * B + trial];
        xn = fmaf(xv, xv, xn);
    }
    sv[warp][lane] = best;
    sg[warp][lane] = best_g;
    sn[warp][lane] = xn;
    __syncthreads();
    best = sv[0][lane];
    best_g = sg[0][lane];
    xn = sn[0][lane];
    for (int w = 1; w < 8; ++w) {
        const float v = sv[w][lane];
        const int g = sg[w][lane];
        if (v > best || (v == best && g < best_g)) {
            best = v;
            best_g = g;
        }
        xn += sn[w][lane];
    }
    // fp32 dot-product error bound: ~dims * 2^-24 * |S_g||x| with |S_g| = 1
    const float eps = 4.0f * (float)dims * 5.97e-8f * sqrtf(xn) + 1e-30f;
    int n_close = 0;
    for (int i = warp; i < ncand; i += 8)
        if (pval[(size_t)i * B + trial] >= best - eps) ++n_close;
    sc[warp][lane] = n_close;
    __syncthreads();
    n_close = 0;
    for (int w = 0; w < 8; ++w) n_close += sc[w][lane];
    if (n_close > 1 && S64 != nullptr) {
        // rare path (a near-tie): every warp redundantly re-scores the same candidates for its lane's trial
        double dbest = -1e300;
        int dg = 0x7fffffff;
        for (int i = 0; i < ncand; ++i) {
            if (pval[(size_t)i * B + trial] >= best - eps) {
                const int g = pidx[(size_t)i * B + trial];
                if (g == 0x7fffffff) continue;
                const double* sgp = S64 + (size_t)g * dims;
                double acc = 0.0;
                // argmax is invariant to the positive normalisation, so the raw float64 query can be used
                if (q64 != nullptr && q0 + trial < n_q) {
                    const double* qr = q64 + (size_t)(q0 + trial) * dims;
                    for (int k = 0; k < dims; ++k) acc += sgp[k] * qr[k];
                } else {
                    for (int k = 0; k < dims; ++k) acc += sgp[k] * (double)cx[(size_t)k * B + trial];
                }
                if (acc > dbest || (acc == dbest && g < dg)) {
                    dbest = acc;
                    dg = g;
                }
            }
        }
        best_g = dg;
    }
    if (out_idx && warp == 0) out_idx[trial] = best_g;
    if (out_rows) {
        const float* sgp = S32 + (size_t)best_g * dpad;
        for (int k = warp; k < dims; k += 8) out_rows[(size_t)k * B + trial] = sgp[k];
    }
}

// --------------------------------------------------------------------------------------
// Gated correction node (slam.py:233-237): x = [p ; q ; flag].  CTA = one trial group x 8 warps;
// warps split the dimensions, the dot product is reduced through shared memory.
// desc: d in_row0 out_vec rate_bits thres_bits atol_bits
__global__ void __launch_bounds__(256) k_gate(SsbCtx c, const int* __restrict__ desc, int item0) {
    __shared__ float part[8][32];
    const int* d = desc + (item0 + blockIdx.y) * 6;
    const int dims = d[0], in_row0 = d[1], out_vec = d[2];
    const float rate = __int_as_float(d[3]), thres = __int_as_float(d[4]), atol = __int_as_float(d[5]);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int trial = blockIdx.x * 32 + lane;
    const size_t B = c.B;
    const SsbStep s = ssb_step(c);
    // pass 1: p - q goes to the output slot, p.q is reduced over the 8 warps
    float dot = 0.f;
    for (int k = warp; k < dims; k += 8) {
        const float p = ssb_row(c, s, in_row0 + k, trial, s.par_old);
        const float q = ssb_row(c, s, in_row0 + dims + k, trial, s.par_old);
        dot = fmaf(p, q, dot);
        c.vec[(size_t)(out_vec + k) * B + trial] = p - q;
    }
    part[warp][lane] = dot;
    __syncthreads();
    dot = 0.f;
    for (int w = 0; w < 8; ++w) dot += part[w][lane];
    const float flag = ssb_row(c, s, in_row0 + 2 * dims, trial, s.par_old);
    const bool open = (fabsf(flag) <= atol) && (dot > thres);
    // pass 2: each thread rescales the values it wrote itself
    for (int k = warp; k < dims; k += 8) {
        float* o = c.vec + (size_t)(out_vec + k) * B + trial;
        *o = open ? rate * *o : 0.f;
    }
}

// --------------------------------------------------------------------------------------
// End-of-step rows: Lowpass updates (y_new = a*y_old + b*u, written to the other half of
// the ping-pong buffer = nengo's update-after-read), probe samples, PES activity traces.
// rows: [csr_row | act_row, kind, dst]; kind 0 filter, 1 probe, 2 activity trace
__global__ void __launch_bounds__(128) k_lin(SsbCtx c, const int* __restrict__ rows, const float* __restrict__ ab, int n_rows) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int r = blockIdx.x * 4 + warp;
    if (r >= n_rows) return;
    const int trial = blockIdx.y * 32 + lane;
    const size_t B = c.B;
    const SsbStep s = ssb_step(c);
    const int src = rows[r * 3], kind = rows[r * 3 + 1], dst = rows[r * 3 + 2];
    const float a = ab[r * 2], b = ab[r * 2 + 1];
    if (kind == 0) {
        const float u = ssb_row(c, s, src, trial, s.par_old);
        const float y = c.vec[(size_t)(1 + dst + s.par_old) * B + trial];
        c.vec[(size_t)(1 + dst + s.par_new) * B + trial] = fmaf(b, u, a * y);
    } else if (kind == 1) {
        const float u = ssb_row(c, s, src, trial, s.par_old);
        c.probe[((size_t)(s.step - c.dyn[2]) * c.n_probe + dst) * B + trial] = u;
    } else {
        const int old_buf = (int)(s.step & 1);
        const float y = c.afilt[((size_t)old_buf * c.n_afilt + dst) * B + trial];
        const float u = c.act[(size_t)src * B + trial];
        c.afilt[((size_t)(1 - old_buf) * c.n_afilt + dst) * B + trial] = fmaf(b, u, a * y);
    }
}

__global__ void k_advance(long long* dyn) { dyn[0] += 1; }

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
    const double w = 6.283185307179586476925286766559 / (double)d;
    for (int m = threadIdx.x; m < d; m += blockDim.x) {
        double acc = 0.0;
        for (int k = 0; k < d; ++k) {
            // exp(i*theta_k) * exp(+2 pi i k m / d); reduce k*m mod d to keep the angle small
            const int km = (int)(((long long)k * m) % d);
            double sn, cn;
            sincospi(2.0 * (double)km / (double)d, &sn, &cn);
            acc += cs[k] * cn - cs[d + k] * sn;
        }
        (void)w;
        out[(size_t)p * d + m] = acc / (double)d;
    }
}

// Normalise query rows (skip if norm < 1e-6) and transpose to [k][N_pad] float for the scan.
__global__ void k_decode_prep(const double* __restrict__ q, float* __restrict__ cx, long long n_q, int B, int d, int dpad,
                              long long q0) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B) return;
    const long long row = q0 + t;
    if (row >= n_q) {
        for (int k = 0; k < dpad; ++k) cx[(size_t)k * B + t] = 0.f;
        return;
    }
    double nrm = 0.0;
    for (int k = 0; k < d; ++k) nrm += q[(size_t)row * d + k] * q[(size_t)row * d + k];
    nrm = sqrt(nrm);
    const double sc = nrm < 1e-6 ? 1.0 : 1.0 / nrm;
    for (int k = 0; k < dpad; ++k) cx[(size_t)k * B + t] = k < d ? (float)(q[(size_t)row * d + k] * sc) : 0.f;
}
