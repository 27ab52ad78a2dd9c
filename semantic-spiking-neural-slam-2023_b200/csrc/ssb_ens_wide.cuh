// sm_100a kernels of the SSP-SLAM step engine: wide ensembles: k_wide_static, k_wide_static_tc (tcgen05), k_wide_voja (per-trial learned encoders).
// Included by ssb_kernels.cuh (after ssb_common.cuh); see that file for the layout rules.
#pragma once
#include "ssb_common.cuh"

// --------------------------------------------------------------------------------------
// Wide ensembles (OVC / memory / recall / error: 970 x 55).  A CTA owns (ensemble, trial group,
// chunk of neurons).  Everything the chunk needs is contiguous in memory and is staged in shared
// memory by TMA bulk copies issued by one thread while all warps evaluate the input vector:
// static encoders [chunk][dpad], bias, direct-current weights, the chunk's 128-byte state rows.
// The input vector is copied to registers (templated widths), each warp walks its quarter of the
// chunk with broadcast float4 encoder reads, and the updated state goes back with a bulk store.
// Output activities go to act[n] for the decode / PES kernels.
// desc: n dims dpad state0 act0 enc_off bias_off in_row0 ntype flags jn_row0 jn_m jn_w voja_row scale_off alpha_bits
struct SsbItemList {
    int n;
    int idx[15];
};

__device__ __forceinline__ int ssb_r4(int x) { return (x + 3) & ~3; }

// Input rows of a wide ensemble -> shared memory [dpad][32].  Each warp takes every nwarps-th row, eight rows per batch
// so that the (L2-resident) loads of a batch are in flight together instead of one dependent load per store.
__device__ __forceinline__ void ssb_stage_rows(float* xs, const float* vg, int row0, int dims, int dpad, int warp,
                                               int nwarps, int lane) {
    for (int k0 = warp; k0 < dpad; k0 += nwarps * 8) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int k = k0 + u * nwarps;
            v[u] = (k < dims) ? vg[(size_t)(row0 + k) * 32] : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int k = k0 + u * nwarps;
            if (k < dpad) xs[k * 32 + lane] = v[u];
        }
    }
}


template <int DP>
__global__ void __launch_bounds__(128) k_wide_static(SsbCtx c, const int* __restrict__ desc, SsbItemList items, int chunk,
                                                      int i_rel) {
    extern __shared__ __align__(128) float sm[];
    __shared__ unsigned long long bar;
    const int* d = desc + items.idx[blockIdx.z] * 16;
    const int n = d[0], dims = d[1], dpad = d[2], state0 = d[3], act0 = d[4], enc_off = d[5], bias_off = d[6];
    const int in_row0 = d[7], jn_row0 = d[10], jn_m = d[11], jn_w = d[12];
    const int n0 = blockIdx.x * chunk;
    if (n0 >= n) return;
    const int cnt = min(chunk, n - n0);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = blockIdx.y;
    const SsbNeuron nt = ssb_neuron(c, d[8]);
    const bool stateful = nt.type == 0;
    float* s_enc = sm;                                  // [chunk][dpad]
    float* s_bias = s_enc + (size_t)chunk * dpad;       // [chunk]
    float* s_jn = s_bias + chunk;                       // [chunk][jn_m]
    float* s_st = s_jn + (size_t)chunk * jn_m;          // [chunk][32]
    float* xs = s_st + (size_t)chunk * 32;              // [dpad][32]
    float* us = xs + (size_t)dpad * 32;                 // [jn_m][32]
    float* stg = c.st + ((size_t)g * c.nn + state0 + n0) * 32;
    if (threadIdx.x == 0) {
        ssb_mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const uint32_t b_enc = (uint32_t)cnt * dpad * 4, b_bias = (uint32_t)ssb_r4(cnt) * 4;
        const uint32_t b_jn = jn_m ? (uint32_t)ssb_r4(cnt * jn_m) * 4 : 0u, b_st = stateful ? (uint32_t)cnt * 128 : 0u;
        ssb_mbar_expect_tx(&bar, b_enc + b_bias + b_jn + b_st);
        ssb_bulk_g2s(s_enc, c.W + enc_off + (size_t)n0 * dpad, b_enc, &bar);
        ssb_bulk_g2s(s_bias, c.W + bias_off + n0, b_bias, &bar);
        if (jn_m) ssb_bulk_g2s(s_jn, c.W + jn_w + (size_t)n0 * jn_m, b_jn, &bar);
        if (stateful) ssb_bulk_g2s(s_st, stg, b_st, &bar);
    }
    const float* vg = ssb_grp(c.vec, c.nv, g, lane);
    ssb_stage_rows(xs, vg, in_row0, dims, dpad, warp, 4, lane);
    for (int m = warp; m < jn_m; m += 4) us[m * 32 + lane] = vg[(size_t)(jn_row0 + m) * 32];
    __syncthreads();                 // xs / us complete, barrier initialised for every thread
    ssb_mbar_wait(&bar, 0);
    float x[DP > 0 ? DP : 1];
    if (DP > 0) {
#pragma unroll
        for (int k = 0; k < DP; ++k) x[k] = xs[k * 32 + lane];
    }
    const int per = chunk >> 2;
    const int i_lo = warp * per, i_hi = min(cnt, i_lo + per);
    float* ag = ssb_grp(c.act, c.n_act, g, lane) + (size_t)(act0 + n0) * 32;
    for (int i = i_lo; i < i_hi; ++i) {
        const float4* e4 = reinterpret_cast<const float4*>(s_enc + (size_t)i * dpad);
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        if (DP > 0) {
#pragma unroll
            for (int k4 = 0; k4 < DP / 4; ++k4) {
                const float4 e = e4[k4];
                a0 = fmaf(e.x, x[4 * k4 + 0], a0);
                a1 = fmaf(e.y, x[4 * k4 + 1], a1);
                a2 = fmaf(e.z, x[4 * k4 + 2], a2);
                a3 = fmaf(e.w, x[4 * k4 + 3], a3);
            }
        } else {
            for (int k4 = 0; k4 < (dpad >> 2); ++k4) {
                const float4 e = e4[k4];
                const float* xk = xs + (k4 * 4) * 32 + lane;
                a0 = fmaf(e.x, xk[0], a0);
                a1 = fmaf(e.y, xk[32], a1);
                a2 = fmaf(e.z, xk[64], a2);
                a3 = fmaf(e.w, xk[96], a3);
            }
        }
        float J = s_bias[i] + ((a0 + a1) + (a2 + a3));
        for (int m = 0; m < jn_m; ++m) J = fmaf(s_jn[i * jn_m + m], us[m * 32 + lane], J);
        float out;
        if (stateful) {
            float sv = s_st[i * 32 + lane];
            out = nt.fast ? ssb_lif_packed<true>(nt, J, sv) : ssb_lif_packed<false>(nt, J, sv);
            s_st[i * 32 + lane] = sv;
        } else {
            out = ssb_rate(nt, J);
        }
        ag[(size_t)i * 32] = out;
        const unsigned any_on = __ballot_sync(0xffffffffu, out != 0.f);          // bit t: trial t of the group is active
        if (lane == 0) c.aflag[(size_t)g * c.n_act + act0 + n0 + i] = (int)any_on;
    }
    if (stateful && i_hi > i_lo) {
        ssb_fence_async();
        __syncwarp();
        if (lane == 0) {
            ssb_bulk_s2g(stg + (size_t)i_lo * 32, s_st + (size_t)i_lo * 32, (uint32_t)(i_hi - i_lo) * 128);
            ssb_bulk_commit();
            ssb_bulk_wait0();
        }
    }
}

// Tensor-core variant of k_wide_static (tcgen05 + TMEM): the input currents of a static wide ensemble are the
// GEMM J[trial][neuron] = X[trial][k] . E[neuron][k] with encoders shared by every trial.  CTA = (ensemble, block
// of 128 trials, chunk of 64-neuron tiles).  A = X (128 x KP, K-major, 3xTF32 hi | lo) is built once from the
// materialised input rows; B = encoder tiles (64 x KP, hi | lo, pre-tiled by the host) arrive by TMA in a
// two-stage ring; D (128 lanes x 64 columns) is double-buffered in TMEM so the MMAs of tile i+1 overlap the
// neuron epilogue of tile i: tcgen05.ld (lane = trial), + bias (+ direct neuron currents), LIF update on the
// packed state rows (coalesced 128-byte loads / stores per neuron), activities to the act arena.
// Et: [n_tiles][hi|lo][k/4][8 row groups][8][4] floats.  dynamic smem: (2*128 + 4*64) * KP floats.
#define SSB_ETC_N 64
template <bool FAST>
__global__ void __launch_bounds__(512, 1)
k_wide_static_tc(SsbCtx c, const int* __restrict__ desc, SsbItemList items, const float* __restrict__ Et_all,
                 const int* __restrict__ et_off, int KP, int tiles_per_chunk) {
    extern __shared__ __align__(1024) float sm[];
    __shared__ unsigned long long full[2], done[2];
    __shared__ uint32_t tmem_slot;
    __shared__ float s_bias[2][SSB_ETC_N], s_jnw[2][4 * SSB_ETC_N];   // per-tile bias / direct-current weights, double-buffered
    const int item = items.idx[blockIdx.z];
    const int* d = desc + item * 16;
    const int n = d[0], dims = d[1], state0 = d[3], act0 = d[4], bias_off = d[6], in_row0 = d[7];
    const int jn_row0 = d[10], jn_m = d[11], jn_w = d[12];
    const float* __restrict__ Et = Et_all + et_off[item];
    const int n_tiles = (n + SSB_ETC_N - 1) / SSB_ETC_N;
    const int t_lo = blockIdx.x * tiles_per_chunk;
    if (t_lo >= n_tiles) return;
    const int my_tiles = min(n_tiles, t_lo + tiles_per_chunk) - t_lo;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int quad = warp & 3, part = warp >> 2;        // 16 warps: TMEM quadrant, 16-column slice of the tile
    const int group = blockIdx.y * 4 + quad;
    const bool live = group < c.G;
    const int g = live ? group : 0;
    const SsbNeuron nt = ssb_neuron(c, d[8]);
    const bool stateful = nt.type == 0;
    const int a_part = 128 * KP, b_part = SSB_ETC_N * KP;
    const uint32_t tile_bytes = 2u * b_part * 4u;
    float* sA = sm;                                         // [hi|lo][a_part]
    float* sB = sm + 2 * a_part;                            // [2 stages][hi|lo][b_part]
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ssb_smem(&tmem_slot)), "r"(128));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (threadIdx.x == 0) {
        ssb_mbar_init(&full[0], 1);
        ssb_mbar_init(&full[1], 1);
        ssb_mbar_init(&done[0], 1);
        ssb_mbar_init(&done[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int i = 0; i < 2 && i < my_tiles; ++i) {
            ssb_mbar_expect_tx(&full[i], tile_bytes);
            ssb_bulk_g2s(sB + (size_t)i * 2 * b_part, Et + (size_t)(t_lo + i) * 2 * b_part, tile_bytes, &full[i]);
        }
    }
    float* vg = ssb_grp(c.vec, c.nv, g, lane);
    {   // A operand: this thread's trial is row r; the four warps of a quadrant alternate 32-column blocks
        const int r = quad * 32 + lane;
        float* a_hi = sA + (r >> 3) * 32 + (r & 7) * 4;
        float* a_lo = a_hi + a_part;
        const float* src = vg + (size_t)in_row0 * 32;
        for (int k0 = part * 32; k0 < KP; k0 += 128) {
            float x[32];
#pragma unroll
            for (int e = 0; e < 32; ++e) x[e] = (live && k0 + e < dims) ? src[(size_t)(k0 + e) * 32] : 0.f;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int k = k0 + 4 * q;
                if (k < KP) {
                    float4 hi, lo;
                    hi.x = ssb_tf32_round(x[4 * q + 0]);
                    hi.y = ssb_tf32_round(x[4 * q + 1]);
                    hi.z = ssb_tf32_round(x[4 * q + 2]);
                    hi.w = ssb_tf32_round(x[4 * q + 3]);
                    lo.x = ssb_tf32_round(x[4 * q + 0] - hi.x);
                    lo.y = ssb_tf32_round(x[4 * q + 1] - hi.y);
                    lo.z = ssb_tf32_round(x[4 * q + 2] - hi.z);
                    lo.w = ssb_tf32_round(x[4 * q + 3] - hi.w);
                    *reinterpret_cast<float4*>(a_hi + (size_t)(k >> 2) * 16 * 32) = hi;
                    *reinterpret_cast<float4*>(a_lo + (size_t)(k >> 2) * 16 * 32) = lo;
                }
            }
        }
    }
    const int jm = min(jn_m, 4);
    auto stage_consts = [&](int i) {                        // tile i's bias / jn weights -> smem stage i & 1
        const int s = i & 1, base = (t_lo + i) * SSB_ETC_N;
        if (threadIdx.x < SSB_ETC_N) {
            const int nn = base + threadIdx.x;
            s_bias[s][threadIdx.x] = nn < n ? __ldg(c.W + bias_off + nn) : 0.f;
        }
        if (threadIdx.x < jm * SSB_ETC_N) {
            const int e = base * jn_m + threadIdx.x;         // jm == jn_m whenever this path is taken (host guarantees jn_m <= 4)
            s_jnw[s][threadIdx.x] = e < n * jn_m ? __ldg(c.W + jn_w + e) : 0.f;
        }
    };
    stage_consts(0);
    ssb_fence_async();
    ssb_tc_fence_before();
    __syncthreads();
    ssb_tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(SSB_ETC_N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    auto issue_mma = [&](int i) {
        const int s = i & 1;
        ssb_mbar_wait(&full[s], (uint32_t)(i >> 1) & 1u);
        ssb_tc_fence_after();
        const float* b_hi = sB + (size_t)s * 2 * b_part;
        const uint32_t dst = tmem + (uint32_t)s * SSB_ETC_N;
#pragma unroll 1
        for (int j = 0; j < KP / 8; ++j) {
            const size_t oa = (size_t)j * 2 * 16 * 32, ob = (size_t)j * 2 * 8 * 32;
            const uint64_t ah = ssb_umma_desc_lbo(sA + oa, 2048), al = ssb_umma_desc_lbo(sA + a_part + oa, 2048);
            const uint64_t bh = ssb_umma_desc_lbo(b_hi + ob, 1024), bl = ssb_umma_desc_lbo(b_hi + b_part + ob, 1024);
            ssb_umma_tf32(dst, al, bh, idesc, j > 0);
            ssb_umma_tf32(dst, ah, bl, idesc, 1);
            ssb_umma_tf32(dst, ah, bh, idesc, 1);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(ssb_smem(&done[s]))
                     : "memory");
    };
    float u_jn[4];                                          // direct neuron currents (inhibition): a few inputs per trial
#pragma unroll
    for (int m = 0; m < 4; ++m) u_jn[m] = (m < jn_m) ? vg[(size_t)(jn_row0 + m) * 32] : 0.f;
    float* sg = ssb_grp(c.st, c.nn, g, lane) + (size_t)state0 * 32;
    float* ag = ssb_grp(c.act, c.n_act, g, lane) + (size_t)act0 * 32;
    if (threadIdx.x == 0) issue_mma(0);
    __syncwarp();
    for (int i = 0; i < my_tiles; ++i) {
        const int s = i & 1;
        if (threadIdx.x == 0 && i + 1 < my_tiles) issue_mma(i + 1);
        __syncwarp();
        if (i + 1 < my_tiles) stage_consts(i + 1);
        const int nn0 = (t_lo + i) * SSB_ETC_N + part * 16;     // first neuron of this thread's 16 columns
        const int nvalid = live ? min(16, max(0, n - nn0)) : 0;
        float* sgt = sg + (size_t)nn0 * 32;
        float* agt = ag + (size_t)nn0 * 32;
        float sv[16];
        if (stateful) {                                         // state rows in flight while the MMAs finish
#pragma unroll
            for (int j = 0; j < 16; ++j) sv[j] = j < nvalid ? __ldcs(sgt + j * 32) : 0.f;
        }
        ssb_mbar_wait(&done[s], (uint32_t)(i >> 1) & 1u);
        ssb_tc_fence_after();
        if (threadIdx.x == 0 && i + 2 < my_tiles) {
            ssb_mbar_expect_tx(&full[s], tile_bytes);
            ssb_bulk_g2s(sB + (size_t)s * 2 * b_part, Et + (size_t)(t_lo + i + 2) * 2 * b_part, tile_bytes, &full[s]);
        }
        __syncwarp();
        float v[16];
        ssb_tmem_ld16(tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)s * SSB_ETC_N + (uint32_t)part * 16, v);
        int* fl = c.aflag + (size_t)g * c.n_act + act0 + nn0;
        auto neuron = [&](int j) {
            float J = v[j] + s_bias[s][part * 16 + j];
            for (int m = 0; m < jm; ++m) J = fmaf(s_jnw[s][(part * 16 + j) * jm + m], u_jn[m], J);
            float out;
            if (stateful) {
                float st = sv[j];
                out = ssb_lif_packed<FAST>(nt, J, st);
                __stcs(sgt + j * 32, st);
            } else {
                out = ssb_rate(nt, J);
            }
            agt[j * 32] = out;
            const unsigned any_on = __ballot_sync(0xffffffffu, out != 0.f);
            if (lane == 0) fl[j] = (int)any_on;
        };
        if (nvalid == 16) {
#pragma unroll
            for (int j = 0; j < 16; ++j) neuron(j);
        } else {
#pragma unroll 1
            for (int j = 0; j < nvalid; ++j) {
                float vj = 0.f, svj = 0.f;                      // ragged last tile: select without dynamic register indexing
#pragma unroll
                for (int q = 0; q < 16; ++q)
                    if (q == j) { vj = v[q]; svj = sv[q]; }
                float J = vj + s_bias[s][part * 16 + j];
                for (int m = 0; m < jm; ++m) J = fmaf(s_jnw[s][(part * 16 + j) * jm + m], u_jn[m], J);
                float out;
                if (stateful) {
                    out = ssb_lif_packed<FAST>(nt, J, svj);
                    __stcs(sgt + j * 32, svj);
                } else {
                    out = ssb_rate(nt, J);
                }
                agt[j * 32] = out;
                const unsigned any_on = __ballot_sync(0xffffffffu, out != 0.f);
                if (lane == 0) fl[j] = (int)any_on;
            }
        }
        ssb_tc_fence_before();
        __syncthreads();
        ssb_tc_fence_after();
    }
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128));
}

// Voja-learned ensemble (associative-memory keys): the scaled encoders are per trial, the `dims` rows
// of one neuron are `dims` consecutive 128-byte lines.  Each warp streams its neurons' encoder tiles
// through a ring of shared-memory tiles (three in flight) with TMA bulk copies; lanes that spiked update their
// column in place and the tile is written back only if some lane spiked (post_synapse=None => the
// delta is row-sparse).  SimVoja: delta = alpha*L*(scale*outer(post, x) - post[:,None]*E), visible
// to the next step.
#define SSB_VOJA_NB 3         // most encoder tiles in flight per warp (ring barriers); the launch uses SSB_VOJA_NB_DEFAULT
// PES = true: the ensemble's activities feed a PES-learned connection whose decode is fused into this kernel (deferred-PES
// form, see ssb_pes.cuh): after its neuron range a warp walks the neurons where some trial spiked and accumulates
//   out[trial][j] += D_base[j][i][trial] * a[i][trial]      and      dots[trial][q] += f_q[i][trial] * a[i][trial]
// in registers — lane = trial, so no cross-lane reduction, and a lane loads only where ITS trial spiked (32-byte sectors).
// The four warps are added through shared memory, the CTA's partial goes to the PES partial arena, and the last CTA of the
// (ensemble, trial group) adds the chunks in order and applies the K history terms.  This replaces k_pes_defer (a second
// pass over flags, activities and decoder rows) on the longest dependency chain of the step.
#define SSB_PESF_NJ 56        // decoder output rows accumulated per pass
template <int DP, bool PES>
__global__ void __launch_bounds__(128) k_wide_voja(SsbCtx c, const int* __restrict__ desc, SsbItemList items, int chunk,
                                                    int i_rel, int nb, SsbPesFuse pf) {
    extern __shared__ __align__(128) float sm[];
    __shared__ unsigned long long wbar[4][SSB_VOJA_NB];
    const int* d = desc + items.idx[blockIdx.z] * 16;
    const int n = d[0], dims = d[1], dpad = d[2], state0 = d[3], act0 = d[4], enc_off = d[5], bias_off = d[6];
    const int in_row0 = d[7], jn_row0 = d[10], jn_m = d[11], jn_w = d[12], voja_row = d[13], scale_off = d[14];
    const int n0 = blockIdx.x * chunk;
    if (n0 >= n) return;
    const int cnt = min(chunk, n - n0);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int g = blockIdx.y;
    const SsbNeuron nt = ssb_neuron(c, d[8]);
    const bool stateful = nt.type == 0;
    float* xs = sm;                                        // [dpad][32]
    float* us = xs + (size_t)dpad * 32;                    // [jn_m][32]
    float* ebuf = us + (size_t)jn_m * 32 + (size_t)warp * nb * dims * 32;   // [nb][dims][32] per warp
    const int per = (chunk + nwarps - 1) / nwarps;
    const int i_lo = warp * per, i_hi = min(cnt, i_lo + per);
    float* eg = c.lenc + ((size_t)g * c.n_lenc + enc_off + (size_t)(n0 + i_lo) * dims) * 32;   // tile of neuron i_lo
    const uint32_t tile_bytes = (uint32_t)dims * 128;
    if (lane == 0) {
        for (int t = 0; t < SSB_VOJA_NB; ++t) ssb_mbar_init(&wbar[warp][t], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int t = 0; t < nb && i_lo + t < i_hi; ++t) {
            ssb_mbar_expect_tx(&wbar[warp][t], tile_bytes);
            ssb_bulk_g2s(ebuf + (size_t)t * dims * 32, eg + (size_t)t * dims * 32, tile_bytes, &wbar[warp][t]);
        }
    }
    const float* vg = ssb_grp(c.vec, c.nv, g, lane);
    float x[DP > 0 ? DP : 1];
    if (DP > 0) {       // the input rows go straight to registers: DP independent L2 loads per lane, no shared-memory hop
#pragma unroll
        for (int k = 0; k < DP; ++k) x[k] = (k < dims) ? vg[(size_t)(in_row0 + k) * 32] : 0.f;
    } else {
        ssb_stage_rows(xs, vg, in_row0, dims, dpad, warp, nwarps, lane);
    }
    for (int m = warp; m < jn_m; m += nwarps) us[m * 32 + lane] = vg[(size_t)(jn_row0 + m) * 32];
    const float aL = __int_as_float(d[15]) * vg[(size_t)voja_row * 32];
    __syncthreads();
    float* sp = ssb_grp(c.st, c.nn, g, lane) + (size_t)(state0 + n0) * 32;
    float* ag = ssb_grp(c.act, c.n_act, g, lane) + (size_t)(act0 + n0) * 32;
    uint32_t phases = 0;
    // the state row (HBM) and the bias of neuron i + 1 are requested while neuron i is computed: eight warps per SM do
    // not hide one memory round trip per neuron
    // flags bit 2: every trial has its own network seed - bias and Voja scale are rows of the per-trial weight arena
    const bool pt = (d[9] & 4) != 0;
    const int pstride = pt ? 32 : 1;
    const float* __restrict__ bias_p = pt ? c.wpt + ((size_t)g * c.n_wpt + bias_off + n0) * 32 + lane : c.W + bias_off + n0;
    const float* __restrict__ scale_p = pt ? c.wpt + ((size_t)g * c.n_wpt + scale_off + n0) * 32 + lane : c.W + scale_off + n0;
    const float* __restrict__ jn_p = pt ? c.wpt + ((size_t)g * c.n_wpt + jn_w + (size_t)n0 * jn_m) * 32 + lane
                                        : c.W + jn_w + (size_t)n0 * jn_m;
    float sv_next = 0.f, bias_next = 0.f;
    if (i_lo < i_hi) {
        if (stateful) sv_next = __ldcs(sp + (size_t)i_lo * 32);
        bias_next = __ldg(bias_p + (size_t)i_lo * pstride);
    }
    for (int i = i_lo; i < i_hi; ++i) {
        const int t = i - i_lo, b = t % nb;
        float* E = ebuf + (size_t)b * dims * 32 + lane;
        float sv = sv_next;
        float J = bias_next;
        if (i + 1 < i_hi) {
            if (stateful) sv_next = __ldcs(sp + (size_t)(i + 1) * 32);
            bias_next = __ldg(bias_p + (size_t)(i + 1) * pstride);
        }
        for (int m = 0; m < jn_m; ++m) J = fmaf(__ldg(jn_p + (size_t)(i * jn_m + m) * pstride), us[m * 32 + lane], J);
        ssb_mbar_wait(&wbar[warp][b], (phases >> b) & 1u);   // phases: one parity bit per buffer
        phases ^= 1u << b;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        if (DP > 0) {
#pragma unroll
            for (int k = 0; k < DP; k += 4) {
                a0 = fmaf(k + 0 < dims ? E[(k + 0) * 32] : 0.f, x[k + 0], a0);
                a1 = fmaf(k + 1 < dims ? E[(k + 1) * 32] : 0.f, x[k + 1], a1);
                a2 = fmaf(k + 2 < dims ? E[(k + 2) * 32] : 0.f, x[k + 2], a2);
                a3 = fmaf(k + 3 < dims ? E[(k + 3) * 32] : 0.f, x[k + 3], a3);
            }
        } else {
            int k = 0;
            for (; k + 4 <= dims; k += 4) {
                a0 = fmaf(E[k * 32], xs[k * 32 + lane], a0);
                a1 = fmaf(E[(k + 1) * 32], xs[(k + 1) * 32 + lane], a1);
                a2 = fmaf(E[(k + 2) * 32], xs[(k + 2) * 32 + lane], a2);
                a3 = fmaf(E[(k + 3) * 32], xs[(k + 3) * 32 + lane], a3);
            }
            for (; k < dims; ++k) a0 = fmaf(E[k * 32], xs[k * 32 + lane], a0);
        }
        J += (a0 + a1) + (a2 + a3);
        float out;
        if (stateful) {
            out = nt.fast ? ssb_lif_packed<true>(nt, J, sv) : ssb_lif_packed<false>(nt, J, sv);
            __stcs(sp + (size_t)i * 32, sv);
        } else {
            out = ssb_rate(nt, J);
        }
        ag[(size_t)i * 32] = out;
        const bool fired = out != 0.f;
        {
            const unsigned any_on = __ballot_sync(0xffffffffu, fired);
            if (lane == 0) c.aflag[(size_t)g * c.n_act + act0 + n0 + i] = (int)any_on;
        }
        const bool learn = fired && aL != 0.f;           // alpha = 0: a static ensemble with per-trial encoders
        if (learn) {
            const float sc = __ldg(scale_p + (size_t)i * pstride);
            if (DP > 0) {
#pragma unroll
                for (int k = 0; k < DP; ++k) {
                    if (k < dims) {
                        const float e = E[k * 32];
                        E[k * 32] = e + aL * (sc * (out * x[k]) - out * e);
                    }
                }
            } else {
                for (int k = 0; k < dims; ++k) {
                    const float e = E[k * 32];
                    E[k * 32] = e + aL * (sc * (out * xs[k * 32 + lane]) - out * e);
                }
            }
        }
        const bool dirty = __any_sync(0xffffffffu, learn);
        if (dirty) ssb_fence_async();
        __syncwarp();
        if (lane == 0) {
            // one bulk group per tile (empty when the tile is clean) keeps the group count in step with the tiles:
            // before buffer b_prev = (t - 1) % NB is refilled, only the group of tile t may still be reading
            if (dirty) ssb_bulk_s2g(eg + (size_t)t * dims * 32, ebuf + (size_t)b * dims * 32, tile_bytes);
            ssb_bulk_commit();
            if (nb == 1) {                                   // single buffer: refill after this tile's own store has read it
                if (i + 1 < i_hi) {
                    ssb_bulk_wait_read0();
                    ssb_mbar_expect_tx(&wbar[warp][0], tile_bytes);
                    ssb_bulk_g2s(ebuf, eg + (size_t)(t + 1) * dims * 32, tile_bytes, &wbar[warp][0]);
                }
            } else if (t >= 1 && i + nb - 1 < i_hi) {
                const int bp = (t - 1) % nb;
                asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                ssb_mbar_expect_tx(&wbar[warp][bp], tile_bytes);
                ssb_bulk_g2s(ebuf + (size_t)bp * dims * 32, eg + (size_t)(t - 1 + nb) * dims * 32, tile_bytes, &wbar[warp][bp]);
            }
        }
    }
    if (lane == 0) ssb_bulk_wait0();
    if (!PES) return;
    const int pit = pf.item[blockIdx.z];
    if (pit < 0) return;
    // ------------------------------------------------------------------ fused deferred-PES decode
    __syncwarp();                                       // the warp's ring buffers are drained: reuse them for the reduction
    __shared__ int pes_last;
    const int* pd = pf.desc + pit * 13;
    const int* hd = pf.hdesc + pit * 4;
    const int size_out = pd[1], d_off = pd[2], a_off = pd[3], err_vec = pd[5], out_vec = pd[6];
    const int K = pf.h.K;
    const SsbStep stp = ssb_step(c, i_rel);
    const int slot = (int)(stp.step % K);      // this step's term is not in the history yet: it is read at its source
    const int JP = (size_out + 3) & ~3;                 // decoder block of neuron i: [trial lane][JP] (see ssb_pes.cuh)
    const float* __restrict__ dl = c.ldec + ((size_t)g * c.n_ldec + d_off) * 32 + (size_t)lane * JP;
    const float* __restrict__ hf = ssb_grp(pf.h.hist_f, pf.h.rows_f, g, lane) + (size_t)hd[1] * 32;
    const float* __restrict__ fcur = ssb_grp(c.afilt, 2 * c.n_afilt, g, lane) + ((size_t)(1 - stp.odd) * c.n_afilt + a_off) * 32;
    const int prow = size_out + K;
    const int n_chunks = (n + chunk - 1) / chunk;
    float* pg = ssb_grp(pf.h.part, pf.h.rows_p, g, lane) + (size_t)hd[2] * 32;
    float* red = ebuf;                                  // [SSB_PESF_NJ + 8][32] of this warp
    for (int j0 = 0; j0 < size_out; j0 += SSB_PESF_NJ) {
        const bool with_dots = j0 == 0;
        float acc[SSB_PESF_NJ], dots[8];
#pragma unroll
        for (int j = 0; j < SSB_PESF_NJ; ++j) acc[j] = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) dots[q] = 0.f;
        for (int i = i_lo; i < i_hi; ++i) {
            const float a = ag[(size_t)i * 32];         // written by this very thread above
            const bool on = a != 0.f;
            if (!__any_sync(0xffffffffu, on)) continue;
            const size_t ni = (size_t)(n0 + i);
            float w[SSB_PESF_NJ], f[8];
            const float4* src4 = reinterpret_cast<const float4*>(dl + ni * JP * 32 + j0);
#pragma unroll
            for (int q4 = 0; q4 < SSB_PESF_NJ / 4; ++q4) {
                const float4 t4 = (on && j0 + 4 * q4 < JP) ? __ldcs(src4 + q4) : make_float4(0.f, 0.f, 0.f, 0.f);
                w[4 * q4 + 0] = t4.x;
                w[4 * q4 + 1] = t4.y;
                w[4 * q4 + 2] = t4.z;
                w[4 * q4 + 3] = t4.w;
            }
            if (with_dots) {
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    f[q] = (on && q < K) ? ((q == slot) ? fcur[ni * 32] : hf[((size_t)q * n + ni) * 32]) : 0.f;
#pragma unroll
                for (int q = 0; q < 8; ++q) dots[q] = fmaf(f[q], a, dots[q]);
            }
#pragma unroll
            for (int j = 0; j < SSB_PESF_NJ; ++j) acc[j] = fmaf(w[j], a, acc[j]);
        }
#pragma unroll
        for (int j = 0; j < SSB_PESF_NJ; ++j) red[j * 32 + lane] = acc[j];
        if (with_dots) {
#pragma unroll
            for (int q = 0; q < 8; ++q) red[(SSB_PESF_NJ + q) * 32 + lane] = dots[q];
        }
        __syncthreads();
        // CTA partial in warp order, parked in the partial arena [chunk][size_out + K]
        const size_t wstride = (size_t)nb * dims * 32;  // distance between the warps' buffers
        const float* red0 = us + (size_t)jn_m * 32;
        const int nrow = SSB_PESF_NJ + (with_dots ? 8 : 0);
        for (int r = warp; r < nrow; r += nwarps) {
            const bool is_dot = r >= SSB_PESF_NJ;
            const int row = is_dot ? size_out + (r - SSB_PESF_NJ) : j0 + r;
            if (is_dot ? (r - SSB_PESF_NJ < K) : (row < size_out)) {
                float t = 0.f;
                for (int wq = 0; wq < nwarps; ++wq) t += red0[wq * wstride + r * 32 + lane];
                pg[(size_t)(blockIdx.x * prow + row) * 32] = t;
            }
        }
        __syncthreads();
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        int* cnt_p = pf.h.counters + hd[3] * c.G + g;
        const int old = atomicAdd(cnt_p, 1);
        const int last = old == n_chunks - 1;
        if (last) *cnt_p = 0;
        pes_last = last;
    }
    __syncthreads();
    if (!pes_last) return;
    __threadfence();
    // the last CTA of this (ensemble, group): history dot products, then every output row (chunks added in order)
    float dsum[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) dsum[q] = 0.f;
    for (int ck = 0; ck < n_chunks; ++ck) {
        float v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = q < K ? __ldcg(pg + (size_t)(ck * prow + size_out + q) * 32) : 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) dsum[q] += v[q];
    }
    const float* __restrict__ he = ssb_grp(pf.h.hist_e, pf.h.rows_e, g, lane) + (size_t)hd[0] * 32;
    float* vgw = ssb_grp(c.vec, c.nv, g, lane);
    const float alpha = stp.step > 0 ? __int_as_float(pd[7]) : 0.f;
    for (int jb = warp * 8; jb < size_out; jb += nwarps * 8) {
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
                float r = t[u];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    if (q < K) {
                        const float e = (q == slot) ? alpha * vgw[(size_t)(err_vec + jb + u) * 32]
                                                    : he[(size_t)(q * size_out + jb + u) * 32];
                        r = fmaf(e, dsum[q], r);
                    }
                }
                vgw[(size_t)(out_vec + jb + u) * 32] = r;
            }
        }
    }
}

// --------------------------------------------------------------------------------------
// Voja ensembles whose encoder rows are too long to keep whole tiles in shared memory (d = 649: one neuron's encoders of
// one trial group are 83 KB).  The per-trial encoders of a warp's neuron range are ONE contiguous run of 128-byte rows
// (row = neuron * dims + k), so each warp streams that run through a ring of SSB_VS_NB sub-tiles of SSB_VS_SUB rows,
// irrespective of neuron boundaries; the dot product of a neuron is closed when its last row has passed.  Shared memory is
// only read: for the (few) lanes that spiked, the Voja update re-reads the neuron's rows from L2 and writes back just
// those lanes' words, so a spike costs 32-byte sector writes instead of a whole-tile write-back.
// CTA = (neuron chunk, trial group, ensemble), 8 warps; dynamic smem: xs [dpad][32] | us [jn_m][32] | 8 rings.
#define SSB_VS_NB 4
#define SSB_VS_SUB 32
__global__ void __launch_bounds__(256, 1) k_wide_voja_stream(SsbCtx c, const int* __restrict__ desc, SsbItemList items,
                                                             int chunk, int i_rel) {
    extern __shared__ __align__(128) float sm[];
    __shared__ unsigned long long wbar[8][SSB_VS_NB];
    const int* d = desc + items.idx[blockIdx.z] * 16;
    const int n = d[0], dims = d[1], dpad = d[2], state0 = d[3], act0 = d[4], enc_off = d[5], bias_off = d[6];
    const int in_row0 = d[7], jn_row0 = d[10], jn_m = d[11], jn_w = d[12], voja_row = d[13], scale_off = d[14];
    const int n0 = blockIdx.x * chunk;
    if (n0 >= n) return;
    const int cnt = min(chunk, n - n0);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int NW = 8;
    const int g = blockIdx.y;
    const SsbNeuron nt = ssb_neuron(c, d[8]);
    const bool stateful = nt.type == 0;
    float* xs = sm;                                        // [dpad][32]
    float* us = xs + (size_t)dpad * 32;                    // [jn_m][32]
    float* ring = us + (size_t)jn_m * 32 + (size_t)warp * SSB_VS_NB * SSB_VS_SUB * 32;
    const int per = (chunk + NW - 1) / NW;
    const int i_lo = warp * per, i_hi = min(cnt, i_lo + per);
    const long long n_rows = (long long)max(0, i_hi - i_lo) * dims;          // rows of this warp's run
    const int n_tiles = (int)((n_rows + SSB_VS_SUB - 1) / SSB_VS_SUB);
    float* eg = c.lenc + ((size_t)g * c.n_lenc + enc_off + (size_t)(n0 + i_lo) * dims) * 32;   // first row of the run
    auto issue = [&](int t) {                              // lane 0: sub-tile t -> stage t % NB
        const int b = t % SSB_VS_NB;
        const uint32_t bytes = (uint32_t)min((long long)SSB_VS_SUB, n_rows - (long long)t * SSB_VS_SUB) * 128u;
        ssb_mbar_expect_tx(&wbar[warp][b], bytes);
        ssb_bulk_g2s(ring + (size_t)b * SSB_VS_SUB * 32, eg + (size_t)t * SSB_VS_SUB * 32, bytes, &wbar[warp][b]);
    };
    if (lane == 0) {
        for (int t = 0; t < SSB_VS_NB; ++t) ssb_mbar_init(&wbar[warp][t], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int t = 0; t < SSB_VS_NB && t < n_tiles; ++t) issue(t);
    }
    const float* vg = ssb_grp(c.vec, c.nv, g, lane);
    ssb_stage_rows(xs, vg, in_row0, dims, dpad, warp, NW, lane);
    for (int m = warp; m < jn_m; m += NW) us[m * 32 + lane] = vg[(size_t)(jn_row0 + m) * 32];
    const float aL = __int_as_float(d[15]) * vg[(size_t)voja_row * 32];
    __syncthreads();
    float* sp = ssb_grp(c.st, c.nn, g, lane) + (size_t)(state0 + n0) * 32;
    float* ag = ssb_grp(c.act, c.n_act, g, lane) + (size_t)(act0 + n0) * 32;
    uint32_t phases = 0;
    int t_cur = 0, pos = 0, tile_rows = 0;                 // current sub-tile, next row inside it, rows it holds
    const float* E = ring + lane;
    if (n_tiles > 0) {
        ssb_mbar_wait(&wbar[warp][0], 0);
        phases ^= 1u;
        tile_rows = (int)min((long long)SSB_VS_SUB, n_rows);
    }
    // flags bit 2: every trial has its own network seed - bias and Voja scale are rows of the per-trial weight arena
    const bool pt = (d[9] & 4) != 0;
    const int pstride = pt ? 32 : 1;
    const float* __restrict__ bias_p = pt ? c.wpt + ((size_t)g * c.n_wpt + bias_off + n0) * 32 + lane : c.W + bias_off + n0;
    const float* __restrict__ scale_p = pt ? c.wpt + ((size_t)g * c.n_wpt + scale_off + n0) * 32 + lane : c.W + scale_off + n0;
    const float* __restrict__ jn_p = pt ? c.wpt + ((size_t)g * c.n_wpt + jn_w + (size_t)n0 * jn_m) * 32 + lane
                                        : c.W + jn_w + (size_t)n0 * jn_m;
    float sv_next = 0.f, bias_next = 0.f;
    if (i_lo < i_hi) {
        if (stateful) sv_next = __ldcs(sp + (size_t)i_lo * 32);
        bias_next = __ldg(bias_p + (size_t)i_lo * pstride);
    }
    for (int i = i_lo; i < i_hi; ++i) {
        float sv = sv_next;
        float J = bias_next;
        if (i + 1 < i_hi) {
            if (stateful) sv_next = __ldcs(sp + (size_t)(i + 1) * 32);
            bias_next = __ldg(bias_p + (size_t)(i + 1) * pstride);
        }
        for (int m = 0; m < jn_m; ++m) J = fmaf(__ldg(jn_p + (size_t)(i * jn_m + m) * pstride), us[m * 32 + lane], J);
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        int k = 0;
        while (k < dims) {
            if (pos == tile_rows) {                        // sub-tile consumed: refill its stage, move to the next one
                __syncwarp();
                if (lane == 0 && t_cur + SSB_VS_NB < n_tiles) issue(t_cur + SSB_VS_NB);
                ++t_cur;
                const int b = t_cur % SSB_VS_NB;
                ssb_mbar_wait(&wbar[warp][b], (phases >> b) & 1u);
                phases ^= 1u << b;
                E = ring + (size_t)b * SSB_VS_SUB * 32 + lane;
                pos = 0;
                tile_rows = (int)min((long long)SSB_VS_SUB, n_rows - (long long)t_cur * SSB_VS_SUB);
            }
            const int seg = min(dims - k, tile_rows - pos);
            const float* e = E + (size_t)pos * 32;
            const float* x = xs + (size_t)k * 32 + lane;
            int q = 0;
            for (; q + 4 <= seg; q += 4) {
                a0 = fmaf(e[(q + 0) * 32], x[(q + 0) * 32], a0);
                a1 = fmaf(e[(q + 1) * 32], x[(q + 1) * 32], a1);
                a2 = fmaf(e[(q + 2) * 32], x[(q + 2) * 32], a2);
                a3 = fmaf(e[(q + 3) * 32], x[(q + 3) * 32], a3);
            }
            for (; q < seg; ++q) a0 = fmaf(e[q * 32], x[q * 32], a0);
            k += seg;
            pos += seg;
        }
        J += (a0 + a1) + (a2 + a3);
        float out;
        if (stateful) {
            out = nt.fast ? ssb_lif_packed<true>(nt, J, sv) : ssb_lif_packed<false>(nt, J, sv);
            __stcs(sp + (size_t)i * 32, sv);
        } else {
            out = ssb_rate(nt, J);
        }
        ag[(size_t)i * 32] = out;
        const bool fired = out != 0.f;
        const unsigned any_on = __ballot_sync(0xffffffffu, fired);
        if (lane == 0) c.aflag[(size_t)g * c.n_act + act0 + n0 + i] = (int)any_on;
        const bool learn = fired && aL != 0.f;           // alpha = 0: a static ensemble with per-trial encoders
        if (__any_sync(0xffffffffu, learn)) {
            // Voja: only the lanes that spiked touch their words (the rows were just streamed: L2 re-read, 32-byte sector
            // writes).  Measured on B200 at d = 649: this lane = trial form 652 us per launch; a lanes = rows form (one
            // spiking trial at a time, every row in flight) 1 275 us - uncoalesced sector accesses cost ~4 LSU cycles each,
            // so the fewer sector operations win, not the fewer round trips.
            const float sc = __ldg(scale_p + (size_t)i * pstride);
            float* Eg = eg + (size_t)(i - i_lo) * dims * 32 + lane;
            for (int k0 = 0; k0 < dims; k0 += 16) {
                float ev[16];
#pragma unroll
                for (int u = 0; u < 16; ++u) ev[u] = (learn && k0 + u < dims) ? __ldcg(Eg + (size_t)(k0 + u) * 32) : 0.f;
#pragma unroll
                for (int u = 0; u < 16; ++u)
                    if (learn && k0 + u < dims)
                        Eg[(size_t)(k0 + u) * 32] = ev[u] + aL * (sc * (out * xs[(k0 + u) * 32 + lane]) - out * ev[u]);
            }
        }
    }
}
