// sm_100a kernels of the SSP-SLAM step engine: static decoders of wide ensembles: split-K epilogue, k_decode (FFMA), k_decode_tc (tcgen05).
// Included by ssb_kernels.cuh (after ssb_common.cuh); see that file for the layout rules.
#pragma once
#include "ssb_common.cuh"

// --------------------------------------------------------------------------------------
// Split-K epilogue shared by the decode and PES kernels.  The neuron range of one (decoder,
// 8-row tile, trial group) is split over n_chunks CTAs; each CTA reduces its 4 warps in shared
// memory and, if it is not alone, parks its partial sums in the `part` arena.  The CTA that
// arrives last (atomic counter, self-resetting) adds the partials in chunk order — a fixed order,
// so the result does not depend on scheduling — and writes the single output slot.
__device__ __forceinline__ void ssb_splitk_finish(const SsbCtx& c, float (*red)[8][32], int* flag, const float (&acc)[8],
                                                  int g, int j0, int size_out, int out_vec, int n_chunks, int chunk,
                                                  int part_off, int counter) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < 8; ++j) red[warp][j][lane] = acc[j];
    __syncthreads();
    float* vg = ssb_grp(c.vec, c.nv, g, lane);
    float* pg = ssb_grp(c.part, c.n_part, g, lane);
    for (int j = warp; j < 8; j += 4) {
        if (j0 + j < size_out) {
            const float t = (red[0][j][lane] + red[1][j][lane]) + (red[2][j][lane] + red[3][j][lane]);
            if (n_chunks == 1) vg[(size_t)(out_vec + j0 + j) * 32] = t;
            else pg[(size_t)(part_off + chunk * size_out + j0 + j) * 32] = t;
        }
    }
    if (n_chunks == 1) return;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const int old = atomicAdd(c.counters + counter, 1);
        const int last = old == n_chunks - 1;
        if (last) c.counters[counter] = 0;
        *flag = last;
    }
    __syncthreads();
    if (!*flag) return;
    __threadfence();
    for (int j = warp; j < 8; j += 4) {
        if (j0 + j < size_out) {
            float t = 0.f;
            for (int ck0 = 0; ck0 < n_chunks; ck0 += 8) {     // 8 independent loads in flight, added in chunk order
                float v[8];
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    v[q] = ck0 + q < n_chunks ? __ldcg(pg + (size_t)(part_off + (ck0 + q) * size_out + j0 + j) * 32) : 0.f;
#pragma unroll
                for (int q = 0; q < 8; ++q) t += v[q];
            }
            vg[(size_t)(out_vec + j0 + j) * 32] = t;
        }
    }
}

// Static decoders of wide ensembles: out[j] = sum_n Wd[n][j] * act[n].  CTA = (decoder, quad of trial
// groups, neuron chunk); each WARP owns one trial group and the whole 56-wide output tile for the chunk, so
// there is no cross-warp reduction: the four warps share the chunk's weight rows [cnt][jpad] (one TMA bulk
// copy, broadcast float4 reads) and each fetches its own group's activity rows [cnt][32] (one bulk copy per
// warp, own mbarrier).  A neuron whose activity is zero in all 32 trials of the group is skipped (spiking
// activity is sparse).  Chunks are combined by the split-K semaphore in chunk order (fixed summation order).
// desc: n size_out jpad act0 w_off out_vec n_chunks part_off counter0
// dynamic smem: per*jpad (weights) + 4*per*32 (activities) floats, per = ceil(n / n_chunks)
#define SSB_DEC_NJ 56
__global__ void __launch_bounds__(128) k_decode(SsbCtx c, const int* __restrict__ desc, int item0) {
    extern __shared__ __align__(128) float sm[];
    __shared__ unsigned long long bar_w, bar_a[4];
    const int* d = desc + (item0 + blockIdx.z) * 9;
    const int n = d[0], size_out = d[1], jpad = d[2], act0 = d[3], w_off = d[4], out_vec = d[5], n_chunks = d[6];
    const int part_off = d[7];
    const int chunk = blockIdx.x;
    if (chunk >= n_chunks) return;
    const int per = (n + n_chunks - 1) / n_chunks;
    const int i_lo = chunk * per, cnt = min(n, i_lo + per) - i_lo;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = blockIdx.y * 4 + warp;
    const bool live = g < c.G;
    float* s_w = sm;                                                   // [per][jpad]
    float* s_a = s_w + (size_t)per * jpad + (size_t)warp * per * 32;   // [per][32] of this warp's group
    if (threadIdx.x == 0) {
        ssb_mbar_init(&bar_w, 1);
        for (int q = 0; q < 4; ++q) ssb_mbar_init(&bar_a[q], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        ssb_mbar_expect_tx(&bar_w, (uint32_t)cnt * jpad * 4);
        ssb_bulk_g2s(s_w, c.W + w_off + (size_t)i_lo * jpad, (uint32_t)cnt * jpad * 4, &bar_w);
    }
    __syncthreads();
    if (!live) return;
    if (lane == 0) {
        ssb_mbar_expect_tx(&bar_a[warp], (uint32_t)cnt * 128);
        ssb_bulk_g2s(s_a, c.act + ((size_t)g * c.n_act + act0 + i_lo) * 32, (uint32_t)cnt * 128, &bar_a[warp]);
    }
    float* vg = ssb_grp(c.vec, c.nv, g, lane);
    float* pg = ssb_grp(c.part, c.n_part, g, lane);
    ssb_mbar_wait(&bar_a[warp], 0);
    ssb_mbar_wait(&bar_w, 0);
    for (int jb = 0; jb < jpad; jb += SSB_DEC_NJ) {
        const int nq = min(SSB_DEC_NJ, jpad - jb) >> 2;      // float4 columns of this pass (jpad is a multiple of 8)
        float acc[SSB_DEC_NJ];
#pragma unroll
        for (int j = 0; j < SSB_DEC_NJ; ++j) acc[j] = 0.f;
        for (int i = 0; i < cnt; ++i) {
            const float a = s_a[i * 32 + lane];
            if (__any_sync(0xffffffffu, a != 0.f)) {
                const float4* w4 = reinterpret_cast<const float4*>(s_w + (size_t)i * jpad + jb);
#pragma unroll
                for (int k = 0; k < SSB_DEC_NJ / 4; ++k) {
                    if (k < nq) {
                        const float4 w = w4[k];
                        acc[4 * k + 0] = fmaf(w.x, a, acc[4 * k + 0]);
                        acc[4 * k + 1] = fmaf(w.y, a, acc[4 * k + 1]);
                        acc[4 * k + 2] = fmaf(w.z, a, acc[4 * k + 2]);
                        acc[4 * k + 3] = fmaf(w.w, a, acc[4 * k + 3]);
                    }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < SSB_DEC_NJ; ++j) {
            if (j < 4 * nq && jb + j < size_out) {
                if (n_chunks == 1) vg[(size_t)(out_vec + jb + j) * 32] = acc[j];
                else pg[(size_t)(part_off + chunk * size_out + jb + j) * 32] = acc[j];
            }
        }
    }
    if (n_chunks == 1) return;
    // split-K: one arrival counter per (decoder, trial group); the warp that arrives last adds the partials
    __threadfence();
    __syncwarp();
    int last = 0;
    if (lane == 0) {
        int* cnt_p = c.counters + d[8] * c.G + g;
        const int old = atomicAdd(cnt_p, 1);
        last = old == n_chunks - 1;
        if (last) *cnt_p = 0;
    }
    last = __shfl_sync(0xffffffffu, last, 0);
    if (!last) return;
    __threadfence();
    // 8 outputs x 8 chunks = 64 independent loads in flight; the additions stay in chunk order
    for (int j = 0; j < size_out; j += 8) {
        float t[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) t[u] = 0.f;
        for (int ck0 = 0; ck0 < n_chunks; ck0 += 8) {
            float v[8][8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const bool ok = ck0 + q < n_chunks && j + u < size_out;
                    v[q][u] = ok ? __ldcg(pg + (size_t)(part_off + (ck0 + q) * size_out + j + u) * 32) : 0.f;
                }
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) {
#pragma unroll
                for (int u = 0; u < 8; ++u) t[u] += v[q][u];
            }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u)
            if (j + u < size_out) vg[(size_t)(out_vec + j + u) * 32] = t[u];
    }
}

// Tensor-core static decoders (tcgen05 + TMEM): out[trial][j] = sum_k act[k][trial] * Wd[k][j] is a dense GEMM
// whose weights are shared by every trial.  CTA = (decoder, block of 128 trials = 4 trial groups, K chunk);
//   A = activities (128 trials x 64 neurons per stage, K-major), gathered by the CTA's 256 threads from the
//       group-tiled act arena (coalesced 128-byte rows) and split on the fly into TF32 hi + lo,
//   B = Wd^T (64 output rows x 64 neurons per stage, K-major) pre-split into hi / lo and pre-tiled by the host in
//       UMMA core-matrix order, one TMA bulk copy per stage,
//   D = 128 lanes x 64 fp32 columns in TMEM, accumulated over the chunk's stages with the 3xTF32 scheme
//       (A_lo.B_hi + A_hi.B_lo + A_hi.B_hi).  Building stage s+1 overlaps the MMAs of stage s (two buffers).
// The epilogue reads D with tcgen05.ld (lane = trial) and writes the output rows (or split-K partial sums,
// combined in chunk order by the last CTA to arrive, as in the FFMA kernel).
// Wt: [n_stages][hi|lo][k/4][8 row groups][8][4] floats (64 rows x 64 columns per part).
// Instantiated for <N = 64 outputs, KS = 64 neurons per stage> and <N = 128, KS = 32> (wider decoders, e.g. d = 97).
template <int SSB_DTC_N, int SSB_DTC_KS>
__global__ void __launch_bounds__(256, 1)
k_decode_tc(SsbCtx c, const int* __restrict__ desc, int item0, const float* __restrict__ Wt_all, const int* __restrict__ wt_off,
            int n_nt) {
    extern __shared__ __align__(1024) float sm[];
    __shared__ unsigned long long full[2], done[2];
    __shared__ uint32_t tmem_slot;
    __shared__ int s_last[4];
    // blockIdx.z = (decoder, tile of SSB_DTC_N output columns): decoders wider than one tile (d = 649) are covered by n_nt
    // column tiles, each with its own operand tiles [n_stages][hi | lo] and its own split-K counter
    const int zi = blockIdx.z / n_nt, nt = blockIdx.z - zi * n_nt;
    const int* d = desc + (item0 + zi) * 9;
    const int n = d[0], size_out = d[1], act0 = d[3], out_vec = d[5], n_chunks = d[6], part_off = d[7];
    const int col0 = nt * SSB_DTC_N;
    const int chunk = blockIdx.x;
    if (chunk >= n_chunks || col0 >= size_out) return;
    const int n_stages = (n + SSB_DTC_KS - 1) / SSB_DTC_KS;
    const float* __restrict__ Wt = Wt_all + wt_off[item0 + zi] + (size_t)nt * n_stages * 2 * SSB_DTC_N * SSB_DTC_KS;
    const int spc = (n_stages + n_chunks - 1) / n_chunks;
    const int s_lo = chunk * spc, s_hi = min(n_stages, s_lo + spc);
    const int my = max(0, s_hi - s_lo);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int quad = warp & 3, half = warp >> 2;
    const int group = blockIdx.y * 4 + quad;
    const bool live = group < c.G;
    const int g = live ? group : 0;
    constexpr int A_PART = 128 * SSB_DTC_KS;            // floats of one A part (hi or lo)
    constexpr int B_PART = SSB_DTC_N * SSB_DTC_KS;
    float* sA = sm;                                     // [2 buffers][hi|lo][A_PART]
    float* sB = sm + 4 * A_PART;                        // [2 buffers][hi|lo][B_PART]
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ssb_smem(&tmem_slot)), "r"(SSB_DTC_N));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (threadIdx.x == 0) {
        ssb_mbar_init(&full[0], 1);
        ssb_mbar_init(&full[1], 1);
        ssb_mbar_init(&done[0], 1);
        ssb_mbar_init(&done[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int i = 0; i < 2 && i < my; ++i) {
            ssb_mbar_expect_tx(&full[i], 2u * B_PART * 4u);
            ssb_bulk_g2s(sB + (size_t)i * 2 * B_PART, Wt + (size_t)(s_lo + i) * 2 * B_PART, 2u * B_PART * 4u, &full[i]);
        }
    }
    ssb_tc_fence_before();
    __syncthreads();
    ssb_tc_fence_after();
    const uint32_t tmem = tmem_slot;
    // D fp32, A/B tf32, both K-major, N = 64, M = 128
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(SSB_DTC_N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const int r = quad * 32 + lane;                     // this thread's trial row; `half` picks its half of the stage's columns
    constexpr int HK = SSB_DTC_KS / 2;                  // activity rows per thread and stage
    const float* ag = ssb_grp(c.act, c.n_act, g, lane) + (size_t)act0 * 32;
    for (int i = 0; i < my; ++i) {
        const int b = i & 1;
        if (i >= 2) {                                   // buffer b was read by the MMAs of stage i - 2
            ssb_mbar_wait(&done[b], (uint32_t)((i - 2) >> 1) & 1u);
            ssb_tc_fence_after();
            if (threadIdx.x == 0) {
                ssb_mbar_expect_tx(&full[b], 2u * B_PART * 4u);
                ssb_bulk_g2s(sB + (size_t)b * 2 * B_PART, Wt + (size_t)(s_lo + i) * 2 * B_PART, 2u * B_PART * 4u, &full[b]);
            }
        }
        {   // A stage: HK activity rows per thread, all loads issued before they are consumed
            const int k0 = (s_lo + i) * SSB_DTC_KS + half * HK;
            float x[HK];
#pragma unroll
            for (int e = 0; e < HK; ++e) x[e] = (live && k0 + e < n) ? ag[(size_t)(k0 + e) * 32] : 0.f;
            float* a_hi = sA + (size_t)b * 2 * A_PART + (r >> 3) * 32 + (r & 7) * 4 + (size_t)(half * (HK / 4)) * 16 * 32;
            float* a_lo = a_hi + A_PART;
#pragma unroll
            for (int q = 0; q < HK / 4; ++q) {
                float4 hi, lo;
                hi.x = ssb_tf32_round(x[4 * q + 0]);
                hi.y = ssb_tf32_round(x[4 * q + 1]);
                hi.z = ssb_tf32_round(x[4 * q + 2]);
                hi.w = ssb_tf32_round(x[4 * q + 3]);
                lo.x = ssb_tf32_round(x[4 * q + 0] - hi.x);
                lo.y = ssb_tf32_round(x[4 * q + 1] - hi.y);
                lo.z = ssb_tf32_round(x[4 * q + 2] - hi.z);
                lo.w = ssb_tf32_round(x[4 * q + 3] - hi.w);
                *reinterpret_cast<float4*>(a_hi + (size_t)q * 16 * 32) = hi;
                *reinterpret_cast<float4*>(a_lo + (size_t)q * 16 * 32) = lo;
            }
        }
        ssb_fence_async();
        ssb_tc_fence_before();
        __syncthreads();
        ssb_tc_fence_after();
        if (threadIdx.x == 0) {
            ssb_mbar_wait(&full[b], (uint32_t)(i >> 1) & 1u);
            ssb_tc_fence_after();
            const float* ah = sA + (size_t)b * 2 * A_PART;
            const float* bh = sB + (size_t)b * 2 * B_PART;
#pragma unroll 1
            for (int j = 0; j < SSB_DTC_KS / 8; ++j) {
                const size_t oa = (size_t)j * 2 * 16 * 32, ob = (size_t)j * 2 * (SSB_DTC_N / 8) * 32;   // two 16-byte K chunks per MMA
                const uint64_t dah = ssb_umma_desc_lbo(ah + oa, 2048), dal = ssb_umma_desc_lbo(ah + A_PART + oa, 2048);
                const uint64_t dbh = ssb_umma_desc_lbo(bh + ob, (SSB_DTC_N / 8) * 128);
                const uint64_t dbl = ssb_umma_desc_lbo(bh + B_PART + ob, (SSB_DTC_N / 8) * 128);
                ssb_umma_tf32(tmem, dal, dbh, idesc, (i > 0 || j > 0) ? 1u : 0u);
                ssb_umma_tf32(tmem, dah, dbl, idesc, 1);
                ssb_umma_tf32(tmem, dah, dbh, idesc, 1);
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(ssb_smem(&done[b]))
                         : "memory");
        }
        __syncwarp();
    }
    float* vg = ssb_grp(c.vec, c.nv, g, lane);
    float* pg = ssb_grp(c.part, c.n_part, g, lane);
    if (my > 0) {
        // the commit of the last stage covers every earlier MMA
        ssb_mbar_wait(&done[(my - 1) & 1], (uint32_t)((my - 1) >> 1) & 1u);
        ssb_tc_fence_after();
#pragma unroll 1
        for (int cb = 0; cb < SSB_DTC_N / 64; ++cb) {       // this warp's half of the columns, 32 at a time
            const int c0 = half * (SSB_DTC_N / 2) + cb * 32;
            float v[32];
            ssb_tmem_ld32(tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)c0, v);
            if (live) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int jo = col0 + c0 + j;
                    if (jo < size_out) {
                        if (n_chunks == 1) vg[(size_t)(out_vec + jo) * 32] = v[j];
                        else pg[(size_t)(part_off + chunk * size_out + jo) * 32] = v[j];
                    }
                }
            }
        }
    } else if (live && n_chunks > 1) {                  // an empty trailing chunk still owns its partial slot
        for (int j = col0 + half * (SSB_DTC_N / 2); j < min(size_out, col0 + (half + 1) * (SSB_DTC_N / 2)); ++j)
            pg[(size_t)(part_off + chunk * size_out + j) * 32] = 0.f;
    }
    ssb_tc_fence_before();
    __threadfence();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(SSB_DTC_N));
    if (n_chunks == 1) return;
    // split-K: one arrival counter per (decoder, trial group); the CTA that arrives last adds the partials in chunk order
    if (half == 0) {
        if (lane == 0) {
            int last = 0;
            if (live) {
                int* cnt_p = c.counters + (d[8] + nt) * c.G + group;
                const int old = atomicAdd(cnt_p, 1);
                last = old == n_chunks - 1;
                if (last) *cnt_p = 0;
            }
            s_last[quad] = last;
        }
    }
    __syncthreads();
    if (!s_last[quad]) return;
    __threadfence();
    for (int j = col0 + half * (SSB_DTC_N / 2); j < min(size_out, col0 + (half + 1) * (SSB_DTC_N / 2)); j += 8) {   // the two warps of a group split the outputs
        float t[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) t[u] = 0.f;
        for (int ck0 = 0; ck0 < n_chunks; ck0 += 4) {
            float w[4][8];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const bool ok = ck0 + q < n_chunks && j + u < size_out;
                    w[q][u] = ok ? __ldcg(pg + (size_t)(part_off + (ck0 + q) * size_out + j + u) * 32) : 0.f;
                }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
#pragma unroll
                for (int u = 0; u < 8; ++u) t[u] += w[q][u];
            }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u)
            if (j + u < size_out) vg[(size_t)(out_vec + j + u) * 32] = t[u];
    }
}

