// sm_100a kernels of the SSP-SLAM step engine: static wide ensembles with very wide inputs (d = 649) on tcgen05.
// Included by ssb_kernels.cuh after ssb_cleanup.cuh (it shares the K-blocked operand format of k_cleanup_scan_tck).
#pragma once
#include "ssb_common.cuh"
#include "ssb_cleanup.cuh"

// --------------------------------------------------------------------------------------
// K-blocked tensor-core encode + neuron update.  The currents of a static wide ensemble are the same GEMM as the grid scan,
//     J[trial][neuron] = X[trial][k] . E[neuron][k]        (E = scaled encoders, shared by every trial),
// so the kernel is k_cleanup_scan_tck with another epilogue: both operands stream through the three-stage ring of
// 32-column K blocks (X tiles from k_scan_xtiles, encoder tiles pre-tiled on the host), the accumulator of a tile of 128
// neurons stays in TMEM across the K blocks, and the eight epilogue warps (two per TMEM lane quadrant = trial group, each
// its half of the tile's columns) add the bias / direct neuron currents, run the LIF (or rate) update on the state row
// of their trial group and store state, activity and the activity flag word.
//   Etk: [n_tiles][n_kb][hi | lo][KB/4][16][8][4] floats of ONE ensemble; Xt: [trial block][n_kb][hi | lo]... of its input
// grid (tile chunks, trial blocks, ensembles) x 320; dynamic smem 3 x 64 KB.
// desc (big-ensemble descriptor): n dims dpad state0 act0 enc_off bias_off in_row0 ntype flags jn_row0 jn_m jn_w ...
struct SsbTckItems {
    int n;
    int idx[15];             // big-ensemble descriptor index
    long long e_off[15];     // float offset of the ensemble's encoder tiles
    long long x_off[15];     // float offset of its X tiles
};

__global__ void __launch_bounds__(320, 1)
k_wide_static_tck(SsbCtx c, const int* __restrict__ desc, SsbTckItems items, const float* __restrict__ Etk_all,
                  const float* __restrict__ Xt_all, int n_kb) {
    extern __shared__ __align__(1024) float sm[];
    __shared__ unsigned long long full[SSB_SCK_NST], empty[SSB_SCK_NST], dfull[2], dfree[2];
    __shared__ uint32_t tmem_slot;
    constexpr int TR = 128;
    const int* d = desc + items.idx[blockIdx.z] * 16;
    const int n = d[0], state0 = d[3], act0 = d[4], bias_off = d[6], jn_row0 = d[10], jn_m = d[11], jn_w = d[12];
    const float* __restrict__ Etk = Etk_all + items.e_off[blockIdx.z];
    const float* __restrict__ Xt = Xt_all + items.x_off[blockIdx.z];
    const int n_tiles = (n + TR - 1) / TR;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int chunk = blockIdx.x, n_chunks = gridDim.x, tb = blockIdx.y;
    const int my_tiles = chunk < n_tiles ? (n_tiles - chunk + n_chunks - 1) / n_chunks : 0;
    constexpr uint32_t blk_bytes = 2u * SSB_SCK_PART * 4u;
    if (warp == 9) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ssb_smem(&tmem_slot)), "r"(2 * TR));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (threadIdx.x == 0) {
        for (int i = 0; i < SSB_SCK_NST; ++i) {
            ssb_mbar_init(&full[i], 1);
            ssb_mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            ssb_mbar_init(&dfull[i], 1);
            ssb_mbar_init(&dfree[i], 8);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    ssb_tc_fence_before();
    __syncthreads();
    ssb_tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const int n_blocks = my_tiles * n_kb;
    if (warp == 8) {
        if (lane == 0) {                                             // TMA producer
            for (int q = 0; q < n_blocks; ++q) {
                const int st = q % SSB_SCK_NST, round = q / SSB_SCK_NST;
                if (round > 0) ssb_mbar_wait(&empty[st], (uint32_t)(round - 1) & 1u);
                const int i = q / n_kb, kb = q - i * n_kb;
                const int tile = chunk + i * n_chunks;
                float* dst = sm + (size_t)st * 4 * SSB_SCK_PART;
                ssb_mbar_expect_tx(&full[st], 2u * blk_bytes);
                ssb_bulk_g2s(dst, Xt + ((size_t)tb * n_kb + kb) * 2 * SSB_SCK_PART, blk_bytes, &full[st]);
                ssb_bulk_g2s(dst + 2 * SSB_SCK_PART, Etk + ((size_t)tile * n_kb + kb) * 2 * SSB_SCK_PART, blk_bytes, &full[st]);
            }
        }
    } else if (warp == 9) {
        if (lane == 0) {                                             // MMA issuer
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TR >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            int q = 0;
            for (int i = 0; i < my_tiles; ++i) {
                const int buf = i & 1;
                if (i >= 2) {
                    ssb_mbar_wait(&dfree[buf], (uint32_t)((i >> 1) - 1) & 1u);
                    ssb_tc_fence_after();
                }
                const uint32_t dst = tmem + (uint32_t)buf * TR;
                for (int kb = 0; kb < n_kb; ++kb, ++q) {
                    const int st = q % SSB_SCK_NST;
                    ssb_mbar_wait(&full[st], (uint32_t)(q / SSB_SCK_NST) & 1u);
                    ssb_tc_fence_after();
                    const float* a_hi = sm + (size_t)st * 4 * SSB_SCK_PART;
                    const float* a_lo = a_hi + SSB_SCK_PART;
                    const float* b_hi = a_hi + 2 * SSB_SCK_PART;
                    const float* b_lo = b_hi + SSB_SCK_PART;
#pragma unroll
                    for (int j = 0; j < SSB_SCK_KB / 8; ++j) {
                        const size_t off = (size_t)j * 2 * 16 * 32;
                        const uint64_t ah = ssb_umma_desc(a_hi + off), al = ssb_umma_desc(a_lo + off);
                        const uint64_t bh = ssb_umma_desc(b_hi + off), bl = ssb_umma_desc(b_lo + off);
                        ssb_umma_tf32(dst, al, bh, idesc, (kb > 0 || j > 0) ? 1u : 0u);
                        ssb_umma_tf32(dst, ah, bl, idesc, 1);
                        ssb_umma_tf32(dst, ah, bh, idesc, 1);
                    }
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                                     ssb_smem(&empty[st]))
                                 : "memory");
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                                 ssb_smem(&dfull[buf]))
                             : "memory");
            }
        }
    } else {
        // ---------------- epilogue: bias + direct currents, neuron update, state / activity / flag stores
        const int quad = warp & 3, half = warp >> 2;
        const int group = tb * 4 + quad;
        const bool live = group < c.G;
        const int g = live ? group : 0;
        const SsbNeuron nt = ssb_neuron(c, d[8]);
        const bool stateful = nt.type == 0;
        const float* vg = ssb_grp(c.vec, c.nv, g, lane);
        const int jm = min(jn_m, 4);
        float u_jn[4];
#pragma unroll
        for (int m = 0; m < 4; ++m) u_jn[m] = (m < jm) ? vg[(size_t)(jn_row0 + m) * 32] : 0.f;
        float* sg = ssb_grp(c.st, c.nn, g, lane) + (size_t)state0 * 32;
        float* ag = ssb_grp(c.act, c.n_act, g, lane) + (size_t)act0 * 32;
        int* flg = c.aflag + (size_t)g * c.n_act + act0;
        for (int i = 0; i < my_tiles; ++i) {
            const int buf = i & 1;
            const int nn_tile = (chunk + i * n_chunks) * TR;
            ssb_mbar_wait(&dfull[buf], (uint32_t)(i >> 1) & 1u);
            ssb_tc_fence_after();
            const uint32_t taddr = tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)buf * TR;
#pragma unroll 1
            for (int b = half * 2; b < half * 2 + 2; ++b) {
                const int nn0 = nn_tile + b * 32;
                float sv[32];
                if (stateful) {                             // the state rows are in flight while TMEM is read
#pragma unroll
                    for (int j = 0; j < 32; ++j) sv[j] = (live && nn0 + j < n) ? __ldcs(sg + (size_t)(nn0 + j) * 32) : 0.f;
                }
                float v[32];
                ssb_tmem_ld32(taddr + b * 32, v);
                if (live) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int nn = nn0 + j;
                        if (nn < n) {
                            float J = v[j] + __ldg(c.W + bias_off + nn);
                            for (int m = 0; m < jm; ++m) J = fmaf(__ldg(c.W + jn_w + nn * jn_m + m), u_jn[m], J);
                            float out;
                            if (stateful) {
                                float st = sv[j];
                                out = nt.fast ? ssb_lif_packed<true>(nt, J, st) : ssb_lif_packed<false>(nt, J, st);
                                __stcs(sg + (size_t)nn * 32, st);
                            } else {
                                out = ssb_rate(nt, J);
                            }
                            ag[(size_t)nn * 32] = out;
                            const unsigned any_on = __ballot_sync(0xffffffffu, out != 0.f);
                            if (lane == 0) flg[nn] = (int)any_on;
                        }
                    }
                }
            }
            ssb_tc_fence_before();
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(ssb_smem(&dfree[buf])) : "memory");
        }
    }
    ssb_tc_fence_before();
    __syncthreads();
    if (warp == 9) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(2 * TR));
}
