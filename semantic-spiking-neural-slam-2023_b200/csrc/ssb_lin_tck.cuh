// sm_100a kernels of the SSP-SLAM step engine: large dense blocks of the row program on tcgen05 (d = 649).
// Included by ssb_kernels.cuh after ssb_lin.cuh and ssb_cleanup.cuh (it shares the K-blocked operand format of the scan).
#pragma once
#include "ssb_common.cuh"
#include "ssb_cleanup.cuh"
#include "ssb_lin.cuh"

// --------------------------------------------------------------------------------------
// A dense block of the row program is out[r] = sum_k T[r][k] * vec[cols[k]] for R sink rows that share one column list -
// with weights T shared by every trial, i.e. a GEMM  D[trial][r] = X[trial][k] . T[r][k].  At d = 55 the blocks are
// 112 x 55 and stay on the FFMA items of k_lin; at d = 649 (BASELINE configs[4]) they are the 1 300 x 649 / 649 x 1 300
// circular-convolution DFT matrices and the 975 x 649 Fourier layouts, 3 x 285 us of FFMA per step.  Blocks with
// R >= 128 and K >= 256 therefore take the K-blocked tcgen05 path of the grid scan:
//   k_lin_xtiles  gathers the source rows of a block (column list of this step's parity) into hi | lo operand tiles,
//   k_lin_tck     the scan's pipeline (TMA producer warp, MMA issuer warp, accumulator in TMEM across the K blocks) with
//                 the row program's store as epilogue: filter update / probe sample / materialised row (ssb_lin_store).
struct SsbLinTcBlock {
    int R, K, n_kb, n_tiles;
    int cols_off, kpad, view, rows_off;
    long long t_off;          // float offset of the block's T tiles  [tile][kb][hi | lo][KB/4][16][8][4]
    long long x_off;          // float offset of its X tiles          [trial block][kb][hi | lo]...
};

// grid (max n_kb, trial blocks, blocks) x 128: thread = one trial (row of the A tile), 32 columns.
__global__ void __launch_bounds__(128)
k_lin_xtiles(SsbCtx c, const SsbLinTcBlock* __restrict__ blocks, const int* __restrict__ dcols, float* __restrict__ Xt_all,
             int i_rel) {
    const SsbLinTcBlock b = blocks[blockIdx.z];
    const int kb = blockIdx.x, tb = blockIdx.y;
    if (kb >= b.n_kb) return;
    const int lane = threadIdx.x & 31, quad = threadIdx.x >> 5;
    const int group = tb * 4 + quad;
    const bool live = group < c.G;
    const int g = live ? group : 0;
    const int r = quad * 32 + lane;
    const SsbStep s = ssb_step(c, i_rel);
    const int* __restrict__ cols = dcols + b.cols_off + ((s.odd ^ b.view) ? b.kpad : 0);
    const float* vg = ssb_grp(c.vec, c.nv, g, lane);
    float* a_hi = Xt_all + b.x_off + ((size_t)tb * b.n_kb + kb) * 2 * SSB_SCK_PART + (r >> 3) * 32 + (r & 7) * 4;
    float* a_lo = a_hi + SSB_SCK_PART;
    const int k0 = kb * SSB_SCK_KB;
    float x[SSB_SCK_KB];
#pragma unroll
    for (int e = 0; e < SSB_SCK_KB; ++e) x[e] = (live && k0 + e < b.K) ? ssb_ld_src(vg + (size_t)__ldg(cols + k0 + e) * 32) : 0.f;
#pragma unroll
    for (int q = 0; q < SSB_SCK_KB / 4; ++q) {
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

// grid (tile chunks, trial blocks, blocks) x 320; dynamic smem 3 x 64 KB.
__global__ void __launch_bounds__(320, 1)
k_lin_tck(SsbCtx c, const SsbLinTcBlock* __restrict__ blocks, const float* __restrict__ Ttk_all,
          const float* __restrict__ Xt_all, const int* __restrict__ drows, int i_rel) {
    extern __shared__ __align__(1024) float sm[];
    __shared__ unsigned long long full[SSB_SCK_NST], empty[SSB_SCK_NST], dfull[2], dfree[2];
    __shared__ uint32_t tmem_slot;
    constexpr int TR = 128;
    const SsbLinTcBlock blk = blocks[blockIdx.z];
    const int n_kb = blk.n_kb, n_tiles = blk.n_tiles;
    const float* __restrict__ Ttk = Ttk_all + blk.t_off;
    const float* __restrict__ Xt = Xt_all + blk.x_off;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int chunk = blockIdx.x, n_chunks = gridDim.x, tb = blockIdx.y;
    const int my_tiles = chunk < n_tiles ? (n_tiles - chunk + n_chunks - 1) / n_chunks : 0;
    if (my_tiles == 0) return;
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
                ssb_bulk_g2s(dst + 2 * SSB_SCK_PART, Ttk + ((size_t)tile * n_kb + kb) * 2 * SSB_SCK_PART, blk_bytes, &full[st]);
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
        // ---------------- epilogue: the row program's store (filter update, probe sample, materialised row)
        const int quad = warp & 3, half = warp >> 2;
        const int group = tb * 4 + quad;
        const bool live = group < c.G;
        const int g = live ? group : 0;
        const SsbStep s = ssb_step(c, i_rel);
        float* vg = ssb_grp(c.vec, c.nv, g, lane);
        const int4* __restrict__ dr = reinterpret_cast<const int4*>(drows) + blk.rows_off;
        for (int i = 0; i < my_tiles; ++i) {
            const int buf = i & 1;
            const int r_tile = (chunk + i * n_chunks) * TR;
            ssb_mbar_wait(&dfull[buf], (uint32_t)(i >> 1) & 1u);
            ssb_tc_fence_after();
            const uint32_t taddr = tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)buf * TR;
#pragma unroll 1
            for (int bq = half * 2; bq < half * 2 + 2; ++bq) {
                float v[32];
                ssb_tmem_ld32(taddr + bq * 32, v);
                const int r0 = r_tile + bq * 32;
                if (live) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        if (r0 + j < blk.R) {
                            const int4 rd = __ldg(dr + r0 + j);
                            ssb_lin_store(c, s, vg, g, lane, rd.x, rd.y, __int_as_float(rd.z), __int_as_float(rd.w), v[j]);
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
