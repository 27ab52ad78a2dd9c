// sm_100a kernels of the SSP-SLAM step engine: grid clean-up: k_cleanup_scan (FFMA), k_cleanup_scan_tc (tcgen05), k_cleanup_pick, and the gate node k_gate.
// Included by ssb_kernels.cuh (after ssb_common.cuh); see that file for the layout rules.
#pragma once
#include "ssb_common.cuh"

// --------------------------------------------------------------------------------------
// Grid clean-up / decode: argmax_g S[g].x with first-maximum-wins.  Scan in fp32 keeping the
// top-4 candidates per (grid chunk, trial); the pick kernel re-scores near-ties in fp64 so
// that the chosen index equals the float64 NumPy argmax on the same input.
// A CTA = 4 warps = 4 different trial groups scanning the SAME grid chunk: the chunk of S is
// staged in shared memory tiles and read back as warp-uniform (broadcast) float4s, the query
// vector sits in registers; two grid rows are scored per iteration.
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

// desc: G d dpad s_off in_row0 out_vec ; scratch: cx[G][dpad][32], pval/pidx[G][n_cand][32]
// A CTA owns grid rows [blockIdx.x*rows_per_chunk, +rows_per_chunk) and walks them in shared-memory
// tiles of tile_rows rows.  dynamic smem: tile_rows*dpad (S tile)
template <int DP, bool CSR_INPUT>
__global__ void __launch_bounds__(128)
k_cleanup_scan(SsbCtx c, const int* __restrict__ d, const float* __restrict__ S, float* __restrict__ cx,
               float* __restrict__ pval, int* __restrict__ pidx, int rows_per_chunk, int tile_rows, int n_groups,
               int n_cand, int i_rel) {
    extern __shared__ float sm[];
    const int G = d[0], dims = d[1], dpad = d[2], in_row0 = d[4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int group = blockIdx.y * 4 + warp;
    const bool live = group < n_groups;
    const int g = live ? group : 0;
    const int g_lo = blockIdx.x * rows_per_chunk;
    const int g_hi = min(G, g_lo + rows_per_chunk);
    float* tile = sm;                              // [tile_rows][dpad]
    float* cxg = cx + ((size_t)g * dpad) * 32 + lane;
    float x[DP > 0 ? DP : 1];
    if (CSR_INPUT) {   // query = materialised vec rows of this step
        const float* vg = ssb_grp(c.vec, c.nv, g, lane);
        if (DP > 0) {
#pragma unroll
            for (int k = 0; k < DP; ++k) {
                x[k] = (k < dims) ? vg[(size_t)(in_row0 + k) * 32] : 0.f;
                if (blockIdx.x == 0 && live) cxg[(size_t)k * 32] = x[k];
            }
        } else if (blockIdx.x == 0 && live) {      // generic width: the query is re-read per tile; keep the copy for the pick
            for (int k = 0; k < dpad; ++k) cxg[(size_t)k * 32] = (k < dims) ? vg[(size_t)(in_row0 + k) * 32] : 0.f;
        }
    } else {
        if (DP > 0) {
#pragma unroll
            for (int k = 0; k < DP; ++k) x[k] = (k < dims) ? cxg[(size_t)k * 32] : 0.f;
        }
    }
    // generic width: source of the query columns (materialised vec rows, or the prepared stand-alone query)
    const float* xsrc = CSR_INPUT ? ssb_grp(c.vec, c.nv, g, lane) + (size_t)in_row0 * 32 : cxg;
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
            int gg = g0;
            for (; gg + 2 <= g1; gg += 2) {
                const float4* s4 = reinterpret_cast<const float4*>(tile + (size_t)(gg - g0) * DP);
                const float4* t4 = s4 + DP / 4;
                float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, b0 = 0.f, b1 = 0.f, b2 = 0.f, b3 = 0.f;
#pragma unroll
                for (int k4 = 0; k4 < DP / 4; ++k4) {
                    const float4 e = s4[k4], f = t4[k4];
                    a0 = fmaf(e.x, x[4 * k4 + 0], a0);
                    a1 = fmaf(e.y, x[4 * k4 + 1], a1);
                    a2 = fmaf(e.z, x[4 * k4 + 2], a2);
                    a3 = fmaf(e.w, x[4 * k4 + 3], a3);
                    b0 = fmaf(f.x, x[4 * k4 + 0], b0);
                    b1 = fmaf(f.y, x[4 * k4 + 1], b1);
                    b2 = fmaf(f.z, x[4 * k4 + 2], b2);
                    b3 = fmaf(f.w, x[4 * k4 + 3], b3);
                }
                ssb_top_push(top, (a0 + a1) + (a2 + a3), gg);
                ssb_top_push(top, (b0 + b1) + (b2 + b3), gg + 1);
            }
            for (; gg < g1; ++gg) {
                const float4* s4 = reinterpret_cast<const float4*>(tile + (size_t)(gg - g0) * DP);
                float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
                for (int k4 = 0; k4 < DP / 4; ++k4) {
                    const float4 e = s4[k4];
                    a0 = fmaf(e.x, x[4 * k4 + 0], a0);
                    a1 = fmaf(e.y, x[4 * k4 + 1], a1);
                    a2 = fmaf(e.z, x[4 * k4 + 2], a2);
                    a3 = fmaf(e.w, x[4 * k4 + 3], a3);
                }
                ssb_top_push(top, (a0 + a1) + (a2 + a3), gg);
            }
        } else {
            // generic width (any d, e.g. 649): the query is streamed in 32-column register chunks while the partial
            // scores of up to 16 tile rows stay in registers: 8 broadcast float4 grid reads per 32 FFMAs
            for (int gg0 = g0; gg0 < g1; gg0 += 16) {
                const int nr = min(16, g1 - gg0);
                float acc[16];
#pragma unroll
                for (int r = 0; r < 16; ++r) acc[r] = 0.f;
                for (int k0 = 0; k0 < dpad; k0 += 32) {
                    float xk[32];
#pragma unroll
                    for (int e = 0; e < 32; ++e) xk[e] = (k0 + e < dims) ? xsrc[(size_t)(k0 + e) * 32] : 0.f;
#pragma unroll
                    for (int r = 0; r < 16; ++r) {
                        if (r < nr) {
                            const float4* s4 = reinterpret_cast<const float4*>(tile + (size_t)(gg0 - g0 + r) * dpad + k0);
#pragma unroll
                            for (int q = 0; q < 8; ++q) {
                                if (k0 + 4 * q < dpad) {
                                    const float4 e = s4[q];
                                    acc[r] = fmaf(e.x, xk[4 * q + 0], acc[r]);
                                    acc[r] = fmaf(e.y, xk[4 * q + 1], acc[r]);
                                    acc[r] = fmaf(e.z, xk[4 * q + 2], acc[r]);
                                    acc[r] = fmaf(e.w, xk[4 * q + 3], acc[r]);
                                }
                            }
                        }
                    }
                }
#pragma unroll
                for (int r = 0; r < 16; ++r)
                    if (r < nr) ssb_top_push(top, acc[r], gg0 + r);
            }
        }
    }
    if (!live) return;
    float* pv = pval + ((size_t)g * n_cand) * 32 + lane;
    int* pi = pidx + ((size_t)g * n_cand) * 32 + lane;
#pragma unroll
    for (int i = 0; i < SSB_TOPK; ++i) {
        pv[(size_t)(blockIdx.x * SSB_TOPK + i) * 32] = top.v[i];
        pi[(size_t)(blockIdx.x * SSB_TOPK + i) * 32] = top.g[i];
    }
}

// --------------------------------------------------------------------------------------
// Tensor-core grid scan (tcgen05 + TMEM).  The similarity scores of a trial block against the sample
// grid are a real GEMM with weights shared by every trial: D[trial][grid row] = X[trial][k] . S[grid row][k].
// One CTA owns 128 trials (4 trial groups = the 128 TMEM lanes) and every n_chunks-th tile of 128 grid rows.
//   A = X  (128 x KP, K-major)  built once per CTA in shared memory from the materialised vec rows,
//   B = S  (128 x KP, K-major)  pre-tiled on the host in the UMMA core-matrix order, fetched by one TMA bulk
//                               copy per tile into a two-stage ring,
//   D      (128 lanes x 128 columns fp32) double-buffered in TMEM: the MMAs of tile i+1 run while the four
//                               warps drain tile i with tcgen05.ld and keep a per-trial top-4.
// fp32 accuracy comes from the 3xTF32 split: x = x_hi + x_lo with both parts exactly representable in
// TF32, D = X_lo.S_hi + X_hi.S_lo + X_hi.S_hi (the dropped lo.lo term is < 2^-22 relative).  Near-ties are
// still re-scored in fp64 by k_cleanup_pick, so the chosen index equals the float64 argmax.
//
// Shared-memory operand layout (UMMA "interleave" / no-swizzle, K-major): 8 rows x 16 bytes core matrices,
//   float offset(row r, column k) = ((k / 4) * 16 + r / 8) * 32 + (r % 8) * 4 + k % 4
// => stride between 8-row groups SBO = 128 B, stride between 16-byte K chunks LBO = 2048 B.

// Stc: [n_tiles][2 (hi, lo)][KP/4][TR/8][8][4] floats, TR = 128 grid rows per tile (64 when 128 does not fit in
// shared memory, e.g. d = 97).  dynamic smem: (2 * 128 + 4 * TR) * KP floats.
// 256 threads: warps w and w + 4 own the same TMEM lane quadrant (the 32 trials of group 4*blockIdx.y + w % 4)
// and drain the two halves of every tile's columns, each into its own top-4 list (candidate slot
// (2 * chunk + half) * 4 + i), so two warps per scheduler hide the insert latency.
// desc: G d dpad s_off in_row0 out_vec
template <bool CSR_INPUT, int TR>
__global__ void __launch_bounds__(256, 1)
k_cleanup_scan_tc(SsbCtx c, const int* __restrict__ d, const float* __restrict__ Stc, float* __restrict__ cx,
                  float* __restrict__ pval, int* __restrict__ pidx, int KP, int n_tiles, int n_groups, int n_cand) {
    extern __shared__ __align__(1024) float sm[];
    __shared__ unsigned long long full[2], done[2];
    __shared__ uint32_t tmem_slot;
    const int G = d[0], dims = d[1], dpad = d[2], in_row0 = d[4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int quad = warp & 3, half = warp >> 2;
    const int group = blockIdx.y * 4 + quad;
    const bool live = group < n_groups;
    const int g = live ? group : 0;
    const int chunk = blockIdx.x, n_chunks = gridDim.x;
    const int my_tiles = chunk < n_tiles ? (n_tiles - chunk + n_chunks - 1) / n_chunks : 0;
    const int part_floats = 128 * KP;                       // one part (hi or lo) of the A operand (128 trials)
    const int b_part = TR * KP;                             // one part of a grid tile (TR rows)
    const uint32_t tile_bytes = 2u * b_part * 4u;           // hi + lo
    float* sA = sm;                                         // [2][part]
    float* sB = sm + 2 * part_floats;                       // [2 stages][2][part]
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ssb_smem(&tmem_slot)), "r"(2 * TR));
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
            ssb_bulk_g2s(sB + (size_t)i * 2 * b_part, Stc + (size_t)(chunk + i * n_chunks) * 2 * b_part, tile_bytes, &full[i]);
        }
    }
    {   // A operand: this thread's trial is row r of the tile; four K columns per 16-byte store.
        // Loads are issued 32 at a time (8 chunks of 4 columns) before anything consumes them.
        const int r = quad * 32 + lane;
        const float* vg = ssb_grp(c.vec, c.nv, g, lane);
        float* cxg = cx + ((size_t)g * dpad) * 32 + lane;
        float* a_hi = sA + (r >> 3) * 32 + (r & 7) * 4;
        float* a_lo = a_hi + part_floats;
        const float* src = CSR_INPUT ? vg + (size_t)in_row0 * 32 : cxg;
        const bool copy_q = CSR_INPUT && live && blockIdx.x == 0;
        for (int k0 = half * 32; k0 < KP; k0 += 64) {   // the two warps of a quadrant alternate 32-column blocks
            float x[32];
#pragma unroll
            for (int e = 0; e < 32; ++e) {
                const int k = k0 + e;
                x[e] = (live && k < dims) ? src[(size_t)k * 32] : 0.f;
            }
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
            if (copy_q) {
#pragma unroll
                for (int e = 0; e < 32; ++e)
                    if (k0 + e < dpad) cxg[(size_t)(k0 + e) * 32] = x[e];
            }
        }
    }
    ssb_fence_async();            // generic-proxy stores of A -> visible to the tensor core (async proxy)
    ssb_tc_fence_before();
    __syncthreads();
    ssb_tc_fence_after();
    const uint32_t tmem = tmem_slot;
    // instruction descriptor: D fp32, A/B tf32, both K-major, N = 128, M = 128
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TR >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    auto issue_mma = [&](int i) {   // one thread: wait for the tile, queue its 3 * KP/8 MMAs, commit
        const int s = i & 1;
        ssb_mbar_wait(&full[s], (uint32_t)(i >> 1) & 1u);
        ssb_tc_fence_after();
        const float* b_hi = sB + (size_t)s * 2 * b_part;
        const float* b_lo = b_hi + b_part;
        const uint32_t dst = tmem + (uint32_t)s * TR;
        for (int j = 0; j < KP / 8; ++j) {
            const size_t off = (size_t)j * 2 * 16 * 32;     // two 16-byte K chunks per MMA
            const size_t ob = (size_t)j * 2 * (TR / 8) * 32;
            const uint64_t ah = ssb_umma_desc(sA + off), al = ssb_umma_desc(sA + part_floats + off);
            const uint64_t bh = ssb_umma_desc_lbo(b_hi + ob, (TR / 8) * 128), bl = ssb_umma_desc_lbo(b_lo + ob, (TR / 8) * 128);
            ssb_umma_tf32(dst, al, bh, idesc, j > 0);
            ssb_umma_tf32(dst, ah, bl, idesc, 1);
            ssb_umma_tf32(dst, ah, bh, idesc, 1);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(ssb_smem(&done[s]))
                     : "memory");
    };
    // per-trial top-4 in registers, sorted by (value desc, index asc); scores arrive in ascending index order
    float tv0 = -INFINITY, tv1 = -INFINITY, tv2 = -INFINITY, tv3 = -INFINITY;
    int tg0 = 0x7fffffff, tg1 = 0x7fffffff, tg2 = 0x7fffffff, tg3 = 0x7fffffff;
    auto push = [&](float val, int gi) {   // branch-free sorted insert (a strict > keeps the earlier index on ties)
        const bool b0 = val > tv0, b1 = val > tv1, b2 = val > tv2, b3 = val > tv3;
        tv3 = b2 ? tv2 : (b3 ? val : tv3);
        tg3 = b2 ? tg2 : (b3 ? gi : tg3);
        tv2 = b1 ? tv1 : (b2 ? val : tv2);
        tg2 = b1 ? tg1 : (b2 ? gi : tg2);
        tv1 = b0 ? tv0 : (b1 ? val : tv1);
        tg1 = b0 ? tg0 : (b1 ? gi : tg1);
        tv0 = b0 ? val : tv0;
        tg0 = b0 ? gi : tg0;
    };
    if (threadIdx.x == 0 && my_tiles > 0) issue_mma(0);
    __syncwarp();
    for (int i = 0; i < my_tiles; ++i) {
        const int s = i & 1;
        if (threadIdx.x == 0 && i + 1 < my_tiles) issue_mma(i + 1);
        __syncwarp();
        ssb_mbar_wait(&done[s], (uint32_t)(i >> 1) & 1u);
        ssb_tc_fence_after();
        if (threadIdx.x == 0 && i + 2 < my_tiles) {           // the MMAs of tile i have consumed stage s
            ssb_mbar_expect_tx(&full[s], tile_bytes);
            ssb_bulk_g2s(sB + (size_t)s * 2 * b_part, Stc + (size_t)(chunk + (i + 2) * n_chunks) * 2 * b_part, tile_bytes,
                         &full[s]);
        }
        __syncwarp();
        const int row0 = (chunk + i * n_chunks) * TR;
        const uint32_t taddr = tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)s * TR;
#pragma unroll 1
        for (int b = half * (TR / 64); b < (half + 1) * (TR / 64); ++b) {
            float v[32];
            ssb_tmem_ld32(taddr + b * 32, v);
            const int gg0 = row0 + b * 32;
            if (gg0 + 32 <= G) {
#pragma unroll
                for (int j = 0; j < 32; ++j) push(v[j], gg0 + j);
            } else {                                   // last tile: rows beyond the grid are padding
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (gg0 + j < G) push(v[j], gg0 + j);
            }
        }
        ssb_tc_fence_before();
        __syncthreads();           // every warp has drained TMEM buffer s before tile i+2 is accumulated into it
        ssb_tc_fence_after();
    }
    if (live) {
        float* pv = pval + ((size_t)g * n_cand) * 32 + lane;
        int* pi = pidx + ((size_t)g * n_cand) * 32 + lane;
        const float tv[4] = {tv0, tv1, tv2, tv3};
        const int tg[4] = {tg0, tg1, tg2, tg3};
#pragma unroll
        for (int i = 0; i < SSB_TOPK; ++i) {
            pv[(size_t)((blockIdx.x * 2 + half) * SSB_TOPK + i) * 32] = tv[i];
            pi[(size_t)((blockIdx.x * 2 + half) * SSB_TOPK + i) * 32] = tg[i];
        }
    }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(2 * TR));
}

// CTA = one trial group x 8 warps: warps split the candidate list, merge through shared memory, then
// candidates within eps of the fp32 maximum are re-scored in fp64 (S64 is the float64 grid) and
// the winning index / grid row are written.  out_base (may be null) is a group-tiled arena.
__global__ void __launch_bounds__(256)
k_cleanup_pick(int dims, int dpad, int ncand, const float* __restrict__ cx, const float* __restrict__ pval,
               const int* __restrict__ pidx, const double* __restrict__ S64, const float* __restrict__ S32,
               float* __restrict__ out_base, int out_rows_per_group, int out_row0, int* __restrict__ out_idx,
               const double* __restrict__ q64, long long q0, long long n_q, float eps_floor_rel) {
    __shared__ float sv[8][32];
    __shared__ int sg[8][32];
    __shared__ float sn[8][32];
    __shared__ int sc[8][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = blockIdx.x;
    const int trial = g * 32 + lane;
    const float* pv = pval + ((size_t)g * ncand) * 32 + lane;
    const int* pi = pidx + ((size_t)g * ncand) * 32 + lane;
    const float* cxg = cx + ((size_t)g * dpad) * 32 + lane;
    float best = -INFINITY;
    int best_g = 0x7fffffff;
    for (int i0 = warp; i0 < ncand; i0 += 64) {   // 8 independent candidate loads in flight per thread
        float v[8];
        int gi[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = i0 + 8 * u;
            v[u] = (i < ncand) ? pv[(size_t)i * 32] : -INFINITY;
            gi[u] = (i < ncand) ? pi[(size_t)i * 32] : 0x7fffffff;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (v[u] > best || (v[u] == best && gi[u] < best_g)) {
                best = v[u];
                best_g = gi[u];
            }
        }
    }
    float xn = 0.f;
    for (int k = warp; k < dims; k += 8) {
        const float xv = cxg[(size_t)k * 32];
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
        const int gi = sg[w][lane];
        if (v > best || (v == best && gi < best_g)) {
            best = v;
            best_g = gi;
        }
        xn += sn[w][lane];
    }
    // fp32 dot-product error bound: ~dims * 2^-24 * |S_g||x| with |S_g| = 1
    // (the 3xTF32 tensor-core scan passes its own relative floor: dropped lo.lo terms + fp32 accumulation)
    const float eps = fmaxf(4.0f * (float)dims * 5.97e-8f, eps_floor_rel) * sqrtf(xn) + 1e-30f;
    int n_close = 0;
    for (int i0 = warp; i0 < ncand; i0 += 64) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = (i0 + 8 * u < ncand) ? pv[(size_t)(i0 + 8 * u) * 32] : -INFINITY;
#pragma unroll
        for (int u = 0; u < 8; ++u) n_close += v[u] >= best - eps;
    }
    sc[warp][lane] = n_close;
    __syncthreads();
    n_close = 0;
    for (int w = 0; w < 8; ++w) n_close += sc[w][lane];
    // Near-ties (rare): lanes with more than one candidate inside the fp32 error band re-score those
    // candidates in fp64.  Every warp re-walks only its own slice of the candidate list, so a tie costs one
    // more pass instead of a serial scan; the per-warp winners are merged in (value desc, index asc) order.
    __shared__ double sd[8][32];
    const bool multi = n_close > 1 && S64 != nullptr;
    double dbest = -1e300;
    int dg = 0x7fffffff;
    if (__any_sync(0xffffffffu, multi)) {
        for (int i = warp; i < ncand; i += 8) {
            const float v = pv[(size_t)i * 32];
            const int gi = pi[(size_t)i * 32];
            if (multi && v >= best - eps && gi != 0x7fffffff) {
                const double* sgp = S64 + (size_t)gi * dims;
                double acc = 0.0;
                // argmax is invariant to the positive normalisation, so the raw float64 query can be used
                if (q64 != nullptr && q0 + trial < n_q) {
                    const double* qr = q64 + (size_t)(q0 + trial) * dims;
                    for (int k = 0; k < dims; ++k) acc += sgp[k] * qr[k];
                } else {
                    for (int k = 0; k < dims; ++k) acc += sgp[k] * (double)cxg[(size_t)k * 32];
                }
                if (acc > dbest || (acc == dbest && gi < dg)) {
                    dbest = acc;
                    dg = gi;
                }
            }
        }
    }
    __syncthreads();          // sg is re-used for the merge
    sd[warp][lane] = dbest;
    sg[warp][lane] = dg;
    __syncthreads();
    if (multi) {
        dbest = sd[0][lane];
        dg = sg[0][lane];
        for (int w = 1; w < 8; ++w) {
            const double v = sd[w][lane];
            const int gi = sg[w][lane];
            if (v > dbest || (v == dbest && gi < dg)) {
                dbest = v;
                dg = gi;
            }
        }
        best_g = dg;
    }
    if (out_idx && warp == 0) out_idx[trial] = best_g;
    if (out_base) {
        const float* sgp = S32 + (size_t)best_g * dpad;
        float* og = out_base + ((size_t)g * out_rows_per_group + out_row0) * 32 + lane;
        for (int k = warp; k < dims; k += 8) og[(size_t)k * 32] = sgp[k];
    }
}

// --------------------------------------------------------------------------------------
// K-blocked tensor-core grid scan for operand widths that do not fit in shared memory (d = 649: BASELINE configs[4]).
// Same GEMM, same 3xTF32 split, same top-4 epilogue as k_cleanup_scan_tc, but BOTH operands stream through a
// three-stage ring of K blocks of SSB_SCK_KB = 32 columns and the accumulator stays in TMEM across the K blocks:
//   k_scan_xtiles     X (the queries of 128 trials) -> hi | lo operand tiles in global memory, once per step:
//                     Xt[trial block][K block][hi | lo][KB/4][16][8][4]  (32 KB per (trial block, K block))
//   k_cleanup_scan_tck  CTA = (grid chunk, trial block); 10 warps:
//                     warps 0-7  epilogue (two per TMEM lane quadrant, as in k_cleanup_scan_tc),
//                     warp 8     TMA producer: per (grid tile, K block) one bulk copy of the X block and one of the
//                                S block into the stage, as soon as the MMAs that read the stage have retired,
//                     warp 9     MMA issuer: 4 x 3 tcgen05.mma (128 x 128 x 8, tf32) per stage; tcgen05.commit frees
//                                the stage; after the last K block a second commit publishes the D buffer.
//                     D (128 lanes x 128 columns) is double-buffered in TMEM: tile i+1 accumulates while tile i drains.
// Stck: [n_tiles][n_kb][hi | lo][KB/4][16][8][4] floats (host pre-tiled).  dynamic smem: 3 x 64 KB.
#define SSB_SCK_KB 32
#define SSB_SCK_NST 3
#define SSB_SCK_PART (128 * SSB_SCK_KB)          // floats of one operand part (hi or lo) of one K block

// grid (n_kb, trial blocks) x 128: thread = one trial (row of the A tile), 32 columns.
template <bool CSR_INPUT>
__global__ void __launch_bounds__(128)
k_scan_xtiles(SsbCtx c, const int* __restrict__ d0, float* __restrict__ cx, float* __restrict__ Xt0, int n_kb, int n_groups,
              long long z_stride) {
    // blockIdx.z: further (desc, tile buffer) pairs of one launch (the wide-ensemble encode tiles several inputs at once)
    const int* d = d0 + blockIdx.z * 6;
    float* Xt = Xt0 + (size_t)blockIdx.z * z_stride;
    const int dims = d[1], dpad = d[2], in_row0 = d[4];
    const int lane = threadIdx.x & 31, quad = threadIdx.x >> 5;
    const int kb = blockIdx.x, tb = blockIdx.y;
    const int group = tb * 4 + quad;
    const bool live = group < n_groups;
    const int g = live ? group : 0;
    const int r = quad * 32 + lane;
    const float* vg = ssb_grp(c.vec, c.nv, g, lane);
    float* cxg = cx + ((size_t)g * dpad) * 32 + lane;        // (cx == nullptr: no query copy wanted, CSR input only)
    const float* src = CSR_INPUT ? vg + (size_t)in_row0 * 32 : cxg;
    float* a_hi = Xt + ((size_t)tb * n_kb + kb) * 2 * SSB_SCK_PART + (r >> 3) * 32 + (r & 7) * 4;
    float* a_lo = a_hi + SSB_SCK_PART;
    const int k0 = kb * SSB_SCK_KB;
    float x[SSB_SCK_KB];
#pragma unroll
    for (int e = 0; e < SSB_SCK_KB; ++e) x[e] = (live && k0 + e < dims) ? src[(size_t)(k0 + e) * 32] : 0.f;
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
    if (CSR_INPUT && live && cx != nullptr) {    // the query copy k_cleanup_pick re-scores near-ties with
#pragma unroll
        for (int e = 0; e < SSB_SCK_KB; ++e)
            if (k0 + e < dpad) cxg[(size_t)(k0 + e) * 32] = x[e];
    }
}

// desc: G d dpad s_off in_row0 out_vec
__global__ void __launch_bounds__(320, 1)
k_cleanup_scan_tck(const int* __restrict__ d, const float* __restrict__ Stck, const float* __restrict__ Xt,
                   float* __restrict__ pval, int* __restrict__ pidx, int n_kb, int n_tiles, int n_groups, int n_cand) {
    extern __shared__ __align__(1024) float sm[];
    __shared__ unsigned long long full[SSB_SCK_NST], empty[SSB_SCK_NST], dfull[2], dfree[2];
    __shared__ uint32_t tmem_slot;
    constexpr int TR = 128;
    const int G = d[0];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int chunk = blockIdx.x, n_chunks = gridDim.x, tb = blockIdx.y;
    const int my_tiles = chunk < n_tiles ? (n_tiles - chunk + n_chunks - 1) / n_chunks : 0;
    constexpr uint32_t blk_bytes = 2u * SSB_SCK_PART * 4u;          // hi + lo of one operand block: 32 KB
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
            ssb_mbar_init(&dfree[i], 8);                               // one arrival per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    ssb_tc_fence_before();
    __syncthreads();
    ssb_tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const int n_blocks = my_tiles * n_kb;                            // (tile, K block) pairs of this CTA, in order
    if (warp == 8) {
        // ---------------- TMA producer
        if (lane == 0) {
            for (int q = 0; q < n_blocks; ++q) {
                const int st = q % SSB_SCK_NST, round = q / SSB_SCK_NST;
                if (round > 0) ssb_mbar_wait(&empty[st], (uint32_t)(round - 1) & 1u);
                const int i = q / n_kb, kb = q - i * n_kb;
                const int tile = chunk + i * n_chunks;
                float* dst = sm + (size_t)st * 4 * SSB_SCK_PART;
                ssb_mbar_expect_tx(&full[st], 2u * blk_bytes);
                ssb_bulk_g2s(dst, Xt + ((size_t)tb * n_kb + kb) * 2 * SSB_SCK_PART, blk_bytes, &full[st]);
                ssb_bulk_g2s(dst + 2 * SSB_SCK_PART, Stck + ((size_t)tile * n_kb + kb) * 2 * SSB_SCK_PART, blk_bytes, &full[st]);
            }
        }
    } else if (warp == 9) {
        // ---------------- MMA issuer
        if (lane == 0) {
            // instruction descriptor: D fp32, A/B tf32, both K-major, N = 128, M = 128
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TR >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            int q = 0;
            for (int i = 0; i < my_tiles; ++i) {
                const int buf = i & 1;
                if (i >= 2) {                                        // the epilogue has drained this D buffer (tile i - 2)
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
                        const size_t off = (size_t)j * 2 * 16 * 32;        // two 16-byte K chunks per MMA
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
        // ---------------- epilogue: per-trial top-4, two warps per TMEM lane quadrant
        const int quad = warp & 3, half = warp >> 2;
        const int group = tb * 4 + quad;
        const bool live = group < n_groups;
        const int g = live ? group : 0;
        float tv0 = -INFINITY, tv1 = -INFINITY, tv2 = -INFINITY, tv3 = -INFINITY;
        int tg0 = 0x7fffffff, tg1 = 0x7fffffff, tg2 = 0x7fffffff, tg3 = 0x7fffffff;
        auto push = [&](float val, int gi) {   // branch-free sorted insert (a strict > keeps the earlier index on ties)
            const bool b0 = val > tv0, b1 = val > tv1, b2 = val > tv2, b3 = val > tv3;
            tv3 = b2 ? tv2 : (b3 ? val : tv3);
            tg3 = b2 ? tg2 : (b3 ? gi : tg3);
            tv2 = b1 ? tv1 : (b2 ? val : tv2);
            tg2 = b1 ? tg1 : (b2 ? gi : tg2);
            tv1 = b0 ? tv0 : (b1 ? val : tv1);
            tg1 = b0 ? tg0 : (b1 ? gi : tg1);
            tv0 = b0 ? val : tv0;
            tg0 = b0 ? gi : tg0;
        };
        for (int i = 0; i < my_tiles; ++i) {
            const int buf = i & 1;
            ssb_mbar_wait(&dfull[buf], (uint32_t)(i >> 1) & 1u);
            ssb_tc_fence_after();
            const int row0 = (chunk + i * n_chunks) * TR;
            const uint32_t taddr = tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)buf * TR;
#pragma unroll 1
            for (int b = half * 2; b < half * 2 + 2; ++b) {
                float v[32];
                ssb_tmem_ld32(taddr + b * 32, v);
                const int gg0 = row0 + b * 32;
                if (gg0 + 32 <= G) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) push(v[j], gg0 + j);
                } else {                                   // last tile: rows beyond the grid are padding
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (gg0 + j < G) push(v[j], gg0 + j);
                }
            }
            ssb_tc_fence_before();
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(ssb_smem(&dfree[buf])) : "memory");
        }
        if (live) {
            float* pv = pval + ((size_t)g * n_cand) * 32 + lane;
            int* pi = pidx + ((size_t)g * n_cand) * 32 + lane;
            const float tv[4] = {tv0, tv1, tv2, tv3};
            const int tg[4] = {tg0, tg1, tg2, tg3};
#pragma unroll
            for (int i = 0; i < SSB_TOPK; ++i) {
                pv[(size_t)((blockIdx.x * 2 + half) * SSB_TOPK + i) * 32] = tv[i];
                pi[(size_t)((blockIdx.x * 2 + half) * SSB_TOPK + i) * 32] = tg[i];
            }
        }
    }
    ssb_tc_fence_before();
    __syncthreads();
    if (warp == 9) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(2 * TR));
}

// --------------------------------------------------------------------------------------
// Gated correction node (slam.py:233-237): x = [p ; q ; flag].  CTA = one trial group x 8 warps;
// warps split the dimensions, the dot product is reduced through shared memory.
// desc: d in_row0 out_vec rate_bits thres_bits atol_bits
__global__ void __launch_bounds__(256) k_gate(SsbCtx c, const int* __restrict__ desc, int item0, int i_rel) {
    __shared__ float part[8][32];
    const int* d = desc + (item0 + blockIdx.y) * 6;
    const int dims = d[0], in_row0 = d[1], out_vec = d[2];
    const float rate = __int_as_float(d[3]), thres = __int_as_float(d[4]), atol = __int_as_float(d[5]);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = blockIdx.x;
    float* vg = ssb_grp(c.vec, c.nv, g, lane);
    // pass 1: p - q goes to the output slot, p.q is reduced over the 8 warps
    float dot = 0.f;
    for (int k = warp; k < dims; k += 8) {
        const float p = vg[(size_t)(in_row0 + k) * 32];
        const float q = vg[(size_t)(in_row0 + dims + k) * 32];
        dot = fmaf(p, q, dot);
        vg[(size_t)(out_vec + k) * 32] = p - q;
    }
    part[warp][lane] = dot;
    __syncthreads();
    dot = 0.f;
    for (int w = 0; w < 8; ++w) dot += part[w][lane];
    const float flag = vg[(size_t)(in_row0 + 2 * dims) * 32];
    const bool open = (fabsf(flag) <= atol) && (dot > thres);
    // pass 2: each thread rescales the values it wrote itself
    for (int k = warp; k < dims; k += 8) {
        float* o = vg + (size_t)(out_vec + k) * 32;
        *o = open ? rate * *o : 0.f;
    }
}

