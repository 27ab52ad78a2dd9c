// sm_100a kernels of the SSP-SLAM step engine (device side).
//
// Layout rule: every per-trial arena is tiled by trial group, arena[g][row][32]: a warp's 32
// lanes are the 32 trials of one group, a row is one 128-byte line, and consecutive rows of one
// group are contiguous, so each warp streams a sequential address range (DRAM-page friendly) and
// whole neuron ranges can be moved by 1-D TMA bulk copies.  Everything indexed by neuron /
// weight is warp-uniform (broadcast from shared memory or L1/L2).  Static weights are shared by
// all trials; learned matrices (Voja encoders, PES decoders) are per-trial rows of the same form.
//
// LIF state is ONE word per neuron: s >= 0 is the membrane voltage of a neuron that is not
// refractory, s < 0 is minus the remaining refractory time (the voltage is exactly 0 then).  With
// nengo's default min_voltage = 0 this is equivalent to the (voltage, refractory_time) pair: the
// refractory time only matters while it is >= dt, and the voltage is pinned to 0 exactly then.
//
// Semantics restate nengo's operators, executed in dependency
// levels instead of one operator at a time; the CPU checker is oracle/nengo_ref_sim.py.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#define SSB_TOPK 4
#define SSB_SCAN_MAX_CHUNKS 512
#define SSB_SM_CH 16          // neurons per TMA-staged chunk of the narrow-ensemble kernel
#define SSB_SM_WMAX 16        // max packed weight stride (floats) of a narrow ensemble

struct SsbCtx {
    int G;                    // trial groups (32 trials each)
    int nv, nf, nt, tab_row0, nn, n_act, n_lenc, n_ldec, n_afilt, n_probe, n_part;
    int tab_cap, probe_cap;
    float dt;
    float* vec;               // [G][nv][32]   0: ones | 1..nf: filters A | nf+1..2nf: filters B | tables | scratch
    const float* tab;         // [G][tab_cap][nt][32]
    float* st;                // [G][nn][32]   packed LIF state
    float* act;               // [G][n_act][32]
    int* aflag;               // [G][n_act]: bit t set = trial t of the group has a non-zero activity (sparse consumers skip on it)
    float* lenc;              // [G][n_lenc][32]
    float* ldec;              // [G][n_ldec][32]
    float* afilt;             // [G][2*n_afilt][32]
    float* probe;             // [G][probe_cap][n_probe][32]
    float* part;              // [G][n_part][32] split-K partial sums of decode / PES launches
    int* counters;            // split-K arrival counters, self-resetting
    const float* W;           // shared static weights
    const float* wpt;         // [G][n_wpt][32] per-trial static weights (trials with their own network seed), or nullptr
    int n_wpt;
    const int* csr_ptr;
    const int2* ent0;         // CSR entries (vec row, coefficient bits) resolved for even steps
    const int2* ent1;         //   ... and for odd steps (filter columns point at the other half)
    const float* ntypes;      // [n][8] = type, tau_rc, tau_ref, min_voltage, amplitude, fast_math, -, -
    const long long* dyn;     // [0] completed steps, [1] first step of resident tables, [2] first step of probe buffer,
                              // [3] first step of the resident synthesis indices
};

struct SsbStep {
    long long step;
    int odd;
    const int2* ent_old;      // rows evaluated on the values this step reads (old filter states)
    const int2* ent_new;      // ... on the half the previous step read (PES error of step-1)
    int par_old, par_new;     // row offset of the filter half read / written this step
};

__device__ __forceinline__ SsbStep ssb_step(const SsbCtx& c, int i_rel) {
    SsbStep s;
    s.step = c.dyn[0] + i_rel;
    s.odd = (int)(s.step & 1);
    s.ent_old = s.odd ? c.ent1 : c.ent0;
    s.ent_new = s.odd ? c.ent0 : c.ent1;
    s.par_old = s.odd ? c.nf : 0;
    s.par_new = s.odd ? 0 : c.nf;
    return s;
}

// Group base pointers (lane already added): element of row r is p[r * 32].
__device__ __forceinline__ float* ssb_grp(float* base, int rows, int g, int lane) {
    return base + ((size_t)g * rows) * 32 + lane;
}

// One sink row: sparse linear combination of vec rows for the 32 trials of a group.  Entries are
// warp-uniform 8-byte loads; the host pads every row to a multiple of 8 entries with (row 0,
// coefficient 0), so the loop has no tail and the 8 per-trial source loads of a batch are independent.
// `asm volatile` loads keep program order, so the compiler cannot re-serialise a batch to save registers:
// all entry loads of a batch are issued, then all source loads, then the multiply-adds.
__device__ __forceinline__ int2 ssb_ld_ent(const int2* p) {
    int2 v;
    asm volatile("ld.global.nc.v2.s32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ float ssb_ld_src(const float* p) {
    float v;
    asm volatile("ld.global.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

template <int NB>
__device__ __forceinline__ float ssb_row_batch(const int2* __restrict__ ent, const float* vg, float acc) {
    int2 e[NB];
    float x[NB];
#pragma unroll
    for (int u = 0; u < NB; ++u) e[u] = ssb_ld_ent(ent + u);
#pragma unroll
    for (int u = 0; u < NB; ++u) x[u] = ssb_ld_src(vg + (size_t)e[u].x * 32);
#pragma unroll
    for (int u = 0; u < NB; ++u) acc = fmaf(__int_as_float(e[u].y), x[u], acc);
    return acc;
}

// --------------------------------------------------------------------------------------
// Neuron models.
struct SsbNeuron {
    int type;                 // 0 LIF, 1 LIFRate, 2 RectifiedLinear
    bool fast;                // LIF with dt/tau_rc <= 1/16: polynomial expm1 / log1p, exact to fp32
    float tau_rc, tau_ref, amp_dt, amp, dt, inv_dt, neg_dt_over_tau, c0;
};

__device__ __forceinline__ SsbNeuron ssb_neuron(const SsbCtx& c, int tid) {
    const float* p = c.ntypes + tid * 8;
    SsbNeuron n;
    n.type = (int)p[0];
    n.tau_rc = p[1];
    n.tau_ref = p[2];
    n.amp = p[4];
    n.fast = p[5] != 0.f;
    n.dt = c.dt;
    n.amp_dt = p[4] / c.dt;
    n.inv_dt = 1.0f / c.dt;
    n.neg_dt_over_tau = (n.type == 0) ? -c.dt / p[1] : 0.f;
    n.c0 = p[2] + c.dt;       // tau_ref + dt
    return n;
}

// expm1(x) for -1/16 <= x <= 0 (x = -delta/tau_rc with delta <= dt): degree-5 Taylor, rel. error < 2e-9.
__device__ __forceinline__ float ssb_expm1_small(float x) {
    float p = 1.f / 120.f;
    p = fmaf(p, x, 1.f / 24.f);
    p = fmaf(p, x, 1.f / 6.f);
    p = fmaf(p, x, 0.5f);
    p = fmaf(p, x, 1.f);
    return p * x;
}

// log1p(-z) for 0 <= z <= 1/16 (z = overshoot / (J - 1) <= 1 - exp(-dt/tau_rc)): 6 terms, rel. error < 1e-8.
__device__ __forceinline__ float ssb_log1p_neg_small(float z) {
    float p = -1.f / 6.f;
    p = fmaf(p, z, -0.2f);
    p = fmaf(p, z, -0.25f);
    p = fmaf(p, z, -1.f / 3.f);
    p = fmaf(p, z, -0.5f);
    p = fmaf(p, z, -1.f);
    return p * z;
}

// nengo LIF.step on the packed state (App. A.4), branch-free.  Returns the output (0 or amplitude/dt).
//   m = min(s, 0) is minus the remaining refractory time, v = max(s, 0) the voltage (one of them is 0).
//   nengo: refractory_time -= dt; delta = clip(dt - refractory_time, 0, dt)  =>  delta/dt = clip(2 + m/dt, 0, 1);
//   the neuron stays refractory (state m + dt) exactly when that clip gives 0, i.e. refractory_time - dt >= dt.
template <bool FAST>
__device__ __forceinline__ float ssb_lif_packed(const SsbNeuron& n, float J, float& s) {
    const float m = fminf(s, 0.f);
    float v = fmaxf(s, 0.f);
    const float dn = __saturatef(fmaf(m, n.inv_dt, 2.f));  // delta / dt
    const float x = dn * n.neg_dt_over_tau;                // -delta / tau_rc
    const float em1 = FAST ? ssb_expm1_small(x) : expm1f(x);
    v = fmaf(v - J, em1, v);                               // v -= (J - v) * expm1(-delta / tau_rc)
    const bool spiked = v > 1.f;
    const float z = __fdividef(v - 1.f, J - 1.f);          // used only when spiked (then J > v > 1)
    const float lp = FAST ? ssb_log1p_neg_small(z) : log1pf(-z);
    const float r_new = fmaf(n.tau_rc, lp, n.c0);          // tau_ref + dt + tau_rc * log1p(-z) > 0
    const float keep = (dn > 0.f) ? fmaxf(v, 0.f) : m + n.dt;
    s = spiked ? -r_new : keep;
    return spiked ? n.amp_dt : 0.f;
}

// Two neurons per lane with sm_100's packed fp32 instructions (FFMA2 / FADD2 / FMUL2: one issue slot for two
// operations).  k_ens_small is issue-bound, and two thirds of the LIF update are fma / add / mul chains.
__device__ __forceinline__ float2 ssb_fma2(float2 a, float2 b, float2 c) {
    float2 r;
    asm("fma.rn.ftz.f32x2 %0, %1, %2, %3;"
        : "=l"(*reinterpret_cast<unsigned long long*>(&r))
        : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)),
          "l"(*reinterpret_cast<unsigned long long*>(&c)));
    return r;
}
__device__ __forceinline__ float2 ssb_mul2(float2 a, float2 b) {
    float2 r;
    asm("mul.rn.ftz.f32x2 %0, %1, %2;"
        : "=l"(*reinterpret_cast<unsigned long long*>(&r))
        : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
    return r;
}
__device__ __forceinline__ float2 ssb_add2(float2 a, float2 b) {
    float2 r;
    asm("add.rn.ftz.f32x2 %0, %1, %2;"
        : "=l"(*reinterpret_cast<unsigned long long*>(&r))
        : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
    return r;
}
__device__ __forceinline__ float2 ssb_splat(float x) { return make_float2(x, x); }

// ssb_lif_packed<true> for two neurons at once (same arithmetic, element-wise): s0/s1 states, J0/J1 currents.
__device__ __forceinline__ float2 ssb_lif_pair(const SsbNeuron& n, float2 J, float2& s) {
    const float2 m = make_float2(fminf(s.x, 0.f), fminf(s.y, 0.f));
    float2 v = make_float2(fmaxf(s.x, 0.f), fmaxf(s.y, 0.f));
    float2 dn = ssb_fma2(m, ssb_splat(n.inv_dt), ssb_splat(2.f));
    dn = make_float2(__saturatef(dn.x), __saturatef(dn.y));
    const float2 x = ssb_mul2(dn, ssb_splat(n.neg_dt_over_tau));
    float2 p = ssb_fma2(ssb_splat(1.f / 120.f), x, ssb_splat(1.f / 24.f));     // expm1(x), degree-5 Taylor
    p = ssb_fma2(p, x, ssb_splat(1.f / 6.f));
    p = ssb_fma2(p, x, ssb_splat(0.5f));
    p = ssb_fma2(p, x, ssb_splat(1.f));
    const float2 em1 = ssb_mul2(p, x);
    const float2 vmj = ssb_fma2(J, ssb_splat(-1.f), v);                        // v - J
    v = ssb_fma2(vmj, em1, v);
    const bool sp0 = v.x > 1.f, sp1 = v.y > 1.f;
    const float2 vm1 = ssb_add2(v, ssb_splat(-1.f)), jm1 = ssb_add2(J, ssb_splat(-1.f));
    float2 rc;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc.x) : "f"(jm1.x));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc.y) : "f"(jm1.y));
    const float2 z = ssb_mul2(vm1, rc);
    float2 q = ssb_fma2(ssb_splat(-1.f / 6.f), z, ssb_splat(-0.2f));            // log1p(-z), 6 terms
    q = ssb_fma2(q, z, ssb_splat(-0.25f));
    q = ssb_fma2(q, z, ssb_splat(-1.f / 3.f));
    q = ssb_fma2(q, z, ssb_splat(-0.5f));
    q = ssb_fma2(q, z, ssb_splat(-1.f));
    const float2 lp = ssb_mul2(q, z);
    const float2 r_new = ssb_fma2(ssb_splat(n.tau_rc), lp, ssb_splat(n.c0));
    const float2 mdt = ssb_add2(m, ssb_splat(n.dt));
    const float keep0 = (dn.x > 0.f) ? fmaxf(v.x, 0.f) : mdt.x;
    const float keep1 = (dn.y > 0.f) ? fmaxf(v.y, 0.f) : mdt.y;
    s.x = sp0 ? -r_new.x : keep0;
    s.y = sp1 ? -r_new.y : keep1;
    return make_float2(sp0 ? n.amp_dt : 0.f, sp1 ? n.amp_dt : 0.f);
}

__device__ __forceinline__ float ssb_rate(const SsbNeuron& n, float J) {
    if (n.type == 1) {
        const float j = J - 1.f;
        return j > 0.f ? n.amp / (n.tau_ref + n.tau_rc * log1pf(1.f / j)) : 0.f;
    }
    return n.amp * fmaxf(J, 0.f);
}

// MODE 0: LIF with polynomial transcendental functions; MODE 1: anything else (uniform run-time switch).
template <int MODE>
__device__ __forceinline__ float ssb_neuron_apply(const SsbNeuron& n, float J, float& s) {
    if (MODE == 0) return ssb_lif_packed<true>(n, J, s);
    if (n.type == 0) return ssb_lif_packed<false>(n, J, s);
    return ssb_rate(n, J);
}

// --------------------------------------------------------------------------------------
// TMA 1-D bulk copies + mbarriers (one elected lane issues; the warp waits on the barrier).
__device__ __forceinline__ uint32_t ssb_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void ssb_mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ssb_smem(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void ssb_mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ssb_smem(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ssb_mbar_wait(unsigned long long* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(ssb_smem(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void ssb_bulk_g2s(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     ssb_smem(dst)),
                 "l"(src), "r"(bytes), "r"(ssb_smem(bar))
                 : "memory");
}
__device__ __forceinline__ void ssb_bulk_s2g(void* dst, const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(ssb_smem(src)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void ssb_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void ssb_bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void ssb_bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void ssb_fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// --------------------------------------------------------------------------------------
// tcgen05 / TMEM helpers shared by the tensor-core kernels (grid scan, static decoders).
#define SSB_TC_ROWS 128

__host__ __device__ __forceinline__ float ssb_tf32_round(float x) {   // round-to-nearest-even to a 10-bit mantissa
#ifdef __CUDA_ARCH__
    uint32_t u = __float_as_uint(x);
#else
    uint32_t u;
    memcpy(&u, &x, 4);
#endif
    u += 0xfffu + ((u >> 13) & 1u);
    u &= 0xffffe000u;
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    float r;
    memcpy(&r, &u, 4);
    return r;
#endif
}

__device__ __forceinline__ uint64_t ssb_umma_desc(const void* smem_ptr) {
    const uint32_t a = ssb_smem(smem_ptr);
    return (uint64_t)((a >> 4) & 0x3fffu) | ((uint64_t)(2048u >> 4) << 16) | ((uint64_t)(128u >> 4) << 32) | (1ull << 46);
}

// same, with the stride between 16-byte K chunks given (= 128 B x row groups of the tile)
__device__ __forceinline__ uint64_t ssb_umma_desc_lbo(const void* smem_ptr, uint32_t lbo_bytes) {
    const uint32_t a = ssb_smem(smem_ptr);
    return (uint64_t)((a >> 4) & 0x3fffu) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(128u >> 4) << 32) | (1ull << 46);
}

__device__ __forceinline__ void ssb_umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                              uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ void ssb_tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    // the registers are valid only after wait::ld; tying them to the wait keeps every use behind it
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                   "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]),
                   "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]),
                   "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                 :
                 : "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void ssb_tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                   "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :
                 : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void ssb_tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void ssb_tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// --------------------------------------------------------------------------------------
// Deferred-PES history arenas (see ssb_pes.cuh) and the hand-over that lets the Voja ensemble kernel decode through the
// PES-learned decoders of its own neurons (k_wide_voja<.., true>): item[k] = PES descriptor fed by the k-th ensemble of the
// launch, or -1.
#define SSB_PES_KMAX 16
struct SsbPesDefer {
    float* hist_e;            // [G][rows_e][32]
    float* hist_f;            // [G][rows_f][32]
    float* part;              // [G][rows_p][32]
    int* counters;
    int rows_e, rows_f, rows_p, K;
};

struct SsbPesFuse {
    const int* desc;          // PES descriptors (13 ints each)
    const int* hdesc;         // e_row0 f_row0 part_row0 counter0 per PES descriptor
    SsbPesDefer h;
    int item[15];
};
