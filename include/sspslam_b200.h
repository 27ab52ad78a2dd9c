/* sspslam_b200 — C-ABI of the B200 (sm_100a) SSP-SLAM step engine.
 *
 * The reference has no FFI: its boundary for this path is nengo's Python backend
 * convention `sim = Simulator(network); with sim: sim.run(T); sim.data[probe]`
 * (/root/reference/experiments/run_slam.py:198-233,243,250; run_pathint.py:147-163;
 * run_slamview.py:148-158) plus the NumPy SSP API `SSPSpace.encode/decode`
 * (/root/reference/sspslam/sspspace.py:252-273,312-358).  The Python `Simulator`
 * in this repo lowers the network into flat arrays and drives the entry points below
 * through ctypes; INTEGRATION.md shows the binding.  Plain pointers and sizes only.
 *
 * Conventions: every function returns 0 on success and a negative code on failure;
 * `ssb_last_error()` gives the message (thread-local).  The caller owns all host
 * buffers; the library owns device memory.  One handle = one GPU + one stream; a
 * handle is not thread-safe.  Trials are padded to a multiple of 32; host-side per-trial
 * arrays are [row][trial] with `trial` contiguous.
 */
#ifndef SSPSLAM_B200_H
#define SSPSLAM_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ssb_sim ssb_sim;

/* Replaces: nengo.Simulator(model) construction (run_slam.py:199).  `n_trials` is the
 * batching extension: independent trials sharing the static weights. */
int ssb_create(int device, int n_trials, ssb_sim** out);

/* Plan upload (output of the Python lowering).  Array names: csr_ptr csr_ent0 csr_ent1
 * weights ens_small ens_big dec pes cleanup gate lin_rows lin_ab stages ntypes
 * cleanup_s64.  Scalar names: dt nv nf nt tab_row0 nn n_act n_lenc n_ldec n_afilt n_probe
 * n_levels chunk_cap n_part n_jtiles. */
int ssb_set_array(ssb_sim* s, const char* name, const void* data, size_t bytes);
int ssb_set_scalar(ssb_sim* s, const char* name, double value);
/* Allocates the per-trial arenas (zero-filled) and uploads the plan. */
int ssb_finalize(ssb_sim* s);

/* Arena access, host rows are [n_rows][n_trials_padded] float32 (the device keeps them tiled
 * [trial_group][row][32]; the library converts).  Arena names:
 * "st" (packed LIF state: s >= 0 voltage, s < 0 minus the remaining refractory time),
 * "lenc" (Voja-learned scaled encoders, row = n*dims+k),
 * "ldec" (PES-learned decoders, row = j*n_pre+i), "afilt" (PES pre-synaptic trace),
 * "vec" (filter states and scratch), "act" (last activities of the wide ensembles),
 * "cidx" (int32 bits: last clean-up argmax per clean-up node).
 * Replaces: sim.data[Probe(conn,"weights")] / initial-state seeding (run_slam.py:195,265). */
int ssb_upload(ssb_sim* s, const char* arena, size_t row0, size_t n_rows, const float* host);
int ssb_download(ssb_sim* s, const char* arena, size_t row0, size_t n_rows, float* host);

/* Input tables for steps [step0, step0+n_steps): host layout [n_steps][nt][n_trials_padded].
 * Replaces the per-step Python input nodes (run_slam.py:164-169; slam.py:451-495). */
int ssb_set_tables(ssb_sim* s, const float* host, long long step0, int n_steps);
/* Re-base the resident tables so that they apply from `step0` on (bench: reuse a chunk). */
int ssb_rebase_tables(ssb_sim* s, long long step0);

/* Replaces: sim.run_steps(n) / sim.run(T) (run_slam.py:232-233).  Asynchronous. */
int ssb_run_steps(ssb_sim* s, int n_steps);
/* ssb_set_tables + ssb_run_steps + ssb_read_probes in one call, software-pipelined over 16-step sub-chunks: tables of
 * the next sub-chunk are copied host->device and probe rows of the previous one device->host while the current one
 * computes.  Both host buffers should be page-locked (ssb_host_alloc).  host_tables [n_steps][nt][n_trials_padded] for
 * the steps following ssb_n_steps(); host_probes [n_steps][n_probe][n_trials_padded] (may be NULL).  Asynchronous:
 * the host buffers belong to the library until ssb_io_wait() returns.
 * Replaces: sim.run_steps(n) + sim.data[probe] with host-resident inputs (run_slam.py:232-233,250). */
int ssb_run_steps_io(ssb_sim* s, const float* host_tables, int n_steps, float* host_probes);
int ssb_io_wait(ssb_sim* s);
/* On-device input synthesis.  Replaces the input tables, i.e. the reference's per-step Python input
 * closures (sspslam/networks/slam.py:442-497 get_slam_input_functions2; experiments/run_slam.py:164-169;
 * run_pathint.py:134-136), by a kernel that evaluates them from per-trial paths and landmarks:
 *   cfg     = {domain_dim, ssp_dim, n_landmarks, path_len, vel_col, init_col, lmvec_col, lmsp_col, nolm_col}
 *             (column = first input-table row of that signal, -1 = absent),
 *   fparams = {view_rad, none_in_view_value},
 *   phases  [ssp_dim][domain_dim] = phase_matrix / length_scale, lm_sp [n_landmarks][ssp_dim] (float64, shared),
 *   path_rows / vel_rows [path_len*domain_dim][n_trials_padded], lm_rows [n_landmarks*domain_dim][n_trials_padded]
 *             (float32, per trial; vel = velocities already multiplied by vel_scaling_factor).
 * Call after ssb_finalize.  ssb_synth_steps then supplies, per chunk, the float-fragile step indices the closures use
 * (idx [n_steps][3] = int((t-dt)/dt), min(floor(t/dt), path_len-2), t < init_time) instead of table rows. */
int ssb_synth_setup(ssb_sim* s, const int* cfg, const float* fparams, const double* phases, const double* lm_sp,
                    const float* path_rows, const float* vel_rows, const float* lm_rows);
int ssb_synth_steps(ssb_sim* s, const int* idx, long long step0, int n_steps);
/* Replaces: sim.data[probe] for node/ensemble probes: host [n_steps][n_probe][n_trials_padded]. */
int ssb_read_probes(ssb_sim* s, float* host, long long step0, int n_steps);
long long ssb_n_steps(ssb_sim* s);
int ssb_n_trials_padded(ssb_sim* s);
int ssb_sync(ssb_sim* s);
/* Replaces: sim.reset() — zero state arenas and the step counter (weights are re-uploaded by the host). */
int ssb_reset(ssb_sim* s);
void ssb_destroy(ssb_sim* s);

/* Measurement support: CUDA-event time (ms) of the last ssb_run_steps call on the
 * library stream; with profiling on, per-kernel-kind accumulated event times. */
int ssb_set_profiling(ssb_sim* s, int on);
int ssb_last_run_ms(ssb_sim* s, float* ms);
/* kinds: 0 ens_small 1 ens_wide 2 decode 3 pes 4 cleanup_scan 5 cleanup_pick 6 gate 7 lin 8 advance 9 begin
 *        10 ens_voja */
int ssb_kernel_times(ssb_sim* s, float* ms_per_kind, long long* launches_per_kind, int n_kinds);
long long ssb_total_launches(ssb_sim* s);
/* Timeline of the launches since ssb_set_profiling(s, 2) ("timeline" profiling: per-launch CUDA events recorded on the
 * dependency streams the launches really use, so concurrency between the streams is preserved): start / end in ms since
 * that call and the kernel kind of every launch, in host launch order.  *n_out = launches available. */
int ssb_timeline(ssb_sim* s, float* start_ms, float* end_ms, int* kinds, int max_n, int* n_out);
/* CUDA-event marks on the library stream (4 slots) and the device time between two of them:
 * brackets a timed region that spans several ssb_run_steps / table / probe calls. */
int ssb_mark(ssb_sim* s, int slot);
int ssb_mark_elapsed_ms(ssb_sim* s, int slot_a, int slot_b, float* ms);

/* Stand-alone SSP kernels.
 * Replaces SSPSpace.encode (sspspace.py:252-273): out[N][d] = IFFT(exp(i A_scaled x)).real,
 * A_scaled = phase_matrix / length_scale, [d][n] row-major; x [N][n]. */
int ssb_ssp_encode(int device, const double* a_scaled, const double* x, double* out,
                   long long n_points, int domain_dim, int ssp_dim);
/* Replaces SSPSpace.decode(...,'from-set') core (sspspace.py:349-358): per query row,
 * normalise (skip if norm < 1e-6) and return argmax_g S[g].u, first maximum wins. */
int ssb_ssp_decode_argmax(int device, const double* sample_ssps, const double* queries, int* idx_out,
                          long long n_queries, long long n_samples, int ssp_dim);

/* Pinned host staging buffers for tables / probes (cudaHostAlloc). */
void* ssb_host_alloc(size_t bytes);
void ssb_host_free(void* p);

const char* ssb_last_error(void);
const char* ssb_version(void);

#ifdef __cplusplus
}
#endif
#endif /* SSPSLAM_B200_H */
