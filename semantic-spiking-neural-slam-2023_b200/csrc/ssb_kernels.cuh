// sm_100a kernels of the SSP-SLAM step engine (device side), split by kernel family:
//   ssb_common.cuh     context, step helpers, neuron models, TMA / mbarrier and tcgen05 / TMEM helpers (layout rules are there)
//   ssb_inputs.cuh     k_begin, k_synth
//   ssb_ens_small.cuh  k_ens_small
//   ssb_ens_wide.cuh   k_wide_static, k_wide_static_tc, k_wide_voja
//   ssb_decode.cuh     k_decode, k_decode_tc
//   ssb_pes.cuh        k_pes, k_pes_hist, k_pes_defer, k_pes_fold, k_pes_clear
//   ssb_cleanup.cuh    k_cleanup_scan, k_cleanup_scan_tc, k_scan_xtiles, k_cleanup_scan_tck, k_cleanup_pick, k_gate
//   ssb_ens_wide_cta.cuh  k_wide_voja_cta (d = 649: the CTA works on one neuron's encoder tile at a time)
//   ssb_ens_wide_tck.cuh  k_wide_static_tck
//   ssb_lin.cuh        k_lin, k_advance
//   ssb_lin_tck.cuh    k_lin_xtiles, k_lin_tck (large dense blocks of the row program on tcgen05)
//   ssb_ssp.cuh        k_ssp_encode, k_decode_prep
#pragma once
#include "ssb_common.cuh"
#include "ssb_inputs.cuh"
#include "ssb_ens_small.cuh"
#include "ssb_ens_wide.cuh"
#include "ssb_ens_wide_cta.cuh"
#include "ssb_decode.cuh"
#include "ssb_pes.cuh"
#include "ssb_cleanup.cuh"
#include "ssb_ens_wide_tck.cuh"
#include "ssb_lin.cuh"
#include "ssb_lin_tck.cuh"
#include "ssb_ssp.cuh"
