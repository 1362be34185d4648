"""dspfe — Python host layer over libdspfe.so (hand-written sm_100a CUDA kernels).

PyTorch supplies device memory and streams only; every feature value is computed by the
kernels in csrc/.  There is no CPU fallback: if libdspfe.so is missing or no CUDA device is
present, the compute entry points raise.
"""
from .binding import (DspfeError, EndpointPlan, MfccPlan, endpoint_decide_host, endpoint_params, frame_counts,  # noqa: F401
                      lib, lib_path, mfcc_params, mfcc_tables_host, num_frames, frames_f64, preemphasis_f64,
                      row_amplitude_f64, row_zcr_f64, delta_f32, amplitude_rule_host, zcr_rule_host, PitchPlan,
                      center_clip_f32, smooth_rows_f32, robust_max_pitch_host, smooth_subsequence_host,
                      sub_endpoint_host, pitch_feature_tail_host, poly_lead_host, dp_max_pitch_host, fir_window_f64, acr_f64, cmvn_pad_batch,
                      amplitude_rule_gated_host, acr_gate_rows_f64, wav_info, wav_scan_paths, ingest_wavs, row_windowed_amplitude_f64, pitch_num_frames_host,
                      FrontendPlan, timing_begin, timing_end, launch_count, dp_max_pitch_f64)
