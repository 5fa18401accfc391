#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_ivf_gpu.py -m gpu -q -x -k "ring_scan or coresident or two_stage_scan_matches or two_stage_results_through_peer or merge_shards_packed or flat_search_ragged or unit_norm_generator or spherical" > gpurun_out/r02p_sanitizer_ivf.log 2>&1; echo "rc=$?" >> gpurun_out/r02p_sanitizer_ivf.log
tail -6 gpurun_out/r02p_sanitizer_ivf.log
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_encoder_gpu.py -m gpu -q -x -k "attention_kernels_all_lengths and (257 or 300 or 512 or 511 or 33 or 256) or fused_epilogues or tiny_encoder" > gpurun_out/r02p_sanitizer_enc.log 2>&1; echo "rc=$?" >> gpurun_out/r02p_sanitizer_enc.log
tail -6 gpurun_out/r02p_sanitizer_enc.log
