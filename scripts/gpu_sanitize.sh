#!/bin/bash
# compute-sanitizer memcheck over small parity cases of every kernel family (3-D TMA, 3-D register, 2-D, viscoelastic 3-D / 2-D)
mkdir -p gpurun_out
CS="compute-sanitizer --tool memcheck --error-exitcode 86 --print-limit 20"
( time timeout 1200 $CS python -m pytest -x -q tests/test_gpu_parity.py -k "test_3d_iso_matches_oracle or test_3d_iso_kmax_pml or test_3d_register_kernels_still_match or test_2d_layered or test_3d_slabs_with_peer_stores" ) > gpurun_out/sanitize_iso.log 2>&1; echo "rc=$?" >> gpurun_out/sanitize_iso.log
( time timeout 1200 $CS python -m pytest -x -q tests/test_gpu_visco.py tests/test_gpu_visco2d.py -k "test_visco_matches_oracle or test_visco_slabs_in_one_process or test_visco2d_matches_oracle" ) > gpurun_out/sanitize_visco.log 2>&1; echo "rc=$?" >> gpurun_out/sanitize_visco.log
