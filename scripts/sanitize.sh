#!/bin/bash
# compute-sanitizer memcheck over one small parity case of every kernel family (the new round-2 kernels first).
# Usage: scripts/sanitize.sh [tag]
tag=${1:-run}
mkdir -p gpurun_out
run() {  # name, pytest -k expression, file
  ( timeout 900 compute-sanitizer --tool memcheck --error-exitcode 99 python -m pytest "$3" -m gpu -q -x -k "$2" ) > gpurun_out/sanitize_${tag}_$1.log 2>&1
  echo "$1: rc=$? $(grep -E 'ERROR SUMMARY|passed|failed' gpurun_out/sanitize_${tag}_$1.log | tr '\n' ' ')"
}
run iso_ws "test_3d_iso_matches_oracle and shape0" tests/test_gpu_parity.py
run iso_xtiles "test_3d_interior_x_tiles and 104" tests/test_gpu_parity.py
run iso_f32 "single_restatement and shape1" tests/test_gpu_f32.py
run visco_ws "test_visco_ws_velocity_tiles and 2-104" tests/test_gpu_visco.py
run twod_ws "test_2d_ws_strips_and_chunks and 3-3-4" tests/test_gpu_parity.py
run multi_one_device "test_multi_handle_isotropic_matches_oracle and 2-False" tests/test_gpu_multi.py
run multi_visco_one_device "test_multi_handle_viscoelastic_matches_oracle and 2-False-4" tests/test_gpu_multi.py
