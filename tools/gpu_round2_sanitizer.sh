#!/usr/bin/env bash
# compute-sanitizer memcheck over the kernels added in round 2: in-place streaming (even / odd steps, fused
# walls), task table, time-dependent boundary values on the device, plugin path, PDL launches
mkdir -p gpurun_out
export PYLBM_B200_HALO_TIMEOUT_S=120
timeout 1100 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 \
    python -m pytest -q -x -p no:cacheprovider \
    "tests/test_gpu_aa.py::test_in_place_streaming_is_bit_identical_to_two_arrays" \
    "tests/test_gpu_aa.py::test_in_place_streaming_with_and_without_the_fused_walls" \
    "tests/test_gpu_parity.py::test_boundary_entries_in_the_fused_kernel_are_bit_identical_to_the_list_kernels" \
    "tests/test_gpu_parity.py::test_time_dependent_boundary_values_on_the_device" \
    "tests/test_gpu_parity.py::test_walls_in_the_fused_kernel_are_bit_identical_to_the_list_kernel" \
    "tests/test_gpu_plugin.py::test_pylbm_simulation_cuda_against_reference_fixture" \
    > gpurun_out/sanitizer_memcheck.log 2>&1
echo "exit code $?" >> gpurun_out/sanitizer_memcheck.log
grep -E "ERROR SUMMARY|passed|failed|exit code|Invalid|out of bounds" gpurun_out/sanitizer_memcheck.log | tail -12
