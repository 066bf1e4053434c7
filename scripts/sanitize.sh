#!/bin/bash
# compute-sanitizer memcheck over a small slice of the GPU suite (new kernels: TMA block matching, pooled fold sort, event matcher)
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 \
  python -m pytest "tests/test_gpu_parity.py::test_block_matching_parity" "tests/test_gpu_parity.py::test_cull_and_fusion_parity" \
  "tests/test_gpu_parity.py::test_mapping_at_time_multi_frame" "tests/test_gpu_mvstereo.py::test_event_matcher_parity" \
  "tests/test_gpu_mvstereo.py::test_vemp2vdp_and_naive_accumulation_parity" -x -q -m gpu > gpurun_out/sanitize.log 2>&1
echo "sanitizer rc=$?" >> gpurun_out/sanitize.log
grep -c "Invalid\|Error" gpurun_out/sanitize.log; tail -15 gpurun_out/sanitize.log
