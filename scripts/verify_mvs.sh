mkdir -p gpurun_out
python -m pytest tests/test_gpu_mvstereo.py tests/test_cpp_shim.py -q -m gpu -s > gpurun_out/v_mvs.log 2>&1; echo "tests rc=$?" >> gpurun_out/v_mvs.log
grep -v "^$" gpurun_out/v_mvs.log | tail -60
