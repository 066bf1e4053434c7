mkdir -p gpurun_out
python -m pytest tests -q -m gpu > gpurun_out/v_all.log 2>&1; echo "tests rc=$?" >> gpurun_out/v_all.log
tail -8 gpurun_out/v_all.log
bash scripts/sanitize.sh
