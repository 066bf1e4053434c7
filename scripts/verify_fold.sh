mkdir -p gpurun_out
python -m pytest tests/test_gpu_full_size.py tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/v_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/v_tests.log
for sd in 10 11 12 13; do python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-parity --no-extras --min-timed-s 0.1 --scene-seed $sd > gpurun_out/v_scene_$sd.json 2> gpurun_out/v_scene_$sd.err; done
tail -5 gpurun_out/v_tests.log; for sd in 10 11 12 13; do python -c "import json,sys; d=json.loads(open('gpurun_out/v_scene_$sd.json').read().strip().splitlines()[-1]); print($sd, d['ms_per_step'], d['value'], d['e2e']['value'])"; done
