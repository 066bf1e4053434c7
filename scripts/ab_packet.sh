mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -x -q -m gpu > gpurun_out/ab_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/ab_tests.log; tail -3 gpurun_out/ab_tests.log
B="python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-parity --no-extras --min-timed-s 0.3"
for v in 0 1 0 1; do ESVO_TS_PACKET_COPY=$v $B > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err; python -c "
import json; d=json.loads(open('gpurun_out/ab_$v.json').read().strip().splitlines()[-1]); print('packet=$v', d['ms_per_step'], d['e2e']['ms_per_step'], d['host_issue_ms'], d['streams'][0]['map_checksum'])"; done
