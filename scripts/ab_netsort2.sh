mkdir -p gpurun_out
ESVO_FOLD_NETSORT=2 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py tests/test_gpu_mvstereo.py -x -q -m gpu > gpurun_out/ab_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/ab_tests.log; tail -3 gpurun_out/ab_tests.log
for v in 1 2; do ESVO_FOLD_NETSORT=$v python scripts/fold_probe.py cfg2 2>&1 | tail -1; done
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-parity --no-extras --min-timed-s 0.3"
for v in 1 2 1 2; do ESVO_FOLD_NETSORT=$v $B > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err; python -c "
import json; d=json.loads(open('gpurun_out/ab_$v.json').read().strip().splitlines()[-1]); print('netsort=$v K=20', d['ms_per_step'], d['e2e']['ms_per_step'], d['streams'][0]['map_checksum'])"; done
