mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py tests/test_gpu_mvstereo.py -x -q -m gpu > gpurun_out/ab_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/ab_tests.log; tail -3 gpurun_out/ab_tests.log
for c in cfg2 cfg3; do for v in 0 1; do ESVO_FOLD_HYBRID=$v python scripts/fold_probe.py $c 2>&1 | tail -1; done; done
B="python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-parity --min-timed-s 0.3"
for v in 0 1 0 1; do ESVO_FOLD_HYBRID=$v $B > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err; python -c "
import json; d=json.loads(open('gpurun_out/ab_$v.json').read().strip().splitlines()[-1]); c3=d['extras'].get('cfg3',{}); print('hybrid=$v', d['ms_per_step'], d['e2e']['ms_per_step'], d['streams'][0]['map_checksum'], 'cfg3', c3.get('ms_per_step'), c3.get('e2e',{}).get('ms_per_step'))"; done
