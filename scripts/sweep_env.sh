#!/bin/bash
# usage: sweep_env.sh VAR "v1 v2 ..." [repeats] -- runs bench.py (no CPU baseline) once per value of an environment
# variable (e.g. ESVO_BENCH_SAMPLER "nvml smi"), prints resident / e2e ms per step and host issue statistics
VAR=$1; VALS=$2; REP=${3:-1}
for v in $VALS; do for r in $(seq $REP); do
  env $VAR=$v timeout 300 python bench.py --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); b=d['breakdown_ms_per_step']; h=d['host_issue_ms']
print('$VAR=$v', 'resident %.3f ms (%.0fM/s)  e2e %.3f ms (%.0fM/s)  lm %.2f fus %.2f  host res %.3f/%.2f/%.1f e2e %.3f/%.2f/%.1f  clk %s n=%s' % (d['ms_per_step'], d['value']/1e6, d['e2e']['ms_per_step'], d['e2e']['value']/1e6, b['depth_lm'], b['fusion_clean_regularise'], h['resident']['mean'],h['resident']['p99'],h['resident']['max'],h['e2e']['mean'],h['e2e']['p99'],h['e2e']['max'], d['clocks']['sm_mhz'], d['clocks']['samples']))"
done; done
