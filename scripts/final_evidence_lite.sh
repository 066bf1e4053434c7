#!/bin/bash
# final_evidence.sh without the K = 100 line and the saturated-LM capture (LM kernel unchanged since the last full run)
mkdir -p gpurun_out
python -m pytest tests -q -m gpu > gpurun_out/r2_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests.log
tail -3 gpurun_out/r2_tests.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err
python bench.py > gpurun_out/r2_bench_1.json 2> gpurun_out/r2_bench_1.err; echo "bench rc=$?"
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_k20.json 2> gpurun_out/r2_bench_k20.err
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-parity --no-extras --min-timed-s 0.001"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1000 -c 440 --csv --log-file gpurun_out/launches_r2.csv $B > gpurun_out/ncu_launches_r2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:"lm2_kernel|bm_tma_kernel|bm_kernel|fuse_fold|fold_order|fuse_stage_kernel|seeds_order|points_order|map_regularize|map_commit|map_cbits|ts_decay_median|ts_ingest|ts_remap|ts_scatter_fix" \
    -s 480 -c 19 -f -o gpurun_out/prof_r2 $B > gpurun_out/ncu_full_r2.log 2>&1
python - <<'PY'
import json
for f in ("r2_bench_1", "r2_bench_k20", "r2_bench_ref"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d.get("ms_per_step"), d.get("value"), (d.get("e2e") or {}).get("value"), (d.get("roofline") or {}).get("frac"))
    except Exception as e:
        print(f, "unreadable", e)
PY
