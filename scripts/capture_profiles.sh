#!/bin/bash
# Round-end evidence run on the GPU box: default bench, reference arm, launch list and one full ncu capture.
# Artefacts land in gpurun_out/; scripts/summarize_ncu.py turns them into profiles/<tag>_*.md here.
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -2 gpurun_out/bench_final.err
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_final_ref.json 2>> gpurun_out/bench_final.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1100 -c 400 --csv --log-file gpurun_out/launches_final.csv \
    python bench.py --steps 12 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:"lm_kernel|bm_kernel|fuse_fold_kernel|ts_decay_median|fuse_stage|ts_ingest|map_regularize_kernel|seeds_order|points_order" \
    -s 385 -c 11 -f -o gpurun_out/prof_final python bench.py --steps 12 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
