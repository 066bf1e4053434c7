#!/bin/bash
# Evidence run on the GPU box (round 2): launch list and `ncu --set full` captures at HEAD.
# Artefacts land in gpurun_out/; scripts/summarize_ncu.py + scripts/ncu_lm_counters.py turn them into profiles/r2_* here.
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-parity --no-extras --min-timed-s 0.001"
# (1) launch list: one timed region (20 steps x 22 launches) behind priming (20) + profiled pass (20) + warm-up (5) steps
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1000 -c 440 --csv --log-file gpurun_out/launches_r2.csv $B > gpurun_out/ncu_launches_r2.log 2>&1
# (2) full capture of one frame's kernels (~16 matching launches per frame; skip ~30 frames)
timeout 1200 ncu --set full --clock-control none --import-source on \
    -k regex:"lm2_kernel|bm_tma_kernel|bm_kernel|fuse_fold|fuse_stage_kernel|seeds_order|points_order|map_regularize|map_commit|map_cbits|ts_decay_median|ts_ingest|ts_remap|ts_scatter_fix" \
    -s 480 -c 18 -f -o gpurun_out/prof_r2 $B > gpurun_out/ncu_full_r2.log 2>&1
tail -2 gpurun_out/ncu_full_r2.log
# (3) the LM kernel saturated (16 frames' seeds in one launch: no tail), same kernel as (2)
timeout 400 ncu --set full --clock-control none --import-source on -k regex:lm2_kernel -s 5 -c 1 -f -o gpurun_out/lm_sat_r2 python scripts/lm_saturation.py --child 12011 > gpurun_out/ncu_lm_sat_r2.log 2>&1
tail -2 gpurun_out/ncu_lm_sat_r2.log
