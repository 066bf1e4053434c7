#!/bin/bash
# Round-end evidence on one B200 (run through gpurun): GPU tests, both bench arms, ncu captures.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
python -m pytest tests -q -m gpu > gpurun_out/r2_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests.log
tail -3 gpurun_out/r2_tests.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err
python bench.py > gpurun_out/r2_bench_1.json 2> gpurun_out/r2_bench_1.err; echo "bench rc=$?"
python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-parity --no-extras > gpurun_out/r2_bench_k100.json 2> gpurun_out/r2_bench_k100.err
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_k20.json 2> gpurun_out/r2_bench_k20.err
bash scripts/capture_profiles.sh
python - <<'PY'
import json
for f in ("r2_bench_1", "r2_bench_k20", "r2_bench_k100", "r2_bench_ref"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d.get("ms_per_step"), d.get("value"), (d.get("e2e") or {}).get("value"), (d.get("roofline") or {}).get("frac"))
    except Exception as e:
        print(f, "unreadable", e)
PY
