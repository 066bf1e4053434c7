mkdir -p gpurun_out
( time python bench.py > gpurun_out/r2_bench_1.json 2> gpurun_out/r2_bench_1.err ) 2> gpurun_out/r2_bench_1.time; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err
cat gpurun_out/r2_bench_1.time | tail -4
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_1.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["e2e"]["value"], d["timing"]["timed_s_total"], d["timing"]["e2e_timed_s_total"], d["roofline"]["frac"])
PY
