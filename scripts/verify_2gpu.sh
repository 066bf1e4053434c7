mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/v2_gpus.txt
python bench.py > gpurun_out/r2_bench_1.json 2> gpurun_out/r2_bench_1.err; echo "bench1 rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err; echo "bench2 rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_1.json").read().strip().splitlines()[-1])
print("1gpu", d["ms_per_step"], d["value"], d["e2e"]["value"], d["extras"].get("sgm_init"))
d = json.loads(open("gpurun_out/r2_bench_2gpu.json").read().strip().splitlines()[-1])
print("2gpu", d["ms_per_step"], d["value"], d["e2e"]["value"], {k: (v.get("value"), v.get("ms_per_step")) for k, v in d.get("extras", {}).items() if isinstance(v, dict)})
PY
