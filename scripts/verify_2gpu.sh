mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/v2_gpus.txt
python -m pytest tests/test_cpp_shim.py -q -m gpu -s -k "nccl" > gpurun_out/v2_nccl.log 2>&1; echo "nccl test rc=$?" >> gpurun_out/v2_nccl.log
tail -6 gpurun_out/v2_nccl.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err; echo "bench2 rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_2gpu.json").read().strip().splitlines()[-1])
print("2gpu", d["ms_per_step"], d["value"], d["e2e"]["value"], {k: (v.get("value"), v.get("ms_per_step")) for k, v in d.get("extras", {}).items() if isinstance(v, dict)})
PY
