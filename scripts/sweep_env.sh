#!/bin/bash
# usage: sweep_env.sh "<tag> VAR=val ..." ...   -- one short bench run per argument with the given environment
for spec in "$@"; do
  set -- $spec; tag=$1; shift
  env "$@" python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-parity --no-extras --min-timed-s 0.25 > gpurun_out/b_x.json 2> gpurun_out/b_x.err
  python - "$tag" <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/b_x.json").read().strip().splitlines()[-1])
    print("%s: ms/step %.4f e2e %.4f" % (sys.argv[1], d["ms_per_step"], d["e2e"]["ms_per_step"]), {k: round(v, 3) for k, v in d["breakdown_ms_per_step"].items() if k != "source"})
except Exception as e:
    print(sys.argv[1], "failed", e, open("gpurun_out/b_x.err").read()[-300:])
PY
done
