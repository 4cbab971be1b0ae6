set -u
mkdir -p gpurun_out
for v in default tune_c8 tune_c10 tune_f5; do
  if [ $v = default ]; then unset FFB200_LIBRARY; else export FFB200_LIBRARY=$PWD/blender_flip_fluids_b200/lib/$v.so; fi
  timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-secondary --no-checksum > gpurun_out/c28_$v.json 2> gpurun_out/c28_$v.err
  python - <<PY
import json
j=json.loads(open("gpurun_out/c28_$v.json").read().strip().splitlines()[-1])
t=j["tolerance_mode"]
print("$v", j["ms_per_step"], {k[:6]: round(x["ms"],2) for k,x in j["roofline"]["stages"].items()}, "tol", t["ms_per_step"], {k[:6]: round(x,2) for k,x in t["stage_ms"].items()})
PY
done
