set -u
mkdir -p gpurun_out
for v in tune_a4_g4 tune_a5_g5 tune_a6_g6; do
  if [ $v = default ]; then unset FFB200_LIBRARY; else export FFB200_LIBRARY=$PWD/blender_flip_fluids_b200/lib/$v.so; fi
  timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-secondary --no-checksum --no-tolerance > gpurun_out/c27_$v.json 2> gpurun_out/c27_$v.err
  python - <<PY
import json
j=json.loads(open("gpurun_out/c27_$v.json").read().strip().splitlines()[-1])
print("$v", j["ms_per_step"], {k[:6]: round(x["ms"],2) for k,x in j["roofline"]["stages"].items()})
PY
done
