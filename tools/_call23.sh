set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/c23_tests.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/c23_tests.log
timeout 600 python bench.py --grid 128 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/c23_bench128.json 2> gpurun_out/c23_bench128.err
FFB200_P2G_MERGED=0 timeout 600 python bench.py --grid 128 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-secondary > gpurun_out/c23_bench128_unmerged.json 2> gpurun_out/c23_bench128_unmerged.err
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/c23_bench512.json 2> gpurun_out/c23_bench512.err
FFB200_P2G_MERGED=0 timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-secondary --no-tolerance > gpurun_out/c23_bench512_unmerged.json 2> gpurun_out/c23_bench512_unmerged.err
python - <<'PY'
import json
for f in ["c23_bench128","c23_bench128_unmerged","c23_bench512","c23_bench512_unmerged"]:
    try:
        j=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, j["ms_per_step"], j["value"], j["roofline"]["frac"], j.get("checksum"), (j.get("tolerance_mode") or {}).get("ms_per_step"), (j.get("e2e") or {}).get("value"))
        print("  stages", j["roofline"].get("stages"))
    except Exception as e:
        print(f, "ERR", e)
PY
