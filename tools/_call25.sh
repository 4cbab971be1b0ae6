set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "pipelined or dropin or fullsize_ref or config2 or substep or slab" > gpurun_out/c25_tests.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/c25_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-secondary --no-tolerance > gpurun_out/c25_bench512.json 2> gpurun_out/c25_bench512.err
FFB200_P2G_PRIORITY=0 FFB200_PIPELINED_UPLOAD=0 timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-secondary --no-tolerance --no-checksum > gpurun_out/c25_bench512_off.json 2> gpurun_out/c25_bench512_off.err
python - <<'PY'
import json
for f in ["c25_bench512","c25_bench512_off"]:
    try:
        j=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, j["ms_per_step"], j["value"], j["roofline"]["frac"], (j.get("e2e") or {}).get("ms_per_step"), (j.get("e2e") or {}).get("value"))
        print("  stages", {k[:12]: round(v["ms"],2) for k,v in j["roofline"].get("stages").items()})
    except Exception as e:
        print(f, "ERR", e)
PY
