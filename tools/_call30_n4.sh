set -u
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29584 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/c30_bench_n4.json 2> gpurun_out/c30_bench_n4.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/c30_bench_n4.json").read().strip().splitlines()[-1])
print("N=4", d["value"] / 1e9, d["ms_per_step"], d["checksum"]["particle_hash"], d["checksum"]["p2g_field_hash"], d["config"].get("exchange_repeats"), (d.get("e2e") or {}).get("value"), (d.get("tolerance_mode") or {}).get("ms_per_step"))
PY
tail -3 gpurun_out/c30_bench_n4.err
