set -u
mkdir -p gpurun_out
for ov in 1; do
FFB200_BENCH_OVERLAP=$ov timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2953$ov bench.py --gpus 8 --steps 20 --warmup 3 --no-e2e --no-tolerance > gpurun_out/c17_bench_n8_ov$ov.json 2> gpurun_out/c17_bench_n8_ov$ov.err
done
python - <<'PY'
import json
for ov in (1,):
    d = json.load(open(f"gpurun_out/c17_bench_n8_ov{ov}.json"))
    print("overlap", ov, d["value"] / 1e9, d["ms_per_step"], d["checksum"]["particle_hash"], d["checksum"]["p2g_field_hash"], d["config"]["exchange_repeats"], d["roofline"]["stage_ms"])
PY
