set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_slab.py -m gpu -q 2>&1 | tail -8 > gpurun_out/c16_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e --no-tolerance > gpurun_out/c16_bench_n2.json 2> gpurun_out/c16_bench_n2.err
tail -3 gpurun_out/c16_pytest.log
python - <<'PY'
import json
d = json.load(open("gpurun_out/c16_bench_n2.json"))
print(2, d["value"] / 1e9, d["ms_per_step"], d["checksum"]["particle_hash"], d["checksum"]["p2g_field_hash"], d["config"]["exchange_repeats"])
PY
grep -i "error" gpurun_out/c16_bench_n2.err | head -5
