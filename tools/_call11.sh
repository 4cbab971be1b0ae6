set -u
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/c11_bench_n1.json 2> gpurun_out/c11_bench_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/c11_bench_n2.json 2> gpurun_out/c11_bench_n2.err
python - <<'PY'
import json
for n in (1, 2):
    d = json.load(open(f"gpurun_out/c11_bench_n{n}.json"))
    print(n, d["value"] / 1e9, d["ms_per_step"], d["checksum"]["particle_hash"], d["checksum"]["p2g_field_hash"], d["tolerance_mode"]["value"] / 1e9, d["tolerance_mode"]["exact_fallback_rate"], d["e2e"]["value"] / 1e9)
PY
tail -3 gpurun_out/c11_bench_n1.err gpurun_out/c11_bench_n2.err
