set -u
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/c8_gpus.txt
timeout 900 python -m pytest tests -m gpu -q -k "slab or tolerance or feature_scenes" 2>&1 | tail -15 > gpurun_out/c8_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/c8_bench_n2.json 2> gpurun_out/c8_bench_n2.err
tail -5 gpurun_out/c8_pytest.log; tail -c 600 gpurun_out/c8_bench_n2.json; tail -5 gpurun_out/c8_bench_n2.err
