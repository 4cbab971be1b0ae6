set -u
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/c9_gpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29516 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/c9_bench_n8.json 2> gpurun_out/c9_bench_n8.err
tail -c 400 gpurun_out/c9_bench_n8.json; tail -5 gpurun_out/c9_bench_n8.err
