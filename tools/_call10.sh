set -u
mkdir -p gpurun_out
FFB200_SLAB_PROFILE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 10 --warmup 3 --no-e2e --no-tolerance --no-checksum > gpurun_out/c10_prof_n8.json 2> gpurun_out/c10_prof_n8.err
grep "slab phases" gpurun_out/c10_prof_n8.err
