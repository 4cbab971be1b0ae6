set -u
mkdir -p gpurun_out
i=0
for cfg in "0 16" "1 16"; do
set -- $cfg
i=$((i+1))
FFB200_BENCH_OVERLAP=$1 NCCL_MIN_P2P_NCHANNELS=$2 NCCL_MAX_P2P_NCHANNELS=32 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2954$i bench.py --gpus 8 --steps 20 --warmup 3 --no-e2e --no-tolerance > gpurun_out/c20_ov$1_ch$2.json 2> gpurun_out/c20_ov$1_ch$2.err
python - <<PY
import json
d = json.load(open("gpurun_out/c20_ov$1_ch$2.json"))
print("overlap $1 channels $2:", d["value"] / 1e9, d["ms_per_step"], d["checksum"]["particle_hash"])
PY
done
