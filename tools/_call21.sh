set -u
mkdir -p gpurun_out
python tools/p2p_probe.py > gpurun_out/c21_p2p.txt 2>&1
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,P2P,SHM python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 tools/p2p_probe.py > gpurun_out/c21_nccl.txt 2>&1
cat gpurun_out/c21_p2p.txt; grep -i "via \|sendrecv\|P2P\b.*disab\|SHM" gpurun_out/c21_nccl.txt | head -20
nvidia-smi topo -m > gpurun_out/c21_topo.txt 2>&1; head -12 gpurun_out/c21_topo.txt
