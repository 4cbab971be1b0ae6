set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/c1_smi.txt; free -g >> gpurun_out/c1_smi.txt; nproc >> gpurun_out/c1_smi.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/c1_pytest.log
bash tools/run_experimental.sh > gpurun_out/c1_experimental.log 2>&1
timeout 120 python tools/try_scene.py 128 20 fixed > gpurun_out/c1_try128_fixed.log 2>&1
timeout 120 python tools/try_scene.py 128 20 evolving > gpurun_out/c1_try128_evolving.log 2>&1
timeout 300 python tools/try_scene.py 256 10 evolving > gpurun_out/c1_try256.log 2>&1
timeout 600 python tools/try_scene.py 512 5 evolving > gpurun_out/c1_try512.log 2>&1
tail -3 gpurun_out/c1_pytest.log gpurun_out/c1_try*.log
