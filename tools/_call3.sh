set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -k "dropin or tolerance" 2>&1 | tail -40 > gpurun_out/c3_pytest.log
timeout 300 python bench.py --grid 128 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/c3_bench128.json 2> gpurun_out/c3_bench128.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary --no-e2e > gpurun_out/c3_bench512.json 2> gpurun_out/c3_bench512.err
tail -8 gpurun_out/c3_pytest.log; tail -3 gpurun_out/c3_bench128.err gpurun_out/c3_bench512.err
