set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/c2_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/c2_bench_n1.json 2> gpurun_out/c2_bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/c2_bench_ref.json 2> gpurun_out/c2_bench_ref.err
tail -5 gpurun_out/c2_pytest.log; tail -c 1500 gpurun_out/c2_bench_n1.json; tail -5 gpurun_out/c2_bench_n1.err
