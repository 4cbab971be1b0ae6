set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > gpurun_out/c6_pytest.log
timeout 300 python bench.py --grid 128 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary > gpurun_out/c6_bench128.json 2> gpurun_out/c6_bench128.err
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"fast|k_advect|k_g2p" -s 10 -c 14 -o gpurun_out/prof_r2c python bench.py --grid 128 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary --no-checksum > gpurun_out/c6_ncu.log 2>&1
tail -4 gpurun_out/c6_pytest.log; tail -2 gpurun_out/c6_ncu.log
