set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "attribute" 2>&1 | tail -15 > gpurun_out/c7_pytest.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"fast|k_advect_exact_list" -s 8 -c 6 -o gpurun_out/prof_r2d python bench.py --grid 128 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary --no-checksum > gpurun_out/c7_ncu.log 2>&1
tail -4 gpurun_out/c7_pytest.log; tail -2 gpurun_out/c7_ncu.log
