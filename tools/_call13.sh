set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/c13_pytest.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 120 --csv --log-file gpurun_out/c13_launches512.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary --no-checksum --no-tolerance > gpurun_out/c13_ncu.log 2>&1
tail -5 gpurun_out/c13_pytest.log; tail -2 gpurun_out/c13_ncu.log; wc -l gpurun_out/c13_launches512.csv
