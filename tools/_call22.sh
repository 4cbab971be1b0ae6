set -u
mkdir -p gpurun_out
# 512^3: launch list with DRAM bytes (one metric pass per kernel), two steady-state steps
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_" -s 110 -c 80 --csv --log-file gpurun_out/c22_launches512.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary --no-checksum --no-tolerance > gpurun_out/c22_ncu512.log 2>&1
# 128^3: same, exact then tolerance steps
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_" -s 110 -c 400 --csv --log-file gpurun_out/c22_launches128.csv python bench.py --grid 128 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary --no-checksum > gpurun_out/c22_ncu128.log 2>&1
# full bench line (N=1), the record for the docs
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/c22_bench_n1.json 2> gpurun_out/c22_bench_n1.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/c22_bench_ref.json 2> gpurun_out/c22_bench_ref.err
wc -l gpurun_out/c22_launches512.csv gpurun_out/c22_launches128.csv; tail -c 300 gpurun_out/c22_bench_n1.json
