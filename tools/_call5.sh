set -u
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"fast" -s 6 -c 4 -o gpurun_out/prof_r2b python bench.py --grid 128 --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary --no-checksum > gpurun_out/c5_ncu.log 2>&1
timeout 300 python bench.py --grid 128 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary > gpurun_out/c5_bench128.json 2> gpurun_out/c5_bench128.err
tail -3 gpurun_out/c5_ncu.log; ls -la gpurun_out/prof_r2b.ncu-rep
