set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/c26_tests.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/c26_tests.log
# the record lines (N=1 and the reference arm), as the driver runs them
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/c26_bench_n1.json 2> gpurun_out/c26_bench_n1.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/c26_bench_ref.json 2> gpurun_out/c26_bench_ref.err
# launch lists with DRAM bytes
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_" -s 110 -c 90 --csv --log-file gpurun_out/c26_launches512.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary --no-checksum --no-tolerance > gpurun_out/c26_ncu512.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_" -s 110 -c 90 --csv --log-file gpurun_out/c26_launches128.csv python bench.py --grid 128 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary --no-checksum --no-tolerance > gpurun_out/c26_ncu128.log 2>&1
# one --set full capture of the exact step's kernels at 128^3 (one step)
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"k_reorder|k_keys_rank|k_place|k_scan_apply|k_p2g_cell_list|k_p2g_cells|k_p2g_nodes|k_g2p_apic|k_advect" -s 60 -c 16 -o gpurun_out/c26_full128 python bench.py --grid 128 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary --no-checksum --no-tolerance > gpurun_out/c26_ncufull.log 2>&1
ls -la gpurun_out/c26_full128.ncu-rep
python tools/launch_traffic.py gpurun_out/c26_launches512.csv | tail -2
tail -c 400 gpurun_out/c26_bench_n1.json; echo; tail -c 300 gpurun_out/c26_bench_ref.json
