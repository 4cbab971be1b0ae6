set -u
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"k_g2p_apic|k_advect|k_reorder|k_p2g_cells|k_p2g_nodes|k_p2g_cell_list|k_keys_rank|k_place|k_scan_apply" -s 60 -c 30 -o gpurun_out/prof_r2a python bench.py --grid 128 --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary --no-checksum > gpurun_out/c4_ncu.log 2>&1
timeout 300 python bench.py --grid 128 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary > gpurun_out/c4_bench128.json 2> gpurun_out/c4_bench128.err
tail -3 gpurun_out/c4_ncu.log; ls -la gpurun_out/prof_r2a.ncu-rep
