set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/c29_tests.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/c29_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/c29_bench_n1.json 2> gpurun_out/c29_bench_n1.err
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_" -s 110 -c 90 --csv --log-file gpurun_out/c29_launches512.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary --no-checksum --no-tolerance > gpurun_out/c29_ncu512.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"k_reorder|k_keys_rank|k_place|k_scan_apply|k_p2g_cell_list|k_p2g_cells|k_p2g_nodes|k_g2p_apic|k_advect" -s 60 -c 16 -o gpurun_out/c29_full128 python bench.py --grid 128 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary --no-checksum --no-tolerance > gpurun_out/c29_ncufull.log 2>&1
python tools/launch_traffic.py gpurun_out/c29_launches512.csv | tail -2
python - <<'PY'
import json
j=json.loads(open("gpurun_out/c29_bench_n1.json").read().strip().splitlines()[-1])
print(j["ms_per_step"], j["value"], j["roofline"]["frac"], j["checksum"]["particle_hash"], j["checksum"]["p2g_field_hash"], j["tolerance_mode"]["ms_per_step"], j["e2e"]["ms_per_step"])
PY
