#!/usr/bin/env bash
# First GPU call of the next round: run what round 1 wrote after its GPU minutes were spent (all of it is off by
# default and pinned on the CPU only), then time the liquid-SDF variants side by side.
#
#   gpurun --timeout 300 -- 'bash tools/run_experimental.sh'
#
# Outputs under gpurun_out/: experimental_pytest.log, liquid_sdf_variant{0,1}.json
set -u
mkdir -p gpurun_out
FFB200_TEST_EXPERIMENTAL=1 timeout 200 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "experimental" 2>&1 | tee gpurun_out/experimental_pytest.log | tail -15
for v in 0 1; do
    FFB200_SDF_VARIANT=$v timeout 60 python tools/bench_liquid_sdf.py --steps 10 > gpurun_out/liquid_sdf_variant$v.json 2> gpurun_out/liquid_sdf_variant$v.err
    cat gpurun_out/liquid_sdf_variant$v.json
done
