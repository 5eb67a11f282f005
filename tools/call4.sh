set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/gpu_tests.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:amis_kernel -c 1 -o gpurun_out/r2_amis_full \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:lm_warp_kernel -c 1 -o gpurun_out/r2_lm_full \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
