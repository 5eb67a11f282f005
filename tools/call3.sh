set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 | tee gpurun_out/gpu_tests.log
timeout 120 python tools/split_probe.py 2>&1 | tail -1 | tee gpurun_out/split_probe_new.jsonl
timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu-baseline 2>gpurun_out/bench.err | tail -1 | tee gpurun_out/bench_new.json | cut -c1-600
