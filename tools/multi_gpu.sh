#!/usr/bin/env bash
# Multi-GPU evidence of a build (everything lands in gpurun_out/):
#   gpurun --gpus 2 --timeout 900 -- 'bash tools/multi_gpu.sh 2'
#   gpurun --gpus 8 --timeout 900 -- 'bash tools/multi_gpu.sh 8'
# 1. the 2-GPU tests (push gather bit for bit against the single-GPU run, NCCL gather, sharded API),
# 2. bench.py exactly as the driver launches it (default flags: push gather, two batches in flight, e2e, CPU arm on rank 0),
# 3. the same without e2e for: push / one batch in flight, NCCL all-gather,
# 4. the single-GPU figures of the SAME box (the efficiency denominator).
set -u
cd "$(dirname "$0")/.."
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m pytest tests -m gpu -q -k "push or two_gpu or sharded or gather" 2>&1 | tail -3 | tee gpurun_out/gpu_tests_${N}gpu.log
run() {   # $1 = output tag, rest = bench flags
  local tag=$1; shift
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29800 + RANDOM % 100)) \
      bench.py --gpus $N --steps 400 --warmup 5 "$@" 2> gpurun_out/${tag}.err | tail -1 > gpurun_out/${tag}.json
  python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    j = json.load(open(f"gpurun_out/{tag}.json"))
    print(tag, round(j["value"]), "obj/s", round(j["ms_per_step"], 4), "ms/step; e2e", round((j.get("e2e") or {}).get("value") or 0),
          "clocks", (j.get("clocks") or {}).get("sm_mhz"), (j.get("clocks") or {}).get("reasons"))
except Exception as e:
    print(tag, "FAILED", e); print(open(f"gpurun_out/{tag}.err").read()[-1200:])
PY
}
run r2_bench_${N}gpu
run r2_bench_${N}gpu_one_batch_in_flight --streams 1 --no-cpu-baseline --no-e2e
run r2_bench_${N}gpu_nccl --gather nccl --no-cpu-baseline --no-e2e
for s in 2 1; do
  timeout 200 python bench.py --steps 400 --warmup 5 --streams $s --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 > gpurun_out/r2_bench_${N}gpu_box_single_${s}_in_flight.json
  python -c "
import json; j=json.load(open('gpurun_out/r2_bench_${N}gpu_box_single_${s}_in_flight.json')); print('single GPU, same box, $s in flight:', round(j['value']), 'obj/s', round(j['ms_per_step'],4))"
done
