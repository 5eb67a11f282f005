set -u
mkdir -p gpurun_out
nvidia-smi -L
timeout 300 python -m pytest tests/test_peer_gather_gpu.py tests/test_push_gather_gpu.py -q 2>&1 | tail -4 | tee gpurun_out/two_gpu_tests.log
i=0
for g in "--gather nccl" "--gather nccl --nccl-max-ctas 4" "--gather peer" "--gather push" "--gather push --streams 2" "--gather nccl --streams 2"; do
  i=$((i + 1))
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$i \
      bench.py --gpus 2 --steps 400 --warmup 5 $g --no-cpu-baseline --no-e2e 2> gpurun_out/two_gpu_$i.err | tail -1 > gpurun_out/two_gpu_$i.json
  python - <<PY
import json
try:
    j=json.load(open("gpurun_out/two_gpu_$i.json")); print("$g", round(j["value"]), "obj/s", round(j["ms_per_step"],4), "ms/step", j["kernels_ms"], (j.get("clocks") or {}).get("sm_mhz"))
except Exception as e:
    print("$g", "FAILED", e); print(open("gpurun_out/two_gpu_$i.err").read()[-1500:])
PY
done
