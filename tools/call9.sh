set -u
mkdir -p gpurun_out
timeout 200 python tools/p2p_probe.py 2>gpurun_out/p2p_probe.err | tee gpurun_out/p2p_probe.jsonl | cut -c1-500
timeout 150 python -m pytest tests/test_push_gather_gpu.py tests/test_peer_gather_gpu.py -q -x 2>&1 | tail -4 | tee gpurun_out/two_gpu_tests.log
i=0
for g in "--gather push" "--gather nccl --nccl-max-ctas 4" "--gather push --streams 2" "--gather nccl --nccl-max-ctas 4 --streams 2"; do
  i=$((i + 1))
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2971$i \
      bench.py --gpus 2 --steps 400 --warmup 5 $g --no-cpu-baseline --no-e2e 2> gpurun_out/two_gpu_$i.err | tail -1 > gpurun_out/two_gpu_$i.json
  python - <<PY
import json
try:
    j=json.load(open("gpurun_out/two_gpu_$i.json")); print("$g", round(j["value"]), "obj/s", round(j["ms_per_step"],4), "ms/step in_loop", j.get("in_loop"))
except Exception as e:
    print("$g", "FAILED", e); print(open("gpurun_out/two_gpu_$i.err").read()[-500:])
PY
done
timeout 100 python bench.py --steps 400 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | python -c "
import sys,json; j=json.loads(sys.stdin.read()); print('single GPU same box', round(j['value']), 'obj/s', round(j['ms_per_step'],4), j['kernels_ms'])"
