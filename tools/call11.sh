set -u
mkdir -p gpurun_out
nvidia-smi -L | wc -l
i=0
for g in "--gather push" "--gather push --streams 1" "--gather nccl --nccl-max-ctas 4"; do
  i=$((i + 1))
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2981$i \
      bench.py --gpus 8 --steps 400 --warmup 5 $g --no-cpu-baseline --no-e2e 2> gpurun_out/eight_gpu_$i.err | tail -1 > gpurun_out/eight_gpu_$i.json
  python - <<PY
import json
try:
    j=json.load(open("gpurun_out/eight_gpu_$i.json")); print("8 GPUs $g:", round(j["value"]), "obj/s", round(j["ms_per_step"],4), "ms/step in_loop", j.get("in_loop"), (j.get("clocks") or {}).get("sm_mhz"), (j.get("clocks") or {}).get("reasons"))
except Exception as e:
    print("$g", "FAILED", e); print(open("gpurun_out/eight_gpu_$i.err").read()[-800:])
PY
done
timeout 100 python bench.py --steps 400 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | python -c "
import sys,json; j=json.loads(sys.stdin.read()); print('single GPU same box (2 in flight)', round(j['value']), 'obj/s', round(j['ms_per_step'],4))"
timeout 100 python bench.py --steps 400 --warmup 5 --streams 1 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | python -c "
import sys,json; j=json.loads(sys.stdin.read()); print('single GPU same box (1 in flight)', round(j['value']), 'obj/s', round(j['ms_per_step'],4))"
