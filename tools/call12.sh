set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/gpu_tests.log
for c in fused dense lm_only amis; do
  timeout 300 python bench.py --config $c --steps 400 --warmup 5 --no-cpu-baseline --no-e2e --streams 1 2>>gpurun_out/bench.err | tail -1 | python -c "
import sys,json; j=json.loads(sys.stdin.read()); print(j['config']['name'], round(j['value']), 'obj/s', round(j['ms_per_step'],4), 'ms kernels', j['kernels_ms']['lm_warp_kernel'], j['kernels_ms']['amis_kernel'])"
done
timeout 300 python bench.py --steps 400 --warmup 5 --no-cpu-baseline --no-e2e 2>>gpurun_out/bench.err | tail -1 | python -c "
import sys,json; j=json.loads(sys.stdin.read()); print('fused, 2 in flight', round(j['value']), 'obj/s', round(j['ms_per_step'],4))"
