set -u
mkdir -p gpurun_out
rm -f gpurun_out/parity_r2.json
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -12 | tee gpurun_out/gpu_tests.log
for c in dense lm_only; do
  timeout 300 python bench.py --config $c --steps 200 --warmup 5 --no-cpu-baseline 2>>gpurun_out/bench.err | tail -1 | python -c "
import sys,json; j=json.loads(sys.stdin.read()); print(j['config']['name'], round(j['value']), 'obj/s', round(j['ms_per_step'],4), 'ms kernels', j['kernels_ms']['lm_warp_kernel'], j['kernels_ms']['amis_kernel'], 'e2e', round(j['e2e']['value']))"
done
timeout 300 python bench.py --steps 400 --warmup 5 --no-cpu-baseline 2>>gpurun_out/bench.err | tail -1 | python -c "
import sys,json; j=json.loads(sys.stdin.read()); print(j['config']['name'], round(j['value']), 'obj/s', round(j['ms_per_step'],4), 'ms kernels', j['kernels_ms']['lm_warp_kernel'], j['kernels_ms']['amis_kernel'], 'e2e', round(j['e2e']['value']), j['e2e']['step_interval_ms'])"
