set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
timeout 400 python bench.py --steps 400 --warmup 5 2>gpurun_out/bench.err | tail -1 > gpurun_out/r2_bench.json; cut -c1-2200 gpurun_out/r2_bench.json
: > gpurun_out/r2_configs.jsonl
for c in lm_only amis dense train; do
  timeout 300 python bench.py --config $c --steps 200 --warmup 5 2>>gpurun_out/bench.err | tail -1 >> gpurun_out/r2_configs.jsonl
done
python - <<'PY'
import json
for l in open("gpurun_out/r2_configs.jsonl"):
    j=json.loads(l); print(j["config"]["name"], round(j["value"]), "obj/s", round(j["ms_per_step"],4), "ms  kernels", j["kernels_ms"]["lm_warp_kernel"], j["kernels_ms"]["amis_kernel"], " e2e", round(j["e2e"]["value"]), " cpu", (j.get("cpu_baseline") or {}).get("value"), (j.get("clocks") or {}).get("sm_mhz"))
PY
for lanes in 1 2 3; do for ch in 4 8 16 0; do
  EPNP_E2E_LANES=$lanes EPNP_E2E_CHUNKS=$ch timeout 120 python bench.py --steps 60 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import sys,json; j=json.loads(sys.stdin.read()); e=j['e2e']; print('e2e lanes',e['calls_in_flight'],'chunks',e['chunks'],round(e['value']),'obj/s', e['step_interval_ms'], 'numa', e['host_buffers_on_gpu_numa_node'])"
done; done 2>&1 | tee gpurun_out/e2e_sweep.txt
EPNP_E2E_NUMA=0 timeout 120 python bench.py --steps 60 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import sys,json; j=json.loads(sys.stdin.read()); e=j['e2e']; print('e2e NUMA off lanes',e['calls_in_flight'],'chunks',e['chunks'],round(e['value']),'obj/s', e['step_interval_ms'])" | tee -a gpurun_out/e2e_sweep.txt
for st in 2 3; do timeout 120 python bench.py --steps 400 --warmup 5 --streams $st --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | python -c "
import sys,json; j=json.loads(sys.stdin.read()); print('streams', j['config']['batches_in_flight'], round(j['value']), 'obj/s')"; done | tee gpurun_out/streams.txt
timeout 100 python tools/pcie_probe.py > gpurun_out/pcie_probe.json 2>&1; tail -c 700 gpurun_out/pcie_probe.json
