set -u
mkdir -p gpurun_out
for lanes in 2 3 4; do for ch in 1 2 4 8; do
  EPNP_E2E_LANES=$lanes EPNP_E2E_CHUNKS=$ch timeout 120 python bench.py --steps 60 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import sys,json; j=json.loads(sys.stdin.read()); e=j['e2e']; print('e2e lanes',e['calls_in_flight'],'chunks',e['chunks'],round(e['value']),'obj/s', {k: round(v,3) for k,v in e['step_interval_ms'].items()})"
done; done 2>&1 | tee gpurun_out/e2e_sweep2.txt
timeout 100 python tools/pcie_probe.py 2>&1 | tail -c 600
