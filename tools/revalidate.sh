#!/usr/bin/env bash
# The evidence set of a build, one GPU (everything lands in gpurun_out/):
#   gpurun --timeout 1700 -- 'bash tools/revalidate.sh'
#   gpurun --timeout 1200 -- 'bash tools/revalidate.sh ncu'     (the ncu captures; gpurun_out is capped at 64 MiB)
#   gpurun --timeout 900 -- 'bash tools/revalidate.sh side'     (only the kernels off the hot path changed)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run_tests() {
  timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/gpu_tests.log
  timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1 | tee -a gpurun_out/gpu_tests.log
}
run_sanitizer() {
  : > gpurun_out/r2_sanitizer.txt
  for tool in memcheck racecheck synccheck; do
    echo "== $tool" >> gpurun_out/r2_sanitizer.txt
    timeout 500 compute-sanitizer --tool $tool python tools/sanitize.py 2>&1 \
        | grep -E "sanitize driver finished|CHECK FAILED|SUMMARY|Error|error|hazard" | head -30 >> gpurun_out/r2_sanitizer.txt
  done
  cut -c1-200 gpurun_out/r2_sanitizer.txt
}
run_side_kernels() {
  # the steps either side of the path, native against what they replaced (RSLM set-up and solve, GN-step backward, MC epilogue)
  EPNP_BENCH_RSLM=1 EPNP_BENCH_GN_PLUS=1 EPNP_BENCH_MC_EPILOGUE=1 timeout 400 python tools/bench_configs.py > gpurun_out/r2_side_kernels.jsonl 2> gpurun_out/configs.err
  grep -E "RSLM|pose_opt_plus|MC pose loss|training step|evaluate_pnp" gpurun_out/r2_side_kernels.jsonl | cut -c1-330
}
if [ "${1:-bench}" = "side" ]; then
  # after a change that touches only the kernels off the hot path: tests, their timings, sanitizer, their ncu captures
  rm -f gpurun_out/parity_r2.json
  run_tests; run_side_kernels; run_sanitizer
  for k in ${2:-rslm_draw_kernel}; do
    timeout 200 ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -o /tmp/r2_$k python tools/kernel_tour.py > /dev/null 2>&1
    ncu -i /tmp/r2_$k.ncu-rep --page raw --csv > gpurun_out/r2_${k}_raw.csv 2>/dev/null
    ncu -i /tmp/r2_$k.ncu-rep --page source --csv > gpurun_out/r2_${k}_source.csv 2>/dev/null
  done
  exit 0
fi
if [ "${1:-bench}" = "ncu" ]; then
  [ "${2:-}" = "with-tests" ] && { rm -f gpurun_out/parity_r2.json; run_tests; run_sanitizer; }
  # one `--set full` capture per kernel; the two hot kernels keep their report (source page), the others leave CSVs
  for k in amis_kernel lm_warp_kernel; do
    timeout 200 ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -o gpurun_out/r2_$k \
        python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --streams 1 > /dev/null 2>&1
  done
  for k in cost_backward_kernel cost_kernel rslm_draw_kernel rslm_kernel gn_plus_backward_kernel adaptive_delta_kernel mc_epilogue_kernel evaluate_full_kernel; do
    timeout 200 ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -o /tmp/r2_$k python tools/kernel_tour.py > /dev/null 2>&1
    ncu -i /tmp/r2_$k.ncu-rep --page raw --csv > gpurun_out/r2_${k}_raw.csv 2>/dev/null
    ncu -i /tmp/r2_$k.ncu-rep --page source --csv > gpurun_out/r2_${k}_source.csv 2>/dev/null
  done
  timeout 200 ncu --set full --clock-control none -k regex:amis_kernel -c 1 -o /tmp/r2_amis_kernel_dense \
      python bench.py --config dense --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --streams 1 > /dev/null 2>&1
  ncu -i /tmp/r2_amis_kernel_dense.ncu-rep --page raw --csv > gpurun_out/r2_amis_kernel_dense_raw.csv 2>/dev/null
  timeout 200 ncu --set full --clock-control none -k regex:lm_warp_kernel -c 1 -o /tmp/r2_lm_warp_kernel_dense \
      python bench.py --config dense --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --streams 1 > /dev/null 2>&1
  ncu -i /tmp/r2_lm_warp_kernel_dense.ncu-rep --page raw --csv > gpurun_out/r2_lm_warp_kernel_dense_raw.csv 2>/dev/null
  du -sh gpurun_out; ls -la gpurun_out | awk '{print $5, $9}' | tail -25
  exit 0
fi
rm -f gpurun_out/parity_r2.json gpurun_out/*.ncu-rep
run_tests
timeout 400 python bench.py --steps 400 --warmup 5 2>gpurun_out/bench.err | tail -1 > gpurun_out/r2_bench_1gpu.json; cut -c1-300 gpurun_out/r2_bench_1gpu.json
timeout 200 python bench.py --steps 400 --warmup 5 --streams 1 --no-cpu-baseline 2>>gpurun_out/bench.err | tail -1 > gpurun_out/r2_bench_1gpu_one_batch_in_flight.json
: > gpurun_out/r2_configs.jsonl
for c in lm_only amis dense train; do
  timeout 300 python bench.py --config $c --steps 200 --warmup 5 2>>gpurun_out/bench.err | tail -1 >> gpurun_out/r2_configs.jsonl
done
python - <<'PY'
import json
for f in ("gpurun_out/r2_bench_1gpu.json", "gpurun_out/r2_bench_1gpu_one_batch_in_flight.json"):
    j = json.load(open(f)); print(j["config"]["name"], j["config"]["batches_in_flight"], "in flight:", round(j["value"]), "obj/s; e2e", round(j["e2e"]["value"]), j["e2e"]["step_interval_ms"], "cpu", (j.get("cpu_baseline") or {}).get("value"))
for l in open("gpurun_out/r2_configs.jsonl"):
    j = json.loads(l); print(j["config"]["name"], round(j["value"]), "obj/s", round(j["ms_per_step"], 4), "ms; kernels", j["kernels_ms"]["lm_warp_kernel"], j["kernels_ms"]["amis_kernel"], " e2e", round(j["e2e"]["value"]), " cpu", (j.get("cpu_baseline") or {}).get("value"))
PY
run_side_kernels
# launch list of the bench command (cold-cache, serialised: shares only)
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 200 python tools/phase_profile.py 4096 512 512 > gpurun_out/r2_phase_cycles_amis.txt 2>&1; tail -9 gpurun_out/r2_phase_cycles_amis.txt
run_sanitizer
