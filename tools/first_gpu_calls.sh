#!/usr/bin/env bash
# The GPU calls that measure what was written after round 1's GPU budget was spent.
# Every step is wrapped in its own timeout and writes under gpurun_out/; run each block as ONE gpurun call:
#
#   gpurun --timeout 1700 -- 'bash tools/first_gpu_calls.sh variants'        # 1 GPU, ~15 min: bench of every variant, then the parity tests of the faster ones
#   gpurun --timeout 1500 -- 'bash tools/first_gpu_calls.sh experimental'    # 1 GPU, ~15 min
#   gpurun --gpus 2 --timeout 1200 -- 'bash tools/first_gpu_calls.sh two_gpu' # 2 GPUs, ~6 min (charged x2)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
case "${1:-}" in
  variants)
    # A/B of the kernel build options: per variant the GPU parity + edge tests, then one bench line
    python tools/variants.py build > gpurun_out/variants_build.log 2>&1
    timeout 1500 python tools/variants.py run 2>&1 | tee gpurun_out/variants_run.log | tail -20
    # where a CTA's time goes in the shipped build and in the two leading candidates
    for v in "" six_ctas_huber_m five_ctas_mma; do
      timeout 120 python tools/phase_profile.py 4096 512 512 $v > gpurun_out/phase_cycles_${v:-shipped}.txt 2>&1
    done
    ;;
  experimental)
    # the two opt-in kernels, their timings, the copy ceiling of the box, the e2e chunk sweep, the tcgen05 probe
    EPNP_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_rslm_fused_gpu.py tests/test_gn_plus_backward_gpu.py tests/test_mc_epilogue_gpu.py -q \
        2>&1 | tee gpurun_out/experimental_tests.log | tail -5
    # memcheck over the kernels that have never run on hardware (epilogue, push with local stand-in peers)
    EPNP_SANITIZE_EXPERIMENTAL=1 timeout 300 compute-sanitizer --tool memcheck python tools/sanitize.py 2>&1 \
        | grep -E "sanitize driver finished|SUMMARY|Error|error" | head -20 | tee gpurun_out/sanitizer_experimental.txt
    EPNP_BENCH_RSLM=1 EPNP_BENCH_GN_PLUS=1 EPNP_BENCH_MC_EPILOGUE=1 timeout 400 python tools/bench_configs.py > gpurun_out/configs.jsonl 2> gpurun_out/configs.err
    timeout 120 python tools/pcie_probe.py > gpurun_out/pcie_probe.json 2>&1
    for c in 4 7 8 14 0; do     # 0 = automatic: one wave of resident CTAs per chunk
      EPNP_E2E_CHUNKS=$c timeout 200 python bench.py --steps 100 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/e2e_chunks_$c.json
    done
    # two host-buffer calls in flight (double-buffered workspace + results): step i+1 uploads under step i's solve
    for c in 4 7; do
      EPNP_E2E_LANES=2 EPNP_E2E_CHUNKS=$c timeout 200 python bench.py --steps 100 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/e2e_lanes2_chunks_$c.json
    done
    # pinned input buffers first-touched on the GPU's NUMA node (upload measured at 40 GB/s in round 1, download 57)
    EPNP_E2E_NUMA=1 EPNP_E2E_CHUNKS=7 timeout 200 python bench.py --steps 100 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/e2e_numa_chunks_7.json
    EPNP_E2E_NUMA=1 EPNP_E2E_LANES=2 EPNP_E2E_CHUNKS=7 timeout 200 python bench.py --steps 100 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/e2e_numa_lanes2_chunks_7.json
    python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/e2e_*.json")):
    try:
        e = json.load(open(f))["e2e"]; print(f, round(e["value"]), "objects/s", e.get("chunks"), e.get("calls_in_flight"), e.get("host_buffers_on_gpu_numa_node"))
    except Exception as x:
        print(f, "unreadable:", x)
PY
    # batches in flight: does the next batch fill the previous batch's last wave?  (1 = the shipped bench configuration)
    for st in 1 2 3; do
      timeout 200 python bench.py --steps 400 --warmup 3 --streams $st --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 > gpurun_out/streams_$st.json
      cut -c1-200 gpurun_out/streams_$st.json
    done
    (cd tools/tc_probe && nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -o tc_probe tc_probe.cu \
        && timeout 60 ./tc_probe) > gpurun_out/tc_probe.json 2>&1
    tail -c 600 gpurun_out/tc_probe.json
    ;;
  two_gpu)
    EPNP_TEST_PEER_GATHER=1 timeout 300 python -m pytest tests/test_peer_gather_gpu.py tests/test_push_gather_gpu.py -q 2>&1 | tee gpurun_out/peer_gather_test.log | tail -3
    i=0
    for g in "--gather nccl" "--gather nccl-coalesced" "--gather nccl --nccl-max-ctas 2" "--gather nccl --nccl-max-ctas 8" "--gather peer" "--gather push" "--gather nccl --streams 2" "--gather peer --streams 2"; do
      i=$((i + 1))
      timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$i \
          bench.py --gpus 2 --steps 300 --warmup 5 $g --no-cpu-baseline --no-e2e 2> gpurun_out/two_gpu_$i.err | tail -1 > gpurun_out/two_gpu_$i.json
      cat gpurun_out/two_gpu_$i.json | cut -c1-300
    done
    ;;
  revalidate)
    # after a variant has been adopted as the default build: the full evidence set again
    #   gpurun --timeout 1500 -- 'bash tools/first_gpu_calls.sh revalidate'
    timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tee gpurun_out/gpu_tests.log | tail -3
    : > gpurun_out/sanitizer.txt
    for tool in memcheck racecheck synccheck initcheck; do
      echo "== $tool" >> gpurun_out/sanitizer.txt
      EPNP_SANITIZE_EXPERIMENTAL=1 timeout 400 compute-sanitizer --tool $tool python tools/sanitize.py 2>&1 \
          | grep -E "sanitize driver finished|SUMMARY|Error|error|hazard" | head -40 >> gpurun_out/sanitizer.txt
    done
    cat gpurun_out/sanitizer.txt | cut -c1-200
    timeout 200 python bench.py --steps 400 --warmup 3 2>/dev/null | tail -1 > gpurun_out/bench.json; cut -c1-400 gpurun_out/bench.json
    timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
        python bench.py --steps 5 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
    timeout 300 ncu --set full --clock-control none --import-source on -k regex:solve_kernel -c 1 -o gpurun_out/fused_full \
        python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > /dev/null 2>&1
    timeout 200 python tools/phase_profile.py 4096 512 512 > gpurun_out/phase_cycles.txt 2>&1; tail -15 gpurun_out/phase_cycles.txt
    # the other BASELINE configs (LM-only, dense, training step, detection) on the adopted build
    timeout 400 python tools/bench_configs.py > gpurun_out/configs.jsonl 2> gpurun_out/configs.err; cut -c1-160 gpurun_out/configs.jsonl
    # afterwards, in the container:  python tools/ncu_summary.py gpurun_out/fused_full.ncu-rep --tag rN --write-traffic
    python tools/sass_identity.py --write epro-pnp_b200/lib/libepropnp_b200.so > /dev/null && cp profiles/validated_sass.json gpurun_out/
    ;;
  *)
    echo "usage: $0 variants | experimental | two_gpu | revalidate" >&2
    exit 2
    ;;
esac
