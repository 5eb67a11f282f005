set -u
mkdir -p gpurun_out
for v in "" exp/lib_u6c5.so exp/lib_u8c4.so exp/lib_u4c4.so exp/lib_u2c5.so exp/lib_u3c6.so; do
  if [ -n "$v" ]; then export EPNP_LIB=$PWD/epro-pnp_b200/lib/$v; else unset EPNP_LIB; fi
  timeout 120 python tools/split_probe.py 2>&1 | tail -1
done | tee gpurun_out/split_probe_unroll.jsonl
