set -u
mkdir -p gpurun_out
for v in "" exp/lib_lm_unstaged.so exp/lib_amis5.so; do
  if [ -n "$v" ]; then export EPNP_LIB=$PWD/epro-pnp_b200/lib/$v; else unset EPNP_LIB; fi
  timeout 120 python tools/split_probe.py 2>&1 | tail -1
done | tee gpurun_out/split_probe_exp.jsonl
unset EPNP_LIB
timeout 200 python tools/phase_profile.py 4096 512 512 2>&1 | tail -10 | tee gpurun_out/phase_cycles_amis.txt
