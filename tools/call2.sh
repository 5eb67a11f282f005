set -u
mkdir -p gpurun_out
V=epro-pnp_b200/lib/variants
for v in default six_ctas_huber_m six_hm_nocf five_hm_nocf four_hm_nocf_packed; do
  EPNP_LIB=$PWD/$V/libepropnp_b200.$v.so timeout 120 python tools/split_probe.py 2>&1 | tail -1 | tee -a gpurun_out/split_probe.jsonl
done
timeout 600 python tools/variants.py run default six_ctas_huber_m six_hm_nocf six_hm_nocf_packed five_hm_nocf five_hm_nocf_packed four_hm_nocf_packed --min-gain 9 2>&1 | tail -9 | cut -c1-330
# ncu full of the fused kernel of the leading variant, and of the shipped build's AMIS-only / LM-only kernels
cp epro-pnp_b200/lib/libepropnp_b200.so /tmp/shipped.so
cp $V/libepropnp_b200.six_hm_nocf.so epro-pnp_b200/lib/libepropnp_b200.so
timeout 300 ncu --set full --clock-control none --import-source on -k regex:solve_kernel -c 1 -o gpurun_out/six_hm_nocf_fused_full \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > /dev/null 2>&1
cp /tmp/shipped.so epro-pnp_b200/lib/libepropnp_b200.so
for v in "" six_ctas_huber_m; do
  timeout 150 python tools/phase_profile.py 4096 512 512 $v > gpurun_out/phase_cycles_${v:-shipped}.txt 2>&1
  tail -14 gpurun_out/phase_cycles_${v:-shipped}.txt
done
ls -la gpurun_out | tail -8
