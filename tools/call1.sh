set -u
mkdir -p gpurun_out
nvidia-smi -L | head -2
timeout 1100 python tools/variants.py run default six_ctas_huber_m six_ctas_plain_sweep six_ctas five_ctas_plain_sweep five_ctas everything four_ctas_same_code lm_cost_first lm_norefine sweep_huber_m sweep_rsq sweep_split sweep_noclamp fast_blocksum lm_packed amis_lse no_lw --max-tested 4 --test-timeout 300 2>&1 | tee gpurun_out/variants_run.log | tail -30
EPNP_TEST_EXPERIMENTAL=1 timeout 400 python -m pytest tests/test_rslm_fused_gpu.py tests/test_gn_plus_backward_gpu.py tests/test_mc_epilogue_gpu.py -q 2>&1 | tee gpurun_out/experimental_tests.log | tail -15
