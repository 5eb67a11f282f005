/* Minimal C host of libepropnp_b200.so: what a non-Python caller binds.  Needs no GPU to run the checks below
 * (version, defaults, capacity query, argument validation); with a device it would go on to allocate the buffers
 * with cudaMalloc and call epnp_lm_amis_fused_f32 / epnp_lm_amis_fused_host_f32 exactly like the ctypes layer.
 *
 *   gcc -std=c99 -Iinclude examples/host_c.c -o host_c -Lepro-pnp_b200/lib -lepropnp_b200 -Wl,-rpath,epro-pnp_b200/lib
 */
#include <stdio.h>
#include <string.h>

#include "epropnp_b200.h"

int main(void) {
    EpnpParams p;
    int failures = 0;
    if (epnp_abi_version() != EPNP_ABI_VERSION) { printf("abi mismatch\n"); ++failures; }
    epnp_default_params(&p, 6);
    if (p.dof != 6 || p.lm_iter != 10 || p.mc_samples != 512 || p.mc_iter != 4 || p.acg_mle_iter != 3) ++failures;
    if (sizeof(EpnpParams) != 64) ++failures;
    printf("abi %d, defaults: lm_iter=%d radius=%.1f eps=%g M=%d I=%d\n", epnp_abi_version(), p.lm_iter,
           p.initial_radius, p.eps, p.mc_samples, p.mc_iter);
    printf("largest resident N: fused(M=512,I=4) %d, LM-only %d\n", epnp_max_points(6, 512, 4), epnp_max_points(6, 0, 0));
    if (epnp_max_points(6, 512, 4) < 4096) ++failures;
    /* argument validation happens before any CUDA call */
    if (epnp_lm_solve_f32(NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, 4, 64, &p, NULL)
        != EPNP_ERR_BAD_ARG) ++failures;
    if (epnp_cost_backward_f32(NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, 0, NULL, NULL, 0, NULL, NULL, NULL,
                               NULL, 4, 64, 6, 0.1f, NULL) != EPNP_ERR_BAD_ARG) ++failures;
    if (strcmp(epnp_error_string(EPNP_ERR_TOO_MANY_POINTS), "") == 0) ++failures;
    if (epnp_fused_workspace_bytes(4096, 512, &p) < (size_t)4096 * 512 * 28) ++failures;
    printf("%s\n", failures ? "FAILED" : "ok");
    return failures;
}
