"""Drop-in replacement of the reference's `epropnp` package (same module and class names) whose
batched hot loops run in hand-written sm_100a CUDA behind libepropnp_b200.so:

    epropnp.epropnp               EProPnP6DoF / EProPnP4DoF / EProPnPBase / cholesky_wrapper
    epropnp.levenberg_marquardt   LMSolver / RSLMSolver
    epropnp.camera                PerspectiveCamera
    epropnp.cost_fun              HuberPnPCost / AdaptiveHuberPnPCost
    epropnp.common                evaluate_pnp, pnp_normalize, pnp_denormalize, rotation helpers
    epropnp.distributions         AngularCentralGaussian / VonMisesUniformMix
    epropnp.builder               build_pnp / build_camera / build_cost_fun + registries (detection-style configs)

Put `<repo>/epro-pnp_b200` on sys.path (instead of the reference checkout) and existing imports keep
working.  The solve / Monte-Carlo paths have no CPU or PyTorch fallback: CPU tensors raise.
"""
