"""Autograd bridge for the differentiable outputs of the layer (cost_init, Monte-Carlo cost of the
samples -> log-weights).  SURVEY.md section 8(f1): the backward kernel is the next row after the forward
path; until it lands, asking for gradients fails loudly instead of silently running a PyTorch path."""


def evaluate_cost_autograd(x3d, x2d, w2d, pose, camera, cost_fun):
    raise NotImplementedError(
        "gradients through the native EPro-PnP cost are not built yet (forward/inference only in this "
        "release); call under torch.no_grad() or detach the inputs")
