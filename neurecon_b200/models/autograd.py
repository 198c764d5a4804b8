"""Training (autograd) entry points of the MLP kernels -- implemented in a later milestone."""


def _todo(*a, **k):
    raise NotImplementedError(
        "neurecon_b200: the training backward of the fused MLPs is not built yet; call under torch.no_grad()")


sdf_forward_autograd = _todo
sdf_forward_with_nablas_autograd = _todo
radiance_forward_autograd = _todo
