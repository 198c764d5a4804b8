"""Training (autograd) path of NeuS volume_render -- built in a later milestone."""


def volume_render_train(*a, **k):
    raise NotImplementedError(
        "neurecon_b200: NeuS volume_render under autograd is not built yet; wrap inference in torch.no_grad()")
