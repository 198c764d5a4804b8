"""Framework dispatcher: ``get_model(args)`` as models/frameworks/__init__.py of the reference (same framework names,
same 5-tuple ``(model, trainer, render_kwargs_train, render_kwargs_test, renderer)``), plus ``get_framework(name)``
for callers that want the module."""


def get_framework(name):
    name = name.lower()
    if name == "neus":
        from . import neus
        return neus
    if name == "volsdf":
        from . import volsdf
        return volsdf
    if name == "unisurf":
        from . import unisurf
        return unisurf
    raise NotImplementedError("unknown framework %r" % name)


def get_model(args):
    if args.model.framework not in ("UNISURF", "NeuS", "VolSDF"):
        raise NotImplementedError
    return get_framework(args.model.framework).get_model(args)
