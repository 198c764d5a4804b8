"""Framework dispatcher, mirrors models/frameworks/__init__.py of the reference."""


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
