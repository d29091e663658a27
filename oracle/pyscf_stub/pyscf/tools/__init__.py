from . import cubegen, molden  # noqa: F401
