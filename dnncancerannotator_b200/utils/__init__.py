from . import losses, load  # noqa: F401
