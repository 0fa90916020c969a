from .losses import losswrapper  # noqa: F401
