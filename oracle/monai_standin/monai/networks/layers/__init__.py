from .factories import Conv  # noqa: F401
