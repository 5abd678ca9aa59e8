from .basic_unet import Down, TwoConv, UpCat  # noqa: F401
