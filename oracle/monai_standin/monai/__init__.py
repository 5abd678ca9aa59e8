"""Minimal MONAI 1.3.0 stand-in (test infrastructure; see ../README.md)."""
__version__ = "1.3.0+standin"
