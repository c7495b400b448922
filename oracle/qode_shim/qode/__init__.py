"""Minimal NumPy stand-in for the external Qode library (see ../README.md)."""
