"""Top-level ``ficp`` module so that ``from ficp import FractionalICP`` (app.py:20,
tests/test_ficp.py:9, tests/test_rigid_2d_operations.py:8 of the reference) resolves to the
B200-native implementation."""
from coregistrationgame_b200.ficp import FractionalICP  # noqa: F401

__all__ = ["FractionalICP"]
