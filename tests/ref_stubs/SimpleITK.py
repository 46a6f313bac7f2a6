"""Stub of SimpleITK for importing the reference's metrics.py (metrics.py:4) in tests; the Hausdorff metrics that
would need it are not on the hot path."""
