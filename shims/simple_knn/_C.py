"""Drop-in shim for ``from simple_knn._C import distCUDA2`` (/root/reference/src/models/gaussian.py:4)."""
from manus_b200.knn import distCUDA2  # noqa: F401
