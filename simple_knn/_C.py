"""`simple_knn._C` as the reference imports it (lib/models/gaussian_model.py:5): distCUDA2 on the B200-native library."""
from gaussianrpg_b200.simple_knn import distCUDA2  # noqa: F401
