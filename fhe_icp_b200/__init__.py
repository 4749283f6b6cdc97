"""B200-native engine for the encrypted-similarity hot path of shipstone-labs/fhe-icp.

Host side mirrors the reference's ``fhe_similarity.py`` / ``batch_operations.py`` /
``fhe_cli.py`` surface; the arithmetic is hand-written CUDA for sm_100a behind the C-ABI in
``include/fhe_b200.h`` (``libfhe_b200.so``, built in-tree by ``_native.build()``).
"""
from .fhe_similarity import FHESimilarityModel  # noqa: F401
from .linear_model import LinearRegression, SGDRegressor  # noqa: F401

__all__ = ["FHESimilarityModel", "LinearRegression", "SGDRegressor"]
