"""Drop-in replacement of the reference package `gridencoder` (put raw_ngp_b200/dropin first on sys.path)."""
from raw_ngp_b200.gridencoder import GridEncoder, grid_encode  # noqa: F401
from raw_ngp_b200.gridencoder import grid  # noqa: F401
