"""Drop-in replacement of the reference package `shencoder`."""
from raw_ngp_b200.shencoder import SHEncoder, sh_encode  # noqa: F401
from raw_ngp_b200.shencoder import sphere_harmonics  # noqa: F401
