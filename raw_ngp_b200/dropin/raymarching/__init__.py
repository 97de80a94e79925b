"""Drop-in replacement of the reference package `raymarching`."""
from raw_ngp_b200.raymarching.raymarching import *  # noqa: F401,F403
