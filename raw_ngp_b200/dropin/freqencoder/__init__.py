"""Drop-in replacement of the reference package `freqencoder`."""
from raw_ngp_b200.freqencoder import FreqEncoder, freq_encode  # noqa: F401
from raw_ngp_b200.freqencoder import freq  # noqa: F401
