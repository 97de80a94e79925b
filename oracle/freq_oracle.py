"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy, fp32) of the frequency encoder
(freqencoder/src/freqencoder.cu:31-58 forward, :60-94 backward).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline may import this.

The reference evaluates __sinf, the fast-math sine of the GPU, whose error against the correctly rounded sine is bounded
by 2^-21.41 absolute on [-pi, pi] and grows with the argument; the oracle uses np.sin, so forward comparisons against the
reference carry that tolerance (tests/golden/freq.npz pins it; the GPU parity test against the reference kernel is
bit-exact because both sides call the same intrinsic)."""
import numpy as np


def forward(inputs, degree):
    """inputs [B, D] -> [B, D + 2 D degree]: x, then per octave f: sin(x 2^f) for all dims, sin(x 2^f + pi/2) for all dims."""
    x = np.asarray(inputs, dtype=np.float32)
    cols = [x]
    half_pi = np.float32(np.float32(3.141592653589793) / np.float32(2))
    for f in range(degree):
        xs = np.ldexp(x, f).astype(np.float32)                      # scalbnf(x, f), exact
        cols.append(np.sin(xs.astype(np.float64)).astype(np.float32))
        cols.append(np.sin((xs + half_pi).astype(np.float32).astype(np.float64)).astype(np.float32))
    return np.concatenate(cols, axis=1)


def backward(grad, outputs, D, degree):
    """grad, outputs [B, C] -> grad_inputs [B, D], the accumulation order of freqencoder.cu:82-90 in fp32."""
    g = np.asarray(grad, dtype=np.float32)
    o = np.asarray(outputs, dtype=np.float32)
    res = g[:, :D].copy()
    for f in range(degree):
        gs, gc = g[:, D + 2 * f * D: D + (2 * f + 1) * D], g[:, D + (2 * f + 1) * D: D + (2 * f + 2) * D]
        os_, oc = o[:, D + 2 * f * D: D + (2 * f + 1) * D], o[:, D + (2 * f + 1) * D: D + (2 * f + 2) * D]
        res = (res + np.float32(2.0 ** f) * (gs * oc - gc * os_).astype(np.float32)).astype(np.float32)
    return res
