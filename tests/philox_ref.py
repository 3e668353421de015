"""Host reference of the device RNG (tests only): Philox4x32-10 (Salmon et al. 2011) with the
counter/key convention of csrc/epi_sampler.cu."""
import numpy as np

M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
MASK = 0xFFFFFFFF


def philox4x32_10(ctr, key):
    c0, c1, c2, c3 = ctr
    k0, k1 = key
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        hi0, lo0, hi1, lo1 = p0 >> 32, p0 & MASK, p1 >> 32, p1 & MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & MASK, lo1, (hi0 ^ c3 ^ k1) & MASK, lo0
        k0, k1 = (k0 + W0) & MASK, (k1 + W1) & MASK
    return c0, c1, c2, c3


def u01(x):
    return (x + 0.5) / 4294967296.0


def de_partners(seed, chain_id, iteration, tot):
    """the two distinct population members k_accept_propose draws for (chain, iteration)"""
    r = philox4x32_10((chain_id & MASK, chain_id >> 32, iteration & MASK, 1), (seed & MASK, seed >> 32))
    a = (r[0] * tot) >> 32
    b = (r[1] * (tot - 1)) >> 32
    if b >= a:
        b += 1
    return a, b
