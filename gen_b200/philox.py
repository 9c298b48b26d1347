"""Host-side Philox4x32-10 with the draw layout of gen_b200/csrc/gsmc_rng.cuh, for the few scalar draws the
host interface itself needs (the chunk-merge Bernoulli of importance_resampling). The per-particle draws never
come from here: they are generated inside the CUDA kernels."""

M0, M1 = 0xD2511F53, 0xCD9E8D57
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = 0xFFFFFFFF
STREAM_NORMAL, STREAM_UNIFORM, STREAM_RESAMPLE, STREAM_SAMPLE = 0, 1, 2, 3


def philox_call(seed, call, t, stream):
    """-> (a, b): the two 64-bit output words of call `call` of (seed, t, stream)."""
    c0, c1, c2, c3 = call & MASK, (call >> 32) & MASK, t & MASK, stream & MASK
    k0, k1 = seed & MASK, (seed >> 32) & MASK
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & MASK, p1 & MASK, ((p0 >> 32) ^ c3 ^ k1) & MASK, p0 & MASK
        k0, k1 = (k0 + W0) & MASK, (k1 + W1) & MASK
    return c0 | (c1 << 32), c2 | (c3 << 32)


def uniform(seed, element, t, stream):
    """Element `element` of the uniform array of (seed, t, stream): call element>>1, word a (even) or b (odd),
    top 53 bits, in [0, 1)."""
    a, b = philox_call(seed, element >> 1, t, stream)
    return float((b if element & 1 else a) >> 11) * 2.0 ** -53
