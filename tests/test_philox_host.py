"""CPU: the host-side Philox (gen_b200/philox.py, used for the chunk-merge draw of importance_resampling) against the
oracle's C implementation and the Random123 known-answer vectors; chunk seeds."""
import numpy as np

from gen_b200 import philox
from gen_b200.inference import CHUNK_EVENT, chunk_seed


def test_random123_known_answers():
    # Random123 kat_vectors, philox4x32-10: counter/key all zero and all ones
    def raw(ctr, key):
        call = ctr[0] | (ctr[1] << 32)
        seed = key[0] | (key[1] << 32)
        a, b = philox.philox_call(seed, call, ctr[2], ctr[3])
        return [a & 0xFFFFFFFF, a >> 32, b & 0xFFFFFFFF, b >> 32]
    assert raw([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    f = 0xFFFFFFFF
    assert raw([f, f, f, f], [f, f]) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert raw([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_host_uniforms_match_the_oracle(orc):
    for seed, t, stream in ((0, 0, 1), (11, CHUNK_EVENT, philox.STREAM_SAMPLE), (2 ** 63 + 5, 77, 3)):
        want = orc.uniforms(seed, t, stream, 0, 64)
        got = np.array([philox.uniform(seed, e, t, stream) for e in range(64)])
        assert np.array_equal(want, got)


def test_chunk_seeds():
    assert chunk_seed(5, 0) == 5
    s = {chunk_seed(5, c) for c in range(1000)}
    assert len(s) == 1000 and all(0 <= v < 2 ** 64 for v in s)
