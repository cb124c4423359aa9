import numpy as np

from oracle import synth


def test_philox_known_answer():
    # Random123 kat_vectors: philox4x32-10, counter 0, key 0
    o = synth.philox4x32_10(np.array([0]), 0, 0)
    assert [int(v[0]) for v in o] == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]


def test_ih8_moments_and_shard_consistency():
    A = synth.make_A(4000, 64)
    assert A.flags["F_CONTIGUOUS"]
    v = A * np.sqrt(64.0)
    assert abs(v.mean()) < 0.01 and abs(v.std() - 1.0) < 0.01
    # a row shard generated on its own equals the same rows of the full matrix (counter = i + j*n_total)
    B = synth.make_A(1000, 64, row0=1500, n_total=4000)
    assert np.array_equal(B, A[1500:2500])
    assert float(synth.IH8_INV_STD).hex() == "0x1.3988e1412ed76p-16"


def test_density_mask():
    A = synth.make_A(2000, 50, density=0.01)
    frac = np.mean(A != 0)
    assert 0.005 < frac < 0.02
