"""Shared helpers for the parity tests: seeded synthetic DNA and result comparison."""
import numpy as np

BASES = np.array([1, 2, 4, 8], dtype=np.uint8)
IUPAC_AMBIG = np.array([3, 5, 9, 6, 10, 12, 7, 11, 13, 14, 15], dtype=np.uint8)


def rng(seed):
    return np.random.Generator(np.random.PCG64(seed))


def random_dna(g, n, gc=0.41):
    p = np.array([(1 - gc) / 2, gc / 2, gc / 2, (1 - gc) / 2])
    return BASES[g.choice(4, size=n, p=p)]


def ends_as_lists(x):
    """oracle MIndex.endIndex() / product ByPos_MIndex.endIndex() -> list of lists."""
    return [None if e is None else np.asarray(e).tolist() for e in x]


def assert_same_ends(got, want):
    got, want = ends_as_lists(got), ends_as_lists(want)
    assert len(got) == len(want)
    for i, (a, b) in enumerate(zip(got, want)):
        assert a == b, f"pattern {i + 1}: got {a}, want {b}"
