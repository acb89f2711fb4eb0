"""The shared synthetic-input generator in numpy, bit-identical to generate_kernel in
csrc/seqa_util.cuh: i.i.d. uniform DNA from a counter-based splitmix64 stream, random access per pair."""
import numpy as np

SEED = 20240607
_M = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(z):
    z = np.asarray(z, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = z + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def key(seed, pair, which):
    with np.errstate(over="ignore"):
        return splitmix64(np.uint64(seed) ^ (np.uint64(2) * np.asarray(pair, dtype=np.uint64) + np.uint64(which)))


def lengths(seed, first_pair, n):
    p = np.arange(first_pair, first_pair + n, dtype=np.uint64)
    l1 = 50 + (splitmix64(key(seed, p, 0) ^ np.uint64(0xC0FFEE)) % np.uint64(951))
    l2 = 50 + (splitmix64(key(seed, p, 1) ^ np.uint64(0xC0FFEE)) % np.uint64(951))
    return l1.astype(np.uint32), l2.astype(np.uint32)


def sequence(seed, pair, which, length):
    """bytes of sequence `which` (0/1) of pair `pair`."""
    nw = (length + 31) // 32
    with np.errstate(over="ignore"):
        words = splitmix64(key(seed, pair, which) + np.arange(nw, dtype=np.uint64))
    shifts = (np.uint64(2) * np.arange(32, dtype=np.uint64))[None, :]
    codes = ((words[:, None] >> shifts) & np.uint64(3)).astype(np.uint8).reshape(-1)[:length]
    return np.frombuffer(b"ACGT", dtype=np.uint8)[codes]


def batch(seed, first_pair, n, len_mode=0, len1=150, len2=150):
    """-> (bases u8, off1 u64, off2 u64, len1 u32, len2 u32) in the C-ABI batch layout (seq1 then seq2 per pair)."""
    if len_mode:
        l1, l2 = lengths(seed, first_pair, n)
    else:
        l1 = np.full(n, len1, dtype=np.uint32)
        l2 = np.full(n, len2, dtype=np.uint32)
    tot = l1.astype(np.uint64) + l2.astype(np.uint64)
    off1 = np.zeros(n, dtype=np.uint64)
    if n > 1:
        off1[1:] = np.cumsum(tot)[:-1]
    off2 = off1 + l1.astype(np.uint64)
    bases = np.zeros(int(tot.sum()) if n else 0, dtype=np.uint8)
    for p in range(n):
        bases[int(off1[p]):int(off1[p]) + int(l1[p])] = sequence(seed, first_pair + p, 0, int(l1[p]))
        bases[int(off2[p]):int(off2[p]) + int(l2[p])] = sequence(seed, first_pair + p, 1, int(l2[p]))
    return bases, off1, off2, l1, l2
