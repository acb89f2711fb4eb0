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


def batch_uniform(seed, first_pair, n, len1=150, len2=150, out=None, chunk=65536):
    """batch(seed, first_pair, n, 0, len1, len2) for large n: the same bytes, vectorised over pairs (the per-pair loop
    of batch() takes ~20 us per pair).  `out` may be a preallocated uint8 array of n * (len1 + len2) (e.g. pinned)."""
    tot = len1 + len2
    bases = out if out is not None else np.empty(n * tot, dtype=np.uint8)
    view = bases[:n * tot].reshape(n, tot)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    shifts = (np.uint64(2) * np.arange(32, dtype=np.uint64))[None, None, :]
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        p = np.arange(first_pair + lo, first_pair + hi, dtype=np.uint64)
        for which, (L, col) in enumerate(((len1, 0), (len2, len1))):
            if L == 0:
                continue
            nw = (L + 31) // 32
            with np.errstate(over="ignore"):
                words = splitmix64(key(seed, p, which)[:, None] + np.arange(nw, dtype=np.uint64)[None, :])
            codes = ((words[:, :, None] >> shifts) & np.uint64(3)).astype(np.uint8).reshape(hi - lo, nw * 32)[:, :L]
            view[lo:hi, col:col + L] = acgt[codes]
    off1 = np.arange(n, dtype=np.uint64) * np.uint64(tot)
    off2 = off1 + np.uint64(len1)
    return bases, off1, off2, np.full(n, len1, dtype=np.uint32), np.full(n, len2, dtype=np.uint32)


RELATED_SALT = 0x52454C41544544  # "RELATED"


def related_sequence(seed, pair, length):
    """Sequence 2 of the RELATED variant of config 4: sequence 1 of the pair with 10 % substitutions,
    2 % insertions and 2 % deletions, every draw from the same counter-based generator.  Exact procedure, position by
    position over sequence 1 (independent draws, so any position can be regenerated on its own):

        r_i = splitmix64((key(seed, pair, 1) ^ RELATED_SALT) + i)           one draw per position i of sequence 1
        u   = r_i % 100
        u <  2        deletion:      emit nothing
        2 <= u <  4   insertion:     emit seq1[i], then the base "ACGT"[(r_i >> 32) & 3]
        4 <= u < 14   substitution:  emit "ACGT"[(code(seq1[i]) + 1 + (r_i >> 32) % 3) & 3]   (always a different base)
        otherwise     copy:          emit seq1[i]

    -> (seq1 bytes, seq2 bytes); len(seq2) is length * (1 + 0.02 - 0.02) on average."""
    s1 = sequence(seed, pair, 0, length)
    code = np.zeros(256, dtype=np.uint8)
    code[np.frombuffer(b"ACGT", dtype=np.uint8)] = np.arange(4, dtype=np.uint8)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    with np.errstate(over="ignore"):
        r = splitmix64((key(seed, pair, 1) ^ np.uint64(RELATED_SALT)) + np.arange(length, dtype=np.uint64))
    u = (r % np.uint64(100)).astype(np.int64)
    hi = (r >> np.uint64(32))
    first = s1.copy()
    sub = (u >= 4) & (u < 14)
    first[sub] = acgt[(code[s1[sub]].astype(np.uint64) + np.uint64(1) + hi[sub] % np.uint64(3)) & np.uint64(3)]
    second = acgt[(hi & np.uint64(3)).astype(np.int64)]
    count = np.ones(length, dtype=np.int64)
    count[u < 2] = 0
    count[(u >= 2) & (u < 4)] = 2
    out = np.repeat(first, count)
    # the second symbol of every insertion is the inserted base
    ends = np.cumsum(count) - 1
    ins = np.nonzero(count == 2)[0]
    out[ends[ins]] = second[ins]
    return s1, out


def related_batch(seed, first_pair, n, length):
    """n related pairs (see related_sequence) in the C-ABI batch layout."""
    seqs = [related_sequence(seed, first_pair + p, length) for p in range(n)]
    l1 = np.array([len(a) for a, _ in seqs], dtype=np.uint32)
    l2 = np.array([len(b) for _, b in seqs], dtype=np.uint32)
    tot = l1.astype(np.uint64) + l2.astype(np.uint64)
    off1 = np.zeros(n, dtype=np.uint64)
    if n > 1:
        off1[1:] = np.cumsum(tot)[:-1]
    off2 = off1 + l1.astype(np.uint64)
    bases = np.zeros(int(tot.sum()) if n else 0, dtype=np.uint8)
    for p, (a, b) in enumerate(seqs):
        bases[int(off1[p]):int(off1[p]) + len(a)] = a
        bases[int(off2[p]):int(off2[p]) + len(b)] = b
    return bases, off1, off2, l1, l2
