"""Executable specification (numpy, CPU) of the GPU builder's algorithm — the same key packing,
active-set filtering and re-rank arithmetic as csrc/sa_build.cu, without any CUDA — checked
against the oracle.  It documents WHY the kernels are right; the kernels themselves are
checked bit-for-bit in test_gpu_parity.py."""
import numpy as np


def model_suffix_array(text, key_bits=64):
    t = np.frombuffer(bytes(text), dtype=np.uint8)
    n = len(t)
    if n == 0:
        return np.zeros(0, np.int32)
    if n == 1:
        return np.zeros(1, np.int32)
    # alphabet -> dense codes 1..sigma, 0 = past the end (presence_kernel + LUT)
    present = np.unique(t)
    lut = np.zeros(256, dtype=np.uint64)
    lut[present] = np.arange(1, len(present) + 1, dtype=np.uint64)
    b = int(len(present)).bit_length()
    m = min(key_bits // b, n)
    code = np.concatenate([lut[t], np.zeros(m, dtype=np.uint64)])
    # round 0 key: m codes packed big-endian (keygen_kernel)
    key = np.zeros(n, dtype=np.uint64)
    for j in range(m):
        key = (key << np.uint64(b)) | code[j:j + n]
    sa = np.full(n, -1, dtype=np.int64)
    isa = np.zeros(n, dtype=np.int64)          # 1-based rank, 0 = past the end
    idx = np.arange(n, dtype=np.int64)
    grp = np.zeros(n, dtype=np.int64)
    first, h = True, m
    while len(idx):
        if not first:                          # gather_kernel: (group rank, rank of suffix i + h)
            j = idx + h
            r2 = np.where(j < n, isa[np.minimum(j, n - 1)], 0)
            key = (grp.astype(np.uint64) << np.uint64(32)) | r2.astype(np.uint64)
        order = np.argsort(key, kind="stable")  # the onesweep sort
        key, idx = key[order], idx[order]
        k = np.arange(len(idx))
        g = np.zeros(len(idx), dtype=np.int64) if first else (key >> np.uint64(32)).astype(np.int64)
        hn = np.ones(len(idx), dtype=bool)
        hn[1:] = key[1:] != key[:-1]                       # new-group heads
        ho = np.zeros(len(idx), dtype=bool)
        ho[0] = True
        if not first:
            ho[1:] = g[1:] != g[:-1]                       # old-group heads
        A = np.maximum.accumulate(np.where(ho, k, 0))      # the 3-component scan of rerank_apply
        B = np.maximum.accumulate(np.where(hn, k, 0))
        p = g + (k - A)                                    # final SA slot inside the old group
        ng = g + (B - A)                                   # rank of the new group
        nhn = np.ones(len(idx), dtype=bool)
        nhn[:-1] = hn[1:]
        single = hn & nhn
        sa[p[single]] = idx[single]                        # settled suffixes leave the active set
        isa[idx] = np.where(single, p, ng) + 1
        idx, grp = idx[~single], ng[~single]
        if not first:
            h *= 2                                         # round 0 sorted m symbols: round 1 looks at i + m
        first = False
        assert h <= 4 * n + 64, "prefix doubling failed to converge"
    return sa.astype(np.int32)


def test_model_matches_oracle(oracle):
    rng = np.random.default_rng(3)
    cases = [b"ab", b"banana", b"mississippi\n", b"aaaaaaaaaaaaaaaa", b"x\x00x", b"\n\n\n", b"abab" * 40 + b"\n"]
    for _ in range(60):
        n = int(rng.integers(2, 300))
        cases.append(bytes(rng.choice(np.array([0, 10, 97, 98, 255], dtype=np.uint8), size=n)))
    for _ in range(10):
        cases.append(bytes(rng.integers(0, 256, size=int(rng.integers(2, 500)), dtype=np.uint8)))
    for t in cases:
        assert model_suffix_array(t).tolist() == oracle.suffix_array_port(t).tolist(), t[:40]


def test_model_with_narrow_keys_needs_more_rounds_but_same_answer(oracle):
    """The packed-prefix width only changes the number of doubling rounds, never the result
    (the PSS_H0 knob)."""
    t = (b"the quick brown fox jumps over the lazy dog\n" * 7) + b"the quick brown cat\n"
    want = oracle.suffix_array_port(t).tolist()
    for bits in (64, 16, 5):
        assert model_suffix_array(t, key_bits=bits).tolist() == want
