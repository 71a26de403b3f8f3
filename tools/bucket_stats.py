#!/usr/bin/env python
"""Planning data for bucketed doubling rounds ("What a next round should build"): runs the
numpy model of the builder (tests/test_algorithm_model.py) on a config-1-like text and reports, per
doubling round, how the active records fall into 2^16 buckets of SA slots (bucket = group rank >> (gbits - 16)).
A group's records stay inside [g, g + size), so a bucket holds at most its width W = n / 2^16 plus the
overhang of its last group; the table shows how many records sit in buckets a CTA could sort in shared
memory (<= 1.125 W records) and how many are left to the global sorter.  CPU only.
usage: bucket_stats.py [n_bytes=2^26] [kind=words|acgt]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools import synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 26
kind = sys.argv[2] if len(sys.argv) > 2 else "words"
t = (synth.zipf_words_text(n, seed=3) if kind == "words" else synth.acgt_text(n)).astype(np.uint8)
W = max(1, n >> 16)
cap = W + W // 8
present = np.unique(t)
lut = np.zeros(256, dtype=np.uint64)
lut[present] = np.arange(1, len(present) + 1, dtype=np.uint64)
b = int(len(present)).bit_length()
m = min(64 // b, n)
code = np.concatenate([lut[t], np.zeros(m, dtype=np.uint64)])
key = np.zeros(n, dtype=np.uint64)
for j in range(m):
    key = (key << np.uint64(b)) | code[j:j + n]
isa = np.zeros(n, dtype=np.int64)
idx = np.arange(n, dtype=np.int64)
grp = np.zeros(n, dtype=np.int64)
first, h, rnd = True, m, 0
print("n = %d, sigma = %d, %d bits/symbol, h0 = %d, bucket width W = %d SA slots, CTA capacity %d records" % (n, len(present), b, m, W, cap))
print("| round | active | groups | largest group | largest bucket | buckets > capacity | records in them |")
print("|---|---|---|---|---|---|---|")
while len(idx):
    if not first:
        j = idx + h
        r2 = np.where(j < n, isa[np.minimum(j, n - 1)], 0)
        # ---- the statistic: bucket occupancy of this round's active set ----
        cnt = np.bincount(grp // W, minlength=(n + W - 1) // W)
        gsz = np.bincount(grp)
        over = cnt > cap
        print("| %d | %d (%.1f %%) | %d | %d | %d | %d of %d | %d (%.1f %%) |" % (
            rnd, len(idx), 100.0 * len(idx) / n, int((gsz > 0).sum()), int(gsz.max()), int(cnt.max()),
            int(over.sum()), int((cnt > 0).sum()), int(cnt[over].sum()), 100.0 * cnt[over].sum() / len(idx)))
        del cnt, gsz
        key = (grp.astype(np.uint64) << np.uint64(32)) | r2.astype(np.uint64)
    order = np.argsort(key, kind="stable")
    key, idx = key[order], idx[order]
    k = np.arange(len(idx))
    g = np.zeros(len(idx), dtype=np.int64) if first else (key >> np.uint64(32)).astype(np.int64)
    hn = np.ones(len(idx), dtype=bool)
    hn[1:] = key[1:] != key[:-1]
    ho = np.zeros(len(idx), dtype=bool)
    ho[0] = True
    if not first:
        ho[1:] = g[1:] != g[:-1]
    A = np.maximum.accumulate(np.where(ho, k, 0))
    B = np.maximum.accumulate(np.where(hn, k, 0))
    p = g + (k - A)
    ng = g + (B - A)
    nhn = np.ones(len(idx), dtype=bool)
    nhn[:-1] = hn[1:]
    single = hn & nhn
    isa[idx] = np.where(single, p, ng) + 1
    idx, grp = idx[~single], ng[~single]
    if not first:
        h *= 2
    first = False
    rnd += 1
