"""Small end-to-end run for compute-sanitizer: builds a few suffix arrays and runs searches
through both search paths, checking against the oracle."""
import os, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pysubstringsearch_b200 import capi as pss
from oracle import oracle as O
from tools import synth
for n, kind in ((50_000, "w"), (300_000, "w"), (70_000, "a"), (5_000_000, "w")):
    t = synth.zipf_words_text(n, seed=n, vocab=512, block=1 << 14) if kind == "w" else synth.acgt_text(n, base_len=4096, mut_every=1024)
    assert np.array_equal(pss.libsais(t), O.suffix_array_port(t)), n
text = synth.zipf_words_text(400_000, seed=3, vocab=512, block=1 << 14)
with tempfile.TemporaryDirectory() as d:
    p = os.path.join(d, "s.idx")
    w = pss.Writer(p, 1 << 17)
    for e in bytes(text).split(b"\n")[:-1]:
        w.add_entry(e)
    w.close()
    o = O.Reader(p)
    batches = ([b"ab"], [b""], [b"e ", b"zz", b"\n"], [bytes(text[k:k + 7]) for k in range(0, 40000, 400)])
    want = [o.search_multiple_tuples(pats) for pats in batches]
    # every geometry of the bounds kernel (a warp per pair, 8 / 4 lanes per pair, with and without the
    # SA look-ahead) and every extraction mode (line records, line directory, text scans)
    for env in ({}, {"PSS_BOUNDS_GROUP": "-4"}, {"PSS_BOUNDS_GROUP": "8"}, {"PSS_BOUNDS_GROUP": "-32"},
                {"PSS_LINE_DIR": "1"}, {"PSS_LINE_DIR": "0"}, {"PSS_SMALL_PATH": "0"}):
        for k in ("PSS_BOUNDS_GROUP", "PSS_LINE_DIR", "PSS_SMALL_PATH"):
            os.environ.pop(k, None)
        os.environ.update(env)
        r = pss.Reader(p)
        for pats, (c, och, ost, oen) in zip(batches, want):
            qo, ch, st, en, _ = r.search_batch(pats)
            assert np.array_equal(ch, och) and np.array_equal(st, ost) and np.array_equal(en, oen), env
        r.close()
print("sanitize_check ok")
