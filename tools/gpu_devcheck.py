#!/usr/bin/env python
"""Developer GPU check (run under gpurun): exercises the radix sort and the SA builder
through the C ABI and compares against numpy / the compiled reference libsais
(oracle/_ref).  Prints a compact report; writes gpurun_out/devcheck.json."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = C.CDLL(os.path.join(ROOT, "pysubstringsearch_b200", "libpss_b200.so"))
lib.pss_last_error.restype = C.c_char_p
ref = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libsais_ref.so"))
ref.libsais.restype = C.c_int32
ref.libsais.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]

report = {"radix": [], "sa": [], "perf": []}
ok_all = True


def err():
    return lib.pss_last_error().decode()


def ref_sa(text: np.ndarray) -> np.ndarray:
    sa = np.empty(len(text), dtype=np.int32)
    rc = ref.libsais(text.ctypes.data, sa.ctypes.data, len(text), 0, None)
    assert rc == 0
    return sa


def check_radix(n, begin, end, iota, seed, kind="random"):
    global ok_all
    rng = np.random.default_rng(seed)
    if kind == "random":
        keys = rng.integers(0, 2**63, size=n, dtype=np.uint64) * 2 + rng.integers(0, 2, size=n, dtype=np.uint64)
    elif kind == "few":
        keys = rng.integers(0, 3, size=n, dtype=np.uint64) << np.uint64(begin)
    elif kind == "const":
        keys = np.full(n, 0x0123456789ABCDEF, dtype=np.uint64)
    vals = np.arange(n, dtype=np.uint32) if iota else rng.integers(0, 2**32, size=n, dtype=np.uint32)
    dk = torch.from_numpy(keys.view(np.int64)).cuda()
    dka = torch.empty_like(dk)
    dv = torch.from_numpy(vals.view(np.int32)).cuda()
    dva = torch.empty_like(dv)
    in_alt = C.c_int32(0)
    npass = C.c_int32(0)
    ms = (C.c_float * 8)()
    rc = lib.pss_radix_sort_pairs(C.c_void_p(dk.data_ptr()), C.c_void_p(dka.data_ptr()),
                                  C.c_void_p(0 if iota else dv.data_ptr()), C.c_void_p(dva.data_ptr()),
                                  C.c_int64(n), begin, end, C.byref(in_alt), ms, C.byref(npass), None)
    torch.cuda.synchronize()
    if rc != 0:
        print("radix rc", rc, err())
        ok_all = False
        return
    gk = (dka if in_alt.value else dk).cpu().numpy().view(np.uint64)
    gv = (dva if (in_alt.value or iota) else dv).cpu().numpy().view(np.uint32)
    width = end - begin
    mask = np.uint64((1 << width) - 1) if width < 64 else np.uint64(0xFFFFFFFFFFFFFFFF)
    sub = (keys >> np.uint64(begin)) & mask
    order = np.argsort(sub, kind="stable")
    good = np.array_equal(gk, keys[order]) and np.array_equal(gv, vals[order])
    ok_all &= bool(good)
    rec = dict(n=n, bits=[begin, end], iota=iota, kind=kind, ok=bool(good), passes=npass.value,
               ms=[round(ms[i], 4) for i in range(npass.value)])
    report["radix"].append(rec)
    print("radix", rec)


class PassStat(C.Structure):
    _fields_ = [("round", C.c_int32), ("pass_", C.c_int32), ("shift", C.c_int32), ("reserved", C.c_int32),
                ("n_records", C.c_int64), ("ms", C.c_float), ("reserved2", C.c_float)]


class BuildStats(C.Structure):
    _fields_ = [("n", C.c_int32), ("sigma", C.c_int32), ("bits_per_symbol", C.c_int32), ("h0", C.c_int32),
                ("rounds", C.c_int32), ("n_passes", C.c_int32), ("n_pass_stats", C.c_int32),
                ("n_kernel_launches", C.c_int32), ("active_per_round", C.c_int64 * 64),
                ("total_ms", C.c_float), ("sort_ms", C.c_float), ("records_sorted", C.c_int64)]


def make_text(kind, n, seed):
    rng = np.random.default_rng(seed)
    if kind == "words":
        vocab = ["".join(chr(97 + c) for c in rng.integers(0, 26, size=rng.integers(3, 11))) for _ in range(4096)]
        p = 1.0 / np.arange(1, len(vocab) + 1) ** 1.1
        p /= p.sum()
        out = bytearray()
        ids = rng.choice(len(vocab), size=n // 5 + 16, p=p)
        seps = rng.random(len(ids)) < (1 / 6)
        for w, s in zip(ids, seps):
            out += vocab[w].encode()
            out += b"\n" if s else b" "
            if len(out) >= n:
                break
        t = np.frombuffer(bytes(out[:n]), dtype=np.uint8).copy()
        t[-1] = 10
        return t
    if kind == "bin":
        return rng.integers(0, 256, size=n, dtype=np.uint8)
    if kind == "tiny":
        return rng.choice(np.array([0, 10, 97, 98, 255], dtype=np.uint8), size=n)
    if kind == "acgt":
        base = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=max(1, n // 8))
        t = np.tile(base, 9)[:n].copy()
        t[::1000] = 10
        return t
    if kind == "same":
        return np.full(n, 97, dtype=np.uint8)
    if kind == "period":
        t = np.tile(np.frombuffer(b"ACGT", dtype=np.uint8), n // 4 + 1)[:n].copy()
        t[-1] = 10
        return t
    raise ValueError(kind)


def check_sa(kind, n, seed, builder, profile=False):
    global ok_all
    text = make_text(kind, n, seed)
    n = len(text)
    want = ref_sa(text)
    got = np.empty(n, dtype=np.int32)
    lib.pss_sa_builder_set_profiling(builder, 1 if profile else 0)
    t0 = time.time()
    rc = lib.pss_sa_builder_build_host(builder, C.c_void_p(text.ctypes.data), n, C.c_void_p(got.ctypes.data))
    dt = time.time() - t0
    if rc != 0:
        print("sa rc", rc, err())
        ok_all = False
        report["sa"].append(dict(kind=kind, n=n, ok=False, err=err()))
        return
    st = BuildStats()
    ps = (PassStat * 512)()
    lib.pss_sa_builder_stats(builder, C.byref(st), ps)
    good = np.array_equal(want, got)
    ok_all &= bool(good)
    rec = dict(kind=kind, n=n, ok=bool(good), sigma=st.sigma, b=st.bits_per_symbol, h0=st.h0, rounds=st.rounds,
               passes=st.n_passes, launches=st.n_kernel_launches, total_ms=round(st.total_ms, 3),
               host_s=round(dt, 4), active=[int(st.active_per_round[i]) for i in range(st.rounds + 1)])
    if not good:
        bad = np.nonzero(want != got)[0]
        rec["first_bad"] = int(bad[0])
        rec["n_bad"] = int(len(bad))
        rec["want"] = want[bad[0]:bad[0] + 8].tolist()
        rec["got"] = got[bad[0]:bad[0] + 8].tolist()
    if profile:
        rec["pass_ms"] = [(ps[i].round, ps[i].shift, int(ps[i].n_records), round(ps[i].ms, 4),
                           round(24.0 * ps[i].n_records / (ps[i].ms * 1e-3) / 1e9, 1) if ps[i].ms > 0 else 0)
                          for i in range(st.n_pass_stats)]
    report["sa"].append(rec)
    print("sa", json.dumps(rec))


def main():
    print("devices", lib.pss_device_count(), torch.cuda.get_device_name(0))
    # --- radix sort -------------------------------------------------------------
    for n in [1, 2, 31, 32, 33, 4095, 4096, 4097, 100000, 1 << 20, (1 << 22) + 12345]:
        check_radix(n, 0, 64, False, n)
    check_radix(1 << 20, 0, 64, True, 1)
    check_radix(1 << 20, 5, 37, False, 2)
    check_radix(1 << 20, 0, 59, True, 3)
    check_radix(1 << 20, 8, 24, False, 4, kind="few")
    check_radix(1 << 20, 0, 64, True, 5, kind="const")
    check_radix(1 << 26, 0, 64, False, 6)
    # --- SA builder ----------------------------------------------------------------
    builder = C.c_void_p()
    rc = lib.pss_sa_builder_create(-1, C.c_int64(1 << 20), C.byref(builder))
    assert rc == 0, err()
    for kind in ["tiny", "words", "bin", "acgt", "same", "period"]:
        for n in [2, 3, 5, 17, 100, 1000, 4096, 4097, 65536, 1 << 20]:
            if kind in ("same", "period") and n > 65536:
                continue
            check_sa(kind, n, n + 7, builder)
    for seed in range(200):
        n = int(np.random.default_rng(seed).integers(2, 300))
        check_sa("tiny", n, 1000 + seed, builder)
    check_sa("words", 1 << 24, 11, builder, profile=True)
    check_sa("words", 1 << 26, 12, builder, profile=True)
    check_sa("acgt", 1 << 24, 13, builder, profile=True)
    check_sa("bin", 1 << 24, 14, builder, profile=True)
    check_sa("period", 1 << 20, 15, builder, profile=True)
    lib.pss_sa_builder_destroy(builder)
    report["ok"] = bool(ok_all)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "devcheck.json"), "w") as f:
        json.dump(report, f, indent=1)
    n_sa_bad = sum(1 for r in report["sa"] if not r["ok"])
    n_rx_bad = sum(1 for r in report["radix"] if not r["ok"])
    print("SUMMARY ok=%s radix_bad=%d sa_bad=%d" % (ok_all, n_rx_bad, n_sa_bad))
    return 0 if ok_all else 1


if __name__ == "__main__":
    sys.exit(main())
