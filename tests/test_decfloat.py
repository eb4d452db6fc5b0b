"""Number parsing of the GPU Matrix Market ingest (cfs_spmv_b200/csrc/
decfloat.cuh), host build, against the functions the reference's loader calls:
atof -> strtod (correctly rounded; Python's float() is the same function) and
atoi -> (int)strtol (reference include/io/mmf.hpp:309-343)."""
import ctypes
import os
import random
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "decfloat_host.cpp")
OUT = os.path.join(ROOT, "build", "test", "libdecfloat_host.so")


@pytest.fixture(scope="module")
def lib():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-x", "c++",
                           "-I" + os.path.join(ROOT, "cfs_spmv_b200", "csrc"),
                           SRC, "-o", OUT])
    return ctypes.CDLL(OUT)


def parse_doubles(lib, tokens):
    blob = "".join(tokens).encode()
    off = np.zeros(len(tokens) + 1, np.int64)
    np.cumsum([len(t.encode()) for t in tokens], out=off[1:])
    out = np.zeros(len(tokens), np.float64)
    status = np.zeros(len(tokens), np.int32)
    lib.cfs_test_parse_doubles(blob, off.ctypes.data_as(ctypes.c_void_p),
                               ctypes.c_long(len(tokens)),
                               out.ctypes.data_as(ctypes.c_void_p),
                               status.ctypes.data_as(ctypes.c_void_p))
    return out, status


def strtod(libc, token):
    libc.strtod.restype = ctypes.c_double
    return libc.strtod(token.encode(), None)


def bits(a):
    return np.asarray(a, np.float64).view(np.uint64)


def random_tokens(rng, n):
    toks = []
    for _ in range(n):
        kind = rng.randrange(8)
        if kind == 0:   # shortest round-trip of a random double
            v = np.frombuffer(rng.getrandbits(64).to_bytes(8, "little"),
                              np.float64)[0]
            if not np.isfinite(v):
                v = 1.0
            toks.append(repr(float(v)))
        elif kind == 1:  # %.17g / %.16e style, as Matrix Market writers emit
            v = rng.uniform(-1e3, 1e3) * 10.0 ** rng.randrange(-30, 30)
            toks.append(("%.17g" if rng.random() < 0.5 else "%.16e") % v)
        elif kind == 2:  # short decimals
            toks.append("%.*f" % (rng.randrange(0, 8), rng.uniform(-100, 100)))
        elif kind == 3:  # integers, signs, leading zeros
            toks.append(rng.choice(["", "+", "-"]) + "0" * rng.randrange(3) +
                        str(rng.randrange(10 ** rng.randrange(1, 19))))
        elif kind == 4:  # 19 digits exactly, random exponent: Eisel-Lemire
            toks.append("%d.%018de%d" % (rng.randrange(1, 10),
                                         rng.randrange(10 ** 18),
                                         rng.randrange(-330, 300)))
        elif kind == 5:  # more than 19 digits
            toks.append("0.%s%se%d" % ("0" * rng.randrange(4),
                                       "".join(rng.choice("0123456789")
                                               for _ in range(rng.randrange(20, 40))),
                                       rng.randrange(-300, 300)))
        elif kind == 6:  # subnormal / overflow borders
            toks.append("%d.%de%d" % (rng.randrange(1, 10), rng.randrange(10 ** 6),
                                      rng.choice([-325, -324, -323, -320, -310,
                                                  -308, -307, 307, 308, 309])))
        else:            # near halfway cases: 2^53 + odd, scaled
            m = (1 << 53) + 2 * rng.randrange(1 << 20) + 1
            toks.append("%d%s" % (m, rng.choice(["", "e0", "0", "e3", "e-3",
                                                 ".5", ".50000", ".4999999",
                                                 ".5000001"])))
    return toks


def test_parse_double_matches_strtod(lib):
    rng = random.Random(12345)
    toks = random_tokens(rng, 400000)
    out, status = parse_doubles(lib, toks)
    want = np.array([float(t) for t in toks])
    decided = status == 0
    # whatever the parser decides is the correctly rounded value, bit for bit
    assert np.array_equal(bits(out[decided]), bits(want[decided]))
    # ... and it decides nearly everything (the rest goes to the host's strtod)
    assert decided.mean() > 0.97
    short = np.array([len(t.lstrip("+-0.").split("e")[0].replace(".", "")) <= 19
                      for t in toks])
    assert decided[short].mean() > 0.9999


def test_parse_double_grammar(lib):
    libc = ctypes.CDLL("libc.so.6")
    cases = ["1", "-1", "+1", "1.", ".5", "-.5e1", "1e", "1e+", "1e5x", "1.5abc",
             "  2.5", "\t3", "1\r", "-0", "-0.0", "0e9999", "1e-9999", "1e9999",
             "123456789012345678901234567890", "0.000000000000000000000001",
             "4.9e-324", "2.4703282292062327e-324", "2.4703282292062328e-324",
             "1.7976931348623157e308", "1.7976931348623159e308",
             "9007199254740993", "9007199254740992.5", "1E5", "5e-1", "00012",
             "0x10", "0X1p3", "inf", "-inf", "nan", "infinity", ".", "", "-",
             "e5", "abc", "+.", "1.e2", "1..2", "1e5.5", "--1"]
    out, status = parse_doubles(lib, cases)
    for t, v, s in zip(cases, out, status):
        if s == 0:
            w = strtod(libc, t)
            assert bits(v) == bits(w), (t, v, w)
    need_host = {t for t, s in zip(cases, status) if s != 0}
    # outside the decimal grammar: reported, never guessed
    assert {"0x10", "0X1p3", "inf", "-inf", "nan", "infinity", ".", "", "-",
            "e5", "abc", "+.", "--1"} <= need_host
    assert not ({"1", "1e", "1e5x", "1.5abc", "1\r", "1e9999", "4.9e-324",
                 "9007199254740993", "00012"} & need_host)


def test_parse_int_matches_atoi(lib):
    libc = ctypes.CDLL("libc.so.6")
    rng = random.Random(7)
    toks = ["0", "1", "-1", "+7", "  42", "\t9", "12abc", "abc", "", "-",
            "2147483647", "2147483648", "4294967297", "-2147483649",
            "999999999999999999", "1234567890123456789", "007", "1.9", "3e5"]
    toks += [str(rng.randrange(-2 ** 40, 2 ** 40)) for _ in range(20000)]
    blob = "".join(toks).encode()
    off = np.zeros(len(toks) + 1, np.int64)
    np.cumsum([len(t) for t in toks], out=off[1:])
    out = np.zeros(len(toks), np.int32)
    status = np.zeros(len(toks), np.int32)
    lib.cfs_test_parse_ints(blob, off.ctypes.data_as(ctypes.c_void_p),
                            ctypes.c_long(len(toks)),
                            out.ctypes.data_as(ctypes.c_void_p),
                            status.ctypes.data_as(ctypes.c_void_p))
    for t, v, s in zip(toks, out, status):
        if s == 0:
            assert int(v) == libc.atoi(t.encode()), t
        else:
            assert len(t.strip().lstrip("+-")) > 18


def test_split_line(lib):
    def ref_split(line):  # src/mmf.cpp:6-44 of the reference, restated
        return [t for t in line.strip(" \t").split(" ") if t]

    lines = ["1 2 3.5", "  1 2 3.5  ", "\t1 2\t", "1  2   3", "1 2", "1", "",
             "   ", "1\t2 3", "1 2 3 4 5", " \t 7 8 9e1 \t "]
    for line in lines:
        b = (ctypes.c_long * 6)()
        n = lib.cfs_test_split_line(line.encode(), ctypes.c_long(len(line)), b)
        want = ref_split(line)
        assert n == len(want), line
        got = [line[b[2 * k]:b[2 * k + 1]] for k in range(min(n, 3))]
        assert got == want[:3], line
