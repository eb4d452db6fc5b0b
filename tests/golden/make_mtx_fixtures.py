"""Writes the small Matrix Market fixtures under tests/golden/mtx/ that probe
the loader semantics of the reference (include/io/mmf.hpp, src/mmf.cpp;
SURVEY.md appendix A.1). Deterministic; run before make_golden.py."""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "mtx")
N = 40
B = "%" + "%"  # banner prefix, kept out of %-formatting


def lower_entries(seed, n=N, extra=60):
    rng = np.random.default_rng(seed)
    ent = {(i, i): 4.0 + (i % 7) * 0.25 for i in range(n)}
    while len(ent) < n + extra:
        i = int(rng.integers(1, n))
        j = int(rng.integers(0, i))
        ent[(i, j)] = round(float(rng.uniform(-1, 1)), 6)
    return sorted(ent.items())


def main():
    os.makedirs(OUT, exist_ok=True)
    ent = lower_entries(1)

    def w(name, text):
        with open(os.path.join(OUT, name), "w") as f:
            f.write(text)

    body = "".join("%d %d %r\n" % (i + 1, j + 1, v) for (i, j), v in ent)
    # 1. plain symmetric, lower triangle, sorted
    w("sym_lower.mtx", B + "MatrixMarket matrix coordinate real symmetric\n"
      "%d %d %d\n" % (N, N, len(ent)) + body)
    # 2. same entries given as UPPER triangle, shuffled, with comments, extra
    #    blanks and leading/trailing tabs
    rng = np.random.default_rng(2)
    order = rng.permutation(len(ent))
    lines = []
    for k in order:
        (i, j), v = ent[k]
        lines.append("\t %d   %d  %r \t\n" % (j + 1, i + 1, v))
    w("sym_upper_shuffled.mtx",
      B + "MatrixMarket matrix coordinate real symmetric\n"
      "% a comment\n%another\n" + ("%d %d %d\n" % (N, N, len(ent)))
      + "".join(lines))
    # 3. general (not symmetric): asking for SSS silently falls back to CSR
    gen = [((i, j), v) for (i, j), v in ent]
    gen += [((j, i), v * 0.5) for (i, j), v in ent if i != j and (i + j) % 3]
    w("general.mtx", B + "MatrixMarket matrix coordinate real general\n"
      "%d %d %d\n" % (N, N, len(gen))
      + "".join("%d %d %r\n" % (i + 1, j + 1, v) for (i, j), v in gen))
    # 4. pattern-style lines (two tokens): every value becomes 0.42
    w("two_tokens.mtx", B + "MatrixMarket matrix coordinate pattern symmetric\n"
      "%d %d %d\n" % (N, N, len(ent))
      + "".join("%d %d\n" % (i + 1, j + 1) for (i, j), _ in ent))
    # 5. zero-based indices announced by the extra header token
    w("base0.mtx", B + "MatrixMarket matrix coordinate real symmetric base-0\n"
      "%d %d %d\n" % (N, N, len(ent))
      + "".join("%d %d %r\n" % (i, j, v) for (i, j), v in ent))
    # 6. no banner: "regular file" mode, first line is the size line -> general
    w("no_banner.mtx", "%d %d %d\n" % (N, N, len(gen))
      + "".join("%d %d %r\n" % (i + 1, j + 1, v) for (i, j), v in gen))
    # 7. duplicates are kept, integer field, exponent notation, extra columns
    dup = list(ent) + [ent[N + 3], ent[N + 9]]
    w("duplicates.mtx", B + "MatrixMarket matrix coordinate integer symmetric\n"
      "%d %d %d\n" % (N, N, len(dup))
      + "".join("%d %d %.3e ignored\n" % (i + 1, j + 1, v)
                for (i, j), v in dup))
    # error fixtures (no goldens; the loader must print + exit(1))
    w("err_no_trailing_newline.mtx",
      B + "MatrixMarket matrix coordinate real symmetric\n"
      "%d %d %d\n" % (N, N, len(ent)) + body.rstrip("\n"))
    w("err_array_format.mtx", B + "MatrixMarket matrix array real general\n2 2\n")
    w("err_bad_banner.mtx", B + "NotMatrixMarket matrix coordinate real general\n")
    w("err_skew.mtx", B + "MatrixMarket matrix coordinate real skew-symmetric\n")
    w("err_short_header.mtx", B + "MatrixMarket matrix coordinate real\n")


if __name__ == "__main__":
    main()
