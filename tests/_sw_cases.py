"""Seeded (reference window, haplotype) pairs for the Smith-Waterman tests: haplotypes derived from the
window by substitutions, insertions and deletions (what the assembler hands the aligner), sub-windows,
unrelated sequences, tiny and maximum-length cases."""
import numpy as np

PARAMS = [[200, -150, -260, -11], [3, -1, -4, -3], [25, -50, -110, -6], [10, -15, -30, -5]]   # intel_smithwaterman.hpp:21-25


def sw_cases(seed, n, max_len=600):
    rng = np.random.default_rng(seed)
    alpha = np.frombuffer(b"ACGT", np.uint8)
    cases = []
    for it in range(n):
        nref = int(rng.integers(1, max_len + 1))
        if it % 17 == 3:
            nref = 1023 if it % 2 else 1                       # the longest the reference can take, and a single base
        ref = alpha[rng.integers(0, 4, nref)]
        mode = int(rng.integers(0, 4))
        if mode == 0:
            alt = alpha[rng.integers(0, 4, int(rng.integers(1, max_len + 1)))]
        else:
            a = int(rng.integers(0, nref)); b = int(rng.integers(a, nref)) + 1
            alt = list(ref[a:b]) if mode > 1 else list(ref)
            for _ in range(int(rng.integers(0, 6))):
                if not alt:
                    break
                k = int(rng.integers(0, len(alt))); t = int(rng.integers(0, 3))
                if t == 0:
                    alt[k] = int(alpha[rng.integers(0, 4)])
                elif t == 1:
                    del alt[k:k + int(rng.integers(1, 12))]
                else:
                    alt[k:k] = [int(x) for x in alpha[rng.integers(0, 4, int(rng.integers(1, 12)))]]
            alt = np.array((alt or [65])[:1023], np.uint8)
        cases.append((ref.tobytes(), alt.tobytes()))
    return cases
