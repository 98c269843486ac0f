"""Generates tests/golden/ref_sw_pairs.json from the COMPILED REFERENCE (oracle/_ref/libref_pairhmm.so,
ref_sw_align == hc::IntelSWAligner::align, smithwaterman/intel_smithwaterman.hpp:29-44): seeded
(reference window, haplotype) pairs with the offset and CIGAR the reference returns.  Run here, where
/root/reference exists; the fixture travels."""
import ctypes as C, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from _sw_cases import sw_cases, PARAMS

lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_pairhmm.so"))
lib.ref_sw_align.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_int]
out = {"source": "hc::IntelSWAligner::align compiled from /root/reference (oracle/ref_harness.cpp: ref_sw_align)",
       "seed": 2024, "n": 80, "params": PARAMS, "results": []}
for k, (ref, alt) in enumerate(sw_cases(out["seed"], out["n"])):
    buf = C.create_string_buffer(16384)
    off = lib.ref_sw_align(ref, len(ref), alt, len(alt), *PARAMS[k % len(PARAMS)], buf, 16384)
    out["results"].append([off, buf.value.decode()])
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "ref_sw_pairs.json"), "w"))
print("wrote", len(out["results"]), "pairs")
