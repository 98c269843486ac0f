"""chrM-style end-to-end wall time (BASELINE config 0 / SURVEY 8f-1,2): the reference's whole driver around
its own engine, around hc::B200PairHMM (one synchronous call per window) and the batched driver
(hc::B200RegionBatcher, assembly on T host threads).  Best of N runs each; prints one JSON line."""
import json, os, re, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package
pkg = load_package()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
length = int(sys.argv[2]) if len(sys.argv) > 2 else 16569
d = tempfile.mkdtemp()
prefix = os.path.join(d, "chrm_like")
pkg.synth.chrm_like(prefix, length=length)
exe = lambda name: os.path.join(ROOT, "oracle", "_ref", name)
def run(name, extra=()):
    best = None
    for _ in range(n):
        out = os.path.join(d, name + ".vcf")
        t0 = time.perf_counter()
        r = subprocess.run([exe(name), "-I", prefix + ".sam", "-R", prefix + ".fa", "-O", out, *extra], capture_output=True, text=True)
        wall = time.perf_counter() - t0
        assert r.returncode == 0, r.stderr[-500:]
        m = re.search(r"init_s=([0-9.]+) do_work_s=([0-9.]+)", r.stderr)
        rec = dict(wall_s=round(wall, 3), init_s=float(m.group(1)), do_work_s=float(m.group(2)))
        if best is None or rec["do_work_s"] < best["do_work_s"]: best = rec
    best["vcf"] = open(out).read()
    return best
res = {"contig_bp": length, "runs_each": n, "host_cores": os.cpu_count()}
res["reference_engine"] = run("hc_e2e_ref")
res["b200_per_window"] = run("hc_e2e_b200")
for t in (1, 4, 16):
    res[f"b200_batched_T{t}"] = run("hc_e2e_b200_batched", ("-T", str(t)))
ref_vcf = res["reference_engine"].pop("vcf")
for k, v in res.items():
    if isinstance(v, dict) and "vcf" in v: v["vcf_identical"] = (v.pop("vcf") == ref_vcf)
print(json.dumps(res))
