"""Turns the artefacts of scripts/gpu_evidence.sh (gpurun_out/ev_<tag>_*) into the committed summaries under profiles/:
bench lines, the launch list, details pages of the ncu captures, DRAM traffic ratios (profiles/traffic.json, read by
bench.py), the per-page-size DRAM counters, the SASS opcode histogram of the shipped library, the sanitizer logs.
usage: collect_profiles.py <tag>          (run in the development container; needs ncu, cuobjdump)"""
import collections
import csv
import io
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
SRC = os.path.join(ROOT, "gpurun_out", f"ev_{tag}_")
DST = os.path.join(ROOT, "profiles")
os.makedirs(DST, exist_ok=True)


def have(name):
    return os.path.exists(SRC + name)


def copy(name, out=None):
    if have(name):
        shutil.copy(SRC + name, os.path.join(DST, out or f"{tag}_{name}"))


def last_json_line(path):
    for line in reversed(open(path).read().strip().split("\n")):
        if line.startswith("{"):
            return json.loads(line)
    return None


# ---- bench lines (one JSON object per file, pretty enough to diff)
for name in ("bench_n1.json", "bench_reference_arm.json", "bench_random.json", "bench_texture.json", "bench_n2.json", "bench_n4.json", "bench_n8.json"):   # (n2 / n8: separate multi-GPU calls)
    if have(name):
        obj = last_json_line(SRC + name)
        if obj:
            json.dump(obj, open(os.path.join(DST, f"{tag}_{name}"), "w"), indent=1)
for name in ("pytest.log", "pytest_multi.log", "config0.log", "kinds.log", "launches.csv", "page_size_sweep.json", "sweep_dram.csv",
             "sanitizer_memcheck.log", "sanitizer_synccheck.log", "sanitizer_racecheck_lockstep.log", "sanitizer_racecheck_production.log"):
    copy(name)


# ---- ncu captures: details page + DRAM traffic per launch
def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def metric_rows(rep):
    rows = ncu_csv(rep, "raw")
    if len(rows) < 3:
        return []
    hdr = rows[0]
    return [dict(zip(hdr, r)) for r in rows[2:]]


traffic_path = os.path.join(DST, "traffic.json")
traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
traffic["_comment"] = ("dram__bytes_read.sum + dram__bytes_write.sum of the decode kernel(s) from `ncu --set full --clock-control none` captures "
                       "(scripts/gpu_evidence.sh; details pages: profiles/<tag>_prof_*_details.csv), as a ratio to the algorithmic bytes of the "
                       "captured launch (compressed bytes read + decompressed bytes written); bench.py multiplies the ratio by the algorithmic "
                       "bytes of its own launch.")
# the 4 KiB-page capture: details page + instructions / stall samples per code region (scripts/ncu_regions.py;
# needs the library the capture was taken with: build/variants/libbgx_<tag>final.so, else the in-tree build)
rep4k = SRC + "page4k.ncu-rep"
if os.path.exists(rep4k):
    det = subprocess.run(["ncu", "-i", rep4k, "--page", "details", "--csv"], capture_output=True, text=True).stdout
    open(os.path.join(DST, f"{tag}_prof_page4k_details.csv"), "w").write(det)
    lib = os.path.join(ROOT, "build", "variants", f"libbgx_{tag}final.so")
    if not os.path.exists(lib):
        lib = os.path.join(ROOT, "brotli_g_sdk_b200", "libbrotlig_b200.so")
    reg = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_regions.py"), rep4k, lib, "196608"], capture_output=True, text=True).stdout
    open(os.path.join(DST, f"{tag}_prof_page4k_regions.txt"), "w").write(
        "# 4 KiB single-page streams, 1 GiB mixed: warp instructions per COMPRESSED page (196 608 of the 262 144 pages) and stall samples per code region\n" + reg)
ALGO = {}   # algorithmic bytes of the captured launches, from the logs of gpu_prof_one.py
for kind, log in (("mixed", "ncu_mixed.log"), ("random", "ncu_raw.log"), ("texture", "ncu_texture.log")):
    rep = SRC + {"mixed": "mixed", "random": "raw", "texture": "texture"}[kind] + ".ncu-rep"
    if not os.path.exists(rep):
        continue
    det = subprocess.run(["ncu", "-i", rep, "--page", "details", "--csv"], capture_output=True, text=True).stdout
    open(os.path.join(DST, f"{tag}_prof_{'raw' if kind == 'random' else kind}_details.csv"), "w").write(det)
    rows = metric_rows(rep)
    rd = sum(float(r.get("dram__bytes_read.sum", 0) or 0) for r in rows)
    wr = sum(float(r.get("dram__bytes_write.sum", 0) or 0) for r in rows)
    unit_r = next((r for r in ncu_csv(rep, "raw")[1:2]), None)
    # units row: bytes may be reported in Mbyte / Gbyte
    hdr = ncu_csv(rep, "raw")[0]
    units = dict(zip(hdr, unit_r)) if unit_r else {}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    rd *= scale.get(units.get("dram__bytes_read.sum", "byte"), 1.0)
    wr *= scale.get(units.get("dram__bytes_write.sum", "byte"), 1.0)
    algo = None
    if have(log):
        m = re.search(r"algorithmic_bytes (\d+)", open(SRC + log).read())
        if m:
            algo = float(m.group(1))
    if algo is None:   # captures made before gpu_prof_one.py printed it: re-encode the same payload here
        sys.path.insert(0, ROOT)
        import brotli_g_sdk_b200 as b
        from brotli_g_sdk_b200 import datagen
        if kind == "texture":
            from brotli_g_sdk_b200.encoder import DataconditionParams
            d = datagen.bc_texture(1024, 1024, 3, seed=21)
            st = b.Encode(d, dcParams=DataconditionParams(precondition=True, swizzle=True, delta_encode=True, format=3, width_blocks=1024, height_blocks=1024))
            algo = float((len(st) + len(d)) * 64)
        else:
            d = (datagen.mixed if kind == "mixed" else datagen.random_bytes)(64 << 20, seed=21)
            algo = float((len(b.Encode(d)) + len(d)) * 16)
    entry = {"measured_bytes": rd + wr, "dram_read_bytes": rd, "dram_write_bytes": wr, "capture": f"ev_{tag}_{kind} ({len(rows)} kernel launch(es) of one decode)"}
    if algo:
        entry["algorithmic_bytes"] = algo
        entry["ratio_to_algorithmic"] = (rd + wr) / algo
    traffic[kind] = entry
json.dump(traffic, open(traffic_path, "w"), indent=1)

# ---- SASS opcode histogram of the shipped library
lib = os.path.join(ROOT, "brotli_g_sdk_b200", "libbrotlig_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
hist, fn = {}, None
for line in sass.split("\n"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = "bgx_decode_pages_kernel" if "decode_pages" in m.group(1) else "bgx_decondition_kernel" if "decondition" in m.group(1) else m.group(1)
        hist[fn] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and fn:
        hist[fn][m.group(1).split(".")[0]] += 1
        if m.group(1).startswith(("UBLKCP", "LDGSTS", "SYNCS", "REDUX", "UTMA")):
            hist[fn]["  " + m.group(1)] += 1
with open(os.path.join(DST, f"{tag}_sass_histogram.txt"), "w") as f:
    f.write(f"SASS opcode histogram of brotli_g_sdk_b200/libbrotlig_b200.so (cuobjdump -sass; sm_100a). Indented rows: full mnemonics of the\n"
            f"sm_100a artefacts -- UBLKCP = cp.async.bulk (TMA bulk copy), LDGSTS = cp.async, SYNCS.* = mbarrier, REDUX = warp reduce.\n")
    for fn, c in hist.items():
        f.write(f"\n== {fn}: {sum(v for k, v in c.items() if not k.startswith('  '))} instructions\n")
        for k, v in sorted(c.items(), key=lambda kv: (kv[0].startswith("  "), -kv[1])):
            f.write(f"{v:6d}  {k}\n")
print("profiles/ updated:", sorted(x for x in os.listdir(DST) if x.startswith(tag) or x == "traffic.json"))
