"""Executed instructions and stall samples of the page kernel per code REGION (the function whose body a SASS row's
line belongs to; rows of small inlined helpers inherit the region of the last such row), from an `ncu --set full
--import-source on` capture (dev aid). usage: ncu_regions.py <capture.ncu-rep> <lib.so> <units in the launch>"""
import collections, csv, io, os, re, subprocess, sys, tempfile
rep, lib, units = sys.argv[1], sys.argv[2], float(sys.argv[3])
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = open(os.path.join(root, "brotli_g_sdk_b200", "csrc", "page_decode.cuh")).read().split("\n")
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
sass = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, check=True, stdout=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(d) if f.endswith(".cubin") and "api" not in f][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cubin)], capture_output=True, text=True).stdout
funcs = []
for i, l in enumerate(src):
    m = re.match(r"(BGX_DEV|BGX_COLD|BGX_DEV_NOINLINE|BGX_HD)\s+.*?(\w+)\(", l)
    if m and not l.strip().endswith(";"):
        funcs.append((i + 1, m.group(2)))
def fn_of(n):
    name = None
    for a, nm in funcs:
        if a <= n: name = nm
        else: break
    return name
BIG = {"load_table", "load_tables", "build_table", "producer_warp", "consumer_warp", "slow_round", "decode_page_cta", "copy_page_cta_bulk",
       "copy_page_cta_ldst", "copy_page_warp", "delta_decode_warp", "cold_flush_bytes", "cold_read_fields", "decode_literals", "copy_page_cta"}
regions, infn, region = [], False, "kernel"
for l in dis.split("\n"):
    if l.startswith("//---") and ".text." in l:
        infn = "decode_pages" in l; region = "kernel"; continue
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        f, n = os.path.basename(m.group(1)), int(m.group(2))
        if f == "page_decode.cuh":
            fn = fn_of(n)
            if fn in BIG: region = fn
        elif f == "bgx_cuda.cu": region = "kernel"
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,6}\*/", l): regions.append(region)
if len(regions) != len(sass): print(f"warning: {len(sass)} rows vs {len(regions)} instructions", file=sys.stderr)
ie, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
names = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
inst, samp, st, static = collections.Counter(), collections.Counter(), {}, collections.Counter()
for reg, r in zip(regions, sass):
    inst[reg] += int(r[ie] or 0); samp[reg] += int(r[isamp] or 0); static[reg] += 1
    d = st.setdefault(reg, collections.Counter())
    for n in names: d[n] += int(r[hdr.index(n)] or 0)
ts = sum(samp.values())
for reg, v in inst.most_common():
    d = st[reg]; tot = max(sum(d.values()), 1)
    print(f"{reg:20s} {v / units:9.1f} inst/unit  {100.0 * samp[reg] / ts:5.1f}% samples  static {static[reg]:5d}  " +
          " ".join(f"{k[6:]}={100.0 * x / tot:.0f}" for k, x in d.most_common(5)))
