"""Dynamic instruction counts / stall samples per source line of the page kernel, from an `ncu --set full
--import-source on` capture (dev aid). The i-th SASS row of the capture is matched with the i-th instruction of
nvdisasm's listing of the same build, whose -lineinfo gives the source line.
usage: ncu_lines.py <capture.ncu-rep> <lib.so> <rounds in the launch> [first_line last_line]"""
import collections, csv, io, os, re, subprocess, sys, tempfile
rep, lib, rounds = sys.argv[1], sys.argv[2], float(sys.argv[3])
lo = int(sys.argv[4]) if len(sys.argv) > 4 else 0
hi = int(sys.argv[5]) if len(sys.argv) > 5 else 10**9
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
ci = {n: hdr.index(n) for n in ("Source", "# Samples", "Instructions Executed", "Thread Instructions Executed")}
sass = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr)]
with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, check=True, stdout=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(d) if f.endswith(".cubin") and "api" not in f][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cubin)], capture_output=True, text=True).stdout
lines_of = []
cur, infn = None, False
for l in dis.split("\n"):
    if l.startswith("//---") and ".text." in l:
        infn = "decode_pages" in l
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
    elif re.match(r"\s+/\*[0-9a-f]{4,6}\*/", l):
        lines_of.append((cur, l.split("*/", 1)[1].strip().rstrip(";").strip()))
if len(lines_of) != len(sass):
    print(f"warning: {len(sass)} rows in the capture vs {len(lines_of)} instructions in the listing", file=sys.stderr)
inst, samp, thr = collections.Counter(), collections.Counter(), collections.Counter()
for (key, _), r in zip(lines_of, sass):
    inst[key] += int(r[ci["Instructions Executed"]] or 0)
    thr[key] += int(r[ci["Thread Instructions Executed"]] or 0)
    samp[key] += int(r[ci["# Samples"]] or 0)
src = {}
for f in ("page_decode.cuh", "bgx_cuda.cu"):
    p = os.path.join(root, "brotli_g_sdk_b200", "csrc", f)
    src[f] = open(p).read().split("\n")
tot_i, tot_s = sum(inst.values()), sum(samp.values())
print(f"kernel: {tot_i} warp instructions ({tot_i / rounds:.0f} per round), {tot_s} samples")
ri = rs = 0
for key in sorted(k for k in inst if k):
    f, n = key
    if f != "page_decode.cuh" or not (lo <= n <= hi) or inst[key] == 0:
        continue
    ri += inst[key]; rs += samp[key]
    print(f"{n:5d} {inst[key] / rounds:7.1f} i/round {100.0 * samp[key] / tot_s:5.1f}% smp  thr/warp {thr[key] / max(inst[key], 1):4.1f}  {src[f][n - 1].strip()[:100]}")
print(f"range: {ri / rounds:.0f} instructions per round, {100.0 * rs / tot_s:.1f}% of the samples")
# ---- sticky attribution to the two role loops: rows of inlined helpers inherit the role of the last row that
#      belongs to a line inside producer_warp / consumer_warp (code is laid out roughly in source order)
def body(name):
    s = src["page_decode.cuh"]
    a = next(i for i, l in enumerate(s) if l.startswith("BGX_DEV void " + name))
    b = next(i for i in range(a + 1, len(s)) if s[i] == "}")
    return a + 1, b + 1
pa, pb = body("producer_warp")
ca, cb = body("consumer_warp")
role, acc, wait = "other", collections.Counter(), collections.Counter()
for (key, text), r in zip(lines_of, sass):
    if key and key[0] == "page_decode.cuh":
        if pa <= key[1] <= pb: role = "producer"
        elif ca <= key[1] <= cb: role = "consumer"
    n = int(r[ci["Instructions Executed"]] or 0)
    if "SYNCS.PHASECHK" in text or (key and key[0] == "page_decode.cuh" and src["page_decode.cuh"][key[1] - 1].strip().startswith(("while (!done)", '"selp.u32'))):
        wait[role] += n
    else:
        acc[role] += n
print("by role (instructions per round):", {k: round(v / rounds, 1) for k, v in acc.items()}, "mbarrier wait loops:", {k: round(v / rounds, 1) for k, v in wait.items()})
if os.environ.get("ROLE"):
    want = os.environ["ROLE"]
    role, per = "other", collections.Counter()
    for (key, text), r in zip(lines_of, sass):
        if key and key[0] == "page_decode.cuh":
            if pa <= key[1] <= pb: role = "producer"
            elif ca <= key[1] <= cb: role = "consumer"
        if role == want and key:
            per[key] += int(r[ci["Instructions Executed"]] or 0)
    print(f"---- {want}: lines with >= 1 instruction per round")
    for key in sorted(per):
        if per[key] / rounds >= 1.0:
            f, n = key
            print(f"{f[:4]}:{n:5d} {per[key] / rounds:7.1f}  {src[f][n - 1].strip()[:120] if f in src else ''}")
if os.environ.get("STALLS"):
    names = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    role, st = "other", {}
    for (key, text), r in zip(lines_of, sass):
        if key and key[0] == "page_decode.cuh":
            if pa <= key[1] <= pb: role = "producer"
            elif ca <= key[1] <= cb: role = "consumer"
        d = st.setdefault(role, collections.Counter())
        for nme in names:
            d[nme] += int(r[hdr.index(nme)] or 0)
    for role, d in st.items():
        tot = sum(d.values())
        print(role, tot, {k[6:]: round(100.0 * v / max(tot, 1), 1) for k, v in d.most_common(9)})

if os.environ.get("REASON"):     # lines ranked by the samples of one stall reason, e.g. REASON=long_sb
    col = next(h for h in hdr if h.startswith("stall_" + os.environ["REASON"]) and "Not Issued" not in h)
    per = collections.Counter()
    for (key, _), r in zip(lines_of, sass):
        per[key] += int(r[hdr.index(col)] or 0)
    tot = max(sum(per.values()), 1)
    print(f"---- {col}: {tot} samples ({100.0 * tot / tot_s:.1f}% of all)")
    for key, v in per.most_common(25):
        if key:
            f, n = key
            print(f"{f[:4]}:{n:5d} {100.0 * v / tot:5.1f}%  {src[f][n - 1].strip()[:110] if f in src else ''}")
