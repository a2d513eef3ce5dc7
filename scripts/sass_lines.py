"""Static SASS instruction count per source line of the page kernel (dev aid; needs nvdisasm and a -lineinfo build).
usage: sass_lines.py <lib.so> [first_line last_line]   -- prints `line  count  source` for page_decode.cuh"""
import os, re, subprocess, sys, tempfile, collections
lib = sys.argv[1]
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3]) if len(sys.argv) > 3 else 10**9
src = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "brotli_g_sdk_b200", "csrc", "page_decode.cuh")
lines = open(src).read().split("\n")
with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, check=True, stdout=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(d) if f.endswith(".cubin") and "api" not in f][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cubin)], capture_output=True, text=True).stdout
cnt = collections.Counter()
cur = None
infn = False
for l in dis.split("\n"):
    if l.startswith("//---") and ".text." in l:
        infn = "decode_pages" in l
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,6}\*/", l) and cur:
        cnt[cur] += 1
tot = 0
for (f, n), c in sorted(cnt.items()):
    if f == "page_decode.cuh" and lo <= n <= hi:
        tot += c
        print(f"{n:5d} {c:4d}  {lines[n-1].strip()[:110]}")
print("total in range:", tot, " kernel total:", sum(cnt.values()))
