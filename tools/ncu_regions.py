"""Per-function cost of one profiled kernel: joins the SASS page of an .ncu-rep with nvdisasm's line info of the in-tree
object that holds the kernel and groups the instructions by the device function (device_path.cuh) or 25-line block
(persistent.cuh) they come from.  Checks that both listings are the same code before trusting the join.
usage: python tools/ncu_regions.py <rep> <object.o> <mangled-kernel-substring>"""
import collections, csv, glob, io, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, obj, kern = sys.argv[1], sys.argv[2], sys.argv[3]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
dis = []
for cub in glob.glob(tmp + "/*.cubin"):
    out = subprocess.run(["nvdisasm", "--print-line-info", cub], capture_output=True, text=True).stdout
    if kern in out:
        dis = out.splitlines()
line_of, text_of, cur, inside = {}, {}, None, False
for l in dis:
    if l.startswith(".text."):
        inside = kern in l
        continue
    if not inside:
        continue
    m = re.search(r'//## File "(.*?)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if l.startswith("$") and l.endswith(":"):
        cur = ("call:" + l[:-1].split("$")[-1][-40:], 0)
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        line_of[int(m.group(1), 16)] = cur
        text_of[int(m.group(1), 16)] = m.group(2).split()[0]
src = open(os.path.join(ROOT, "lumillyrender_b200", "csrc", "device_path.cuh")).read().splitlines()
funcs = []
for i, l in enumerate(src, 1):
    m = re.match(r"(?:static __device__ __noinline__|LR_DEV|LR_COLD|LR_GGX)\s+[\w:<>]+\s+(\w+)\s*\(", l)
    if m:
        funcs.append((i, m.group(1)))


def fn(ln):
    name = "?"
    for i, n in funcs:
        if i <= ln:
            name = n
        else:
            break
    return name


raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
base = int(rows[2][0], 16)
mism = sum(1 for r in rows[2:] if text_of.get(int(r[0], 16) - base, "?") != r[1].split()[0].rstrip(";"))
print("instructions %d, opcode mismatches vs the object: %d" % (len(rows) - 2, mism))
agg = collections.defaultdict(lambda: [0, 0, 0])
tot = [0, 0, 0]
for r in rows[2:]:
    s_ = line_of.get(int(r[0], 16) - base)
    if s_ is None:
        key = "?"
    elif s_[0] == "device_path.cuh":
        key = "dp:" + fn(s_[1])
    elif s_[0] == "persistent.cuh":
        key = "pk:%d" % (s_[1] // 25 * 25)
    elif s_[0] == "path_vertex.inc":
        key = "pv:%d" % (s_[1] // 25 * 25)
    elif s_[0] in ("pool.cuh", "cpool.cuh"):
        g = int(os.environ.get("REGION_LINES", "10"))
        key = "%s:%d" % (s_[0], s_[1] // g * g)
    else:
        key = s_[0]
    wi, ti, sm = int(r[ix["Instructions Executed"]]), int(r[ix["Thread Instructions Executed"]]), int(r[ix["# Samples"]])
    a = agg[key]
    a[0] += wi; a[1] += ti; a[2] += sm
    tot[0] += wi; tot[1] += ti; tot[2] += sm
print("total warp instr %.3e  thread instr %.3e  avg threads %.2f" % (tot[0], tot[1], tot[1] / tot[0]))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(os.environ.get("REGION_TOP", "32"))]:
    print("%-46s warp-instr %6.2f%%  thr %5.1f  samples %5.2f%%" % (k, 100 * a[0] / tot[0], a[1] / max(a[0], 1), 100 * a[2] / tot[2]))
