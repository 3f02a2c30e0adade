"""Where a kernel waits for instructions: per source region (as tools/ncu_regions.py cuts them) the `stall_no_inst` samples of
one profiled launch, next to the warp instructions executed there and the region's static size; plus the hot footprint of the
kernel (static instructions that account for 90 / 99 % of the dynamic ones).
usage: python tools/ncu_ifetch.py <rep> <object.o> <mangled-kernel-substring>"""
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
line_of, cur, inside = {}, None, False
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
src = open(os.path.join(ROOT, "lumillyrender_b200", "csrc", "device_path.cuh")).read().splitlines()
funcs = [(i, m.group(1)) for i, l in enumerate(src, 1)
         for m in [re.match(r"(?:static __device__ __noinline__|LR_DEV|LR_COLD|LR_GGX)\s+[\w:<>]+\s+(\w+)\s*\(", l)] if m]


def fn(ln):
    name = "?"
    for i, n in funcs:
        if i <= ln:
            name = n
    return name


raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
ix = {h: i for i, h in enumerate(rows[1])}
base = int(rows[2][0], 16)
agg = collections.defaultdict(lambda: [0, 0, 0, 0])        # static instructions, warp instructions, no_inst samples, all samples
per = []
for r in rows[2:]:
    s_ = line_of.get(int(r[0], 16) - base)
    if s_ is None:
        key = "?"
    elif s_[0] == "device_path.cuh":
        key = "dp:" + fn(s_[1])
    elif s_[0] in ("path_vertex.inc", "pool.cuh", "persistent.cuh"):
        key = "%s:%d" % (s_[0].split(".")[0], s_[1] // 25 * 25)
    else:
        key = s_[0]
    wi, ni, sm = int(r[ix["Instructions Executed"]]), int(r[ix["stall_no_inst"]] or 0), int(r[ix["# Samples"]] or 0)
    a = agg[key]
    a[0] += 1; a[1] += wi; a[2] += ni; a[3] += sm
    per.append(wi)
tot_w, tot_n, tot_s = sum(per), sum(a[2] for a in agg.values()), sum(a[3] for a in agg.values())
per.sort(reverse=True)
acc, n90, n99 = 0, None, None
for k, w in enumerate(per, 1):
    acc += w
    if n90 is None and acc >= 0.90 * tot_w:
        n90 = k
    if n99 is None and acc >= 0.99 * tot_w:
        n99 = k
print("static instructions %d (%.0f KB), executed at least once %d; 90 %% of the dynamic instructions come from %d of them (%.0f KB), 99 %% from %d (%.0f KB)" % (
    len(per), len(per) * 16 / 1024, sum(1 for w in per if w > 0), n90, n90 * 16 / 1024, n99, n99 * 16 / 1024))
print("no_instruction samples: %d of %d samples (%.1f %%)" % (tot_n, tot_s, 100.0 * tot_n / max(tot_s, 1)))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][2])[:28]:
    print("%-40s static %5d  warp-instr %6.2f%%  no_inst samples %6.2f%% of all no_inst, %5.1f%% of the region's own samples" % (
        k, a[0], 100 * a[1] / tot_w, 100 * a[2] / max(tot_n, 1), 100 * a[2] / max(a[3], 1)))
