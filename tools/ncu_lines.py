"""Per-source-line cost of one profiled kernel: joins the SASS page of an .ncu-rep with nvdisasm's line info.
usage: python tools/ncu_lines.py <rep> <mangled-kernel-substring> [top_n]
Prints, per source line: warp instructions executed, share, average active threads, stall samples."""
import csv, io, os, re, subprocess, sys, tempfile, glob, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 45
# every object of the in-tree build holds one cubin; disassemble them all and keep the one with the kernel
dis = []
for obj in sorted(glob.glob(os.environ.get("NCU_OBJ_GLOB") or os.path.join(ROOT, "lumillyrender_b200", "build", "*.o"))):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", obj], cwd=tmp, capture_output=True)
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
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        line_of[int(m.group(1), 16)] = (cur, m.group(2).strip())
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
allrows = list(csv.reader(io.StringIO(raw)))
# one section per profiled launch: a "Kernel Name" row, a header row, then one row per SASS instruction
starts = [i for i, r in enumerate(allrows) if r and r[0] == "Kernel Name"] + [len(allrows)]
want = os.environ.get("KERNEL_MATCH", "")
rows = None
for a, b in zip(starts[:-1], starts[1:]):
    if want in allrows[a][1]:
        rows = allrows[a:b]
        break
print("kernel:", rows[0][1][:100])
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
base = int(rows[2][0], 16)
agg = collections.defaultdict(lambda: [0, 0, 0, 0])
tot = [0, 0, 0]
for r in rows[2:]:
    off = int(r[0], 16) - base
    src = line_of.get(off, (None, ""))[0]
    wi, ti, sm = int(r[ix["Instructions Executed"]]), int(r[ix["Thread Instructions Executed"]]), int(r[ix["# Samples"]])
    a = agg[src]
    a[0] += wi; a[1] += ti; a[2] += sm; a[3] += 1
    tot[0] += wi; tot[1] += ti; tot[2] += sm
print("total warp instr %.3e  thread instr %.3e  avg threads %.2f  samples %d" % (tot[0], tot[1], tot[1] / max(tot[0], 1), tot[2]))
print("%-28s %12s %7s %7s %8s %7s" % ("line", "warp_instr", "share", "thr", "samples", "smp%"))
for src, a in sorted(agg.items(), key=lambda kv: -kv[1][2])[:top]:
    print("%-28s %12.3e %6.2f%% %7.2f %8d %6.2f%%" % ("%s:%s" % src if src else "?", a[0], 100.0 * a[0] / tot[0], a[1] / max(a[0], 1), a[2], 100.0 * a[2] / max(tot[2], 1)))
