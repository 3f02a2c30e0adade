"""Fills the results table of BASELINE.md §4 from bench.py JSON lines: the N = 1 line (headline + `configs`) and, if given, the
N = 2 / 4 / 8 lines of the same workload.  usage: python tools/fill_baseline_table.py n1.json [n2.json n4.json n8.json]"""
import json, os, re, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load(p):
    with open(p) as f:
        for ln in f:
            ln = ln.strip()
            if ln.startswith("{"):
                return json.loads(ln)
    raise SystemExit("no JSON line in " + p)


n1 = load(sys.argv[1])
multi = {l["n_gpus"]: l for l in (load(p) for p in sys.argv[2:])}
rows = []
parity = {
    "1 primitive": "≥ 99.99 % of 4.19 M px / 0 (bit-equal) / ≥ 99 % / see log (full size)",
    "2 new-cbox": "≥ 99.99 % / 0 (bit-equal) / ≥ 99 % / see log (full size)",
    "3 brdf": "≥ 99.99 % of 518 k px / 0 (bit-equal) / ≥ 99 % / see log (full size)",
    "4 welcome-2018 (144k)": "100 % of 3.28 M px / 0 (bit-equal) / 99.75 % / 43.2 (full size)",
    "4 welcome-2018 (1M)": "100 % of 3.28 M px / 0 (bit-equal) / replay exact vs tie rule / —",
}
for c in n1.get("configs", []):
    cpu = c["cpu"]
    rows.append("| %s, %dx%d, %d spp | %.2f · %.1f (%d) | **%.0f · %.0f** | — | — | — | — | %.3f (tree only %.3f) | %.3f | %s |" % (
        c["config"], c["resolution"][0], c["resolution"][1], c["spp"], cpu["msamples_per_s"], cpu["mrays_per_s"], cpu["cores"],
        c["msamples_per_s"], c["mrays_per_s"], c["l2_frac"], c["l2_frac_tree_only"], c["hbm_frac"], parity.get(c["config"], "see tests")))
cb = n1.get("cpu_baseline", {})
r, rh = n1["roofline"], n1.get("roofline_hbm", {})


def cell(n):
    l = multi.get(n)
    return "%.0f · %.0f (e2e %.0f)" % (l["value"], l["mrays_per_s"], l["e2e"]["value"]) if l else "—"


eff = "—"
if 8 in multi:
    b = multi.get(1, n1)                                  # the N = 1 run of the same box, if given
    eff = "%.1f %% (e2e %.1f %%)%s" % (100.0 * multi[8]["value"] / (8 * b["value"]), 100.0 * multi[8]["e2e"]["value"] / (8 * b["e2e"]["value"]),
                                       " vs %.0f on the same box" % b["value"] if 1 in multi else "")
rows.append("| 5 sample.toml, %dx%d, %d spp total (strong scaling) | %.2f · %.1f (%d) | **%.0f · %.0f** (e2e %.0f) | %s | %s | %s | %s | %.3f (tree only %.3f) | %.3f | 100 %% of 2.63 M px / 0 (bit-equal) / 99.95 %% / 1.14 (full size) |" % (
    n1["config"]["resolution"][0], n1["config"]["resolution"][1], n1["config"]["spp_total"], cb.get("value", float("nan")), cb.get("mrays_per_s", float("nan")),
    cb.get("cores", 0), n1["value"], n1["mrays_per_s"], n1["e2e"]["value"], cell(2), cell(4), cell(8), eff, r["frac"], r["frac_tree_only"], rh.get("frac", float("nan"))))
table = ("| config | CPU restatement Msamples/s · Mrays/s (cores) | 1×B200 Msamples/s · Mrays/s | 2× | 4× | 8× | scaling eff. 1→8 | L2 frac | HBM frac | "
         "parity (hit-index %, max rel Δt, 3σ pass %, relMSE) |\n|---|---|---|---|---|---|---|---|---|---|\n" + "\n".join(rows))
path = os.path.join(ROOT, "BASELINE.md")
with open(path) as f:
    s = f.read()
head = s[:s.index("Results table")]
note = ("Results table (filled from `bench.py`'s own JSON lines by `tools/fill_baseline_table.py`; one B200 box, clocks %s MHz, no throttle; the CPU column is the\n"
        "C++ restatement of the reference algorithm on the box's host cores, timed in the same run on a bounded pixel subset; L2 / HBM frac = algorithmic bytes per ray ×\n"
        "rays/s against the measured L2 read peak (%.1f TB/s, in-run microbenchmark) / `MEASURED_PEAKS.json` HBM; parity columns from `tests/test_gpu_fullsize.py` and\n"
        "`tests/test_gpu_parity.py`, logs under `profiles/`):\n\n" % (n1.get("clocks", {}).get("sm_mhz"), (r.get("peak") or 0) / 1e3))
with open(path, "w") as f:
    f.write(head + note + table + "\n")
print(table)
