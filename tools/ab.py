"""A/B timing of library variants inside ONE gpurun call (different calls land on different boxes / clocks):
  here:    python tools/ab.py build name1:DEF1,DEF2 name2: ...        (builds lumillyrender_b200/variants/lib_<name>.so)
  on GPU:  python tools/ab.py run name1 name2 ... [--rounds 2]        (interleaved rounds of tools/sweep.py per variant)"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if sys.argv[1] == "build":
    from lumillyrender_b200 import build
    for spec in sys.argv[2:]:
        name, _, defs = spec.partition(":")
        print(build.build_variant(name, [d for d in defs.split(",") if d]))
else:
    args = sys.argv[2:]
    rounds = 2
    if "--rounds" in args:
        i = args.index("--rounds")
        rounds = int(args[i + 1])
        del args[i:i + 2]
    configs = None                       # --configs CFG ...: sweep.py configurations (LR_* knobs) run for every variant
    if "--configs" in args:
        i = args.index("--configs")
        configs = args[i + 1:]
        del args[i:]
    names = args
    for r in range(rounds):
        for n in names:
            env = dict(os.environ, LUMILLY_LIB=os.path.join(ROOT, "lumillyrender_b200", "variants", "lib_%s.so" % n))
            cfgs = ["%s,%s" % (n, c) for c in configs] if configs else [n]
            out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sweep.py")] + cfgs, env=env, capture_output=True, text=True)
            print(out.stdout.strip() or out.stderr[-400:], flush=True)
