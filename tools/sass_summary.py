"""profiles/sass_summary.md: per-kernel counts of the SASS mnemonics that prove the Blackwell paths (tcgen05 MMA, tensor-memory
loads/stores, TMA) in the in-tree libwca_b200.so.  Runs in the build container (cuobjdump, no GPU needed).
usage: python tools/sass_summary.py [out.md]"""
import collections, os, re, subprocess, sys

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(root, "whisper_char_alignment_b200", "libwca_b200.so")
out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(root, "profiles", "sass_summary.md")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
WATCH = ["UTCHMMA", "UTCMMA", "LDTM", "STTM", "UTMALDG", "UTMAPF", "UTCBAR", "SYNCS", "LDGSTS", "MUFU.EX2", "SHFL", "REDUX", "BAR.SYNC"]
kernels, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kernels[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        kernels[cur]["_total"] += 1
        for w in WATCH:
            if op.startswith(w):
                kernels[cur][w] += 1
demangled = {}
try:
    names = subprocess.run(["cu++filt"] + list(kernels), capture_output=True, text=True, check=True).stdout.splitlines()
    demangled = dict(zip(kernels, names))
except Exception:
    pass
cols = [w for w in WATCH if any(k[w] for k in kernels.values())]
with open(out_path, "w") as f:
    f.write("# SASS summary of `libwca_b200.so` (sm_100a)\n\n")
    f.write("`python tools/sass_summary.py` (cuobjdump -sass of the in-tree library, in the build container).  `UTCHMMA` = tcgen05.mma "
            "(kind::tf32), `LDTM` / `STTM` = tcgen05.ld / tcgen05.st (tensor memory), `UTMALDG` = TMA tensor-map load, `UTMAPF` = TMA L2 "
            "prefetch, `UTCBAR` = tcgen05.commit, `SYNCS` = mbarrier operations, `LDGSTS` = cp.async.\n\n")
    f.write("| kernel | instructions | " + " | ".join(cols) + " |\n|---|---|" + "---|" * len(cols) + "\n")
    total = collections.Counter()
    for k, c in kernels.items():
        name = demangled.get(k, k)
        name = re.sub(r"\((?:[^()]|\([^()]*\))*\)\s*$", "", name).replace("void ", "").replace("(int)", "").replace("(bool)", "")
        f.write(f"| `{name}` | {c['_total']} | " + " | ".join(str(c[w]) if c[w] else "" for w in cols) + " |\n")
        total.update(c)
    f.write(f"| **all kernels** | {total['_total']} | " + " | ".join(str(total[w]) for w in cols) + " |\n")
print(open(out_path).read())
