"""profiles/sass_r02.txt: per-kernel opcode histogram of the built library (cuobjdump -sass): the instructions that prove the
FP64 tensor path (DMMA), TMA (UTMALDG), cp.async (LDGSTS), mbarriers (SYNCS), cluster / PDL control (ACQBULK ...)."""
import collections, os, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "abstractbayesopt.jl_b200", "libabo_cuda.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEYS = ["DMMA", "UTMALDG", "LDGSTS", "SYNCS", "DFMA", "DADD", "DMUL", "MUFU", "LDS", "STS", "LDG", "STG", "SHFL", "BAR", "ATOM", "RED", "STL", "LDL"]
kern, hist, arch = None, collections.OrderedDict(), None
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
        kern = re.sub(r"\(.*", "", kern)
        hist[kern] = collections.Counter()
        continue
    m = re.search(r"arch = (sm_\w+)", ln)
    if m:
        arch = m.group(1)
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", ln)
    if m and kern:
        hist[kern][m.group(1)] += 1
print(f"# cuobjdump -sass opcode histogram per kernel, {os.path.basename(lib)}, cubin arch {arch}; columns: total instructions, then {', '.join(KEYS)}")
for kname, h in hist.items():
    tot = sum(h.values())
    print(f"{kname[:90]:<90} {tot:6d} " + " ".join(f"{k}={h[k]}" for k in KEYS if h[k]))
