"""Per-kernel SASS census of libmixvae_b200.so: which kernels carry tcgen05 (UTC*MMA), TMA (UTMALDG), tensor-memory
loads/stores (LDTM/STTM), legacy warp MMAs (HMMA) and how many instructions.  usage: sass_census.py [lib.so] > csv"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "distributed-vae_b200", "mmidas_b200", "libmixvae_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur, rows = None, collections.OrderedDict()
pats = {"UTCMMA": r"\bUTC[A-Z]*MMA", "UTMALDG": r"\bUTMALDG", "LDTM": r"\bLDTM", "STTM": r"\bSTTM", "HMMA": r"\bHMMA", "SYNCS": r"\bSYNCS",
        "LDGSTS": r"\bLDGSTS", "DADD/DFMA/DMUL": r"\bD(ADD|FMA|MUL)\b", "STL/LDL": r"\b(STL|LDL)",
        "PDL (ACQBULK/PREEXIT)": r"\b(ACQBULK|PREEXIT)"}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        rows[cur] = collections.Counter()
        continue
    if cur and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
        rows[cur]["instructions"] += 1
        for k, p in pats.items():
            if re.search(p, line):
                rows[cur][k] += 1
def demangle(n):
    r = subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
    r = r.replace("(anonymous namespace)::", "")
    r = re.sub(r"\(.*", "", r).replace("void ", "").replace("mvae::", "")
    return r
print("# SASS census of libmixvae_b200.so (cuobjdump -sass, sm_100a): instruction counts per kernel")
print("kernel,instructions," + ",".join(pats))
for n, c in rows.items():
    print('"%s",%d,%s' % (demangle(n), c["instructions"], ",".join(str(c[k]) for k in pats)))
