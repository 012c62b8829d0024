"""Per-kernel SASS mnemonic counts of lib/librbslam.so (cuobjdump -sass, sm_100a) -> profiles/sass_summary_r2.txt"""
import subprocess, re, collections, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "rao-blackwellized-slam-smoothing_b200", "lib", "librbslam.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cols = ["UBLKCP", "SYNCS", "DMMA", "LDGSTS", "DFMA", "LDS", "STS", "LDG", "STG", "LDL", "STL", "BAR.SYNC"]
kern, cnt, order = None, collections.defaultdict(collections.Counter), []
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = m.group(1); order.append(kern); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and kern:
        op = m.group(1)
        cnt[kern]["instrs"] += 1
        for c in cols:
            if op == c or op.startswith(c + "."):
                cnt[kern][c] += 1
names = subprocess.run(["c++filt"] + order, capture_output=True, text=True).stdout.splitlines()
def short(n):
    n = re.sub(r"\(.*\)$", "", n)            # drop the parameter list
    n = re.sub(r"^void ", "", n).replace("rb::", "").replace("(anonymous namespace)::", "")
    return re.sub(r"\((int|bool)\)", "", n)
out = ["# SASS mnemonic counts per kernel of lib/librbslam.so (cuobjdump -sass, sm_100a), round 2 final build (tools/sass_summary.py)",
       "# UBLKCP = cp.async.bulk (TMA, non-tensor form), SYNCS = mbarrier ops, DMMA = fp64 tensor-core MMA (mma.sync.m8n8k4.f64),",
       "# LDGSTS = cp.async (Ampere-style async copy), LDL/STL = local memory (spills / local arrays).  No UTMALDG (tensor-map TMA:",
       "# the 1-D bulk form is the right one for contiguous slab streams) and no UTC*MMA (tcgen05 has no f64 kind).",
       "\t".join(["kernel", "instrs"] + cols)]
for k, n in sorted(zip(order, names), key=lambda kn: short(kn[1])):
    out.append("\t".join([short(n), str(cnt[k]["instrs"])] + [str(cnt[k][c]) for c in cols]))
open(os.path.join(ROOT, "profiles", "sass_summary_r2.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(l for l in out if any(w in l for w in ("k_stream_fam_pt", "k_chol_inv", "k_dgemm", "kernel\t", "k_stream_fam<", "k_peer"))))
