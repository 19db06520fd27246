"""Opcode histogram of the shipped cubins (cuobjdump -sass of libbicgstab_b200.so) -> profiles/sass_summary.txt.
Evidence that the kernels are native sm_100a code using the 1-D TMA bulk-copy / mbarrier path (UBLKCP, SYNCS), fp64
FMAs (DFMA), system-scope stores for the NVLink mailboxes, L1 invalidation after acquire fences (CCTL.IVALL); no
tensor-core opcode is expected (no dense contraction on this path)."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "mpi-bicgstab_b200", "libbicgstab_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
per, cur, arch = collections.defaultdict(collections.Counter), None, set()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
        cur = cur.replace("(anonymous namespace)::", "")
        cur = cur[:cur.rfind("(")] if "(" in cur else cur
        continue
    m = re.search(r"arch = (sm_\w+)", line)
    if m:
        arch.add(m.group(1))
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_]*(?:\.[A-Z0-9_.]+)?)", line)
    if m and cur:
        per[cur][m.group(1)] += 1
interesting = ["UBLKCP", "SYNCS", "DFMA", "DADD", "DMUL", "LDG", "STG", "LDS", "CCTL", "MEMBAR", "FENCE", "ATOM", "RED", "SHFL", "BAR", "HMMA", "UTMALDG", "UTCHMMA", "NANOSLEEP"]
lines = [f"# cuobjdump -sass {os.path.relpath(so, ROOT)}   arch: {sorted(arch)}", f"# kernels: {len(per)}", ""]
tot = collections.Counter()
for k in per:
    tot.update(per[k])
def fam(counter, prefix):
    return sum(v for op, v in counter.items() if op.split(".")[0] == prefix or op.startswith(prefix + "."))
lines.append("## whole library (opcode family: count)")
lines.append("  " + "  ".join(f"{p}:{fam(tot, p)}" for p in interesting))
lines.append("")
lines.append("## variants worth seeing verbatim")
for op, v in sorted(tot.items()):
    if any(t in op for t in ("UBLKCP", "SYNCS", "CCTL", "STRONG.SYS", "MEMBAR", "FENCE", "ATOM", ".128")):
        lines.append(f"  {op}: {v}")
lines.append("")
lines.append("## per kernel (families)")
for k in sorted(per):
    if not any(t in k for t in ("mega", "spmv", "vec_kernel", "merge", "dep")):
        continue
    c = per[k]
    lines.append(f"{k}\n    total {sum(c.values())}  " + "  ".join(f"{p}:{fam(c, p)}" for p in interesting if fam(c, p)))
path = os.path.join(ROOT, "profiles", "sass_summary.txt")
open(path, "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:14]))
