"""SASS opcode summary per kernel of the in-tree library (how profiles/r0N_sass_summary.txt is made).
usage: python scripts/sass_summary.py [lib.so] > profiles/r02_sass_summary.txt"""
import collections, os, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ocaml-hnsw_b200", "libhnsw_b200.so")
KEEP = ("UBLKCP", "UBLKPF", "SYNCS", "UTC", "LDTM", "STTM", "UTMA", "FFMA2", "FADD2", "LDGSTS", "BAR", "ATOMG", "ATOMS", "REDG", "LDG", "LDS", "STS",
        "SHFL", "VOTE", "HMMA", "CCTL")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
demangle = lambda s: subprocess.run(["c++filt", s], capture_output=True, text=True).stdout.strip()
kern, ops = None, None
res = []
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = demangle(m.group(1)); ops = collections.Counter(); res.append((kern, ops)); continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and ops is not None:
        ops[m.group(1)] += 1
print(f"# SASS opcode summary of {os.path.relpath(lib)} (cuobjdump -sass, sm_100a)")
print("# per kernel: total instructions, then the opcodes that prove where data moves and what computes")
print("# (UBLKCP = cp.async.bulk 1-D TMA copy, UBLKPF = bulk L2 prefetch, SYNCS.* = mbarrier ops, UTCHMMA = tcgen05.mma,")
print("#  LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, FFMA2 / FADD2 = packed fp32, LDGSTS = cp.async, BAR = named barrier)\n")
for kern, ops in res:
    name = re.sub(r"\(.*", "", kern)
    if "cub::" in name or "thrust::" in name:
        continue
    agg = collections.Counter()
    for op, c in ops.items():
        base = op.split(".")[0]
        if base in ("SYNCS", "UBLKCP", "UBLKPF") or base.startswith("UTC"):
            base = ".".join(op.split(".")[:4]) if base == "SYNCS" else ".".join(op.split(".")[:3])
        if any(base.startswith(k) for k in KEEP):
            agg[base] += c
    print(f"{name}: {sum(ops.values())} instructions")
    print("    " + "  ".join(f"{k}={v}" for k, v in sorted(agg.items())))
