"""SASS evidence for the hot kernels: per kernel the counts of the mnemonics that prove the design
(TMA bulk copies, mbarrier ops, packed FP32, 128-bit streaming stores, NO local-memory traffic) and
excerpts of the producer loop and of one consumer node body.  Reads the built .so with cuobjdump.

    python tools/sass_excerpt.py > profiles/r02_sass_hot.txt"""
import os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "ultrasphere_b200", "lib", "libultrasphere_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs = {}
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        funcs[cur] = []
    elif cur and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
        funcs[cur].append(re.sub(r"\s*/\* 0x[0-9a-f]+ \*/", "", line).rstrip())
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
KEYS = ("UBLKCP", "SYNCS.ARRIVE", "SYNCS.PHASECHK", "VOTE", "ELECT", "FFMA2", "FMUL2", "FADD2", "DFMA", "MUFU", "LDS.128", "STS.128", "STG.E.EF.128", "LDL", "STL", "LDG", "BAR.SYNC")
hot = [n for n in funcs if re.search(r"cartesian_(tiled|ring)|wreduce_tma", n)]
print("# SASS of the hot kernels (cuobjdump -sass of the shipped libultrasphere_b200.so)\n")
print("Counts per kernel.  UBLKCP = cp.async.bulk (TMA), SYNCS = mbarrier, FFMA2/FMUL2/FADD2 = packed FP32,")
print("STG.E.EF.128 = 128-bit evict-first store; LDL/STL = local memory (the deferred-branch stack of round 1): must be 0.\n")
print("| kernel | instr | " + " | ".join(KEYS) + " |")
print("|---|---|" + "---|" * len(KEYS))
for n in sorted(hot, key=demangle):
    body = funcs[n]
    ops = [re.sub(r"^\s*/\*[0-9a-f]+\*/\s*(@!?U?P\d+\s+)?", "", l) for l in body]
    print(f"| `{demangle(n)[:88]}` | {len(body)} | " + " | ".join(str(sum(1 for o in ops if o.startswith(k))) for k in KEYS) + " |")


def excerpt(name_re, start_re, n_lines, title):
    for n in sorted(funcs):
        if re.search(name_re, n):
            body = funcs[n]
            for i, l in enumerate(body):
                if re.search(start_re, l):
                    print(f"\n## {title}\n`{demangle(n)}` from offset {l.split('*/')[0].split('/*')[1]}\n```")
                    print("\n".join(body[max(0, i - 6): i + n_lines]))
                    print("```")
                    return


excerpt(r"from_cartesian_tiledIfLi4ELi1E", r"UBLKCP", 26, "producer loop of the tile kernels: TMA bulk copies issued from uniform registers, four rows per iteration")
excerpt(r"from_cartesian_tiledIfLi4ELi1E", r"VOTE\.ALL", 10, "consumer wait: mbarrier try_wait, warp vote, no divergence")
excerpt(r"from_cartesian_tiledIfLi4ELi1E", r"MUFU\.RCP", 70, "float32 chain node (V = 4): packed FFMA2 polynomial, MUFU.RCP / MUFU.RSQ seeds, LDS.128 in, STG.E.EF.128 out")
excerpt(r"from_cartesian_tiledIdLi2ELi4E", r"MUFU\.RCP64H", 60, "float64 create_hopf(3) walk (compile-time recursion): descriptors as constant-bank operands, no stack")
excerpt(r"from_cartesian_ringIfLi2ELi8ELi2E", r"STS\.128", 12, "row-ring kernel: a deferred branch parked in shared memory (STS.128), not in local memory")
excerpt(r"wreduce_tma", r"UBLKCP", 12, "weighted reduction: one TMA bulk copy per slab")
