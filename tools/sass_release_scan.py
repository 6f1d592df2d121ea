"""Static check of the shipped SASS: a consumer's mbarrier arrive (the one that hands a shared-memory
slot / buffer back to the TMA producer) must not issue before the loads from that slot have returned.
For every SYNCS.ARRIVE...A1T0 of every kernel that has one, each LDS since the previous such arrive
must have one of its destination registers read (or a MEMBAR crossed) before the arrive: in-order issue
then guarantees the load's data is back.  python tools/sass_release_scan.py [library.so]"""
import os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def scan(text: str):
    """text: `cuobjdump -sass` output.  Returns (violations, loads checked, report lines)."""
    funcs, cur = {}, None
    for line in text.splitlines():
        m = re.match(r"\s+Function : (\S+)", line)
        if m:
            cur = funcs.setdefault(m.group(1), [])
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and cur is not None:
            cur.append((m.group(1), m.group(2)))
    bad = total = 0
    report = []
    for fn, L in sorted(funcs.items()):
        arrives = [i for i, (_, ins) in enumerate(L) if "SYNCS.ARRIVE" in ins and "A1T0" in ins]
        if not arrives:
            continue
        checked = issues = 0
        for n, i in enumerate(arrives):
            lo = arrives[n - 1] + 1 if n else 0
            for k in range(lo, i):
                m = re.search(r"\bLDS(?:\.\w+)* R(\d+), \[", L[k][1])
                if not m:
                    continue
                width = 4 if ".128" in L[k][1] else 2 if ".64" in L[k][1] else 1
                regs = [f"R{int(m.group(1)) + t}" for t in range(width)]
                ok = False
                for t in range(k + 1, i):
                    body = L[t][1]
                    if "MEMBAR" in body:
                        ok = True
                        break
                    words = body.split()
                    op = words[1] if words and words[0].startswith("@") and len(words) > 1 else (words[0] if words else "")
                    # sources: everything after the first comma; stores read all their register operands
                    src = body if op.startswith("ST") else (body.split(",", 1)[1] if "," in body else "")
                    if any(re.search(r"\b" + r + r"\b", src) for r in regs):
                        ok = True
                        break
                checked += 1
                if not ok:
                    issues += 1
                    report.append(f"  {fn}: arrive at {L[i][0]} can issue before the load at {L[k][0]} returns: {L[k][1]}")
        total += checked
        bad += issues
        report.append(f"{fn[:90]}: {len(arrives)} consumer arrives, {checked} loads checked, {issues} released before use")
    return bad, total, report


if __name__ == "__main__":
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "ultrasphere_b200", "lib", "libultrasphere_b200.so")
    bad, total, report = scan(subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout)
    print("\n".join(report))
    print("release-before-use:", bad, "of", total)
    sys.exit(1 if bad else 0)
