"""Per-kernel summary of an ncu report's source page: where the warp samples sit (stall reasons),
the executed opcode mix and the hottest instructions.  python tools/ncu_source_stats.py <report.ncu-rep> [top N]"""
import csv, subprocess, sys

rep = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 20
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
sections, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        sections.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
for sec in sections:
    hdr, body = sec["rows"][0], sec["rows"][1:]
    ix = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = dict.fromkeys(stalls, 0)
    total = inst = 0
    data, mix = [], {}
    for r in body:
        if len(r) < len(hdr):
            continue
        samp, ex = int(r[ix["# Samples"]] or 0), int(r[ix["Instructions Executed"]] or 0)
        total += samp
        inst += ex
        st = {s: int(r[ix[s]] or 0) for s in stalls}
        for s in stalls:
            tot[s] += st[s]
        src = r[ix["Source"]].strip()
        toks = src.split()
        op = (toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")).split(".")[0]
        mix[op] = mix.get(op, 0) + ex
        data.append((r[ix["Address"]][-5:], src, samp, ex, st))
    print(f"## {sec['name']}\nwarp samples {total}, warp instructions executed {inst}")
    print("stall reasons %:", {s[6:]: round(100 * v / max(total, 1), 1) for s, v in sorted(tot.items(), key=lambda x: -x[1]) if v * 200 > total})
    print("opcode mix %:", {k: round(100 * v / max(inst, 1), 1) for k, v in sorted(mix.items(), key=lambda x: -x[1])[:24]})
    for a, src, samp, ex, st in sorted(data, key=lambda x: -x[2])[:top_n]:
        print(" ", a, samp, ex, src[:64].ljust(64), [(k[6:], v) for k, v in sorted(st.items(), key=lambda x: -x[1])[:3] if v])
    print()
