"""Summarise an ncu report's source page: stall reasons, opcode histogram, top stalled instructions.
Usage: python tools/ncu_src_summary.py report.ncu-rep [ntop]"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
byop, bystall, execs = collections.Counter(), collections.Counter(), collections.Counter()
tot = 0
for r in data:
    n = int(r[ix["# Samples"]] or 0)
    tot += n
    src = r[ix["Source"]].strip()
    op = (src.split()[1] if src.startswith("@") else src.split()[0]).split(".")[0]
    byop[op] += n
    execs[op] += int(r[ix["Instructions Executed"]] or 0)
    for h in stall_cols:
        bystall[h] += int(r[ix[h]] or 0)
print(rows[0][1][:90])
print("samples", tot, "warp-instructions", sum(execs.values()))
print("stalls:", ", ".join("%s %d" % (k[6:], v) for k, v in bystall.most_common(10)))
print("opcodes (samples / executed):", ", ".join("%s %d/%d" % (k, v, execs[k]) for k, v in byop.most_common(16)))
for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[:ntop]:
    st = sorted(((h[6:], int(r[ix[h]] or 0)) for h in stall_cols), key=lambda kv: -kv[1])[:2]
    print(r[ix["Address"]][-5:], r[ix["# Samples"]].rjust(6), r[ix["Instructions Executed"]].rjust(9), r[ix["Source"]][:60].ljust(60), st)
