"""Condense an `ncu --page source --csv --print-source sass` export (tens of MB) into the summary kept under profiles/:
stall-reason totals, the SASS lines with the most stall samples, and a census of the SASS mnemonics that matter here.
usage: python tests/ncu_source_summary.py gpurun_out/<tag>_ncu_source_C2.csv [launch index] > profiles/<tag>_sass_stalls_C2.txt"""
import csv
import re
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Address"]          # one section per captured launch
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hdr_i = starts[which]
end = starts[which + 1] - 1 if which + 1 < len(starts) else len(rows)
hdr = rows[hdr_i]
body = [r for r in rows[hdr_i + 1:end] if len(r) == len(hdr)]
col = {n: i for i, n in enumerate(hdr)}
stalls = [n for n in hdr if n.startswith("stall_") and "(Not Issued)" not in n]
tot = Counter()
for r in body:
    for s in stalls:
        tot[s] += int(r[col[s]] or 0)
all_s = sum(tot.values())
print(f"launch {which + 1} of {len(starts)} in the capture; kernel: {rows[hdr_i - 1][1] if rows[hdr_i - 1] else '?'}   SASS instructions: {len(body)}   stall samples: {all_s}")
print("\nstall reasons (all samples):")
for s, v in tot.most_common():
    if v:
        print(f"  {s:<24}{v:>10}  {100.0 * v / all_s:5.1f} %")
print("\ntop 25 SASS lines by samples:   samples  share  executed(warp)  top reason   instruction")
samp = col["# Samples"]
for r in sorted(body, key=lambda r: -int(r[samp] or 0))[:25]:
    top = max(stalls, key=lambda s: int(r[col[s]] or 0))
    print(f"  {int(r[samp]):>8} {100.0 * int(r[samp]) / all_s:5.1f} % {int(r[col['Instructions Executed']]):>12}  {top:<18} {r[col['Source']].strip()}")
mn = Counter()
for r in body:
    m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[col["Source"]])
    if m:
        mn[m.group(1).split(".")[0]] += 1
print("\nSASS census (static instruction counts):")
keys = ["BAR", "WARPSYNC", "REDUX", "RED", "ATOM", "ATOMS", "ATOMG", "MEMBAR", "FENCE", "CCTL", "LDG", "STG", "LDS", "STS", "LDL", "STL",
        "SHFL", "VOTE", "MATCH", "PRMT", "LOP3", "IMAD", "FFMA", "DFMA", "BSSY", "BSYNC", "CALL", "NANOSLEEP", "UTMALDG", "UBLKCP", "HMMA", "UTCMMA"]
print("  " + "  ".join(f"{k}={mn.get(k, 0)}" for k in keys))
ex = Counter()
for r in body:
    m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[col["Source"]])
    if m:
        ex[m.group(1).split(".")[0]] += int(r[col["Instructions Executed"]] or 0)
tot_ex = sum(ex.values())
print("\nexecuted warp instructions by mnemonic (top 15):")
for k, v in ex.most_common(15):
    print(f"  {k:<10}{v:>14}  {100.0 * v / tot_ex:5.1f} %")
