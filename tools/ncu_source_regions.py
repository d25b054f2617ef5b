#!/usr/bin/env python
"""Per-region summary of an `ncu --page source --csv` export: samples, instructions executed, shared-memory wavefronts and the
dominant stall reasons between given SASS offsets.  usage: ncu_source_regions.py file.csv off0 off1 off2 ... (hex offsets)"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
base = int(rows[2][0], 16)
cuts = [int(x, 16) for x in sys.argv[2:]] + [1 << 30]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot_s = sum(int(r[ix["# Samples"]]) for r in rows[2:] if len(r) > 10)
tot_i = sum(int(r[ix["Instructions Executed"]]) for r in rows[2:] if len(r) > 10)
print(f"total samples {tot_s} instructions {tot_i}")
lo = 0
for hi in cuts:
    sel = [r for r in rows[2:] if len(r) > 10 and lo <= int(r[0], 16) - base < hi]
    if sel:
        s = sum(int(r[ix["# Samples"]]) for r in sel); n = sum(int(r[ix["Instructions Executed"]]) for r in sel)
        wf = sum(int(r[ix["L1 Wavefronts Shared"]]) for r in sel); wfi = sum(int(r[ix["L1 Wavefronts Shared Ideal"]]) for r in sel)
        st = sorted(((sum(int(r[ix[k]]) for r in sel), k) for k in stalls), reverse=True)[:4]
        print(f"{lo:#7x}-{min(hi, int(sel[-1][0],16)-base+16):#7x} samples {100*s/tot_s:5.1f}% instr {100*n/tot_i:5.1f}% ({n/1e6:7.1f}M) shwf {wf/1e6:6.1f}M ideal {wfi/1e6:6.1f}M  " +
              " ".join(f"{k[6:]}={100*v/max(s,1):.0f}%" for v, k in st))
    lo = hi
