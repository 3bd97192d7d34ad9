"""Decode an NM_TC_TRACE dump (CTA 0 event log of the tcgen05 MLP kernel) into a per-layer timeline summary."""
import sys
import numpy as np

raw = np.fromfile(sys.argv[1], dtype=np.uint64)
rec = raw[1:].reshape(-1, 5)
rec = rec[rec[:, 0] != 0]
n = rec.shape[0]
hdr = rec[:, 0]
kind, rid, idx, gl = (hdr >> 48) & 0xFFFF, (hdr >> 40) & 0xFF, (hdr >> 24) & 0xFFFF, hdr & 0xFFFFFF
t = rec[:, 1:].astype(np.int64)
t0 = t[t > 0].min()
t = np.where(t > 0, t - t0, -1)
lo, hi = int(sys.argv[2]) if len(sys.argv) > 2 else 11, int(sys.argv[3]) if len(sys.argv) > 3 else 24
print(f"{n} records; layer-instances {lo}..{hi}")
for g in range(lo, hi):
    m = gl == g
    iss = [(int(rid[i]), int(idx[i]), *[int(x) for x in t[i]]) for i in np.nonzero(m & (kind == 1))[0]]
    epi = [(int(rid[i]), int(idx[i]), *[int(x) for x in t[i]]) for i in np.nonzero(m & (kind == 2))[0]]
    if not iss and not epi:
        continue
    start = min([x[2] for x in iss] + [x[2] for x in epi])
    end = max([x[5] for x in iss] + [x[5] for x in epi])
    print(f"--- layer-instance {g}: span {start}..{end} ({end - start} cyc)")
    for w, b, a, bb, c, d in sorted(iss, key=lambda x: x[2]):
        print(f"   issuer{w} blk{b:4d}: start {a - start:6d}  group-wait {bb - a:5d}  w_full-wait {c - bb:5d}  issue {d - c:5d}  -> done {d - start:6d}")
    for s_, nn, a, bb, c, d in sorted(epi, key=lambda x: x[2]):
        print(f"   epi set{s_} chunk{nn}: start {a - start:6d}  d_full-wait {bb - a:5d}  ld+math {max(c - bb, 0):5d}  kb_free+st+arrive {d - max(c, bb):5d}  -> arrive {d - start:6d}")
