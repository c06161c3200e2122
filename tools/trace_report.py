"""Reads a BVC_REC_TRACE dump of the recurrent kernel and prints a per-phase timeline of one frame.

    BVC_REC_TRACE=/tmp/t.bin python tools/run_recurrent_once.py ... ; python tools/trace_report.py /tmp/t.bin [frame]

Events (ns, %globaltimer): copy thread 0 phase start, 1 weight prefetch issued, 2 phase barrier passed, 3 activations
issued, 4 all weights issued; MMA thread 5 first activation chunk landed, 6 last chunk landed, 7 last MMA issued;
epilogue 8 first accumulator complete, 9 last accumulator complete, 10 last accumulator read from TMEM, 11 last partials
sent and signalled, 14 peers' partials received, 15 partials summed, 12 last tile finalised, 13 arrived at the phase barrier.
"""
import sys
import numpy as np

path = sys.argv[1]
frame = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 3
CLK = "--clk" in sys.argv     # stamps are clock64 of the CTA's SM (BVC_REC_DEBUG=64): per-CTA intervals from the barrier (event 2)
hdr = np.fromfile(path, dtype=np.int32, count=4)
n_cta, n_fr, n_ph, n_ev = [int(v) for v in hdr]
d = np.fromfile(path, dtype=np.uint64, offset=16).reshape(n_cta, n_fr, n_ph, n_ev).astype(np.float64)
d[d == 0] = np.nan
fr = d[:, frame]
if CLK:
    ev = [("Aiss", 3), ("c0", 16), ("c1", 17), ("c2", 18), ("c3", 19),
          ("mmaIss", 7), ("acc0", 8), ("accL", 9), ("tmem", 10), ("stores", 0), ("sent", 11),
          ("recv", 14), ("sum", 15), ("freed", 1), ("f.add", 20), ("f.elu", 21), ("f.img", 22), ("f.out", 23), ("fin", 12), ("barred", 36), ("arr", 13)]
    print("frame %d: %d CTAs; clk after the CTA's own phase-barrier pass (median / max over busy CTAs)" % (frame, n_cta))
    print("phase busy | " + " ".join("%6s" % n for n, _ in ev))
    for ph in range(n_ph):
        if np.all(np.isnan(fr[:, ph, 13])):
            continue
        busy = ~np.isnan(fr[:, ph, 7]) & ~np.isnan(fr[:, ph, 2])
        if not busy.any():
            continue
        ref = fr[busy, ph, 2]
        med = [np.nanmedian(fr[busy, ph, e] - ref) if e < n_ev else np.nan for _, e in ev]
        mx = [np.nanmax(fr[busy, ph, e] - ref) if e < n_ev else np.nan for _, e in ev]
        print("%3d  %4d  | " % (ph, int(busy.sum())) + " ".join("%6.0f" % v for v in med))
        if n_ev > 30 and not np.all(np.isnan(fr[busy, ph, 24])):
            # weight ring probe on chunk 0 of the LAST tile of a multi-tile phase: copy warp 24 starts waiting for the slot, 25 slot
            # free, 26 copy issued; MMA warp 30 previous tile's chunk 0 issued, 27 starts waiting for the chunk, 28 chunk landed, 29 its MMAs issued
            print("       weight ring (last tile, chunk 0): prev chunk-0 MMAs issued %.0f | copy: wait slot %.0f, slot free %.0f, issued %.0f | MMA: wait %.0f, landed %.0f, issued %.0f" % tuple(
                np.nanmedian(fr[busy, ph, e] - ref) for e in (30, 24, 25, 26, 27, 28, 29)))
        print("       max | " + " ".join("%6.0f" % v for v in mx))
    sys.exit(0)
prev_done = np.nanmax(d[:, frame - 1, :, 13]) if frame > 0 else np.nanmin(fr[:, 0, 0])
order = [2, 3, 4, 5, 6, 7, 8, 9, 10, 0, 11, 14, 15, 1, 12, 13]
names = ["bar", "Aiss", "Wall", "A0", "Alast", "mma", "acc0", "accL", "tmem", "stores", "sent", "recv", "sum", "freed", "fin", "arr"]
print("frame %d: %d CTAs; times in us relative to the previous phase's last barrier arrival" % (frame, n_cta))
print("phase  dur   busy | " + " ".join("%6s" % n for n in names) + "   (median over busy CTAs)   last-arr-cta")
total = 0.0
for ph in range(n_ph):
    if np.all(np.isnan(fr[:, ph, 13])):
        continue
    done = np.nanmax(fr[:, ph, 13])
    busy = ~np.isnan(fr[:, ph, 8])
    med = [np.nanmedian(fr[busy, ph, e]) - prev_done if busy.any() else np.nan for e in order]
    mx = [np.nanmax(fr[busy, ph, e]) - prev_done if busy.any() else np.nan for e in order]
    print("%3d  %6.2f  %4d | " % (ph, (done - prev_done) / 1e3, int(busy.sum())) + " ".join("%6.2f" % (v / 1e3) for v in med) +
          "   cta %d" % int(np.nanargmax(fr[:, ph, 13])))
    print("                  max " + " ".join("%6.2f" % (v / 1e3) for v in mx))
    total += (done - prev_done) / 1e3
    prev_done = done
print("frame total %.2f us" % total)
