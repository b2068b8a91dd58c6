"""Top CUDA source lines of a kernel by executed warp instructions (first matching launch):
    python profiles/ncu_source_inst.py rep.ncu-rep <kernel-regex> [top]"""
import csv, io, subprocess, sys, collections
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                      "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, seen, agg, src_of = None, 0, collections.Counter(), {}
samples = collections.Counter()
for r in rows:
    if len(r) >= 2 and r[0] == "Kernel Name":
        seen += 1
        if seen > 1: break
    if len(r) >= 2 and r[0] == "Line No":
        hdr = r; continue
    if hdr and len(r) > 10 and r[0].strip().isdigit():
        col = {h: i for i, h in enumerate(hdr)}
        off = len(r) - len(hdr)                      # source text with embedded commas/quotes shifts the columns
        def f(k):
            try:
                v = r[col[k] + off]
                return float(v.replace(",", "") or 0) if v not in ("-", "") else 0.0
            except (ValueError, IndexError):
                return 0.0
        key = (r[0], ",".join(r[1:2 + off]).strip()[:100])
        agg[key] += f("Instructions Executed"); samples[key] += f("# Samples")
tot = sum(agg.values()) or 1; ts = sum(samples.values()) or 1
print("total warp insts %d, samples %d" % (tot, ts))
for (ln, src), n in agg.most_common(top):
    print("%5.1f%% inst %5.1f%% smp  L%-4s %s" % (100 * n / tot, 100 * samples[(ln, src)] / ts, ln, src))
