"""Hot CUDA source lines of a kernel from an ncu report (needs -lineinfo and --import-source on):
    python profiles/ncu_source_hot.py rep.ncu-rep <kernel-regex> [top]"""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                      "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
cur_file, hdr, out, seen_kernel = None, None, [], 0
for r in rows:
    if len(r) >= 2 and r[0] == "Kernel Name":
        seen_kernel += 1
        if seen_kernel > 1:
            break
    if len(r) >= 2 and r[0] == "File Name":
        cur_file = r[1]; continue
    if len(r) >= 2 and r[0] == "Line No":
        hdr = r; continue
    if hdr and len(r) > 10 and r[0].strip().isdigit():
        col = {h: i for i, h in enumerate(hdr)}
        num = lambda k: float(r[col[k]].replace(",", "") or 0) if r[col[k]] not in ("-", "") else 0.0
        try:
            s, inst = num("# Samples"), num("Instructions Executed")
        except ValueError:          # a source line with embedded quotes (inline asm) breaks the CSV row: skip it
            continue
        try:
            stalls = {k: num(k) for k in hdr if k.startswith("stall_") and "Not Issued" not in k}
        except ValueError:
            continue
        out.append((s, inst, (cur_file or "?").split("/")[-1], r[0], r[1].strip()[:88], stalls))
tot = sum(o[0] for o in out) or 1
print("total samples %d, total warp insts %d" % (tot, sum(o[1] for o in out)))
for s, inst, f, ln, src, st in sorted(out, reverse=True)[:top]:
    best = sorted(st.items(), key=lambda kv: -kv[1])[:2]
    print("%5.1f%% inst=%8d %s:%s  %-88s %s" % (100 * s / tot, inst, f, ln, src, " ".join("%s=%d" % (k[6:], v) for k, v in best if v)))
