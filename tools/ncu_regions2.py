"""Per CUDA-source-line totals (instructions executed, warp samples, top stall reasons) from an ncu --import-source report
(`--print-source cuda,sass`: a source line row followed by its SASS rows).  Needs no GPU.
usage: python tools/ncu_regions2.py report.ncu-rep [min_sample_pct]"""
import csv, subprocess, sys, collections, io
rep = sys.argv[1]; minp = float(sys.argv[2]) if len(sys.argv) > 2 else 0.8
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = [ln[1:-1].split('","') if ln.startswith('"') else [ln] for ln in out.splitlines()]   # source text holds unescaped quotes: split by hand
fpath = ""; hdr = None; lines = []          # (file, line, text, exec, samples, stalls dict)
for r in rows:
    if not r: continue
    if r[0] == "File Path": fpath = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; ix = {}; [ix.setdefault(h, i) for i, h in enumerate(hdr)]; stall = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]; continue
    if hdr is None or len(r) < len(hdr): continue
    if len(r) > len(hdr): r = [r[0], '","'.join(r[1:len(r) - len(hdr) + 2])] + r[len(r) - len(hdr) + 2:]   # commas / quotes inside the source text
    if r[0] != "":                            # a CUDA source line (its totals are the sum of its SASS rows)
        def f(k):
            try: return float(r[ix[k]] or 0)
            except ValueError: return 0.0
        lines.append((fpath, int(r[0]), r[1].strip(), f("Instructions Executed"), f("# Samples"), {s[6:]: f(s) for s in stall}))
tot_e = sum(l[3] for l in lines) or 1; tot_s = sum(l[4] for l in lines) or 1
print("total warp instructions %.1f M, samples %d" % (tot_e / 1e6, tot_s))
print("| file:line | exec % | smp % | top stalls | source |\n|---|---|---|---|---|")
for l in sorted(lines, key=lambda l: (l[0], l[1])):
    if 100 * l[4] / tot_s < minp and 100 * l[3] / tot_e < minp: continue
    st = sorted(l[5].items(), key=lambda x: -x[1])[:3]; ss = sum(l[5].values()) or 1
    print("| %s:%d | %.1f | %.1f | %s | `%s` |" % (l[0], l[1], 100 * l[3] / tot_e, 100 * l[4] / tot_s, " ".join("%s %.0f%%" % (k, 100 * v / ss) for k, v in st if v), l[2][:90]))
if len(sys.argv) > 3:                                     # ranges "file:a-b,..." -> totals
    print()
    for spec in sys.argv[3].split(","):
        fn, rg = spec.split(":"); a, b = [int(v) for v in rg.split("-")]
        sel = [l for l in lines if l[0] == fn and a <= l[1] <= b]
        st = collections.Counter()
        for l in sel: st.update(l[5])
        ss = sum(st.values()) or 1
        print("%-22s exec %5.1f %%  smp %5.1f %%  %s" % (spec, 100 * sum(l[3] for l in sel) / tot_e, 100 * sum(l[4] for l in sel) / tot_s,
              " ".join("%s %.0f%%" % (k, 100 * v / ss) for k, v in st.most_common(5))))
