"""Per-kernel instruction mix and stall picture from the SASS source page of an ncu capture taken with --import-source on:

    ncu -i gpurun_out/prof_full.ncu-rep --page source --csv > /tmp/all_src.csv
    python tools/ncu_source_mix.py /tmp/all_src.csv <tag> > profiles/r01_<tag>_source_mix.md

For every captured launch: warp instructions executed, the share of the sampled stall reasons, the opcodes that execute most
and the opcodes the warps are sampled on most (a large samples/executed ratio = an instruction warps wait ON or AT), and the
shared-memory wavefronts against their ideal.  Needs no GPU."""
import collections, csv, sys

path, tag = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
rows = list(csv.reader(open(path)))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
seen = collections.Counter()
print("# Round 1, %s -- SASS-level instruction mix and stall samples per kernel (`ncu --page source`, tools/ncu_source_mix.py)\n" % tag)
print("One forward + decode at C2 (model 101, 513x513, OS16, batch 64, bf16), launches in step order.  `exec %` = share of the warp")
print("instructions executed, `smp %` = share of the warp-state samples taken at that opcode.\n")
for n, s in enumerate(starts):
    e = starts[n + 1] if n + 1 < len(starts) else len(rows)
    name = rows[s][1].split("(")[0].replace("void ", "").replace("pn::", "") + ("".join(rows[s][1].split("(")[1:3]) if "<" in rows[s][1] else "")
    name = rows[s][1]
    short = name.split("(CUtensorMap")[0].split("(pn::")[0].split("(pn_map")[0].replace("void ", "").replace("pn::", "").replace("(int)", "").replace("(bool)", "")
    hdr, data = rows[s + 1], [r for r in rows[s + 2:e] if len(r) > 10]
    ix = {h: i for i, h in enumerate(hdr)}
    if "Instructions Executed" not in ix or not data:
        continue
    src0 = data[0][ix["Source"]].strip()
    if not src0 or not src0.split()[0].replace("@", "").replace("!", "").replace("P", "").replace("U", "").isalnum() or src0.startswith(("//", ".", "{")):
        continue                                       # the PTX / CUDA-C view of the same launch
    if any(k in src0 for k in ("ld.param", ".reg", "mov.u32", "cvta")):
        continue
    sig = (short, sum(float(r[ix["Instructions Executed"]] or 0) for r in data if r[ix["Instructions Executed"]].replace(".", "").isdigit()))
    if n and sig == globals().get("_last_sig"):
        continue                                       # ncu lists a launch once per source view; identical counts = same launch
    globals()["_last_sig"] = sig
    seen[short] += 1
    launch_no = globals().get("_launch_no", -1) + 1
    globals()["_launch_no"] = launch_no

    def f(r, k):
        try:
            return float(r[ix[k]])
        except (ValueError, KeyError, IndexError):
            return 0.0
    tot_i = sum(f(r, "Instructions Executed") for r in data)
    tot_s = sum(f(r, "# Samples") for r in data)
    if tot_i == 0 or tot_s == 0:
        continue
    stalls = {h: sum(f(r, h) for r in data) for h in hdr if h.startswith("stall_") and "Not Issued" not in h}
    mix, smp = collections.Counter(), collections.Counter()
    for r in data:
        t = r[ix["Source"]].strip().split()
        if not t:
            continue
        op = (t[1] if t[0].startswith("@") and len(t) > 1 else t[0]).split(".")[0].rstrip(";")
        mix[op] += f(r, "Instructions Executed")
        smp[op] += f(r, "# Samples")
    wf, wfi = sum(f(r, "L1 Wavefronts Shared") for r in data), sum(f(r, "L1 Wavefronts Shared Ideal") for r in data)
    print("## %d. `%s`%s\n" % (launch_no, short, " (launch %d of this kernel)" % seen[short] if seen[short] > 1 else ""))
    print("%.1f M warp instructions over %d SASS lines; %d samples.  Stalls: %s.  Shared-memory wavefronts %.1f M (ideal %.1f M).\n" % (
        tot_i / 1e6, len(data), tot_s, ", ".join("%s %.0f %%" % (k[6:], 100 * v / tot_s) for k, v in sorted(stalls.items(), key=lambda x: -x[1])[:6]),
        wf / 1e6, wfi / 1e6))
    print("| opcode | exec % | smp % |\n|---|---|---|")
    for op, v in mix.most_common(12):
        print("| %s | %.1f | %.1f |" % (op, 100 * v / tot_i, 100 * smp[op] / tot_s))
    hot = [(op, v) for op, v in smp.most_common(6) if op not in dict(mix.most_common(12))]
    for op, v in hot:
        print("| %s | %.1f | %.1f |" % (op, 100 * mix[op] / tot_i, 100 * v / tot_s))
    print()
