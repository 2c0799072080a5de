"""Per-CUDA-source-line view of an ncu report that was captured with `--set full --import-source on`:

    ncu -i report.ncu-rep --page source --csv --print-source cuda,sass > src.csv
    python tools/ncu_source_lines.py src.csv [path/to/file.cu] [top]

Aggregates the SASS rows under every source line: warp-stall samples (share of the kernel), executed warp instructions,
the dominant stall reasons, plus the kernel's opcode mix.  Used to find the hot lines of the attention / MLP kernels."""
import collections
import csv
import sys


def main(path, src_path=None, top=30):
    rows = list(csv.reader(open(path)))
    hdr = cur_line = cur_file = fn = None
    agg = collections.OrderedDict()
    ops = collections.Counter()
    tot = totinst = 0
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if r[0] == "Function Name":
            fn = r[1]
            continue
        if r[0] != "":
            cur_line = (cur_file, r[0])
            continue
        if hdr is None or len(r) < len(hdr) - 5:
            continue
        try:
            smp, inst = int(r[hdr.index("# Samples")]), int(r[hdr.index("Instructions Executed")])
        except Exception:
            continue
        a = agg.setdefault(cur_line, [0, 0, collections.Counter()])
        a[0] += smp
        a[1] += inst
        tot += smp
        totinst += inst
        toks = r[3].split()
        if toks:
            op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
            ops[op.split(".")[0]] += inst
        for i, h in enumerate(hdr):
            if h.startswith("stall_") and "Not Issued" not in h:
                try:
                    a[2][h[6:]] += int(r[i] or 0)
                except Exception:
                    pass
    print((fn or "?")[:100])
    print("stall samples %d, warp instructions %d" % (tot, totinst))
    print("opcode mix:", ", ".join("%s %.1f%%" % (k, 100.0 * v / max(totinst, 1)) for k, v in ops.most_common(16)))
    src = {}
    if src_path:
        for i, line in enumerate(open(src_path)):
            src[str(i + 1)] = line.rstrip()
    base = src_path.split("/")[-1] if src_path else None
    for (f, ln), (s, inst, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        text = src.get(ln, "") if (base and f == base) else ""
        print("%5.1f%% smp %5.1f%% inst  %s:%s  %-70s %s" % (100.0 * s / max(tot, 1), 100.0 * inst / max(totinst, 1), f[:14], ln,
                                                             text.strip()[:70], " ".join("%s=%d" % kv for kv in st.most_common(3))))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None, int(sys.argv[3]) if len(sys.argv) > 3 else 30)
