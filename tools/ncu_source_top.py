"""Top source lines by warp-stall samples from `ncu --page source --csv` (first kernel or kernel index)."""
import csv
import sys


def main(path, which=0, top=25):
    kernels = []
    cur = None
    with open(path, newline="") as f:
        for row in csv.reader(f):
            if not row:
                continue
            if row[0] == "Kernel Name":
                cur = {"name": row[1], "hdr": None, "rows": []}
                kernels.append(cur)
            elif cur is not None and cur["hdr"] is None:
                cur["hdr"] = row
            elif cur is not None:
                cur["rows"].append(row)
    k = kernels[which]
    hdr = k["hdr"]
    i_src, i_smp = hdr.index("Source"), hdr.index("# Samples")
    i_inst = hdr.index("Instructions Executed")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    rows = []
    for r in k["rows"]:
        try:
            n = int(r[i_smp])
        except Exception:
            continue
        rows.append((n, r))
    tot = sum(n for n, _ in rows)
    print("kernel %d of %d: %s  total samples %d" % (which, len(kernels), k["name"][:60], tot))
    for n, r in sorted(rows, key=lambda x: -x[0])[:top]:
        stalls = sorted(((int(r[i] or 0), hdr[i]) for i in stall_cols), reverse=True)[:2]
        print("%6d %5.1f%%  inst=%-7s %-70s %s" % (n, 100.0 * n / max(tot, 1), r[i_inst], r[i_src].strip()[:70],
                                               " ".join("%s=%d" % (h[6:], v) for v, h in stalls if v)))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0, int(sys.argv[3]) if len(sys.argv) > 3 else 25)
