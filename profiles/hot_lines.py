"""Per-source-line instruction / stall-sample shares of one kernel in an .ncu-rep (read here, no GPU).

usage: python profiles/hot_lines.py report.ncu-rep libheist_b200.so kernel-substring [top [stall|inst [mangled]]]
Joins the ncu SASS source page with nvdisasm line info of the same build (-lineinfo).  `mangled`: substring of the
.text section of the instantiation that was captured (e.g. k_walkILi1ELi1E) -- templates have several, and their
offsets overlap."""
import collections, csv, glob, io, os, re, subprocess, sys, tempfile


def main():
    rep, so, kern = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    by_stall = len(sys.argv) > 5 and sys.argv[5] == "stall"   # sort by stall samples instead of instructions
    mangled = sys.argv[6] if len(sys.argv) > 6 else kern
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + kern, "-c", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, data = rows[1], rows[2:]
    for i, r in enumerate(data):   # several launches of the kernel in the report: the first one
        if r and r[0] == "Kernel Name":
            data = data[:i]
            break
    ia, iex, ismp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    ithr = hdr.index("Avg. Threads Executed")
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
    cubin = glob.glob(tmp + "/*.cubin")[0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
    a2l, cur, inside = {}, None, False
    for l in dis:
        if l.startswith("//-") and ".text." in l:
            inside = mangled in l
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+\S", l)
        if m and cur:
            a2l[int(m.group(1), 16)] = cur
    base = int(data[0][ia], 16)
    agg, smp, thr = collections.Counter(), collections.Counter(), collections.Counter()
    for r in data:
        ln = a2l.get(int(r[ia], 16) - base, ("?", 0))
        agg[ln] += int(r[iex]); smp[ln] += int(r[ismp]); thr[ln] += float(r[ithr]) * int(r[iex])
    tot, tots = sum(agg.values()), max(1, sum(smp.values()))
    print(f"total warp-instructions {tot}, samples {tots}")
    for ln, c in sorted(agg.items(), key=lambda x: -(smp[x[0]] if by_stall else x[1]))[:top]:
        print(f"{ln[0]}:{ln[1]:4d}  inst {c / tot * 100:5.2f}%  stall-samples {smp[ln] / tots * 100:5.2f}%  lanes {thr[ln] / max(c, 1):4.1f}")


if __name__ == "__main__":
    main()
