#!/usr/bin/env python
"""Per-source-line stall samples of one kernel launch in an .ncu-rep: joins ncu's SASS page (samples per instruction) with
nvdisasm's line info of the same cubin (extracted from libcalico_b200.so, which must be the build that was profiled).

  python scripts/ncu_lines.py gpurun_out/prof.ncu-rep 'cr_level_kernel<0>' [launch_index] [top_n]
"""
import csv, io, os, re, subprocess, sys, tempfile, collections

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sass_lines(mangled_filter):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "calico_b200", "libcalico_b200.so")], cwd=tmp, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    funcs, cur, line = {}, None, None
    for ln in out.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
        if m:
            cur = m.group(1); funcs[cur] = {}; line = None; continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            line = (os.path.basename(m.group(1)), int(m.group(2))); continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m and cur:
            funcs[cur][int(m.group(1), 16)] = (line, m.group(2).strip())
    return funcs


def main():
    rep, kre = sys.argv[1], sys.argv[2]
    idx = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
    filt = [] if kre == "-" else ["--kernel-name", "regex:" + kre]     # "-" = no name filter (reports holding a single kernel)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", *filt, "--launch-skip", str(idx), "--launch-count", "1",
                          "--print-kernel-base", "mangled"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h = [i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r][0]
    hdr = rows[h]
    ai, si, ki = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples")
    mangled = None
    for r in rows[:h]:
        for c in r:
            m = re.search(r"(_ZN\S+)", c)
            if m:
                mangled = m.group(1)
    inst = []
    seen = set()
    for r in rows[h + 1:]:
        try:
            a = int(r[ai], 16)
        except (ValueError, IndexError):
            continue
        if a in seen:
            continue
        seen.add(a)
        inst.append((a, float(r[ki] or 0), r[si]))
    a0 = min(a for a, _, _ in inst)
    funcs = sass_lines(mangled)
    # pick the function whose instruction text matches best at the first few offsets
    best, score = None, -1
    for name, tab in funcs.items():
        sc = sum(1 for a, _, src in inst[:400] if (a - a0) in tab and tab[a - a0][1].split()[0] == src.strip().split()[0])
        if sc > score:
            best, score = name, sc
    tab = funcs[best]
    by_line = collections.Counter()
    tot = sum(s for _, s, _ in inst)
    for a, s, _ in inst:
        ln = tab.get(a - a0, (None, ""))[0]
        by_line[ln] += s
    print(f"kernel {best}  launch {idx}: {tot:.0f} samples, matched {score}/400 instructions")
    srcs = {}
    for (ln, s) in by_line.most_common(top):
        if ln is None:
            print(f"{100 * s / tot:5.1f}%  ?")
            continue
        f, n = ln
        if f not in srcs:
            p = os.path.join(ROOT, "calico_b200", "csrc", f)
            srcs[f] = open(p).read().splitlines() if os.path.exists(p) else []
        text = srcs[f][n - 1].strip()[:110] if n - 1 < len(srcs[f]) else ""
        print(f"{100 * s / tot:5.1f}%  {f}:{n}  {text}")


if __name__ == "__main__":
    main()
