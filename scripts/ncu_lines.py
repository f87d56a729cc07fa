#!/usr/bin/env python
"""Per-source-line view of an ncu report: joins `ncu --page source --csv` (SASS-level
instruction counts and stall samples) with `nvdisasm -g` line info of the in-tree library.

    python scripts/ncu_lines.py gpurun_out/<rep>.ncu-rep <kernel substring> [launch index] [top N]

Prints, per source line, the executed warp-instructions and stall samples (+ the top stall reason).
Needs the SAME build of libdiffphar_b200.so that produced the report.
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "cmd_gen_b200", "libdiffphar_b200.so")


def sass_lines(kernel_sub, want_len=None):
    """SASS (address, text, source line, inline chain) of the library function whose name contains kernel_sub; with
    several template instantiations, the one whose instruction count matches the report's (want_len)."""
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, check=True, capture_output=True)
    found = {}
    for f in sorted(os.listdir(tmp)):
        if not f.endswith(".cubin"):
            continue
        out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        cur_fn, cur_line, res, inl = None, None, [], None
        for ln in out.splitlines():
            m = re.match(r"\s*\.text\.(\S+):", ln)
            if m:
                if res and cur_fn and kernel_sub in cur_fn:
                    found[cur_fn] = res
                cur_fn, res, cur_line = m.group(1), [], None
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
            if m:
                cur_line = (os.path.basename(m.group(1)), int(m.group(2)))
                # "inlined at" chains: keep the outermost location inside tc_path/our files if present
                m2 = re.findall(r'inlined at "([^"]+)", line (\d+)', m.group(3))
                inl = [(os.path.basename(a), int(b)) for a, b in m2]
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
            if m and cur_fn:
                res.append((int(m.group(1), 16), m.group(2).strip(), cur_line, inl))
        if res and cur_fn and kernel_sub in cur_fn:
            found[cur_fn] = res
    if not found:
        raise SystemExit(f"kernel {kernel_sub} not found in {LIB}")
    if want_len is None:
        return next(iter(found.values()))
    name = min(found, key=lambda k: abs(len(found[k]) - want_len))
    print(f"# library function: {name} ({len(found[name])} instructions)")
    return found[name]


def main():
    rep, ksub = sys.argv[1], sys.argv[2]
    idx = sys.argv[3] if len(sys.argv) > 3 else "1"
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", f":::{idx}"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    his = [i for i, r in enumerate(rows) if r and r[0].strip() == "Address"]
    if not his:
        raise SystemExit("no source table in the report output:\n" + out[:400])
    hi = his[0]
    hdr = rows[hi]
    print("#", rows[0][1] if len(rows[0]) > 1 else "")
    ci = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    data = rows[hi + 1:]
    sass = sass_lines(ksub, len(data))
    if len(sass) != len(data):
        print(f"# warning: {len(data)} instructions in the report vs {len(sass)} in the library (rebuilt since?)")
    agg = collections.defaultdict(lambda: [0, 0, collections.Counter(), 0])
    tot_inst = tot_samp = 0
    for k, r in enumerate(data):
        if k >= len(sass):
            break
        loc = sass[k][2] or ("?", 0)
        if sass[k][3]:
            loc = sass[k][3][-1]          # attribute inlined code to its outermost call site
        inst = int(r[ci["Instructions Executed"]] or 0)
        samp = int(r[ci["# Samples"]] or 0)
        a = agg[loc]
        a[0] += inst; a[1] += samp; a[3] += 1
        for s in stall_cols:
            v = int(r[ci[s]] or 0)
            if v:
                a[2][s] += v
        tot_inst += inst; tot_samp += samp
    print(f"# total warp-instructions {tot_inst}, samples {tot_samp}")
    src_cache = {}
    def src(loc):
        f, l = loc
        for d in ("cmd_gen_b200/csrc", "include"):
            p = os.path.join(ROOT, d, f)
            if os.path.exists(p):
                if p not in src_cache:
                    src_cache[p] = open(p).read().splitlines()
                if 0 < l <= len(src_cache[p]):
                    return src_cache[p][l - 1].strip()[:90]
        return ""
    for loc, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        st = ", ".join(f"{k[6:]}={v}" for k, v in a[2].most_common(3))
        print(f"{loc[0]}:{loc[1]:<4d} inst={a[0]:9d} ({100.0 * a[0] / max(tot_inst, 1):5.1f}%) samples={a[1]:6d} "
              f"({100.0 * a[1] / max(tot_samp, 1):5.1f}%) sass={a[3]:4d} [{st}] | {src(loc)}")


if __name__ == "__main__":
    main()
