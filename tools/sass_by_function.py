#!/usr/bin/env python
"""Attribute the executed instructions / stall samples of an ncu SASS page to source FUNCTIONS.
   nvdisasm -g -c <cubin> section of one kernel (text file) + `ncu --page source --csv` of the same kernel."""
import csv
import re
import sys


def ranges(path):
    """[(first_line, last_line, name)] of the functions / kernels of a source file (brace matching; a function is a
    block opened at namespace depth whose header holds a '(')."""
    out, depth, base, header, hstart, open_fn = [], 0, 0, "", 0, None
    for i, ln in enumerate(open(path), 1):
        code = ln.split("//")[0]
        if re.match(r"\s*namespace\s+\w+\s*\{", code) or re.match(r'\s*extern\s+"C"\s*\{', code):
            base += 1
            depth += 1
            header = ""
            continue
        if depth == base and open_fn is None:
            if not header.strip():
                hstart = i
            header += " " + code
            if "{" in code and "(" in header:
                hdr = re.sub(r"__launch_bounds__\s*\([^)]*\)|__align__\s*\([^)]*\)|template\s*<[^>]*>", " ", header)
                m = re.search(r"(\w+)\s*\(", hdr)
                open_fn = (hstart, m.group(1) if m else "?")
            if ";" in code and "{" not in code:
                header = ""
        depth += code.count("{") - code.count("}")
        if open_fn and depth == base:
            out.append((open_fn[0], i, open_fn[1]))
            open_fn, header = None, ""
        elif depth < base:
            base = depth
            header = ""
        elif depth == base and "}" in code:
            header = ""
    return out


def main():
    sass, prof_csv, core, engine = sys.argv[1:5]
    units = float(sys.argv[5]) if len(sys.argv) > 5 else 1.0
    fr = {"gobblet_core.cuh": ranges(core), "gobblet_engine.cu": ranges(engine)}
    ins, cur = [], None
    for ln in open(sass):
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split('/')[-1], int(m.group(2)))
            continue
        if re.match(r'\s*/\*[0-9a-f]{4,}\*/\s+', ln):
            ins.append(cur)
    rows = list(csv.reader(open(prof_csv)))
    hdr = rows[1]
    iex, ismp = hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
    prof = [(int(r[iex]), int(r[ismp])) for r in rows[2:] if r[iex].isdigit()]
    assert len(prof) == len(ins), (len(prof), len(ins))
    agg = {}
    for (f, l), (ex, sm) in zip(ins, prof):
        name = f"{f}:?"
        for a, b, n in fr.get(f, []):
            if a <= l <= b:
                name = n
        if f not in fr:
            name = f
        e = agg.setdefault(name, [0, 0])
        e[0] += ex
        e[1] += sm
    te, ts = sum(v[0] for v in agg.values()), sum(v[1] for v in agg.values())
    for n, (ex, sm) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{n:28s} instr/unit {ex / units:8.1f} ({100 * ex / te:5.1f}%)   samples {sm:7d} ({100 * sm / ts:5.1f}%)")


if __name__ == "__main__":
    main()
