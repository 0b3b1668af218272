#!/usr/bin/env python
"""Per-phase breakdown of a barrier-structured kernel from an .ncu-rep captured with --import-source on:
the SASS listing is cut at every BAR.SYNC and the per-instruction counters are summed per segment.
usage: python tools_ncu_phases.py x.ncu-rep [out.json]"""
import csv
import io
import json
import subprocess
import sys


def main():
    rep = sys.argv[1]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    col = {h: i for i, h in enumerate(hdr)}
    segs, cur = [], {"first": None, "n": 0, "inst": 0, "samples": 0, "wave": 0, "excess": 0, "ops": {}}
    for r in rows[hdr_i + 1:]:
        if len(r) < len(hdr):
            continue
        src = r[col["Source"]].strip()
        op = src.split()[0] if not src.startswith("@") else src.split()[1]
        op = op.split(".")[0]
        inst = int(r[col["Instructions Executed"]] or 0)
        cur["n"] += 1
        cur["inst"] += inst
        cur["samples"] += int(r[col["# Samples"]] or 0)
        cur["wave"] += int(r[col["L1 Wavefronts Shared"]] or 0)
        cur["excess"] += int(r[col["L1 Wavefronts Shared Excessive"]] or 0)
        cur["ops"][op] = cur["ops"].get(op, 0) + inst
        if cur["first"] is None:
            cur["first"] = r[col["Address"]]
        if src.startswith("BAR.SYNC"):
            segs.append(cur)
            cur = {"first": None, "n": 0, "inst": 0, "samples": 0, "wave": 0, "excess": 0, "ops": {}}
    segs.append(cur)
    tot_i = sum(s["inst"] for s in segs) or 1
    tot_s = sum(s["samples"] for s in segs) or 1
    out = []
    for k, s in enumerate(segs):
        top = sorted(s["ops"].items(), key=lambda kv: -kv[1])[:6]
        rec = {"segment": k, "sass_lines": s["n"], "warp_inst": s["inst"], "inst_share": round(s["inst"] / tot_i, 4),
               "sample_share": round(s["samples"] / tot_s, 4), "smem_wavefronts": s["wave"],
               "smem_excess_wavefronts": s["excess"], "top_ops": {k2: v for k2, v in top}}
        out.append(rec)
        print(json.dumps(rec))
    if len(sys.argv) > 2:
        json.dump({"report": rep, "segments": out}, open(sys.argv[2], "w"), indent=1)


if __name__ == "__main__":
    main()
