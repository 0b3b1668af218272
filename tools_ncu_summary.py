#!/usr/bin/env python
"""Summarise an .ncu-rep (read with `ncu -i ... --page raw --csv`) into a small JSON for profiles/.
usage: python tools_ncu_summary.py gpurun_out/x.ncu-rep profiles/rNN_x.json "how it was captured" """
import csv
import io
import json
import subprocess
import sys

KEEP = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum',
        'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic', 'launch__shared_mem_per_block_static',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'lts__t_bytes.sum',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__inst_executed.sum', 'l1tex__throughput.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed']


def main():
    rep, out, how = sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else ""
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    keep = KEEP + [h for h in hdr if 'issue_stalled' in h and h.endswith('per_issue_active.ratio')]
    launches = []
    for r in rows[2:]:
        d = {}
        for k in keep:
            if k in hdr:
                v = r[hdr.index(k)]
                if 'issue_stalled' in k:
                    try:
                        if float(v) < 0.05:
                            continue
                    except ValueError:
                        pass
                d[k] = (v + ' ' + units[hdr.index(k)]).strip()
        launches.append(d)
    json.dump({"source": how, "report": rep, "launches": launches}, open(out, "w"), indent=1)
    for d in launches:
        print(json.dumps(d)[:2000])


if __name__ == "__main__":
    main()
