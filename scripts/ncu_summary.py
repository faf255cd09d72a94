#!/usr/bin/env python
"""Summarise an ncu report (--page raw --csv) into the handful of numbers DESIGN.md cites.
usage: python scripts/ncu_summary.py gpurun_out/prof_X.ncu-rep [out.txt]"""
import csv
import io
import subprocess
import sys

KEYS = ['Kernel Name', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'dram__bytes_read.sum',
        'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sector_hit_rate.pct',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_sector_hit_rate.pct',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__data_pipe_lsu_wavefronts.sum',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__thread_inst_executed_per_inst_executed.ratio']


def traffic(rep, config, out_json):
    """--traffic REP CONFIG OUT: per-launch DRAM bytes of the doc / term / log-likelihood pass."""
    import json
    import os
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], stdout=subprocess.PIPE,
                         stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]

    def val(r, key):
        i = hdr.index(key)
        v = float(r[i])
        return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[units[i]]
    names = {'0': 'doc_pass', '3': 'doc_pass', '1': 'term_pass', '2': 'loglik_pass'}
    data = json.load(open(out_json)) if os.path.exists(out_json) else {}
    entry = data.setdefault(config, {})
    for r in rows[2:]:
        kn = r[hdr.index('Kernel Name')]
        if 'row_pass_kernel' not in kn:
            continue
        mode = kn.split('<')[1].split(',')[2].strip()
        entry[names.get(mode, mode)] = {
            'kernel': kn.strip(), 'dram_bytes_read': val(r, 'dram__bytes_read.sum'),
            'dram_bytes_write': val(r, 'dram__bytes_write.sum'),
            'duration_us_under_ncu': float(r[hdr.index('gpu__time_duration.sum')]),
            'source': os.path.basename(rep)}
        if 'lts__t_sectors_srcunit_tex_op_read.sum' in hdr:   # 32-byte sectors L2 -> SMs
            entry[names.get(mode, mode)]['lts_sectors_read_from_sm'] = float(
                r[hdr.index('lts__t_sectors_srcunit_tex_op_read.sum')])
        if 'l1tex__t_sector_hit_rate.pct' in hdr:
            entry[names.get(mode, mode)]['l1_sector_hit_rate_pct'] = float(
                r[hdr.index('l1tex__t_sector_hit_rate.pct')])
    json.dump(data, open(out_json, 'w'), indent=1, sort_keys=True)
    print(json.dumps(entry, indent=1))


def main():
    if sys.argv[1] == '--traffic':
        return traffic(sys.argv[2], sys.argv[3], sys.argv[4])
    rep = sys.argv[1]
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], stdout=subprocess.PIPE,
                         stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                out.append(f"{k:82s} {r[i]:>18s} {units[i]}")
        out.append("")
    text = "\n".join(out)
    if len(sys.argv) > 2:
        with open(sys.argv[2], "a") as f:
            f.write(text + "\n")
    print(text)


if __name__ == "__main__":
    main()
