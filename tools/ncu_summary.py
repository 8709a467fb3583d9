#!/usr/bin/env python
"""
Prints the handful of ncu metrics this project reads, from one or more .ncu-rep files (runs `ncu -i ... --page raw --csv`).

    python tools/ncu_summary.py a.ncu-rep [b.ncu-rep ...]                      text, one block per kernel launch
    python tools/ncu_summary.py --json OUT.json [--frames-per-launch F] [--pairs-per-launch P] a.ncu-rep[@F] ...

--json writes the machine-readable summary that bench.py reads for `roofline.traffic`, `issue_active` and
`slots_per_pair` (profiles/ncu_r02.json): per kernel the DRAM bytes, issue-slot utilisation, executed warp instructions
and duration of ONE launch, the hash of the CUDA sources the capture was taken from, and the capture date.
"""
import csv
import datetime
import hashlib
import json
import os
import re
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__waves_per_multiprocessor', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__inst_executed.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed_op_shared_atom.sum', 'sm__cycles_elapsed.max',
        'lts__t_sector_hit_rate.pct', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_requests_pipe_lsu_mem_global_op_red.sum', 'lts__t_sectors_op_red.sum']
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCALE = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12, 'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3,
         'usecond': 1e-3, 'msecond': 1.0, 'nsecond': 1e-6, 'second': 1e3}


def csrc_sha():
    h = hashlib.sha1()
    d = os.path.join(ROOT, "mdcraft_b200", "csrc")
    for name in sorted(os.listdir(d)):
        if name.endswith((".cu", ".cuh")):
            with open(os.path.join(d, name), "rb") as fh:
                h.update(name.encode() + b"\0" + fh.read())
    return h.hexdigest()[:16]


def rows_of(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    return hdr, units, [dict(zip(hdr, vals)) for vals in rows[2:]]


def num(d, units, key):
    v = d.get(key, '')
    if v == '':
        return None
    return float(v.replace(',', '')) * SCALE.get(units.get(key, ''), 1.0)


def short_name(full):
    m = re.search(r'(\w+)(<[^(]*>)?\(', full)
    return (m.group(1) + (m.group(2) or '')) if m else full


def main(argv):
    as_json, frames, pairs = None, 1.0, None
    files = []
    it = iter(argv)
    for a in it:
        if a == '--json':
            as_json = next(it)
        elif a == '--frames-per-launch':
            frames = float(next(it))
        elif a == '--pairs-per-launch':
            pairs = float(next(it))
        else:
            files.append(a)
    kernels = []
    default_frames = frames
    for path in files:
        frames = default_frames
        if '@' in path:                      # file.ncu-rep@F: F frames per launch in that capture
            path, f = path.rsplit('@', 1)
            frames = float(f)
        hdr, units_row, launches = rows_of(path)
        units = dict(zip(hdr, units_row))
        for d in launches:
            if not as_json:
                print('==', d.get('Kernel Name'))
                for h in hdr:
                    if h in KEYS:
                        print(f'  {h:82s} {units[h]:16s} {d[h]}')
                stalls = sorted(((float(d[h]), h) for h in hdr if 'issue_stalled' in h and h.endswith('per_issue_active.ratio') and d[h]),
                                reverse=True)[:7]
                for v, h in stalls:
                    print(f'  stall {h.split("issue_stalled_")[1].split("_per_issue")[0]:30s} {v:.3f}')
                continue
            name = short_name(d.get('Kernel Name', ''))
            rd, wr = num(d, units, 'dram__bytes_read.sum'), num(d, units, 'dram__bytes_write.sum')
            inst = num(d, units, 'smsp__inst_executed.sum')
            k = {
                "name": name, "file": os.path.basename(path),
                "duration_ms": num(d, units, 'gpu__time_duration.sum'),
                "dram_bytes": None if rd is None or wr is None else rd + wr,
                "dram_bytes_per_frame": None if rd is None or wr is None else (rd + wr) / frames,
                "frames_per_launch": frames,
                "issue_active_pct": num(d, units, 'smsp__issue_active.avg.pct_of_peak_sustained_active'),
                "warp_instructions": inst,
                "lts_throughput_pct": num(d, units, 'lts__throughput.avg.pct_of_peak_sustained_elapsed'),
                "dram_throughput_pct": num(d, units, 'dram__throughput.avg.pct_of_peak_sustained_elapsed'),
                "shared_wavefronts_pct": num(d, units, 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed'),
                "registers_per_thread": num(d, units, 'launch__registers_per_thread'),
                "warps_active_pct": num(d, units, 'sm__warps_active.avg.pct_of_peak_sustained_active'),
            }
            if pairs and 'rdf_pairs' in name and inst:
                k["pairs_per_launch"] = pairs
                k["slots_per_algorithmic_pair"] = inst * 32.0 / pairs
            kernels.append(k)
    if as_json:
        out = {"captured": datetime.date.today().isoformat(), "csrc_sha": csrc_sha(), "tool": "ncu --set full --clock-control none",
               "kernels": kernels}
        with open(as_json, 'w') as fh:
            json.dump(out, fh, indent=1)
        print(f"wrote {as_json}: {len(kernels)} kernel launches")


if __name__ == '__main__':
    main(sys.argv[1:])
