"""Turns the raw ncu artefacts a gpurun call brings back into the small, tracked summaries under profiles/.

    python profiles/summarize.py launches gpurun_out/launches.csv profiles/rN_launches.json
    python profiles/summarize.py full gpurun_out/x.ncu-rep profiles/rN_x_ncu.json

`launches`: the `--metrics gpu__time_duration.sum` CSV -> per-kernel launch count, total and share of the
profiled time (cold-cache, serialised: only the SHARES are comparable with bench.py's CUDA-event numbers).
`full`: selected raw metrics of every launch captured with `--set full`.
"""
import csv
import json
import re
import subprocess
import sys
from collections import OrderedDict

KEEP = [
    'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
    'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
    'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
    'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__waves_per_multiprocessor',
    'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
    'smsp__sass_inst_executed_op_local_ld.sum', 'smsp__sass_inst_executed_op_local_st.sum',
    # memory-hierarchy detail (round 2): L1 hit rate, L1 / L2 sector counts, L2 -> SM bytes, stall reasons per issue
    'l1tex__t_sector_hit_rate.pct', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
    'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'lts__t_sectors.sum', 'l1tex__m_xbar2l1tex_read_bytes.sum',
    'smsp__thread_inst_executed_per_inst_executed.ratio',
    'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
]


def short(name: str) -> str:
    name = re.sub(r'\(anonymous namespace\)::|<unnamed>::', '', name)
    name = re.sub(r'^void ', '', name)
    return re.sub(r'\(.*$', '', name)[:96]


def launches(src, dst):
    rows = [r for r in csv.reader(open(src, errors='replace')) if len(r) > 14 and r[0].isdigit()]
    agg = OrderedDict()
    for r in rows:
        d = agg.setdefault(short(r[4]), [0, 0.0])
        d[0] += 1
        d[1] += float(r[14])
    tot = sum(v[1] for v in agg.values())
    out = {'source': src, 'launches': len(rows), 'total_us': tot / 1e3,
           'kernels': [{'kernel': k, 'launches': v[0], 'us': v[1] / 1e3, 'share': v[1] / tot}
                       for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])]}
    json.dump(out, open(dst, 'w'), indent=1)
    for k in out['kernels'][:12]:
        print(f"{k['share']:6.1%} {k['us']:10.1f} us {k['launches']:5d}  {k['kernel']}")


def full(src, dst):
    """src: an .ncu-rep, or the `ncu -i x.ncu-rep --page raw --csv` text of one made on the GPU box (reports above the
    64 MiB gpurun_out/ limit are converted there and only the CSV travels back)."""
    if src.endswith('.csv'):
        txt = open(src, errors='replace').read()
    else:
        txt = subprocess.run(['ncu', '-i', src, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = {'kernel': short(r[hdr.index('Kernel Name')]), 'grid': r[hdr.index('Grid Size')],
             'block': r[hdr.index('Block Size')]}
        for k in KEEP:
            if k in hdr:
                i = hdr.index(k)
                d[k] = f'{r[i]} {units[i]}'.strip()
        out.append(d)
    json.dump({'source': src, 'launches': out}, open(dst, 'w'), indent=1)
    print(f'{len(out)} launches -> {dst}')


if __name__ == '__main__':
    {'launches': launches, 'full': full}[sys.argv[1]](sys.argv[2], sys.argv[3])
