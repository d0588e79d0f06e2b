#!/usr/bin/env python
"""
usage: scratch/ncu_summary.py <raw.csv from `ncu -i X.ncu-rep --page raw --csv`> [out.json]
Key metrics of every kernel in an `ncu --set full` capture: duration, DRAM bytes, issue utilisation, occupancy, stall
reasons per issued instruction, pipe utilisation.  The committed profiles/r02_*.json files are this script's output.
"""
import csv
import json
import sys

KEYS = {
    'gpu__time_duration.sum': 'duration',
    'dram__bytes_read.sum': 'dram_read',
    'dram__bytes_write.sum': 'dram_write',
    'lts__t_sector_hit_rate.pct': 'l2_hit_pct',
    'smsp__inst_executed.sum': 'warp_instructions',
    'smsp__issue_active.avg.pct_of_peak_sustained_active': 'issue_active_pct',
    'sm__warps_active.avg.pct_of_peak_sustained_active': 'warps_active_pct',
    'launch__registers_per_thread': 'registers',
    'launch__grid_size': 'grid',
    'launch__block_size': 'block',
    'launch__occupancy_limit_registers': 'ctas_per_sm_by_registers',
    'launch__occupancy_limit_shared_mem': 'ctas_per_sm_by_smem',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed': 'dram_throughput_pct_of_ncu_peak',
    'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active': 'pipe_fp64_pct',
    'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active': 'pipe_xu_pct',
    'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active': 'pipe_lsu_pct',
    'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active': 'pipe_alu_pct',
    'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active': 'pipe_fma_pct',
    'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum': 'smem_bank_conflicts',
}
STALL = 'smsp__average_warps_issue_stalled_'


def num(v):
    try:
        return float(v.replace(',', ''))
    except ValueError:
        return v


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        k = {'kernel': d.get('Kernel Name', '')[:160]}
        for key, name in KEYS.items():
            if key in d and d[key] != '':
                k[name] = num(d[key])
                if name in ('duration', 'dram_read', 'dram_write'):
                    k[name + '_unit'] = u[key]
        stalls = {key[len(STALL):-len('_per_issue_active.ratio')]: round(num(v), 3) for key, v in d.items()
                  if key.startswith(STALL) and key.endswith('_per_issue_active.ratio') and v not in ('', '0')}
        k['stalls_per_issue'] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:8])
        out.append(k)
    text = json.dumps(out, indent=1)
    if len(sys.argv) > 2:
        open(sys.argv[2], 'w').write(text + '\n')
    for k in out:
        print(k['kernel'][:70], k.get('duration'), k.get('duration_unit'), 'rd', k.get('dram_read'), k.get('dram_read_unit'),
              'wr', k.get('dram_write'), k.get('dram_write_unit'), 'issue', k.get('issue_active_pct'), 'regs', k.get('registers'))


if __name__ == '__main__':
    main()
