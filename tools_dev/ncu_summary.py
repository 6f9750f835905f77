#!/usr/bin/env python
"""Summarise `ncu --set full` captures (.ncu-rep) as markdown tables for profiles/ (developer tool; runs where ncu is
installed, no GPU needed).   python tools_dev/ncu_summary.py rep1.ncu-rep [rep2 ...] > profiles/<name>.md"""
import csv
import io
import subprocess
import sys

KEEP = [
    'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
    'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
    'sm__warps_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
    'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sectors_srcunit_tex_op_red.sum', 'lts__t_sectors_srcunit_tex_op_write.sum',
    'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__thread_inst_executed_per_inst_executed.ratio',
    'lts__d_atomic_input_cycles_active.avg.pct_of_peak_sustained_elapsed', 'lts__d_atomic_input_cycles_active.max.pct_of_peak_sustained_elapsed',
    'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
    'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed_op_shared_ld.sum', 'smsp__inst_executed_op_shared_st.sum',
]
STALL = 'smsp__average_warps_issue_stalled_'


def rows_of(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True, check=True).stdout
    rd = list(csv.reader(io.StringIO(out)))
    header, units = rd[0], rd[1]
    return header, units, rd[2:]


def main():
    for rep in sys.argv[1:]:
        header, units, rows = rows_of(rep)
        idx = {h: i for i, h in enumerate(header)}
        for row in rows:
            print(f"## {row[idx['Kernel Name']]}  ({rep.split('/')[-1]})\n")
            print('| metric | value |\n|---|---|')
            for k in KEEP:
                if k in idx:
                    print(f'| {k} | {row[idx[k]]} {units[idx[k]]} |')
            stalls = []
            for h, i in idx.items():
                if h.startswith(STALL) and h.endswith('_per_warp_active.pct') is False and h.endswith('.ratio'):
                    try:
                        stalls.append((float(row[i].replace(',', '')), h[len(STALL):-len('.ratio')]))
                    except ValueError:
                        pass
            if stalls:
                print('\nWarp stall breakdown (warps per issued instruction):\n\n| reason | ratio |\n|---|---|')
                for v, name in sorted(stalls, reverse=True)[:12]:
                    print(f'| {name} | {v:.3f} |')
            try:
                rd_b = float(row[idx['dram__bytes_read.sum']].replace(',', '')); wr_b = float(row[idx['dram__bytes_write.sum']].replace(',', ''))
                print(f"\nDRAM traffic: read {rd_b:.4g} {units[idx['dram__bytes_read.sum']]} + write {wr_b:.4g} {units[idx['dram__bytes_write.sum']]}\n")
            except (KeyError, ValueError):
                pass


if __name__ == '__main__':
    main()
