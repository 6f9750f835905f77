#!/usr/bin/env python
"""profiles/sgns_traffic.json from an `ncu --set full` capture of the S3 bench kernel: DRAM bytes per launch of the dominant kernel,
stamped with a hash of the kernel sources so that bench.py reports `roofline.traffic` only while the kernel is the one that was
profiled.   python tools_dev/make_traffic_json.py gpurun_out/r02_s3_sgns.ncu-rep"""
import csv
import hashlib
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SOURCES = ['deepwalk-and-node2vec_b200/csrc/sgns_win.cuh', 'deepwalk-and-node2vec_b200/csrc/sgns_common.cuh', 'deepwalk-and-node2vec_b200/csrc/common.cuh',
           'deepwalk-and-node2vec_b200/csrc/sgns_win_g32.cu']


def kernel_source_sha():
    h = hashlib.sha256()
    for rel in SOURCES:
        h.update(open(os.path.join(ROOT, rel), 'rb').read())
    return h.hexdigest()[:16]


def main():
    rep = sys.argv[1]
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True, check=True).stdout
    rd = list(csv.reader(io.StringIO(out)))
    header, units, rows = rd[0], rd[1], rd[2:]
    idx = {h: i for i, h in enumerate(header)}
    row = next(r for r in rows if 'sgns_win_kernel' in r[idx['Kernel Name']])

    def gb(name):
        v, u = float(row[idx[name]].replace(',', '')), units[idx[name]]
        return v * {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0, 'Tbyte': 1e12}[u]
    rd_b, wr_b = gb('dram__bytes_read.sum'), gb('dram__bytes_write.sum')
    js = {'dram_bytes_per_launch': rd_b + wr_b, 'dram_bytes_read': rd_b, 'dram_bytes_write': wr_b, 'kernel': row[idx['Kernel Name']],
          'kernel_ms_under_ncu': float(row[idx['gpu__time_duration.sum']].replace(',', '')) * (1e-6 if units[idx['gpu__time_duration.sum']] == 'ns' else 1.0 if units[idx['gpu__time_duration.sum']] == 'ms' else 1e-3),
          'source': f'ncu --set full, one launch of the S3 bench step ({os.path.basename(rep)}); see profiles/r02_ncu_summary.md',
          'kernel_source_sha': kernel_source_sha()}
    json.dump(js, open(os.path.join(ROOT, 'profiles', 'sgns_traffic.json'), 'w'), indent=1)
    print(js)


if __name__ == '__main__':
    main()
