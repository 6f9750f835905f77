"""Host-side logic of bench.py (no GPU): algorithmic byte counts, configuration strings, the reference-arm line."""
import json
import subprocess
import sys
import os

import bench

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _args(argv):
    old = sys.argv
    sys.argv = ['bench.py'] + argv
    try:
        return bench.parse_args()
    finally:
        sys.argv = old


def test_algorithmic_bytes_per_pair():
    # 2 * 4E * (1 + K + 1/N) = 6246.4 B at E=128, K=5, N=10; window kernel: 2 * 4E * (K + 2/N) = 5324.8 B
    assert abs(bench.bytes_per_pair(128, 5, 5) - 6246.4) < 1e-9
    assert abs(bench.bytes_per_pair(128, 5, 5, window=True) - 5324.8) < 1e-9
    assert abs(bench.bytes_per_pair(128, 5, 2) - 6400.0) < 1e-9
    assert abs(bench.bytes_per_pair(48, 3, 5) - 1574.4) < 1e-9


def test_workload_config_names_the_parallelism():
    a = _args(['--gpus', '8'])
    cfg = bench.workload_config(a, 8)
    assert cfg['nodes'] == 10_000_000 and cfg['edges'] == 250_000_000 and cfg['walk_len'] == 80 and cfg['emb'] == 128
    # the N > 1 headline keeps the reference's negative distribution (uniform over the WHOLE table), like N = 1: owner-computes on striped tables
    assert 'row-striped' in cfg['parallelism'] and 'whose W_out row it owns' in cfg['parallelism'] and 'se_sgns_update_pairs_owned' in cfg['parallelism']
    assert cfg['negative_sampling'] == 'uniform (reference)' == bench.workload_config(_args([]), 1)['negative_sampling']
    a = _args(['--gpus', '8', '--negatives', 'local'])
    cfg = bench.workload_config(a, 8)
    assert 'rows each GPU owns' in cfg['parallelism'] and 'owned by the GPU' in cfg['negative_sampling']
    a = _args(['--gpus', '8', '--multi', 'synced'])
    assert 'reduce-scatter + all-gather' in bench.workload_config(a, 8)['parallelism']
    a = _args(['--gpus', '2', '--negatives', 'global'])
    assert 'uniform (reference)' == bench.workload_config(a, 2)['negative_sampling']
    a = _args(['--gpus', '2', '--multi', 'a2a'])
    assert 'NCCL BASELINE' in bench.workload_config(a, 2)['parallelism']
    assert bench.workload_config(_args([]), 1)['parallelism'] == 'single GPU'


def test_reference_arm_prints_one_json_line_on_a_tiny_sample():
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '1', '--cpu-nodes', '2000',
                          '--cpu-walks-per-step', '8', '--walk-len', '12', '--emb', '16'], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line['impl'] == 'reference' and line['metric'] == 'sgns_pairs_per_s' and line['value'] > 0
    from oracle import ref_import
    want = 'reference' if ref_import.reference_root() else 'port'          # the unmodified reference wherever it can be imported
    assert line['cpu_baseline']['kind'] == want and line['e2e']['h2d_bytes_per_step'] == 0 and line['gpu_launches'] == 0
    assert ('UNMODIFIED reference' in line['cpu_baseline']['sample']) == (want == 'reference')


def test_reference_arm_can_still_time_the_port():
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--cpu-port', '--steps', '1', '--warmup', '1',
                          '--cpu-nodes', '2000', '--cpu-walks-per-step', '8', '--walk-len', '12', '--emb', '16'], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line['cpu_baseline']['kind'] == 'port' and line['value'] > 0


def test_recorded_dram_traffic_belongs_to_the_current_window_kernel_sources():
    # profiles/sgns_traffic.json is stamped with a hash of the window kernel's source set (tools_dev/make_traffic_json.py); bench.py reports
    # roofline.traffic only while that stamp matches.  A change to those files must come with a new `ncu --set full` capture.
    rec = bench.recorded_traffic()
    assert rec is not None and rec['dram_bytes_per_launch'], rec
    algorithmic = 183_500_800 * bench.bytes_per_pair(128, 5, 5, window=True)
    assert 0.9 < rec['dram_bytes_per_launch'] / algorithmic < 1.1            # DRAM traffic = the kernel's algorithmic bytes (no wasted re-reads)
