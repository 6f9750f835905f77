#!/usr/bin/env python
"""
bench.py -- the hot path of BASELINE.json on B200: node2vec walks over a CSR graph -> skip-gram windows ->
negative sampling -> in-place SGNS update of the input and context tables.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload "S3"): synthetic power-law graph, 10 M nodes / 250 M undirected edges, node2vec p=0.5 q=2
(reference code rule), walk_len 80, dim 128, r=5, K=5 uniform negatives (the reference's distribution).  One STEP =
one batch of `--walks-per-step` walks (start nodes follow the reference's schedule: shuffled node list, 10
consecutive walks per node) pushed through the walk kernel and the fused window/negatives/SGNS kernel.
value = positive (centre, context) pairs per second over the whole step (walk time included), all GPUs.

N > 1: one process per GPU, weak scaling (every GPU processes `--walks-per-step` walks per step).  Walks shard by walk id
against a replicated CSR with no communication.  SGNS (`--multi sharded`, default): ONE pair of tables, rows striped over
the HBM of the N GPUs and mapped into every process (shallow_encoders/word2vec/sharded.py); the same fused kernel gathers
rows and scatters red.global.add.v4.f32 updates straight into peer HBM over NVLink -- no staging, no separate all-to-all,
no inter-GPU barrier inside a step.  `--negatives local` (default when sharded) draws the K negatives of a pair among the
rows the GPU owns; `--negatives global` keeps the reference's uniform draw over the whole table (5/6 of the rows then
cross NVLink) and is reported beside it.  `--multi replicas`: per-GPU replicas averaged by an NCCL all-reduce each step.

--impl reference: the UNMODIFIED reference's CPU path (oracle/ref_pipeline.py imports it from /root/reference or from
the git-ignored byte-for-byte copy baseline/_ref/: Node2Vec.walk on all host cores -> tokenize -> collate ->
Word2VecTrainer.training_step -> backward -> Adam) on a bounded sample of the same workload; rank 0 only.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, 'deepwalk-and-node2vec_b200'))
sys.path.insert(0, ROOT)

METRIC = 'sgns_pairs_per_s'
UNIT = 'pairs/s'


_REAL_STDOUT = None


def quiet_stdout():
    """Route everything libraries print on fd 1 (e.g. NCCL's version banner) to stderr: stdout carries exactly ONE JSON line."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + '\n').encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_REAL_STDOUT, data)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='s3', choices=['s3', 's4'],
                    help='s3: power-law graph walks -> SGNS (the headline, BASELINE configs[3]); s4: Zipf token stream of the wiki-103 shape '
                         '-> SGNS with unigram^0.75 alias negatives (BASELINE configs[4]; single GPU)')
    ap.add_argument('--s4-vocab', type=int, default=267_735)
    ap.add_argument('--s4-sentences-per-step', type=int, default=65_536)
    ap.add_argument('--s4-sentence-len', type=int, default=128)
    ap.add_argument('--s4-zipf', type=float, default=1.0)
    ap.add_argument('--s4-power', type=float, default=0.75, help='negative-sampling exponent (0 = the reference\'s uniform draw)')
    ap.add_argument('--window-refresh', type=int, default=None, help='S4: SE_SGNS_WINDOW_REFRESH (default on for the token stream)')
    ap.add_argument('--s4-force-alias', action='store_true', help='dev: with --s4-power 0, still draw through an alias table (isolates its cost)')
    ap.add_argument('--nodes', type=int, default=10_000_000)
    ap.add_argument('--edges', type=int, default=250_000_000)
    ap.add_argument('--walks-per-step', type=int, default=262_144)
    ap.add_argument('--walks-per-node', type=int, default=10)
    ap.add_argument('--walk-len', type=int, default=80)
    ap.add_argument('--emb', type=int, default=128)
    ap.add_argument('--radius', type=int, default=5)
    ap.add_argument('--neg', type=int, default=5)
    ap.add_argument('--p', type=float, default=0.5)
    ap.add_argument('--q', type=float, default=2.0)
    ap.add_argument('--lr', type=float, default=None, help='per-pair SGD step (default 0.025 on S3; 0.0025 on the Zipf-hot S4 stream)')
    ap.add_argument('--scatter', default='red', choices=['red', 'store'])
    ap.add_argument('--seed', type=int, default=0)
    ap.add_argument('--kernel', default='window', choices=['window', 'context'],
                    help='window: context rows resident in shared memory while in the window; context: gathered per pair')
    ap.add_argument('--multi', default='sharded', choices=['sharded', 'synced', 'hybrid', 'replicas', 'a2a'],
                    help='N > 1: synced = the reference\'s global negative draw on per-GPU working copies + row-sharded masters, one fused '
                         'reduce-scatter/all-gather kernel over peer memory per step (product, reference-exact draw); sharded = one striped table '
                         'pair gathered / red.added per pair over NVLink (capacity mode; --negatives local|global|owner); a2a = the NCCL all-to-all '
                         'baseline; replicas = NCCL all-reduce averaging; hybrid = W_out ONE striped table updated by the GPU that owns each negative row '
                         '(owner-computes, reference draw, no staleness), W_in a working copy per GPU merged by the sync kernel every step')
    ap.add_argument('--merge', default=None, help='--multi synced: how the GPUs\' updates of a step combine in the sync kernel: mean (local SGD with '
                    'model averaging), sum (synchronous SGD with summed updates), stable (weight min(1, 2/G): cannot overshoot) or a weight in (0, 1]; '
                    'default: stable for --multi synced, sum for --multi hybrid (only W_in is merged there)')
    ap.add_argument('--a2a-micro-walks', type=int, default=8192, help='a2a baseline: walks per exchange micro-batch')
    ap.add_argument('--negatives', default='auto', choices=['auto', 'local', 'global', 'owner'],
                    help='sharded tables: owner (auto) = the reference\'s draw over the whole table, every GPU processes the negatives whose rows it owns for all '
                         'GPUs\' centres; local = draw negatives among the rows the GPU owns; global = reference draw, rows fetched over NVLink per pair')
    ap.add_argument('--tables', default='torch', choices=['torch', 'vmm'], help='N = 1: torch tensor or a 1-shard VMM table')
    ap.add_argument('--extra-steps', type=int, default=5, help='sharded: steps of the other negative mode timed after the main run')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--cpu-nodes', type=int, default=100_000, help='node count of the CPU sample graph (same mean degree)')
    ap.add_argument('--cpu-walks-per-step', type=int, default=64, help='reference batch_size (walks per step)')
    ap.add_argument('--cpu-steps', type=int, default=8, help='timed steps of the cpu_baseline sample (about 10-30 s of CPU work)')
    ap.add_argument('--cpu-port', action='store_true', help='time oracle/cpu_port.py (the restated pattern) instead of the unmodified reference')
    return ap.parse_args()


def parallelism(a, n_gpus):
    if n_gpus == 1:
        return 'single GPU' + (' (tables in a 1-shard VMM mapping)' if a.tables == 'vmm' else '')
    if a.multi == 'replicas':
        return f'dp{n_gpus}: walks sharded by id, table replicas averaged by NCCL all-reduce every step'
    if a.multi == 'hybrid':
        return (f'dp{n_gpus}: walks sharded by id (replicated CSR) and all-gathered (4 B per token); W_out is ONE table row-striped over {n_gpus} HBMs and every '
                f'negative pair (reference draw over the whole table) is computed by the GPU that owns the negative row, against its own HBM; W_in is a working '
                f'copy per GPU + row-sharded masters merged every step by ONE kernel over NVLink peer memory (csrc/replica.cu); positives on the walk\'s home GPU')
    if a.multi == 'synced':
        return (f'dp{n_gpus}: walks sharded by id (replicated CSR, no communication); both tables row-sharded into {n_gpus} master chunks (rows by node id) '
                f'+ one working copy per GPU; every GPU runs the single-GPU fused kernel with the reference\'s uniform draw over the WHOLE table, then '
                f'ONE kernel per table sums the updates of all copies into the masters and writes the rows back over NVLink peer memory '
                f'(fused reduce-scatter + all-gather, csrc/replica.cu); no NCCL on the data path')
    neg = ('local' if a.multi == 'a2a' else 'owner') if a.negatives == 'auto' else a.negatives
    if a.multi == 'a2a':
        return (f'dp{n_gpus} NCCL BASELINE: tables row-sharded by row % {n_gpus}; per micro-batch of {a.a2a_micro_walks} walks: unique ids -> '
                f'all_to_all ids / rows -> se_sgns_grad on compact tables -> all_to_all gradients -> owners apply; negatives '
                + ('among the rows each GPU owns' if neg == 'local' else 'uniform over the whole table (reference)'))
    if neg == 'owner':
        return (f'dp{n_gpus}: walks sharded by id (replicated CSR) and all-gathered (4 B per token, the only collective); ONE pair of tables row-striped '
                f'(2 MiB stripes) over {n_gpus} HBMs; negatives drawn uniformly over the whole table (reference); every GPU buckets the centres of the '
                'gathered batch by table row and computes EVERY pair -- positive or negative -- whose W_out row it owns against its own HBM '
                '(se_sgns_update_pairs_owned): W_out never crosses NVLink, a W_in centre row is read / reduced over NVLink once per run of equal rows')
    return (f'dp{n_gpus}: walks sharded by id (replicated CSR, no communication); ONE pair of tables row-striped (2 MiB stripes) over '
            f'{n_gpus} HBMs, fused kernel gathers / red.adds peer rows over NVLink; negatives drawn '
            + {'local': 'among the rows each GPU owns', 'global': 'uniformly over the whole table (reference)',
               'owner': 'uniformly over the whole table (reference); walks all-gathered (4 B per token), centres bucketed by table row on every GPU, '
                        'and EVERY pair -- positive or negative -- computed by the GPU that owns its W_out row against its own HBM '
                        '(se_sgns_update_pairs_owned): W_out never crosses NVLink, a W_in centre row crosses once per run of equal rows'}[neg])


def workload_config(a, n_gpus):
    return {
        'workload': 'S3 synthetic power-law graph: node2vec walks -> windows -> negatives -> SGNS update',
        'nodes': a.nodes, 'edges': a.edges, 'method': 'node2vec', 'p': a.p, 'q': a.q, 'rule': 'reference-code',
        'walk_len': a.walk_len, 'walks_per_node': a.walks_per_node, 'walks_per_step_per_gpu': a.walks_per_step,
        'emb': a.emb, 'context_radius': a.radius, 'negatives': a.neg,
        'negative_sampling': ('uniform over the rows owned by the GPU (walks are dealt to GPUs by id)'
                              if (n_gpus > 1 and ((a.multi == 'a2a' and a.negatives in ('auto', 'local')) or (a.multi == 'sharded' and a.negatives == 'local')))
                              else 'uniform (reference)'),
        'table_sync': (f'every step: master += beta * sum over GPUs of (working copy - master), written back to all copies; merge = {a.merge}'
                       if n_gpus > 1 and a.multi == 'synced' else None),
        'optimizer': 'in-place SGD (Hogwild, red.global.add.v4.f32)' if a.scatter == 'red' else 'in-place SGD (Hogwild, plain stores)',
        'parallelism': parallelism(a, n_gpus),
        'l2': 'inputs exceed L2 (tables 2 x %.2f GB, CSR ~%.1f GB); no flush' % ((a.nodes + 1) * a.emb * 4 / 1e9, (2 * a.edges * 4 + a.nodes * 8) / 1e9),
    }


def sgns_kernel_label(emb, neg, window):
    """Name of the kernel `se_sgns_update_walks` dispatches to for this shape (csrc/sgns.cu::launch, launch_win)."""
    t = 1 + neg
    if window and emb % 4 == 0 and 128 < emb <= 256 and 1 <= neg <= 7:
        return f'sgns_winw_kernel<R=2, T={t}> (E={emb}; csrc/sgns_win_wide.cu)'
    if window and emb % 4 == 0 and 16 <= emb <= 128 and neg <= 7:
        g = 32 if emb > 64 else (16 if emb > 32 else 8)
        return f'sgns_win_kernel<G={g}, T={t}, EXACT={"true" if emb == 128 else "false"}> (E={emb})'
    if emb % 4 == 0 and 32 < emb <= 128 and neg <= 7:
        return f'sgns_ctx_kernel<MODE_WALK, T={t}> (E={emb})'
    if emb % 4 == 0 and 64 < emb <= 1024:
        return f'sgns_fast_kernel<MODE_WALK, R={-(-emb // 128)}> (E={emb})'
    return f'sgns_kernel<MODE_WALK> generic (E={emb})'


def bytes_per_pair(emb, neg, radius, window=False):
    """Algorithmic HBM bytes per positive pair, fp32 rows read + written once per use.
    Per-pair kernel: the K negative rows and the context row per pair, the centre row amortised over its 2r
    contexts: 2 * 4E * (1 + K + 1/N).
    Window-resident kernel: a token's context row is fetched and scattered once per window pass (= once per centre, i.e.
    1/N per pair) instead of once per pair: 2 * 4E * (K + 2/N)."""
    n = 2 * radius
    return 2.0 * 4.0 * emb * ((neg + 2.0 / n) if window else (1.0 + neg + 1.0 / n))


# --------------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock and clock-event reasons through NVML while the timed region runs."""
    REASONS = {0x1: 'gpu_idle', 0x2: 'applications_clocks_setting', 0x4: 'sw_power_cap', 0x8: 'hw_slowdown',
               0x10: 'sync_boost', 0x20: 'sw_thermal_slowdown', 0x40: 'hw_thermal_slowdown', 0x80: 'hw_power_brake_slowdown',
               0x100: 'display_clock_setting'}

    def __init__(self, index, period=0.1):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz, self.power = [], set(), None, []
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:   # noqa: BLE001
            self.err = repr(e)

    def run(self):
        if not self.ok:
            return
        while not self._stop_evt.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                try:
                    mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:   # noqa: BLE001
                    mask = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
                self.power.append(self.nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            except Exception:   # noqa: BLE001
                pass
            self._stop_evt.wait(self.period)

    def stop(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        if not self.samples:
            return {'sm_mhz': None, 'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons), 'samples': 0}
        s = sorted(self.samples)
        return {'sm_mhz': s[len(s) // 2], 'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons - {'gpu_idle'}),
                'samples': len(s), 'power_w_max': max(self.power) if self.power else None}


def nvlink_bytes(index):
    """(tx, rx) NVLink data bytes of GPU `index` since driver start (NVML field values, all links), or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        vals = pynvml.nvmlDeviceGetFieldValues(h, [(pynvml.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX, 0xFFFFFFFF),
                                                   (pynvml.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_RX, 0xFFFFFFFF)])
        out = []
        for v in vals:
            if v.nvmlReturn != 0:
                return None
            out.append(int(v.value.ullVal) * 1024)          # the counters are in KiB
        return tuple(out)
    except Exception:   # noqa: BLE001
        return None


def nvml_index(local_rank):
    vis = os.environ.get('CUDA_VISIBLE_DEVICES')
    if vis:
        parts = [p.strip() for p in vis.split(',') if p.strip()]
        if local_rank < len(parts) and parts[local_rank].isdigit():
            return int(parts[local_rank])
    return local_rank


def measured_peak():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        try:
            return float(json.load(open(path))['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
        except Exception:   # noqa: BLE001
            pass
    return 6650.0, 'fallback (B200_PROFILING.md)'


def recorded_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture -- only while the kernel sources are the
    ones that were profiled (profiles/sgns_traffic.json carries their hash, tools_dev/make_traffic_json.py); otherwise None."""
    path = os.path.join(ROOT, 'profiles', 'sgns_traffic.json')
    try:
        js = json.load(open(path))
        import hashlib
        h = hashlib.sha256()
        for rel in ('sgns_win.cuh', 'sgns_common.cuh', 'common.cuh', 'sgns_win_g32.cu'):
            h.update(open(os.path.join(ROOT, 'deepwalk-and-node2vec_b200', 'csrc', rel), 'rb').read())
        if js.get('kernel_source_sha') != h.hexdigest()[:16]:
            return {'dram_bytes_per_launch': None, 'source': 'stale: profiles/sgns_traffic.json was captured for other kernel sources; re-run tools_dev/make_traffic_json.py'}
        return js
    except Exception:   # noqa: BLE001
        return None


# --------------------------------------------------------------------------------------------------------------
def _reference_sample(a, steps, warmup):
    """The CPU arm on a bounded sample of the S3 workload.  Preferred: the UNMODIFIED reference (oracle/ref_pipeline.py in its own
    process: /root/reference, or the byte-for-byte copy under baseline/_ref/ that travels to the GPU box), kind "reference".
    Only if the reference cannot be imported: oracle/cpu_port.py (same pattern restated; faster than the reference because it tests
    membership on sets), kind "port"."""
    cores = os.cpu_count() or 1
    nodes = min(a.cpu_nodes, a.nodes)
    edges = int(round(a.edges * (nodes / a.nodes)))
    why = None
    if not a.cpu_port:
        from oracle import ref_pipeline
        r = ref_pipeline.call(nodes=nodes, edges=edges, method='node2vec', p=a.p, q=a.q, walk_len=a.walk_len,
                              walks_per_node=a.walks_per_node, walks_per_step=a.cpu_walks_per_step, radius=a.radius, emb=a.emb,
                              neg=a.neg, optimizer='adam', lr=0.1, steps=steps, warmup=warmup, workers=cores, seed=a.seed)
        if 'unavailable' not in r:
            what = (f'{steps} steps x {a.cpu_walks_per_step} walks (reference batch_size; {r["seconds"]:.1f} s CPU) on a {r["graph"]["nodes"]}-node / '
                    f'{r["graph"]["edges"]}-edge power-law sample of S3 (same generator, same mean degree); the UNMODIFIED reference from '
                    f'{"baseline/_ref (byte-for-byte copy)" if "baseline" in r["reference_root"] else r["reference_root"]}: Node2Vec.walk in {cores} worker '
                    f'processes -> tokenize -> vocab -> W2VCollateFunctional -> Word2VecTrainer.training_step -> backward -> torch.optim.Adam '
                    f'({r["torch_threads"]} threads); walk {r["walk_steps_per_s"]:.3g} steps/s, sgns {r["sgns_pairs_per_s"]:.3g} pairs/s')
            return r, 'reference', what, cores
        why = r['unavailable']
    from oracle import cpu_port
    g = cpu_port.powerlaw_graph_host(nodes, edges, a.seed)
    r = cpu_port.run_reference_pipeline(g, a.cpu_walks_per_step, steps, warmup, a.walk_len, a.p, a.q, True, a.radius,
                                        a.emb, a.neg, cores, optimizer='adam', seed=a.seed)
    what = (f'{steps} steps x {a.cpu_walks_per_step} walks ({r["seconds"]:.1f} s CPU) on a {nodes}-node / {edges}-edge power-law sample of S3; '
            f'oracle/cpu_port.py = the reference pattern restated (python node2vec walks on {cores} processes, python collate, torch CPU '
            f'SkipGram + loss + backward + dense Adam, {cores} threads)' + (f'; reference itself unavailable: {why}' if why else ''))
    return r, 'port', what, cores


def run_reference(a, rank, world):
    """CPU arm (`--impl reference`): the reference's own CPU implementation of the path on all host cores, bounded sample of
    the same workload; rank 0 only."""
    if rank != 0:
        return
    r, kind, sample, cores = _reference_sample(a, a.steps, a.warmup)
    cfg = workload_config(a, 1)
    cfg['parallelism'] = f'{cores} host cores'
    cfg['optimizer'] = 'torch.optim.Adam on dense tables (as every shipped YAML, e.g. configs/sge_sg_cora.yaml:32-34)'
    cfg['cpu_sample'] = {'nodes': min(a.cpu_nodes, a.nodes), 'walks_per_step': a.cpu_walks_per_step}
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': r['pairs_per_s'], 'unit': UNIT, 'n_gpus': a.gpus, 'steps': a.steps,
        'warmup': a.warmup, 'ms_per_step': r['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic', 'config': cfg,
        'cpu_baseline': {'value': r['pairs_per_s'], 'unit': UNIT, 'cores': cores, 'kind': kind, 'sample': sample},
        'e2e': {'value': r['pairs_per_s'], 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'walk_steps_per_s': r['walk_steps_per_s'], 'sgns_pairs_per_s': r['sgns_pairs_per_s'], 'gpu_launches': 0,
    }
    emit(line)


def cpu_baseline(a):
    steps = max(3, int(a.cpu_steps))
    r, kind, sample, cores = _reference_sample(a, steps, 1)
    return {'value': r['pairs_per_s'], 'unit': UNIT, 'cores': cores, 'kind': kind, 'sample': sample,
            'walk_steps_per_s': r['walk_steps_per_s'], 'sgns_pairs_per_s': r['sgns_pairs_per_s']}


# --------------------------------------------------------------------------------------------------------------
def run_b200(a, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    from shallow_encoders import _native as nat
    from shallow_encoders.graph.synthetic import powerlaw_graph_device

    assert torch.cuda.is_available(), 'bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm'
    dev = torch.device('cuda', local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    nat.load()

    # ---- resident state: replicated CSR, tables ----------------------------------------------------------------
    t_build = time.perf_counter()
    csr = powerlaw_graph_device(a.nodes, a.edges, a.seed, dev)
    torch.cuda.synchronize()
    t_build = time.perf_counter() - t_build
    print(f'[bench] rank {rank}: graph built on the device in {t_build:.2f} s (se_csr_build: {csr.n_nodes} nodes, {csr.nnz} CSR entries, max degree {csr.max_degree})',
          file=sys.stderr, flush=True)
    vocab = a.nodes + 1                                   # row 0 = '<unk>' (torch_dataset.py:99-110)
    bound = (6.0 / (vocab + a.emb)) ** 0.5                # xavier_uniform_ (model.py:26-27)
    fallback_note = None
    sharded = (world > 1 and a.multi == 'sharded') or (world == 1 and a.tables == 'vmm')
    a2a = world > 1 and a.multi == 'a2a'
    synced = world > 1 and a.multi == 'synced'
    hybrid = world > 1 and a.multi == 'hybrid'
    if a.merge is None:
        a.merge = 'sum' if a.multi == 'hybrid' else 'stable'
    ex = None
    neg_mode = 'global'
    if (sharded or a2a) and world > 1:
        neg_mode = ('local' if a2a else 'owner') if a.negatives == 'auto' else a.negatives
        assert not (a2a and neg_mode == 'owner'), 'the NCCL baseline has no owner-computes mode'
    local_neg = neg_mode == 'local'
    if a2a:
        from shallow_encoders.word2vec.row_exchange import RowShardedTables
        tables = RowShardedTables(vocab, a.emb, rank, world, dev)
        tables.fill_uniform(bound, a.seed + 101, a.seed + 102)
        w_in = w_out = None
    elif sharded or synced or hybrid:
        # Peer-mapped tables need CUDA VMM handle export between processes (POSIX fds over unix sockets).  If the platform
        # refuses that on ANY rank, every rank drops to the replica mode (still the same CUDA kernels) and the line says so.
        from shallow_encoders.word2vec.sharded import ReplicatedTable, ShardedTable, make_exchange
        w_in = w_out = None
        try:
            ex = make_exchange(rank, world)
            w_in = (ReplicatedTable if (synced or hybrid) else ShardedTable)(vocab, a.emb, dev, rank, world, ex)
            w_out = (ReplicatedTable if synced else ShardedTable)(vocab, a.emb, dev, rank, world, ex)
            ok, why = 1, ''
        except Exception as e:   # noqa: BLE001
            ok, why = 0, repr(e)
        if world > 1:
            flag = torch.tensor([ok], dtype=torch.int32, device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            ok = int(flag.item())
        if not ok:
            if world == 1:
                raise RuntimeError(f'striped-table set-up failed: {why}')
            for t in (w_in, w_out):
                if t is not None:
                    t.close()
            sharded, synced, hybrid, local_neg, neg_mode = False, False, False, False, 'global'
            fallback_note = f'peer-mapped tables unavailable ({why or "on another rank"}): ran --multi replicas'
            a.multi = 'replicas'
            w_in = torch.empty((vocab, a.emb), dtype=torch.float32, device=dev)
            w_out = torch.empty((vocab, a.emb), dtype=torch.float32, device=dev)
    else:
        w_in = torch.empty((vocab, a.emb), dtype=torch.float32, device=dev)
        w_out = torch.empty((vocab, a.emb), dtype=torch.float32, device=dev)
    if synced or hybrid:
        w_in.fill_uniform(bound, a.seed + 101)                # every rank fills its working copy and adopts its master chunk
        if synced:
            w_out.fill_uniform(bound, a.seed + 102)
        else:
            nat.table_fill_uniform(w_out, bound, a.seed + 102)
            neg_mode = 'owner'
    elif not a2a:
        nat.table_fill_uniform(w_in, bound, a.seed + 101)     # same content on every rank / for every sharding
        nat.table_fill_uniform(w_out, bound, a.seed + 102)
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()                                        # nobody touches a peer's rows before they are initialised
    T = {'w_in': w_in, 'w_out': w_out, 'synced': synced, 'sharded': sharded, 'hybrid': hybrid}
    flags = nat.SCATTER_RED if a.scatter == 'red' else nat.SCATTER_STORE
    if a.kernel == 'context':
        flags |= nat.NO_WINDOW

    # ---- schedule: shuffled node list, walks_per_node consecutive walks per node (graph/datasets.py:45,76) -----
    total_steps = a.warmup + 2 * a.steps + 2 + 3 * (a.extra_steps + 1)
    n_walks = a.walks_per_step
    g_cpu = torch.Generator()
    g_cpu.manual_seed(a.seed)
    order = torch.randperm(a.nodes, generator=g_cpu, dtype=torch.int64)
    n_cen = a.walk_len - 2 * a.radius

    def starts_for(step):
        # global walk index of this rank's j-th walk in `step`: ((step * world + rank) * n_walks + j)
        base = (step * world + rank) * n_walks
        idx = (torch.arange(base, base + n_walks, dtype=torch.int64) // a.walks_per_node) % a.nodes
        return order[idx].to(torch.int32), base

    host_starts = [starts_for(s) for s in range(total_steps)]
    pinned = [(st.pin_memory(), base) for st, base in host_starts]
    dev_starts = [(st.to(dev), base) for st, base in host_starts]
    walks = torch.empty((n_walks, a.walk_len), dtype=torch.int32, device=dev)
    stats = torch.zeros(nat.STATS_LEN, dtype=torch.float64, device=dev)
    scratch = {'starts': torch.empty(n_walks, dtype=torch.int32, device=dev), 'walks': walks, 'stats': stats}
    stats_host = torch.zeros(nat.STATS_LEN, dtype=torch.float64).pin_memory()

    ev = lambda: torch.cuda.Event(enable_timing=True)   # noqa: E731
    sgns_events, walk_events, sync_events = [], [], []

    def sync_tables(record=False):
        if world == 1 or a2a or T['sharded']:
            return
        if record:
            s0, s1 = ev(), ev()
            s0.record()
        if T['synced'] or T['hybrid']:
            from shallow_encoders.word2vec.sharded import sync_replicated
            sync_replicated([T['w_in'], T['w_out']] if T['synced'] else [T['w_in']], merge=a.merge)
        else:
            dist.all_reduce(T['w_in'], op=dist.ReduceOp.AVG)
            dist.all_reduce(T['w_out'], op=dist.ReduceOp.AVG)
        if record:
            s1.record()
            sync_events.append((s0, s1))

    gather_buf = torch.empty((world * n_walks, a.walk_len), dtype=torch.int32, device=dev) if world > 1 and not a2a else None

    def sgns_stage(base, step, mode):
        if a2a:
            tables.step(walks, a.radius, a.neg, 1, a.lr, a.seed + 1, draw_id_base=base * n_cen * 2 * a.radius * a.neg,
                        micro_walks=a.a2a_micro_walks, local_negatives=mode == 'local', stats=stats)
        elif mode == 'owner':
            from shallow_encoders.word2vec.sharded import sgns_update_walks_owner_computes
            sgns_update_walks_owner_computes(T['w_in'], T['w_out'], walks, a.radius, a.neg, 1, a.lr, a.seed + 1, step * world * n_walks * n_cen,
                                             rank, world, stats=stats, gather_buf=gather_buf)
        else:
            nat.sgns_update_walks(T['w_in'], T['w_out'], walks, a.radius, a.neg, 1, a.lr, a.seed + 1, centre_id_base=base * n_cen,
                                  flags=flags, stats=stats, local_negatives=mode == 'local')

    def device_step(step, record=False, mode=None):
        mode = neg_mode if mode is None else mode
        st, base = dev_starts[step]
        if record:
            e0, e1, e2 = ev(), ev(), ev()
            e0.record()
        nat.walk(csr, st, a.walk_len, a.p, a.q, True, nat.RULE_REFERENCE, a.seed, walk_id_base=base, out=walks)
        if record:
            e1.record()
        sgns_stage(base, step, mode)
        if record:
            e2.record()
            walk_events.append((e0, e1))
            sgns_events.append((e1, e2))
        sync_tables(record)

    def host_step(step):
        st, base = pinned[step]
        if a2a or neg_mode == 'owner':     # a collective sits inside the step: the host-buffer contract assembled from the device calls
            scratch['starts'].copy_(st, non_blocking=True)
            stats.zero_()
            nat.walk(csr, scratch['starts'], a.walk_len, a.p, a.q, True, nat.RULE_REFERENCE, a.seed, walk_id_base=base, out=walks)
            sgns_stage(base, step, neg_mode)
            sync_tables()
            stats_host.copy_(stats)
            torch.cuda.synchronize()
            return
        nat.host_walk_sgns_step(csr, st, a.walk_len, a.p, a.q, True, nat.RULE_REFERENCE, a.seed, base, T['w_in'], T['w_out'], a.radius,
                                a.neg, 1, a.lr, scratch, stats_host, flags=flags, local_negatives=local_neg)
        sync_tables()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident timing (`value`) ---------------------------------------------------------------------
    for s in range(a.warmup):
        device_step(s)
    sampler = ClockSampler(nvml_index(local_rank))
    barrier()
    launches0 = nat.launches()
    sampler.start()
    nvl0 = nvlink_bytes(nvml_index(local_rank)) if world > 1 else None
    torch.cuda.profiler.start()          # `ncu --profile-from-start off` captures exactly the timed region
    t0, t1 = ev(), ev()
    t0.record()
    for s in range(a.warmup, a.warmup + a.steps):
        device_step(s, record=True)
    t1.record()
    barrier()
    torch.cuda.profiler.stop()
    nvl1 = nvlink_bytes(nvml_index(local_rank)) if world > 1 else None
    clocks = sampler.stop()
    launches = nat.launches() - launches0
    ms_total = max_over_ranks(t0.elapsed_time(t1))
    pairs_per_step = n_walks * n_cen * 2 * a.radius
    walk_steps_per_step = n_walks * (a.walk_len - 1)
    value = world * pairs_per_step * a.steps / (ms_total / 1e3)
    sgns_ms = max_over_ranks(sum(x.elapsed_time(y) for x, y in sgns_events) / len(sgns_events))
    walk_ms = max_over_ranks(sum(x.elapsed_time(y) for x, y in walk_events) / len(walk_events))
    sync_ms = max_over_ranks(sum(x.elapsed_time(y) for x, y in sync_events) / len(sync_events)) if sync_events else None
    stat_vals = stats.tolist()

    # ---- end-to-end through the host-buffer C-ABI entry (`e2e`) -------------------------------------------------
    host_step(a.warmup + a.steps)       # warm
    barrier()
    w0 = time.perf_counter()
    e0, e1 = ev(), ev()
    e0.record()
    for s in range(a.warmup + a.steps + 1, a.warmup + 2 * a.steps + 1):
        host_step(s)
    e1.record()
    barrier()
    e2e_wall = time.perf_counter() - w0
    e2e_ms = max_over_ranks(max(e0.elapsed_time(e1), 0.0))
    e2e_ms = max(e2e_ms, max_over_ranks(e2e_wall * 1e3) if world > 1 else e2e_wall * 1e3)   # host copies + sync are on the clock
    e2e_value = world * pairs_per_step * a.steps / (e2e_ms / 1e3)

    # ---- the other multi-GPU modes, a few steps each, reported beside the headline ---------------------------------
    other = None
    if (sharded or a2a or synced or hybrid) and world > 1 and a.extra_steps > 0:
        names = {'local': 'striped tables, negatives among the rows the GPU owns (GraphVite-style partitioned sampler: not the reference\'s per-pair draw; '
                          'accuracy = 1 GPU at 8 GPUs, profiles/r02_multi_gpu.md)',
                 'global': 'striped tables, reference draw, negative rows fetched / red.added over NVLink per pair',
                 'owner': 'striped tables, reference draw, owner-computes (every pair on the GPU that owns its W_out row, centres bucketed by row; accuracy = 1 GPU)',
                 'synced': 'working copy per GPU + row-sharded masters, reference draw, one fused reduce-scatter/all-gather kernel per step, merge = stable '
                           '(throughput mode: statistically inefficient at this step size on 8 GPUs, profiles/r02_multi_gpu.md)'}

        def swap_tables(kind):
            """Replace the resident table pair by a freshly initialised one of the other kind (striped <-> working copies)."""
            from shallow_encoders.word2vec.sharded import ReplicatedTable, ShardedTable
            T['w_in'].close(); T['w_out'].close()
            torch.cuda.empty_cache()
            make = ReplicatedTable if kind == 'synced' else ShardedTable
            T['w_in'], T['w_out'] = make(vocab, a.emb, dev, rank, world, ex), make(vocab, a.emb, dev, rank, world, ex)
            if kind == 'synced':
                T['w_in'].fill_uniform(bound, a.seed + 101); T['w_out'].fill_uniform(bound, a.seed + 102)
            else:
                nat.table_fill_uniform(T['w_in'], bound, a.seed + 101); nat.table_fill_uniform(T['w_out'], bound, a.seed + 102)
            T['synced'], T['hybrid'], T['sharded'] = kind == 'synced', False, kind != 'synced'
            barrier()

        main_mode = 'synced' if synced else ('hybrid' if hybrid else neg_mode)
        if a2a:
            plan = [m for m in ('local', 'global') if m != main_mode]
        else:
            plan = [m for m in ('owner', 'local', 'synced') if m != main_mode]
        other, first = [], a.warmup + 2 * a.steps + 1
        merge_main = a.merge
        for mode in plan:
            if not a2a:
                want = 'synced' if mode == 'synced' else 'striped'
                have = 'synced' if T['synced'] else ('striped' if T['sharded'] else 'other')
                if want != have:
                    swap_tables(want)
            a.merge = 'stable' if mode == 'synced' else merge_main
            step_mode = 'global' if mode == 'synced' else mode
            device_step(first, mode=step_mode)
            barrier()
            x0, x1 = ev(), ev()
            x0.record()
            for s in range(first + 1, first + 1 + a.extra_steps):
                device_step(s, mode=step_mode)
            x1.record()
            barrier()
            other_ms = max_over_ranks(x0.elapsed_time(x1))
            other.append({'mode': names[mode], 'value': world * pairs_per_step * a.extra_steps / (other_ms / 1e3), 'unit': UNIT,
                          'steps': a.extra_steps, 'ms_per_step': other_ms / a.extra_steps})
            first += a.extra_steps + 1
        a.merge = merge_main

    def close_tables():
        if T['sharded'] or T['synced'] or T.get('hybrid'):
            T['w_in'].close(); T['w_out'].close()
        if ex is not None:
            ex.close()

    if rank != 0:
        if world > 1:
            barrier()
            close_tables()
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak()
    window = a.kernel == 'window' and not a2a
    owner_pairs = world > 1 and not a2a and neg_mode == 'owner'      # the stage is se_sgns_update_pairs_owned (+ all-gather + bucketing)
    bpp = bytes_per_pair(a.emb, a.neg, a.radius, window)
    if owner_pairs:
        bpp = 2.0 * 4.0 * a.emb * (1.0 + a.neg)        # every output row of a pair is read and reduced once, in the owner's HBM
    bpp_survey = bytes_per_pair(a.emb, a.neg, a.radius, False)
    achieved = pairs_per_step * bpp / (sgns_ms / 1e3) / 1e9
    # the ncu capture is one launch of THIS configuration (E = 128, K = 5, r = 5, 262,144 walks of 80): other shapes carry no traffic figure
    traffic = recorded_traffic() if (a.emb, a.neg, a.radius, a.kernel, pairs_per_step) == (128, 5, 5, 'window', 183_500_800) else None
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': a.steps, 'warmup': a.warmup,
        'ms_per_step': ms_total / a.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
        'data': 'synthetic', 'config': workload_config(a, world),
        'walk_steps_per_s': world * walk_steps_per_step / (walk_ms / 1e3),
        'sgns_kernel_pairs_per_s': world * pairs_per_step / (sgns_ms / 1e3),
        'kernel_ms': {'walk_kernel': walk_ms, 'sgns_kernel': sgns_ms},
        'roofline': {
            'bound': 'hbm', 'kernel': 'se_sgns_grad inside the NCCL row exchange (baseline)' if a2a else (
                f'sgns_owned_pairs_kernel<EXACT={"true" if a.emb == 128 else "false"}, POS=true> via se_sgns_update_pairs_owned (stage time includes the '
                'all-gather of the walks and the centre bucketing kernels)' if owner_pairs else
                sgns_kernel_label(a.emb, a.neg, a.kernel == 'window') + ' via se_sgns_update_walks'), 'achieved': achieved, 'peak': peak,
            'unit': 'GB/s', 'frac': achieved / peak, 'peak_source': peak_src,
            'algorithmic_bytes_per_pair': bpp, 'pairs_per_launch': pairs_per_step,
            'bytes_per_pair_formula': ('2*4E*(1 + K): each GPU processes (1 + K) / G of the output rows of all G GPUs\' pairs = (1 + K) rows per pair of its '
                                       'own share; centre rows (once per run of equal rows) not counted') if owner_pairs else (
                '2*4E*(K + 2/N): context rows resident per window' if window else '2*4E*(1 + K + 1/N) (SURVEY 8d)'),
            'survey_unit': {'bytes_per_pair': bpp_survey, 'frac': pairs_per_step * bpp_survey / (sgns_ms / 1e3) / 1e9 / peak},
            'traffic': None if owner_pairs else (traffic or {}).get('dram_bytes_per_launch'),
            'traffic_source': None if owner_pairs else (traffic or {}).get('source'),
            # the window kernel fetches a token's context row once per window instead of once per pair, so its DRAM traffic
            # is below the per-pair algorithmic figure; the DRAM-side rate is traffic / launch time
            'dram_traffic_gbs': ((traffic or {}).get('dram_bytes_per_launch') or 0) / (sgns_ms / 1e3) / 1e9 if traffic and a.kernel == 'window' and world == 1 else None,
        },
        'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': n_walks * 4, 'd2h_bytes_per_step': nat.STATS_LEN * 8,
                'ms_per_step': e2e_ms / a.steps, 'api': 'se_host_walk_sgns_step (pinned host start nodes in, loss statistics out)'},
        'gpu_launches': launches,
        'clocks': clocks,
        'train_stats': {'loss': (stat_vals[0] + stat_vals[1]) / max(stat_vals[4], 1), 'pairs': stat_vals[4]},
        'graph': {'n_nodes': csr.n_nodes, 'nnz': csr.nnz, 'max_degree': csr.max_degree, 'device_build_s': t_build},
        'library': nat.version(),
    }
    if fallback_note:
        line['multi_gpu_fallback'] = fallback_note
    if other is not None:
        line['other_multi_gpu_modes'] = other
    if a2a:
        line['a2a'] = {'micro_walks': a.a2a_micro_walks, 'exchanged_bytes_sent_per_rank': tables.exchanged_bytes}
    if sharded:
        line['tables'] = {'kind': 'vmm-striped', 'stripe_bytes': w_in.stripe_bytes, 'stripes_per_table': w_in.n_stripes,
                          'bytes_per_gpu': 2 * w_in.n_stripes * w_in.stripe_bytes // world}
    if synced or hybrid:
        table_bytes = vocab * a.emb * 4
        moved = (2 if synced else 1) * 2 * table_bytes * (world - 1) / world          # per GPU per direction per step (both tables / W_in only)
        line['tables'] = {'kind': ('working copy per GPU + row-sharded masters (vmm peer-mapped)' if synced else
                                   'W_in: working copy per GPU + row-sharded masters; W_out: one table striped over the GPUs (vmm peer-mapped)'),
                          'bytes_per_gpu': (2 if synced else 1) * w_in.seg_bytes + 2 * table_bytes // world, 'merge': a.merge}
        line['table_sync'] = {'ms_per_step': sync_ms, 'nvlink_bytes_per_gpu_per_direction': moved,
                              'achieved_gbs_per_direction': (moved / (sync_ms / 1e3) / 1e9) if sync_ms else None,
                              'reference_gbs_per_direction': 770.0, 'reference_source': 'B200_PROFILING.md measured peer copy'}
    if nvl0 and nvl1:
        line['nvlink_counters_rank0'] = {'tx_bytes_per_step': (nvl1[0] - nvl0[0]) / a.steps, 'rx_bytes_per_step': (nvl1[1] - nvl0[1]) / a.steps,
                                         'source': 'NVML NVLINK_THROUGHPUT_DATA_TX/RX over the timed region'}
    if world == 1 and not a.no_cpu_baseline:
        if sharded:
            w_in.close(); w_out.close()
            T['sharded'] = False
        del w_in, w_out, csr
        T.clear(); T.update({'sharded': False, 'synced': False})
        torch.cuda.empty_cache()
        try:
            line['cpu_baseline'] = cpu_baseline(a)
        except Exception as e:   # noqa: BLE001
            line['cpu_baseline'] = {'value': None, 'unit': UNIT, 'cores': os.cpu_count(), 'kind': 'port', 'sample': f'failed: {e!r}'}
    emit(line)
    if world > 1:
        barrier()
        close_tables()
        dist.destroy_process_group()


def run_s4(a):
    """S4 (BASELINE configs[4], w2v_sg_wiki_text_103.yaml SHAPE): SGNS over a synthetic Zipf token stream, 267,735-word
    vocabulary (+ '<unk>' row 0), sentences of max_length = 128 tokens, r = 5, K = 5, E = 128, negatives from a
    unigram^0.75 alias table.  No walk stage: a step = one batch of sentences through the fused kernel.  The tables
    (2 x 137 MB) largely live in the 126 MB L2 and the stream is Zipf-hot, so `frac` is an L2-assisted figure and may
    exceed what HBM alone could deliver; the HBM-roofline claim is made on S3."""
    import numpy as np
    import torch
    from shallow_encoders import _native as nat
    assert torch.cuda.is_available(), 'bench.py needs a CUDA device (no CPU fallback)'
    dev = torch.device('cuda', 0)
    torch.cuda.set_device(dev)
    nat.load()
    vocab = a.s4_vocab + 1
    n_seq, L = a.s4_sentences_per_step, a.s4_sentence_len
    weights = 1.0 / np.arange(1, a.s4_vocab + 1, dtype=np.float64) ** a.s4_zipf
    cdf = torch.from_numpy(np.cumsum(weights) / weights.sum()).to(dev)
    counts = np.concatenate([[0.0], weights / weights.sum() * 100e6])          # expected counts over 100 M tokens; '<unk>' never drawn
    alias = nat.alias_build(counts, a.s4_power, dev) if (a.s4_power != 0 or a.s4_force_alias) else None
    gen = torch.Generator(device=dev)
    gen.manual_seed(a.seed)
    total_steps = a.warmup + 2 * a.steps + 2

    def batch():
        u = torch.rand(n_seq * L, generator=gen, device=dev, dtype=torch.float64)
        return torch.searchsorted(cdf, u).clamp_(max=a.s4_vocab - 1).to(torch.int32).reshape(n_seq, L)      # token id = rank

    dev_tokens = [batch() for _ in range(min(total_steps, 8))]                  # 8 distinct batches (34 MB each), reused round-robin
    pinned = [t.cpu().pin_memory() for t in dev_tokens]
    bound = (6.0 / (vocab + a.emb)) ** 0.5
    w_in = torch.empty((vocab, a.emb), dtype=torch.float32, device=dev)
    w_out = torch.empty((vocab, a.emb), dtype=torch.float32, device=dev)
    nat.table_fill_uniform(w_in, bound, a.seed + 101)
    nat.table_fill_uniform(w_out, bound, a.seed + 102)
    refresh = (bool(a.window_refresh) if a.window_refresh is not None else True) and a.kernel == 'window'
    flags = nat.SCATTER_RED | (nat.NO_WINDOW if a.kernel == 'context' else 0) | (nat.WINDOW_REFRESH if refresh else 0)
    stats = torch.zeros(nat.STATS_LEN, dtype=torch.float64, device=dev)
    stats_host = torch.zeros(nat.STATS_LEN, dtype=torch.float64).pin_memory()
    tok_scratch = torch.empty((n_seq, L), dtype=torch.int32, device=dev)
    n_cen = L - 2 * a.radius
    pairs_per_step = n_seq * n_cen * 2 * a.radius

    def device_step(i):
        nat.sgns_update_walks(w_in, w_out, dev_tokens[i % len(dev_tokens)], a.radius, a.neg, 1, a.lr, a.seed + 1,
                              centre_id_base=i * n_seq * n_cen, alias=alias, flags=flags, stats=stats)

    def host_step(i):
        nat.host_sgns_update_tokens(pinned[i % len(pinned)], w_in, w_out, a.radius, a.neg, 1, a.lr, a.seed + 1, tok_scratch, stats,
                                    stats_host, centre_id_base=i * n_seq * n_cen, alias=alias, flags=flags)

    ev = lambda: torch.cuda.Event(enable_timing=True)   # noqa: E731
    for i in range(a.warmup):
        device_step(i)
    sampler = ClockSampler(nvml_index(0))
    torch.cuda.synchronize()
    launches0 = nat.launches()
    sampler.start()
    torch.cuda.profiler.start()
    t0, t1 = ev(), ev()
    t0.record()
    for i in range(a.warmup, a.warmup + a.steps):
        device_step(i)
    t1.record()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    clocks = sampler.stop()
    launches = nat.launches() - launches0
    ms_total = t0.elapsed_time(t1)
    stat_vals = stats.tolist()
    host_step(a.warmup + a.steps)
    torch.cuda.synchronize()
    w0 = time.perf_counter()
    for i in range(a.warmup + a.steps + 1, a.warmup + 2 * a.steps + 1):
        host_step(i)
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - w0) * 1e3
    peak, peak_src = measured_peak()
    window = a.kernel == 'window'
    bpp = bytes_per_pair(a.emb, a.neg, a.radius, window and '_win' in sgns_kernel_label(a.emb, a.neg, window))
    achieved = pairs_per_step * bpp / (ms_total / a.steps / 1e3) / 1e9
    line = {
        'metric': METRIC, 'value': pairs_per_step * a.steps / (ms_total / 1e3), 'unit': UNIT, 'n_gpus': 1, 'steps': a.steps, 'warmup': a.warmup,
        'ms_per_step': ms_total / a.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'S4 synthetic Zipf token stream of the wiki-103 shape: windows -> alias negatives -> SGNS update',
                   'vocab': a.s4_vocab, 'zipf_s': a.s4_zipf, 'sentence_len': L, 'sentences_per_step': n_seq, 'tokens_per_step': n_seq * L,
                   'emb': a.emb, 'context_radius': a.radius, 'negatives': a.neg,
                   'negative_sampling': f'alias table, unigram^{a.s4_power}' if alias else 'uniform (reference)', 'window_refresh': refresh,
                   'optimizer': f'in-place SGD lr {a.lr} (Hogwild, red.global.add.v4.f32)', 'parallelism': 'single GPU',
                   'l2': 'tables 2 x %.0f MB vs 126 MB L2: largely L2-resident, no flush (hot set is the point of this workload)' % (vocab * a.emb * 4 / 1e6)},
        'roofline': {'bound': 'hbm', 'kernel': sgns_kernel_label(a.emb, a.neg, window) + ' via se_sgns_update_walks',
                     'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak, 'peak_source': peak_src,
                     'algorithmic_bytes_per_pair': bpp, 'pairs_per_launch': pairs_per_step, 'traffic': None,
                     'note': 'L2-assisted: algorithmic bytes are served mostly from L2 on this workload'},
        'e2e': {'value': pairs_per_step * a.steps / (e2e_ms / 1e3), 'unit': UNIT, 'h2d_bytes_per_step': n_seq * L * 4,
                'd2h_bytes_per_step': nat.STATS_LEN * 8, 'ms_per_step': e2e_ms / a.steps,
                'api': 'se_host_sgns_update_tokens (pinned host token ids in, loss statistics out)'},
        'gpu_launches': launches, 'clocks': clocks,
        'train_stats': {'loss': (stat_vals[0] + stat_vals[1]) / max(stat_vals[4], 1), 'pairs': stat_vals[4]},
        'tokens_per_s': n_seq * L * a.steps / (ms_total / 1e3), 'library': nat.version(),
    }
    emit(line)


def main():
    a = parse_args()
    if a.lr is None:
        a.lr = 0.025 if a.workload == 's3' else 0.0025
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if world == 1 and a.gpus > 1 and a.impl != 'reference':
        # launched without torchrun: re-exec under it (the children write the JSON line to this process's stdout)
        import subprocess
        cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={a.gpus}',
               '--master-addr', '127.0.0.1', '--master-port', os.environ.get('MASTER_PORT', '29517'), *sys.argv]
        sys.exit(subprocess.call(cmd))
    quiet_stdout()
    if a.impl == 'reference':
        run_reference(a, rank, world)
        return
    if a.workload == 's4':
        if rank == 0:
            run_s4(a)
        return
    run_b200(a, rank, local_rank, world)


if __name__ == '__main__':
    main()
