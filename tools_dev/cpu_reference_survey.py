"""the UNMODIFIED reference timed on this box's host cores (oracle/ref_pipeline.py, one fresh process per
measurement) -- walk steps/s on S1 karate / S2 Cora-shape / an S3 sample at 1 process and all cores, training_step + backward + optimizer
pairs/s at the S1 / S2 / S4 shapes with the YAML's Adam and with SGD, and the reference's own DataLoader path with num_workers 0 and 8.
Writes gpurun_out/cpu_reference_survey.json (copied to profiles/)."""
import json
import os
import platform
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_pipeline  # noqa: E402

cores = os.cpu_count() or 1
out = {'host': {'cpu_count': cores, 'machine': platform.machine(), 'processor': platform.processor(), 'python': sys.version.split()[0]}, 'walks': [], 'sgns': [],
       'dataloader': []}
try:
    out['host']['model'] = [l.split(':')[1].strip() for l in open('/proc/cpuinfo') if l.startswith('model name')][0]
except Exception:   # noqa: BLE001
    pass
for graph, method, p, q, L, nodes, edges in (('karate', 'node2vec', 1, 0.5, 10, 0, 0), ('karate', 'deepwalk', 1, 1, 10, 0, 0),
                                             ('cora_shape', 'node2vec', 1, 0.5, 10, 0, 0), ('cora_shape', 'deepwalk', 1, 1, 10, 0, 0),
                                             ('powerlaw', 'node2vec', 0.5, 2.0, 80, 100_000, 2_500_000)):
    for workers in (1, cores):
        steps = 2 if graph == 'powerlaw' else 20
        r = ref_pipeline.call(mode='walks', graph=graph, method=method, p=p, q=q, walk_len=L, nodes=max(nodes, 1), edges=max(edges, 1),
                              walks_per_step=64 if graph == 'powerlaw' else 256, steps=steps * (workers if graph != 'powerlaw' else 1), workers=workers)
        out['walks'].append(r)
        print('walks', graph, method, workers, r.get('walk_steps_per_s'), flush=True)
for name, vocab, emb, rows, radius, neg in (('S1 karate', 35, 2, 384, 2, 1), ('S2 cora-shape E=128', 2709, 128, 384, 2, 5), ('S2 cora yaml E=8', 2709, 8, 384, 2, 5),
                                            ('S4 wiki-103 shape E=128 K=5', 267736, 128, 1888, 5, 5), ('S4 wiki-103 yaml E=48 K=3', 267736, 48, 1888, 5, 3)):
    for optimizer in ('adam', 'sgd'):
        r = ref_pipeline.call(mode='sgns', vocab=vocab, emb=emb, batch_rows=rows, radius=radius, neg=neg, optimizer=optimizer, steps=5 if vocab > 10000 else 30, warmup=1)
        r['shape'] = name
        out['sgns'].append(r)
        print('sgns', name, optimizer, r.get('pairs_per_s'), flush=True)
for workers in (0, 8):
    r = ref_pipeline.call(mode='dataloader', workers=workers, steps=2)
    out['dataloader'].append(r)
    print('dataloader', workers, r.get('pairs_per_s'), r.get('rows_per_epoch'), flush=True)
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, 'gpurun_out', 'cpu_reference_survey.json'), 'w'), indent=1)
