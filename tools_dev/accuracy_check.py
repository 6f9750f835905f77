"""North-star gate 3: downstream node-classification accuracy of embeddings trained on the B200 path vs the reference's
CPU pattern (oracle/cpu_port.train_reference_cpu) at identical settings, seed-averaged.  Writes gpurun_out/accuracy.json (copied to profiles/)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'deepwalk-and-node2vec_b200'))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from oracle import cpu_port
from shallow_encoders.config_parser import load_config
from shallow_encoders.config_parser.core import instantiate
from tools.downstream import edge_classification, node_classification
from tools.train import train

CASES = {
    # name: (yaml, overrides common to both sides, fused overrides, seeds, n_experiments)
    'karate': ('sge_sg_karate_club', [], ['train.fused_lr=40.0'], 5, 100),
    'triplets': ('sge_sg_graph_triplets', ['train.max_epochs=60', 'train.optimizer.lr=0.05', 'train.scheduler.step_size=30'], ['train.fused_lr=4.0'], 8, 10),
    'cora_synthetic': ('sge_sg_cora', ['train.max_epochs=8', 'train.scheduler.step_size=4'], ['train.fused_lr=60.0'], 2, 10),
    # BASELINE.json configs[1]: node2vec p=1 q=0.5, dim 128 -- the shape that runs the window-resident hot kernel
    'cora_synthetic_dim128': ('sge_sg_cora', ['train.max_epochs=8', 'train.scheduler.step_size=4', 'model.embedding_size=128',
                                              'datamodule.additional_parameters.method_params.q=0.5'], ['train.fused_lr=60.0'], 2, 10),
}
if len(sys.argv) > 1:
    CASES = {k: v for k, v in CASES.items() if k in sys.argv[1:]}
out = {}
workers = os.cpu_count() or 1
for name, (yaml_name, common, fused_over, seeds, n_exp) in CASES.items():
    res = {'cpu_port': [], 'b200_reference_engine': [], 'b200_fused_adam_engine': [], 'b200_fused_engine': [], 'b200_fused_engine_edge_classification': []}
    for seed in range(seeds):
        base = common + [f'path.output_dir=/tmp/se_acc/{name}_{seed}']
        cfg = load_config(yaml_name, base)
        nc = cfg.downstream['node_classification']
        ap = cfg.datamodule.additional_parameters
        mp = ap.get('method_params', {}) or {}
        sched = cfg.train.scheduler
        # CPU: the reference pattern
        torch.manual_seed(seed)
        t0 = time.time()
        ds = cfg.datamodule.instantiate_dataset()
        w, names, _ = cpu_port.train_reference_cpu(
            ds.graph, ap['walks_per_node'], ap['walk_length'], ap.get('method', 'deepwalk'), float(mp.get('p', 1)), float(mp.get('q', 1)),
            cfg.datamodule.context_radius, cfg.model['embedding_size'], cfg.train.loss.negative_samples, cfg.datamodule.batch_size,
            cfg.train.optimizer['lr'], cfg.train.max_epochs, sched['step_size'], sched['gamma'], workers=workers, seed=seed)
        assert ['<unk>'] + names == ds.vocab.get_itos()
        acc = node_classification(w, ds.vocab.get_itos(), ds.labels, instantiate(nc['split_algorithm']), n_exp, nc.get('classifier_params'))
        res['cpu_port'].append(acc[0])
        t1 = time.time()
        for key, over in (('b200_reference_engine', ['train.engine=reference']), ('b200_fused_adam_engine', ['train.engine=fused']),
                          ('b200_fused_engine', ['train.engine=fused'] + fused_over)):
            torch.manual_seed(seed)
            cfg2 = load_config(yaml_name, base + over)
            tr, ds2 = train(cfg2, quiet=True)
            acc = node_classification(tr.model.input_embedding.numpy(), ds2.vocab.get_itos(), ds2.labels,
                                      instantiate(nc['split_algorithm']), n_exp, nc.get('classifier_params'))
            res[key].append(acc[0])
            ec = cfg.downstream.get('edge_classification')
            if key == 'b200_fused_engine' and ec:       # link prediction on the same embeddings (reference README: karate 69.5 %, triplets 85.8 %, real Cora 81.7 %)
                eacc = edge_classification(tr.model.tables[0], ds2._dataset.walk_generator.csr, ec['train_ratio'], min(ec['n_experiments'], 20),
                                           ec['operator_name'], ec.get('classifier_params'), seed=seed)
                res['b200_fused_engine_edge_classification'].append(eacc[0])
        print(name, seed, {k: round(v[-1], 4) for k, v in res.items()}, f'cpu {t1 - t0:.0f}s gpu {time.time() - t1:.0f}s', flush=True)
    out[name] = {k: {'mean': float(np.mean(v)), 'std': float(np.std(v)), 'runs': v} for k, v in res.items() if v}
    out[name]['delta_reference_engine_pp'] = 100 * (out[name]['b200_reference_engine']['mean'] - out[name]['cpu_port']['mean'])
    out[name]['delta_fused_engine_pp'] = 100 * (out[name]['b200_fused_engine']['mean'] - out[name]['cpu_port']['mean'])
    out[name]['delta_fused_adam_engine_pp'] = 100 * (out[name]['b200_fused_adam_engine']['mean'] - out[name]['cpu_port']['mean'])
    out[name]['settings'] = {'yaml': yaml_name, 'overrides': common, 'fused_overrides': fused_over, 'seeds': seeds, 'n_experiments': n_exp}
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, 'gpurun_out', 'accuracy.json'), 'w'), indent=1)
print(json.dumps({k: {kk: (round(vv['mean'], 4) if isinstance(vv, dict) and 'mean' in vv else vv) for kk, vv in v.items() if kk != 'settings'} for k, v in out.items()}, indent=1))
