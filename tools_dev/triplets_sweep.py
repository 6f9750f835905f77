import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'deepwalk-and-node2vec_b200')); sys.path.insert(0, ROOT)
import numpy as np, torch
from shallow_encoders.config_parser import load_config
from shallow_encoders.config_parser.core import instantiate
from tools.downstream import node_classification
from tools.train import train
for engine, lrs in (('fused', [5, 20, 80, 200, 500]), ('reference', [0.05])):
    for lr in lrs:
        for epochs in (20, 60):
            accs = []
            for seed in range(8):
                torch.manual_seed(seed)
                over = [f'path.output_dir=/tmp/se_tri/{seed}', f'train.engine={engine}', f'train.max_epochs={epochs}', 'train.scheduler.step_size=%d' % (epochs // 2),
                        'train.optimizer.lr=0.05', f'train.fused_lr={lr}']
                cfg = load_config('sge_sg_graph_triplets', over)
                tr, ds = train(cfg, quiet=True)
                nc = cfg.downstream['node_classification']
                accs.append(node_classification(tr.model.input_embedding.numpy(), ds.vocab.get_itos(), ds.labels, instantiate(nc['split_algorithm']), 10)[0])
            print(engine, lr, epochs, round(float(np.mean(accs)), 4), [round(a, 2) for a in accs], flush=True)
