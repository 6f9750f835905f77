"""
Downstream node classification -- the accuracy yardstick of the north star.  CPU / sklearn evaluation, restating
tools/graph_model_downstream_classification.py:94-148 of the reference (that module imports hydra + matplotlib and cannot be
imported here): X = input embedding rows 1.. ('<unk>' skipped, :116), optional feature concat (:120-123), labels mapped to
integers, `n_experiments` splits with random_state = i (:134-136), LogisticRegression(**classifier_params) fit on the train
split and scored on the test split (:85-91); mean and best accuracy (:146-148).

Command line, like the reference's tool (tools/graph_model_downstream_classification.py:300-335):

    python tools/downstream.py --config-name=sge_sg_karate_club [a.b=c ...] [--device-classifier]

loads the experiment's checkpoint (`downstream.checkpoint`, default last.ckpt; the reference reads `analysis.checkpoint`, accepted too),
runs node classification when the data set has labels and `downstream.node_classification.enable`, edge classification when
`downstream.edge_classification.enable`, prints the reference's result lines and writes them to <experiment>/analysis/downstream.json.
"""
import os
import sys
from typing import Dict, List, Optional, Tuple

import numpy as np
from sklearn.linear_model import LogisticRegression

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def node_classification_device(embedding, itos: List[str], labels: Dict[str, str], split_algorithm, n_experiments: int,
                               classifier_params: Optional[dict] = None, return_all: bool = False):
    """The same yardstick with the embedding matrix staying in HBM: `embedding` is a CUDA tensor [V x E] (row 0 = '<unk>' skipped, :116),
    the splits are the reference's (its split class is applied to the row INDICES, which shuffles exactly like applying it to the rows),
    the classifier is tools/device_classifier.DeviceLogisticRegression (same objective as sklearn's default LogisticRegression, two
    tensor-core GEMMs per gradient).  Returns (mean accuracy, best accuracy) [and the per-experiment list]."""
    import torch
    from tools.device_classifier import DeviceLogisticRegression
    x = embedding[1:].to(torch.float32).contiguous()
    vertices = itos[1:]
    classes = {c: i for i, c in enumerate(sorted(set(labels.values())))}
    y_host = np.array([classes[labels[v]] for v in vertices], dtype=np.float32)
    y = torch.from_numpy(y_host.astype(np.int64)).to(x.device)
    index = np.arange(len(vertices), dtype=np.int64).reshape(-1, 1)
    accs = []
    for i in range(n_experiments):
        split_algorithm.random_state = i
        split = split_algorithm(index, y_host)
        tr = torch.from_numpy(split['X_train'].reshape(-1)).to(x.device)
        te = torch.from_numpy(split['X_test'].reshape(-1)).to(x.device)
        clf = DeviceLogisticRegression(**(classifier_params or {})).fit(x[tr], y[tr])
        accs.append(clf.score(x[te], y[te]))
    out = (float(np.mean(accs)), float(np.max(accs)))
    return out + (accs,) if return_all else out


def node_classification(embedding: np.ndarray, itos: List[str], labels: Dict[str, str], split_algorithm,
                        n_experiments: int, classifier_params: Optional[dict] = None,
                        features: Optional[Dict[str, np.ndarray]] = None) -> Tuple[float, float]:
    """Returns (mean accuracy, best accuracy) over `n_experiments` splits."""
    X = np.asarray(embedding)[1:, :]
    vertices = itos[1:]
    if features is not None:
        X = np.concatenate([X, np.stack([features[v] for v in vertices])], axis=1)
    classes = {c: i for i, c in enumerate(sorted(set(labels.values())))}
    y = np.array([classes[labels[v]] for v in vertices], dtype=np.float32)
    total, best = 0.0, 0.0
    for i in range(n_experiments):
        split_algorithm.random_state = i
        split = split_algorithm(X, y)
        clf = LogisticRegression(**(classifier_params or {}))
        clf.fit(split['X_train'], split['y_train'])
        acc = float(np.equal(clf.predict(split['X_test']), split['y_test']).astype(np.float32).mean())
        total += acc
        best = max(best, acc)
    return total / n_experiments, best


def edge_classification(embedding, csr, train_ratio: float, n_experiments: int, edge_operator_name: str,
                        classifier_params: Optional[dict] = None, row_offset: int = 1, seed: int = 0) -> Tuple[float, float]:
    """Link prediction, restating perform_edge_classification (tools/graph_model_downstream_classification.py:227-299):
    per experiment a random `train_ratio` of the graph's edges are the positive training set, as many sampled non-edges the
    negative one (`sample_negative_edges`, :170-200); the classifier is scored on ALL edges plus the training and validation
    negatives (:268-287).  Features are built on the device (`se_edge_features`, `se_sample_negative_edges`); the logistic
    regression is sklearn on the host, as in the reference.  `embedding`: [vocab x emb] CUDA tensor (row = node id + row_offset);
    `csr`: the graph (CSRGraph).  Returns (mean accuracy, best accuracy)."""
    import torch
    from shallow_encoders import _native as nat
    from shallow_encoders.graph.edge_operators import edge_embeddings
    dev = csr.device
    embedding = embedding.to(dev, torch.float32).contiguous()
    deg = (csr.rowptr[1:] - csr.rowptr[:-1])
    src_all = torch.repeat_interleave(torch.arange(csr.n_nodes, device=dev), deg)
    dst_all = csr.col_sorted.to(torch.int64)
    keep = src_all <= dst_all                                      # each undirected edge once, like graph.edges
    e_src, e_dst = src_all[keep], dst_all[keep]
    n_edges = int(e_src.numel())
    n_train = round(train_ratio * n_edges)
    n_val = n_edges - n_train
    gen = torch.Generator(device=dev)
    total, best = 0.0, 0.0
    for i in range(n_experiments):
        gen.manual_seed(seed * 1_000_003 + i)
        perm = torch.randperm(n_edges, generator=gen, device=dev)                       # random.shuffle(edges), :252
        tr_pos = perm[:n_train]
        neg_src, neg_dst = nat.sample_negative_edges(csr, n_train + n_val, seed * 7919 + i)
        neg_src, neg_dst = neg_src.to(torch.int64), neg_dst.to(torch.int64)
        tr_s = torch.cat([e_src[tr_pos], neg_src[:n_train]]) + row_offset
        tr_d = torch.cat([e_dst[tr_pos], neg_dst[:n_train]]) + row_offset
        all_s = torch.cat([e_src, neg_src]) + row_offset
        all_d = torch.cat([e_dst, neg_dst]) + row_offset
        X_train = edge_embeddings(embedding, tr_s, tr_d, edge_operator_name).cpu().numpy()
        X = edge_embeddings(embedding, all_s, all_d, edge_operator_name).cpu().numpy()
        y_train = np.array(n_train * [1] + n_train * [0], dtype=np.float32)
        y = np.array(n_edges * [1] + (n_train + n_val) * [0], dtype=np.float32)
        clf = LogisticRegression(**(classifier_params or {}))
        clf.fit(X_train, y_train)
        acc = float(np.equal(clf.predict(X), y).astype(np.float32).mean())
        total += acc
        best = max(best, acc)
    return total / n_experiments, best


def run_downstream(cfg, device_classifier: bool = False) -> Dict[str, Dict[str, float]]:
    """The body of the reference's `main` (:301-331) for a loaded config: checkpoint -> model -> the enabled downstream tasks."""
    import json
    from shallow_encoders.config_parser.core import instantiate
    assert cfg.datamodule.is_graph, 'This script supports only graph datasets!'
    base = os.path.join(cfg.path.output_dir, cfg.datamodule.dataset_name, cfg.train.experiment)
    dataset = cfg.datamodule.instantiate_dataset()
    checkpoint = cfg.downstream.get('checkpoint') or (cfg.analysis or {}).get('checkpoint', 'last.ckpt')
    trainer = cfg.instantiate_trainer(dataset=dataset, checkpoint_path=os.path.join(base, 'checkpoints', checkpoint))
    results: Dict[str, Dict[str, float]] = {}
    nc = cfg.downstream.get('node_classification') or {}
    if nc.get('enable') and dataset.has_labels:
        split_algorithm = instantiate(nc['split_algorithm'])
        itos = dataset.vocab.get_itos()
        if device_classifier and not dataset.has_features:      # (features live on the host: the sklearn path concatenates them)
            mean_acc, best_acc = node_classification_device(trainer.model.tables[0], itos, dataset.labels, split_algorithm, nc['n_experiments'],
                                                            nc.get('classifier_params'))
        else:
            mean_acc, best_acc = node_classification(trainer.model.input_embedding.numpy(), itos, dataset.labels, split_algorithm,
                                                     nc['n_experiments'], nc.get('classifier_params'),
                                                     features=dataset.features if dataset.has_features else None)      # :120-123
        results['node_classification'] = {'mean_accuracy': mean_acc, 'best_accuracy': best_acc}
        print(f"Node classification accuracy: {100 * mean_acc:.2f}% (averaged over {nc['n_experiments']} experiments).")
        print(f'Best accuracy score: {100 * best_acc:.2f}%.')
    ec = cfg.downstream.get('edge_classification') or {}
    if ec.get('enable'):
        csr = dataset._dataset.walk_generator.csr
        mean_acc, best_acc = edge_classification(trainer.model.tables[0], csr, ec['train_ratio'], ec['n_experiments'], ec['operator_name'],
                                                 ec.get('classifier_params'), row_offset=dataset.row_offset)
        results['edge_classification'] = {'mean_accuracy': mean_acc, 'best_accuracy': best_acc}
        print(f"Edge classification accuracy: {100 * mean_acc:.2f}% (averaged over {ec['n_experiments']} experiments).")
        print(f'Best accuracy score: {100 * best_acc:.2f}%.')
    os.makedirs(os.path.join(base, 'analysis'), exist_ok=True)
    with open(os.path.join(base, 'analysis', 'downstream.json'), 'w', encoding='utf-8') as f:
        json.dump(results, f, indent=1)
    return results


def main(argv=None):
    import argparse
    from shallow_encoders.config_parser import load_config
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument('--config-name', required=True)
    ap.add_argument('--device-classifier', action='store_true',
                    help='fit the node classifier on the device (tools/device_classifier.py) instead of sklearn on a host copy of the embeddings')
    ap.add_argument('overrides', nargs='*')
    a = ap.parse_args(argv)
    return run_downstream(load_config(a.config_name, a.overrides), device_classifier=a.device_classifier)


if __name__ == '__main__':
    main()
