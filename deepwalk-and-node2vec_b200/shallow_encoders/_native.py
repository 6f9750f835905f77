"""
ctypes binding of the C ABI in include/se_b200.h (deepwalk-and-node2vec_b200/lib/libse_b200.so).

This is the ONLY compute path of the package: every wrapper raises if the library is missing, if a symbol
declared in the header is not exported, or if its tensors are not CUDA tensors.  torch is used for device
memory and streams only -- pointers and sizes cross the boundary, never torch types.
"""
import ctypes
import os
import re
from typing import Dict, Optional

import torch

PKG_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REPO_ROOT = os.path.dirname(PKG_ROOT)
LIB_PATH = os.environ.get('SE_B200_LIB', os.path.join(PKG_ROOT, 'lib', 'libse_b200.so'))
HEADER_PATH = os.path.join(REPO_ROOT, 'include', 'se_b200.h')

SE_OK = 0
RULE_REFERENCE = 0
RULE_PAPER = 1
SCATTER_RED = 0
SCATTER_STORE = 1
GENERIC_KERNEL = 2
NO_WINDOW = 4
WHOLE_SEQUENCES = 8
WINDOW_REFRESH = 16
BATCHED_POSITIVES = 32
STATS_LEN = 6
WALK_AUTO, WALK_WARP, WALK_THREAD = 0, 1, 2
EDGE_OPS = {'average': 0, 'hadamard': 1, 'weighted_l1': 2, 'weighted_l2': 3}

_lib = None
_launches = 0    # kernels launched through this module (bench.py reports it as gpu_launches)

c_i64, c_i32, c_int, c_f64, c_f32, c_u64, c_p = (ctypes.c_int64, ctypes.c_int32, ctypes.c_int, ctypes.c_double,
                                                 ctypes.c_float, ctypes.c_uint64, ctypes.c_void_p)

_SIGNATURES = {
    'se_version': (ctypes.c_char_p, []),
    'se_last_error': (ctypes.c_char_p, []),
    'se_device_info': (c_int, [c_p, c_p, c_p]),
    'se_walk_exact_scratch_bytes': (c_i64, [c_i64, c_i64]),
    'se_walk_exact': (c_int, [c_p, c_p, c_p, c_p, c_int, c_i64, c_i64, c_p, c_i64, c_int, c_f64, c_f64, c_int, c_int,
                              c_p, c_p, c_i64, c_p, c_p]),
    'se_walk': (c_int, [c_p, c_p, c_p, c_i64, c_int, c_p, c_i64, c_int, c_f64, c_f64, c_int, c_int, c_u64, c_i64, c_i64,
                        c_p, c_p, c_int, c_p]),
    'se_alias_build_host': (c_int, [c_p, c_i64, c_f64, c_p, c_p]),
    'se_sample_negatives': (c_int, [c_p, c_p, c_i64, c_u64, c_i64, c_i64, c_p, c_p]),
    'se_skipgram_scores': (c_int, [c_p, c_p, c_i64, c_int, c_p, c_p, c_i64, c_int, c_int, c_p, c_p]),
    'se_skipgram_scores_backward': (c_int, [c_p, c_p, c_i64, c_int, c_p, c_p, c_i64, c_int, c_p, c_p, c_p, c_p]),
    'se_cbow_scores': (c_int, [c_p, c_p, c_i64, c_int, c_p, c_p, c_i64, c_int, c_int, c_int, c_p, c_p]),
    'se_cbow_grad': (c_int, [c_p, c_p, c_i64, c_int, c_p, c_p, c_p, c_i64, c_int, c_int, c_int, c_p, c_p, c_p, c_p]),
    'se_ns_loss': (c_int, [c_p, c_p, c_i64, c_int, c_int, c_p, c_p, c_p, c_p]),
    'se_sgns_grad': (c_int, [c_p, c_p, c_i64, c_int, c_p, c_p, c_p, c_i64, c_int, c_int, c_p, c_p, c_p, c_p]),
    'se_sgns_adam_step': (c_int, [c_p, c_p, c_i64, c_int, c_p, c_p, c_p, c_i64, c_int, c_int, c_p, c_f32, c_f32, c_f32, c_f32, c_p, c_p]),
    'se_sgns_step': (c_int, [c_p, c_p, c_i64, c_int, c_p, c_p, c_p, c_i64, c_int, c_int, c_p, c_p, c_f32, c_u64, c_i64,
                             c_int, c_p, c_p]),
    'se_sgns_update_walks': (c_int, [c_p, c_p, c_i64, c_int, c_p, c_i64, c_int, c_int, c_int, c_int, c_p, c_p, c_f32,
                                     c_u64, c_i64, c_int, c_p, c_p]),
    'se_host_walk_sgns_step': (c_int, [c_p, c_p, c_p, c_i64, c_int, c_p, c_i64, c_int, c_f64, c_f64, c_int, c_int, c_u64,
                                       c_i64, c_p, c_p, c_i64, c_int, c_int, c_int, c_int, c_p, c_p, c_f32, c_int,
                                       c_p, c_p, c_p, c_p, c_p, c_p]),
    'se_shard_granularity': (c_int, [c_p]),
    'se_shard_reserve': (c_int, [c_i64, c_p]),
    'se_shard_unreserve': (c_int, [c_u64, c_i64]),
    'se_shard_create': (c_int, [c_i64, c_p]),
    'se_shard_release': (c_int, [c_u64]),
    'se_shard_export_fd': (c_int, [c_u64, c_p]),
    'se_shard_import_fd': (c_int, [c_int, c_p]),
    'se_shard_map': (c_int, [c_u64, c_i64, c_u64]),
    'se_shard_unmap': (c_int, [c_u64, c_i64]),
    'se_shard_local_rows': (c_int, [c_i64, c_p, c_p]),
    'se_sgns_update_walks_sharded': (c_int, [c_p, c_p, c_i64, c_int, c_p, c_i64, c_int, c_int, c_int, c_int, c_p, c_p,
                                             c_f32, c_u64, c_i64, c_int, c_p, c_p, c_p]),
    'se_sgns_update_negatives_owned': (c_int, [c_p, c_p, c_i64, c_int, c_p, c_i64, c_int, c_int, c_int, c_int, c_p, c_p, c_f32, c_u64,
                                               c_i64, c_p, c_p, c_p]),
    'se_pairs_owned_scratch_bytes': (c_i64, [c_i64, c_i64, c_int, c_int]),
    'se_sgns_update_pairs_owned': (c_int, [c_p, c_p, c_i64, c_int, c_p, c_i64, c_int, c_int, c_int, c_int, c_int, c_p, c_p, c_f32,
                                           c_u64, c_i64, c_p, c_p, c_i64, c_p, c_p]),
    'se_host_walk_sgns_step_sharded': (c_int, [c_p, c_p, c_p, c_i64, c_int, c_p, c_i64, c_int, c_f64, c_f64, c_int, c_int,
                                               c_u64, c_i64, c_p, c_p, c_i64, c_int, c_int, c_int, c_int, c_p, c_p, c_f32,
                                               c_int, c_p, c_p, c_p, c_p, c_p, c_p, c_p]),
    'se_host_sgns_update_tokens': (c_int, [c_p, c_i64, c_int, c_p, c_p, c_i64, c_int, c_int, c_int, c_int, c_p, c_p, c_f32, c_u64, c_i64,
                                           c_int, c_p, c_p, c_p, c_p, c_p]),
    'se_edge_features': (c_int, [c_p, c_i64, c_int, c_p, c_p, c_i64, c_int, c_p, c_p]),
    'se_edge_op': (c_int, [c_p, c_p, c_i64, c_int, c_p, c_p]),
    'se_sample_negative_edges': (c_int, [c_p, c_p, c_i64, c_i64, c_u64, c_i64, c_p, c_p, c_p, c_p]),
    'se_check_ids': (c_int, [c_p, c_i64, c_i64, c_i64, c_p, c_p]),
    'se_replica_chunk': (c_int, [c_i64, c_int, c_int, c_p, c_p]),
    'se_replica_sync': (c_int, [c_p, c_i64, c_int, c_int, c_i64, c_p, c_int, c_f32, c_p]),
    'se_table_renorm_rows': (c_int, [c_p, c_int, c_p, c_i64, c_f32, c_p]),
    'se_softmax_xent': (c_int, [c_p, c_p, c_p, c_i64, c_int, c_f32, c_p, c_p, c_p]),
    'se_csr_build_scratch_bytes': (c_i64, [c_i64, c_i64, c_int]),
    'se_csr_build': (c_int, [c_p, c_p, c_p, c_i64, c_i64, c_int, c_p, c_i64, c_p, c_p, c_p, c_p, c_p, c_p]),
    'se_gemm_nt': (c_int, [c_p, c_p, c_i64, c_i64, c_int, c_p, c_p, c_p, c_p]),
    'se_row_inv_norms': (c_int, [c_p, c_i64, c_int, c_p, c_p]),
    'se_cosine_similarity': (c_int, [c_p, c_p, c_i64, c_i64, c_int, c_p, c_p, c_p]),
    'se_topk_rows': (c_int, [c_p, c_i64, c_i64, c_int, c_p, c_p, c_p]),
    'se_transpose': (c_int, [c_p, c_i64, c_i64, c_p, c_p]),
    'se_shared_negatives_scratch_floats': (c_i64, [c_i64, c_i64, c_int]),
    'se_sgns_step_shared_negatives': (c_int, [c_p, c_p, c_i64, c_int, c_p, c_i64, c_p, c_i64, c_int, c_int, c_f32, c_p, c_i64, c_p, c_p]),
    'se_table_fill_uniform': (c_int, [c_p, c_i64, c_f32, c_u64, c_i64, c_int, c_int, c_p]),
    'se_table_gather_rows': (c_int, [c_p, c_int, c_p, c_i64, c_p, c_p]),
    'se_table_scatter_rows': (c_int, [c_p, c_int, c_p, c_i64, c_p, c_p]),
}


class ShardSpec(ctypes.Structure):
    """struct se_shard_spec (include/se_b200.h)."""
    _fields_ = [('world', c_i32), ('rank', c_i32), ('stripe_rows', c_i64), ('local_negatives', c_i32), ('reserved', c_i32)]


class AdamStateC(ctypes.Structure):
    """struct se_adam_state (include/se_b200.h)."""
    _fields_ = [('m_in', c_p), ('v_in', c_p), ('m_out', c_p), ('v_out', c_p), ('g_in', c_p), ('g_out', c_p), ('t_in', c_p), ('t_out', c_p),
                ('touched_in', c_p), ('touched_out', c_p), ('list_in', c_p), ('list_out', c_p), ('list_in_capacity', c_i64),
                ('list_out_capacity', c_i64), ('counts', c_p)]


def header_symbols():
    """Function names declared in include/se_b200.h."""
    text = open(HEADER_PATH).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(se_[a-z0-9_]+)\s*\(', text)))


def load():
    """dlopen the native library and bind every symbol of the header.  Does not need a GPU."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f'native library not found at {LIB_PATH}: build it with `python __graft_entry__.py` '
            f'(nvcc -gencode arch=compute_100a,code=sm_100a). There is no CPU fallback.')
    lib = ctypes.CDLL(LIB_PATH)
    declared = header_symbols()
    missing = [s for s in declared if not hasattr(lib, s)]
    if missing:
        raise RuntimeError(f'{LIB_PATH} does not export {missing} declared in {HEADER_PATH}; rebuild it')
    unbound = [s for s in declared if s not in _SIGNATURES]
    if unbound:
        raise RuntimeError(f'_native.py has no ctypes signature for {unbound}')
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def launches() -> int:
    return _launches


def _check(rc: int):
    if rc != SE_OK:
        msg = load().se_last_error().decode()
        if rc == -1:
            raise AssertionError(msg)      # the reference signals bad arguments with `assert`
        raise RuntimeError(f'se_b200 error {rc}: {msg}')


def _ptr(t: Optional[torch.Tensor], dtype=None, name='tensor') -> Optional[int]:
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError(f'{name} must be a CUDA tensor: the B200 path has no CPU fallback')
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f'{name} must be {dtype}, got {t.dtype}')
    if not t.is_contiguous():
        raise ValueError(f'{name} must be contiguous')
    return t.data_ptr()


def _on(t: torch.Tensor):
    """Device guard; refuses CPU tensors before anything else happens."""
    if not t.is_cuda:
        raise RuntimeError('expected a CUDA tensor: the B200 path has no CPU fallback')
    return torch.cuda.device(t.device)


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _table(t, name='table'):
    """(device pointer, vocab, emb, shard spec or None, device guard) of a torch table or a sharded table
    (shallow_encoders/word2vec/sharded.py: anything with `.ptr`, `.vocab`, `.emb`, `.spec()`, `.device`)."""
    if isinstance(t, torch.Tensor):
        return _ptr(t, torch.float32, name), t.shape[0], t.shape[1], None, t.device
    if hasattr(t, 'ptr') and hasattr(t, 'spec'):
        return int(t.ptr), int(t.vocab), int(t.emb), t.spec(), t.device
    raise TypeError(f'{name} must be a CUDA float32 tensor or a ShardedTable, got {type(t)}')


def _same_sharding(si, so):
    if (si is None) != (so is None):
        raise ValueError('w_in and w_out must both be torch tensors or both be ShardedTables')
    if si is not None and si.world == 1 and so.world > 1:
        # hybrid: W_in is this GPU's working copy of a ReplicatedTable (a local table to the kernels), W_out one striped table
        return so
    if si is not None and (si.world, si.rank, si.stripe_rows) != (so.world, so.rank, so.stripe_rows):
        raise ValueError('w_in and w_out must be sharded the same way')
    return si


def version() -> str:
    return load().se_version().decode()


def device_info() -> Dict[str, int]:
    sm, major, minor = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    _check(load().se_device_info(ctypes.byref(sm), ctypes.byref(major), ctypes.byref(minor)))
    return {'sm_count': sm.value, 'cc_major': major.value, 'cc_minor': minor.value}


# ----------------------------------------------------------------------------------------------------------------
# walks
# ----------------------------------------------------------------------------------------------------------------
def walk_exact(csr, starts: torch.Tensor, walk_len: int, p: float, q: float, node2vec: bool, rule: int,
               uniforms: torch.Tensor) -> torch.Tensor:
    """Reference-exact walks under supplied uniforms (one per transition) -> int32 [n_walks, walk_len]."""
    global _launches
    lib = load()
    n = starts.numel()
    out = torch.empty((n, walk_len), dtype=torch.int32, device=starts.device)
    nbytes = lib.se_walk_exact_scratch_bytes(csr.max_degree, n)
    scratch = torch.empty(max(int(nbytes), 8), dtype=torch.uint8, device=starts.device)
    with _on(starts):
        _check(lib.se_walk_exact(
            _ptr(csr.rowptr, torch.int64, 'rowptr'), _ptr(csr.col, torch.int32, 'col'),
            _ptr(csr.col_sorted, torch.int32, 'col_sorted'), _ptr(csr.w, torch.float64, 'w'), int(csr.w_is_int),
            csr.n_nodes, csr.max_degree, _ptr(starts, torch.int32, 'starts'), n, int(walk_len), float(p), float(q),
            int(bool(node2vec)), int(rule), _ptr(uniforms, torch.float64, 'uniforms'), scratch.data_ptr(),
            scratch.numel(), out.data_ptr(), _stream()))
    _launches += 1
    return out


def walk(csr, starts: torch.Tensor, walk_len: int, p: float, q: float, node2vec: bool, rule: int, seed: int,
         walk_id_base: int = 0, walk_id_stride: int = 1, out: Optional[torch.Tensor] = None,
         err_count: Optional[torch.Tensor] = None, kernel: int = WALK_AUTO) -> torch.Tensor:
    """Philox/rejection walks -> int32 [n_walks, walk_len]."""
    global _launches
    lib = load()
    n = starts.numel()
    if out is None:
        out = torch.empty((n, walk_len), dtype=torch.int32, device=starts.device)
    assert out.numel() >= n * walk_len
    with _on(starts):
        _check(lib.se_walk(
            _ptr(csr.rowptr, torch.int64, 'rowptr'), _ptr(csr.col_sorted, torch.int32, 'col_sorted'),
            _ptr(csr.wcdf, torch.float32, 'wcdf'), csr.n_nodes, int(csr.symmetric), _ptr(starts, torch.int32, 'starts'),
            n, int(walk_len), float(p), float(q), int(bool(node2vec)), int(rule), int(seed) & (2 ** 64 - 1),
            int(walk_id_base), int(walk_id_stride), _ptr(out, torch.int32, 'out'),
            _ptr(err_count, torch.int32, 'err_count'), int(kernel), _stream()))
    _launches += 1
    return out


# ----------------------------------------------------------------------------------------------------------------
# negatives
# ----------------------------------------------------------------------------------------------------------------
def alias_build(counts, power: float, device) -> Dict[str, torch.Tensor]:
    """Vose alias table for weights counts**power (host set-up), uploaded to `device`."""
    import numpy as np
    counts = np.ascontiguousarray(counts, dtype=np.float64)
    prob = np.empty(len(counts), dtype=np.float32)
    alias = np.empty(len(counts), dtype=np.int32)
    _check(load().se_alias_build_host(counts.ctypes.data, len(counts), float(power), prob.ctypes.data, alias.ctypes.data))
    return {'prob': torch.from_numpy(prob).to(device), 'alias': torch.from_numpy(alias).to(device)}


def sample_negatives(n: int, vocab: int, seed: int, device, alias: Optional[Dict[str, torch.Tensor]] = None,
                     draw_id_base: int = 0) -> torch.Tensor:
    global _launches
    out = torch.empty(n, dtype=torch.int64, device=device)
    with _on(out):
        _check(load().se_sample_negatives(
            _ptr(alias['prob'], torch.float32) if alias else None, _ptr(alias['alias'], torch.int32) if alias else None,
            int(vocab), int(seed) & (2 ** 64 - 1), int(draw_id_base), int(n), out.data_ptr(), _stream()))
    _launches += 1
    return out


# ----------------------------------------------------------------------------------------------------------------
# SGNS
# ----------------------------------------------------------------------------------------------------------------
def _stats_dict(stats: torch.Tensor) -> Dict[str, float]:
    s = stats.tolist()
    pairs = max(s[4], 1.0)
    return {
        'loss': (s[0] + s[1]) / pairs, 'positive-loss': s[0] / pairs, 'negative-loss': s[1] / pairs,
        'recall': s[2] / pairs, 'precision': 1.0 - s[3] / max(s[5], 1.0), 'pairs': int(s[4]), 'negatives': int(s[5]),
    }


def skipgram_scores(w_in, w_out, inputs: torch.Tensor, outputs: torch.Tensor, proba: bool) -> torch.Tensor:
    """SkipGram.forward on torch tables or striped ShardedTables."""
    global _launches
    batch, m = outputs.shape
    p_in, vocab, emb, _, dev = _table(w_in, 'w_in')
    p_out, _, _, _, _ = _table(w_out, 'w_out')
    out = torch.empty((batch, m), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _check(load().se_skipgram_scores(
            p_in, p_out, vocab, emb,
            _ptr(inputs.reshape(-1), torch.int64, 'inputs'), _ptr(outputs, torch.int64, 'outputs'), batch, m,
            int(bool(proba)), out.data_ptr(), _stream()))
    _launches += 1
    return out


def skipgram_scores_backward(w_in: torch.Tensor, w_out: torch.Tensor, inputs: torch.Tensor, outputs: torch.Tensor,
                             grad_scores: torch.Tensor, grad_in: torch.Tensor, grad_out: torch.Tensor) -> None:
    """Accumulate the dense gradients of the raw scores into grad_in / grad_out."""
    global _launches
    batch, m = outputs.shape
    with _on(w_in):
        _check(load().se_skipgram_scores_backward(
            _ptr(w_in, torch.float32, 'w_in'), _ptr(w_out, torch.float32, 'w_out'), w_in.shape[0], w_in.shape[1],
            _ptr(inputs.reshape(-1), torch.int64, 'inputs'), _ptr(outputs, torch.int64, 'outputs'), batch, m,
            _ptr(grad_scores, torch.float32, 'grad_scores'), _ptr(grad_in, torch.float32, 'grad_in'),
            _ptr(grad_out, torch.float32, 'grad_out'), _stream()))
    _launches += 1


def cbow_scores(w_in: torch.Tensor, w_out: torch.Tensor, inputs: torch.Tensor, outputs: torch.Tensor, proba: bool) -> torch.Tensor:
    """CBOW.forward: inputs (B, N) context ids, outputs (B, M) -> scores (B, M)."""
    global _launches
    batch, m = outputs.shape
    out = torch.empty((batch, m), dtype=torch.float32, device=w_in.device)
    with _on(w_in):
        _check(load().se_cbow_scores(_ptr(w_in, torch.float32, 'w_in'), _ptr(w_out, torch.float32, 'w_out'), w_in.shape[0], w_in.shape[1],
                                     _ptr(inputs, torch.int64, 'inputs'), _ptr(outputs, torch.int64, 'outputs'), batch, inputs.shape[1], m,
                                     int(bool(proba)), out.data_ptr(), _stream()))
    _launches += 1
    return out


def cbow_grad(w_in: torch.Tensor, w_out: torch.Tensor, inputs: torch.Tensor, targets: torch.Tensor, noise: torch.Tensor,
              want_grads: bool = True, stats: Optional[torch.Tensor] = None) -> Dict:
    """CBOW training step: loss dict (+ dense grads of the mean loss) for inputs (B, N), targets (B, M), noise (B, M, K)."""
    global _launches
    batch, m = targets.shape
    n_neg = noise.shape[2] if noise is not None and noise.dim() == 3 else 0
    own_stats = stats is None
    if own_stats:
        stats = torch.zeros(STATS_LEN, dtype=torch.float64, device=w_in.device)
    g_in = torch.zeros_like(w_in) if want_grads else None
    g_out = torch.zeros_like(w_out) if want_grads else None
    with _on(w_in):
        _check(load().se_cbow_grad(_ptr(w_in, torch.float32, 'w_in'), _ptr(w_out, torch.float32, 'w_out'), w_in.shape[0], w_in.shape[1],
                                   _ptr(inputs, torch.int64, 'inputs'), _ptr(targets, torch.int64, 'targets'),
                                   _ptr(noise, torch.int64, 'noise') if n_neg else None, batch, inputs.shape[1], m, n_neg, stats.data_ptr(),
                                   _ptr(g_in), _ptr(g_out), _stream()))
    _launches += 1
    out = _stats_dict(stats) if own_stats else {}
    out['grad_in'], out['grad_out'] = g_in, g_out
    return out


def ns_loss(pos_logits: torch.Tensor, neg_logits: torch.Tensor, want_grads: bool = True):
    """(stats double[6] on device, grad_pos, grad_neg) for logits (B,N) / (B,N,K)."""
    global _launches
    batch, n_ctx = pos_logits.shape
    n_neg = neg_logits.shape[2] if neg_logits.dim() == 3 else 0
    stats = torch.zeros(STATS_LEN, dtype=torch.float64, device=pos_logits.device)
    g_pos = torch.empty_like(pos_logits) if want_grads else None
    g_neg = torch.empty_like(neg_logits) if want_grads else None
    with _on(pos_logits):
        _check(load().se_ns_loss(
            _ptr(pos_logits, torch.float32, 'pos_logits'), _ptr(neg_logits, torch.float32, 'neg_logits') if n_neg else None,
            batch, n_ctx, n_neg, stats.data_ptr(), _ptr(g_pos), _ptr(g_neg) if n_neg else None, _stream()))
    _launches += 1
    return stats, g_pos, g_neg


def sgns_grad(w_in: torch.Tensor, w_out: torch.Tensor, inputs: torch.Tensor, targets: torch.Tensor,
              noise: torch.Tensor, want_grads: bool = True, stats: Optional[torch.Tensor] = None) -> Dict:
    """Loss dict (+ dense grads of the mean loss) for an explicit (inputs (B,1), targets (B,N), noise (B,N,K)) batch."""
    global _launches
    batch, n_ctx = targets.shape
    n_neg = noise.shape[2] if noise is not None and noise.dim() == 3 else 0
    own_stats = stats is None
    if own_stats:
        stats = torch.zeros(STATS_LEN, dtype=torch.float64, device=w_in.device)
    g_in = torch.zeros_like(w_in) if want_grads else None
    g_out = torch.zeros_like(w_out) if want_grads else None
    with _on(w_in):
        _check(load().se_sgns_grad(
            _ptr(w_in, torch.float32, 'w_in'), _ptr(w_out, torch.float32, 'w_out'), w_in.shape[0], w_in.shape[1],
            _ptr(inputs.reshape(-1), torch.int64, 'inputs'), _ptr(targets, torch.int64, 'targets'),
            _ptr(noise, torch.int64, 'noise') if n_neg else None, batch, n_ctx, n_neg, stats.data_ptr(),
            _ptr(g_in), _ptr(g_out), _stream()))
    _launches += 1
    out = _stats_dict(stats) if own_stats else {}
    out['grad_in'], out['grad_out'] = g_in, g_out
    return out


class AdamState:
    """Device state of the row-sparse Adam (se_adam_state): moments, gradient accumulators, per-row step counts, touched-row
    flags and lists.  Everything is zero-initialised; the row lists grow with the largest batch seen."""

    def __init__(self, vocab: int, emb: int, device):
        z = lambda *shape, dtype=torch.float32: torch.zeros(shape, dtype=dtype, device=device)   # noqa: E731
        self.vocab, self.emb, self.device = int(vocab), int(emb), device
        self.m_in, self.v_in, self.m_out, self.v_out = z(vocab, emb), z(vocab, emb), z(vocab, emb), z(vocab, emb)
        self.g_in, self.g_out = z(vocab, emb), z(vocab, emb)
        self.t_in, self.t_out = z(vocab, dtype=torch.int32), z(vocab, dtype=torch.int32)
        self.touched_in, self.touched_out = z(vocab, dtype=torch.int32), z(vocab, dtype=torch.int32)
        self.counts = z(4, dtype=torch.int32)
        self.list_in = self.list_out = None

    def ensure(self, batch: int, n_ctx: int, n_neg: int) -> None:
        need_in, need_out = min(self.vocab, batch), min(self.vocab, batch * n_ctx * (1 + n_neg))
        if self.list_in is None or self.list_in.numel() < need_in:
            self.list_in = torch.zeros(need_in, dtype=torch.int32, device=self.device)
        if self.list_out is None or self.list_out.numel() < need_out:
            self.list_out = torch.zeros(need_out, dtype=torch.int32, device=self.device)

    def c_struct(self) -> AdamStateC:
        p = lambda t: t.data_ptr()   # noqa: E731
        return AdamStateC(p(self.m_in), p(self.v_in), p(self.m_out), p(self.v_out), p(self.g_in), p(self.g_out), p(self.t_in), p(self.t_out),
                          p(self.touched_in), p(self.touched_out), p(self.list_in), p(self.list_out), self.list_in.numel(),
                          self.list_out.numel(), p(self.counts))


def sgns_adam_step(w_in: torch.Tensor, w_out: torch.Tensor, inputs: torch.Tensor, targets: torch.Tensor, noise: Optional[torch.Tensor],
                   state: AdamState, lr: float, beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8,
                   stats: Optional[torch.Tensor] = None) -> Optional[Dict[str, float]]:
    """Row-sparse Adam on an explicit (inputs (B,1), targets (B,N), noise (B,N,K)) batch: gradient of the mean loss + update of the
    touched rows of both tables, three launches."""
    global _launches
    batch, n_ctx = targets.shape
    n_neg = noise.shape[2] if noise is not None and noise.dim() == 3 else 0
    state.ensure(batch, n_ctx, n_neg)
    own_stats = stats is None
    if own_stats:
        stats = torch.zeros(STATS_LEN, dtype=torch.float64, device=w_in.device)
    cs = state.c_struct()
    with _on(w_in):
        _check(load().se_sgns_adam_step(
            _ptr(w_in, torch.float32, 'w_in'), _ptr(w_out, torch.float32, 'w_out'), w_in.shape[0], w_in.shape[1],
            _ptr(inputs.reshape(-1), torch.int64, 'inputs'), _ptr(targets, torch.int64, 'targets'),
            _ptr(noise, torch.int64, 'noise') if n_neg else None, batch, n_ctx, n_neg, ctypes.byref(cs), float(lr), float(beta1),
            float(beta2), float(eps), stats.data_ptr(), _stream()))
    _launches += 3
    return _stats_dict(stats) if own_stats else None


def sgns_step(w_in: torch.Tensor, w_out: torch.Tensor, inputs: torch.Tensor, targets: torch.Tensor,
              noise: Optional[torch.Tensor], n_neg: int, lr: float, seed: int = 0, pair_id_base: int = 0,
              alias: Optional[Dict[str, torch.Tensor]] = None, flags: int = SCATTER_RED,
              stats: Optional[torch.Tensor] = None) -> Optional[Dict[str, float]]:
    """Fused in-place SGD on an explicit batch; `lr` multiplies the un-averaged per-pair gradient."""
    global _launches
    batch, n_ctx = targets.shape
    own_stats = stats is None
    if own_stats:
        stats = torch.zeros(STATS_LEN, dtype=torch.float64, device=w_in.device)
    with _on(w_in):
        _check(load().se_sgns_step(
            _ptr(w_in, torch.float32, 'w_in'), _ptr(w_out, torch.float32, 'w_out'), w_in.shape[0], w_in.shape[1],
            _ptr(inputs.reshape(-1), torch.int64, 'inputs'), _ptr(targets, torch.int64, 'targets'),
            _ptr(noise, torch.int64, 'noise'), batch, n_ctx, int(n_neg),
            _ptr(alias['prob'], torch.float32) if alias else None, _ptr(alias['alias'], torch.int32) if alias else None,
            float(lr), int(seed) & (2 ** 64 - 1), int(pair_id_base), int(flags), stats.data_ptr(), _stream()))
    _launches += 1
    return _stats_dict(stats) if own_stats else None


def check_ids(ids: torch.Tensor, lo: int, hi: int, what: str = 'ids') -> None:
    """Raise IndexError (as the reference's nn.Embedding does) if any id is outside [lo, hi); synchronises."""
    global _launches
    bad = torch.zeros(1, dtype=torch.int32, device=ids.device)
    with _on(ids):
        _check(load().se_check_ids(_ptr(ids, torch.int32, what), ids.numel(), int(lo), int(hi), bad.data_ptr(), _stream()))
    _launches += 1
    n_bad = int(bad.item())
    if n_bad:
        raise IndexError(f'{n_bad} of {ids.numel()} {what} are outside [{lo}, {hi})')


def sgns_update_walks(w_in, w_out, tokens: torch.Tensor, radius: int, n_neg: int,
                      row_offset: int, lr: float, seed: int, centre_id_base: int = 0,
                      alias: Optional[Dict[str, torch.Tensor]] = None, flags: int = SCATTER_RED,
                      stats: Optional[torch.Tensor] = None, local_negatives: bool = False,
                      check_tokens: bool = False) -> Optional[Dict[str, float]]:
    """The fused hot path on tokens int32 [n_seq, L]: windows + negatives + in-place SGNS update.
    w_in / w_out: CUDA float32 tensors, or ShardedTables striped over several GPUs (then `local_negatives` selects
    negatives among the rows this GPU owns; `alias`, if given, must be built over those local rows).
    `check_tokens`: validate token + row_offset against the table first (one extra small launch and a sync) -- for tokens
    that do not come from `walk` (which only emits valid node ids)."""
    global _launches
    n_seq, seq_len = tokens.shape
    p_in, vocab, emb, s_in, dev = _table(w_in, 'w_in')
    if check_tokens:
        check_ids(tokens, -row_offset, vocab - row_offset, 'tokens')
    p_out, vocab_o, emb_o, s_out, _ = _table(w_out, 'w_out')
    if (vocab, emb) != (vocab_o, emb_o):
        raise ValueError('w_in and w_out must have the same shape')
    spec = _same_sharding(s_in, s_out)
    if spec is not None:
        spec = ShardSpec(spec.world, spec.rank, spec.stripe_rows, int(bool(local_negatives)), 0)
    elif local_negatives:
        raise ValueError('local_negatives needs sharded tables')
    own_stats = stats is None
    if own_stats:
        stats = torch.zeros(STATS_LEN, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _check(load().se_sgns_update_walks_sharded(
            p_in, p_out, vocab, emb,
            _ptr(tokens, torch.int32, 'tokens'), n_seq, seq_len, int(radius), int(n_neg), int(row_offset),
            _ptr(alias['prob'], torch.float32) if alias else None, _ptr(alias['alias'], torch.int32) if alias else None,
            float(lr), int(seed) & (2 ** 64 - 1), int(centre_id_base), int(flags),
            ctypes.byref(spec) if spec is not None else None, stats.data_ptr(), _stream()))
    _launches += 1
    return _stats_dict(stats) if own_stats else None


_pairs_scratch: Dict[int, torch.Tensor] = {}


def sgns_update_negatives_owned(w_in, w_out, tokens: torch.Tensor, radius: int, n_neg: int, row_offset: int, lr: float, seed: int,
                                centre_id_base: int = 0, alias: Optional[Dict[str, torch.Tensor]] = None,
                                stats: Optional[torch.Tensor] = None, grouped: bool = True,
                                scratch: Optional[torch.Tensor] = None) -> None:
    """Owner-computes negatives on striped tables: processes, for every centre of `tokens` (any GPU's walks), the
    negatives whose rows this rank owns (global negative distribution, same Philox keys as `sgns_update_walks`).
    grouped (default): centres bucketed by table row first (`sgns_update_pairs_owned` without the positives); False: walk order
    (`se_sgns_update_negatives_owned`)."""
    global _launches
    if grouped:
        return sgns_update_pairs_owned(w_in, w_out, tokens, radius, n_neg, row_offset, lr, seed, centre_id_base=centre_id_base, alias=alias,
                                       stats=stats, positives=False, scratch=scratch)
    n_seq, seq_len = tokens.shape
    p_in, vocab, emb, s_in, dev = _table(w_in, 'w_in')
    p_out, _, _, s_out, _ = _table(w_out, 'w_out')
    spec = _same_sharding(s_in, s_out)
    if spec is None:
        raise ValueError('sgns_update_negatives_owned needs striped tables (ShardedTable)')
    spec = ShardSpec(spec.world, spec.rank, spec.stripe_rows, 0, 0)
    with torch.cuda.device(dev):
        _check(load().se_sgns_update_negatives_owned(
            p_in, p_out, vocab, emb, _ptr(tokens, torch.int32, 'tokens'), n_seq, seq_len, int(radius), int(n_neg), int(row_offset),
            _ptr(alias['prob'], torch.float32) if alias else None, _ptr(alias['alias'], torch.int32) if alias else None,
            float(lr), int(seed) & (2 ** 64 - 1), int(centre_id_base), ctypes.byref(spec),
            stats.data_ptr() if stats is not None else None, _stream()))
    _launches += 1


def sgns_update_pairs_owned(w_in, w_out, tokens: torch.Tensor, radius: int, n_neg: int, row_offset: int, lr: float, seed: int,
                            centre_id_base: int = 0, alias: Optional[Dict[str, torch.Tensor]] = None,
                            stats: Optional[torch.Tensor] = None, positives: bool = True,
                            scratch: Optional[torch.Tensor] = None) -> None:
    """Owner-computes on striped tables, centres bucketed by table row (`se_sgns_update_pairs_owned`): for every centre of `tokens`
    (the gathered walks of all GPUs) the negatives -- and with `positives` the context tokens -- whose W_out rows this rank owns.
    Called on every rank with the same tokens and keys, every pair of the batch is computed exactly once.  `scratch`: a uint8 device
    tensor of `pairs_owned_scratch_bytes` (default: one cached per device)."""
    global _launches
    n_seq, seq_len = tokens.shape
    p_in, vocab, emb, s_in, dev = _table(w_in, 'w_in')
    p_out, _, _, s_out, _ = _table(w_out, 'w_out')
    spec = _same_sharding(s_in, s_out)
    if spec is None:
        raise ValueError('sgns_update_pairs_owned needs striped tables (ShardedTable)')
    spec = ShardSpec(spec.world, spec.rank, spec.stripe_rows, 0, 0)
    with torch.cuda.device(dev):
        need = pairs_owned_scratch_bytes(vocab, n_seq, seq_len, radius)
        if scratch is None:
            key = torch.device(dev).index or 0
            scratch = _pairs_scratch.get(key)
            if scratch is None or scratch.numel() < need:
                scratch = _pairs_scratch[key] = torch.empty(need, dtype=torch.uint8, device=dev)
        assert scratch.is_cuda and scratch.dtype == torch.uint8 and scratch.numel() >= need, 'pairs_owned scratch too small'
        _check(load().se_sgns_update_pairs_owned(
            p_in, p_out, vocab, emb, _ptr(tokens, torch.int32, 'tokens'), n_seq, seq_len, int(radius), int(n_neg), int(bool(positives)),
            int(row_offset), _ptr(alias['prob'], torch.float32) if alias else None, _ptr(alias['alias'], torch.int32) if alias else None,
            float(lr), int(seed) & (2 ** 64 - 1), int(centre_id_base), ctypes.byref(spec), scratch.data_ptr(), scratch.numel(),
            stats.data_ptr() if stats is not None else None, _stream()))
    _launches += 9          # count, 5 scan kernels, coarse + fine fill, update


def pairs_owned_scratch_bytes(vocab: int, n_seq: int, seq_len: int, radius: int) -> int:
    n = int(load().se_pairs_owned_scratch_bytes(int(vocab), int(n_seq), int(seq_len), int(radius)))
    if n < 0:
        raise ValueError('pairs_owned_scratch_bytes: bad sizes')
    return n


def host_walk_sgns_step(csr, starts_host: torch.Tensor, walk_len: int, p: float, q: float, node2vec: bool, rule: int,
                        seed: int, walk_id_base: int, w_in, w_out, radius: int, n_neg: int,
                        row_offset: int, lr: float, scratch: Dict[str, torch.Tensor], stats_host: torch.Tensor,
                        alias: Optional[Dict[str, torch.Tensor]] = None, flags: int = SCATTER_RED,
                        walks_host: Optional[torch.Tensor] = None, local_negatives: bool = False) -> None:
    """HOST-buffer pipeline step (H2D starts -> walk -> fused SGNS -> D2H stats [+ walks]); synchronises."""
    global _launches
    assert not starts_host.is_cuda and starts_host.dtype == torch.int32 and starts_host.is_contiguous()
    assert not stats_host.is_cuda and stats_host.dtype == torch.float64 and stats_host.numel() >= STATS_LEN
    n = starts_host.numel()
    p_in, vocab, emb, s_in, dev = _table(w_in, 'w_in')
    p_out, vocab_o, emb_o, s_out, _ = _table(w_out, 'w_out')
    if (vocab, emb) != (vocab_o, emb_o):
        raise ValueError('w_in and w_out must have the same shape')
    spec = _same_sharding(s_in, s_out)
    if spec is not None:
        spec = ShardSpec(spec.world, spec.rank, spec.stripe_rows, int(bool(local_negatives)), 0)
    with torch.cuda.device(dev):
        _check(load().se_host_walk_sgns_step_sharded(
            _ptr(csr.rowptr, torch.int64), _ptr(csr.col_sorted, torch.int32), _ptr(csr.wcdf, torch.float32), csr.n_nodes,
            int(csr.symmetric), starts_host.data_ptr(), n, int(walk_len), float(p), float(q), int(bool(node2vec)),
            int(rule), int(seed) & (2 ** 64 - 1), int(walk_id_base), p_in, p_out,
            vocab, emb, int(radius), int(n_neg), int(row_offset),
            _ptr(alias['prob'], torch.float32) if alias else None, _ptr(alias['alias'], torch.int32) if alias else None,
            float(lr), int(flags), ctypes.byref(spec) if spec is not None else None,
            _ptr(scratch['starts'], torch.int32), _ptr(scratch['walks'], torch.int32),
            _ptr(scratch['stats'], torch.float64), walks_host.data_ptr() if walks_host is not None else None,
            stats_host.data_ptr(), _stream()))
    _launches += 2


def host_sgns_update_tokens(tokens_host: torch.Tensor, w_in, w_out, radius: int, n_neg: int, row_offset: int, lr: float, seed: int,
                            tokens_dev: torch.Tensor, stats_dev: torch.Tensor, stats_host: torch.Tensor, centre_id_base: int = 0,
                            alias: Optional[Dict[str, torch.Tensor]] = None, flags: int = SCATTER_RED,
                            local_negatives: bool = False) -> None:
    """HOST-buffer step on token-id sequences int32 [n_seq, L] (H2D tokens -> fused SGNS -> D2H stats); synchronises."""
    global _launches
    assert not tokens_host.is_cuda and tokens_host.dtype == torch.int32 and tokens_host.is_contiguous()
    assert not stats_host.is_cuda and stats_host.dtype == torch.float64 and stats_host.numel() >= STATS_LEN
    n_seq, seq_len = tokens_host.shape
    assert tokens_dev.numel() >= n_seq * seq_len
    p_in, vocab, emb, s_in, dev = _table(w_in, 'w_in')
    p_out, _, _, s_out, _ = _table(w_out, 'w_out')
    spec = _same_sharding(s_in, s_out)
    if spec is not None:
        spec = ShardSpec(spec.world, spec.rank, spec.stripe_rows, int(bool(local_negatives)), 0)
    with torch.cuda.device(dev):
        _check(load().se_host_sgns_update_tokens(
            tokens_host.data_ptr(), n_seq, seq_len, p_in, p_out, vocab, emb, int(radius), int(n_neg), int(row_offset),
            _ptr(alias['prob'], torch.float32) if alias else None, _ptr(alias['alias'], torch.int32) if alias else None,
            float(lr), int(seed) & (2 ** 64 - 1), int(centre_id_base), int(flags),
            ctypes.byref(spec) if spec is not None else None, _ptr(tokens_dev, torch.int32), _ptr(stats_dev, torch.float64),
            stats_host.data_ptr(), _stream()))
    _launches += 1


# ----------------------------------------------------------------------------------------------------------------
# table utilities (local torch tensors and sharded tables alike)
# ----------------------------------------------------------------------------------------------------------------
def table_fill_uniform(table, bound: float, seed: int) -> None:
    """Xavier-style uniform(-bound, bound) keyed by the global element index (model.py:26-27 semantics, Philox draws);
    a ShardedTable fills only the stripes its rank owns."""
    global _launches
    ptr, vocab, emb, spec, dev = _table(table)
    with torch.cuda.device(dev):
        _check(load().se_table_fill_uniform(ptr, vocab * emb, float(bound), int(seed) & (2 ** 64 - 1),
                                            spec.stripe_rows * emb if spec is not None else 0,
                                            spec.world if spec is not None else 1, spec.rank if spec is not None else 0, _stream()))
    _launches += 1


def table_gather_rows(table, rows: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    global _launches
    ptr, vocab, emb, _, dev = _table(table)
    rows = rows.reshape(-1)
    if out is None:
        out = torch.empty((rows.numel(), emb), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _check(load().se_table_gather_rows(ptr, emb, _ptr(rows, torch.int64, 'rows'), rows.numel(), _ptr(out, torch.float32, 'out'), _stream()))
    _launches += 1
    return out


def table_scatter_rows(table, rows: torch.Tensor, src: torch.Tensor) -> None:
    global _launches
    ptr, vocab, emb, _, dev = _table(table)
    rows = rows.reshape(-1)
    assert src.shape == (rows.numel(), emb)
    with torch.cuda.device(dev):
        _check(load().se_table_scatter_rows(ptr, emb, _ptr(rows, torch.int64, 'rows'), rows.numel(), _ptr(src, torch.float32, 'src'), _stream()))
    _launches += 1


def table_renorm_rows(table, rows: torch.Tensor, max_norm: float) -> None:
    """nn.Embedding(max_norm) look-up side effect: renormalise the (de-duplicated) rows whose norm exceeds max_norm, in place."""
    global _launches
    ptr, vocab, emb, _, dev = _table(table)
    rows = torch.unique(rows.reshape(-1).to(dev, torch.int64))
    with torch.cuda.device(dev):
        _check(load().se_table_renorm_rows(ptr, emb, _ptr(rows, torch.int64, 'rows'), rows.numel(), float(max_norm), _stream()))
    _launches += 1


def csr_build(src: torch.Tensor, dst: torch.Tensor, n_nodes: int, weights: Optional[torch.Tensor] = None, symmetrize: bool = True):
    """Device edge list -> (rowptr int64 [n+1], col int32 [nnz], w fp64 [nnz] | None, wcdf fp32 [nnz] | None, max_degree, skipped)
    with networkx's simple-graph semantics (csrc/ingest.cu); one synchronisation to learn nnz."""
    global _launches
    dev = src.device
    n_edges = src.numel()
    slots = n_edges * (2 if symmetrize else 1)
    lib = load()
    scratch = torch.empty(max(int(lib.se_csr_build_scratch_bytes(n_nodes, n_edges, int(symmetrize))), 16), dtype=torch.uint8, device=dev)
    rowptr = torch.empty(n_nodes + 1, dtype=torch.int64, device=dev)
    col = torch.empty(max(slots, 1), dtype=torch.int32, device=dev)
    w_out = torch.empty(max(slots, 1), dtype=torch.float64, device=dev) if weights is not None else None
    wcdf = torch.empty(max(slots, 1), dtype=torch.float32, device=dev) if weights is not None else None
    info = torch.zeros(3, dtype=torch.int64, device=dev)
    with _on(src):
        _check(lib.se_csr_build(_ptr(src, torch.int32, 'src'), _ptr(dst, torch.int32, 'dst'), _ptr(weights, torch.float64, 'weights'), n_edges,
                                int(n_nodes), int(symmetrize), scratch.data_ptr(), scratch.numel(), rowptr.data_ptr(), col.data_ptr(),
                                _ptr(w_out), _ptr(wcdf), info.data_ptr(), _stream()))
    _launches += 16
    nnz, max_degree, skipped = (int(x) for x in info.tolist())
    skipped &= 0xffffffff
    return rowptr, col[:nnz].clone(), (w_out[:nnz].clone() if w_out is not None else None), (wcdf[:nnz].clone() if wcdf is not None else None), max_degree, skipped


def replica_chunk(n_elems: int, world: int, rank: int):
    """[lo, hi) element range of the flat table whose master lives on `rank` (host arithmetic, no GPU)."""
    lo, hi = c_i64(), c_i64()
    _check(load().se_replica_chunk(int(n_elems), int(world), int(rank), ctypes.byref(lo), ctypes.byref(hi)))
    return lo.value, hi.value


def replica_sync(base_ptr: int, stride_elems: int, world: int, rank: int, n_elems: int, master: torch.Tensor, mode: int = 0,
                 beta: float = 1.0) -> None:
    """One rank's share of the fused reduce-scatter + all-gather over the working copies (csrc/replica.cu)."""
    global _launches
    with _on(master):
        _check(load().se_replica_sync(int(base_ptr), int(stride_elems), int(world), int(rank), int(n_elems),
                                      _ptr(master, torch.float32, 'master'), int(mode), float(beta), _stream()))
    _launches += 1


# ----------------------------------------------------------------------------------------------------------------
# tensor-core contraction (tcgen05): cosine similarity / closest pairs, shared-negatives scoring
# ----------------------------------------------------------------------------------------------------------------
def gemm_nt(a: torch.Tensor, b: torch.Tensor, scale_a: Optional[torch.Tensor] = None, scale_b: Optional[torch.Tensor] = None,
            out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[i, j] = scale_a[i] * scale_b[j] * <a_i, b_j> for fp32 row-major a [m, k], b [n, k] on the tensor cores (3xTF32)."""
    global _launches
    m, kd = a.shape
    n = b.shape[0]
    assert b.shape[1] == kd
    if out is None:
        out = torch.empty((m, n), dtype=torch.float32, device=a.device)
    with _on(a):
        _check(load().se_gemm_nt(_ptr(a, torch.float32, 'a'), _ptr(b, torch.float32, 'b'), m, n, kd, _ptr(scale_a, torch.float32, 'scale_a'),
                                 _ptr(scale_b, torch.float32, 'scale_b'), _ptr(out, torch.float32, 'out'), _stream()))
    _launches += 1
    return out


def cosine_similarity(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """pairwise_cosine_similarity(x, y) (utils/func.py:7-20) -> [m, n] on the device."""
    global _launches
    m, n, emb = x.shape[0], y.shape[0], x.shape[1]
    out = torch.empty((m, n), dtype=torch.float32, device=x.device)
    norms = torch.empty(m + n, dtype=torch.float32, device=x.device)
    with _on(x):
        _check(load().se_cosine_similarity(_ptr(x, torch.float32, 'x'), _ptr(y, torch.float32, 'y'), m, n, emb, norms.data_ptr(), out.data_ptr(), _stream()))
    _launches += 3
    return out


def topk_rows(x: torch.Tensor, k: int):
    """(indices int64 [rows, k], values [rows, k]): the k largest entries per row, descending."""
    global _launches
    rows, cols = x.shape
    idx = torch.empty((rows, k), dtype=torch.int64, device=x.device)
    val = torch.empty((rows, k), dtype=torch.float32, device=x.device)
    with _on(x):
        _check(load().se_topk_rows(_ptr(x, torch.float32, 'x'), rows, cols, int(k), idx.data_ptr(), val.data_ptr(), _stream()))
    _launches += 1
    return idx, val


def sgns_step_shared_negatives(w_in: torch.Tensor, w_out: torch.Tensor, inputs: torch.Tensor, shared: torch.Tensor, n_ctx: int, n_neg: int,
                               lr: float, stats: Optional[torch.Tensor] = None, scratch: Optional[torch.Tensor] = None) -> Optional[Dict[str, float]]:
    """Negative half of an SGNS step with ONE shared set of negative rows for the whole batch (three tcgen05 GEMMs); see the header."""
    global _launches
    batch, s = inputs.numel(), shared.numel()
    need = load().se_shared_negatives_scratch_floats(batch, s, w_in.shape[1])
    if scratch is None or scratch.numel() < need:
        scratch = torch.empty(max(int(need), 4), dtype=torch.float32, device=w_in.device)
    own_stats = stats is None
    if own_stats:
        stats = torch.zeros(STATS_LEN, dtype=torch.float64, device=w_in.device)
    with _on(w_in):
        _check(load().se_sgns_step_shared_negatives(
            _ptr(w_in, torch.float32, 'w_in'), _ptr(w_out, torch.float32, 'w_out'), w_in.shape[0], w_in.shape[1],
            _ptr(inputs.reshape(-1), torch.int64, 'inputs'), batch, _ptr(shared.reshape(-1), torch.int64, 'shared'), s, int(n_ctx), int(n_neg),
            float(lr), scratch.data_ptr(), scratch.numel(), stats.data_ptr(), _stream()))
    stats[5] += float(batch * n_ctx * n_neg)        # negatives the step stands for
    _launches += 11
    return _stats_dict(stats) if own_stats else None


def softmax_xent(logits: torch.Tensor, labels: torch.Tensor, bias: Optional[torch.Tensor], grad_scale: float, loss_sum: Optional[torch.Tensor] = None,
                 n_correct: Optional[torch.Tensor] = None) -> None:
    """In place: logits [n, C] (C = 1: binary) -> grad_scale * d loss / d logits; loss / correct counts accumulated."""
    global _launches
    n, c = logits.shape
    with _on(logits):
        _check(load().se_softmax_xent(_ptr(logits, torch.float32, 'logits'), _ptr(labels, torch.int32, 'labels'), _ptr(bias, torch.float32, 'bias'), n, c,
                                      float(grad_scale), _ptr(loss_sum, torch.float64, 'loss_sum'), _ptr(n_correct, torch.int32, 'n_correct'), _stream()))
    _launches += 1


def transpose(x: torch.Tensor) -> torch.Tensor:
    global _launches
    rows, cols = x.shape
    out = torch.empty((cols, rows), dtype=torch.float32, device=x.device)
    with _on(x):
        _check(load().se_transpose(_ptr(x, torch.float32, 'x'), rows, cols, out.data_ptr(), _stream()))
    _launches += 1
    return out


# ----------------------------------------------------------------------------------------------------------------
# link-prediction features
# ----------------------------------------------------------------------------------------------------------------
def _edge_op_code(op) -> int:
    if isinstance(op, str):
        name = op.lower()
        assert name in EDGE_OPS, f'Operator "{op}" is not supported. Available: {list(EDGE_OPS.keys())}'
        return EDGE_OPS[name]
    return int(op)


def edge_features(table: torch.Tensor, src_rows: torch.Tensor, dst_rows: torch.Tensor, op) -> torch.Tensor:
    """out[i] = op(table[src_rows[i]], table[dst_rows[i]]) -> float32 [n_edges, emb] on the device."""
    global _launches
    n = src_rows.numel()
    out = torch.empty((n, table.shape[1]), dtype=torch.float32, device=table.device)
    with _on(table):
        _check(load().se_edge_features(_ptr(table, torch.float32, 'table'), table.shape[0], table.shape[1],
                                       _ptr(src_rows.reshape(-1), torch.int64, 'src_rows'), _ptr(dst_rows.reshape(-1), torch.int64, 'dst_rows'),
                                       n, _edge_op_code(op), out.data_ptr(), _stream()))
    _launches += 1
    return out


def edge_op(lhs: torch.Tensor, rhs: torch.Tensor, op) -> torch.Tensor:
    global _launches
    assert lhs.shape == rhs.shape
    out = torch.empty_like(lhs)
    with _on(lhs):
        _check(load().se_edge_op(_ptr(lhs, torch.float32, 'lhs'), _ptr(rhs, torch.float32, 'rhs'), lhs.numel(), _edge_op_code(op),
                                 out.data_ptr(), _stream()))
    _launches += 1
    return out


def sample_negative_edges(csr, n: int, seed: int, sample_id_base: int = 0):
    """(src, dst) int32 node ids of n sampled non-edges; raises if some sample found no non-neighbour."""
    global _launches
    dev = csr.device
    src = torch.empty(n, dtype=torch.int32, device=dev)
    dst = torch.empty(n, dtype=torch.int32, device=dev)
    fail = torch.zeros(1, dtype=torch.int32, device=dev)
    with _on(src):
        _check(load().se_sample_negative_edges(_ptr(csr.rowptr, torch.int64), _ptr(csr.col_sorted, torch.int32), csr.n_nodes, int(n),
                                               int(seed) & (2 ** 64 - 1), int(sample_id_base), src.data_ptr(), dst.data_ptr(),
                                               fail.data_ptr(), _stream()))
    _launches += 1
    if int(fail.item()):
        raise RuntimeError(f'{int(fail.item())} negative-edge samples found no non-neighbour (graph too dense)')
    return src, dst
