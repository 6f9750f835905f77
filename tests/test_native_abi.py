"""CPU: the C-ABI library loads, exports every symbol include/se_b200.h declares, and validates arguments
(no compute call is made without a GPU)."""
import ctypes

import numpy as np
import pytest

from shallow_encoders import _native as nat


def test_library_exports_every_header_symbol():
    lib = nat.load()
    declared = nat.header_symbols()
    assert len(declared) >= 12
    for name in declared:
        assert hasattr(lib, name), name
        assert name in nat._SIGNATURES
    assert b'sm_100a' in lib.se_version()


def test_alias_table_host_reconstructs_distribution():
    rng = np.random.default_rng(0)
    counts = rng.integers(0, 1000, 5000).astype(np.float64)
    counts[:5] = [0, 1, 10 ** 6, 3, 0]
    for power in (0.75, 1.0, 0.0):
        prob = np.empty(len(counts), dtype=np.float32)
        alias = np.empty(len(counts), dtype=np.int32)
        rc = nat.load().se_alias_build_host(counts.ctypes.data, len(counts), power, prob.ctypes.data, alias.ctypes.data)
        assert rc == 0
        w = np.ones_like(counts) if power == 0.0 else np.where(counts > 0, counts ** power, 0.0)
        want = w / w.sum()
        got = np.zeros(len(counts))
        np.add.at(got, np.arange(len(counts)), prob.astype(np.float64))
        np.add.at(got, alias, 1.0 - prob.astype(np.float64))
        got /= len(counts)
        np.testing.assert_allclose(got, want, atol=2e-7)
        assert alias.min() >= 0 and alias.max() < len(counts)


def test_argument_validation_without_gpu():
    lib = nat.load()
    one = ctypes.c_void_p(16)   # never dereferenced: validation fails first
    # text shorter than 2r+1 (torch_dataset.py:298 assert)
    rc = lib.se_sgns_update_walks(one, one, 10, 4, one, 1, 4, 2, 1, 0, None, None, 0.1, 0, 0, 0, None, None)
    assert rc == -1 and b'Text is too short' in lib.se_last_error()
    # walk length must be >= 1 (random_walk_generator.py:21 assert)
    rc = lib.se_walk(one, one, None, 5, 1, one, 1, 0, 1.0, 1.0, 0, 0, 0, 0, 1, one, None, 0, None)
    assert rc == -1 and b'walk length' in lib.se_last_error()
    rc = lib.se_walk(one, one, None, 5, 1, one, 1, 5, 0.0, 1.0, 1, 0, 0, 0, 1, one, None, 0, None)
    assert rc == -1 and b'positive' in lib.se_last_error()
    rc = lib.se_sgns_grad(None, one, 10, 4, one, one, one, 1, 1, 1, None, None, None, None)
    assert rc == -1


def test_wrappers_refuse_cpu_tensors():
    import torch
    t = torch.zeros(4, 4)
    idx = torch.zeros((2, 1), dtype=torch.int64)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        nat.skipgram_scores(t, t, idx, idx, True)


def test_product_path_fails_loudly_when_the_native_library_is_missing(monkeypatch, tmp_path):
    """No CPU fallback anywhere: with the shared library absent (or exporting fewer symbols than include/se_b200.h declares) `load()` raises
    and names the build command; nothing routes around it."""
    from shallow_encoders import _native as nat
    monkeypatch.setattr(nat, '_lib', None)
    monkeypatch.setattr(nat, 'LIB_PATH', str(tmp_path / 'libse_b200.so'))
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        nat.load()
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        nat.version()
