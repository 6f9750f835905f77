"""numpy restatement of csrc/common.cuh's Philox4x32-10 keying, so tests can predict the in-kernel draws."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
STREAM_WALK, STREAM_NEG, STREAM_DRAW, STREAM_NEG_COIN = 0x10000000, 0x20000000, 0x30000000, 0x40000000
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = (np.asarray(x, dtype=np.uint64) & MASK for x in (c0, c1, c2, c3))
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & MASK, p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return c0.astype(np.uint32), c1.astype(np.uint32), c2.astype(np.uint32), c3.astype(np.uint32)


def philox(seed, ids, sub, stream):
    ids = np.asarray(ids, dtype=np.uint64)
    sub = np.broadcast_to(np.asarray(sub, dtype=np.uint64), ids.shape)
    return philox4x32_10(ids & MASK, ids >> np.uint64(32), sub, np.full(ids.shape, stream, dtype=np.uint64),
                         seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)


def u01(r):
    return (r >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)


def draw_row(r0, r1, vocab, prob=None, alias=None):
    j = ((r0.astype(np.uint64) * np.uint64(vocab)) >> np.uint64(32)).astype(np.int64)
    if prob is not None:
        take_alias = u01(r1) >= prob[j]
        j = np.where(take_alias, alias[j].astype(np.int64), j)
    return j


def negatives(seed, centre_ids, n_ctx, n_neg, vocab, prob=None, alias=None):
    """(len(centre_ids), n_ctx, n_neg) int64, keyed as csrc/common.cuh `neg_words`: negative k of context n of centre c
    uses word (n & 3) of Philox(seed; c, n >> 2, STREAM_NEG | k) (and STREAM_NEG_COIN | k for the alias coin)."""
    c = np.asarray(centre_ids, dtype=np.uint64)
    shape = (len(c), n_ctx, n_neg)
    ids = np.broadcast_to(c[:, None, None], shape)
    n = np.broadcast_to(np.arange(n_ctx)[None, :, None], shape)
    k = np.broadcast_to(np.arange(n_neg)[None, None, :], shape)

    def word(stream):
        w = philox4x32_10(ids & MASK, ids >> np.uint64(32), (n >> 2).astype(np.uint64), (stream | k).astype(np.uint64),
                          seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
        return np.choose(n & 3, w)

    r0 = word(STREAM_NEG)
    r1 = word(STREAM_NEG_COIN) if prob is not None else r0
    return draw_row(r0, r1, vocab, prob, alias)


def draws(seed, n, vocab, prob=None, alias=None, base=0):
    ids = np.arange(base, base + n, dtype=np.uint64)
    r0, r1, _, _ = philox(seed, ids, 0, STREAM_DRAW)
    return draw_row(r0, r1, vocab, prob, alias)
