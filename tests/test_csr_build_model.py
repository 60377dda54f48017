"""CPU model of csrc/csr_build.cu: the block scan, the per-tile digit histogram (digit-major), and the
stable in-tile ranking (peer masks per 32 consecutive items, warp-private digit counters advanced by
the group leader, exclusive prefix over the 8 warps) restated thread for thread in numpy and compared
with the oracle's host builder (numpy stable argsort).  It pins the ALGORITHM the kernels implement --
the kernels themselves are checked bit for bit on the GPU (tests/test_gpu_csr_build.py)."""
import numpy as np
import pytest

import glnn_oracle as O

T, ITEMS, SCAN_ITEMS = 256, 8, 16
TILE, SCAN_TILE = T * ITEMS, T * SCAN_ITEMS


def block_excl_scan(v):
    """warp_incl_scan by shuffles + warp totals through shared memory, as block_excl_scan() does."""
    inc = v.copy()
    for w in range(8):
        seg = inc[w * 32:(w + 1) * 32]
        o = 1
        while o < 32:
            up = np.concatenate([np.zeros(o, dtype=seg.dtype), seg[:-o]])
            seg = seg + np.where(np.arange(32) >= o, up, 0)
            o <<= 1
        inc[w * 32:(w + 1) * 32] = seg
    s_warp = inc[31::32]
    before = np.repeat(np.concatenate([[0], np.cumsum(s_warp)[:-1]]), 32)
    return before + inc - v, int(s_warp.sum())


def exclusive_scan(x):
    n = len(x)
    nb = (n + SCAN_TILE - 1) // SCAN_TILE
    pad = np.zeros(nb * SCAN_TILE, dtype=np.int64)
    pad[:n] = x
    bsum = np.zeros(nb, dtype=np.int64)
    for b in range(nb):                      # scan_reduce_kernel: item i of thread t = base + i*T + t
        blk = pad[b * SCAN_TILE:(b + 1) * SCAN_TILE].reshape(SCAN_ITEMS, T)
        bsum[b] = block_excl_scan(blk.sum(0))[1]
    carry = 0                                # scan_bsums_kernel
    for base in range(0, nb, T):
        v = np.zeros(T, dtype=np.int64)
        m = min(T, nb - base)
        v[:m] = bsum[base:base + m]
        ex, tot = block_excl_scan(v)
        bsum[base:base + m] = carry + ex[:m]
        carry += tot
    out = np.zeros(nb * SCAN_TILE, dtype=np.int64)
    for b in range(nb):                      # scan_apply_kernel: thread t owns 16 consecutive items
        blk = pad[b * SCAN_TILE:(b + 1) * SCAN_TILE].reshape(T, SCAN_ITEMS)
        ex, _ = block_excl_scan(blk.sum(1))
        run = bsum[b] + ex
        for i in range(SCAN_ITEMS):
            out[b * SCAN_TILE + np.arange(T) * SCAN_ITEMS + i] = run
            run = run + blk[:, i]
    return out[:n]


def peer_masks(d, valid):
    """Eight ballots refine the mask of lanes with the same digit (what replaces MATCH.ANY)."""
    vm = sum(1 << l for l in range(32) if valid[l])
    out = []
    for l in range(32):
        grp = vm if valid[l] else ~vm & 0xffffffff
        for b in range(8):
            m = sum(1 << x for x in range(32) if (d[x] >> b) & 1)
            grp &= m if (d[l] >> b) & 1 else ~m & 0xffffffff
        out.append(grp)
    return out


def radix_pass(key, val, shift):
    e = len(key)
    nb = (e + TILE - 1) // TILE
    hist = np.zeros(256 * nb, dtype=np.int64)
    for b in range(nb):                      # radix_hist_kernel, digit-major
        dig = (key[b * TILE:(b + 1) * TILE] >> shift) & 255
        hist[np.arange(256) * nb + b] = np.bincount(dig, minlength=256)
    goff = exclusive_scan(hist)
    out_k, out_v = np.zeros(e, dtype=np.int64), np.zeros(e, dtype=np.int64)
    for b in range(nb):                      # radix_scatter_kernel
        cnt = np.zeros((8, 256), dtype=np.int64)
        rank = np.zeros((8, ITEMS, 32), dtype=np.int64)
        for w in range(8):
            base = b * TILE + w * 32 * ITEMS
            for i in range(ITEMS):
                j = base + i * 32 + np.arange(32)
                valid = j < e
                d = [int((key[x] >> shift) & 255) if ok else 0xffffffff for x, ok in zip(j, valid)]
                grp = peer_masks(d, valid)
                old = [0] * 32
                for l in range(32):
                    if valid[l] and grp[l] & ((1 << l) - 1) == 0:       # group leader
                        old[l] = cnt[w][d[l]]
                        cnt[w][d[l]] += bin(grp[l]).count("1")
                for l in range(32):
                    leader = (grp[l] & -grp[l]).bit_length() - 1
                    rank[w, i, l] = old[leader] + bin(grp[l] & ((1 << l) - 1)).count("1")
        for t in range(256):                 # thread t = digit t
            run = goff[t * nb + b]
            for w in range(8):
                c = cnt[w][t]
                cnt[w][t] = run
                run += c
        for w in range(8):
            base = b * TILE + w * 32 * ITEMS
            for i in range(ITEMS):
                for l in range(32):
                    j = base + i * 32 + l
                    if j < e:
                        pos = cnt[w][(key[j] >> shift) & 255] + rank[w, i, l]
                        out_k[pos], out_v[pos] = key[j], val[j]
    return out_k, out_v


def passes(n):
    bits = 0
    while bits < 31 and (1 << bits) < n:
        bits += 1
    return (bits + 7) // 8


def test_scan_model_equals_cumsum():
    rng = np.random.default_rng(0)
    for n in (1, 255, 4096, 4097, 3 * SCAN_TILE + 77):
        x = rng.integers(0, 7, n)
        assert np.array_equal(exclusive_scan(x), np.concatenate([[0], np.cumsum(x)[:-1]]))


@pytest.mark.parametrize("n,e", [(1, 10), (2, 50), (256, 3000), (257, 5000), (300, 2048), (300, 2049),
                                 (70000, 6000)])
def test_radix_model_equals_stable_sort_by_destination(n, e):
    rng = np.random.default_rng(n + e)
    src = rng.integers(0, n, e)
    dst = np.minimum((n * rng.random(e) ** 2).astype(np.int64), n - 1)
    key, val = dst.copy(), src.copy()
    for p in range(passes(n)):
        key, val = radix_pass(key, val, 8 * p)
    want_ptr, want_idx = O.csr_from_edges(src, dst, n)
    assert np.array_equal(val, want_idx)
    cnt = np.concatenate([np.bincount(dst, minlength=n), [0]])
    assert np.array_equal(exclusive_scan(cnt), want_ptr)
