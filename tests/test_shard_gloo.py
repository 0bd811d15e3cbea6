"""The N > 1 path on the CPU: world_size 2 over gloo.

What the library does with NCCL on the GPUs is restated here with the oracle as the per-rank engine and
gloo as the transport: contiguous sample shards (kmerlr_b200/shard.py), a MAX all-reduce of the
observed-class bitmap so that every rank derives the same column numbering, an int64 SUM all-reduce of
the fixed-point gradient (bit-identical to the single-rank sum, in any order), and region sharding with
no collective for genomic scoring.  Host logic only: no GPU, no CUDA library calls.
"""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _dense_ids(cfg, k, code):
    """dense class id = level offset (multiples of 32, as in extract.cu) + code"""
    off, d = {}, 0
    for kk in range(cfg.M, cfg.N + 1):
        off[kk] = d
        d += ((1 << (2 * kk)) + 31) // 32 * 32
    return np.array([off[int(a)] + int(b) for a, b in zip(k, code)], dtype=np.int64), d, off


def _fixed_point_gradient(O, mat, labels, theta, n_global, scale):
    """sum_i round(w_i v_ij S) per column as int64 (logistic.cu: fused kernels), bias in slot 0"""
    r = O.log_pdf(mat, theta)
    w = np.where(labels.astype(bool), (np.exp(r) - 1.0), np.exp(r)) / n_global
    rp, col, val = mat.rows()
    G = np.zeros(mat.m + 1, dtype=np.int64)
    G[0] = np.rint(w * scale).astype(np.int64).sum()
    rows = np.repeat(np.arange(mat.n), np.diff(rp))
    np.add.at(G, col + 1, np.rint(w[rows] * scale * val).astype(np.int64))
    return G


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from kmerlr_b200 import shard, synth
    from oracle import oracle as O

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n_fg, n_bg, L = 37, 30, 120
        fb, fo = synth.sequences(n_fg, L, 1, planted=True)
        bb, bo = synth.sequences(n_bg, L, 2)
        cfg = O.make_config(1, 6, revcomp=True)
        fpart, bpart, labels = shard.shard_training_set((fb, fo), (bb, bo), rank, world)
        buf = np.concatenate([fpart[0], bpart[0]])
        off = np.concatenate([fpart[1], bpart[1][1:] + fpart[1][-1]])
        # 1. local extraction, bitmap of observed classes, MAX all-reduce = OR
        local = O.extract(cfg, (buf, off))
        ids, nbits, lvl = _dense_ids(cfg, *local.classes())
        bm = torch.zeros(nbits, dtype=torch.uint8)
        bm[torch.from_numpy(ids)] = 1
        dist.all_reduce(bm, op=dist.ReduceOp.MAX)
        gids = np.nonzero(bm.numpy())[0]
        gk = np.zeros(len(gids), dtype=np.int32)
        for kk, o in lvl.items():
            gk[gids >= o] = kk
        gcode = (gids - np.array([lvl[int(k)] for k in gk])).astype(np.uint64)
        mine = O.extract(cfg, (buf, off), frozen=(gk, gcode))
        # 2. fixed-point gradient, int64 SUM all-reduce
        n_global = n_fg + n_bg
        theta = np.linspace(-0.02, 0.02, mine.m + 1)
        scale = float(2 ** 50)
        G = torch.from_numpy(_fixed_point_gradient(O, mine, labels, theta, n_global, scale))
        dist.all_reduce(G, op=dist.ReduceOp.SUM)
        lsum = torch.tensor([float(np.sum(-np.where(labels.astype(bool), O.log_pdf(mine, theta),
                                                      np.log1p(-np.exp(O.log_pdf(mine, theta))))))], dtype=torch.float64)
        parts = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(parts, lsum)
        loss = sum(float(p) for p in parts) / n_global          # rank order: identical on every rank
        # 3. region sharding for window scoring, no collective
        regions = [synth.random_bases(ln, 3, offset=1000 * i).tobytes().decode() for i, ln in enumerate([400, 90, 260, 333, 150])]
        owner = shard.assign_regions([len(r) for r in regions], world)
        feats = [(i, i) for i in range(mine.m)]
        md = dict(cfg=cfg, class_k=gk, class_code=gcode, features=feats, theta=theta)
        scored = {i: O.score_windows([md], [regions[i]], 100, 10) for i in range(len(regions)) if owner[i] == rank}
        np.savez(os.path.join(out_dir, "rank%d.npz" % rank), rp=mine.rows()[0], col=mine.rows()[1], val=mine.rows()[2],
                 gk=gk, gcode=gcode, G=G.numpy(), loss=loss, owner=owner,
                 **{"score%d" % i: v for i, v in scored.items()})
    finally:
        dist.destroy_process_group()


def test_partition_helpers():
    sys.path.insert(0, ROOT)
    from kmerlr_b200 import shard
    for n in (0, 1, 7, 64, 1001):
        for world in (1, 2, 3, 8):
            r = [shard.sample_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            sizes = [hi - lo for lo, hi in r]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
    with pytest.raises(ValueError):
        shard.sample_range(10, 2, 2)
    owner = shard.assign_regions([5, 100, 7, 50, 49, 1], 2)
    assert owner.tolist() == [0, 0, 1, 1, 1, 0]          # 100+5+1 | 50+49+7: longest first to the least loaded
    load = np.bincount(shard.assign_regions([10] * 24, 8), minlength=8)
    assert load.tolist() == [3] * 8


def test_world2_gloo_matches_single_rank(tmp_path):
    import torch.multiprocessing as mp
    sys.path.insert(0, ROOT)
    from kmerlr_b200 import synth
    from oracle import oracle as O

    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    parts = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % r)) for r in range(world)]
    # single-rank reference
    n_fg, n_bg, L = 37, 30, 120
    fb, fo = synth.sequences(n_fg, L, 1, planted=True)
    bb, bo = synth.sequences(n_bg, L, 2)
    buf = np.concatenate([fb, bb])
    off = np.concatenate([fo, bo[1:] + fo[-1]])
    labels = np.concatenate([np.ones(n_fg, dtype=np.uint8), np.zeros(n_bg, dtype=np.uint8)])
    cfg = O.make_config(1, 6, revcomp=True)
    full = O.extract(cfg, (buf, off))
    k, code = full.classes()
    for p in parts:                                     # same column numbering on every rank
        assert np.array_equal(p["gk"], k) and np.array_equal(p["gcode"], code)
    rp, col, val = full.rows()
    cat_col = np.concatenate([p["col"] for p in parts])
    cat_val = np.concatenate([p["val"] for p in parts])
    cat_rp = np.concatenate([parts[0]["rp"], parts[1]["rp"][1:] + parts[0]["rp"][-1]])
    assert np.array_equal(cat_rp, rp) and np.array_equal(cat_col, col) and np.array_equal(cat_val, val)
    theta = np.linspace(-0.02, 0.02, full.m + 1)
    G1 = _fixed_point_gradient(O, full, labels, theta, n_fg + n_bg, float(2 ** 50))
    for p in parts:
        assert np.array_equal(p["G"], G1)               # integer sums: bit identical for any world size
    g = O.gradient(full, labels, theta)
    assert np.allclose(parts[0]["G"] / float(2 ** 50), g, rtol=1e-9, atol=1e-12)
    assert parts[0]["loss"] == parts[1]["loss"]
    assert abs(parts[0]["loss"] - O.loss(full, labels, theta)) <= 1e-12 * abs(parts[0]["loss"])
    # window scoring: every region scored by exactly one rank, results equal the unsharded call
    regions = [synth.random_bases(ln, 3, offset=1000 * i).tobytes().decode() for i, ln in enumerate([400, 90, 260, 333, 150])]
    feats = [(i, i) for i in range(full.m)]
    md = dict(cfg=cfg, class_k=k, class_code=code, features=feats, theta=theta)
    owner = parts[0]["owner"]
    assert np.array_equal(owner, parts[1]["owner"])
    for i, reg in enumerate(regions):
        ref = O.score_windows([md], [reg], 100, 10)
        got = parts[int(owner[i])]["score%d" % i]
        assert "score%d" % i not in parts[1 - int(owner[i])].files
        assert np.array_equal(got, ref)


def test_bind_host_to_gpu_is_harmless_without_a_gpu():
    """no GPU / no NVML here: the helper must change nothing and say so"""
    import os
    from kmerlr_b200 import shard
    before = os.sched_getaffinity(0)
    cpus = shard.bind_host_to_gpu(0)
    assert cpus is None or set(cpus) <= before
    if cpus is None:
        assert os.sched_getaffinity(0) == before
