"""Multi-GPU parity check (not collected by pytest; needs N >= 2 B200s):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 tests/multi_gpu_check.py

Every rank extracts its contiguous shard with KMERLR_FLAG_SHARDED; the class numbering, the gradient
(int64 fixed-point all-reduce), the loss and a few proximal-gradient iterations must equal what ONE GPU
computes on the whole set -- the gradient bit for bit.  Rank 0 prints one line per check.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import kmerlr_b200 as K
    from kmerlr_b200 import api, shard, synth

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    K.init(local)
    ok = True
    for M, N, L, flags in [(1, 8, 500, dict(revcomp=True)), (1, 10, 200, dict(revcomp=True, binarize=True)), (3, 7, 120, dict())]:
        n_fg, n_bg = 3001, 2500
        fb, fo = synth.sequences(n_fg, L, 1, planted=True)
        bb, bo = synth.sequences(n_bg, L, 2)
        kc = K.NewKmerCounter(M, N, **flags)
        binz = flags.get("binarize", False)
        # one GPU, whole set (no communicator yet / not sharded)
        full = K.compile_training_data(None, kc, None, None, True, binz, (fb, fo), (bb, bo))
        rng = np.random.default_rng(17)
        theta = rng.normal(scale=0.01, size=full.m + 1)
        cw = (0.9, 1.2)
        lr = K.logisticRegression(theta, cw, 0.0)
        g_full, l_full = lr.Gradient(None, full), lr.Loss(full)
        est = K.KmerLrEstimator(Epsilon=0.0, EpsilonLoss=0.0, MaxIterations=4)
        est.Theta = np.zeros(full.m + 1); est.ClassWeights = np.array(cw)
        est.estimate_proximal(full, 1e-3)
        th_full = est.Theta.copy()
        # reduced matrix (what the leapfrog path iterates on): 40 columns, 300 iterations
        sel = np.unique(np.concatenate([[0], np.linspace(1, full.m, 40).astype(np.int64)]))
        rfull = K.select_data(full, sel)
        rfull.SetLabels(np.concatenate([np.ones(n_fg, dtype=np.uint8), np.zeros(n_bg, dtype=np.uint8)]))
        est = K.KmerLrEstimator(Epsilon=0.0, EpsilonLoss=0.0, MaxIterations=300)
        est.Theta = np.zeros(len(sel)); est.ClassWeights = np.array(cw)
        est.estimate_proximal(rfull, 1e-3)
        thr_full = est.Theta.copy()
        # long rows (the classes of the short k-mers): the solver runs on the sliced + column-major views
        sel_long = np.unique(np.concatenate([[0], np.arange(1, min(31, full.m)), np.linspace(1, full.m, 20).astype(np.int64)]))
        rlong = K.select_data(full, sel_long)
        rlong.SetLabels(np.concatenate([np.ones(n_fg, dtype=np.uint8), np.zeros(n_bg, dtype=np.uint8)]))
        est = K.KmerLrEstimator(Epsilon=0.0, EpsilonLoss=0.0, MaxIterations=200)
        est.Theta = np.zeros(len(sel_long)); est.ClassWeights = np.array(cw)
        est.estimate_proximal(rlong, 1e-3)
        thl_full, long_density = est.Theta.copy(), rlong.nnz / max(rlong.n, 1)
        rlong.free()
        est = K.KmerLrEstimator(Epsilon=0.0, EpsilonLoss=1e-9, MaxIterations=100000)
        est.Theta = np.zeros(len(sel)); est.ClassWeights = np.array(cw)
        it_full, _ = est.estimate_proximal(rfull, 1e-3)
        thc_full = est.Theta.copy()
        rfull.free()
        # a short leapfrog path with EpsilonLambda > 0 (the `ok` of Select then depends on L1Reg = lambda n: n must be
        # the row count of the WHOLE set on every rank, or the ranks leave the epoch loop at different epochs)
        full.SetLabels(np.concatenate([np.ones(n_fg, dtype=np.uint8), np.zeros(n_bg, dtype=np.uint8)]))
        est = K.KmerLrEstimator(EpsilonLoss=1e-7, EpsilonLambda=1e-9, MaxIterations=20000, MaxEpochs=6, tie=K.TIE_INDEX)
        est.estimate_loop(full, 4)
        path_full, act_full, thp_full = [p[:3] for p in est.path], est.active_idx.copy(), est.Theta.copy()
        k_full, c_full = full.Kmers()
        full.free()
        # sharded
        K.comm_init_torch()
        fpart, bpart, labels = shard.shard_training_set((fb, fo), (bb, bo), rank, world)
        mine = K.compile_training_data(None, kc, None, None, True, binz, fpart, bpart, sharded=True)
        k_s, c_s = mine.Kmers()
        g_s, l_s = lr.Gradient(None, mine), lr.Loss(mine)
        est = K.KmerLrEstimator(Epsilon=0.0, EpsilonLoss=0.0, MaxIterations=4)
        est.Theta = np.zeros(mine.m + 1); est.ClassWeights = np.array(cw)
        est.estimate_proximal(mine, 1e-3)
        rmine = K.select_data(mine, sel)
        rmine.SetLabels(labels)
        est2 = K.KmerLrEstimator(Epsilon=0.0, EpsilonLoss=0.0, MaxIterations=300)
        est2.Theta = np.zeros(len(sel)); est2.ClassWeights = np.array(cw)
        est2.estimate_proximal(rmine, 1e-3)
        est3 = K.KmerLrEstimator(Epsilon=0.0, EpsilonLoss=1e-9, MaxIterations=100000)
        est3.Theta = np.zeros(len(sel)); est3.ClassWeights = np.array(cw)
        it_s, _ = est3.estimate_proximal(rmine, 1e-3)
        rmine.free()
        rlmine = K.select_data(mine, sel_long)
        rlmine.SetLabels(labels)
        est5 = K.KmerLrEstimator(Epsilon=0.0, EpsilonLoss=0.0, MaxIterations=200)
        est5.Theta = np.zeros(len(sel_long)); est5.ClassWeights = np.array(cw)
        est5.estimate_proximal(rlmine, 1e-3)
        rlmine.free()
        est4 = K.KmerLrEstimator(EpsilonLoss=1e-7, EpsilonLambda=1e-9, MaxIterations=20000, MaxEpochs=6, tie=K.TIE_INDEX)
        est4.estimate_loop(mine, 4)
        # the ABI from another host thread (cgo moves goroutines between OS threads): the library re-binds its device
        import threading
        box = {}
        th = threading.Thread(target=lambda: box.update(l=lr.Loss(mine)))
        th.start(); th.join()
        checks = {
            "leapfrog path with EpsilonLambda > 0 (%d epochs)" % len(est4.path): [p[:3] for p in est4.path] == path_full
                and np.array_equal(est4.active_idx, act_full) and np.array_equal(est4.Theta, thp_full),
            "loss from a second host thread": box.get("l") == l_s,
            "reduced theta after 300 iterations": np.array_equal(est2.Theta, thr_full),
            "long-row reduced theta after 200 iterations (%.0f entries per row)" % long_density: np.array_equal(est5.Theta, thl_full),
            "reduced converged (%d vs %d iterations)" % (it_s, it_full): abs(it_s - it_full) <= 2 and np.allclose(est3.Theta, thc_full, rtol=1e-6, atol=1e-12),
            "classes": mine.m == len(k_full) and np.array_equal(k_s, k_full) and np.array_equal(c_s, c_full),
            "gradient bit-identical": np.array_equal(g_s, g_full),
            "loss": abs(l_s - l_full) <= 1e-14 * abs(l_full),
            "theta after 4 iterations": np.array_equal(est.Theta, th_full),
        }
        rows = torch.tensor([mine.n], device="cuda")
        dist.all_reduce(rows)
        checks["rows"] = int(rows.item()) == n_fg + n_bg
        mine.free()
        K.comm_destroy()
        flat = torch.tensor([int(all(checks.values()))], device="cuda")
        dist.all_reduce(flat, op=dist.ReduceOp.MIN)
        if rank == 0:
            print("k=%d..%d L=%d %s world=%d: %s" % (M, N, L, flags, world, checks), flush=True)
        ok = ok and bool(flat.item())
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("MULTI_GPU_CHECK", "PASS" if ok else "FAIL", flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
