"""Host-side partitioning for the multi-GPU path: one process per GPU.

Training data: rank r owns a contiguous range of the fg||bg concatenation (the order
compile_training_data builds, kmerLr_data.go:306-325), so that concatenating the shards in rank order
gives back the single-GPU matrix row for row.  Genomic scoring: regions (contigs) are spread over the
ranks greedily by length (longest first to the least loaded rank); no collective, results are gathered
on the host in region order.  Pure integer logic: no device work here.
"""
import numpy as np


def sample_range(n, rank, world):
    """[lo, hi) of the n samples owned by `rank`: sizes differ by at most one, lower ranks get the extra"""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world of %d" % (rank, world))
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sequences(buf, off, rank, world):
    """(buffer, offsets) -> the rank's contiguous slice, offsets rebased to 0"""
    off = np.asarray(off, dtype=np.int64)
    lo, hi = sample_range(len(off) - 1, rank, world)
    return buf[off[lo]:off[hi]], off[lo:hi + 1] - off[lo]


def shard_training_set(fg, bg, rank, world):
    """fg, bg = (buffer, offsets).  Returns (fg_part, bg_part, labels_part) of this rank's slice of fg||bg."""
    fb, fo = fg
    bb, bo = bg
    fo = np.asarray(fo, dtype=np.int64)
    bo = np.asarray(bo, dtype=np.int64)
    nfg, nbg = len(fo) - 1, len(bo) - 1
    lo, hi = sample_range(nfg + nbg, rank, world)
    f_lo, f_hi = min(lo, nfg), min(hi, nfg)
    b_lo, b_hi = max(lo - nfg, 0), max(hi - nfg, 0)
    fpart = (fb[fo[f_lo]:fo[f_hi]], fo[f_lo:f_hi + 1] - fo[f_lo])
    bpart = (bb[bo[b_lo]:bo[b_hi]], bo[b_lo:b_hi + 1] - bo[b_lo])
    labels = np.concatenate([np.ones(f_hi - f_lo, dtype=np.uint8), np.zeros(b_hi - b_lo, dtype=np.uint8)])
    return fpart, bpart, labels


def assign_regions(lengths, world):
    """greedy longest-first assignment of regions to ranks; returns owner[i] for every region"""
    lengths = np.asarray(lengths, dtype=np.int64)
    owner = np.zeros(len(lengths), dtype=np.int64)
    load = np.zeros(world, dtype=np.int64)
    for i in np.argsort(-lengths, kind="stable"):
        r = int(np.argmin(load))
        owner[i] = r
        load[r] += lengths[i]
    return owner


def bind_host_to_gpu(device):
    """Pin this process (and the pinned host buffers it allocates afterwards, by first touch) to the CPUs
    NVML reports as local to CUDA device `device`: with one process per GPU the host-to-device copies of
    kmerlr_extract then stay on the GPU's own socket.  Returns the CPU list, or None when NVML, the PCI id
    or the affinity call is not available (nothing is changed then)."""
    import os
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        pr = torch.cuda.get_device_properties(device)
        try:
            bus = "%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
            h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(device)
        ncpu = os.cpu_count() or 1
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in range(ncpu) if (mask[c // 64] >> (c % 64)) & 1 and c in allowed]
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None
