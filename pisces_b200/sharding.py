"""Interval sharding of the per-locus path across the GPUs of one box, and the single gather of call records at the end.

The reference already shards by chromosome (src/lib/Pisces.Processing/Logic/BaseGenomeProcessor.cs:60-72) and concatenates per-chromosome
VCFs in genome order (src/exe/Pisces/Logic/Processing/GenomeProcessor.cs:156-186). Loci are independent once their counts are complete, so a
chromosome can be cut further into contiguous runs of 1000-bp blocks (GlobalConstants.RegionSize), one run per rank: no collective on the
data path, one all-gather of fixed-size call records when every rank is done. Works with any torch.distributed backend:
NCCL on the GPUs (bench.py), gloo on CPU (tests/test_sharding_gloo.py).
"""
import numpy as np
import torch
import torch.distributed as dist

from . import _native

BLOCK = 1000  # GlobalConstants.RegionSize
RECORD_BYTES = 96


def shard_plan(pos0, first_position, last_position, max_read_span, n_shards):
    """pb2_shard_plan (the library owns the plan, a .NET host calls the same entry point): block-aligned cuts balanced by the position-sorted reads' starts
    (pos0 may be None: balanced by positions), halo of two blocks + max_read_span. Returns a list of dicts own_lo, own_hi, stage_lo, stage_hi, read_first,
    read_end."""
    import ctypes as C
    L = _native.load()
    out = (_native.Shard * n_shards)()
    p0 = None if pos0 is None else np.ascontiguousarray(pos0, dtype=np.int32)
    rc = L.pb2_shard_plan(None if p0 is None else p0.ctypes.data, 0 if p0 is None else len(p0), int(first_position), int(last_position), int(max_read_span), int(n_shards), out)
    if rc != 0:
        raise ValueError(f"pb2_shard_plan: bad argument ({rc})")
    return [dict(own_lo=s.own_lo, own_hi=s.own_hi, stage_lo=s.stage_lo, stage_hi=s.stage_hi, read_first=s.read_first, read_end=s.read_end) for s in out]


def shard_reads(d, shard):
    """The reads of one shard (views of the struct of arrays `d`, offsets left absolute: pb2_push_reads accepts windows of larger arrays)."""
    a, b = int(shard["read_first"]), int(shard["read_end"])
    out = dict(pos0=d["pos0"][a:b], flag=d["flag"][a:b], cigar_off=d["cigar_off"][a:b + 1], cigar=d["cigar"], seq_off=d["seq_off"][a:b + 1], bases=d["bases"], quals=d["quals"])
    if d.get("collapsed") is not None:
        out["collapsed"] = d["collapsed"][a:b]
    if d.get("base_dirs") is not None:
        out["base_dirs"] = d["base_dirs"]
    return out


def shard_loci(weights, world_size, block=BLOCK, first_position=1):
    """Cut loci [0, n) into world_size contiguous shards balanced by `weights` (entries per locus), cutting only where a new 1000-bp block of
    reference positions starts (so that a shard owns whole RegionState blocks). Returns [(lo, hi)] with hi exclusive; shards may be empty."""
    w = np.asarray(weights, dtype=np.float64)
    n = len(w)
    if n == 0:
        return [(0, 0)] * world_size
    cum = np.concatenate([[0.0], np.cumsum(w)])
    # candidate cut points: locus indices whose position is the first of a block
    first_block_start = (-(first_position - 1)) % block
    cuts = np.arange(first_block_start, n, block, dtype=np.int64)
    cuts = cuts[cuts > 0]
    bounds = [0]
    for r in range(1, world_size):
        target = cum[-1] * r / world_size
        if len(cuts) == 0:
            bounds.append(bounds[-1])
            continue
        k = int(np.argmin(np.abs(cum[cuts] - target)))
        bounds.append(max(int(cuts[k]), bounds[-1]))
    bounds.append(n)
    return [(bounds[i], bounds[i + 1]) for i in range(world_size)]


def gather_call_records(local_records, group=None, device=None):
    """All ranks contribute their (position-sorted) call records and receive every rank's, concatenated in rank order — which is genome
    order for interval shards. `local_records`: numpy structured array (dtype _native.RECORD_DTYPE) or a uint8 torch tensor of n*96 bytes.
    Two collectives: the counts, then one all_gather of the payload padded to the largest count."""
    world = dist.get_world_size(group)
    if isinstance(local_records, np.ndarray):
        raw = torch.from_numpy(np.ascontiguousarray(local_records).view(np.uint8).reshape(-1).copy())
    else:
        raw = local_records.reshape(-1)
    if device is None:
        device = raw.device
    raw = raw.to(device)
    n_local = raw.numel() // RECORD_BYTES
    counts = torch.zeros(world, dtype=torch.int64, device=device)
    mine = torch.tensor([n_local], dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(counts, mine, group=group)
    cap = int(counts.max().item())
    padded = torch.zeros(max(cap, 1) * RECORD_BYTES, dtype=torch.uint8, device=device)
    padded[: raw.numel()] = raw
    out = torch.empty(world * max(cap, 1) * RECORD_BYTES, dtype=torch.uint8, device=device)
    dist.all_gather_into_tensor(out, padded, group=group)
    parts = []
    for r in range(world):
        k = int(counts[r].item())
        parts.append(out[r * max(cap, 1) * RECORD_BYTES: r * max(cap, 1) * RECORD_BYTES + k * RECORD_BYTES])
    merged = torch.cat(parts) if parts else out[:0]
    return np.frombuffer(merged.cpu().numpy().tobytes(), dtype=_native.RECORD_DTYPE)
