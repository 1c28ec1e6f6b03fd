"""Multi-GPU plumbing for independent sliding windows (BASELINE config 5).

A window is a closed problem, so the path shards with NO data-path collective: window i goes to
rank i mod G.  torch.distributed (NCCL on GPUs, gloo in the CPU tests) carries only the start
barrier, the max-over-ranks of the device time and the gather of per-window summaries."""
import torch
import torch.distributed as dist


def shard_windows(n_windows, world, rank):
    """Indices of the windows owned by `rank` (round-robin, as DESIGN.md 6)."""
    return list(range(rank, n_windows, world))


def max_over_ranks(value, device="cpu"):
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_summaries(local, n_windows):
    """local: {window_index: summary-dict}.  Returns the list of all summaries on every rank."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        merged = dict(local)
    else:
        parts = [None] * dist.get_world_size()
        dist.all_gather_object(parts, local)
        merged = {}
        for p in parts:
            for k, v in p.items():
                assert k not in merged, "window assigned to two ranks"
                merged[k] = v
    assert sorted(merged) == list(range(n_windows)), "windows missing from the gather"
    return [merged[i] for i in range(n_windows)]
