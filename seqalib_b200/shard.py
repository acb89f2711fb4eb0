"""Multi-GPU plumbing for one-process-per-GPU runs (torchrun): the DP path has no exchange step -- pairs are
independent -- so the only distributed operations are the static shard assignment and the reductions a benchmark
or a driver needs (max of per-rank times, sum of per-rank cell counts).  Works with the nccl backend (CUDA
tensors) and, for the CPU test-suite, with gloo."""
import torch
import torch.distributed as dist


def shard_range(n_total, rank, world):
    """Contiguous static split of n_total pairs: rank r owns [lo, hi)."""
    lo = n_total * rank // world
    hi = n_total * (rank + 1) // world
    return lo, hi


def weak_shard(pairs_per_rank, rank):
    """Weak scaling: every rank owns pairs_per_rank pairs of the global synthetic stream; -> first pair id."""
    return pairs_per_rank * rank


def _device():
    if dist.is_initialized() and dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def reduce_max(x):
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(x)
    t = torch.tensor([float(x)], dtype=torch.float64, device=_device())
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def reduce_sum(x):
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(x)
    t = torch.tensor([float(x)], dtype=torch.float64, device=_device())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def barrier():
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
