"""Sample sharding across ranks.

Linear CorEx touches X only through X~ A^T and X~^T (X~ A^T), so rows of X shard across GPUs with
no data movement: each rank keeps its row block, and the only exchange per pass pair is a sum of
the (m*n + m) partial moments.  `Reducer` is that sum -- `torch.distributed.all_reduce` over NCCL
(NVLink/NVSwitch) on the GPU box, over gloo in the CPU tests.  Column statistics of `preprocess`
(sums, counts, squared deviations) go through the same object.
"""


def shard_rows(n_rows, rank, world):
    """Contiguous, balanced row block [lo, hi) of rank `rank` out of `world`."""
    base, extra = divmod(int(n_rows), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class Reducer(object):
    """In-place sum over the ranks of a process group; the identity for a single rank."""

    def __init__(self, comm=None):
        self.group, self.world, self.rank, self.backend = None, 1, 0, None
        if comm is None or comm is False:
            return
        import torch.distributed as dist
        if not dist.is_available() or not dist.is_initialized():
            raise RuntimeError("comm was given but torch.distributed is not initialised")
        self.group = None if comm is True else comm
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.backend = dist.get_backend(self.group)

    def sum_(self, tensor):
        """All-reduce(sum) `tensor` in place on the current stream (NCCL) or synchronously (gloo)."""
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(tensor, op=dist.ReduceOp.SUM, group=self.group)
        return tensor

    def max_scalar(self, value):
        """Maximum of a host scalar over the ranks (used to make per-rank decisions collective)."""
        if self.world == 1:
            return value
        import torch
        import torch.distributed as dist
        dev = torch.device("cuda", torch.cuda.current_device()) if self.backend == "nccl" else torch.device("cpu")
        t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return t.item()

    def min_scalar(self, value):
        """Minimum of a host scalar over the ranks."""
        return -self.max_scalar(-value)

    def sum_scalar(self, value):
        if self.world == 1:
            return value
        import torch
        dev = torch.device("cuda", torch.cuda.current_device()) if self.backend == "nccl" else torch.device("cpu")
        t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
        self.sum_(t)
        return t.item()
