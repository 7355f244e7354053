"""Per-GPU batch sharding (one process per GPU, torch.distributed for the plumbing).

Curves are independent, so the fit itself needs NO collective: every rank fits the contiguous
slice `shard_range(nbatch, rank, world)` of the batch with its own Engine.  The only optional
exchange is the final gather of the per-curve summary (n, fp, ier: 16 bytes per curve), which
runs over NCCL/NVLink on GPUs and over gloo on CPU tensors (used by the CPU tests).
"""
import torch
import torch.distributed as dist


def shard_range(nbatch, rank, world, align=128):
    """Contiguous slice [b0, b1) of rank `rank`; slice starts are multiples of `align` curves
    (a CTA of the fit kernels handles 128 curves) and the slices cover [0, nbatch) exactly."""
    if world <= 1:
        return 0, nbatch
    per = -(-nbatch // world)
    per = -(-per // align) * align
    b0 = min(nbatch, per * rank)
    b1 = min(nbatch, per * (rank + 1))
    return b0, b1


def gather_fit_summary(n, fp, ier, nbatch=None, group=None):
    """All-gather the per-curve summary of every rank's slice.

    n, ier: int32 [nb_local]; fp: float64 [nb_local] (device tensors for NCCL, CPU tensors for
    gloo).  Slices may have different lengths (last ranks can be short or empty); returns
    (n, fp, ier) of the whole batch in curve order on every rank.
    """
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return n, fp, ier
    world = dist.get_world_size(group)
    dev = n.device
    nb_local = torch.tensor([n.numel()], device=dev, dtype=torch.int64)
    sizes = [torch.zeros(1, device=dev, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, nb_local, group=group)
    sizes = [int(s.item()) for s in sizes]
    cap = max(max(sizes), 1)
    # pack (n, ier) as two int32 in one int64 word next to fp: one collective, 16 B per curve
    packed = torch.zeros((cap, 2), device=dev, dtype=torch.float64)
    if n.numel():
        ni = torch.stack([n.to(torch.int32), ier.to(torch.int32)], dim=1).contiguous()
        packed[:n.numel(), 0] = ni.view(torch.int64).view(-1).view(torch.float64)
        packed[:n.numel(), 1] = fp
    out = [torch.empty_like(packed) for _ in range(world)]
    dist.all_gather(out, packed, group=group)
    ns, fps, iers = [], [], []
    for r in range(world):
        k = sizes[r]
        if k == 0:
            continue
        ni = out[r][:k, 0].contiguous().view(torch.int64).view(torch.int32).view(k, 2)
        ns.append(ni[:, 0])
        iers.append(ni[:, 1])
        fps.append(out[r][:k, 1])
    n_all, fp_all, ier_all = torch.cat(ns), torch.cat(fps), torch.cat(iers)
    if nbatch is not None:
        assert n_all.numel() == nbatch
    return n_all, fp_all, ier_all
