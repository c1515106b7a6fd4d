"""Multi-GPU execution of the three-point estimators: one process per GPU.

The reference has no distributed mode (``trv::sys::currTask`` is the constant 0,
I/monitor.hpp:249-250).  Here every rank makes the same estimator call with
``part_rank``/``part_count`` set.  Deterministic mode, survey catalogues, the 3PCF and
the CPU tests: the mesh is replicated and the independent entries of the data vector --
the (k1, k2) bin pairs of every (m1, m2, M) term, S/threept.cpp:1902-1904, 2137-2139 --
are dealt to the ranks; each entry is produced by exactly one rank (zeros elsewhere), so
a single small all-reduce(sum) of ``4 * dv_dim`` doubles completes the result and is
bit-identical to the single-GPU one.  Throughput mode on a box: the ranks split the
x-planes of the sub-grid (and, with the NCCL communicator attached, of the mesh itself:
DESIGN.md section 6) and every entry is a sum over the ranks.

The exchange itself lives behind the C API (``trv_comm_init`` / ``trvb_allreduce``,
NCCL bound inside ``libtrvb.so``): once :func:`init_comm` has attached the
communicator, ``core.threept(..., part_rank=r, part_count=R)`` returns the complete
measurement on every rank with no Python in the data path.  ``torch.distributed``
only ships the 128-byte NCCL id at start-up (any backend) and serves the CPU tests
(``gloo``), where :func:`allreduce_result` sums the partial vectors instead.
"""
import numpy as np


def init_comm(group=None):
    """Attach an NCCL communicator spanning the ranks of ``group`` to this process's
    estimator calls (collective).  Needs an initialised ``torch.distributed`` group of
    any backend to ship the id; returns the communicator size."""
    import torch
    import torch.distributed as dist
    from . import core
    if not (dist.is_available() and dist.is_initialized()):
        return 1
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if world == 1:
        return 1
    ident = [core.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ident, src=0, group=group)
    core.comm_init(world, rank, ident[0])
    return core.comm_size()

_STAT_KEYS = {"bispec": ("bk_raw", "bk_shot"), "3pcf": ("zeta_raw", "zeta_shot")}


def owners(form, degrees, num_bins, world_size, idx_bin=0):
    """Owner rank of every data-vector entry: the rule of ``active_entries`` in
    src/threept.cpp (``trv::partition_owners``) -- compact blocks of the bin-pair
    matrix with equal numbers of distinct shell fields (not of entries): the shell
    transforms are what a rank's share costs."""
    from . import core
    return core.partition_owners(form, degrees, num_bins, world_size, idx_bin=idx_bin)


def shot_noise_owners(form, degrees, num_bins, world_size, idx_bin=0, distributed_mesh=False):
    """Rank that computes the SHOT NOISE of every bispectrum entry.  Replicated mesh (CPU
    tests, survey catalogues, deterministic mode): with two or more ranks the last rank
    does, for every entry (its full-grid inverse FFT does not depend on the pair
    partition).  ``distributed_mesh`` (box catalogues with the NCCL communicator attached,
    src/threept.cpp ``dist_mesh``): every rank takes part in xi(r) and the entries are
    dealt round robin.  The raw-bispectrum owners of a pair-block run are :func:`owners`
    with the last rank's share cut short by the cost of that branch (``bispec_share``);
    in x-slab mode every rank contributes to every entry."""
    own = owners(form, degrees, num_bins, max(world_size, 1), idx_bin=idx_bin)
    if world_size < 2:
        return own
    if distributed_mesh:
        return np.arange(len(own)) % world_size
    return np.full_like(own, world_size - 1)


def local_entries(form, degrees, num_bins, rank, world_size, idx_bin=0):
    return np.nonzero(owners(form, degrees, num_bins, world_size, idx_bin=idx_bin) == rank)[0]


def pack(out, stat):
    """Partial statistics of one rank -> flat float64 buffer (re, im interleaved)."""
    raw, shot = _STAT_KEYS[stat]
    return np.concatenate([np.ascontiguousarray(out[raw]).view(np.float64),
                           np.ascontiguousarray(out[shot]).view(np.float64)])


def unpack(buf, out, stat):
    raw, shot = _STAT_KEYS[stat]
    n = len(out[raw])
    out = dict(out)
    out[raw] = np.array(buf[:2 * n]).view(np.complex128)
    out[shot] = np.array(buf[2 * n:4 * n]).view(np.complex128)
    return out


def allreduce_result(out, stat, group=None, device=None):
    """Sum the ranks' partial result vectors; every rank gets the full result."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return out
    buf = torch.from_numpy(pack(out, stat))
    if device is not None:
        buf = buf.to(device)
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return unpack(buf.cpu().numpy(), out, stat)


def threept(stat, *args, group=None, device=None, **kwargs):
    """``core.threept`` / ``core.threept_box_arrays`` on this rank's share of the
    entries; the sum over ranks happens inside the call when :func:`init_comm` attached
    the NCCL communicator, else through ``torch.distributed``.  Pass ``x_ptr=...`` style
    arguments through ``kwargs`` exactly as for the single-GPU call."""
    import torch.distributed as dist
    from . import core
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    kwargs.update(part_rank=rank, part_count=world)
    fn = core.threept_box_arrays if kwargs.pop("_arrays", False) else core.threept
    out = fn(stat, *args, **kwargs)
    if core.comm_size() == world:
        return out
    return allreduce_result(out, stat, group=group, device=device)
