"""Multi-GPU host logic: shard the observation axis, all-reduce only the scalar block.

Every function of the path is independent per observation and the scalars are plain sums
(SURVEY §8e), so rank g owns the contiguous block [g*N/G, (g+1)*N/G) (whole rows for the
Categorical likelihood), state / inputs / β, γ stay sharded, RNG counters use the GLOBAL
observation index (pass `i0=lo` to the sampling verbs), and the only collective is ONE
ncclAllReduce(sum, double, 8) of the scalar block, issued by libaugcuda on the ctx stream.
torch.distributed is used for the rendezvous (broadcast of the NCCL unique id) only.
"""
from __future__ import annotations

import ctypes as C

import torch

from ._lib import NSCALARS, check


def shard_bounds(n: int, world: int, rank: int):
    """Contiguous block partition: [lo, hi) of rank `rank`; sizes differ by at most one and the
    lower ranks take the remainder.  lo doubles as the global index offset `i0` of the shard."""
    if world < 1 or not (0 <= rank < world) or n < 0:
        raise ValueError("bad shard request")
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def combine_scalars_host(blocks):
    """Reference semantics of the collective for host-side checks: element-wise sum over ranks of the
    additive slots; slot 2 (= slot 0 + slot 1) and slot 5 (= slot 3 + slot 4) stay consistent."""
    out = [0.0] * NSCALARS
    for b in blocks:
        for k in range(NSCALARS):
            out[k] += float(b[k])
    return out


def init_comm(ctx, group=None):
    """Attach an NCCL communicator to ctx: rank 0 creates the unique id, torch.distributed broadcasts it."""
    import torch.distributed as dist
    if not dist.is_initialized():
        raise RuntimeError("torch.distributed is not initialised")
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    uid = (C.c_char * 128)()
    if rank == 0:
        check(ctx.lib.aug_comm_get_unique_id(uid))
    payload = [bytes(uid.raw)]
    dist.broadcast_object_list(payload, src=0, group=group)
    uid2 = (C.c_char * 128).from_buffer_copy(payload[0])
    check(ctx.lib.aug_comm_init(ctx.h, world, rank, uid2))
    ctx.comm_ready = True
    ctx.world, ctx.rank = world, rank
    return ctx


def allreduce_scalars_(ctx, scal: torch.Tensor):
    """In-place sum over ranks of the device scalar block (ncclAllReduce on the ctx stream)."""
    if not ctx.comm_ready:
        raise RuntimeError("communicator not initialised (call init_comm)")
    ctx.enter()
    check(ctx.lib.aug_allreduce_scalars(ctx.h, C.c_void_p(scal.data_ptr()), int(scal.numel())))
    ctx.leave()
    return scal


def init_p2p(ctx, group=None, fused=True):
    """Attach the peer-memory mailbox (NVLink / NVSwitch): every rank exports a cudaIpc handle of its mailbox (slots + bulk area),
    torch.distributed all-gathers the handles, every rank maps its peers.  With fused=True the scalar-producing
    verbs (cavi_step_ with want_elbo, expected_elbo_terms, sampled_loglik_terms) return the sums over ALL ranks,
    exchanged inside their own reducing kernel — no separate all-reduce launch (include/augcuda.h)."""
    import torch.distributed as dist
    if not dist.is_initialized():
        raise RuntimeError("torch.distributed is not initialised")
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    h = (C.c_char * 64)()
    check(ctx.lib.aug_comm_p2p_export(ctx.h, h, None))
    gathered = [None] * world
    dist.all_gather_object(gathered, bytes(h.raw), group=group)
    check(ctx.lib.aug_comm_p2p_attach(ctx.h, world, rank, b"".join(gathered)))
    dist.barrier(group=group)                      # nobody starts exchanging before everybody has mapped everybody
    ctx.p2p_ready = True
    ctx.world, ctx.rank = world, rank
    if fused:
        set_fused(ctx, True)
    return ctx


def set_fused(ctx, on=True):
    check(ctx.lib.aug_comm_set_fused(ctx.h, 1 if on else 0))
    ctx.fused = bool(on)
    if not on:
        ctx.deferred = getattr(ctx, "deferred", False)
    return ctx


def set_deferred(ctx, on=True):
    """Split-phase exchange (fused mode): cavi_step_ / the expected-ELBO verbs only publish their sums; the gather rides in
    the next aux_sample_ launch or in ctx.flush() / ctx.sync().  Read device scalars only after one of those
    (the float-returning verbs flush by themselves).  include/augcuda.h: aug_comm_set_deferred."""
    check(ctx.lib.aug_comm_set_deferred(ctx.h, 1 if on else 0))
    ctx.deferred = bool(on)
    return ctx


def allreduce_scalars_p2p_(ctx, scal: torch.Tensor, count: int = 7):
    """In-place sum over ranks of the first `count` (<= 7) doubles of the scalar block through the mailbox."""
    if not getattr(ctx, "p2p_ready", False):
        raise RuntimeError("mailbox not attached (call init_p2p)")
    ctx.enter()
    check(ctx.lib.aug_allreduce_scalars_p2p(ctx.h, C.c_void_p(scal.data_ptr()), int(count)))
    ctx.leave()
    return scal


def bind_host_to_gpu_numa_node(device: int):
    """Pin this process (threads it starts later included) to the CPUs of the NUMA node the GPU hangs off, so that the
    pinned host buffers it allocates from now on (first touch) and the library's staging copies stay on the socket that
    owns the GPU's PCIe root port.  One process per GPU: without it every rank's staging memory can land on one node and
    the host-buffer verbs of 8 ranks share that node's memory bandwidth (round-1 SCALE: 86 -> 16 GB/s per GPU).
    Returns {"node": k, "cpus": n} or None when the platform exposes no NUMA placement (single node, VM)."""
    import os
    try:
        bus = torch.cuda.get_device_properties(device).pci_bus_id        # torch >= 2.x
        dom = getattr(torch.cuda.get_device_properties(device), "pci_domain_id", 0)
        dev = getattr(torch.cuda.get_device_properties(device), "pci_device_id", 0)
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            return None
        nodes = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()]
        if len(nodes) < 2:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return {"node": node, "cpus": len(cpus)}
    except (OSError, ValueError, AttributeError, RuntimeError):
        return None
