"""Pattern-sharded execution over torch.distributed: one process per GPU, every rank
holds a contiguous word slice of every bit plane, topology / plans / RNG are replicated, and the
only exchange is an in-place int32 SUM all-reduce of small count vectors, which the library
requests through the mpgpu_allreduce_fn callback installed here (NCCL over NVLink on GPUs; the
same code runs on the gloo backend with host buffers for the CPU tests of the plumbing)."""
import ctypes as C

import numpy as np

ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p)


class _DevInt32:
    """Zero-copy view of a device int32 vector for torch (CUDA array interface)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i4", "data": (ptr, False), "version": 2}


def shard_words(total_words, world, pad=128):
    """Words per plane of each shard and the first word of every shard: the library's rule
    (mpgpu_api.cu build_planes): pad the plane to a multiple of pad*world, split evenly."""
    quantum = pad * world
    glob = max((total_words + quantum - 1) // quantum * quantum, quantum)
    wl = glob // world
    return wl, [r * wl for r in range(world)]


def make_allreduce(group=None, device="cuda"):
    """Returns (callback object, stats dict).  Keep the callback object alive as long as the
    context uses it.  device="cpu": dev_buf is host memory (gloo tests of the plumbing)."""
    import torch
    import torch.distributed as dist
    stats = {"calls": 0, "elements": 0}

    def _cb(_user, ptr, count, stream):
        try:
            stats["calls"] += 1
            stats["elements"] += int(count)
            if device == "cpu":
                buf = (C.c_int32 * count).from_address(ptr)
                t = torch.from_numpy(np.frombuffer(buf, dtype=np.int32))
                dist.all_reduce(t, group=group)
            else:
                t = torch.as_tensor(_DevInt32(ptr, int(count)), device="cuda")
                cur = torch.cuda.current_stream()
                if stream and cur.cuda_stream != stream:
                    with torch.cuda.stream(torch.cuda.ExternalStream(stream)):
                        dist.all_reduce(t, group=group)
                else:
                    dist.all_reduce(t, group=group)
            return 0
        except Exception as e:          # never let an exception cross the C boundary
            import sys
            print("mpboot_b200.sharded: all-reduce failed: %r" % (e,), file=sys.stderr)
            return 1

    return ALLREDUCE_FN(_cb), stats


def connect_peers(eng, group=None, capacity=1 << 16):
    """The exchange step inside the library (mpgpu_peer_prepare / mpgpu_peer_connect): every rank exports the CUDA IPC
    handle of its exchange region, the handles travel once over torch.distributed, and from then on the library sums its
    count vectors itself with a one-shot all-reduce kernel over NVLink peer memory -- no NCCL launch, no host callback."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    mine = torch.from_numpy(eng.peer_prepare(capacity).copy())
    if dist.get_backend(group) == "nccl":
        mine = mine.cuda()
    allh = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allh, mine, group=group)
    eng.peer_connect(np.stack([h.cpu().numpy() for h in allh]))
    dist.barrier(group=group)                   # every rank has mapped every region before anyone pushes into one


def sharded_engine(device, stream=None, group=None, exchange="nccl"):
    """An Engine holding this rank's word slice with its exchange step installed: every call then returns complete
    results on every rank.  exchange = "nccl": the all-reduce callback over torch.distributed; "peer": the library's own
    one-shot all-reduce over NVLink peer memory (one NVSwitch box, one process per GPU)."""
    import torch.distributed as dist
    from . import engine
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    eng = engine.Engine(device=device, stream=stream, shard_rank=rank, shard_count=world)
    if exchange == "peer":
        connect_peers(eng, group)
        eng.allreduce_stats = None
    else:
        cb, stats = make_allreduce(group)
        eng.set_allreduce(cb)
        eng.allreduce_stats = stats
    return eng
