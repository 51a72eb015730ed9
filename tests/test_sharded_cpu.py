"""CPU suite for the N > 1 path (world_size 2, gloo): the all-reduce plumbing the library calls back
into (mpboot_b200.sharded.make_allreduce), the word-slice layout, and the additivity the pattern
sharding rests on -- per-shard partial scores of the oracle sum to the oracle's total.  The CUDA
library itself needs a GPU; its sharded results are checked against unsharded ones by
tools/sharded_check.py on 2+ GPUs."""
import ctypes as C
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from mpboot_b200 import sharded
    from oracle import portlib
    from tests.helpers import make_case
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cb, stats = sharded.make_allreduce(device="cpu")
        # 1. the callback reduces a host int32 vector in place, exactly like the library calls it
        buf = np.arange(10, dtype=np.int32) * (rank + 1)
        assert cb(None, buf.ctypes.data, len(buf), None) == 0
        assert np.array_equal(buf, np.arange(10, dtype=np.int32) * sum(range(1, world + 1)))
        # 2. word slices partition the padded plane
        c = make_case(20, 3000, 1, 17)
        o = portlib.OracleEngine(c["codes"], c["weights"], 1)
        o.set_ring(c["bn"], c["bs"]); W = o.allocate(True)
        total = o.evaluate_full(True)
        wl, starts = sharded.shard_words(W, world)
        assert wl % 128 == 0 and wl * world >= W and starts == [r * wl for r in range(world)]
        # 3. additivity: partial score over this rank's expanded sites, all-reduced = total score
        pp, sm = o.pattern_parsimony(c["n_inf"])
        site_score = np.repeat(pp.astype(np.int64), c["weights"][: c["n_inf"]])
        lo, hi = starts[rank] * 32, (starts[rank] + wl) * 32
        part = np.array([int(site_score[lo:hi].sum()), len(site_score[lo:hi])], dtype=np.int32)
        assert cb(None, part.ctypes.data, 2, None) == 0
        assert part[0] == total == sm and part[1] == len(site_score)
        # 4. per-replicate REPS partials add up too (no segment wraps in this case)
        rng = np.random.default_rng(3)
        boot = rng.integers(0, 4, size=(6, c["n_inf"])).astype(np.int64)
        site_ptn = np.repeat(np.arange(c["n_inf"]), c["weights"][: c["n_inf"]])
        first = np.concatenate([[0], np.cumsum(c["weights"][: c["n_inf"]])[:-1]])
        mine = (first >= lo) & (first < hi)                      # patterns whose first site is in my slice
        reps = (boot[:, mine] * pp[mine].astype(np.int64)).sum(axis=1).astype(np.int32)
        assert cb(None, reps.ctypes.data, len(reps), None) == 0
        want = portlib.reps(pp, boot.astype(np.uint16), np.array([c["n_inf"]], dtype=np.int32))
        assert np.array_equal(reps, want)
        assert stats["calls"] == 3 and site_ptn.shape[0] == len(site_score)
        q.put((rank, "ok"))
    except Exception as e:          # report instead of hanging the peer
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_allreduce_and_additivity():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_engine_module_has_no_cpu_path():
    from mpboot_b200 import engine
    if engine.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(engine.MpGpuError):
        engine.Engine(shard_rank=0, shard_count=2)
