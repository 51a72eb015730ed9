"""Multi-GPU parity check, run under torchrun on N GPUs of one box:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/sharded_check.py
Every rank holds one word slice (pattern sharding) and also builds the unsharded engine on its own
GPU; everything the sharded context returns -- view lengths, scores, every insertion score, pattern
scores, replicate scores, and whole plain / -bb SPR searches -- must equal the unsharded result."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from mpboot_b200 import engine, sharded  # noqa: E402
from tests.helpers import make_boot, make_case  # noqa: E402
import bench  # noqa: E402


def bb_run(eng, c, boot, seg, seed):
    B = boot.shape[0]
    bl = np.full(B, -float(np.iinfo(np.int64).max)); bc = np.zeros(B, dtype=np.int32); bt = np.full(B, -1, dtype=np.int32)
    tl = engine.Treels(c["n"])
    rs = engine.HostRng(seed)
    ret, bn, bs, nins, ncalls, nreps = eng.optimize_spr_bb(c["bn"], c["bs"], tl.hooks(rs.fn, rs.user), bl, bc, bt, 0.0)
    return ret, bn, bs, nins, ncalls, nreps, bl, bc, bt, tl.logl(), tl.materialized()


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    st = torch.cuda.Stream(); torch.cuda.set_stream(st)
    ok = True
    exchanges = os.environ.get("SHARDED_EXCHANGE", "peer,nccl").split(",")
    # (n, sites, datatype, seed, compress patterns?)  -- with and without site == pattern identity
    for exchange, (n, L, dt, seed, compress) in [(x, cs) for x in exchanges for cs in
                                                 [(40, 9000, 1, 11, False), (24, 6000, 2, 5, True), (64, 20000, 1, 41, True), (20, 5000, 6, 9, False)]]:
        c = make_case(n, L, dt, seed)
        if not compress:
            from mpboot_b200 import hostprep, synth
            chars = synth.evolve_alignment(n, L, dt, 0.05, seed)
            prep = hostprep.prepare(chars, dt, compress=False)
            c.update(chars=prep["chars"], codes=prep["codes"], weights=prep["weights"], n_inf=prep["n_inf"])
        one = engine.Engine(device=local, stream=st.cuda_stream)
        sh = sharded.sharded_engine(local, stream=st.cuda_stream, exchange=exchange)
        for e in (one, sh):
            e.load_alignment(c["codes"], c["weights"], dt)
            e.set_tree(c["bn"], c["bs"])
        assert sh.shard_words * world >= one.shard_words
        s1, s2 = one.tree_score(), sh.tree_score()
        assert s1 == s2, (s1, s2)
        for node in (n + 1, 2 * n - 2):
            for slot in range(3):
                assert one.view_length(node, slot) == sh.view_length(node, slot)
        p1, q1 = one.pattern_parsimony(); p2, q2 = sh.pattern_parsimony()
        assert q1 == q2 and np.array_equal(p1, p2)
        order = one.visit_order()
        assert np.array_equal(order, sh.visit_order())
        a = one.scan_visits(order, 1, 2 * n - 2, 1, 6); b = sh.scan_visits(order, 1, 2 * n - 2, 1, 6)
        assert all(np.array_equal(x, y) for x, y in zip(a, b))
        # replicate scores, with wrap-prone replicates
        ninf = c["n_inf"]
        hv = [(1, 3, 40000), (2, 5, 300), (2, 6, 65535)] + [(3, k, 900) for k in range(0, 40)]
        boot = make_boot(c, 48, seed, heavy=hv)
        seg = bench.do_segmenting(p1[:ninf], c["weights"], ninf)
        for e in (one, sh):
            e.load_replicates(boot, seg)
        assert np.array_equal(one.reps_current_tree(), sh.reps_current_tree())
        idx = np.arange(-1, min(len(a[1]), 700), dtype=np.int32)
        assert np.array_equal(one.reps_candidates(idx), sh.reps_candidates(idx))
        # plain and -bb searches: same decisions, same draws (same RNG seed on every rank), same result
        r1 = engine.HostRng(77); r2 = engine.HostRng(77)
        x = one.optimize_spr(c["bn"], c["bs"], r1.fn, 1, 6, rng_user=r1.user)
        y = sh.optimize_spr(c["bn"], c["bs"], r2.fn, 1, 6, rng_user=r2.user)
        assert x[0] == y[0] and np.array_equal(x[1], y[1]) and np.array_equal(x[2], y[2]) and x[3] == y[3]
        assert r1.state.value == r2.state.value
        u = bb_run(one, c, boot, seg, 5); v = bb_run(sh, c, boot, seg, 5)
        for k in range(len(u)):
            assert np.array_equal(np.asarray(u[k]), np.asarray(v[k])), "bb result %d differs" % k
        # -cost (Sankoff): fresh contexts (the cost matrix goes in before any replicates); scores, node scores, pattern vector,
        # every insertion score with its early-exit bound, whole searches in both modes, RAS
        S = one.S
        rng = np.random.default_rng(seed)
        cost = rng.integers(1, 5, size=(S, S)); cost = np.minimum(cost, cost.T); np.fill_diagonal(cost, 0)
        cseg = np.array([x for x in range(160, ninf, 160)] + [ninf], dtype=np.int32)
        one2 = engine.Engine(device=local, stream=st.cuda_stream)
        sh2 = sharded.sharded_engine(local, stream=st.cuda_stream, exchange=exchange)
        for e in (one2, sh2):
            e.load_alignment(c["codes"], c["weights"], dt)
            e.set_cost_matrix(cost, cseg)
            e.set_tree(c["bn"], c["bs"])
        cs1, cs2 = one2.tree_score(), sh2.tree_score()
        assert cs1 == cs2, (cs1, cs2)
        assert np.array_equal(one2.sankoff_layout()[1], sh2.sankoff_layout()[1])
        for node in (n + 1, 2 * n - 2):
            for slot in range(3):
                assert one2.view_length(node, slot) == sh2.view_length(node, slot)
        p1c, q1c = one2.pattern_parsimony(); p2c, q2c = sh2.pattern_parsimony()
        assert q1c == q2c and np.array_equal(p1c[:ninf], p2c[:ninf])
        ac = one2.scan_visits(order, 1, 2 * n - 2, 1, 6); bc_ = sh2.scan_visits(order, 1, 2 * n - 2, 1, 6)
        assert all(np.array_equal(x, y) for x, y in zip(ac, bc_))
        assert np.array_equal(one2.scan_bounds(len(ac[1])), sh2.scan_bounds(len(ac[1])))
        for exact in (0, 1):
            one2.set_option("sankoff_exact", exact); sh2.set_option("sankoff_exact", exact)
            r1 = engine.HostRng(91); r2 = engine.HostRng(91)
            x = one2.optimize_spr(c["bn"], c["bs"], r1.fn, 1, 6, rng_user=r1.user)
            y = sh2.optimize_spr(c["bn"], c["bs"], r2.fn, 1, 6, rng_user=r2.user)
            assert x[0] == y[0] and np.array_equal(x[1], y[1]) and np.array_equal(x[2], y[2]) and x[3] == y[3]
            assert r1.state.value == r2.state.value
        one2.set_option("sankoff_exact", 0); sh2.set_option("sankoff_exact", 0)
        r1 = engine.HostRng(92); r2 = engine.HostRng(92)
        x = one2.stepwise_addition(4242, 5, r1.fn, rng_user=r1.user); y = sh2.stepwise_addition(4242, 5, r2.fn, rng_user=r2.user)
        assert x[0] == y[0] and np.array_equal(x[1], y[1]) and np.array_equal(x[2], y[2]) and r1.state.value == r2.state.value
        one2.close(); sh2.close()
        if rank == 0:
            print("   -cost: sharded x%d == unsharded (score %d, %d insertions, searches in both modes, RAS score %d)" % (world, cs1, len(ac[1]), x[0]), flush=True)
        if rank == 0:
            calls, elems = (sh.allreduce_stats["calls"], sh.allreduce_stats["elements"]) if sh.allreduce_stats else sh.peer_stats()[:2]
            print("[%s] case n=%d L=%d dt=%d identity=%s: sharded x%d == unsharded (score %d, %d insertions, %d exchange steps, %d int32)"
                  % (exchange, n, L, dt, not compress, world, s1, len(a[1]), calls, elems), flush=True)
        if exchange == "peer":
            assert sh.peer_stats()[2] == 0, "peer exchange timed out waiting for a shard"
        one.close(); sh.close()
    # ---- replicate shards (mpgpu_set_replicate_shards): whole alignment on every GPU, B / world replicates each ----
    for (n, L, dt, seed, B) in [(30, 4000, 1, 21, 50), (22, 1500, 2, 22, 37), (26, 3000, 1, 23, 64)]:
        c = make_case(n, L, dt, seed)
        ninf = c["n_inf"]
        one = engine.Engine(device=local, stream=st.cuda_stream)
        rs = engine.Engine(device=local, stream=st.cuda_stream)
        rs.set_replicate_shards(rank, world)
        sharded.connect_peers(rs)
        for e in (one, rs):
            e.load_alignment(c["codes"], c["weights"], dt)
            e.set_tree(c["bn"], c["bs"])
        p1, _ = one.pattern_parsimony()
        hv = [(1, 3, 40000), (2, 5, 300), (B - 1, 6, 65535)] + [(3, k, 900) for k in range(0, 40)]
        boot = make_boot(c, B, seed, heavy=hv)
        seg = bench.do_segmenting(p1[:ninf], c["weights"], ninf)
        for e in (one, rs):
            e.load_replicates(boot, seg)
        assert np.array_equal(one.reps_current_tree(), rs.reps_current_tree())
        order = one.visit_order()
        a = one.scan_visits(order, 1, 2 * n - 2, 1, 6); b = rs.scan_visits(order, 1, 2 * n - 2, 1, 6)
        assert all(np.array_equal(x, y) for x, y in zip(a, b))
        idx = np.arange(-1, min(len(a[1]), 500), dtype=np.int32)
        assert np.array_equal(one.reps_candidates(idx), rs.reps_candidates(idx))
        u = bb_run(one, c, boot, seg, 5); v = bb_run(rs, c, boot, seg, 5)
        for k in range(len(u)):
            assert np.array_equal(np.asarray(u[k]), np.asarray(v[k])), "replicate shards: bb result %d differs" % k
        # -cost -bb: refused on pattern shards, runs on replicate shards
        S = one.S
        rng = np.random.default_rng(seed)
        cost = rng.integers(1, 5, size=(S, S)); cost = np.minimum(cost, cost.T); np.fill_diagonal(cost, 0)
        cseg = np.array([x for x in range(160, ninf, 160)] + [ninf], dtype=np.int32)
        one2 = engine.Engine(device=local, stream=st.cuda_stream)
        rs2 = engine.Engine(device=local, stream=st.cuda_stream)
        rs2.set_replicate_shards(rank, world)
        sharded.connect_peers(rs2)
        boot2 = make_boot(c, B, seed + 1)
        for e in (one2, rs2):
            e.load_alignment(c["codes"], c["weights"], dt)
            e.set_cost_matrix(cost, cseg)
            e.set_tree(c["bn"], c["bs"])
            e.load_replicates(boot2, cseg)
        ac = one2.scan_visits(order, 1, 2 * n - 2, 1, 6); bc_ = rs2.scan_visits(order, 1, 2 * n - 2, 1, 6)
        assert all(np.array_equal(x, y) for x, y in zip(ac, bc_))
        idx = np.arange(-1, min(len(ac[1]), 300), dtype=np.int32)
        assert np.array_equal(one2.reps_candidates(idx), rs2.reps_candidates(idx))
        u2 = bb_run(one2, c, boot2, cseg, 6); v2 = bb_run(rs2, c, boot2, cseg, 6)
        for k in range(len(u2)):
            assert np.array_equal(np.asarray(u2[k]), np.asarray(v2[k])), "replicate shards: -cost -bb result %d differs" % k
        assert rs.peer_stats()[2] == 0 and rs2.peer_stats()[2] == 0
        if rank == 0:
            print("[replicate shards] case n=%d L=%d dt=%d B=%d: x%d == unsharded (REPS vectors, whole -bb search: %d calls, %d vectors; -cost -bb search: %d vectors; %d exchange steps)"
                  % (n, L, dt, B, world, u[4], u[5], u2[5], rs.peer_stats()[0]), flush=True)
        for e in (one, rs, one2, rs2):
            e.close()
    dist.barrier()
    if rank == 0:
        print("SHARDED_CHECK_OK world=%d" % world, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
