"""Multi-rank host logic of the shuffle batch: sharding, the fixed-capacity gather buffer and
its unpacking into the original shuffle order.  CPU: world_size 2 over gloo (the buffer is
filled from the oracle's matrices, standing in for the kernels).  GPU: the same buffer filled
by rp_batch_sparse_device on one rank."""
import ctypes as C
import os
import socket

import numpy as np
import pytest

from conftest import has_gpu, rand_seq


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _pairs():
    rng = np.random.default_rng(42)
    return [(rand_seq(rng, int(rng.integers(14, 30)), "GCGCAU"), rand_seq(rng, int(rng.integers(14, 30)), "GCGCAU"))
            for _ in range(7)]  # odd count: ranks get 4 and 3 shuffles


def oracle_lists(oracle, s1, s2, opts):
    """Thresholded lists x, y, z, v, w from the oracle's matrices, in the reference's creation order
    (src/ractip.cpp:557-567, 578-588, 598-609, 619-628, 639-648)."""
    bp1, up1 = oracle.rnafold(s1, opts.max_w)
    bp2, up2 = oracle.rnafold(s2, opts.max_w)
    hp = oracle.rnaduplex(s1, s2, opts.th_hy)

    def xs(bp, L):
        out = []
        for j in range(1, L):
            for i in range(j - 1, -1, -1):
                p = bp[(i + 1) * (2 * L + 1 - (i + 1)) // 2 + j + 1]
                if p > np.float32(opts.th_ss):
                    out.append((i, j, p))
        return out

    def vs(up):
        if not (opts.min_w > 1 and opts.max_w >= opts.min_w):
            return []
        return [(i, j, up[i][j]) for i in range(up.shape[0]) for j in range(opts.min_w - 1, opts.max_w)
                if up[i][j] > np.float32(opts.th_ac)]
    z = [(i, j, hp[i + 1][j + 1]) for i in range(len(s1)) for j in range(len(s2))
         if hp[i + 1][j + 1] > np.float32(opts.th_hy)]
    return xs(bp1, len(s1)), xs(bp2, len(s2)), z, vs(up1), vs(up2)


def _fill_from_oracle(plan, oracle, opts):
    """Write the oracle's thresholded lists into a local gather buffer, at the plan's layout."""
    from ractip_b200.dist import CNT_BYTES
    buf = np.zeros(plan.nbytes, dtype=np.uint8)
    o_rec, o_cnt = plan.section_offsets()
    recs = plan.rec_view(buf)
    cnts = buf[o_cnt:o_cnt + len(plan.my_pairs) * CNT_BYTES].view(np.int32).reshape(-1, CNT_BYTES // 4)
    for k, (s1, s2) in enumerate(plan.my_pairs):
        S = plan.layouts[plan.rank][k]
        x, y, z, v, w = oracle_lists(oracle, s1, s2, opts)
        for off, lst in ((S.x, x), (S.y, y), (S.z, z), (S.v, v), (S.w, w)):
            for t, (i, j, p) in enumerate(lst):
                recs[off + t] = (i, j, p)
        cnts[k] = (len(x), len(y), len(z), 0, len(v), len(w))
    return buf


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle.oracle import Oracle
        from ractip_b200 import default_model, default_opts
        from ractip_b200.dist import ShardPlan
        opts = default_opts(max_w=6, min_w=3)
        pairs = _pairs()
        plan = ShardPlan(pairs, opts, rank, world)
        oracle = Oracle(default_model())
        local = torch.from_numpy(_fill_from_oracle(plan, oracle, opts))
        gathered = plan.gather(local)
        res = plan.unpack(gathered.numpy())
        # every rank must now hold every shuffle, in the original order
        summary = [(r.x.tolist(), r.y.tolist(), r.z.tolist(), r.v.tolist(), r.w.tolist()) for r in res]
        q.put((rank, summary, plan.shards, plan.nbytes))
    finally:
        dist.destroy_process_group()


def test_two_ranks_gloo_gather_matches_single_process(oracle):
    import torch.multiprocessing as mp
    from ractip_b200 import default_opts
    from ractip_b200.dist import ShardPlan, shard_indices
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    got.sort()
    (r0, s0, shards0, nb0), (r1, s1, shards1, nb1) = got
    assert s0 == s1 and nb0 == nb1
    assert shards0 == [[0, 2, 4, 6], [1, 3, 5]] == [shard_indices(7, 0, 2), shard_indices(7, 1, 2)]
    # single-process truth
    opts = default_opts(max_w=6, min_w=3)
    plan = ShardPlan(_pairs(), opts, 0, 1)
    truth = plan.unpack(_fill_from_oracle(plan, oracle, opts))
    summary = [(r.x.tolist(), r.y.tolist(), r.z.tolist(), r.v.tolist(), r.w.tolist()) for r in truth]
    assert summary == s0
    assert any(len(t[0]) for t in summary) and any(len(t[2]) for t in summary) and any(len(t[3]) for t in summary)


def test_shard_plan_capacities_are_rank_independent():
    from ractip_b200 import default_opts
    from ractip_b200.dist import ShardPlan
    opts = default_opts()
    pairs = _pairs()
    plans = [ShardPlan(pairs, opts, r, 4) for r in range(4)]
    assert len({p.nbytes for p in plans}) == 1
    assert sorted(i for p in plans for i in p.shards[p.rank]) == list(range(len(pairs)))
    assert all(p.rec_bytes % 256 == 0 and p.cnt_bytes % 256 == 0 for p in plans)


@pytest.mark.gpu
def test_device_resident_records_roundtrip(stage, bundled):
    """rp_batch_sparse_device writes the gather buffer in device memory; unpacking it equals rp_run_sparse."""
    import torch
    from ractip_b200 import default_opts
    from ractip_b200.dist import ShardPlan
    opts = default_opts()
    pairs = [(bundled["sequences"][a], bundled["sequences"][b]) for a, b in bundled["pairs"]]
    plan = ShardPlan(pairs, opts, 0, 1)
    buf = torch.zeros(plan.nbytes, dtype=torch.uint8, device="cuda")
    b = stage.batch(plan.my_pairs, opts)
    b.run()
    plan.fill(b, buf)
    b.sync()
    res = plan.unpack(plan.gather(buf).cpu().numpy())
    ref = stage.run_sparse(pairs, opts)
    b.close()
    for a, r in zip(res, ref):
        assert a.x.tolist() == r.x.tolist() and a.y.tolist() == r.y.tolist() and a.z.tolist() == r.z.tolist()
        assert a.v.tolist() == r.v.tolist() and a.w.tolist() == r.w.tolist()
        assert len(r.v) > 0 and len(r.w) > 0
