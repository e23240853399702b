"""CPU-side logic of the training engine: flat storage, bucket planning, fused-optimizer reference
path, gloo data-parallel equivalence (world_size 2)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from edl_b200 import ops
from edl_b200.models import ResNetVd
from edl_b200.parallel import ElasticDataParallel, FlatParams, plan_buckets, choose_algo


def test_flat_params_views_and_order():
    m = torch.nn.Sequential(torch.nn.Linear(10, 7), torch.nn.Linear(7, 3))
    w0 = m[0].weight.detach().clone()
    flat = FlatParams(m)
    g = flat.groups[torch.float32]
    assert torch.equal(m[0].weight, w0)
    # reverse registration order: last layer's bias first
    assert g.entries[0].name == "1.bias" and g.entries[-1].name == "0.weight"
    for e in g.entries:
        assert e.offset % 128 == 0
        assert e.param.data_ptr() == g.param.data_ptr() + e.offset * 4
        assert e.param.grad.data_ptr() == g.grad.data_ptr() + e.offset * 4
    m(torch.randn(4, 10)).sum().backward()
    assert g.grad.abs().sum() > 0
    flat.zero_grad()
    assert g.grad.abs().sum() == 0


def test_bucket_plan_covers_everything():
    m = ResNetVd(18, class_dim=10, width_mult=0.25)
    flat = FlatParams(m)
    buckets = plan_buckets(flat, 64 * 1024)
    for dt, g in flat.groups.items():
        bs = sorted([b for b in buckets if b.dtype == dt], key=lambda b: b.start)
        assert bs[0].start == 0
        for a, b in zip(bs, bs[1:]):
            assert a.start + a.numel == b.start
        assert bs[-1].start + bs[-1].numel == g.numel
    assert [b.order for b in buckets] == sorted(b.order for b in buckets)
    assert sum(len(b.entry_ids) for b in buckets) == len(list(flat.entries()))


def test_choose_algo():
    assert choose_algo(1 << 20, 1, False) == "none"
    assert choose_algo(1 << 20, 8, True) == "multimem"
    assert choose_algo(1 << 20, 8, False) == "twoshot"
    assert choose_algo(1024, 8, True) == "twoshot"
    assert choose_algo(1 << 20, 6, True, prefer="twoshot") == "twoshot"


def test_sgd_cpu_matches_torch():
    torch.manual_seed(0)
    m = torch.nn.Linear(20, 5)
    ref = torch.nn.Linear(20, 5)
    ref.load_state_dict(m.state_dict())
    flat = FlatParams(m)
    opt = ops.FlatSGDMomentum(flat, lr=0.1, momentum=0.9, weight_decay=1e-3)
    ropt = torch.optim.SGD(ref.parameters(), lr=0.1, momentum=0.9, weight_decay=1e-3)
    x = torch.randn(8, 20)
    for _ in range(4):
        flat.zero_grad()
        ropt.zero_grad()
        m(x).pow(2).sum().backward()
        ref(x).pow(2).sum().backward()
        opt.step()
        ropt.step()
    for p, q in zip(m.parameters(), ref.parameters()):
        assert torch.allclose(p, q, atol=1e-5)


def _dp_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    m = ResNetVd(18, class_dim=8, width_mult=0.125)
    dp = ElasticDataParallel(m, bucket_cap_mb=0.05)
    opt = ops.FlatSGDMomentum(dp.flat, lr=0.05)
    torch.manual_seed(100)
    xs = torch.randn(4, 3, 32, 32)
    ts = torch.softmax(torch.randn(4, 8), -1)
    x, t = xs[rank * 2:(rank + 1) * 2], ts[rank * 2:(rank + 1) * 2]
    for _ in range(2):
        dp.zero_grad()
        loss = ops.soft_cross_entropy(dp(x.contiguous(memory_format=torch.channels_last)), t)
        loss.backward()
        dp.finish()
        opt.step()
    flat = torch.cat([g.param.flatten() for g in dp.flat.groups.values()])
    gathered = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    if rank == 0:
        q.put(float((gathered[0] - gathered[1]).abs().max()))
    dist.destroy_process_group()


def test_gloo_data_parallel_keeps_ranks_in_sync():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert q.get(timeout=5) < 1e-6


def test_trainer_cpu_step_eval_and_state_roundtrip():
    from edl_b200.models import to_train_dtype
    from edl_b200.trainer import StudentTrainer

    torch.manual_seed(0)
    m = to_train_dtype(ResNetVd(18, class_dim=8, width_mult=0.125), torch.float32)
    tr = StudentTrainer(m, 4, image_shape=(3, 32, 32), num_classes=8, lr=0.05, target_kind="labels", dtype=torch.float32)
    x = torch.randn(4, 3, 32, 32).contiguous(memory_format=torch.channels_last)
    y = torch.tensor([0, 1, 2, 3])
    l0 = float(tr.step(x, y))
    for _ in range(10):
        l1 = float(tr.step(x, y))
    assert l1 < l0
    ev = tr.evaluate([(x, y)])
    assert ev["n"] == 4 and 0.0 <= ev["acc1"] <= ev["acc5"] <= 1.0
    sd = tr.state_dict()
    m2 = to_train_dtype(ResNetVd(18, class_dim=8, width_mult=0.125), torch.float32)
    tr2 = StudentTrainer(m2, 4, image_shape=(3, 32, 32), num_classes=8, lr=0.01, target_kind="labels", dtype=torch.float32)
    tr2.load_state_dict(sd)
    assert tr2.opt.lr == tr.opt.lr
    a, b = float(tr.step(x, y)), float(tr2.step(x, y))
    assert abs(a - b) < 1e-4      # identical params + momentum => identical next step


def test_recompute_matches_plain_backward():
    from edl_b200.models import to_train_dtype

    torch.manual_seed(1)
    a = to_train_dtype(ResNetVd(18, class_dim=6, width_mult=0.125), torch.float32)
    b = to_train_dtype(ResNetVd(18, class_dim=6, width_mult=0.125, recompute=True), torch.float32)
    b.load_state_dict(a.state_dict())
    x = torch.randn(3, 3, 32, 32)
    for m in (a, b):
        m.train()
        m(x).square().mean().backward()
    for (n, p), (_, q) in zip(a.named_parameters(), b.named_parameters()):
        assert torch.allclose(p.grad, q.grad, atol=1e-5, rtol=1e-4), n


def test_loss_scaling_skips_overflow_and_matches_unscaled():
    from edl_b200.models import to_train_dtype
    from edl_b200.trainer import StudentTrainer

    def make(scale):
        torch.manual_seed(3)
        m = to_train_dtype(ResNetVd(18, class_dim=8, width_mult=0.125), torch.float32)
        return StudentTrainer(m, 4, image_shape=(3, 32, 32), num_classes=8, lr=0.05, target_kind="labels",
                              dtype=torch.float32, loss_scaling=scale)

    x = torch.randn(4, 3, 32, 32).contiguous(memory_format=torch.channels_last)
    y = torch.tensor([0, 1, 2, 3])
    plain, scaled = make(None), make(1024.0)
    for _ in range(3):
        a, b = float(plain.step(x, y)), float(scaled.step(x, y))
        assert abs(a - b) < 1e-3 * max(1.0, abs(a))
    # poison one step: the update must be skipped and the scale halved
    before = [p.detach().clone() for p in scaled.model.parameters()]
    bad = x.clone()
    bad[0, 0, 0, 0] = float("inf")
    scaled.step(bad, y)
    for p, q in zip(scaled.model.parameters(), before):
        assert torch.equal(p, q)
    assert float(scaled.scaler.scale) == 512.0


def test_dgc_world1_sparse_steps_reduce_loss():
    from edl_b200.models import to_train_dtype
    from edl_b200.parallel import DGCMomentum
    from edl_b200.trainer import StudentTrainer

    torch.manual_seed(5)
    m = to_train_dtype(ResNetVd(18, class_dim=8, width_mult=0.125), torch.float32)
    tr = StudentTrainer(m, 4, image_shape=(3, 32, 32), num_classes=8, target_kind="labels", dtype=torch.float32,
                        optimizer=lambda flat: DGCMomentum(flat, lr=0.05, momentum=0.9, weight_decay=0.0,
                                                           rampup_begin_step=2, rampup_step=4,
                                                           sparsity=(0.5, 0.9)))
    x = torch.randn(4, 3, 32, 32).contiguous(memory_format=torch.channels_last)
    y = torch.tensor([0, 1, 2, 3])
    losses = [float(tr.step(x, y)) for _ in range(25)]
    assert losses[-1] < losses[0]
    assert tr.opt.t == 25 and tr.opt.current_sparsity() == 0.9
    numel = sum(g.numel for g in tr.dp.flat.groups.values())
    assert tr.opt.sent_elems < 23 * numel * 0.6      # compressed steps shipped a fraction of the gradient


def _dgc_worker(rank, world, port, q):
    from edl_b200.parallel import DGCMomentum

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    m = ResNetVd(18, class_dim=8, width_mult=0.125)
    dp = ElasticDataParallel(m, bucket_cap_mb=0.05)
    opt = DGCMomentum(dp.flat, dp=dp, lr=0.05, weight_decay=0.0, rampup_begin_step=1, rampup_step=2, sparsity=(0.9,))
    torch.manual_seed(100)
    xs = torch.randn(4, 3, 32, 32)
    ts = torch.softmax(torch.randn(4, 8), -1)
    x, t = xs[rank * 2:(rank + 1) * 2], ts[rank * 2:(rank + 1) * 2]
    enabled = []
    for _ in range(4):
        dp.zero_grad()
        loss = ops.soft_cross_entropy(dp(x.contiguous(memory_format=torch.channels_last)), t)
        loss.backward()
        enabled.append(dp.enabled)
        dp.finish()
        opt.step()
    flat = torch.cat([g.param.flatten() for g in dp.flat.groups.values()])
    gathered = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    if rank == 0:
        q.put((float((gathered[0] - gathered[1]).abs().max()), enabled))
    dist.destroy_process_group()


def test_dgc_gloo_ranks_stay_in_sync_with_sparse_exchange():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_dgc_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    diff, enabled = q.get(timeout=5)
    assert diff < 1e-6                         # sparse all-gather applies the same update everywhere
    assert enabled == [True, False, False, False]   # dense all-reduce only during the ramp-up step


def _elastic_worker(rank, world, port, q):
    from edl_b200.models import to_train_dtype
    from edl_b200.trainer import StudentTrainer

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    m = to_train_dtype(ResNetVd(18, class_dim=8, width_mult=0.125), torch.float32)
    tr = StudentTrainer(m, 2, image_shape=(3, 32, 32), num_classes=8, lr=0.05, target_kind="labels", dtype=torch.float32,
                        bucket_cap_mb=0.05)
    torch.manual_seed(50 + rank)
    x = torch.randn(2, 3, 32, 32).contiguous(memory_format=torch.channels_last)
    y = torch.randint(0, 8, (2,))

    def spread():
        flat = torch.cat([g.param.flatten() for g in tr.dp.flat.groups.values()])
        g = [torch.zeros_like(flat) for _ in range(world)]
        dist.all_gather(g, flat)
        return float((g[0] - g[1]).abs().max())

    tr.step(x, y)
    d0 = spread()
    solo = [dist.new_group(ranks=[r]) for r in range(world)][rank]
    tr.rebuild(solo)
    tr.step(x, y)
    d1 = spread()
    tr.rebuild(None)
    tr.sync_from(0)
    tr.step(x, y)
    d2 = spread()
    if rank == 0:
        q.put((d0, d1, d2, tr.dp.world))
    dist.destroy_process_group()


def test_elastic_rebuild_in_place_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + os.getpid() % 2000
    procs = [ctx.Process(target=_elastic_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    d0, d1, d2, world = q.get(timeout=5)
    assert d0 < 1e-6 and d1 > 1e-6 and d2 < 1e-6 and world == 2


def test_step_pipelined_cpu_fallback_matches_step():
    from edl_b200.models import ResNetVd, to_train_dtype
    from edl_b200.trainer import StudentTrainer

    def run(pipelined):
        torch.manual_seed(0)
        m = to_train_dtype(ResNetVd(18, class_dim=10, width_mult=0.125), torch.float32, torch.device("cpu")).train()
        tr = StudentTrainer(m, 4, image_shape=(3, 32, 32), num_classes=10, lr=0.05, use_graph=False, dtype=torch.float32)
        g = torch.Generator().manual_seed(1)
        x = torch.randn(4, 3, 32, 32, generator=g).contiguous(memory_format=torch.channels_last)
        t = torch.softmax(torch.randn(4, 10, generator=g), -1)
        if pipelined:
            return [tr.step_pipelined(x, t).item() for _ in range(3)]
        return [float(tr.step(x, t)) for _ in range(3)]

    assert run(True) == run(False)


def _hier_worker(rank, world, port, q, fake_hosts):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    if fake_hosts:
        os.environ["EDL_FAKE_HOST"] = "node%d" % (rank // 2)       # 2 "hosts" x 2 ranks
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    m = torch.nn.Sequential(torch.nn.Linear(37, 301), torch.nn.Tanh(), torch.nn.Linear(301, 53), torch.nn.Tanh(),
                            torch.nn.Linear(53, 8)).double()          # odd sizes: slices that do not divide evenly
    dp = ElasticDataParallel(m, bucket_cap_mb=0.01)
    assert dp.hier == bool(fake_hosts) and (not fake_hosts or (dp.local_world == 2 and dp.local_rank == rank % 2))
    assert len(dp.buckets) >= 2
    opt = ops.FlatSGDMomentum(dp.flat, lr=0.05)
    torch.manual_seed(100)
    xs = torch.randn(8, 37).double()
    ts = torch.softmax(torch.randn(8, 8), -1).double()
    x, t = xs[rank * 2:(rank + 1) * 2], ts[rank * 2:(rank + 1) * 2]
    for _ in range(3):
        dp.zero_grad()
        loss = ops.soft_cross_entropy(dp(x), t)
        loss.backward()
        dp.finish()
        opt.step()
    assert dp.agree(1.0 if rank == 3 else 0.0) == (1.0, 0)
    dp.broadcast_parameters(0)
    flat = torch.cat([(g.master if g.master is not None else g.param).flatten().double() for g in dp.flat.groups.values()])
    gathered = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    if rank == 0:
        q.put((float(max((gathered[0] - g).abs().max() for g in gathered)), flat.clone()))
    dist.barrier()
    dist.destroy_process_group()


def test_hierarchical_allreduce_matches_the_flat_one():
    """Ranks on two (fake) hosts: sum inside the host, 1/L slice per rank across hosts, second local sum -- same
    parameters as the flat all-reduce after three optimizer steps (float64: summation order is the only difference)."""
    ctx = mp.get_context("spawn")
    results = []
    for fake in (False, True):
        q = ctx.Queue()
        port = 29500 + (os.getpid() + 7 + int(fake)) % 2000
        procs = [ctx.Process(target=_hier_worker, args=(r, 4, port, q, fake)) for r in range(4)]
        for p in procs:
            p.start()
        spread, flat = q.get(timeout=90)
        for p in procs:
            p.join(60)
            assert p.exitcode == 0
        assert spread == 0.0                     # every rank holds identical parameters
        results.append(flat)
    assert torch.allclose(results[0], results[1], rtol=1e-9, atol=1e-12)


def test_buckets_tile_the_flat_buffer_in_16_byte_multiples():
    """Entries start on 128-element boundaries; the alignment gap behind a bucket's last tensor belongs to that
    bucket (the two-shot kernels want multiples of 16 bytes -- a DeepFM table with an odd row count broke that)."""
    from edl_b200.parallel.ddp import plan_buckets
    from edl_b200.parallel.flat import FlatParams

    m = torch.nn.Sequential(torch.nn.Linear(13, 7), torch.nn.Linear(7, 5), torch.nn.Embedding(1001, 3), torch.nn.Linear(5, 3))
    flat = FlatParams(m)
    for cap in (64, 4096, 1 << 20):
        buckets = plan_buckets(flat, cap)
        for g in flat.groups.values():
            mine = sorted((b for b in buckets if b.dtype == g.dtype), key=lambda b: b.start)
            assert mine[0].start == 0 and sum(b.numel for b in mine) == g.numel
            assert all(b.numel % 8 == 0 and b.start % 8 == 0 for b in mine)
            for b, nxt in zip(mine, mine[1:]):
                assert b.start + b.numel == nxt.start
