"""Fused all-reduce kernels over NVSwitch peer memory (needs >= 2 GPUs): one-shot, two-shot,
NVLS multimem, broadcast, scalar all-gather, and the ElasticDataParallel engine end to end."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]


def _worker(rank, world, port, q):
    try:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
        from edl_b200.ops import native
        from edl_b200.parallel.symm import SymmetricPool

        C = native()
        pool = SymmetricPool(64 << 20, device=dev)
        res = {"multicast": pool.has_multicast}
        for dtype, tol in ((torch.bfloat16, 2e-2), (torch.float32, 1e-5)):
            for n in (8 * 1024, 1_000_000 // 8 * 8, 6_000_000):
                sl = pool.alloc(n, dtype)
                torch.manual_seed(rank)
                src = torch.randn(n, device=dev).to(dtype)
                gathered = [torch.empty_like(src) for _ in range(world)]
                dist.all_gather(gathered, src)
                ref = sum(g.float() for g in gathered) / world
                for algo in (["twoshot", "multimem"] if pool.has_multicast else ["twoshot"]):
                    sl.tensor.copy_(src)
                    torch.cuda.synchronize()
                    dist.barrier()
                    found = torch.zeros(1, dtype=torch.int32, device=dev)
                    sq = torch.zeros(1, device=dev)
                    C.allreduce_twoshot(sl.data_ptrs, sl.sig_ptrs, sl.mc_ptr, rank, sl.tensor, n, 1.0 / world,
                                        found, sq, algo == "multimem", 32, 20.0)
                    torch.cuda.synchronize()
                    err = ((sl.tensor.float() - ref).norm() / ref.norm()).item()
                    assert err < tol, (algo, dtype, n, err)
                    assert int(found.item()) == 0
                    tot = sq.clone()
                    dist.all_reduce(tot)
                    assert abs(tot.item() - (ref ** 2).sum().item()) / (ref ** 2).sum().item() < 2e-2
                # one-shot into a separate output
                sl.tensor.copy_(src)
                torch.cuda.synchronize()
                dist.barrier()
                out = torch.empty_like(src)
                C.allreduce_oneshot(sl.data_ptrs, sl.sig_ptrs, rank, out, n, 1.0 / world, None, None, 16, 20.0)
                torch.cuda.synchronize()
                err = ((out.float() - ref).norm() / ref.norm()).item()
                assert err < tol, ("oneshot", dtype, n, err)
                # bit-identical across ranks
                outs = [torch.empty_like(out) for _ in range(world)]
                dist.all_gather(outs, out)
                assert all(torch.equal(outs[0], o) for o in outs)
        # inf detection
        sl = pool.alloc(4096, torch.bfloat16)
        sl.tensor.fill_(1.0)
        if rank == world - 1:
            sl.tensor[77] = float("inf")
        found = torch.zeros(1, dtype=torch.int32, device=dev)
        torch.cuda.synchronize()
        dist.barrier()
        C.allreduce_twoshot(sl.data_ptrs, sl.sig_ptrs, 0, rank, sl.tensor, 4096, 1.0, found, None, False, 4, 20.0)
        torch.cuda.synchronize()
        f = found.clone()
        dist.all_reduce(f, op=dist.ReduceOp.MAX)
        assert int(f.item()) == 1
        # broadcast + scalar allgather
        sl = pool.alloc(1 << 20, torch.float32)
        sl.tensor.fill_(float(rank + 1))
        torch.cuda.synchronize()
        dist.barrier()
        C.comm_broadcast(sl.data_ptrs, sl.sig_ptrs, rank, world - 1, sl.tensor.numel() * 4, 16, 20.0, sl.tensor)
        torch.cuda.synchronize()
        assert torch.all(sl.tensor == float(world))
        inp = torch.tensor([rank + 0.5, 2.0 * rank], device=dev)
        out = torch.zeros(world * 2, device=dev)
        C.comm_allgather_scalars(sl.data_ptrs, sl.sig_ptrs, rank, inp, out, 20.0)
        torch.cuda.synchronize()
        exp = torch.tensor([v for r in range(world) for v in (r + 0.5, 2.0 * r)], device=dev)
        assert torch.equal(out, exp)
        assert pool.check_error() == 0

        # ---- engine end to end: 2 ranks with different data stay bit-identical and match 1-proc math
        from edl_b200.models import ResNetVd, to_train_dtype
        from edl_b200.trainer import StudentTrainer

        torch.manual_seed(0)
        m = to_train_dtype(ResNetVd(18, class_dim=16, width_mult=0.25), torch.bfloat16, dev).train()
        for use_graph in (False, True):
            tr = StudentTrainer(m, 8, image_shape=(3, 32, 32), num_classes=16, lr=0.05, use_graph=use_graph,
                                bucket_cap_mb=0.25)
            torch.manual_seed(100 + rank)
            x = torch.randn(8, 3, 32, 32).bfloat16().contiguous(memory_format=torch.channels_last).pin_memory()
            t = torch.softmax(torch.randn(8, 16), -1).bfloat16().pin_memory()
            for _ in range(4):
                loss = tr.step(x, t)
            torch.cuda.synchronize()
            flat = torch.cat([g.param.flatten().float() for g in tr.dp.flat.groups.values()])
            outs = [torch.empty_like(flat) for _ in range(world)]
            dist.all_gather(outs, flat)
            assert all(torch.equal(outs[0], o) for o in outs), "ranks diverged (graph=%s)" % use_graph
            assert torch.isfinite(flat).all()
            assert tr.dp.comm_launches > 0
            res["comm_launches_graph_%s" % use_graph] = tr.dp.comm_launches

        # ---- elastic resize without restarting the processes: 2 -> 1 (each rank alone) -> 2
        solo = [dist.new_group(ranks=[r]) for r in range(world)][rank]
        tr.rebuild(solo)
        assert tr.dp.world == 1
        for _ in range(2):
            tr.step(x, t)                      # different data per rank: replicas drift apart
        tr.rebuild(None)                       # back to the full group: new slab, new bucket plan
        assert tr.dp.world == world
        tr.sync_from(0)                        # "joiners" take params + momentum from rank 0
        for _ in range(2):
            tr.step(x, t)
        torch.cuda.synchronize()
        flat = torch.cat([g.param.flatten().float() for g in tr.dp.flat.groups.values()])
        outs = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(outs, flat)
        assert all(torch.equal(outs[0], o) for o in outs), "ranks diverged after the elastic resize"
        assert tr.dp.check_comm_error() == 0
        if rank == 0:
            q.put(("ok", res))
        dist.barrier()
        dist.destroy_process_group()
    except Exception as e:  # noqa
        import traceback

        q.put(("fail", "rank %d: %s\n%s" % (rank, e, traceback.format_exc())))
        raise


@pytest.mark.parametrize("world", [2, 4, 8])
def test_allreduce_kernels_and_engine(world):
    if torch.cuda.device_count() < world:
        pytest.skip("not enough GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    status, payload = q.get(timeout=540)
    for p in procs:
        p.join(60)
    assert status == "ok", payload
    print(payload)
