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

        res["pool"] = pool.describe()
        # ---- fused reduce-scatter -> SGD-momentum -> parameter all-gather kernel against a single-process fp32
        #      reference of the same update (mean gradient over the ranks, then the plain momentum rule)
        n = 3_000_000 // 8 * 8
        gsl, psl = pool.alloc(n, torch.bfloat16), pool.alloc(n, torch.bfloat16)
        for algo in (["twoshot", "multimem"] if pool.has_multicast else ["twoshot"]):
            torch.manual_seed(7)
            master = torch.randn(n, device=dev)
            mom = torch.randn(n, device=dev) * 0.1
            torch.manual_seed(50 + rank)
            grad = (torch.randn(n, device=dev) * 0.5).bfloat16()
            gathered = [torch.empty_like(grad) for _ in range(world)]
            dist.all_gather(gathered, grad)
            gref = sum(g.float() for g in gathered) / world
            mref = 0.9 * mom + (gref + 1e-4 * master)
            wref = master - 0.1 * mref
            gsl.tensor.copy_(grad)
            psl.tensor.copy_(master.bfloat16())
            lr = torch.tensor([0.1], device=dev)
            torch.cuda.synchronize()
            dist.barrier()
            C.allreduce_sgd(gsl.data_ptrs, gsl.sig_ptrs, gsl.mc_ptr, psl.data_ptrs, psl.mc_ptr, rank, master, mom, None,
                            lr, 1.0 / world, None, None, None, 0.9, 1e-4, False, algo == "multimem", 32, 20.0)
            torch.cuda.synchronize()
            dist.barrier()
            # every rank holds the new bf16 parameters of EVERY slice; master / momentum only of its own slice
            err = ((psl.tensor.float() - wref).abs().max() / wref.abs().max()).item()
            assert err < 1e-2, ("fused params", algo, err)
            nvec = n // 8
            cap = -(-nvec // world)
            lo, hi = min(nvec, cap * rank) * 8, min(nvec, cap * (rank + 1)) * 8
            # P2P: fp32 sum of the bf16 gradients goes straight into the update; NVLS: the switch hands back the
            # fp32-accumulated sum rounded to bf16 (what the unfused path stores, too)
            tol = 1e-5 if algo == "twoshot" else 4e-3
            assert torch.allclose(master[lo:hi], wref[lo:hi], rtol=tol, atol=tol), ("fused master", algo)
            assert torch.allclose(mom[lo:hi], mref[lo:hi], rtol=tol, atol=tol), ("fused momentum", algo)
            outs = [torch.empty_like(psl.tensor) for _ in range(world)]
            dist.all_gather(outs, psl.tensor)
            assert all(torch.equal(outs[0], o) for o in outs), "parameter shadows differ between ranks"

        # ---- engine end to end, against ONE process doing the same global batch in fp32 arithmetic on the reduction
        import copy

        from edl_b200.models import ResNetVd, to_train_dtype
        from edl_b200.trainer import StudentTrainer

        torch.manual_seed(0)
        m0 = to_train_dtype(ResNetVd(18, class_dim=16, width_mult=0.25), torch.bfloat16, dev).train()
        torch.manual_seed(100 + rank)
        x = torch.randn(8, 3, 32, 32).bfloat16().contiguous(memory_format=torch.channels_last).pin_memory()
        t = torch.softmax(torch.randn(8, 16), -1).bfloat16().pin_memory()

        def flat_params(tr):
            return torch.cat([g.param.flatten().float() for g in tr.dp.flat.groups.values()])

        def flat_grads(tr):
            return torch.cat([g.grad.flatten().float() for g in tr.dp.flat.groups.values()])

        # reference gradient of step 1: every rank runs its own batch WITHOUT communication, fp32 mean over ranks
        solo_group = [dist.new_group(ranks=[r]) for r in range(world)][rank]
        ref_tr = StudentTrainer(copy.deepcopy(m0), 8, image_shape=(3, 32, 32), num_classes=16, lr=0.0, use_graph=False,
                                group=solo_group, fused_optimizer=False, weight_decay=0.0)
        ref_tr.step(x, t, sync=True)
        torch.cuda.synchronize()
        g_local = flat_grads(ref_tr)
        g_all = [torch.empty_like(g_local) for _ in range(world)]
        dist.all_gather(g_all, g_local)
        g_ref = sum(g_all) / world
        for fused in (False, True):
            tr = StudentTrainer(copy.deepcopy(m0), 8, image_shape=(3, 32, 32), num_classes=16, lr=0.0, use_graph=False,
                                bucket_cap_mb=0.25, fused_optimizer=fused, weight_decay=0.0)
            assert tr.dp.bucket_opt == fused
            tr.step(x, t, sync=True)
            torch.cuda.synchronize()
            if not fused:       # the reduced gradient itself (the fused kernel never writes it back)
                err = ((flat_grads(tr) - g_ref).norm() / g_ref.norm()).item()
                assert err < 2e-2, ("engine gradient vs fp32 single-process mean", err)
                res["engine_grad_err"] = err
            res["algos_fused_%s" % fused] = sorted({a for a, _, _ in tr.dp.last_algos})
        params = {}
        for use_graph, fused in ((False, True), (True, True), (True, False), (True, "again")):
            again, fused = fused == "again", fused is True
            tr = StudentTrainer(copy.deepcopy(m0), 8, image_shape=(3, 32, 32), num_classes=16, lr=0.05,
                                use_graph=use_graph, bucket_cap_mb=0.25, fused_optimizer=fused)
            for _ in range(4):
                loss = tr.step(x, t)
            float(loss)
            torch.cuda.synchronize()
            flat = flat_params(tr)
            outs = [torch.empty_like(flat) for _ in range(world)]
            dist.all_gather(outs, flat)
            assert all(torch.equal(outs[0], o) for o in outs), "ranks diverged (graph=%s fused=%s)" % (use_graph, fused)
            assert torch.isfinite(flat).all()
            assert tr.dp.comm_launches > 0
            res["comm_launches_graph_%s_fused_%s" % (use_graph, fused)] = tr.dp.comm_launches
            params["again" if again else (use_graph, fused)] = flat
        # the fused and the unfused engines follow the same trajectory (fp32 vs bf16-rounded reduced gradient); the
        # yardstick is the run-to-run drift of the unfused engine itself (float atomics in the BatchNorm statistics)
        drift = ((params[(True, True)] - params[(True, False)]).norm() / params[(True, False)].norm()).item()
        noise = ((params["again"] - params[(True, False)]).norm() / params[(True, False)].norm()).item()
        assert drift < max(2e-2, 3 * noise), ("fused vs unfused parameters after 4 steps", drift, noise)
        res["fused_vs_unfused_drift"] = [drift, noise]
        # sharded optimizer state -> complete again (checkpoint / planned rescale)
        tr = StudentTrainer(copy.deepcopy(m0), 8, image_shape=(3, 32, 32), num_classes=16, lr=0.05, use_graph=False,
                            bucket_cap_mb=0.25, fused_optimizer=True)
        for _ in range(3):
            tr.step(x, t)
        torch.cuda.synchronize()
        try:
            tr.state_dict()
            raise AssertionError("state_dict() of a sharded optimizer state must refuse")
        except RuntimeError:
            pass
        tr.consolidate()
        for dt, g in tr.dp.flat.groups.items():
            if g.master is not None:
                assert torch.equal(g.master.bfloat16(), g.param), "consolidated masters do not match the parameters"
                outs = [torch.empty_like(tr.opt.state[dt]["mom"]) for _ in range(world)]
                dist.all_gather(outs, tr.opt.state[dt]["mom"])
                assert all(torch.equal(outs[0], o) for o in outs), "momentum differs between ranks after consolidate"
        tr.state_dict()

        # ---- global-norm clipping against the norm of the fp32 reference gradient
        trc = StudentTrainer(copy.deepcopy(m0), 8, image_shape=(3, 32, 32), num_classes=16, lr=0.0, use_graph=False,
                             bucket_cap_mb=0.25, clip_norm=0.05, weight_decay=0.0)
        assert not trc.dp.bucket_opt
        trc.step(x, t, sync=True)
        torch.cuda.synchronize()
        want = float(g_ref.norm())
        got = float(trc.dp.grad_norm_t.item())
        assert abs(got - want) / want < 2e-2, ("global gradient norm", got, want)
        scale = float(trc.dp.clip_scale_t.item())
        assert abs(scale - min(1.0, 0.05 / (want + 1e-6))) < 2e-2 * max(scale, 1e-3), (scale, want)
        res["clip"] = {"norm": got, "ref_norm": want, "scale": scale}

        # ---- elastic resize without restarting the processes: 2 -> 1 (each rank alone) -> 2
        tr.prepare_rescale()                   # collective over the old stage: sharded optimizer state made complete
        tr.rebuild(solo_group)
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


@pytest.mark.slow
@pytest.mark.parametrize("leave", ["scale_in", "kill"])
def test_inplace_rescale_through_the_launcher_on_gpus(tmp_path, leave):
    """BASELINE config 4 in miniature: ResNet50_vd, pod A = GPU 0, pod B = GPU 1, the real launcher, in-place mode with the
    fabric backend (no process group, no NCCL): B joins, then leaves by the leader's ScaleIn RPC or dies by SIGKILL; the
    survivor's trainer must keep its process both times (hot recovery in the SIGKILL case)."""
    import json
    import subprocess
    import sys

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = str(tmp_path / "elastic.json")
    env = dict(os.environ, EDL_COMM_TIMEOUT="5")
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "bench_elastic_launch.py"), "--native-store",
                        "--trainer", "resnet", "--gpus-per-pod", "1", "--leave", leave, "--modes", "inplace", "--out", out],
                       env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    run = json.load(open(out))["runs"][0]
    assert "error" not in run, run
    assert run["survivor_process_kept"] is True, run
    assert run["leave_stall_s"] < (15.0 if leave == "kill" else 5.0), run
