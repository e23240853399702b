"""NVSwitch-direct distillation link (needs 2 GPUs): student ships images into the teacher's HBM, the
teacher ships logits (+ row softmax stats) back, the student's fused loss kernel consumes them; loss
and gradient must match the plain PyTorch soft-CE on the same logits."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]


def _worker(rank, world, port, q):
    try:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
        from edl_b200.distill.device_feed import DeviceDistillLink, pool_bytes_needed
        from edl_b200.parallel.symm import SymmetricPool

        B, C, shape, T = 16, 1000, (3, 32, 32), 2.0
        pool = SymmetricPool(2 * pool_bytes_needed(B, shape, C, slots=2), device=dev)
        role = "student" if rank == 0 else "teacher"
        links = {f: DeviceDistillLink(pool, peer_rank=1 - rank, role=role, batch=B, image_shape=shape, num_classes=C,
                                      slots=2, temperature=T, timeout_s=20.0, fused_fc=f) for f in (False, True)}
        torch.manual_seed(7)
        proj = (torch.randn(3 * 32 * 32, C, device=dev) * 0.05).bfloat16()      # same "teacher" on both ranks
        proj_t = proj.t().contiguous()                                           # [C, K] for the fused GEMM
        for step in range(12):
            fused = step >= 6            # first the copy-kernel path, then GEMM -> peer ship
            link = links[fused]
            slot, seq = step % 2, step % 6 + 1
            if role == "student":
                torch.manual_seed(100 + step)
                x = torch.randn(B, *shape, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
                z = (torch.randn(B, C, device=dev) * 2).bfloat16().requires_grad_(True)
                link.send_images(x, slot, seq_imm=seq)
                loss = link.loss(z, slot, seq_imm=seq, kl=(step % 2 == 1))
                loss.backward()
                torch.cuda.synchronize()
                # reference: recompute the teacher logits locally with the same weights
                t_logits = (x.permute(0, 2, 3, 1).reshape(B, -1).float() @ proj.float()).bfloat16().float()
                p = F.softmax(t_logits / T, -1)
                zr = z.detach().float().requires_grad_(True)
                ref = -(p * F.log_softmax(zr, -1)).sum(-1).mean()
                if step % 2 == 1:
                    ref = ref + (p * torch.log(p.clamp_min(1e-30))).sum(-1).mean()
                ref.backward()
                assert abs(loss.item() - ref.item()) < 5e-3 * max(1.0, abs(ref.item())), (step, loss.item(), ref.item())
                rel = ((z.grad.float() - zr.grad).norm() / zr.grad.norm()).item()
                assert rel < 3e-2, (step, rel)
            else:
                img = link.wait_images(slot, seq_imm=seq)                          # NCHW view of NHWC memory
                feats = img.permute(0, 2, 3, 1).reshape(B, -1)
                if fused:
                    link.ship_linear(feats.contiguous(), proj_t, None, slot, seq_imm=seq)
                else:
                    logits = (feats.float() @ proj.float()).bfloat16()
                    link.send_logits(logits, slot, seq_imm=seq)
                torch.cuda.synchronize()
            assert link.check_error() == 0, (step, fused)
        dist.barrier()
        if rank == 0:
            q.put(("ok", ""))
        dist.destroy_process_group()
    except Exception as e:  # noqa
        import traceback
        q.put(("fail", "rank %d: %s\n%s" % (rank, e, traceback.format_exc())))
        raise


def test_device_distill_link():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29800 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    status, payload = q.get(timeout=300)
    for p in procs:
        p.join(60)
    assert status == "ok", payload
