"""torchrun / single-process entry: checks GradientSynchronizer (deqsci_b200/distributed.py, csrc/optim.cu) --
flat gradients + ONE fused kernel = cross-GPU one-shot all-reduce over NVLink peer memory + 1/world scale + Adam
-- against the plain recipe: all_gather the per-rank gradients, average, torch.optim.Adam.step().

    python scripts/fused_adam_check.py                                  # one GPU: fused scale + Adam
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 scripts/fused_adam_check.py
Prints FUSED_ADAM_OK <mode> on success (rank 0)."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
from deqsci_b200.distributed import GradientSynchronizer  # noqa: E402

SHAPES = [(64, 5, 3, 3), (64,), (64, 64, 3, 3), (7,), (4, 64, 3, 3), (1, 3)]       # odd sizes: padding between slices
modes = os.environ.get("DEQSCI_CHECK_MODES", "p2p,nccl").split(",") if world > 1 else ["single"]
ok = True
for mode in modes:
    torch.manual_seed(3)                                   # same initial parameters on every rank
    ps = [torch.nn.Parameter(torch.randn(s, device=dev) * 0.1) for s in SHAPES]
    ref = [p.detach().clone().requires_grad_() for p in ps]
    opt = torch.optim.Adam(ref, lr=1e-3)
    opt_adopted = torch.optim.Adam(ps, lr=1e-3)
    sync = GradientSynchronizer.adopt(opt_adopted) if mode != "nccl" else GradientSynchronizer(ps, lr=1e-3, p2p=False)
    assert sync is not None
    if world > 1 and mode == "p2p" and sync.mode != "p2p":
        if rank == 0:
            print("peer mapping unavailable on this box (mode %s): p2p leg skipped" % sync.mode)
        continue
    g = torch.Generator(device="cpu").manual_seed(100 + rank)   # different gradients per rank
    ms = []
    for it in range(4):
        sync.zero_grad()
        x = [torch.randn(s, generator=g).to(dev) for s in SHAPES]
        loss = sum((p * xi).sum() + 0.5 * (p * p).sum() for p, xi in zip(ps, x))
        loss.backward()                                    # accumulates into the flat buffer in place
        assert all(p.grad.data_ptr() == sync.flat_g.data_ptr() + 4 * o for p, o in zip(ps, sync.offsets))
        local_g = [p.grad.detach().clone() for p in ps]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        sync.step()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
        for r_, lg in zip(ref, local_g):                   # the plain recipe
            if world > 1:
                parts = [torch.empty_like(lg) for _ in range(world)]
                dist.all_gather(parts, lg)
                lg = torch.stack(parts).mean(0)
            r_.grad = lg
        opt.step()
        err = max(float((p.detach() - r_.detach()).abs().max() / r_.detach().abs().max()) for p, r_ in zip(ps, ref))
        ok = ok and err < 2e-6
    sync.check_barrier()
    st = opt_adopted.state_dict() if mode != "nccl" else None
    if st is not None:                                     # the adopted optimizer's state is the live moment buffers
        e = st["state"][0]
        ok = ok and float(e["step"]) == 4.0 and float((e["exp_avg"] - opt.state[ref[0]]["exp_avg"]).abs().max()) < 1e-6
    if world > 1:                                          # every rank must hold identical parameters (bitwise)
        flat = sync.flat_p.clone()
        parts = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(parts, flat)
        ok = ok and all(torch.equal(parts[0], q) for q in parts)
    if rank == 0:
        print("mode %s (%s): max rel param err vs all-gather-mean + torch Adam %.2e, step %.3f ms (first %.3f)" % (
            mode, sync.mode, err, min(ms), ms[0]))
    sync.close()
if world > 1:
    t = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    ok = bool(t.item())
    dist.destroy_process_group()
if rank == 0:
    print("FUSED_ADAM_OK" if ok else "FUSED_ADAM_BAD")
