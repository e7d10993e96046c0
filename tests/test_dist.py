"""world_size-2 gloo test (CPU) of the batch-shard + single all-reduce plumbing: two ranks each take
half the batch, compute their local gradients with the CPU oracle, all-reduce [grad_weight || grad_bias]
through GradBucket, and must reproduce the full-batch oracle; per-sample gradients must equal the
corresponding slices of the full-batch result."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _worker(rank, world, port, ret):
    for p in (os.path.join(ROOT, "modulated-deform-conv_b200"), ROOT, HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle
        from cases import Case, make_inputs, rel_err
        import mdconv_dist
        case = Case("dist", 2, True, B=5, C=4, Co=6, size=(6, 5), dg=2, seed=77)
        t = make_inputs(case, torch.float64)
        a = (case.s, case.p, case.d, case.g, case.dg)
        full = oracle.backward(t["x"], t["weight"], t["offset"], t["mask"], t["grad_output"], *a, variant=0)
        xs, offs, ms, gos = mdconv_dist.shard_batch([t["x"], t["offset"], t["mask"], t["grad_output"]], rank, world)
        lo, hi = mdconv_dist.shard_range(case.B, rank, world)
        loc = oracle.backward(xs, t["weight"], offs, ms, gos, *a, variant=0)
        w = t["weight"].clone().requires_grad_()
        b = t["bias"].clone().requires_grad_()
        w.grad = torch.from_numpy(loc["grad_weight"]).clone()
        b.grad = torch.from_numpy(loc["grad_bias"]).clone()
        mdconv_dist.all_reduce_param_grads([w, b])
        ok = rel_err(w.grad.numpy(), full["grad_weight"]) < 1e-6 and rel_err(b.grad.numpy(), full["grad_bias"]) < 1e-6
        # the in-backward reducer (GradSync) on the same local gradients: CPU tensors -> plain all-reduce, in place
        gw2 = torch.from_numpy(loc["grad_weight"]).clone()
        gb2 = torch.from_numpy(loc["grad_bias"]).clone()
        sync = mdconv_dist.GradSync(None, torch.device("cpu"))
        ok = ok and sync.active()
        sync.reduce([gw2, gb2, None])
        ok = ok and rel_err(gw2.numpy(), full["grad_weight"]) < 1e-12 and rel_err(gb2.numpy(), full["grad_bias"]) < 1e-12
        ok = ok and rel_err(loc["grad_input"], full["grad_input"][lo:hi]) < 1e-12
        ok = ok and rel_err(loc["grad_offset"], full["grad_offset"][lo:hi]) < 1e-12
        ok = ok and rel_err(loc["grad_mask"], full["grad_mask"][lo:hi]) < 1e-12
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_shard_range_partitions_exactly():
    sys.path.insert(0, os.path.join(ROOT, "modulated-deform-conv_b200"))
    import mdconv_dist
    for n in (1, 5, 32, 256):
        for world in (1, 2, 3, 8):
            spans = [mdconv_dist.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
    with pytest.raises(ValueError):
        mdconv_dist.shard_range(4, 2, 2)


def test_two_rank_gloo_allreduce_matches_full_batch():
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}


def _worker_attach(rank, world, port, ret):
    for p in (os.path.join(ROOT, "modulated-deform-conv_b200"), ROOT, HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import mdconv_dist
        import modulated_deform_conv as mdc
        torch.manual_seed(0)
        pack = mdc.ModulatedDeformConv2dPack(4, 4, 3, padding=1, bias=True)
        sync = mdconv_dist.attach(pack)
        ok = sync.active() and len(pack._grad_sync_handles) == 4     # conv_offset.{weight,bias}, conv_mask.{weight,bias}
        # the producers' gradients are all-reduced when autograd accumulates them (sum over ranks)
        x = torch.full((1, 4, 5, 5), float(rank + 1))
        (pack.conv_offset(x).sum() + pack.conv_mask(x).sum()).backward()
        ref = mdc.ModulatedDeformConv2dPack(4, 4, 3, padding=1, bias=True)
        ref.load_state_dict(pack.state_dict())
        xs = torch.cat([torch.full((1, 4, 5, 5), float(r + 1)) for r in range(world)])
        (ref.conv_offset(xs).sum() + ref.conv_mask(xs).sum()).backward()
        for a, b in ((pack.conv_offset.weight, ref.conv_offset.weight), (pack.conv_mask.bias, ref.conv_mask.bias)):
            ok = ok and torch.allclose(a.grad, b.grad, rtol=1e-5, atol=1e-6)
        # average=True is DDP's convention
        mdconv_dist.detach(pack)
        ok = ok and pack._grad_sync is None
        # prepare_ddp: deformable-conv parameters go on DDP's ignore list, everything else stays with DDP
        model = torch.nn.Sequential(torch.nn.Conv2d(4, 4, 3, padding=1), mdc.ModulatedDeformConv2dPack(4, 4, 3, padding=1))
        ignored = mdconv_dist.prepare_ddp(model)
        ok = ok and set(ignored) == {"1.weight", "1.conv_offset.weight", "1.conv_offset.bias", "1.conv_mask.weight",
                                     "1.conv_mask.bias"}
        ok = ok and model[1]._grad_sync.average
        ddp = torch.nn.parallel.DistributedDataParallel(model)
        ok = ok and set(ddp.parameters_to_ignore) == set(ignored)
        # GradSync on a backward_ex-style dict with the fp32 flat buffer: exchange fp32, round once
        gw = torch.full((4, 4, 3, 3), float(rank + 1), dtype=torch.float16)
        gb = torch.full((4,), 0.5 * (rank + 1), dtype=torch.float16)
        flat = torch.cat([gw.float().reshape(-1) + 1e-4, gb.float()])
        s2 = mdconv_dist.GradSync(None, torch.device("cpu"))
        g = {"grad_weight": gw, "grad_bias": gb, "flat_f32": flat}
        s2.reduce(g, True)          # CPU tensors: in-place all-reduce in the gradients' own dtype
        ok = ok and float(gw[0, 0, 0, 0]) == 3.0 and float(gb[0]) == 1.5
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_attach_reduces_pack_producers_and_prepare_ddp_ignores_them():
    world = 2
    port = 31500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker_attach, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}
