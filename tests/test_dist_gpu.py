"""2-rank NCCL test (-m gpu, needs two GPUs; skipped otherwise) of the in-backward overlapped reduction:
each rank runs the module on its half of the batch with mdconv_dist.attach; the reduced parameter gradients must
equal the gradients of ONE process on the full batch (fp16: bit-equal up to one rounding, because the ranks exchange
the fp32 values before rounding), the per-sample gradients must equal the corresponding slices, and a Pack module's
producer gradients must be reduced too."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _worker(rank, world, port, ret):
    for p in (os.path.join(ROOT, "modulated-deform-conv_b200"), ROOT, HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import mdconv_dist
        import modulated_deform_conv as mdc
        torch.manual_seed(0)
        B, C, Co, H = 4, 64, 64, 16
        x = torch.randn(B, C, H, H).half()
        off = torch.randn(B, 18, H, H).half()
        msk = torch.sigmoid(torch.randn(B, 9, H, H)).half()
        go = torch.randn(B, Co, H, H).half()
        mod = mdc.ModulatedDeformConv2d(C, Co, 3, padding=1, bias=True).half().to(dev)

        def run(xb, ob, mb, gb):
            lv = [t.to(dev).requires_grad_() for t in (xb, ob, mb)]
            mod.weight.grad = None
            mod.bias.grad = None
            mod(*lv).backward(gb.to(dev))
            torch.cuda.synchronize()
            return [t.grad.clone() for t in lv], mod.weight.grad.clone(), mod.bias.grad.clone()

        full_data, full_w, full_b = run(x, off, msk, go)                  # one process, whole batch
        sync = mdconv_dist.attach(mod, max_ctas=4)
        lo, hi = mdconv_dist.shard_range(B, rank, world)
        data, w, b = run(x[lo:hi], off[lo:hi], msk[lo:hi], go[lo:hi])     # sharded, reduced inside the backward
        ok = sync.active()
        rel = lambda a, r: float((a.float() - r.float()).abs().max() / r.float().abs().max())
        # fp32 exchange + one rounding: at most one fp16 ulp away from the single-process result
        ok = ok and rel(w, full_w) < 1.0e-3 and rel(b, full_b) < 1.0e-3
        for a, r in zip(data, full_data):
            ok = ok and torch.equal(a, r[lo:hi])
        # Pack: the producers' gradients are reduced by the post-accumulate hooks
        pack = mdc.ModulatedDeformConv2dPack(C, Co, 3, padding=1).half().to(dev)
        xs = x.to(dev)
        pack(xs).float().sum().backward()
        ref_off = pack.conv_offset.weight.grad.clone()
        ref_w = pack.weight.grad.clone()
        for p_ in pack.parameters():
            p_.grad = None
        mdconv_dist.attach(pack, max_ctas=4)
        pack(xs[lo:hi]).float().sum().backward()
        torch.cuda.synchronize()
        ok = ok and rel(pack.conv_offset.weight.grad, ref_off) < 2e-2 and rel(pack.weight.grad, ref_w) < 2e-2
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_two_rank_nccl_in_backward_reduction_matches_full_batch():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    world = 2
    port = 33500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}
