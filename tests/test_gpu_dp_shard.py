"""The peer-memory data-parallel optimizer step (csrc/dp_step.cu, vyomai_b200/dp_shard.py) on ONE GPU: the "ranks" are separate
buffers on the same device, so the kernels' peer pointer tables are exercised without NVLink (tools/dp_check.py runs the
real two-process version under torchrun). Reference for the arithmetic: torch.optim.AdamW + clip_grad_norm_ on the mean
gradient — what accelerate's DDP + clip + AdamW does in Examples/vyom-ai-accelerate-multimodel-2t4.ipynb cell 1."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ranks(world, n, dtype, seed):
    from vyomai_b200 import dp_shard
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(seed)
    w0 = (0.05 * torch.randn(n, generator=g)).to(dtype)
    flats = [w0.clone().to(dev) for _ in range(world)]
    grads = [torch.zeros(n, dtype=dtype, device=dev) for _ in range(world)]
    flags = [torch.zeros(64, dtype=torch.int32, device=dev) for _ in range(world)]
    scal = [torch.zeros(64, dtype=torch.float32, device=dev) for _ in range(world)]
    hyper = dict(lr=1e-2, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01, max_grad_norm=1.0)
    steps = [dp_shard.ShardedStep(world, r, flats[r], [t.data_ptr() for t in flats], grads[r], [t.data_ptr() for t in grads],
                                  flags[r], [t.data_ptr() for t in flags], scal[r], [t.data_ptr() for t in scal], **hyper)
             for r in range(world)]
    return w0, flats, grads, steps, hyper, g


@pytest.mark.parametrize("world,dtype,n", [(1, torch.bfloat16, 8 * 1000), (2, torch.bfloat16, 8 * 12345), (3, torch.float32, 8 * 4097),
                                           (8, torch.bfloat16, 8 * 50001)])
def test_sharded_step_matches_torch_adamw_on_the_mean_gradient(world, dtype, n):
    w0, flats, grads, steps, hyper, g = _ranks(world, n, dtype, 7)
    ref_w = torch.nn.Parameter(w0.float().cuda())
    opt = torch.optim.AdamW([ref_w], lr=hyper["lr"], betas=hyper["betas"], eps=hyper["eps"], weight_decay=hyper["weight_decay"])
    step_dev = torch.zeros(1, dtype=torch.int32, device="cuda")
    for it in range(1, 4):
        scale = 3.0 if it == 1 else 0.01  # step 1 clips (norm > 1), the later ones do not
        gs = [(scale * torch.randn(n, generator=g)).to(dtype) for _ in range(world)]
        for r in range(world):
            grads[r].copy_(gs[r])
        step_dev.add_(1)
        # phase by phase in stream order (what the device barriers order between real ranks)
        for s in steps:
            s.reduce()
        for s in steps:
            s.adamw(it, step_dev if it % 2 else None)
        torch.cuda.synchronize()
        mean = torch.stack([x.float() for x in gs]).sum(0).cuda() / world
        ref_w.grad = mean.clone()
        torch.nn.utils.clip_grad_norm_([ref_w], hyper["max_grad_norm"])
        opt.step()
        for r in range(1, world):
            assert torch.equal(flats[r], flats[0]), "replicas must hold bit-identical parameters"
        got = torch.cat([s.master if s.master is not None else flats[0][s.lo:s.hi] for s in steps]).float()
        assert got.numel() == n
        err = float((got - ref_w.detach()).abs().max())
        assert err < 2e-6, f"step {it}: fp32 weights differ from torch AdamW by {err}"
        if dtype == torch.bfloat16:
            assert torch.equal(flats[0], got.to(torch.bfloat16)), "bf16 parameters are the rounded fp32 master shard"
        # the published partial norms: every rank sees the same world values, their sum is |sum of gradients|^2
        tot = float(torch.stack([x.float() for x in gs]).sum(0).double().pow(2).sum())
        for s in steps:
            assert torch.equal(s.scalars[:world], steps[0].scalars[:world])
        assert abs(float(steps[0].scalars[:world].double().sum()) - tot) <= 1e-4 * tot


def test_device_barrier_single_rank_and_two_ranks_on_two_streams():
    _, flats, grads, steps, _, _ = _ranks(1, 64, torch.bfloat16, 1)
    for _ in range(5):
        steps[0].barrier()
    torch.cuda.synchronize()
    assert int(steps[0].epoch) == 5 and int(steps[0].error) == 0 and int(steps[0].flags[0]) == 5
    # two ranks as two streams of one device: each barrier kernel is one 32-thread CTA, so both are resident and the spin ends
    _, flats, grads, steps, _, _ = _ranks(2, 64, torch.bfloat16, 2)
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    for k in range(3):
        for r in (0, 1):
            with torch.cuda.stream(streams[r]):
                steps[r].barrier()
    torch.cuda.synchronize()
    for r in (0, 1):
        steps[r].check()
        assert int(steps[r].epoch) == 3 and steps[r].flags[:2].tolist() == [3, 3]


def test_sharded_step_argument_checks():
    from vyomai_b200 import _lib, dp_shard
    with pytest.raises(_lib.VyomError):
        _ranks(9, 8 * 100, torch.bfloat16, 3)[3][0].barrier()  # world > 8
    assert dp_shard.shard_bounds(8 * 10, 3, 0) == (0, 32) and dp_shard.shard_bounds(8 * 10, 3, 2) == (64, 80)
