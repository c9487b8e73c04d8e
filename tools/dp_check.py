"""Data-parallel correctness check for 2 ranks (torchrun, NCCL): the overlapped, bucketed gradient all-reduce must give
bit-identical parameters to a single all-reduce after backward (at world size 2 a sum is order-independent), the ranks
must stay in lock step, and overwrite-mode gradients must survive their first-step verification.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/dp_check.py
"""
import copy
import io
import os
import sys
from contextlib import redirect_stdout
from dataclasses import make_dataclass

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vyomai_b200 import VisionLanguageModel, Vit  # noqa: E402
from vyomai_b200.trainer import Trainer, caption_labels  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    T = make_dataclass("T", [("hidden_size", int, 256), ("num_attention_heads", int, 4), ("num_key_value_heads", int, 2),
                             ("max_position_embeddings", int, 64), ("num_hidden_layers", int, 3), ("vocab_size", int, 5003),
                             ("hidden_dropout_prob", float, 0.0), ("layer_norm_eps", float, 1e-5), ("hidden_act", str, "gelu")])
    Vc = make_dataclass("Vc", [("hidden_size", int, 256), ("num_attention_heads", int, 4), ("image_size", tuple, (32, 32)),
                               ("patch_size", tuple, (8, 8)), ("num_channels", int, 4), ("num_hidden_layers", int, 2),
                               ("hidden_dropout_prob", float, 0.0), ("layer_norm_eps", float, 1e-5), ("hidden_act", str, "gelu")])
    torch.manual_seed(0)
    with redirect_stdout(io.StringIO()):
        base = VisionLanguageModel(T(), encoder=Vit(Vc()), pos_embedding_type="rope", attention_type="gqa")
    base = base.to(dev).to(torch.bfloat16).train()
    models = [base, copy.deepcopy(base), copy.deepcopy(base), copy.deepcopy(base)]
    w_init = None
    # tiny buckets (64 KB) so that many boundaries fall inside the model; overlap on / off
    trainers = [Trainer(models[0], lr=1e-3, weight_decay=0.01, max_grad_norm=1.0, use_graph=False, bucket_mb=0.0625, overlap=True, dp_mode="nccl"),
                Trainer(models[1], lr=1e-3, weight_decay=0.01, max_grad_norm=1.0, use_graph=False, bucket_mb=0.0625, overlap=False, dp_mode="nccl"),
                Trainer(models[2], lr=1e-3, weight_decay=0.01, max_grad_norm=1.0, use_graph=False, bucket_mb=0.0625, overlap=False, dp_mode="nccl"),
                # the peer-memory sharded step (dp_shard.py): no NCCL in the step, fp32 gradient sum
                Trainer(models[3], lr=1e-3, weight_decay=0.01, max_grad_norm=1.0, use_graph=False, dp_mode="p2p")]
    assert trainers[3].dp_mode == "p2p", f"symmetric memory unavailable: dp_mode {trainers[3].dp_mode}"
    w_init = trainers[1].fp.flat.float().clone()
    g = torch.Generator().manual_seed(100 + rank)
    ok = True
    for step in range(3):
        px = torch.rand(4, 4, 32, 32, generator=g).to(dev)
        ids = torch.randint(3, 5003, (4, 12), generator=g).to(dev)
        mask = torch.ones(4, 12, dtype=torch.long)
        mask[1, 8:] = 0
        mask = mask.to(dev)
        labels = caption_labels(ids, mask)
        losses = [float(t.caption_step(px, ids, mask, labels)) for t in trainers]
        d_ab = float((trainers[0].fp.flat.float() - trainers[1].fp.flat.float()).abs().max())
        d_bc = float((trainers[1].fp.flat.float() - trainers[2].fp.flat.float()).abs().max())  # run-to-run noise (atomics)
        n_ab = int((trainers[0].fp.flat != trainers[1].fp.flat).sum())
        n_bc = int((trainers[1].fp.flat != trainers[2].fp.flat).sum())
        same = d_ab <= 2.0 * d_bc + 1e-6 and n_ab <= 2 * n_bc + 16
        if step == 0:
            ga, gb = trainers[0].fp.grad.float(), trainers[1].fp.grad.float()
            idx = (ga != gb).nonzero().flatten()
            if idx.numel():
                i0, i1 = int(idx[0]), int(idx[-1])
                sel = idx[:200000]
                ratio = (ga[sel] / gb[sel]).nan_to_num(0.0)
                print(f"   [rank {rank}] gradients differ in {idx.numel()} elements, flat range [{i0}, {i1}]; ratio overlap/plain: "
                      f"median {float(ratio.median()):.4f} min {float(ratio.min()):.4f} max {float(ratio.max()):.4f}; "
                      f"|overlap| {float(ga[sel].abs().mean()):.3e} |plain| {float(gb[sel].abs().mean()):.3e}", flush=True)
            else:
                print(f"   [rank {rank}] gradients identical", flush=True)
        if rank == 0 and step == 0:
            names = {id(q): n for n, q in models[0].named_parameters()}
            for qa, qb, off in zip(trainers[0].fp.params, trainers[1].fp.params, trainers[0].fp.offsets):
                nd = int((qa != qb).sum())
                if nd:
                    print(f"   differs: {names[id(qa)]} offset {off} numel {qa.numel()} bucket {trainers[0].exchange.bucket_of(off)}: {nd} elements, "
                          f"max {float((qa.float() - qb.float()).abs().max()):.3e}", flush=True)
            print("   buckets:", trainers[0].exchange.buckets[:8], "...", flush=True)
        ref = trainers[0].fp.flat.clone()
        dist.broadcast(ref, src=0)
        lock = torch.equal(ref, trainers[0].fp.flat)
        if rank == 1:
            print(f"   [rank 1] step {step}: parameters equal to rank 0's: {lock}", flush=True)
        # sharded peer-memory step: replicas bit-identical; against the NCCL path the only difference is the gradient sum's
        # rounding (fp32 here, bf16 inside NCCL), so the UPDATE so far must agree closely
        ref3 = trainers[3].fp.flat.clone()
        dist.broadcast(ref3, src=0)
        lock3 = torch.equal(ref3, trainers[3].fp.flat)
        trainers[3].shard.check()
        up_n, up_p = trainers[1].fp.flat.float() - w_init, trainers[3].fp.flat.float() - w_init
        rel = float((up_p - up_n).norm() / up_n.norm())
        cos = float(torch.nn.functional.cosine_similarity(up_p, up_n, dim=0))
        p2p_ok = lock3 and cos > 0.98 and abs(losses[3] - losses[1]) < 0.02
        if rank == 0:
            print(f"step {step}: peer-memory sharded step: loss {losses[3]:.5f}; replicas bit-identical: {lock3}; update vs NCCL path: "
                  f"rel l2 {rel:.3e}, cosine {cos:.5f}", flush=True)
        ok &= same and lock and p2p_ok
        if rank == 0:
            print(f"step {step}: loss overlap {losses[0]:.5f} / single all-reduce {losses[1]:.5f}; parameters identical to the "
                  f"non-overlapped run: {same} (max diff {d_ab:.3e} in {n_ab} elements; two non-overlapped runs differ by {d_bc:.3e} in {n_bc}); ranks in lock step: {lock}; buckets {len(trainers[0].exchange.buckets)}; "
                  f"grad_overwrite {trainers[0].grad_overwrite}/{trainers[1].grad_overwrite}", flush=True)
    ok &= trainers[0].grad_overwrite and trainers[1].grad_overwrite
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("DP CHECK", "PASS" if int(flag) == 1 else "FAIL", flush=True)
    dist.barrier()
    torch.cuda.synchronize()
    os._exit(0 if int(flag) == 1 else 1)


if __name__ == "__main__":
    main()
