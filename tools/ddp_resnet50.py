#!/usr/bin/env python
"""BASELINE config C3: torch DDP ResNet-50 step launched through the kt distributed launcher.

    python tools/ddp_resnet50.py --gpus 8            # kt.Compute(gpus=8).distribute("pytorch", num_proc=8)
    torchrun --nproc-per-node 8 tools/ddp_resnet50.py --direct   # the honest comparator: same function, plain torchrun

The framework only launches ranks with the reference's env contract (MASTER_ADDR/PORT, RANK, WORLD_SIZE,
LOCAL_RANK — kt/serving/spmd/pytorch_process.py:18-29); DDP's NCCL all-reduce over NVLink is the user's.
Synthetic 224x224 batches, random-init torchvision resnet50, SGD(0.1, momentum 0.9), bf16 autocast,
channels_last."""
import argparse
import json
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def train_resnet50(steps: int = 20, warmup: int = 5, batch: int = 256):
    """One rank of the DDP job. Returns {"rank", "images_per_sec" (this rank), "loss", "param_checksum"}."""
    import torch
    import torch.distributed as dist
    import torchvision
    from torch.nn.parallel import DistributedDataParallel as DDP

    local_rank = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    model = torchvision.models.resnet50(weights=None).to(dev).to(memory_format=torch.channels_last)
    model = DDP(model, device_ids=[local_rank])
    opt = torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9)
    loss_fn = torch.nn.CrossEntropyLoss()
    g = torch.Generator(device=dev).manual_seed(1234 + dist.get_rank())
    x = torch.randn(batch, 3, 224, 224, device=dev, generator=g).contiguous(memory_format=torch.channels_last)
    y = torch.randint(0, 1000, (batch,), device=dev, generator=g)

    def step():
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss = loss_fn(model(x), y)
        loss.backward()
        opt.step()
        return loss

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        loss = step()
    torch.cuda.synchronize()
    dist.barrier()
    dt = time.perf_counter() - t0
    checksum = float(sum(p.detach().double().sum() for p in model.parameters()))
    return {"rank": dist.get_rank(), "world": dist.get_world_size(), "images_per_sec": batch * steps / dt,
            "loss": float(loss), "param_checksum": checksum, "seconds": dt}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--direct", action="store_true", help="run as a torchrun rank (comparator)")
    args = ap.parse_args()
    if args.direct:
        import torch.distributed as dist

        r = train_resnet50(args.steps, args.warmup, args.batch)
        if r["rank"] == 0:
            print(json.dumps({"what": "ddp_resnet50_torchrun", "n_gpus": r["world"],
                              "images_per_sec": r["images_per_sec"] * r["world"], "loss": r["loss"]}), flush=True)
        dist.destroy_process_group()
        return
    import math

    import kubetorch_b200 as kt

    t0 = time.perf_counter()
    remote = kt.fn(train_resnet50, name="ddp-resnet50").to(
        kt.Compute(gpus=args.gpus, launch_timeout=600).distribute("pytorch", workers=1, num_proc=args.gpus, port=29577))
    t_launch = time.perf_counter() - t0
    try:
        results = remote(args.steps, args.warmup, args.batch)
    finally:
        remote.teardown()
    assert [r["rank"] for r in results] == list(range(args.gpus))  # rank-ordered, one entry per rank
    assert all(math.isfinite(r["loss"]) for r in results)
    assert len({round(r["param_checksum"], 3) for r in results}) == 1, "DDP replicas diverged"
    total = sum(r["images_per_sec"] for r in results)
    print(json.dumps({"what": "ddp_resnet50_kt_launcher", "n_gpus": args.gpus, "images_per_sec": total,
                      "per_rank": [round(r["images_per_sec"], 1) for r in results], "loss": results[0]["loss"],
                      "launch_seconds": t_launch, "batch_per_gpu": args.batch, "steps": args.steps,
                      "dtype": "bf16 autocast, channels_last", "data": "synthetic"}), flush=True)


if __name__ == "__main__":
    main()
