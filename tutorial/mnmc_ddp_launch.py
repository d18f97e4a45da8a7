"""(3/6) Multi node / multi device, one process per device, started by a launcher:
    torchrun --nproc_per_node=4 tutorial/mnmc_ddp_launch.py
    # two "nodes" on one host (reference README.md:123-142):
    CUDA_VISIBLE_DEVICES=0,1 torchrun --nnodes=2 --node_rank=0 --nproc_per_node=2 --master_addr=127.0.0.1 --master_port=29500 tutorial/mnmc_ddp_launch.py
    CUDA_VISIBLE_DEVICES=2,3 torchrun --nnodes=2 --node_rank=1 --nproc_per_node=2 --master_addr=127.0.0.1 --master_port=29500 tutorial/mnmc_ddp_launch.py
Counterpart of reference tutorial/mnmc_ddp_launch.py: env:// rendezvous (RANK / WORLD_SIZE / LOCAL_RANK set by the
launcher), DistributedSampler + set_epoch, LR scaled by world size, gradients averaged by the data-parallel wrapper."""
import os

import torch
import torch.distributed as dist

from common import base_parser, cifar10, make_net, pick_device, train_one_epoch
from distribuuuu_b200.parallel import BucketedDataParallel


def main():
    args = base_parser(__doc__).parse_args()
    rank, world, local_rank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    device = pick_device(args.device, local_rank)
    if device.type == "cuda":
        torch.cuda.set_device(device)
    dist.init_process_group(backend="nccl" if device.type == "cuda" else "gloo")  # env:// rendezvous
    print(f"[init] == local rank: {local_rank}, global rank: {rank} ==", flush=True)

    net = BucketedDataParallel(make_net().to(device))   # broadcasts rank 0's weights, averages gradients in buckets
    ds = cifar10(args.data, args.synthetic)
    sampler = torch.utils.data.distributed.DistributedSampler(ds, shuffle=True)
    loader = torch.utils.data.DataLoader(ds, batch_size=args.batch_size, sampler=sampler, num_workers=args.workers,
                                         pin_memory=device.type == "cuda")
    opt = torch.optim.SGD(net.parameters(), lr=args.lr * world, momentum=0.9, weight_decay=1e-4, nesterov=True)
    opt.register_step_pre_hook(lambda *_: net.finish_backward())   # wait for the bucket all-reduces before stepping
    if rank == 0:
        print("            =======  Training  ======= \n")
    for ep in range(1, args.epochs + 1):
        train_one_epoch(net, loader, opt, device, ep, rank, args.print_freq, args.max_iters, sampler)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
