"""(4/6) Same as (3) but the processes are created with ``torch.multiprocessing.spawn`` and rendezvous over tcp://
    python tutorial/mnmc_ddp_mp.py --nodes 1 --nproc-per-node 2 [--ip 127.0.0.1 --port 23456 --node-rank 0]
Counterpart of reference tutorial/mnmc_ddp_mp.py (args :20-38, spawn :47, tcp init :58-63)."""
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from common import base_parser, cifar10, make_net, pick_device, train_one_epoch
from distribuuuu_b200.parallel import BucketedDataParallel


def worker(local_rank, args):
    rank = args.node_rank * args.nproc_per_node + local_rank
    world = args.nodes * args.nproc_per_node
    device = pick_device(args.device, local_rank)
    if device.type == "cuda":
        torch.cuda.set_device(device)
    dist.init_process_group(backend="nccl" if device.type == "cuda" else "gloo",
                            init_method=f"tcp://{args.ip}:{args.port}", world_size=world, rank=rank)
    print(f"[init] == local rank: {local_rank}, global rank: {rank} ==", flush=True)
    net = BucketedDataParallel(make_net().to(device))
    ds = cifar10(args.data, args.synthetic)
    sampler = torch.utils.data.distributed.DistributedSampler(ds, shuffle=True)
    loader = torch.utils.data.DataLoader(ds, batch_size=args.batch_size, sampler=sampler, num_workers=args.workers,
                                         pin_memory=device.type == "cuda")
    opt = torch.optim.SGD(net.parameters(), lr=args.lr * world, momentum=0.9, weight_decay=1e-4, nesterov=True)
    opt.register_step_pre_hook(lambda *_: net.finish_backward())
    for ep in range(1, args.epochs + 1):
        train_one_epoch(net, loader, opt, device, ep, rank, args.print_freq, args.max_iters, sampler)
    dist.destroy_process_group()


def main():
    ap = base_parser(__doc__)
    ap.add_argument("--nodes", type=int, default=1)
    ap.add_argument("--nproc-per-node", type=int, default=2)
    ap.add_argument("--node-rank", type=int, default=0)
    ap.add_argument("--ip", default="127.0.0.1")
    ap.add_argument("--port", default="23456")
    args = ap.parse_args()
    mp.spawn(worker, nprocs=args.nproc_per_node, args=(args,))


if __name__ == "__main__":
    main()
