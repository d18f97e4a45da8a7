"""(5/6) One process per device under Slurm:
    srun --partition=<p> -n16 --gres=gpu:8 --ntasks-per-node=8 --job-name=demo python -u tutorial/mnmc_ddp_slurm.py
Counterpart of reference tutorial/mnmc_ddp_slurm.py (setup_distributed :21-42, lr x world :93): rank / world / master
address come from SLURM_PROCID / SLURM_NTASKS / SLURM_NODELIST; ``utils.setup_distributed`` handles both Slurm and
launcher environments, so the same script also runs under torchrun."""
import os

import torch
import torch.distributed as dist

from common import base_parser, cifar10, make_net, train_one_epoch
from distribuuuu_b200 import config, utils
from distribuuuu_b200.parallel import BucketedDataParallel


def main():
    args = base_parser(__doc__).parse_args()
    config.cfg.B200.DEVICE = args.device
    utils.setup_distributed(port=os.environ.get("MASTER_PORT", 29500))
    rank, world = dist.get_rank(), dist.get_world_size()
    device = utils.resolve_device()
    print(f"[init] == local rank: {os.environ['LOCAL_RANK']}, global rank: {rank} ==", flush=True)
    net = BucketedDataParallel(make_net().to(device))
    ds = cifar10(args.data, args.synthetic)
    sampler = torch.utils.data.distributed.DistributedSampler(ds, shuffle=True)
    loader = torch.utils.data.DataLoader(ds, batch_size=args.batch_size, sampler=sampler, num_workers=args.workers,
                                         pin_memory=device.type == "cuda")
    opt = torch.optim.SGD(net.parameters(), lr=args.lr * world, momentum=0.9, weight_decay=1e-4, nesterov=True)
    opt.register_step_pre_hook(lambda *_: net.finish_backward())
    for ep in range(1, args.epochs + 1):
        train_one_epoch(net, loader, opt, device, ep, rank, args.print_freq, args.max_iters, sampler)
    utils.shutdown()


if __name__ == "__main__":
    main()
