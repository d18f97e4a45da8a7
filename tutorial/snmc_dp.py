"""(2/6) Single node, multiple devices, ONE process: ``nn.DataParallel`` (scatter / replicate / gather per step).
    CUDA_VISIBLE_DEVICES=0,1 python tutorial/snmc_dp.py
Counterpart of reference tutorial/snmc_dp.py:21.  Kept for completeness -- the one-process-per-GPU scripts that
follow are faster (no GIL contention, no per-step replication) and are what the framework itself uses."""
import torch
import torch.nn as nn

from common import base_parser, cifar10, make_net, pick_device, train_one_epoch


def main():
    args = base_parser(__doc__).parse_args()
    device = pick_device(args.device)
    net = make_net().to(device)
    n_dev = torch.cuda.device_count() if device.type == "cuda" else 1
    if n_dev > 1:
        net = nn.DataParallel(net)  # splits each batch along dim 0 over the visible GPUs
    loader = torch.utils.data.DataLoader(cifar10(args.data, args.synthetic), batch_size=args.batch_size * n_dev,
                                         shuffle=True, num_workers=args.workers, pin_memory=device.type == "cuda")
    opt = torch.optim.SGD(net.parameters(), lr=args.lr, momentum=0.9, weight_decay=1e-4, nesterov=True)
    print(f"            =======  Training ({n_dev} device(s), DataParallel)  ======= \n")
    for ep in range(1, args.epochs + 1):
        train_one_epoch(net, loader, opt, device, ep, 0, args.print_freq, args.max_iters)


if __name__ == "__main__":
    main()
