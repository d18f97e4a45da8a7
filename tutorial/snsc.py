"""(1/6) Single node, single device.   python tutorial/snsc.py [--synthetic]
Counterpart of reference tutorial/snsc.py: plain loop, no parallelism."""
import torch

from common import base_parser, cifar10, make_net, pick_device, train_one_epoch


def main():
    args = base_parser(__doc__).parse_args()
    device = pick_device(args.device)
    net = make_net().to(device)
    loader = torch.utils.data.DataLoader(cifar10(args.data, args.synthetic), batch_size=args.batch_size, shuffle=True,
                                         num_workers=args.workers, pin_memory=device.type == "cuda")
    opt = torch.optim.SGD(net.parameters(), lr=args.lr, momentum=0.9, weight_decay=1e-4, nesterov=True)
    print("            =======  Training  ======= \n")
    for ep in range(1, args.epochs + 1):
        train_one_epoch(net, loader, opt, device, ep, 0, args.print_freq, args.max_iters)
    print("\n            =======  Training Finished  ======= \n")


if __name__ == "__main__":
    main()
