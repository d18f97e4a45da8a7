"""(6/6) Minimal ImageNet data-parallel training with the save -> barrier -> load-on-every-rank pattern.
    torchrun --nproc_per_node=8 tutorial/imagenet.py --data ./data/ILSVRC [--synthetic]
Counterpart of reference tutorial/imagenet.py (setup :21-55, save :147-154, barrier :159, load with
map_location :160-165).  On a B200 box pass ``--engine native`` to run the same loop on the fused sm_100a path."""
import os

import torch
import torch.distributed as dist

from common import base_parser
from distribuuuu_b200 import config, models, utils


def main():
    ap = base_parser(__doc__)
    ap.add_argument("--arch", default="resnet18")
    ap.add_argument("--engine", default="torch", choices=["torch", "native"])
    ap.add_argument("--ckpt", default="./imagenet_tutorial.pth.tar")
    ap.set_defaults(batch_size=256, lr=0.1, epochs=1)
    args = ap.parse_args()
    config.cfg.B200.DEVICE = args.device
    utils.setup_distributed()
    rank, world, local_rank = dist.get_rank(), dist.get_world_size(), int(os.environ["LOCAL_RANK"])
    device = utils.resolve_device()

    if args.synthetic or not os.path.isdir(os.path.join(args.data, "train")):
        ds = utils.DummyDataset(args.batch_size * world * 4, [3, 224, 224])
    else:
        import torchvision
        import torchvision.transforms as T
        ds = torchvision.datasets.ImageFolder(os.path.join(args.data, "train"), T.Compose([
            T.RandomResizedCrop(224), T.RandomHorizontalFlip(), T.ToTensor(),
            T.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])]))
    sampler = torch.utils.data.distributed.DistributedSampler(ds, shuffle=True)
    loader = torch.utils.data.DataLoader(ds, batch_size=args.batch_size, sampler=sampler, num_workers=args.workers,
                                         pin_memory=device.type == "cuda", drop_last=True)
    net = models.build_model(args.arch).to(device)
    if args.engine == "native":
        from distribuuuu_b200.parallel.native_engine import NativeEngine
        engine = NativeEngine(net, device)
        opt = engine.make_optimizer(lr=args.lr * world, momentum=0.9, dampening=0.0, weight_decay=1e-4, nesterov=True)
    else:
        from distribuuuu_b200.trainer import TorchEngine
        engine = TorchEngine(net)
        opt = torch.optim.SGD(engine.parameters(), lr=args.lr * world, momentum=0.9, weight_decay=1e-4, nesterov=True)

    for ep in range(args.epochs):
        sampler.set_epoch(ep)
        engine.train()
        for idx, (x, y) in enumerate(loader):
            if args.max_iters and idx >= args.max_iters:
                break
            x, y = x.to(device, non_blocking=True), torch.as_tensor(y).to(device, non_blocking=True)
            loss, hits1, _ = engine.train_step(x, y, opt, 5)
            if rank == 0 and (idx + 1) % args.print_freq == 0:
                print(f"   == step: [{idx + 1:3d}/{len(loader)}] [{ep}] | loss: {loss.item():.3f} | "
                      f"acc: {100.0 * hits1.item() / y.size(0):6.3f}%", flush=True)
        # rank 0 saves; everyone waits; every rank loads the file onto ITS device
        if rank == 0:
            torch.save(utils.unwrap_model(engine).state_dict(), args.ckpt)
        dist.barrier()
        state = torch.load(args.ckpt, map_location=device)
        utils.unwrap_model(engine).load_state_dict(state)
        if hasattr(engine, "on_weights_loaded"):
            engine.on_weights_loaded()
    utils.shutdown()


if __name__ == "__main__":
    main()
