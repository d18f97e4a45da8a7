"""Evaluate a trained model on the validation split.

    python -m torch.distributed.run --nproc_per_node=8 test_net.py --cfg config/resnet18.yaml MODEL.WEIGHTS best.pth.tar
"""
from distribuuuu_b200.cli import main

if __name__ == "__main__":
    main("test")
