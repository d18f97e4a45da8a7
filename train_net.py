"""Train a classification model.

    python -m torch.distributed.run --nproc_per_node=8 train_net.py --cfg config/resnet50.yaml [KEY VALUE ...]
    srun ... python -u train_net.py --cfg config/resnet18.yaml        # Slurm
"""
from distribuuuu_b200.cli import main

if __name__ == "__main__":
    main("train")
