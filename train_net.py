"""Train a classification model.

    python -m torch.distributed.run --nproc_per_node=8 train_net.py --cfg config/resnet50.yaml [KEY VALUE ...]
    srun ... python -u train_net.py --cfg config/resnet18.yaml        # Slurm
(entry contract of reference train_net.py:6-9)."""
from distribuuuu_b200 import config, trainer


def main():
    config.load_cfg_fom_args("Train a classification model.")
    config.cfg.freeze()
    trainer.train_model()


if __name__ == "__main__":
    main()
