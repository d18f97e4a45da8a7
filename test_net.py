"""Evaluate a trained model: ``... test_net.py --cfg config/resnet18.yaml MODEL.WEIGHTS path``
(entry contract of reference test_net.py:6-9)."""
from distribuuuu_b200 import config, trainer


def main():
    config.load_cfg_fom_args("Test a classification model.")
    config.cfg.freeze()
    trainer.test_model()


if __name__ == "__main__":
    main()
