"""`timm` is imported at the top of the reference trainer (trainer.py:4) but only used for archs outside its own
zoo (RegNet / EfficientNet).  It cannot be installed offline; this stub lets the import succeed and fails loudly
if such an arch is requested."""


def create_model(model_name, pretrained=False, num_classes=1000, **kwargs):
    raise RuntimeError(f"timm is not available offline: cannot build '{model_name}' for the reference arm")
