"""Config system: schema parity with reference config.py:10-63, precedence, typing, freeze."""
import os

import pytest

from distribuuuu_b200 import config
from distribuuuu_b200.config import CfgNode

REFERENCE_DEFAULTS = {
    "MODEL.ARCH": "resnet18", "MODEL.NUM_CLASSES": 1000, "MODEL.PRETRAINED": False, "MODEL.SYNCBN": False,
    "MODEL.WEIGHTS": None, "MODEL.DUMMY_INPUT": False, "TRAIN.BATCH_SIZE": 32, "TRAIN.IM_SIZE": 224,
    "TRAIN.DATASET": "./data/ILSVRC/", "TRAIN.SPLIT": "train", "TRAIN.AUTO_RESUME": True, "TRAIN.LOAD_OPT": True,
    "TRAIN.WORKERS": 4, "TRAIN.PIN_MEMORY": True, "TRAIN.PRINT_FREQ": 30, "TRAIN.TOPK": 5,
    "TEST.DATASET": "./data/ILSVRC/", "TEST.SPLIT": "val", "TEST.BATCH_SIZE": 200, "TEST.IM_SIZE": 256,
    "TEST.PRINT_FREQ": 10, "CUDNN.BENCHMARK": True, "CUDNN.DETERMINISTIC": False, "OPTIM.MAX_EPOCH": 100,
    "OPTIM.LR_POLICY": "cos", "OPTIM.BASE_LR": 0.2, "OPTIM.MIN_LR": 0.0, "OPTIM.STEPS": [], "OPTIM.LR_MULT": 0.1,
    "OPTIM.MOMENTUM": 0.9, "OPTIM.DAMPENING": 0.0, "OPTIM.NESTEROV": True, "OPTIM.WARMUP_FACTOR": 0.1,
    "OPTIM.WARMUP_EPOCHS": 5, "OPTIM.WEIGHT_DECAY": 5e-5, "OUT_DIR": "./exp", "CFG_DEST": "config.yaml",
    "RNG_SEED": None,
}


def _get(node, dotted):
    for part in dotted.split("."):
        node = node[part]
    return node


def test_defaults_match_reference_schema(fresh_cfg):
    for key, val in REFERENCE_DEFAULTS.items():
        assert _get(fresh_cfg, key) == val, key


@pytest.mark.parametrize("name,arch,lr,wd,bs", [
    ("resnet18", "resnet18", 0.2, 5e-5, 32), ("resnet50", "resnet50", 0.2, 5e-5, 32),
    ("botnet50", "botnet50", 0.2, 5e-5, 32), ("efficientnet_b0", "efficientnet_b0", 0.4, 1e-5, 64),
    ("regnetx_160", "regnetx_160", 0.4, 5e-5, 64), ("regnety_160", "regnety_160", 0.4, 5e-5, 64),
    ("regnety_320", "regnety_320", 0.4, 5e-5, 64)])
def test_presets(fresh_cfg, name, arch, lr, wd, bs):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    config.merge_from_file(os.path.join(root, "config", f"{name}.yaml"))
    assert fresh_cfg.MODEL.ARCH == arch and fresh_cfg.OPTIM.BASE_LR == lr
    assert fresh_cfg.OPTIM.WEIGHT_DECAY == pytest.approx(wd) and fresh_cfg.TRAIN.BATCH_SIZE == bs
    assert fresh_cfg.OUT_DIR == f"./{name}"


def test_full_yacs_style_dump_roundtrips(fresh_cfg, tmp_path):
    """A complete dump (what the reference ships as config/*.yaml) loads unchanged."""
    fresh_cfg.MODEL.ARCH = "resnet50"
    fresh_cfg.OPTIM.STEPS = [0, 30, 60]
    path = tmp_path / "full.yaml"
    path.write_text(fresh_cfg.dump())
    config.reset_cfg()
    config.merge_from_file(str(path))
    assert config.cfg.MODEL.ARCH == "resnet50" and config.cfg.OPTIM.STEPS == [0, 30, 60]
    assert config.cfg.MODEL.WEIGHTS is None


def test_reference_yaml_without_b200_section_loads(fresh_cfg, tmp_path):
    ref_yaml = "CFG_DEST: config.yaml\nCUDNN:\n  BENCHMARK: true\nMODEL:\n  ARCH: resnet50\n  WEIGHTS: null\n" \
               "OPTIM:\n  WEIGHT_DECAY: 5.0e-05\n  STEPS: []\nRNG_SEED: null\n"
    p = tmp_path / "ref.yaml"
    p.write_text(ref_yaml)
    config.merge_from_file(str(p))
    assert fresh_cfg.MODEL.ARCH == "resnet50" and fresh_cfg.B200.ENGINE == "auto"


def test_cli_precedence_and_literals(fresh_cfg, tmp_path):
    p = tmp_path / "a.yaml"
    p.write_text("OPTIM:\n  BASE_LR: 0.4\nTRAIN:\n  BATCH_SIZE: 64\n")
    config.load_cfg_fom_args(argv=["--cfg", str(p), "--local_rank", "3", "OPTIM.BASE_LR", "0.8", "MODEL.SYNCBN",
                                   "True", "OPTIM.STEPS", "[0, 30]", "MODEL.WEIGHTS", "w.pth", "RNG_SEED", "7",
                                   "OPTIM.MAX_EPOCH", "5"])
    assert fresh_cfg.OPTIM.BASE_LR == 0.8 and fresh_cfg.TRAIN.BATCH_SIZE == 64
    assert fresh_cfg.MODEL.SYNCBN is True and fresh_cfg.OPTIM.STEPS == [0, 30]
    assert fresh_cfg.MODEL.WEIGHTS == "w.pth" and fresh_cfg.RNG_SEED == 7 and fresh_cfg.OPTIM.MAX_EPOCH == 5


def test_unknown_key_and_type_mismatch_rejected(fresh_cfg):
    with pytest.raises(KeyError):
        fresh_cfg.merge_from_list(["MODEL.NOPE", "1"])
    with pytest.raises(ValueError):
        fresh_cfg.merge_from_list(["TRAIN.BATCH_SIZE", "big"])
    with pytest.raises(KeyError):
        fresh_cfg.merge_from_other_cfg(CfgNode({"BOGUS": 1}))
    fresh_cfg.merge_from_list(["OPTIM.BASE_LR", "1"])  # int -> float is an allowed cast
    assert fresh_cfg.OPTIM.BASE_LR == 1.0 and isinstance(fresh_cfg.OPTIM.BASE_LR, float)


def test_freeze_clone_reset(fresh_cfg):
    fresh_cfg.OPTIM.BASE_LR = 3.0
    snap = fresh_cfg.clone()
    fresh_cfg.freeze()
    with pytest.raises(AttributeError):
        fresh_cfg.OPTIM.BASE_LR = 1.0
    with pytest.raises(AttributeError):
        fresh_cfg.merge_from_list(["OPTIM.BASE_LR", "1.0"])
    assert not snap.is_frozen() and snap.OPTIM.BASE_LR == 3.0
    config.reset_cfg()
    assert config.cfg.OPTIM.BASE_LR == 0.2


def test_dump_cfg_writes_file(fresh_cfg, tmp_path):
    fresh_cfg.OUT_DIR = str(tmp_path / "o")
    config.dump_cfg()
    loaded = CfgNode.load_cfg(open(tmp_path / "o" / "config.yaml"))
    assert loaded.MODEL.ARCH == "resnet18"


# ---- property tests: the KEY VALUE override parser and the dump/load round trip -------------------------------
from hypothesis import given, settings, strategies as st  # noqa: E402


@settings(max_examples=60, deadline=None)
@given(lr=st.floats(min_value=1e-6, max_value=10, allow_nan=False, allow_infinity=False),
       epochs=st.integers(min_value=1, max_value=10_000),
       steps=st.lists(st.integers(min_value=0, max_value=500), min_size=0, max_size=5),
       syncbn=st.booleans(),
       arch=st.sampled_from(["resnet18", "resnet50", "botnet50", "regnety_160", "efficientnet_b0"]))
def test_override_strings_parse_to_typed_values(lr, epochs, steps, syncbn, arch):
    """Every value arrives as a string on the command line (reference config.py:95-100) and must come back typed."""
    from distribuuuu_b200 import config
    config.reset_cfg()
    c = config.cfg
    c.merge_from_list(["OPTIM.BASE_LR", repr(lr), "OPTIM.MAX_EPOCH", str(epochs), "OPTIM.STEPS", str(steps),
                       "MODEL.SYNCBN", str(syncbn), "MODEL.ARCH", arch])
    assert isinstance(c.OPTIM.BASE_LR, float) and c.OPTIM.BASE_LR == lr
    assert isinstance(c.OPTIM.MAX_EPOCH, int) and c.OPTIM.MAX_EPOCH == epochs
    assert list(c.OPTIM.STEPS) == steps and c.MODEL.SYNCBN is syncbn and c.MODEL.ARCH == arch
    # dump -> load reproduces the node exactly
    again = config.CfgNode.load_cfg(c.dump()) if hasattr(config.CfgNode, "load_cfg") else None
    if again is not None:
        assert again.OPTIM.BASE_LR == lr and again.OPTIM.MAX_EPOCH == epochs and list(again.OPTIM.STEPS) == steps
    config.reset_cfg()


@settings(max_examples=40, deadline=None)
@given(e=st.integers(min_value=0, max_value=299), warm=st.integers(min_value=0, max_value=20),
       base=st.floats(min_value=1e-4, max_value=5.0), min_frac=st.floats(min_value=0.0, max_value=0.5))
def test_cosine_lr_is_bounded_and_monotone_after_warmup(e, warm, base, min_frac):
    """lr(e) stays within [MIN_LR*BASE_LR, BASE_LR] and never increases once warm-up is over (reference utils.py:286-310)."""
    from distribuuuu_b200 import config, utils
    config.reset_cfg()
    c = config.cfg
    c.OPTIM.LR_POLICY, c.OPTIM.MAX_EPOCH, c.OPTIM.BASE_LR = "cos", 300, base
    c.OPTIM.WARMUP_EPOCHS, c.OPTIM.MIN_LR = warm, min_frac
    lr0, lr1 = utils.get_epoch_lr(e), utils.get_epoch_lr(e + 1)
    assert 0.0 <= lr0 <= base * (1 + 1e-9)
    if e >= warm:
        assert lr1 <= lr0 * (1 + 1e-9) and lr0 >= base * min_frac * (1 - 1e-9)
    config.reset_cfg()
