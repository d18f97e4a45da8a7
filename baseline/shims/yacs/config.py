"""`yacs.config.CfgNode` stand-in for running the UNMODIFIED reference offline.
The reference uses: CN(), attribute set/get, clone, freeze, load_cfg, merge_from_other_cfg, merge_from_list, dump.
This is configuration plumbing only -- no model, kernel or engine code of this repo is on the reference's path."""
from distribuuuu_b200.config import CfgNode  # noqa: F401
