"""Run-to-run spread of the engine-vs-fp32 loss deviation (tests: test_native_engine_tracks_fp32_reference) for each stem path;
REP_ARCH / REP_BATCH / REP_SIZE / REP_LR choose the configuration.  Used to size that test's configuration and tolerance."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import json

import torch  # noqa: F401
from distribuuuu_b200 import selftest as st
out = {}
for tag, env in (("gather", "1"), ("nogather", "0"), ("im2col_stem", None)):
    if env is None:
        os.environ["B200_STEM_S2D"] = "0"
    else:
        os.environ["B200_STEM_S2D"] = "1"; os.environ["B200_STEM_GATHER"] = env
    vals = []
    for i in range(12):
        if i % 3 == 0:   # disturb the allocator state like the test sequence does
            try: st.check_stem_s2d(N=16, H=224, W=224)
            except Exception as e: vals.append("s2dFAIL:" + str(e)[:100])
        try:
            r = st.check_engine_vs_torch(os.environ.get("REP_ARCH", "resnet50"), batch=int(os.environ.get("REP_BATCH", "8")), size=int(os.environ.get("REP_SIZE", "64")),
                                          lr=float(os.environ.get("REP_LR", "0.01")), tol=10.0)
            vals.append(round(r["max_rel_loss_diff"], 4))
        except Exception as e:
            vals.append("FAIL:" + str(e)[:200])
    out[tag] = vals
    print(tag, vals, flush=True)
json.dump(out, open("gpurun_out/rep_engine.json", "w"))
