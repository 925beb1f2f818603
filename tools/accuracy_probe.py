"""Full-width accuracy of the bf16 hot path against the REFERENCE's own forward (tests/golden/unet_full512.npz),
under the precision switches of UNetModel.  Prints one line per setting.

    python tools/accuracy_probe.py
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "open-pandora_b200")]
from pandora_b200 import UNetModel, config  # noqa: E402
from pandora_b200.synthetic import randomize_module_, synth_inputs  # noqa: E402


def rel(a, b):
    a, b = a.double().cpu(), torch.as_tensor(b).double()
    return float((a - b).norm() / b.norm())


def main():
    z = np.load(os.path.join(ROOT, "tests", "golden", "unet_full512.npz"))
    meta = json.loads(str(z["meta"]))
    with torch.device("cuda"):
        m = UNetModel(**config.UNET_512).eval()
    randomize_module_(m, meta["seed_w"])
    inp = {k: v.cuda() for k, v in synth_inputs(1, tuple(meta["latent"]), 1024, meta["seed_x"]).items()}
    t = torch.tensor([499], device="cuda")
    fs = torch.tensor([meta["fs"]], device="cuda")
    m.compute_mode = "bf16"
    for flag in (False, True):
        m.bf16_norm_inputs = flag
        m.invalidate_packed_weights()
        y = m(inp["x"], t, context=inp["context"], fs=fs, x_concat=[inp["c_concat"]])
        print(json.dumps(dict(bf16_norm_inputs=flag, rel_l2_vs_reference=rel(y, z["y"]), rel_l2_vs_reference_fp64=rel(y, z["y64"]))))


if __name__ == "__main__":
    main()
