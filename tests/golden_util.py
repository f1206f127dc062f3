"""Helpers shared by the parity tests: golden fixture loading (tests/golden/*.npz)."""
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

SGD = dict(lr=0.1, momentum=0.9, dampening=0.0, nesterov=True, weight_decay=5e-4)

# must mirror tests/golden/make_golden.py::CASES (the fixtures do not store hyper-parameters)
CASES = {
    "v1_tiny": dict(spec="c3,8,3,1,1 n a r1 r1 r1 ap8,1,0 fc32,10", preact=False, use_proj=False,
                    dropout=0.0),
    "wrn_tiny": dict(spec="c3,16,3,1,1 r2 r1 n a ap16,1,0 fc32,10", preact=True, use_proj=True,
                     dropout=0.0),
    "v2_bottleneck_tiny": dict(spec="c3,32,3,1,1 b2 b1 n a ap16,1,0 fc64,10", preact=True,
                               use_proj=True, dropout=0.0),
    "imagenet_style_tiny": dict(spec="c3,32,7,2,3 n a mp3,2,1 b1 b1 ap4,1,0 fc64,12", preact=False,
                                use_proj=True, dropout=0.0),
    "wrn_dropout_tiny": dict(spec="c3,16,3,1,1 r1 r1 n a ap16,1,0 fc32,10", preact=True, use_proj=True,
                             dropout=0.3),
}


def load_case(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    out = {"x": torch.from_numpy(z["x"]), "y": torch.from_numpy(z["y"]), "init": {}, "grad": {},
           "after": {}, "metric": {}}
    for k in z.files:
        if "/" in k:
            grp, key = k.split("/", 1)
            out[grp][key] = torch.from_numpy(z[k])
        elif k not in ("x", "y"):
            out[k] = torch.from_numpy(z[k])
    return out


def rel_l2(a, b):
    a = a.float()
    b = b.float()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()
