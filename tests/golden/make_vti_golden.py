"""Decode the reference's two .vti fixtures into a compact array fixture (build container only).

    python tests/golden/make_vti_golden.py [/root/reference]

`example_data/fuel.vti` (64^3) and `hydrogen.vti` (128^3) are the inputs of BASELINE.json configs[1].
The GPU box has no /root/reference, so their decoded payloads travel as `vti_volumes.npz`: both hold
integral values 0..255 stored as float32, so uint8 is lossless.  The decoder used here is the repo's
std-lib reader; its output is pinned by the sha256 of the decoded float32 bytes that SURVEY.md section
8(c) recorded from an independent probe (`586231c5...`, `6657a75a...`) -- asserted below -- and by the
properties the reference's own loader tests assert (tests/test_dataloaders/test_vtk_loader.py:14-76).
"""

import hashlib
import json
import os
import sys

import numpy as np

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(OUT)))

from pyvr_b200.dataloaders import read_vti  # noqa: E402

SURVEY_SHA_PREFIX = {"fuel": "586231c59df490bd", "hydrogen": "6657a75aa7704b27"}

arrays, meta = {}, {}
for name in ("fuel", "hydrogen"):
    dims, spacing, found = read_vti(os.path.join(REF, "example_data", f"{name}.vti"))
    flat, ncomp = found["Scalars_"]
    assert ncomp == 1 and flat.dtype == np.float32
    digest = hashlib.sha256(flat.astype("<f4").tobytes()).hexdigest()
    assert digest.startswith(SURVEY_SHA_PREFIX[name]), (name, digest)
    assert np.all(flat == np.round(flat)) and flat.min() >= 0 and flat.max() <= 255
    arrays[name] = flat.astype(np.uint8).reshape(dims[2], dims[1], dims[0])
    meta[name] = {"dims_xyz": list(dims), "spacing_xyz": list(spacing), "sha256_f32": digest,
                  "min": float(flat.min()), "max": float(flat.max()), "nonzero": int(np.count_nonzero(flat)),
                  "distinct": int(len(np.unique(flat)))}
np.savez_compressed(os.path.join(OUT, "vti_volumes.npz"), **arrays)
with open(os.path.join(OUT, "vti_meta.json"), "w") as f:
    json.dump(meta, f, indent=1)
print(meta)
