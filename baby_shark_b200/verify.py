"""Order-sensitive fingerprints of a downloaded volume and of a vertex soup (numpy + hashlib only).

`bench.py` asserts the fingerprints of its own output against `tests/golden/config_hashes.json`, which
`tests/golden/make_config_hashes.py` wrote from the CPU checker at the same size: a benchmark whose result is not the
reference's result is not a result."""
import hashlib

import numpy as np


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:32]


def active_bits(masks):
    m = np.ascontiguousarray(masks, np.uint64)
    return np.unpackbits(m.view(np.uint8).reshape(-1, 8, 8), axis=-1, bitorder="little").reshape(-1, 512).astype(bool)


def fingerprint_volume(d):
    """d: dict from Volume.download() (origins, values, masks in the reference's leaf visit order)."""
    act = active_bits(d["masks"])
    v = np.ascontiguousarray(d["values"], np.float32).reshape(-1, 512)[act]
    neg = np.signbit(v)
    return {
        "n_bricks": int(d["origins"].shape[0]), "origins_sha": _sha(np.asarray(d["origins"], np.int32)),
        "masks_sha": _sha(np.asarray(d["masks"], np.uint64)), "n_active": int(v.size), "n_negative": int(neg.sum()),
        "abs_sha": _sha(np.abs(v).view(np.uint32)), "sign_sha": _sha(np.packbits(neg)),
    }


def fingerprint_soup(verts):
    v = np.ascontiguousarray(verts, np.float32).reshape(-1, 3)
    return {"n_verts": int(v.shape[0]), "soup_sha": _sha(v.view(np.uint32))}
