"""oracle/weights.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Seeded, reference-free reconstruction of a model ``state_dict`` from a small
JSON manifest (names, shapes, roles) that tests/golden/make_golden.py writes
from the UNMODIFIED reference model classes.  The same (manifest, seed) pair
gives bit-identical tensors in this container and on the GPU box, so multi-MB
checkpoints never have to be committed.  BN statistics are deliberately
non-trivial so that BN-folding mistakes are visible.
"""
import json
import zlib

import numpy as np
import torch


def classify(name, tensor):
    leaf = name.rsplit(".", 1)[-1]
    if leaf == "num_batches_tracked":
        return "counter"
    if leaf == "running_mean":
        return "bn_mean"
    if leaf == "running_var":
        return "bn_var"
    if leaf == "kernel_points":
        return "kpoints"
    if leaf == "weight" and tensor.dim() == 1:
        return "bn_weight"
    if leaf == "bias":
        return "bias"
    return "weight"


def manifest_from_state_dict(sd):
    out = []
    for k, v in sd.items():
        kind = classify(k, v)
        scale = float(v.abs().max()) if kind == "kpoints" else 0.0
        out.append(dict(name=k, shape=list(v.shape), kind=kind, scale=scale))
    return out


def _fan_in(name, shape):
    if name.endswith("KPConv.weights"):  # [K, Cin, Cout]
        return shape[0] * shape[1]
    if "decoder" in name and name.endswith("conv.weight") and len(shape) == 4:
        return shape[0]  # ConvTranspose2d [in,out,1,1] (randlanet.py:486-491)
    if "deblocks" in name and len(shape) == 4:
        return shape[0]  # ConvTranspose2d [in,out,k,k]; each output sees Cin taps
    n = 1
    for s in shape[1:]:
        n *= s
    return max(n, 1)


def seeded_state_dict(manifest, seed=0):
    """-> dict name -> float32 (or int64 counter) CPU tensor."""
    sd = {}
    for e in manifest:
        name, shape, kind = e["name"], tuple(e["shape"]), e["kind"]
        rng = np.random.default_rng([seed, zlib.crc32(name.encode())])
        if kind == "counter":
            sd[name] = torch.zeros(shape, dtype=torch.int64)
            continue
        if kind == "bn_mean":
            a = rng.standard_normal(shape) * 0.1
        elif kind == "bn_var":
            a = rng.uniform(0.5, 2.0, shape)
        elif kind == "bn_weight":
            a = rng.uniform(0.6, 1.4, shape)
        elif kind == "bias":
            a = rng.standard_normal(shape) * 0.1
        elif kind == "kpoints":
            d = rng.standard_normal(shape)
            d /= np.linalg.norm(d, axis=-1, keepdims=True)
            a = d * (rng.uniform(0.2, 1.0, shape[:-1] + (1,)) ** (1 / 3)) * e["scale"] / 1.0
            a[0] = 0.0  # fixed_kernel_points: center (kpconv.py:1909-1999)
        else:
            lim = np.sqrt(6.0 / _fan_in(name, shape))
            a = rng.uniform(-lim, lim, shape)
        sd[name] = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
    return sd


def save_manifest(path, manifest, extra=None):
    with open(path, "w") as f:
        json.dump(dict(params=manifest, **(extra or {})), f, indent=0, separators=(",", ":"))


def load_manifest(path):
    with open(path) as f:
        d = json.load(f)
    return d["params"], {k: v for k, v in d.items() if k != "params"}
