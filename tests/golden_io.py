"""Loaders for the committed golden vectors (made by oracle/make_golden.py from the reference)."""
import os

import numpy as np
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _deq(m, scale):
    return torch.from_numpy(m.astype(np.float32) * np.float32(scale))


def load_features():
    z = np.load(os.path.join(GOLD, "features.npz"))
    clips, o = [], 0
    for n, s in zip(z["lens"], z["scales"]):
        clips.append(_deq(z["mant"][o:o + n], s))
        o += int(n)
    return clips, torch.from_numpy(z["features"]), z


def load_model():
    z = np.load(os.path.join(GOLD, "model.npz"))
    sd0 = {k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd0/")}
    sd1 = dict(sd0)
    sd1.update({k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd1/")})
    return sd0, sd1, z


def load_stream():
    z = np.load(os.path.join(GOLD, "stream.npz"))
    return _deq(z["mant"], z["scale"]), z


def load_engine():
    return np.load(os.path.join(GOLD, "engine.npz"))


def load_resample():
    z = np.load(os.path.join(GOLD, "resample.npz"))
    out = {}
    for k in z.files:
        if k.startswith("out_"):
            sr = int(k[4:])
            x = _deq(z[f"in_{sr}_mant"], z[f"in_{sr}_scale"]).reshape(2, -1)
            out[sr] = (x, torch.from_numpy(z[k]))
    return out


def load_feature_flags():
    """{"pre" | "pcen" | "pre_pcen": (32, 90, 101)} for the clips of load_features(), plus the both-flags stream case."""
    z = np.load(os.path.join(GOLD, "features_flags.npz"))
    return {k[len("features_"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("features_")}, \
        torch.from_numpy(z["stream_pre_pcen_hop100"])


def load_delta_delta_rows():
    """{"dd" | "pre_pcen_dd": (32, 13, 101)}: rows 90..102 of the reference's 103-row output (rows 0..89 equal the 90-row case)."""
    z = np.load(os.path.join(GOLD, "features_flags.npz"))
    return {k[len("ddrows_"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("ddrows_")}


def load_delta_delta_logits():
    """Reference logits of the 103-row ("dd") golden features, default-init and randomised weights: {name: (32, 2)}."""
    z = np.load(os.path.join(GOLD, "features_flags.npz"))
    return {k[len("logits_dd_"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("logits_dd_")}


def load_other_model(kind: str):
    """(state_dict, logits on the 90-row golden features, logits on the 103-row ones) of the reference's 'small' / 'standard'
    network with randomised BatchNorm statistics (oracle/make_golden.py models)."""
    z = np.load(os.path.join(GOLD, "models_other.npz"))
    pre = f"{kind}/sd/"
    sd = {k[len(pre):]: torch.from_numpy(z[k].astype(np.float32)) for k in z.files if k.startswith(pre)}
    for k in list(sd):
        if k.endswith("running_mean"):
            sd[k[: -len("running_mean")] + "num_batches_tracked"] = torch.zeros((), dtype=torch.int64)
    return sd, torch.from_numpy(z[f"{kind}/logits90"]), torch.from_numpy(z[f"{kind}/logits103"])


def load_spec_augment():
    """[(rows, SpecAugment kwargs, zero-cell masks (8 calls, rows, 101) bool)] per case, and the shared-mask batch case
    (3, 1, 90, 101) bool, recorded from the reference's SpecAugment (oracle/make_golden.py augment)."""
    from oracle.make_golden import AUGMENT_CASES
    z = np.load(os.path.join(GOLD, "spec_augment.npz"))
    cases = []
    for ci, case in enumerate(AUGMENT_CASES):
        n = case["rows"] * 101
        bits = np.stack([np.unpackbits(r)[:n] for r in z[f"zero_{ci}"]]).astype(bool).reshape(8, case["rows"], 101)
        cases.append((case["rows"], case["kw"], bits))
    batch = np.unpackbits(z["zero_batch"])[: 3 * 90 * 101].astype(bool).reshape(3, 1, 90, 101)
    return cases, batch
