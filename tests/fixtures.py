"""Loading of the committed golden fixtures (tests/golden/*.npz, produced by make_golden.py from the reference)."""
from __future__ import annotations

import glob
import json
import os

import numpy as np
import torch

from oracle import layoutdm_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NAMES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz")))
_cache = {}


class Fixture:
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN, name + ".npz"))
        self.meta = json.loads(str(z["meta"]))
        self.z = z
        m = self.meta
        self.vocab = O.RICO25 if m["dataset"] == "rico25" else O.PUBLAYNET
        self.spec = O.ModelSpec(T=m["T"])
        self.plan = [tuple(p) for p in m["plan"]]
        self.B = m["B"]
        self.x_in = torch.from_numpy(z["x_in"].astype(np.int64))
        self.x_out = torch.from_numpy(z["x_out"].astype(np.int64))
        self.ids_final = torch.from_numpy(z["ids_final"].astype(np.int64))
        self.trace_steps = m["trace_steps"]
        self.cfg = O.SamplingCfg(name=m["sampling"], temperature=m["temperature"], top_p=m["top_p"], top_k=m["top_k"],
                                 num_timesteps=m["T_eval"], time_difference=m["time_difference"])
        self.cfg_dict = dict(name=m["sampling"], temperature=m["temperature"], top_p=m["top_p"], top_k=m["top_k"],
                             num_timesteps=m["T_eval"], time_difference=m["time_difference"])
        self.cond = None
        if m["cond"]:
            self.cond = dict(seq=torch.from_numpy(z["cond_seq"].astype(np.int64)), mask=torch.from_numpy(z["cond_mask"]), type=m["cond"])
            if "cond_seq_orig" in z:
                self.cond["seq_orig"] = torch.from_numpy(z["cond_seq_orig"].astype(np.int64))
                self.cond["refine_table"] = torch.from_numpy(z["refine_table"])

    def weights(self):
        key = (self.meta["dataset"], self.meta["T"], self.meta["weight_seed"], self.meta["weight_scale"])
        if key not in _cache:
            sd = O.make_weights(self.vocab, self.spec, seed=self.meta["weight_seed"], scale=self.meta["weight_scale"])
            chk = O.weights_checksum(sd)
            assert abs(chk - self.meta["weights_checksum"]) <= 1e-6 * abs(chk), "synthetic weight generator drifted from the fixture"
            _cache.clear()
            _cache[key] = sd
        return _cache[key]

    def logits(self, i):
        return torch.from_numpy(self.z[f"logits_{i}"])

    def logp(self, i):
        return torch.from_numpy(self.z[f"logp_{i}"])

    def noise(self, i):
        m = self.meta
        u = ug = None
        if m["sampling"] != "deterministic":
            u = O.uniforms(m["noise_seed"], i, 0, 0, self.B, self.vocab.S, self.vocab.C)
        if m["sampling"] == "gumbel":
            ug = O.uniforms(m["noise_seed"], i, 1, 0, self.B, self.vocab.S, self.vocab.C)
        return u, ug
