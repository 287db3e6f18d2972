"""Helpers shared by the golden-vector tests: load a case, rebuild its StaticBatch."""
import glob
import os

import numpy as np

from lcsim_b200.packer import StaticBatch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def case_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


class Case:
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.name = name
        S, A, T, E = (int(v) for v in z["dims"])
        start, interval, total = z["clock"]
        kw = {k[len("static_"):]: np.ascontiguousarray(z[k]) for k in z.files if k.startswith("static_")}
        self.static = StaticBatch(S=S, A=A, T=T, E=E, start=float(start), interval=float(interval),
                                  total_steps=int(total), **kw)
        self.reactive, self.no_static, self.allow_new_obj = (bool(v) for v in z["control"])
        self.n_steps = int(z["n_steps"][0])
        self.ref = {k[len("ref_"):]: z[k] for k in z.files if k.startswith("ref_")}
        self.ads_actions = z["ads_actions"] if "ads_actions" in z.files else None
        self.replan = None
        if "replan_tick" in z.files:
            self.replan = dict(tick=int(z["replan_tick"][0]), xy=z["replan_xy"], vel=z["replan_vel"],
                               heading=z["replan_heading"], mask=z["replan_mask"])
        self.ads = int(self.static.ads_index[0])

    def actions_for(self, k):
        """(act_agent, act_type, act_data) arrays of shape (1,1[,5]) for tick k, or Nones."""
        if self.ads_actions is None:
            return None, None, None
        aa = np.full((1, 1), self.ads, np.int32)
        at = np.full((1, 1), 1, np.int32)  # BICYCLE
        ad = np.zeros((1, 1, 5))
        # reference: policy/type.py:100-102 init_with_normalized
        ad[0, 0, 0] = float(self.ads_actions[k, 0]) * 6.0
        ad[0, 0, 1] = float(self.ads_actions[k, 1]) * 0.3
        return aa, at, ad
