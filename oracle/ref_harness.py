"""Runs the REAL reference engine (imported from /root/reference) and dumps its per-tick state.

TEST INFRASTRUCTURE.  Works only where /root/reference exists (the build container);
nothing under tests/ -m gpu, smoke() or bench.py imports this module.  It is used by
oracle/gen_golden.py to freeze golden vectors under tests/golden/, and by
tests/test_oracle_vs_reference.py (skipped when the reference is absent).

The reference needs four small shims to import in this image:
``np.bool8`` (removed in NumPy 2), and stub modules for torch_geometric.data,
matplotlib and motion_diff (none of which the stepping path executes).
"""
from __future__ import annotations

import copy
import os
import sys
import types
import warnings

import numpy as np

REFERENCE_ROOT = os.environ.get("LCSIM_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "lcsim"))


class _Any:
    def __init__(self, *a, **k):
        pass

    def __getattr__(self, k):
        return _Any()

    def __call__(self, *a, **k):
        return _Any()


def _install_shims():
    warnings.simplefilter("ignore")
    if not hasattr(np, "bool8"):
        np.bool8 = np.bool_

    def stub(name, **kw):
        if name in sys.modules:
            return sys.modules[name]
        m = types.ModuleType(name)
        m.__dict__.update(kw)
        sys.modules[name] = m
        return m

    stub("torch_geometric")
    stub("torch_geometric.data", Batch=_Any, HeteroData=_Any, Dataset=_Any)
    m = stub("matplotlib", axes=_Any(), figure=_Any())
    stub("matplotlib.patheffects")
    m.pyplot = stub("matplotlib.pyplot")
    stub("motion_diff")
    stub("motion_diff.model", MotionDiff=_Any)
    stub("motion_diff.modules")
    stub("motion_diff.modules.guidance", RealGuidance=_Any)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def load_reference():
    """Returns the reference's (Engine, Action) classes."""
    _install_shims()
    from lcsim.engine.engine import Engine  # type: ignore
    from lcsim.entity.agent.policy import Action  # type: ignore

    return Engine, Action


DEFAULT_CONFIG = {
    "input": {"data_dir": "", "model_path": "none", "model_config_path": "none"},
    "control": {
        "step": {"start": 0, "total": 91, "interval": 0.1},
        "enable_poly_emb": False, "enable_plan": False, "plan_interval": 80, "plan_init_step": 10,
        "device": "cpu", "allow_new_obj": True, "reactive": False, "no_static": False, "guidance": False,
    },
    "reward": {"forward": 2, "collision": 5, "smooth": 1, "destination": 1},
    "render": {"center_id": "ads", "center_type": "agent", "fig_size": [8, 8], "range": [-75, 75, -75, 75],
               "plot_id": False, "plot_ref_traj": False, "axises": False},
}

REWARD_KEYS = ["forward", "route_progress", "smooth", "collision", "offroad", "destination"]
STATUS = {"WAITING": 0, "ACTIVE": 1, "FINISHED": 2, "IDLE": 3}


def make_config(data_dir: str, **control) -> dict:
    cfg = copy.deepcopy(DEFAULT_CONFIG)
    cfg["input"]["data_dir"] = data_dir
    cfg["control"].update(control)
    return cfg


def run_reference(data_dir: str, n_steps: int = 90, ads_actions: np.ndarray | None = None,
                  replan: dict | None = None, quiet: bool = True, **control) -> dict:
    """Steps the reference engine n_steps times and records, AFTER each step:

    active list (agent indices in list order), status/x/y/h/v/cursor of every agent,
    check_collision_all() counts, junction.check_offroads flags + nearest edge index for the
    active agents, TrajIDM lead observations, and for the ADS: _compute_rewards terms,
    terminated, check_collision, is_offroad, get_route.

    ads_actions: optional (n_steps, 2) normalised BICYCLE actions for the ADS
    (reference: envs/gym_warpper.py:78-87).
    replan: optional {"tick": k, "xy": (A,T,2) f32, "vel": ..., "heading": ...} applied through
    Agent.set_ref_trajectory to dynamic agents before step k (reference: engine.py:293-297).
    """
    Engine, Action = load_reference()
    cfg = make_config(data_dir, **control)
    if quiet:
        import contextlib
        import io

        with contextlib.redirect_stdout(io.StringIO()):
            eng = Engine(cfg)
    else:
        eng = Engine(cfg)
    eng.reset()
    am = eng.agent_manager
    ids = list(am.agents.keys())
    index_of = {aid: i for i, aid in enumerate(ids)}
    A = len(ids)
    ads = am.get_ads()
    junc = eng.junction_manager.get_unique_junction()
    edge_pts = np.array([p for e in junc.road_edges for p in e[0]]).reshape(-1, 2)

    K = n_steps
    out = dict(
        n_active=np.zeros(K + 1, np.int32), active=np.full((K + 1, A), -1, np.int32),
        status=np.zeros((K + 1, A), np.int8), cursor=np.zeros((K + 1, A), np.int32),
        x=np.zeros((K + 1, A)), y=np.zeros((K + 1, A)), h=np.zeros((K + 1, A)), v=np.zeros((K + 1, A)),
        coll=np.zeros((K + 1, A), np.int32), offroad=np.zeros((K + 1, A), np.uint8),
        nearest=np.full((K + 1, A), -1, np.int32),
        lead_valid=np.zeros((K + 1, A), np.uint8), lead_dis=np.zeros((K + 1, A)), lead_v=np.zeros((K + 1, A)),
        rewards=np.zeros((K + 1, 6)), terminated=np.zeros(K + 1, np.uint8), ads_collision=np.zeros(K + 1, np.uint8),
        ads_offroad=np.zeros(K + 1, np.uint8), route=np.zeros((K + 1, 10, 2)), n_updates=np.zeros(K + 1, np.int64),
        traj_d=np.zeros((K + 1, A)), traj_s=np.zeros((K + 1, A)),
    )

    def record(k):
        act = am.active_agents
        out["n_active"][k] = len(act)
        for i, a in enumerate(act):
            out["active"][k, i] = index_of[a.id]
        for aid, a in am.agents.items():
            i = index_of[aid]
            out["status"][k, i] = a.runtime.status.value
            out["cursor"][k, i] = a.ref_trajectory.index
            out["traj_d"][k, i] = a.ref_trajectory.d
            out["traj_s"][k, i] = a.ref_trajectory.s
            out["x"][k, i], out["y"][k, i] = a.runtime.xy[0], a.runtime.xy[1]
            out["h"][k, i], out["v"][k, i] = a.runtime.heading, a.runtime.v
        if len(act):
            cc = am.check_collision_all()
            poses = am.get_agent_bbox([a.id for a in act])[..., :2]
            off = junc.check_offroads(poses) if len(edge_pts) else np.zeros(len(act), bool)
            for i, a in enumerate(act):
                j = index_of[a.id]
                out["coll"][k, j] = int(cc[i])
                out["offroad"][k, j] = bool(off[i])
                if len(edge_pts):
                    # same expression as junction.py:381 on the same concatenated point array
                    out["nearest"][k, j] = int(np.argmin(np.linalg.norm(edge_pts - a.runtime.xy, axis=1)))
                ahead = getattr(a.obs, "ahead_vehicle", None)
                if ahead is not None:
                    out["lead_valid"][k, j] = 1
                    out["lead_dis"][k, j], out["lead_v"][k, j] = ahead[0], ahead[1]
        rew, term = eng._compute_rewards(REWARD_KEYS, am.ads_id)
        out["rewards"][k] = [float(np.asarray(rew[key]).reshape(-1)[0]) if key in rew else 0.0 for key in REWARD_KEYS]
        out["terminated"][k] = bool(term)
        out["ads_collision"][k] = bool(am.check_collision(am.ads_id))
        out["ads_offroad"][k] = bool(ads.is_offroad())
        out["route"][k] = ads.get_route()

    record(0)  # state right after reset()
    for k in range(n_steps):
        out["n_updates"][k + 1] = len(am.active_agents)
        if replan is not None and replan["tick"] == k:
            for i, aid in enumerate(ids):
                a = am.agents[aid]
                if a.dynamic_flag and replan["mask"][i]:
                    a.set_ref_trajectory(replan["xy"][i], replan["vel"][i], replan["heading"][i])
        actions = {}
        if ads_actions is not None:
            data = Action.BicycleAction()
            data.init_with_normalized(float(ads_actions[k, 0]), float(ads_actions[k, 1]))
            actions = {ads.id: Action(type=Action.ActionType.BICYCLE, data=data)}
        eng.step(actions)
        record(k + 1)
    # final history rings
    ring = np.zeros((A, 11, 5))
    ring_valid = np.zeros(A, np.int32)
    for aid, a in am.agents.items():
        i = index_of[aid]
        hs = a.history_states
        ring[i, :, 0:2] = hs.xyz[:, :2]
        ring[i, :, 2:4] = hs.vel
        ring[i, :, 4] = hs.heading[:, 0]
        ring_valid[i] = int(hs.valid.sum())
    out["ring"] = ring
    out["ring_valid"] = ring_valid
    out["agent_ids"] = np.asarray(ids, np.int32)
    out["dynamic"] = np.asarray([bool(am.agents[i].dynamic_flag) for i in ids], np.uint8)
    return out
