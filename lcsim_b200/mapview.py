"""Read-only views with the reference's map-manager attribute surface (``engine.lane_manager / road_manager /
junction_manager``, lcsim/engine/engine.py:27-51,100-107).

The CUDA engine keeps the map as packed arrays (road-edge points, candidate lists, polyline centres); callers of the
reference that only READ the map (renderers, feature builders) get the same names here: ``lane_manager.lanes[id]`` /
``get_lane_by_id`` (lane/manager.py:14-29), ``road_manager.roads`` (road/manager.py:16-27, empty for WOMD-shaped
scenarios: waymo_convert.py writes zero roads), ``junction_manager.junctions[id]`` / ``get_junction_by_id`` /
``get_unique_junction`` (junction/manager.py:15-35).  Nothing here computes on the CPU for the stepping path: geometry
queries (nearest edge, off-road, collisions) go through the engine (``Agent.is_offroad``, ``AgentManager.*``).
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np


class LaneView:
    """reference: lane/lane.py:20-78 (id, centre line, length, parent)"""

    def __init__(self, lane_id: int, nodes: np.ndarray, parent_junction: "JunctionView"):
        self.id = int(lane_id)
        self.center_line = np.asarray(nodes, np.float64)
        self.center_line.setflags(write=False)
        seg = np.diff(self.center_line, axis=0)
        self.line_lengths = np.concatenate([[0.0], np.cumsum(np.linalg.norm(seg, axis=1))]) if len(seg) else np.zeros(1)
        self.length = float(self.line_lengths[-1])
        self.parent_junction = parent_junction
        self.parent_road = None

    def in_junction(self) -> bool:
        return self.parent_junction is not None


class LaneManagerView:
    def __init__(self):
        self.lanes: Dict[int, LaneView] = {}

    def get_lane_by_id(self, lane_id: int) -> Optional[LaneView]:
        return self.lanes.get(lane_id)


class RoadManagerView:
    def __init__(self):
        self.roads: Dict[int, object] = {}

    def get_road_by_id(self, road_id: int):
        return self.roads.get(road_id)


class JunctionView:
    """reference: junction/junction.py:40-89 (id, lanes, road lines / edges, the 0.5 m-resampled edge points)"""

    def __init__(self, junction_id: int, sc):
        self.id = int(junction_id)
        self.lanes: Dict[int, LaneView] = {}
        self.road_lines = [(np.asarray(p, np.float64), int(t)) for p, t in sc.lines]
        self.road_edges = [(np.asarray(p, np.float64), int(t)) for p, t in sc.edges_raw]
        # junction.py:59-78: what check_offroad searches (concatenated resampled edge points + unit directions)
        self.edge_points, self.edge_dirs = sc.edge_xy, sc.edge_dir
        self.edge_polyline_start = sc.edge_poly_start

    @property
    def lane_ids(self) -> List[int]:
        return list(self.lanes.keys())


class JunctionManagerView:
    def __init__(self):
        self.junctions: Dict[int, JunctionView] = {}

    def get_junction_by_id(self, junction_id: int) -> Optional[JunctionView]:
        return self.junctions.get(junction_id)

    def get_unique_junction(self) -> JunctionView:
        assert len(self.junctions) == 1  # junction/manager.py:33-35
        return next(iter(self.junctions.values()))


def build(sc):
    """(lane_manager, road_manager, junction_manager) views of one packed scenario (packer.ScenarioArrays)."""
    lm, rm, jm = LaneManagerView(), RoadManagerView(), JunctionManagerView()
    j = JunctionView(getattr(sc, "junction_id", 0), sc)
    jm.junctions[j.id] = j
    for nodes, lane_id in sc.lanes:
        lv = LaneView(lane_id, nodes, j)
        lm.lanes[lv.id] = lv
        j.lanes[lv.id] = lv
    return lm, rm, jm
