"""Runtime protobuf schema for the LCSim scenario format (``map.pb`` / ``agents.pb``).

The scenario format is part of the drop-in boundary: the same
bytes must feed the reference engine and this one.  The reference ships
protoc-generated modules (reference: lcsim/protos/{geo,agent,map,routing}/*.proto,
lcsim/entity/gen/*); ``protoc`` is not available here and generated code is not
ours to copy, so the wire schema (message names, field numbers, wire types —
facts of the format) is restated below as a small table and turned into message
classes with protobuf's descriptor API at import time.  Only fields the hot path
or the scenario writer touch are declared; unknown fields are skipped by the
protobuf runtime when parsing, which is what keeps real converted scenarios
loadable.
"""
from __future__ import annotations

from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

_T = descriptor_pb2.FieldDescriptorProto

_SCALARS = {
    "double": _T.TYPE_DOUBLE,
    "int32": _T.TYPE_INT32,
    "bool": _T.TYPE_BOOL,
    "string": _T.TYPE_STRING,
}

# message -> [(name, number, type, flags)]  flags: "r" repeated, "o" proto3 optional
# type is a scalar name, or ".pkg.Message" / "enum" (enums travel as int32 varints;
# declaring them int32 is wire-compatible and keeps this table short).
_SCHEMA = {
    "geo": {
        "XYPosition": [("x", 1, "double", ""), ("y", 2, "double", ""), ("z", 3, "double", "o")],
        "LongLatPosition": [("longitude", 1, "double", ""), ("latitude", 2, "double", ""), ("z", 3, "double", "o")],
        "LanePosition": [("lane_id", 1, "int32", ""), ("s", 2, "double", "")],
        "Position": [
            ("lane_position", 1, ".geo.LanePosition", "o"),
            ("longlat_position", 2, ".geo.LongLatPosition", "o"),
            ("xy_position", 3, ".geo.XYPosition", "o"),
        ],
    },
    "routing": {
        "DrivingJourneyBody": [("road_ids", 2, "int32", "r"), ("eta", 3, "double", "")],
        "Journey": [("type", 1, "enum", ""), ("driving", 2, ".routing.DrivingJourneyBody", "o")],
    },
    "agent": {
        "AgentAttribute": [
            ("type", 1, "enum", ""),
            ("length", 2, "double", ""),
            ("width", 3, "double", ""),
            ("max_speed", 4, "double", ""),
            ("max_acceleration", 5, "double", ""),
            ("max_braking_acceleration", 6, "double", ""),
            ("usual_acceleration", 7, "double", ""),
            ("usual_braking_acceleration", 8, "double", ""),
        ],
        "VehicleAttribute": [("lane_change_length", 1, "double", ""), ("min_gap", 2, "double", "")],
        "AgentState": [
            ("position", 1, ".geo.XYPosition", ""),
            ("velocity_x", 2, "double", ""),
            ("velocity_y", 3, "double", ""),
            ("heading", 4, "double", ""),
        ],
        "Trip": [
            ("mode", 1, "enum", ""),
            ("end", 2, ".geo.Position", ""),
            ("departure_time", 3, "double", "o"),
            ("wait_time", 4, "double", "o"),
            ("arrival_time", 5, "double", "o"),
            ("activity", 6, "string", "o"),
            ("routes", 7, ".routing.Journey", "r"),
            ("agent_states", 8, ".agent.AgentState", "r"),
        ],
        "Schedule": [
            ("trips", 1, ".agent.Trip", "r"),
            ("loop_count", 2, "int32", ""),
            ("departure_time", 3, "double", "o"),
            ("wait_time", 4, "double", "o"),
        ],
        "Agent": [
            ("id", 1, "int32", ""),
            ("attribute", 2, ".agent.AgentAttribute", ""),
            ("home", 3, ".geo.Position", ""),
            ("schedules", 4, ".agent.Schedule", "r"),
            ("vehicle_attribute", 5, ".agent.VehicleAttribute", "o"),
        ],
        "Agents": [("agents", 1, ".agent.Agent", "r"), ("ads_index", 2, "int32", "o")],
    },
    "map": {
        "Polyline": [("nodes", 1, ".geo.XYPosition", "r")],
        "Header": [
            ("name", 1, "string", ""),
            ("date", 2, "string", ""),
            ("north", 3, "double", ""),
            ("south", 4, "double", ""),
            ("east", 5, "double", ""),
            ("west", 6, "double", ""),
            ("projection", 7, "string", ""),
        ],
        "LaneConnection": [("id", 1, "int32", ""), ("type", 2, "enum", "")],
        "Lane": [
            ("id", 1, "int32", ""),
            ("type", 2, "enum", ""),
            ("turn", 3, "enum", ""),
            ("max_speed", 4, "double", ""),
            ("length", 5, "double", ""),
            ("width", 6, "double", ""),
            ("center_line", 7, ".map.Polyline", ""),
            ("predecessors", 10, ".map.LaneConnection", "r"),
            ("successors", 11, ".map.LaneConnection", "r"),
            ("left_lane_ids", 12, "int32", "r"),
            ("right_lane_ids", 13, "int32", "r"),
            ("parent_id", 14, "int32", ""),
            ("center_line_type", 16, "enum", ""),
        ],
        "RoadLine": [("type", 1, "enum", ""), ("nodes", 2, ".geo.XYPosition", "r")],
        "RoadEdge": [("type", 1, "enum", ""), ("nodes", 2, ".geo.XYPosition", "r")],
        "RoadConnection": [
            ("type", 1, "enum", ""),
            ("nodes", 2, ".geo.XYPosition", "r"),
            ("road_id", 3, "int32", ""),
        ],
        "Road": [
            ("id", 1, "int32", ""),
            ("lane_ids", 2, "int32", "r"),
            ("name", 4, "string", ""),
            ("road_lines", 5, ".map.RoadLine", "r"),
            ("road_edges", 6, ".map.RoadEdge", "r"),
            ("road_connections", 7, ".map.RoadConnection", "r"),
        ],
        "JunctionLaneGroup": [
            ("in_road_id", 1, "int32", ""),
            ("in_angle", 2, "double", ""),
            ("out_road_id", 3, "int32", ""),
            ("out_angle", 4, "double", ""),
            ("lane_ids", 5, "int32", "r"),
            ("turn", 6, "enum", ""),
        ],
        "Junction": [
            ("id", 1, "int32", ""),
            ("lane_ids", 2, "int32", "r"),
            ("driving_lane_groups", 3, ".map.JunctionLaneGroup", "r"),
            ("road_lines", 6, ".map.RoadLine", "r"),
            ("road_edges", 7, ".map.RoadEdge", "r"),
        ],
        "Map": [
            ("header", 1, ".map.Header", ""),
            ("lanes", 2, ".map.Lane", "r"),
            ("roads", 3, ".map.Road", "r"),
            ("junctions", 4, ".map.Junction", "r"),
        ],
    },
}

# enum values the path reads (reference: protos/agent/agent.proto:10-16, map.proto:150-160)
AGENT_TYPE_UNSPECIFIED = 0
AGENT_TYPE_VEHICLE = 1
AGENT_TYPE_PEDESTRIAN = 2
AGENT_TYPE_CYCLIST = 3
AGENT_TYPE_OTHER = 4
ROAD_EDGE_TYPE_BOUNDARY = 1
ROAD_EDGE_TYPE_MEDIAN = 2
LANE_TYPE_DRIVING = 1
LANE_TURN_STRAIGHT = 1
CENTER_LINE_TYPE_SURFACE_STREET = 2
TRIP_MODE_DRIVE_ONLY = 2
JOURNEY_TYPE_DRIVING = 1

_DEPS = {"geo": [], "routing": [], "agent": ["geo", "routing"], "map": ["geo"]}


def _build_pool() -> descriptor_pool.DescriptorPool:
    pool = descriptor_pool.DescriptorPool()
    for pkg in ("geo", "routing", "agent", "map"):
        fd = descriptor_pb2.FileDescriptorProto()
        fd.name = f"lcsim_b200/{pkg}.proto"
        fd.package = pkg
        fd.syntax = "proto3"
        for dep in _DEPS[pkg]:
            fd.dependency.append(f"lcsim_b200/{dep}.proto")
        for msg_name, fields in _SCHEMA[pkg].items():
            m = fd.message_type.add()
            m.name = msg_name
            for name, number, typ, flags in fields:
                f = m.field.add()
                f.name = name
                f.number = number
                f.label = _T.LABEL_REPEATED if "r" in flags else _T.LABEL_OPTIONAL
                if typ == "enum":
                    f.type = _T.TYPE_INT32
                elif typ.startswith("."):
                    f.type = _T.TYPE_MESSAGE
                    f.type_name = typ
                else:
                    f.type = _SCALARS[typ]
                if "o" in flags:
                    f.proto3_optional = True
                    f.oneof_index = len(m.oneof_decl)
                    m.oneof_decl.add().name = "_" + name
        pool.Add(fd)
    return pool


_POOL = _build_pool()


def _cls(full_name: str):
    return message_factory.GetMessageClass(_POOL.FindMessageTypeByName(full_name))


Map = _cls("map.Map")
Lane = _cls("map.Lane")
Junction = _cls("map.Junction")
Road = _cls("map.Road")
Agents = _cls("agent.Agents")
Agent = _cls("agent.Agent")
Trip = _cls("agent.Trip")
Schedule = _cls("agent.Schedule")
AgentState = _cls("agent.AgentState")


def load_map(path: str):
    m = Map()
    with open(path, "rb") as f:
        m.ParseFromString(f.read())
    return m


def load_agents(path: str):
    a = Agents()
    with open(path, "rb") as f:
        a.ParseFromString(f.read())
    return a
