"""CPU restatement of the reference's LaneIDM / lane-graph stepping — TEST INFRASTRUCTURE.

Pure-Python loops over the CSR arrays of lcsim_b200.city.CityStatic (small cases only).  PINNED: tests/golden_city/*.npz
holds per-tick outputs of the REAL reference engine on MOSS-shaped cities (oracle/gen_golden_city.py);
tests/test_city.py holds this restatement to them (integers bit-exact, floats 1e-9).  Only tests/ imports this module.

Every method cites the reference lines it follows, including the forced-lane-change branch
(idm_policy.py:64-84,118-167, agent.py:331-368) with its order-dependent reads of already-updated neighbours
(the loop below runs in list order and updates `v` in place, exactly like AgentManager.update).
`lc_required` records that some agent needed a lane change (the CUDA engine of round 1 flags instead of changing).
"""
from __future__ import annotations

import math

import numpy as np

WAITING, ACTIVE, FINISHED, IDLE = 0, 1, 2, 3
MIN_VIEW_DISTANCE, VIEW_DISTANCE_FACTOR, CLOSE_TO_END = 50.0, 12.0, 5.0  # agent.py:86-88
HEADWAY = 2.0  # idm_policy.py:35
H = 11


def _corners(b):  # polygon.py:238-268 corners_from_bboxes
    x, y, length, width, yaw = b[..., 0], b[..., 1], b[..., 2], b[..., 3], b[..., 4]
    cos, sin = np.cos(yaw), np.sin(yaw)
    xc = np.stack([x + length / 2 * cos + width / 2 * sin, x - length / 2 * cos + width / 2 * sin,
                   x - length / 2 * cos - width / 2 * sin, x + length / 2 * cos - width / 2 * sin], axis=-1)
    yc = np.stack([y + length / 2 * sin - width / 2 * cos, y - length / 2 * sin - width / 2 * cos,
                   y - length / 2 * sin + width / 2 * cos, y + length / 2 * sin + width / 2 * cos], axis=-1)
    return np.stack([xc, yc], axis=-1)


def _overlap_a_over_b(first, second):  # polygon.py:206-230
    c, s = np.cos(first[..., 4]), np.sin(first[..., 4])
    normals_t = np.stack([np.stack([c, -s], axis=-1), np.stack([s, c], axis=-1)], axis=-2)
    proj_a = np.matmul(_corners(first), normals_t)
    proj_b = np.matmul(_corners(second), normals_t)
    distance = np.minimum(proj_a.max(axis=-2), proj_b.max(axis=-2)) - np.maximum(proj_a.min(axis=-2), proj_b.min(axis=-2))
    return np.all(distance > 0, axis=-1)


def collision_counts(boxes):
    """AgentManager.check_collision_all (manager.py:313-327) for boxes (n,5) [x, y, l, w, heading]."""
    a, b = boxes[:, None, :], boxes[None, :, :]
    hit = np.logical_and(_overlap_a_over_b(a, b), _overlap_a_over_b(b, a))
    return np.where(np.eye(len(boxes), dtype=bool), False, hit).sum(axis=-1)


class CityOracle:
    def __init__(self, st, with_metrics=True):
        self.st = st
        self.N = st.n_agents
        self.with_metrics = with_metrics
        self.reset()

    # ------------------------------------------------------------------ geometry of a lane (lane.py:269-291)
    def _seg(self, lane, s):
        st = self.st
        a, b = st.lane_node_start[lane], st.lane_node_start[lane + 1]
        s = max(0.0, min(s, st.lane_length[lane]))
        i = 1
        for i in range(1, b - a):
            if s <= st.node_cum[a + i]:
                break
        return a + i, s

    def position(self, lane, s):
        st = self.st
        i, s = self._seg(lane, s)
        p1, p2 = st.node_xy[i - 1], st.node_xy[i]
        ratio = (s - st.node_cum[i - 1]) / (st.node_cum[i] - st.node_cum[i - 1])
        return p1 + ratio * (p2 - p1)

    def direction(self, lane, s):
        i, _ = self._seg(lane, s)
        return float(self.st.node_dir[i])

    # ------------------------------------------------------------------ Route (route.py:102-203)
    def _group(self, a, k):
        """junc_lane_groups[k] of agent a's route, or None"""
        st = self.st
        n_groups = st.route_start[a + 1] - st.route_start[a] - 1
        g = self.g_pop[a] + k
        return int(st.route_grp[st.route_start[a] - a + g]) if g < n_groups else None

    def junc_lane_by_pre_lane(self, a, pre_lane, k):  # route.py:191-203
        st = self.st
        g = self._group(a, k)
        if g is None:
            return None
        best, index = 1e6, 0
        for i in range(st.grp_start[g], st.grp_start[g + 1]):
            d = abs(int(st.lane_offset[pre_lane]) - int(st.grp_pre_offset[i]))
            if d < best:
                best, index = d, i
        return int(st.grp_lane[index])

    def get_lc(self, a, lane):  # route.py:138-189 -> (in_candidate, side, count, force_lc_length)
        st = self.st
        off = int(st.lane_offset[lane])
        g = self._group(a, 0)
        if g is None:
            delta = int(st.end_offset[a]) - off
            if delta == 0:
                return True, 0, 0, 0.0
            return (False, 0, -delta, 0.0) if delta < 0 else (False, 1, delta, 0.0)
        pre = [int(x) for x in st.grp_pre_offset[st.grp_start[g]:st.grp_start[g + 1]]]
        if pre[0] - off > 0 or pre[-1] - off < 0:
            min_delta, cnt = 1e6, 0
            for po in pre:
                delta = po - off
                if abs(delta) < abs(min_delta):
                    min_delta, cnt = delta, abs(delta)
            force = min(max(st.lane_max_speed[lane] * 3, 10.0), 50.0)  # LC_FACTOR, MIN/MAX_FORCE_LC_LENGTH
            return False, (0 if min_delta < 0 else 1), cnt, force
        return True, 0, 0, 0.0

    def lc_in_candidate(self, a, lane):
        return self.get_lc(a, lane)[0]

    def project_from_lane(self, lane, other, other_s):  # lane.py:293-300
        assert self.st.lane_road[lane] == self.st.lane_road[other] and self.st.lane_road[lane] >= 0, "not on the same road"
        s = other_s / self.st.lane_length[other] * self.st.lane_length[lane]
        return max(0.0, min(s, self.st.lane_length[lane]))

    def route_next(self, a, lane):  # route.py:102-136
        st = self.st
        if self.at_road[a]:
            if self._group(a, 0) is None:
                return None
            in_cand, side, _, _ = self.get_lc(a, lane)
            g = self._group(a, 0)
            if in_cand:
                nxt = self.junc_lane_by_pre_lane(a, lane, 0)
            else:  # left change pending: the group's rightmost lane is the nearest; right change: the leftmost
                self.lc_required = True
                nxt = int(st.grp_lane[st.grp_start[g + 1] - 1] if side == 0 else st.grp_lane[st.grp_start[g]])
            self.r_pop[a] += 1
        else:
            nxt = int(st.lane_succ0[lane])
            self.g_pop[a] += 1
        self.at_road[a] = not self.at_road[a]
        return nxt

    # ------------------------------------------------------------------ lanes' agent lists (lane.py:171-175,209-236)
    def _next_agent_by_s(self, lane, s):
        best = None
        for k, aid in self.lane_agents[lane].items():
            if k > s and (best is None or k < best[0]):
                best = (k, aid)
        return best

    def _first_agent(self, lane):
        d = self.lane_agents[lane]
        if not d:
            return None
        k = min(d)
        return (k, d[k])

    # ------------------------------------------------------------------ Engine.reset (engine.py:132-143, agent.py:106-171)
    def reset(self):
        st, N = self.st, self.N
        self.tick = 0
        self.status = np.full(N, WAITING, np.int32)
        self.lane = st.home_lane.astype(np.int32).copy()
        self.s = st.home_s.astype(np.float64).copy()
        self.v = np.zeros(N)
        self.xy = np.stack([self.position(int(st.home_lane[a]), float(st.home_s[a])) for a in range(N)]) if N else np.zeros((0, 2))
        self.h = np.asarray([self.direction(int(st.home_lane[a]), float(st.home_s[a])) for a in range(N)])
        self.r_pop = np.zeros(N, np.int32)
        self.g_pop = np.zeros(N, np.int32)
        self.at_road = np.ones(N, bool)
        self.lead_valid = np.zeros(N, np.uint8)
        self.lead_dis = np.zeros(N)
        self.lead_v = np.zeros(N)
        self.is_lc = np.zeros(N, np.uint8)  # LaneIDM.LCRuntime (idm_policy.py:18-27)
        self.lc_total = np.zeros(N)
        self.lc_completed = np.zeros(N)
        self.lc_target = np.full(N, -1, np.int32)
        self.lc_target_s = np.zeros(N)
        self.ring = np.zeros((N, H, 5))
        self.ring_valid = np.zeros(N, np.int32)
        for a in range(N):
            self._ring_push(a)
        self.active = []
        self.wait_ptr = 0
        self.lane_agents = [dict() for _ in range(st.n_lanes)]
        self.lane_new = [dict() for _ in range(st.n_lanes)]
        self.lc_required = False
        self.agent_steps = 0
        self.coll_count = np.zeros(N, np.int32)
        self.offroad = np.zeros(N, np.uint8)
        # Engine.reset (engine.py:139-143) checks departures and prepares observations but, unlike __init__ (:128-130)
        # and prepare() (:210), never calls lane_manager.prepare(): the departed agents stay in the lanes' pending
        # lists, nobody sees a leader at tick 0, and their home entries surface as ghosts one tick later.
        self._prepare(swap_lanes=False)

    def _ring_push(self, a):  # agent.py:67-84,159-165,419-426
        self.ring[a, :-1] = self.ring[a, 1:]
        self.ring[a, -1] = [self.xy[a, 0], self.xy[a, 1], self.v[a] * np.cos(self.h[a]), self.v[a] * np.sin(self.h[a]), self.h[a]]
        self.ring_valid[a] = min(H, self.ring_valid[a] + 1)

    # ------------------------------------------------------------------ Engine.prepare (engine.py:203-218)
    def _prepare(self, swap_lanes=True):
        st = self.st
        # check_departure_and_arrival (manager.py:50-96); single-trip schedules: FINISHED -> IDLE (schedule.py:22-40)
        arrival = [i for i, a in enumerate(self.active) if self.status[a] == FINISHED]
        for i in arrival:
            self.status[self.active[i]] = IDLE
        moved = 0
        while self.wait_ptr < self.N and st.depart_tick[st.wait_order[self.wait_ptr]] <= self.tick:
            a = int(st.wait_order[self.wait_ptr])
            self.wait_ptr += 1
            self.status[a] = ACTIVE
            self.lane_new[self.lane[a]][float(self.s[a])] = a  # add_to_lane (agent.py:454-460)
            if moved < len(arrival):
                self.active[arrival[moved]] = a
            else:
                self.active.append(a)
            moved += 1
        for i in sorted(arrival[moved:], reverse=True):
            self.active.pop(i)
        # lane_manager.prepare (lane.py:171-175)
        if swap_lanes:
            self.lane_agents, self.lane_new = self.lane_new, [dict() for _ in range(st.n_lanes)]
        # prepare_obs, LaneIDM branch (agent.py:198-243)
        for a in self.active:
            lane, s = int(self.lane[a]), float(self.s[a])
            view = max(MIN_VIEW_DISTANCE, self.v[a] * VIEW_DISTANCE_FACTOR)
            look = st.lane_length[lane] - s
            cur, junc_index, ahead_lanes = lane, 0, []
            while look < view:
                if st.lane_junction[cur] >= 0:
                    cur = int(st.lane_succ0[cur])
                    junc_index += 1
                else:
                    cur = self.junc_lane_by_pre_lane(a, cur, junc_index)
                if cur is None:
                    break
                ahead_lanes.append((look, cur))
                look += st.lane_length[cur]
            obs = None
            ahead = self._next_agent_by_s(lane, s)
            if ahead is not None:
                # NB reference: ahead[0] is the leader's ABSOLUTE s on the lane, not the gap (agent.py:227-230)
                obs = (ahead[0] - (st.length[a] + st.length[ahead[1]]) / 2.0, self.v[ahead[1]])
            else:
                for look_dis, al in ahead_lanes:
                    ahead = self._first_agent(al)
                    if ahead is not None:
                        obs = (look_dis + ahead[0] - (st.length[a] + st.length[ahead[1]]) / 2.0, self.v[ahead[1]])
                        break
            self.lead_valid[a] = obs is not None
            self.lead_dis[a], self.lead_v[a] = obs if obs is not None else (math.inf, 0.0)
        if self.with_metrics:
            self.metrics()

    # ------------------------------------------------------------------ LaneIDM.get_action (idm_policy.py:53-116)
    def _car_follow(self, a, self_v, target_v, ahead_v, distance):
        ub, mb, ma, _, gap = self.st.idm[a]
        if distance < 0:
            acc = -math.inf
        else:
            s_star = gap + max(self_v * HEADWAY + self_v * (self_v - ahead_v) / (2 * math.sqrt(-ub * ma)), 0)
            acc = ma * (1 - (self_v / target_v) ** 4 - (s_star / distance) ** 2)
        return max(mb, min(ma, acc))

    # ------------------------------------------------------------------ Engine.step (engine.py:145-158)
    def step(self):
        st = self.st
        dt = st.interval
        self.agent_steps += len(self.active)
        for a in self.active:  # AgentManager.update (manager.py:105-115)
            lane = int(self.lane[a])
            distance, ahead_v = (self.lead_dis[a], self.lead_v[a]) if self.lead_valid[a] else (math.inf, 0.0)
            acc = self._car_follow(a, self.v[a], min(st.idm[a][3], st.lane_max_speed[lane]), ahead_v, distance)
            if self.is_lc[a]:  # lane changing: also follow the leader on the target lane (idm_policy.py:67-84)
                tl = int(self.lc_target[a])
                ahead = self._next_agent_by_s(tl, float(self.lc_target_s[a]))
                if ahead is not None:
                    # ahead_a.runtime.v is read LIVE: already this tick's value if that agent precedes a in the list
                    d2 = ahead[0] - (st.length[a] + st.length[ahead[1]]) / 2.0
                    acc = min(acc, self._car_follow(a, self.v[a], st.lane_max_speed[tl], self.v[ahead[1]], d2))
            if not self.is_lc[a] and st.lane_road[lane] >= 0:  # _policy_lane_change (idm_policy.py:118-167)
                cur_s, cur_v = float(self.s[a]), float(self.v[a])
                reverse_s = st.lane_length[lane] - cur_s
                lc_length = max(3 * cur_v, st.length[a])  # LC_LENGTH_FACTOR
                in_cand, side, count, force = self.get_lc(a, lane)
                if not in_cand and reverse_s - force <= lc_length * count:
                    self.lc_required = True
                    tl = int(st.lane_left[lane] if side == 0 else st.lane_right[lane])
                    assert tl >= 0
                    neighbor_s = self.project_from_lane(tl, lane, cur_s)
                    ahead = self._next_agent_by_s(tl, neighbor_s)
                    if ahead is not None:
                        d2 = ahead[0] - (st.length[a] + st.length[ahead[1]]) / 2.0
                        acc = min(acc, self._car_follow(a, cur_v, st.lane_max_speed[tl], self.v[ahead[1]], d2))
                    self.is_lc[a], self.lc_total[a], self.lc_completed[a] = 1, lc_length, 0.0
                    self.lc_target[a], self.lc_target_s[a] = tl, neighbor_s
            # update_runtime_by_action, LANE branch (agent.py:314-376) + _compute_v_and_dis (:496-501)
            v0 = float(self.v[a])
            dv = acc * dt
            if v0 + dv < 0:
                v, ds = 0.0, -v0 * v0 / (2 * acc)
            else:
                v, ds = v0 + dv, (v0 + 0.5 * dv) * dt
            self.v[a] = v
            s = float(self.s[a]) + ds
            cur = lane
            if s > st.lane_length[lane]:
                while s > st.lane_length[cur]:
                    s -= st.lane_length[cur]
                    nxt = self.route_next(a, cur)
                    if nxt is None:
                        break
                    cur = nxt
            self.lane[a], self.s[a] = cur, s
            if self.is_lc[a]:  # agent.py:331-346
                tl = int(self.lc_target[a])
                if self.lc_completed[a] + ds >= self.lc_total[a]:
                    self.s[a] = self.project_from_lane(tl, cur, s)
                    self.lane[a] = tl
                    self.is_lc[a], self.lc_total[a], self.lc_completed[a], self.lc_target[a], self.lc_target_s[a] = 0, 0.0, 0.0, -1, 0.0
                else:
                    self.lc_completed[a] += ds
                    self.lc_target_s[a] = self.project_from_lane(tl, cur, s)
                    self.lane_new[tl][float(self.lc_target_s[a])] = a
            xy = self.position(cur, s)  # the lane walked on, even when the change just completed (agent.py:347-348)
            heading = self.direction(cur, s)
            if self.is_lc[a]:  # agent.py:349-369: blend towards the target lane
                tl = int(self.lc_target[a])
                xy_lc = self.position(tl, float(self.lc_target_s[a]))
                ratio = self.lc_completed[a] / self.lc_total[a]
                xy = xy * (1 - ratio) + xy_lc * ratio
                bias = np.arctan2((st.lane_width[cur] + st.lane_width[tl]) / 2.0, self.lc_total[a]) * (1 - abs(ratio * 2 - 1))
                if tl == st.lane_left[cur]:
                    heading += bias
                elif tl == st.lane_right[cur]:
                    heading -= bias
                else:
                    raise ValueError("lane changing error")
                heading = float(np.arctan2(np.sin(heading), np.cos(heading)))
            self.xy[a] = xy
            self.h[a] = heading
            cur = int(self.lane[a])
            s = float(self.s[a])
            # check_close_to_end (agent.py:434-441): returns before add_to_lane and the history append
            if st.lane_road[cur] >= 0 and st.lane_road[cur] == st.end_road[a] and st.end_s[a] - s < CLOSE_TO_END:
                self.status[a] = FINISHED
                continue
            self.lane_new[cur][s] = a
            self._ring_push(a)
        self._prepare()
        self.tick += 1

    # ------------------------------------------------------------------ metrics (manager.py:313-345, agent.py:482-494)
    def metrics(self):
        st = self.st
        self.coll_count[:] = 0
        self.offroad[:] = 0
        act = self.active
        n = len(act)
        if n == 0:
            return
        boxes = np.stack([self.xy[act, 0], self.xy[act, 1], st.length[act], st.width[act], self.h[act]], axis=1)
        self.coll_count[act] = collision_counts(boxes)
        for a in act:
            es = int(st.lane_eset[self.lane[a]])
            pts = st.edge_xy[st.eset_start[es]:st.eset_start[es + 1]]
            if len(pts) == 0:
                continue  # the reference would raise on an empty edge set (road.py:357); generator never makes one
            k = int(np.argmin(np.linalg.norm(pts - self.xy[a], axis=1)))
            d = st.edge_dir[st.eset_start[es] + k]
            p = pts[k] - self.xy[a]
            self.offroad[a] = (p[0] * d[1] - p[1] * d[0]) < 0
