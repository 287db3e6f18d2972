// lcsim_city.cuh — LaneIDM stepping on a lane graph (BASELINE configs[3]: one MOSS-shaped city, up to
// 200k vehicles).  Included by lcsim_b200.cu; same two builds as the rest (nvcc sm_100a / g++ -DLCSIM_HOSTEMU).
//
// One scenario, N agents: the parallel axis is the agent (active-list slot), not the scenario, so a tick is a short
// sequence of flat kernels (one thread per slot / pending lane entry / grid cell) with two prefix sums in between:
//   update -> arrivals/departures (ordered compaction of the active list) -> lane buckets (the reference's per-lane
//   SortedDict, rebuilt per tick as a counting sort by lane) -> observations (leader search) -> metrics (uniform
//   spatial hash + float64 SAT, per-lane edge sets) -> clock.
// Every decision is the reference's float64 expression with pinned roundings (lcs_dmul / lcs_dadd / ...).
// The forced-lane-change branch (idm_policy.py:64-84,118-167, agent.py:331-368) reads neighbours' runtime.v LIVE
// (already updated when the neighbour precedes the agent in the list).  Agents without such a read are
// finished by the first update kernel; the few "readers" are resolved in dependency rounds (a reader waits only for
// readers that precede it in the list), then by a serial pass for whatever chain is left.
#pragma once

struct LccP {
    int32_t N, NL, n_cells, cell_w, cell_h;
    double dt, cell_x0, cell_y0, cell_inv;
    int32_t with_metrics;
    // ---- static: lanes
    const double *lane_length, *lane_max_speed, *lane_width;
    const int32_t *lane_road, *lane_junction, *lane_offset, *lane_succ0, *lane_node_start, *lane_eset, *lane_left, *lane_right;
    const double *node_xy, *node_cum, *node_dir;
    const int32_t* eset_start;
    const double *edge_xy, *edge_dir;
    const int32_t *grp_start, *grp_lane, *grp_pre_offset;
    // ---- static: agents
    const double *length, *width, *idm, *home_s, *end_s;
    const int32_t *home_lane, *end_road, *end_offset, *depart_tick, *wait_order, *route_start, *route_grp;
    // ---- state: agents
    int32_t *status, *lane, *r_pop, *g_pop, *at_road, *ring_head, *ring_count, *coll_count;
    double *s, *v, *xy, *h, *lead_dis, *lead_v, *ring;
    // lane change runtime (LaneIDM.LCRuntime, idm_policy.py:18-27) and the update's scratch
    int32_t *is_lc, *lc_target, *slot_of, *rd_ahead, *rd_lane, *done_round;
    double *lc_total, *lc_completed, *lc_target_s, *v2, *acc, *rd_dist;
    uint8_t *lead_valid, *offroad;
    // ---- state: lists
    int32_t *active, *active2, *fin, *rank; // [N]
    int32_t* scal; // [16] device scalars, see LCC_* below
    int64_t* stats; // [LCS_NSTAT]
    // pending lane entries (Lane.new_agents of all lanes) and the buckets built from them (Lane.agents)
    int32_t *ent_lane, *ent_agent, *ent_order; // [2N]
    double* ent_s;
    int32_t *lane_start, *lane_cur; // [NL+1]
    int32_t *bkt_agent, *bkt_order; // [2N]
    double* bkt_s;
    // spatial hash of the active agents
    int32_t *cell_start, *cell_cur, *cell_agent; // [n_cells+1], [n_cells+1], [N]
    double* cell_rec; // [N][8] collision records in cell order
    int32_t mode; // 1 = reset
    // 1: the per-agent kernels map thread -> AGENT index (its slot from slot_of) instead of thread -> list slot: every
    // per-agent array is then read and written coalesced (the list order is the order of departures, i.e. random
    // against the agent index: a slot-ordered warp touches 32 sectors per array).  LCSIM_B200_CITY_ORDER=slot: 0
    int32_t by_agent;
};
enum { LCC_N_ACTIVE = 0, LCC_WAIT_PTR, LCC_TICK, LCC_N_ENT, LCC_N_ARR, LCC_M_POP, LCC_LC_REQUIRED, LCC_LC_ERROR, LCC_N_OLD, LCC_N_PENDING,
       LCC_SYNC /* blocks of lcc_place_k that are done */ };
// thread t of a per-agent kernel -> (agent, slot) or false when it has none
LCS_D bool lcc_member(const LccP& c, int t, int* a, int* i) {
    if (c.by_agent) {
        *a = t;
        *i = c.slot_of[t];
        return *i >= 0;
    }
    if (t >= c.scal[LCC_N_ACTIVE]) return false;
    *i = t;
    *a = c.active[t];
    return true;
}
// the same after lcc_place (the new list is active2 until the end of the tick)
LCS_D bool lcc_member2(const LccP& c, int t, int* a) {
    if (c.by_agent) {
        *a = t;
        return c.slot_of[t] >= 0;
    }
    if (t >= c.scal[LCC_N_ACTIVE]) return false;
    *a = c.active2[t];
    return true;
}


#ifndef LCSIM_HOSTEMU
#define LCC_KERNEL(name)                                          \
    __global__ void name##_k(const LccP c, int n) {               \
        const int i = blockIdx.x * blockDim.x + threadIdx.x;      \
        if (i < n) name(c, i);                                    \
    }
// the same with a floor on the resident blocks per SM (a register cap): a kernel bound by the latency of its
// dependent loads wants all of its blocks resident at once, not one and a third waves of them
#define LCC_KERNEL_LB(name, minb)                                                     \
    __global__ void __launch_bounds__(256, minb) name##_k(const LccP c, int n) {      \
        const int i = blockIdx.x * blockDim.x + threadIdx.x;                          \
        if (i < n) name(c, i);                                                        \
    }
#define LCC_RUN(name, c, n, st)                                                                  \
    do {                                                                                         \
        if ((n) > 0) name##_k<<<((n) + 255) / 256, 256, 0, st>>>(c, n);                          \
    } while (0)
LCS_D double lcc_div(double a, double b) { return __ddiv_rn(a, b); }
LCS_D void atomicAdd_i64(int64_t* p, int64_t v) { atomicAdd((unsigned long long*)p, (unsigned long long)v); }
#else
#define LCC_KERNEL(name)
#define LCC_KERNEL_LB(name, minb)
#define LCC_RUN(name, c, n, st)                          \
    do {                                                 \
        for (int _i = 0; _i < (n); _i++) name(c, _i);    \
    } while (0)
LCS_D double lcc_div(double a, double b) { return a / b; }
LCS_D void atomicAdd_i64(int64_t* p, int64_t v) { *p += v; }
#endif

// ------------------------------------------------------------------ lane geometry (lane/lane.py:269-291)
LCS_D int lcc_segment(const LccP& c, int lane, double& s) {
    const int a = c.lane_node_start[lane], b = c.lane_node_start[lane + 1];
    const double L = c.lane_length[lane];
    s = lcs_pymax(0.0, lcs_pymin(s, L)); // max(0.0, min(s, self.length))
    int i = 1;
    for (int k = 1; k < b - a; k++) {
        i = k;
        if (s <= c.node_cum[a + k]) break;
    }
    return a + i;
}
LCS_D void lcc_position(const LccP& c, int lane, double s, double* x, double* y, double* dir) {
    const int i = lcc_segment(c, lane, s);
    const double x1 = c.node_xy[2 * (i - 1)], y1 = c.node_xy[2 * (i - 1) + 1], x2 = c.node_xy[2 * i], y2 = c.node_xy[2 * i + 1];
    const double ratio = lcc_div(lcs_dsub(s, c.node_cum[i - 1]), lcs_dsub(c.node_cum[i], c.node_cum[i - 1]));
    *x = lcs_dadd(x1, lcs_dmul(ratio, lcs_dsub(x2, x1)));
    *y = lcs_dadd(y1, lcs_dmul(ratio, lcs_dsub(y2, y1)));
    *dir = c.node_dir[i];
}

// ------------------------------------------------------------------ Route (agent/route.py:102-203)
LCS_D int lcc_group(const LccP& c, int a, int k) { // junc_lane_groups[k] of the remaining route, -1 = none
    const int n_groups = c.route_start[a + 1] - c.route_start[a] - 1;
    const int g = c.g_pop[a] + k;
    return g < n_groups ? c.route_grp[c.route_start[a] - a + g] : -1;
}
LCS_D int lcc_junc_lane_by_pre_lane(const LccP& c, int a, int pre_lane, int k) { // route.py:191-203
    const int g = lcc_group(c, a, k);
    if (g < 0) return -1;
    int best = 1000000, index = c.grp_start[g];
    const int off = c.lane_offset[pre_lane];
    for (int i = c.grp_start[g]; i < c.grp_start[g + 1]; i++) {
        int d = off - c.grp_pre_offset[i];
        d = d < 0 ? -d : d;
        if (d < best) { best = d; index = i; }
    }
    return c.grp_lane[index];
}
// Route.get_lc (route.py:138-189): in_candidate, else side (0 left, 1 right), count and force_lc_length
struct LccLC {
    bool in_candidate;
    int side, count;
    double force_len;
};
LCS_D LccLC lcc_get_lc(const LccP& c, int a, int lane) {
    LccLC r;
    r.in_candidate = true; r.side = 0; r.count = 0; r.force_len = 0.0;
    const int off = c.lane_offset[lane];
    const int g = lcc_group(c, a, 0);
    if (g < 0) { // reach the end: compare with the end lane
        const int delta = c.end_offset[a] - off;
        if (delta != 0) { r.in_candidate = false; r.side = delta < 0 ? 0 : 1; r.count = delta < 0 ? -delta : delta; }
        return r;
    }
    const int b = c.grp_start[g], e = c.grp_start[g + 1];
    if (c.grp_pre_offset[b] - off > 0 || c.grp_pre_offset[e - 1] - off < 0) {
        int min_delta = 1000000, cnt = 0;
        for (int i = b; i < e; i++) {
            const int delta = c.grp_pre_offset[i] - off, ad = delta < 0 ? -delta : delta;
            if (ad < (min_delta < 0 ? -min_delta : min_delta)) { min_delta = delta; cnt = ad; }
        }
        r.in_candidate = false;
        r.count = cnt;
        r.side = min_delta < 0 ? 0 : 1;
        r.force_len = lcs_pymin(lcs_pymax(lcs_dmul(c.lane_max_speed[lane], 3.0), 10.0), 50.0);
    }
    return r;
}
// Lane.project_from_lane (lane.py:293-300); the reference raises when the lanes are not on the same road
LCS_D double lcc_project_from_lane(const LccP& c, int lane, int other, double other_s) {
    if (c.lane_road[lane] < 0 || c.lane_road[lane] != c.lane_road[other]) c.scal[LCC_LC_ERROR] = 1;
    const double s = lcs_dmul(lcc_div(other_s, c.lane_length[other]), c.lane_length[lane]);
    return lcs_pymax(0.0, lcs_pymin(s, c.lane_length[lane]));
}
LCS_D int lcc_route_next(const LccP& c, int a, int lane) { // route.py:102-136
    int nxt;
    if (c.at_road[a]) {
        const int g = lcc_group(c, a, 0);
        if (g < 0) return -1;
        const LccLC lc = lcc_get_lc(c, a, lane);
        if (lc.in_candidate)
            nxt = lcc_junc_lane_by_pre_lane(c, a, lane, 0);
        else // a left change is pending: the group's rightmost lane is the nearest; a right change: its leftmost
            nxt = lc.side == 0 ? c.grp_lane[c.grp_start[g + 1] - 1] : c.grp_lane[c.grp_start[g]];
        c.r_pop[a] += 1;
    } else {
        nxt = c.lane_succ0[lane];
        c.g_pop[a] += 1;
    }
    c.at_road[a] = !c.at_road[a];
    return nxt;
}

// ------------------------------------------------------------------ helpers
LCS_D void lcc_add_to_lane(const LccP& c, int lane, double s, int a, int order) { // Lane.add_agent (lane.py:207-208)
    const int e = lcs_atomic_add(&c.scal[LCC_N_ENT], 1);
    c.ent_lane[e] = lane;
    c.ent_s[e] = s;
    c.ent_agent[e] = a;
    c.ent_order[e] = order;
}
LCS_D void lcc_ring_push(const LccP& c, int a, double v) { // agent.py:67-84,159-165,419-426 (circular instead of shifting)
    const int head = c.ring_head[a];
    double* r = c.ring + ((size_t)a * LCS_H + head) * 5;
    double sh, ch;
    lcs_sincos(c.h[a], &sh, &ch);
    r[0] = c.xy[2 * a];
    r[1] = c.xy[2 * a + 1];
    r[2] = lcs_dmul(v, ch);
    r[3] = lcs_dmul(v, sh);
    r[4] = c.h[a];
    c.ring_head[a] = head + 1 == LCS_H ? 0 : head + 1;
    if (c.ring_count[a] < LCS_H) c.ring_count[a] += 1;
}

// ------------------------------------------------------------------ reset: Agent.__init__ (agent.py:106-171)
LCS_D void lcc_reset_agent(const LccP& c, int a) {
    c.status[a] = LCS_WAITING;
    c.lane[a] = c.home_lane[a];
    c.s[a] = c.home_s[a];
    c.v[a] = 0.0;
    double x, y, dir;
    lcc_position(c, c.home_lane[a], c.home_s[a], &x, &y, &dir);
    c.xy[2 * a] = x;
    c.xy[2 * a + 1] = y;
    c.h[a] = dir;
    c.r_pop[a] = 0;
    c.g_pop[a] = 0;
    c.at_road[a] = 1;
    c.lead_valid[a] = 0;
    c.lead_dis[a] = lcs_inf_d();
    c.lead_v[a] = 0.0;
    c.is_lc[a] = 0;
    c.lc_total[a] = 0.0;
    c.lc_completed[a] = 0.0;
    c.lc_target[a] = -1;
    c.lc_target_s[a] = 0.0;
    c.slot_of[a] = -1;
    c.v2[a] = 0.0;
    c.coll_count[a] = 0;
    c.offroad[a] = 0;
    c.ring_head[a] = 0;
    c.ring_count[a] = 0;
    for (int k = 0; k < LCS_H * 5; k++) c.ring[(size_t)a * LCS_H * 5 + k] = 0.0;
    lcc_ring_push(c, a, 0.0);
    c.fin[a] = 0;
    c.active[a] = -1;
}
LCC_KERNEL(lcc_reset_agent)

// ------------------------------------------------------------------ update: LaneIDM.get_action + LANE branch
// idm_policy.py:92-116 _policy_car_follow
LCS_D double lcc_car_follow(const double* idm, double self_v, double target_v, double ahead_v, double distance) {
    const double ub = idm[0], mb = idm[1], ma = idm[2], gap = idm[4];
    double acc;
    if (distance < 0) {
        acc = -lcs_inf_d();
    } else {
        const double den = lcs_dmul(2.0, sqrt(lcs_dmul(-ub, ma)));
        const double t = lcs_dadd(lcs_dmul(self_v, 2.0 /* headway, idm_policy.py:35 */), lcc_div(lcs_dmul(self_v, lcs_dsub(self_v, ahead_v)), den));
        const double s_star = lcs_dadd(gap, lcs_pymax(t, 0.0));
        const double r1 = lcc_div(self_v, target_v), r2 = lcc_div(s_star, distance);
        acc = lcs_dmul(ma, lcs_dsub(lcs_dsub(1.0, lcs_pow4(r1)), lcs_dmul(r2, r2)));
    }
    return lcs_pymax(mb, lcs_pymin(ma, acc));
}
// forward declarations of the bucket lookups (lane.py:210-234), defined with the lane buckets below
LCS_D bool lcc_next_agent(const LccP& c, int lane, double s, double* ks, int* agent);

// everything after the action is known: update_runtime_by_action, LANE branch (agent.py:314-376, 419-426)
LCS_D void lcc_finalize(const LccP& c, int i, int a, double acc) {
    const int lane = c.lane[a];
    const double v0 = c.v[a];
    // _compute_v_and_dis (agent.py:496-501)
    const double dt = c.dt, dv = lcs_dmul(acc, dt);
    double v, ds;
    if (lcs_dadd(v0, dv) < 0) {
        v = 0.0;
        ds = lcc_div(lcs_dmul(-v0, v0), lcs_dmul(2.0, acc));
    } else {
        v = lcs_dadd(v0, dv);
        ds = lcs_dmul(lcs_dadd(v0, lcs_dmul(0.5, dv)), dt);
    }
    c.v2[a] = v; // c.v keeps the value of the tick start until lcc_commit_v: other agents may still have to read it
    double s = lcs_dadd(c.s[a], ds);
    int cur = lane;
    if (s > c.lane_length[lane]) {
        while (s > c.lane_length[cur]) {
            s = lcs_dsub(s, c.lane_length[cur]);
            const int nxt = lcc_route_next(c, a, cur);
            if (nxt < 0) break;
            cur = nxt;
        }
    }
    int new_lane = cur;
    double new_s = s;
    bool lc = c.is_lc[a] != 0;
    const int tl = c.lc_target[a];
    if (lc) { // agent.py:331-346
        if (lcs_dadd(c.lc_completed[a], ds) >= c.lc_total[a]) {
            new_s = lcc_project_from_lane(c, tl, cur, s);
            new_lane = tl;
            c.is_lc[a] = 0; c.lc_total[a] = 0.0; c.lc_completed[a] = 0.0; c.lc_target[a] = -1; c.lc_target_s[a] = 0.0;
            lc = false;
        } else {
            c.lc_completed[a] = lcs_dadd(c.lc_completed[a], ds);
            const double ts = lcc_project_from_lane(c, tl, cur, s);
            c.lc_target_s[a] = ts;
            lcc_add_to_lane(c, tl, ts, a, (1 << 28) | (i << 1)); // target_lane.add_agent, before the own lane's entry
        }
    }
    c.lane[a] = new_lane;
    c.s[a] = new_s;
    double x, y, dir;
    lcc_position(c, cur, s, &x, &y, &dir); // the lane walked on, even when the change just completed (agent.py:347-348)
    if (lc) { // agent.py:349-369: blend towards the target lane
        double xl, yl, dl;
        lcc_position(c, tl, c.lc_target_s[a], &xl, &yl, &dl);
        const double ratio = lcc_div(c.lc_completed[a], c.lc_total[a]), om = lcs_dsub(1.0, ratio);
        x = lcs_dadd(lcs_dmul(x, om), lcs_dmul(xl, ratio));
        y = lcs_dadd(lcs_dmul(y, om), lcs_dmul(yl, ratio));
        const double bias = lcs_dmul(atan2(lcc_div(lcs_dadd(c.lane_width[cur], c.lane_width[tl]), 2.0), c.lc_total[a]),
                                     lcs_dsub(1.0, fabs(lcs_dsub(lcs_dmul(ratio, 2.0), 1.0))));
        if (tl == c.lane_left[cur])
            dir = lcs_dadd(dir, bias);
        else if (tl == c.lane_right[cur])
            dir = lcs_dsub(dir, bias);
        else
            c.scal[LCC_LC_ERROR] = 1; // reference: ValueError("lane changing error")
        double sh, ch;
        lcs_sincos(dir, &sh, &ch);
        dir = atan2(sh, ch);
    }
    c.xy[2 * a] = x;
    c.xy[2 * a + 1] = y;
    c.h[a] = dir;
    // check_close_to_end (agent.py:434-441): returns before add_to_lane and the history append
    if (c.lane_road[new_lane] >= 0 && c.lane_road[new_lane] == c.end_road[a] && lcs_dsub(c.end_s[a], new_s) < 5.0) {
        c.status[a] = LCS_IDLE; // FINISHED now, IDLE after this tick's check_departure_and_arrival (single-trip schedules)
        c.fin[i] = 1;
        c.coll_count[a] = 0;
        c.offroad[a] = 0;
        return;
    }
    c.fin[i] = 0;
    lcc_add_to_lane(c, new_lane, new_s, a, (1 << 28) | (i << 1) | 1);
    lcc_ring_push(c, a, v); // history append with the new speed (agent.py:419-426 reads runtime.v after the update)
}
// slot i of the active list: manager.py:105-115 -> LaneIDM.get_action (idm_policy.py:53-90).  Agents whose action
// needs no live read of another agent are finished here; the others leave (acc, whom to read) for the rounds.
LCS_D void lcc_update(const LccP& c, int t) {
    if (t >= c.scal[LCC_N_ACTIVE]) c.fin[t] = 0; // (t as a slot beyond the list)
    int a, i;
    if (!lcc_member(c, t, &a, &i)) return;
    const int lane = c.lane[a];
    const double* idm = c.idm + 5 * (size_t)a;
    const bool lv = c.lead_valid[a] != 0;
    const double distance = lv ? c.lead_dis[a] : lcs_inf_d(), ahead_v = lv ? c.lead_v[a] : 0.0;
    const double v0 = c.v[a];
    const double acc = lcc_car_follow(idm, v0, lcs_pymin(idm[3], c.lane_max_speed[lane]), ahead_v, distance);
    int tl = -1;
    double ts = 0.0;
    if (c.is_lc[a]) { // lane changing: also follow the leader on the target lane (idm_policy.py:67-84)
        tl = c.lc_target[a];
        ts = c.lc_target_s[a];
    } else if (c.lane_road[lane] >= 0) { // _policy_lane_change (idm_policy.py:118-167)
        const double cur_s = c.s[a];
        const double reverse_s = lcs_dsub(c.lane_length[lane], cur_s);
        const double lc_length = lcs_pymax(lcs_dmul(3.0, v0), c.length[a]);
        const LccLC lc = lcc_get_lc(c, a, lane);
        if (!lc.in_candidate && lcs_dsub(reverse_s, lc.force_len) <= lcs_dmul(lc_length, (double)lc.count)) {
            c.scal[LCC_LC_REQUIRED] = 1; // informational: this scenario exercises lane changes
            tl = lc.side == 0 ? c.lane_left[lane] : c.lane_right[lane];
            if (tl < 0) {
                c.scal[LCC_LC_ERROR] = 1; // reference: assert target_lane is not None
            } else {
                ts = lcc_project_from_lane(c, tl, lane, cur_s);
                c.is_lc[a] = 1;
                c.lc_total[a] = lc_length;
                c.lc_completed[a] = 0.0;
                c.lc_target[a] = tl;
                c.lc_target_s[a] = ts;
            }
        }
    }
    c.acc[a] = acc;
    c.rd_ahead[a] = -1;
    if (tl >= 0) {
        double ks;
        int other;
        if (lcc_next_agent(c, tl, ts, &ks, &other)) { // NB the reference uses the leader's absolute s (idm_policy.py:77,157)
            c.rd_ahead[a] = other;
            c.rd_dist[a] = lcs_dsub(ks, lcc_div(lcs_dadd(c.length[a], c.length[other]), 2.0));
            c.rd_lane[a] = tl;
        }
    }
    if (c.rd_ahead[a] < 0) {
        c.done_round[a] = 0;
        lcc_finalize(c, i, a, acc);
    } else {
        c.done_round[a] = 0x7fffffff;
        lcs_atomic_add(&c.scal[LCC_N_PENDING], 1);
    }
}
#ifndef LCC_UPDATE_MINB
#define LCC_UPDATE_MINB 0
#endif
#if LCC_UPDATE_MINB > 0
LCC_KERNEL_LB(lcc_update, LCC_UPDATE_MINB)
#else
LCC_KERNEL(lcc_update)
#endif
// reader of slot i: ahead_a.runtime.v is this tick's value if that agent precedes it in the list, else last tick's
LCS_D bool lcc_reader_step(const LccP& c, int i, int a, int round, bool serial) {
    if (c.done_round[a] != 0x7fffffff) return false;
    const int x = c.rd_ahead[a];
    const int sx = c.slot_of[x];
    double vx;
    if (sx >= 0 && sx < i) {
        if (!serial && c.done_round[x] >= round) return false; // not finished in an EARLIER launch: wait
        vx = c.v2[x];
    } else {
        vx = c.v[x];
    }
    const double acc2 = lcc_car_follow(c.idm + 5 * (size_t)a, c.v[a], c.lane_max_speed[c.rd_lane[a]], vx, c.rd_dist[a]);
    lcc_finalize(c, i, a, lcs_pymin(c.acc[a], acc2));
    c.done_round[a] = round;
    return true;
}
LCS_D void lcc_update_round(const LccP& c, int t) {
    if (c.scal[LCC_N_PENDING] == 0) return;
    int a, i;
    if (!lcc_member(c, t, &a, &i)) return;
    if (lcc_reader_step(c, i, a, c.mode /* round number travels in mode */, false)) lcs_atomic_add(&c.scal[LCC_N_PENDING], -1);
}
LCC_KERNEL(lcc_update_round)
// whatever dependency chain is longer than the rounds: one thread, list order (normally no work at all)
LCS_D void lcc_update_serial(const LccP& c, int unused) {
    if (c.scal[LCC_N_PENDING] == 0) return;
    const int n = c.scal[LCC_N_ACTIVE];
    for (int i = 0; i < n; i++)
        if (lcc_reader_step(c, i, c.active[i], 0x7ffffffe, true)) c.scal[LCC_N_PENDING] -= 1; // list order: predecessors are finished
}
LCC_KERNEL(lcc_update_serial)
LCS_D void lcc_commit_v(const LccP& c, int t) {
    int a, i;
    if (!lcc_member(c, t, &a, &i)) return;
    c.v[a] = c.v2[a];
}
LCC_KERNEL(lcc_commit_v)

// ------------------------------------------------------------------ arrivals / departures (manager.py:50-96)
// waiting offset j: the popped entries are a prefix of the (sorted) waiting list; the last popped one reports m
LCS_D void lcc_pop_count(const LccP& c, int j) {
    const int wp = c.scal[LCC_WAIT_PTR], tick = c.scal[LCC_TICK];
    if (j == 0) {
        const bool any = wp < c.N && c.depart_tick[c.wait_order[wp]] <= tick;
        if (!any) c.scal[LCC_M_POP] = 0;
    }
    const int p = wp + j;
    if (p >= c.N) return;
    if (c.depart_tick[c.wait_order[p]] <= tick && (p + 1 == c.N || c.depart_tick[c.wait_order[p + 1]] > tick)) c.scal[LCC_M_POP] = j + 1;
}
LCC_KERNEL(lcc_pop_count)
// one launch for two independent per-index jobs of a tick: the speed commit of list slot i and the pop test of waiting
// offset i
LCS_D void lcc_commit_pop(const LccP& c, int i) {
    lcc_commit_v(c, i);
    lcc_pop_count(c, i);
}
LCC_KERNEL(lcc_commit_pop)
// thread i: list slot i and, independently, popped waiting entry i
LCS_D void lcc_place(const LccP& c, int i) {
    const int n = c.scal[LCC_N_ACTIVE], n_arr = c.scal[LCC_N_ARR], m = c.scal[LCC_M_POP], wp = c.scal[LCC_WAIT_PTR];
    if (i < n) {
        const int r = c.rank[i]; // finished slots before i
        if (!c.fin[i]) {
            const int pos = i - (r > m ? r - m : 0);
            c.active2[pos] = c.active[i];
            c.slot_of[c.active[i]] = pos;
        } else {
            c.slot_of[c.active[i]] = -1;
            if (r < m) { // departure r takes the slot of the r-th finished agent
                c.active2[i] = c.wait_order[wp + r];
                c.slot_of[c.wait_order[wp + r]] = i;
            }
        }
    }
    if (i < m) {
        const int a = c.wait_order[wp + i];
        c.status[a] = LCS_ACTIVE;
        if (i >= n_arr) {
            c.active2[n + (i - n_arr)] = a;
            c.slot_of[a] = n + (i - n_arr);
        }
        lcc_add_to_lane(c, c.lane[a], c.s[a], a, ((c.mode == 1 ? 0 : 2) << 28) | i); // add_to_lane at the home position
    }
}
LCC_KERNEL(lcc_place)
LCS_D void lcc_place_end(const LccP& c, int i) {
    const int n = c.scal[LCC_N_ACTIVE], n_arr = c.scal[LCC_N_ARR], m = c.scal[LCC_M_POP];
    c.scal[LCC_N_OLD] = n;
    c.scal[LCC_N_PENDING] = 0;
    c.scal[LCC_N_ACTIVE] = n - n_arr + m;
    c.scal[LCC_WAIT_PTR] += m;
    if (c.mode == 0) {
        c.stats[0] += n;
        c.stats[1] += n - n_arr + m;
        c.stats[4] += n_arr;
        c.stats[6] += 1;
    }
    c.stats[5] += m;
}
LCC_KERNEL(lcc_place_end)
#ifndef LCSIM_HOSTEMU
// lcc_place for every index, then lcc_place_end by the block that finishes last (its threads' reads of the old counters
// are over by then: every block passed its barrier before it took a ticket)
__global__ void lcc_place_all_k(const LccP c, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) lcc_place(c, i);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const int done = atomicAdd(&c.scal[LCC_SYNC], 1);
        if (done == (int)gridDim.x - 1) {
            c.scal[LCC_SYNC] = 0;
            __threadfence();
            lcc_place_end(c, 0);
        }
    }
}
#endif
// end of a tick on the main branch: the new list becomes the list, the clock advances
LCS_D void lcc_finish(const LccP& c, int i) {
    c.active[i] = c.active2[i];
    if (i == 0 && c.mode == 0) c.scal[LCC_TICK] += 1;
}
LCC_KERNEL(lcc_finish)

// ------------------------------------------------------------------ lane buckets = Lane.prepare (lane.py:171-175)
LCS_D void lcc_bucket_count(const LccP& c, int e) {
    if (e < c.scal[LCC_N_ENT]) lcs_atomic_add(&c.lane_start[c.ent_lane[e]], 1);
}
LCC_KERNEL(lcc_bucket_count)
LCS_D void lcc_bucket_scatter(const LccP& c, int e) {
    if (e >= c.scal[LCC_N_ENT]) return;
    const int pos = lcs_atomic_add(&c.lane_cur[c.ent_lane[e]], 1);
    c.bkt_s[pos] = c.ent_s[e];
    c.bkt_agent[pos] = c.ent_agent[e];
    c.bkt_order[pos] = c.ent_order[e];
}
LCC_KERNEL(lcc_bucket_scatter)
LCS_D void lcc_bucket_end(const LccP& c, int i) { c.scal[LCC_N_ENT] = 0; }
LCC_KERNEL(lcc_bucket_end)

// SortedDict semantics: one entry per key; a later add_agent with the same s replaced the earlier one
// (lane.py:207-208), so among equal keys the entry with the highest insertion order is the one that exists.
// get_next_agent_by_s (lane.py:210-217): smallest key > s
LCS_D bool lcc_next_agent(const LccP& c, int lane, double s, double* ks, int* agent) {
    double best = lcs_inf_d();
    int bo = -1, ba = -1;
    for (int e = c.lane_start[lane]; e < c.lane_start[lane + 1]; e++) {
        const double k = c.bkt_s[e];
        if (k > s && (k < best || (k == best && c.bkt_order[e] > bo))) { best = k; bo = c.bkt_order[e]; ba = c.bkt_agent[e]; }
    }
    *ks = best;
    *agent = ba;
    return ba >= 0;
}
// get_first_agent (lane.py:228-234): smallest key
LCS_D bool lcc_first_agent(const LccP& c, int lane, double* ks, int* agent) {
    double best = lcs_inf_d();
    int bo = -1, ba = -1;
    for (int e = c.lane_start[lane]; e < c.lane_start[lane + 1]; e++) {
        const double k = c.bkt_s[e];
        if (ba < 0 || k < best || (k == best && c.bkt_order[e] > bo)) { best = k; bo = c.bkt_order[e]; ba = c.bkt_agent[e]; }
    }
    *ks = best;
    *agent = ba;
    return ba >= 0;
}

// ------------------------------------------------------------------ prepare_obs, LaneIDM branch (agent.py:198-243)
LCS_D void lcc_observe(const LccP& c, int i) {
    if (i == 0 && c.mode == 0) c.scal[LCC_N_ENT] = 0; // (lcc_bucket_end) the pending lane entries were all bucketed
    int a;
    if (!lcc_member2(c, i, &a)) return;
    const int lane = c.lane[a];
    const double s = c.s[a], len_a = c.length[a];
    double ks;
    int other;
    bool valid = false;
    double dis = lcs_inf_d(), lv = 0.0;
    if (lcc_next_agent(c, lane, s, &ks, &other)) {
        // NB reference: ahead[0] is the leader's absolute s on the lane, not the gap (agent.py:227-230)
        dis = lcs_dsub(ks, lcc_div(lcs_dadd(len_a, c.length[other]), 2.0));
        lv = c.v[other];
        valid = true;
    } else {
        const double view = lcs_pymax(50.0, lcs_dmul(c.v[a], 12.0));
        double look = lcs_dsub(c.lane_length[lane], s);
        int cur = lane, junc_index = 0;
        while (look < view) {
            if (c.lane_junction[cur] >= 0) {
                cur = c.lane_succ0[cur];
                junc_index++;
            } else {
                cur = lcc_junc_lane_by_pre_lane(c, a, cur, junc_index);
            }
            if (cur < 0) break;
            if (lcc_first_agent(c, cur, &ks, &other)) {
                dis = lcs_dsub(lcs_dadd(look, ks), lcc_div(lcs_dadd(len_a, c.length[other]), 2.0));
                lv = c.v[other];
                valid = true;
                break;
            }
            look = lcs_dadd(look, c.lane_length[cur]);
        }
    }
    c.lead_valid[a] = valid ? 1 : 0;
    c.lead_dis[a] = dis;
    c.lead_v[a] = lv;
}
LCC_KERNEL(lcc_observe)

// ------------------------------------------------------------------ metrics
// uniform spatial hash of the active agents; cell >= the largest bounding-circle diameter, so every overlapping
// pair lies in neighbouring cells (check_collision_all, manager.py:313-327, is all pairs in the reference)
LCS_D int lcc_cell_of(const LccP& c, int a) {
    int cx = (int)floor(lcs_dmul(lcs_dsub(c.xy[2 * a], c.cell_x0), c.cell_inv));
    int cy = (int)floor(lcs_dmul(lcs_dsub(c.xy[2 * a + 1], c.cell_y0), c.cell_inv));
    cx = cx < 0 ? 0 : (cx >= c.cell_w ? c.cell_w - 1 : cx);
    cy = cy < 0 ? 0 : (cy >= c.cell_h ? c.cell_h - 1 : cy);
    return cy * c.cell_w + cx;
}
LCS_D void lcc_cell_count(const LccP& c, int i) {
    int a;
    if (lcc_member2(c, i, &a)) lcs_atomic_add(&c.cell_start[lcc_cell_of(c, a)], 1);
}
LCC_KERNEL(lcc_cell_count)
// slot -> its cell, as a 64-byte record {x, y, heading, length, width, bounding radius, agent}: the collision kernel
// then reads its candidates from three contiguous runs (the three cells of a grid row are neighbours in the table)
LCS_D void lcc_cell_scatter(const LccP& c, int i) {
    int a;
    if (!lcc_member2(c, i, &a)) return;
    const int e = lcs_atomic_add(&c.cell_cur[lcc_cell_of(c, a)], 1);
    c.cell_agent[e] = a;
    double* r = c.cell_rec + 8 * (size_t)e;
    const double L = c.length[a], W = c.width[a];
    r[0] = c.xy[2 * a]; r[1] = c.xy[2 * a + 1]; r[2] = c.h[a]; r[3] = L; r[4] = W;
    r[5] = 0.5 * sqrt(L * L + W * W);
    r[6] = (double)a;
}
LCC_KERNEL(lcc_cell_scatter)
// thread e: the e-th entry of the cell table (cell order: its own 64-byte record and its neighbours' come from the
// same few lines); the trigonometry runs only for pairs whose bounding circles touch
LCS_D void lcc_collide(const LccP& c, int e0) {
    if (e0 >= c.scal[LCC_N_ACTIVE]) return;
    const double* me = c.cell_rec + 8 * (size_t)e0;
    const double mx = me[0], my = me[1], mh = me[2], mL = me[3], mW = me[4], ra = me[5];
    const int a = (int)me[6];
    double ea[8], eb[8];
    bool have_ea = false;
    int cx = (int)floor(lcs_dmul(lcs_dsub(mx, c.cell_x0), c.cell_inv)), cy = (int)floor(lcs_dmul(lcs_dsub(my, c.cell_y0), c.cell_inv));
    cx = cx < 0 ? 0 : (cx >= c.cell_w ? c.cell_w - 1 : cx);
    cy = cy < 0 ? 0 : (cy >= c.cell_h ? c.cell_h - 1 : cy);
    const int x0 = cx > 0 ? cx - 1 : 0, x1 = cx < c.cell_w - 1 ? cx + 1 : c.cell_w - 1;
    int cnt = 0;
    for (int y = cy - 1; y <= cy + 1; y++) {
        if (y < 0 || y >= c.cell_h) continue;
        const int b0 = c.cell_start[y * c.cell_w + x0], b1 = c.cell_start[y * c.cell_w + x1 + 1];
        for (int e = b0; e < b1; e++) {
            if (e == e0) continue;
            const double* r = c.cell_rec + 8 * (size_t)e;
            const double dx = r[0] - mx, dy = r[1] - my;
            const double rr = (ra + r[5]) * 1.000001 + 1e-6; // bounding circles apart: the SAT cannot overlap
            if (dx * dx + dy * dy > rr * rr) continue;
            double sh, ch;
            if (!have_ea) {
                lcs_sincos(mh, &sh, &ch);
                ea[0] = mx; ea[1] = my; ea[2] = 0.0; ea[3] = ch; ea[4] = sh; ea[5] = mL; ea[6] = mW; ea[7] = mh;
                have_ea = true;
            }
            lcs_sincos(r[2], &sh, &ch);
            eb[0] = r[0]; eb[1] = r[1]; eb[2] = 0.0; eb[3] = ch; eb[4] = sh; eb[5] = r[3]; eb[6] = r[4]; eb[7] = r[2];
            if (lcs_collide(ea, eb)) cnt++;
        }
    }
    c.coll_count[a] = cnt;
    if (cnt) atomicAdd_i64(&c.stats[2], cnt);
}
#ifndef LCC_COLLIDE_MINB
#define LCC_COLLIDE_MINB 0
#endif
#if LCC_COLLIDE_MINB > 0
LCC_KERNEL_LB(lcc_collide, LCC_COLLIDE_MINB)
#else
LCC_KERNEL(lcc_collide)
#endif
// Agent.is_offroad (agent.py:482-494) -> Road.check_offroad (road.py:342-362, raw nodes) /
// Junction.check_offroad (junction.py:365-385, resampled): first-index argmin, side by the 2-D cross product
LCS_D void lcc_offroad(const LccP& c, int i) {
    int a;
    if (!lcc_member2(c, i, &a)) return;
    const int es = c.lane_eset[c.lane[a]];
    const int b = c.eset_start[es], e = c.eset_start[es + 1];
    const double px = c.xy[2 * a], py = c.xy[2 * a + 1];
    uint8_t off = 0;
    if (e > b) {
        const int k = b + lcs_argmin_exact(c.edge_xy + 2 * (size_t)b, e - b, px, py);
        const double ax = lcs_dsub(c.edge_xy[2 * k], px), ay = lcs_dsub(c.edge_xy[2 * k + 1], py);
        off = lcs_dsub(lcs_dmul(ax, c.edge_dir[2 * k + 1]), lcs_dmul(ay, c.edge_dir[2 * k])) < 0 ? 1 : 0;
    }
    c.offroad[a] = off;
    if (off) atomicAdd_i64(&c.stats[3], 1);
}
LCC_KERNEL(lcc_offroad)
LCS_D void lcc_tick_end(const LccP& c, int i) { c.scal[LCC_TICK] += 1; }
LCC_KERNEL(lcc_tick_end)
