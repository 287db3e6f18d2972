/*
 * lcsim_oracle.c — TEST INFRASTRUCTURE, not product code.
 *
 * A plain-C, float64, brute-force restatement of the reference's per-tick
 * agent-stepping path (tsinghua-fib-lab/LCSim, pure Python/NumPy).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library; the product (lcsim_b200/) never does.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks this file against
 * golden vectors produced by running the real reference engine (imported from
 * /root/reference with the shims of oracle/ref_harness.py) on the two bundled
 * WOMD scenarios and on synthetic scenarios, in both control modes
 * (tests/golden/*.npz, generator: oracle/gen_golden.py).
 *
 * Every function cites the reference file:line it follows.  Floating-point
 * operation order follows the reference expression by expression.  Where the
 * reference goes through BLAS (np.dot / np.matmul / 1-D np.linalg.norm) the
 * two-term dot product is evaluated the way OpenBLAS does on FMA hardware,
 *     acc = a0*b0; acc = fma(a1, b1, acc)
 * (probed against numpy 2.3.5 in the build container: 100% agreement, versus
 * ~75% for the unfused form).  Everything else is unfused; build with
 * -ffp-contract=off (oracle/Makefile) so the compiler adds no contractions.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_WAITING 0
#define ORC_ACTIVE 1
#define ORC_FINISHED 2
#define ORC_IDLE 3

#define ORC_ACT_NONE 0
#define ORC_ACT_BICYCLE 1 /* reference: policy/type.py:79-84 */
#define ORC_ACT_DELTA 2
#define ORC_ACT_STATE 3

#define ORC_H 11 /* reference: agent.py:89 NUM_HISTORICAL_STEPS */
#define ORC_ROUTE 10 /* reference: agent.py:90 ROUTE_LENGTH */

typedef struct {
    int32_t S, A, T, E;
    double start, interval;
    int32_t total_steps;
    int32_t reactive, no_static, allow_new_obj;
    const int32_t *n_agents, *ads_index, *agent_id, *agent_type;
    const double *length, *width;
    const uint8_t *dynamic;
    const int32_t *traj_len, *depart_tick, *wait_order;
    const double *home_xy, *traj_xy, *traj_vel, *traj_h;
    const int32_t *n_edge;
    const double *edge_xy, *edge_dir;
} orc_static;

typedef struct {
    orc_static st;
    /* mutable copies of the reference trajectories (set_ref_trajectory overwrites them) */
    int32_t *traj_len;
    double *traj_xy, *traj_vel, *traj_h;
    uint8_t *traj_f32; /* 1 after a re-plan: reference holds float32 arrays then (engine.py:290-297) */
    /* Agent.obs.ref_trajectory keeps pointing at the OLD Trajectory object until the next
     * prepare_obs (agent.py:247,250 vs :462-466): the first policy action after a re-plan is
     * taken from the old trajectory's next waypoint. */
    uint8_t *stale_valid, *stale_f32;
    double *stale_wp; /* [S][A][5] x,y,vx,vy,h */
    /* per scenario */
    int32_t *tick, *n_active, *wait_ptr, *active; /* active: [S][A] slot -> agent */
    /* per agent */
    int32_t *status, *cursor;
    double *x, *y, *h, *v;
    uint8_t *lead_valid;
    double *lead_dis, *lead_v;
    double *ring; /* [S][A][H][5] x,y,vx,vy,h, shift-register order */
    int32_t *ring_valid; /* number of valid trailing entries, <= H */
    /* metrics (filled by orc_metrics) */
    int32_t *coll_count, *nearest_edge;
    uint8_t *offroad;
    int64_t *agent_steps; /* [S] running count of agent updates */
} orc_engine;

/* ---------------------------------------------------------------- helpers */

/* reference: utils/polygon.py:191-192  (angle + pi) % (2*pi) - pi, NumPy/Python floored modulo */
static double orc_wrap_angle(double a) {
    const double two_pi = 2 * M_PI;
    double t = a + M_PI;
    double r = fmod(t, two_pi);
    if (r != 0.0) {
        if (r < 0.0) r += two_pi;
    } else {
        r = copysign(0.0, two_pi);
    }
    return r - M_PI;
}

/* two-term dot product as BLAS evaluates it (see header) */
static inline double orc_dot2(double a0, double b0, double a1, double b1) { return fma(a1, b1, a0 * b0); }

/* np.linalg.norm over a trailing axis of length 2: sqrt(dx*dx + dy*dy), unfused */
static inline double orc_norm2_axis(double dx, double dy) { return sqrt(dx * dx + dy * dy); }

/* np.linalg.norm of a 1-D 2-vector: sqrt(x.dot(x)) through BLAS ddot */
static inline double orc_norm2_vec(double dx, double dy) { return sqrt(orc_dot2(dx, dx, dy, dy)); }

/* numpy's pairwise summation (np.sum / np.add.reduce over a contiguous 1-D double array) */
static double orc_np_sum(const double* a, int64_t n) {
    if (n < 8) {
        double r = 0.0;
        for (int64_t i = 0; i < n; i++) r += a[i];
        return r;
    } else if (n <= 128) {
        double r[8];
        int64_t i;
        for (i = 0; i < 8; i++) r[i] = a[i];
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; j++) r[j] += a[i + j];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; i++) res += a[i];
        return res;
    } else {
        int64_t n2 = n / 2;
        n2 -= n2 % 8;
        return orc_np_sum(a, n2) + orc_np_sum(a + n2, n - n2);
    }
}
/* same algorithm on float32 data (a float32 reference trajectory after a re-plan) */
static float orc_np_sumf(const float* a, int64_t n) {
    if (n < 8) {
        float r = 0.0f;
        for (int64_t i = 0; i < n; i++) r += a[i];
        return r;
    } else if (n <= 128) {
        float r[8];
        int64_t i;
        for (i = 0; i < 8; i++) r[i] = a[i];
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; j++) r[j] += a[i + j];
        float res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; i++) res += a[i];
        return res;
    } else {
        int64_t n2 = n / 2;
        n2 -= n2 % 8;
        return orc_np_sumf(a, n2) + orc_np_sumf(a + n2, n - n2);
    }
}
double orc_test_np_sum(const double* a, int64_t n) { return orc_np_sum(a, n); }
double orc_test_wrap_angle(double a) { return orc_wrap_angle(a); }

#define IDX_SA(e, s, a) ((int64_t)(s) * (e)->st.A + (a))
#define TRAJ(e, s, a) (((int64_t)(s) * (e)->st.A + (a)) * (e)->st.T)

/* reference: utils/polygon.py:271-293 frenet_projection: first-index argmin of the
 * point distances (sqrt taken BEFORE the compare), d = norm(point - closest), s = sum of
 * the segment lengths before the closest index. */
static int32_t orc_argmin_pts(const double* pts, int32_t n, double px, double py) {
    int32_t best = 0;
    double bd = INFINITY;
    for (int32_t k = 0; k < n; k++) {
        double d = orc_norm2_axis(pts[2 * k] - px, pts[2 * k + 1] - py);
        if (d < bd) { bd = d; best = k; }
    }
    return best;
}

/* f32: the trajectory is a float32 array (re-plan): `reference_line[1:] - reference_line[:-1]`, its
 * norm and np.sum then run in float32; the point differences stay float64 (pos is float64). */
static double orc_seg_sum(const double* pts, int32_t n_seg_total, int32_t k, int f32) {
    if (k > n_seg_total) k = n_seg_total;
    if (k <= 0) return 0.0;
    if (f32) {
        float* sg = (float*)malloc(sizeof(float) * k);
        for (int32_t i = 0; i < k; i++) {
            float dx = (float)pts[2 * (i + 1)] - (float)pts[2 * i], dy = (float)pts[2 * (i + 1) + 1] - (float)pts[2 * i + 1];
            sg[i] = sqrtf(dx * dx + dy * dy);
        }
        float r = orc_np_sumf(sg, k);
        free(sg);
        return (double)r;
    }
    double* sg = (double*)malloc(sizeof(double) * k);
    for (int32_t i = 0; i < k; i++)
        sg[i] = orc_norm2_axis(pts[2 * (i + 1)] - pts[2 * i], pts[2 * (i + 1) + 1] - pts[2 * i + 1]);
    double r = orc_np_sum(sg, k);
    free(sg);
    return r;
}

static void orc_frenet(const double* pts, int32_t n, double px, double py, double* d, double* s, int32_t* idx, int f32) {
    int32_t k = orc_argmin_pts(pts, n, px, py);
    *idx = k;
    *d = orc_norm2_vec(px - pts[2 * k], py - pts[2 * k + 1]);
    *s = orc_seg_sum(pts, n - 1, k, f32);
}

/* reference: utils/polygon.py:296-305 get_length */
static double orc_polyline_length(const double* pts, int32_t n, int f32) { return orc_seg_sum(pts, n - 1, n - 1, f32); }

/* reference: agent/agent.py:67-84 HistoryState.update (shift left, append) */
static void orc_ring_push(orc_engine* e, int s, int a, double x, double y, double h, double v, int f32) {
    double* r = e->ring + IDX_SA(e, s, a) * ORC_H * 5;
    memmove(r, r + 5, sizeof(double) * 5 * (ORC_H - 1));
    double* last = r + 5 * (ORC_H - 1);
    /* reference: agent.py:419-426 — velocity re-synthesised from (v, heading) */
    last[0] = x;
    last[1] = y;
    if (f32) { /* runtime.v / heading are np.float32 after a STATE action read from a float32 re-plan */
        last[2] = (double)((float)v * cosf((float)h));
        last[3] = (double)((float)v * sinf((float)h));
    } else {
        last[2] = v * cos(h);
        last[3] = v * sin(h);
    }
    last[4] = h;
    int32_t* nv = e->ring_valid + IDX_SA(e, s, a);
    if (*nv < ORC_H) (*nv)++;
}

/* ------------------------------------------------ departure / arrival (A12) */

/* reference: agent/manager.py:50-96 check_departure_and_arrival, single-trip schedules
 * (agent/schedule.py:22-40: next_trip -> False -> IDLE).  The float64 clock compare
 * `departure_time <= current_time` is pre-evaluated on the host as depart_tick
 * (lcsim_b200/packer.py depart_tick, reference engine.py:157). */
static void orc_departure_arrival(orc_engine* e, int s, int tick, int allow_new) {
    const orc_static* st = &e->st;
    int A = st->A;
    int32_t* act = e->active + (int64_t)s * A;
    int n = e->n_active[s];
    int arr[4096];
    int n_arr = 0;
    for (int i = 0; i < n; i++) {
        int a = act[i];
        if (e->status[IDX_SA(e, s, a)] != ORC_FINISHED) continue;
        e->status[IDX_SA(e, s, a)] = ORC_IDLE;
        arr[n_arr++] = i;
    }
    int moved = 0;
    if (allow_new) {
        const int32_t* wo = st->wait_order + (int64_t)s * A;
        while (e->wait_ptr[s] < st->n_agents[s] &&
               st->depart_tick[IDX_SA(e, s, wo[e->wait_ptr[s]])] <= tick) {
            int a = wo[e->wait_ptr[s]++];
            if (st->no_static && !st->dynamic[IDX_SA(e, s, a)] && a != st->ads_index[s]) continue;
            e->status[IDX_SA(e, s, a)] = ORC_ACTIVE;
            if (moved < n_arr)
                act[arr[moved]] = a;
            else
                act[n++] = a;
            moved++;
        }
    }
    /* drop the finished slots nobody took over, preserving order */
    if (moved < n_arr) {
        int w = 0, q = moved;
        for (int i = 0; i < n; i++) {
            if (q < n_arr && arr[q] == i) { q++; continue; }
            act[w++] = act[i];
        }
        n = w;
    }
    e->n_active[s] = n;
}

/* ------------------------------------------------------ lead search (A7) */

/* reference: agent/agent.py:249-299 TrajIDM branch of Agent.prepare_obs */
static void orc_lead_search(orc_engine* e, int s, int a) {
    const orc_static* st = &e->st;
    int64_t ia = IDX_SA(e, s, a);
    double px = e->x[ia], py = e->y[ia], heading = e->h[ia];
    double L = st->length[ia], W = st->width[ia];
    double ch = cos(heading), sh = sin(heading);
    double best = INFINITY, best_v = 0.0;
    const int32_t* act = e->active + (int64_t)s * st->A;
    for (int i = 0; i < e->n_active[s]; i++) {
        int o = act[i];
        if (o == a) continue;
        int64_t io = IDX_SA(e, s, o);
        double dx = e->x[io] - px, dy = e->y[io] - py;
        /* np.dot(xy, rot_mat), rot_mat = [[c,-s],[s,c]] */
        double rx = orc_dot2(dx, ch, dy, sh);
        double ry = orc_dot2(dx, -sh, dy, ch);
        double la = st->length[io], wa = st->width[io];
        double rel = orc_wrap_angle(e->h[io] - heading);
        double c = cos(rel), sn = sin(rel);
        double yc[4] = {ry + la / 2.0 * sn - wa / 2.0 * c, ry + la / 2.0 * sn + wa / 2.0 * c,
                        ry - la / 2.0 * sn + wa / 2.0 * c, ry - la / 2.0 * sn - wa / 2.0 * c};
        double ymax = yc[0], ymin = yc[0];
        for (int k = 1; k < 4; k++) {
            if (yc[k] > ymax) ymax = yc[k];
            if (yc[k] < ymin) ymin = yc[k];
        }
        if (ymax < -W / 2.0 || ymin > W / 2.0) continue;
        double xc[4] = {rx + la / 2.0 * c + wa / 2.0 * sn, rx + la / 2.0 * c - wa / 2.0 * sn,
                        rx - la / 2.0 * c - wa / 2.0 * sn, rx - la / 2.0 * c + wa / 2.0 * sn};
        double xmin = xc[0];
        for (int k = 1; k < 4; k++)
            if (xc[k] < xmin) xmin = xc[k];
        if (xmin < L / 2.0) continue;
        double dis = xmin - L / 2.0;
        if (dis < best) {
            best = dis;
            best_v = e->v[io] * cos(rel);
        }
    }
    e->lead_valid[ia] = best != INFINITY;
    e->lead_dis[ia] = best;
    e->lead_v[ia] = best_v;
}

/* reference: agent/manager.py:98-103 + agent.py:193-302 */
static void orc_prepare_obs(orc_engine* e, int s) {
    for (int a = 0; a < e->st.A; a++) e->stale_valid[IDX_SA(e, s, a)] = 0; /* obs snapshot refreshed */
    if (!e->st.reactive) return; /* StateExpert branch is a plain snapshot (agent.py:244-248) */
    const int32_t* act = e->active + (int64_t)s * e->st.A;
    for (int i = 0; i < e->n_active[s]; i++) orc_lead_search(e, s, act[i]);
}

/* ---------------------------------------------------------- policies (A4/A6) */

/* reference: policy/type.py:32-40 next_waypoint */
static int orc_next_wp(orc_engine* e, int s, int a) {
    int64_t ia = IDX_SA(e, s, a);
    return e->traj_len[ia] == 1 ? e->cursor[ia] : e->cursor[ia] + 1;
}

/* python: max(a, b) returns a unless b > a; min(a, b) returns a unless b < a */
static inline double py_max(double a, double b) { return b > a ? b : a; }
static inline double py_min(double a, double b) { return b < a ? b : a; }
/* np.clip(x, lo, hi) = minimum(maximum(x, lo), hi) (NaN propagates) */
static inline double np_clip(double x, double lo, double hi) {
    if (isnan(x)) return x;
    double m = x < lo ? lo : x;
    return m > hi ? hi : m;
}

/* reference: policy/idm_policy.py:226-245 TrajIDM._get_acc (class constants, Q19).
 * `**` on floats is C pow(); glibc's pow(x, 2.0) equals x*x for every double probed. */
static double orc_idm_acc(double self_v, double target_v, double ahead_v, double distance) {
    const double min_gap = 10.0, headway = 2.0, max_a = 2.0, usual_brake = -2.0, max_brake = -2.0;
    double acc;
    if (distance - min_gap < 0) {
        acc = -INFINITY;
    } else {
        double s_star = min_gap + py_max(self_v * headway + self_v * (self_v - ahead_v) / (2 * sqrt(-usual_brake * max_a)), 0.0);
        acc = max_a * (1 - pow(self_v / target_v, 4) - pow(s_star / distance, 2));
    }
    return py_max(max_brake, py_min(max_a, acc));
}

/* obs.ref_trajectory.next_waypoint() as the policy sees it (stale after a re-plan, see orc_engine) */
static int orc_policy_waypoint(orc_engine* e, int s, int a, double* wp /* x,y,vx,vy,h */) {
    int64_t ia = IDX_SA(e, s, a);
    if (e->stale_valid[ia]) {
        memcpy(wp, e->stale_wp + 5 * ia, sizeof(double) * 5);
        return e->stale_f32[ia];
    }
    int w = orc_next_wp(e, s, a);
    int64_t t0 = TRAJ(e, s, a);
    wp[0] = e->traj_xy[2 * (t0 + w)];
    wp[1] = e->traj_xy[2 * (t0 + w) + 1];
    wp[2] = e->traj_vel[2 * (t0 + w)];
    wp[3] = e->traj_vel[2 * (t0 + w) + 1];
    wp[4] = e->traj_h[t0 + w];
    return e->traj_f32[ia];
}

/* reference: policy/idm_policy.py:196-224 TrajIDM.get_action */
static void orc_trajidm_action(orc_engine* e, int s, int a, double dt, double* acc_out, double* steer_out) {
    int64_t ia = IDX_SA(e, s, a);
    double wp[5];
    int wp_f32 = orc_policy_waypoint(e, s, a, wp);
    double vx = wp[2], vy = wp[3];
    double wh = wp[4];
    double v = orc_norm2_vec(vx, vy); /* np.linalg.norm(vel) on a 1-D vector */
    if (wp_f32) { /* float32 vel array after a re-plan: the norm is evaluated in float32 */
        float fvx = (float)vx, fvy = (float)vy;
        v = (double)sqrtf(fmaf(fvy, fvy, fvx * fvx));
    }
    double cur_v = e->v[ia], cur_h = e->h[ia];
    double acc = (v - cur_v) / dt;
    acc = np_clip(acc, -3.0, 3.0);
    double dh = orc_wrap_angle(wh - cur_h);
    double steer;
    if (cur_v < 0.6 || v < 0.6)
        steer = 0.0;
    else
        steer = dh / (cur_v * dt + 0.5 * acc * (dt * dt)); /* dt**2 == dt*dt */
    steer = np_clip(steer, -0.3, 0.3);
    if (e->lead_valid[ia]) {
        double idm = orc_idm_acc(cur_v, 40.0, e->lead_v[ia], e->lead_dis[ia]);
        acc = py_max(py_min(acc, idm), -3.0);
    }
    *acc_out = acc;
    *steer_out = steer;
}

/* ------------------------------------------------------------- update (A4/A5) */

static void orc_cursor_update(orc_engine* e, int s, int a, double px, double py) {
    int64_t ia = IDX_SA(e, s, a);
    int n = e->traj_len[ia];
    e->cursor[ia] = orc_argmin_pts(e->traj_xy + 2 * TRAJ(e, s, a), n, px, py);
    if (e->cursor[ia] >= n - 1) e->status[ia] = ORC_FINISHED; /* type.py:50-51 reach_end */
}

/* reference: agent/agent.py:310-426 update_runtime_by_action (STATE / BICYCLE / DELTA branches) */
static void orc_apply_action(orc_engine* e, int s, int a, int type, const double* data, double dt, int f32) {
    int64_t ia = IDX_SA(e, s, a);
    int ring_f32 = 0;
    if (type == ORC_ACT_STATE) {
        /* agent.py:377-388 — cursor from the NEW position */
        double x = data[0], y = data[1], vx = data[2], vy = data[3];
        double v;
        if (f32) { /* float32 scalars after a re-plan: np.sqrt(vx**2+vy**2) stays float32 */
            float fvx = (float)vx, fvy = (float)vy;
            v = (double)sqrtf(fvx * fvx + fvy * fvy);
            ring_f32 = 1;
        } else {
            v = sqrt(vx * vx + vy * vy);
        }
        e->x[ia] = x;
        e->y[ia] = y;
        e->v[ia] = v;
        e->h[ia] = data[4];
        orc_cursor_update(e, s, a, x, y);
    } else if (type == ORC_ACT_BICYCLE) {
        /* agent.py:389-405 — cursor from the OLD position; no speed clamp (Q6) */
        double acc = data[0], steer = data[1];
        double x = e->x[ia], y = e->y[ia], v = e->v[ia], heading = e->h[ia];
        double delta = v * dt + 0.5 * acc * (dt * dt);
        double dheading = steer * (v * dt + 0.5 * acc * (dt * dt));
        double nh = orc_wrap_angle(heading + dheading);
        e->x[ia] = x + delta * cos(nh);
        e->y[ia] = y + delta * sin(nh);
        e->v[ia] = v + acc * dt;
        e->h[ia] = nh;
        orc_cursor_update(e, s, a, x, y);
    } else if (type == ORC_ACT_DELTA) {
        /* agent.py:406-416 */
        double x = e->x[ia], y = e->y[ia];
        e->x[ia] = x + data[0];
        e->y[ia] = y + data[1];
        e->h[ia] = orc_wrap_angle(e->h[ia] + data[2]);
        orc_cursor_update(e, s, a, x, y);
    }
    orc_ring_push(e, s, a, e->x[ia], e->y[ia], e->h[ia], e->v[ia], ring_f32);
}

/* reference: engine.py:145-158 Engine.step = update + prepare + clock; manager.py:105-115 */
static void orc_step_scenario(orc_engine* e, int s, const int32_t* act_agent, const int32_t* act_type,
                              const double* act_data, int K) {
    const orc_static* st = &e->st;
    double dt = st->interval;
    const int32_t* act = e->active + (int64_t)s * st->A;
    int n = e->n_active[s];
    e->agent_steps[s] += n;
    for (int i = 0; i < n; i++) {
        int a = act[i];
        int ext = -1;
        for (int k = 0; k < K; k++)
            if (act_agent && act_agent[(int64_t)s * K + k] == a && act_type[(int64_t)s * K + k] != ORC_ACT_NONE) ext = k;
        if (ext >= 0) {
            orc_apply_action(e, s, a, act_type[(int64_t)s * K + ext], act_data + ((int64_t)s * K + ext) * 5, dt, 0);
        } else if (st->reactive) {
            double d[5] = {0, 0, 0, 0, 0};
            orc_trajidm_action(e, s, a, 0.1 /* get_action(obs, dt=0.1) default, idm_policy.py:196 */, &d[0], &d[1]);
            orc_apply_action(e, s, a, ORC_ACT_BICYCLE, d, dt, 0);
        } else {
            /* reference: policy/expert_policy.py:18-33 StateExpert */
            double d[5];
            int wp_f32 = orc_policy_waypoint(e, s, a, d);
            orc_apply_action(e, s, a, ORC_ACT_STATE, d, dt, wp_f32);
        }
    }
    orc_departure_arrival(e, s, e->tick[s], st->allow_new_obj);
    orc_prepare_obs(e, s);
    e->tick[s]++;
}

/* reference: agent/agent.py:106-171 Agent.__init__ (xy home), manager.py:117-126 reset,
 * engine.py:132-143 Engine.reset */
static void orc_reset_scenario(orc_engine* e, int s) {
    const orc_static* st = &e->st;
    int A = st->A, T = st->T;
    for (int a = 0; a < A; a++) {
        int64_t ia = IDX_SA(e, s, a);
        e->status[ia] = ORC_WAITING;
        e->cursor[ia] = 0;
        e->lead_valid[ia] = 0;
        e->lead_dis[ia] = INFINITY;
        e->lead_v[ia] = 0;
        e->coll_count[ia] = 0;
        e->nearest_edge[ia] = -1;
        e->offroad[ia] = 0;
        e->ring_valid[ia] = 0;
        memset(e->ring + ia * ORC_H * 5, 0, sizeof(double) * ORC_H * 5);
        e->traj_len[ia] = st->traj_len[ia];
        e->traj_f32[ia] = 0;
        e->stale_valid[ia] = 0;
        memcpy(e->traj_xy + 2 * ia * T, st->traj_xy + 2 * ia * T, sizeof(double) * 2 * T);
        memcpy(e->traj_vel + 2 * ia * T, st->traj_vel + 2 * ia * T, sizeof(double) * 2 * T);
        memcpy(e->traj_h + ia * T, st->traj_h + ia * T, sizeof(double) * T);
        if (a >= st->n_agents[s]) {
            e->x[ia] = e->y[ia] = e->h[ia] = e->v[ia] = 0;
            continue;
        }
        double vx = st->traj_vel[2 * ia * T], vy = st->traj_vel[2 * ia * T + 1];
        e->x[ia] = st->home_xy[2 * ia];
        e->y[ia] = st->home_xy[2 * ia + 1];
        e->v[ia] = sqrt(vx * vx + vy * vy); /* agent.py:144-146 */
        e->h[ia] = st->traj_h[ia * T];
        orc_ring_push(e, s, a, e->x[ia], e->y[ia], e->h[ia], e->v[ia], 0); /* agent.py:159-165 */
    }
    e->tick[s] = 0;
    e->n_active[s] = 0;
    e->wait_ptr[s] = 0;
    e->agent_steps[s] = 0;
    orc_departure_arrival(e, s, 0, 1);
    orc_prepare_obs(e, s);
}

/* ----------------------------------------------------------- collision (A8) */

/* reference: utils/polygon.py:238-268 corners_from_bboxes */
static void orc_corners(double x, double y, double l, double w, double yaw, double* X, double* Y) {
    double c = cos(yaw), s = sin(yaw);
    X[0] = x + l / 2 * c + w / 2 * s;
    X[1] = x - l / 2 * c + w / 2 * s;
    X[2] = x - l / 2 * c - w / 2 * s;
    X[3] = x + l / 2 * c - w / 2 * s;
    Y[0] = y + l / 2 * s - w / 2 * c;
    Y[1] = y - l / 2 * s - w / 2 * c;
    Y[2] = y - l / 2 * s + w / 2 * c;
    Y[3] = y + l / 2 * s + w / 2 * c;
}

/* reference: utils/polygon.py:206-230 _overlap_a_over_b: project both boxes' corners on
 * `first`'s axes (np.matmul with [[c,-s],[s,c]]), overlap iff both extents are > 0 */
static int orc_overlap_a_over_b(const double* fX, const double* fY, double fyaw, const double* sX, const double* sY) {
    double c = cos(fyaw), s = sin(fyaw);
    for (int axis = 0; axis < 2; axis++) {
        double n0 = axis == 0 ? c : -s, n1 = axis == 0 ? s : c;
        double min_a = INFINITY, max_a = -INFINITY, min_b = INFINITY, max_b = -INFINITY;
        for (int k = 0; k < 4; k++) {
            double pa = orc_dot2(fX[k], n0, fY[k], n1);
            double pb = orc_dot2(sX[k], n0, sY[k], n1);
            if (pa < min_a) min_a = pa;
            if (pa > max_a) max_a = pa;
            if (pb < min_b) min_b = pb;
            if (pb > max_b) max_b = pb;
        }
        double hi = max_a < max_b ? max_a : max_b; /* np.minimum(max_a, max_b) */
        double lo = min_a > min_b ? min_a : min_b; /* np.maximum(min_a, min_b) */
        if (!(hi - lo > 0)) return 0;
    }
    return 1;
}

static int orc_collide(orc_engine* e, int s, int a, int b) {
    int64_t ia = IDX_SA(e, s, a), ib = IDX_SA(e, s, b);
    double aX[4], aY[4], bX[4], bY[4];
    orc_corners(e->x[ia], e->y[ia], e->st.length[ia], e->st.width[ia], e->h[ia], aX, aY);
    orc_corners(e->x[ib], e->y[ib], e->st.length[ib], e->st.width[ib], e->h[ib], bX, bY);
    return orc_overlap_a_over_b(aX, aY, e->h[ia], bX, bY) && orc_overlap_a_over_b(bX, bY, e->h[ib], aX, aY);
}

/* reference: agent/manager.py:313-327 check_collision_all (per-agent overlap COUNT, Q8) */
static void orc_collision_all(orc_engine* e, int s) {
    const int32_t* act = e->active + (int64_t)s * e->st.A;
    int n = e->n_active[s];
    for (int a = 0; a < e->st.A; a++) e->coll_count[IDX_SA(e, s, a)] = 0;
    for (int i = 0; i < n; i++) {
        int cnt = 0;
        for (int j = 0; j < n; j++)
            if (i != j && orc_collide(e, s, act[i], act[j])) cnt++;
        e->coll_count[IDX_SA(e, s, act[i])] = cnt;
    }
}

/* reference: agent/manager.py:289-311 check_collision(agent_id) for lane=None agents */
static int orc_collision_one(orc_engine* e, int s, int a) {
    if (e->status[IDX_SA(e, s, a)] != ORC_ACTIVE) return 0;
    const int32_t* act = e->active + (int64_t)s * e->st.A;
    for (int j = 0; j < e->n_active[s]; j++)
        if (act[j] != a && orc_collide(e, s, a, act[j])) return 1;
    return 0;
}

/* ------------------------------------------------------------- off-road (A9) */

/* reference: junction/junction.py:365-385 check_offroad: nearest resampled edge point
 * (first-index argmin), off-road iff cross(pt - pos, dir[pt]) < 0 */
static int orc_offroad_pos(const orc_static* st, int s, double px, double py, int32_t* nearest) {
    int n = st->n_edge[s];
    if (n == 0) {
        *nearest = -1;
        return 0;
    }
    const double* pts = st->edge_xy + 2 * (int64_t)s * st->E;
    const double* dir = st->edge_dir + 2 * (int64_t)s * st->E;
    int k = orc_argmin_pts(pts, n, px, py);
    *nearest = k;
    double ax = pts[2 * k] - px, ay = pts[2 * k + 1] - py;
    double cross = ax * dir[2 * k + 1] - ay * dir[2 * k]; /* np.cross on 2-vectors */
    return cross < 0;
}

/* ----------------------------------------------------------------- public API */

#define ALLOC(p, n) (p) = calloc((size_t)(n), sizeof(*(p)))

orc_engine* orc_create(const orc_static* st) {
    orc_engine* e = (orc_engine*)calloc(1, sizeof(orc_engine));
    e->st = *st;
    int64_t S = st->S, A = st->A, T = st->T, SA = S * A;
    ALLOC(e->traj_len, SA);
    ALLOC(e->traj_xy, SA * T * 2);
    ALLOC(e->traj_vel, SA * T * 2);
    ALLOC(e->traj_h, SA * T);
    ALLOC(e->traj_f32, SA);
    ALLOC(e->stale_valid, SA);
    ALLOC(e->stale_f32, SA);
    ALLOC(e->stale_wp, SA * 5);
    ALLOC(e->tick, S);
    ALLOC(e->n_active, S);
    ALLOC(e->wait_ptr, S);
    ALLOC(e->active, SA);
    ALLOC(e->status, SA);
    ALLOC(e->cursor, SA);
    ALLOC(e->x, SA);
    ALLOC(e->y, SA);
    ALLOC(e->h, SA);
    ALLOC(e->v, SA);
    ALLOC(e->lead_valid, SA);
    ALLOC(e->lead_dis, SA);
    ALLOC(e->lead_v, SA);
    ALLOC(e->ring, SA * ORC_H * 5);
    ALLOC(e->ring_valid, SA);
    ALLOC(e->coll_count, SA);
    ALLOC(e->nearest_edge, SA);
    ALLOC(e->offroad, SA);
    ALLOC(e->agent_steps, S);
    for (int s = 0; s < S; s++) orc_reset_scenario(e, s);
    return e;
}

void orc_destroy(orc_engine* e) {
    if (!e) return;
    free(e->traj_len); free(e->traj_xy); free(e->traj_vel); free(e->traj_h); free(e->traj_f32);
    free(e->stale_valid); free(e->stale_f32); free(e->stale_wp);
    free(e->tick); free(e->n_active); free(e->wait_ptr); free(e->active); free(e->status); free(e->cursor);
    free(e->x); free(e->y); free(e->h); free(e->v); free(e->lead_valid); free(e->lead_dis); free(e->lead_v);
    free(e->ring); free(e->ring_valid); free(e->coll_count); free(e->nearest_edge); free(e->offroad);
    free(e->agent_steps);
    free(e);
}

void orc_reset(orc_engine* e, const int32_t* scenario_ids, int n) {
    if (!scenario_ids)
        for (int s = 0; s < e->st.S; s++) orc_reset_scenario(e, s);
    else
        for (int i = 0; i < n; i++) orc_reset_scenario(e, scenario_ids[i]);
}

void orc_step(orc_engine* e, const int32_t* act_agent, const int32_t* act_type, const double* act_data, int K) {
    for (int s = 0; s < e->st.S; s++) orc_step_scenario(e, s, act_agent, act_type, act_data, K);
}

/* collision counts + nearest edge + off-road flag for every ACTIVE agent (post-step state);
 * reference: manager.py:313-345 check_collision_all / get_offroad_rate -> junction.py:387-409 */
static void orc_metrics_scenario(orc_engine* e, int s) {
    orc_collision_all(e, s);
    const int32_t* act = e->active + (int64_t)s * e->st.A;
    for (int a = 0; a < e->st.A; a++) {
        e->nearest_edge[IDX_SA(e, s, a)] = -1;
        e->offroad[IDX_SA(e, s, a)] = 0;
    }
    for (int i = 0; i < e->n_active[s]; i++) {
        int64_t ia = IDX_SA(e, s, act[i]);
        e->offroad[ia] = (uint8_t)orc_offroad_pos(&e->st, s, e->x[ia], e->y[ia], &e->nearest_edge[ia]);
    }
}

void orc_metrics(orc_engine* e) {
    for (int s = 0; s < e->st.S; s++) orc_metrics_scenario(e, s);
}

/* reference: agent.py:482-494 Agent.is_offroad for ANY agent (active or not) */
int orc_is_offroad(orc_engine* e, int s, int a, int32_t* nearest) {
    int64_t ia = IDX_SA(e, s, a);
    return orc_offroad_pos(&e->st, s, e->x[ia], e->y[ia], nearest);
}

int orc_check_collision(orc_engine* e, int s, int a) { return orc_collision_one(e, s, a); }

/* reference: policy/type.py:53-62 Trajectory.get_route + agent.py:468-480 Agent.get_route:
 * rows cursor+1.. of traj_xy minus the current xy, times [[c,-s],[s,c]], zero padded to (10,2) */
void orc_route(orc_engine* e, int s, int a, double* out /* [10][2] */) {
    int64_t ia = IDX_SA(e, s, a);
    memset(out, 0, sizeof(double) * ORC_ROUTE * 2);
    double c = cos(e->h[ia]), sn = sin(e->h[ia]);
    const double* xy = e->traj_xy + 2 * TRAJ(e, s, a);
    int n = e->traj_len[ia];
    for (int r = 0; r < ORC_ROUTE; r++) {
        int k = e->cursor[ia] + 1 + r;
        if (k >= n) break;
        double dx = xy[2 * k] - e->x[ia], dy = xy[2 * k + 1] - e->y[ia];
        out[2 * r] = orc_dot2(dx, c, dy, sn);
        out[2 * r + 1] = orc_dot2(dx, -sn, dy, c);
    }
}

/* reference: engine.py:301-357 Engine._compute_rewards for agent a of scenario s.
 * terms: forward, route_progress, smooth, collision, offroad, destination.
 * flags: bit0 terminated; truncated is the caller's (engine.py:197-200). */
void orc_rewards(orc_engine* e, int s, int a, double* terms /* [6] */, int32_t* terminated) {
    int64_t ia = IDX_SA(e, s, a);
    for (int k = 0; k < 6; k++) terms[k] = 0.0;
    const double* r = e->ring + ia * ORC_H * 5;
    const double* xy = e->traj_xy + 2 * TRAJ(e, s, a);
    int n = e->traj_len[ia];
    if (e->ring_valid[ia] > 1) {
        const double* p0 = r + 5 * (ORC_H - 2);
        const double* p1 = r + 5 * (ORC_H - 1);
        double cur_v = orc_norm2_vec(p1[2], p1[3]);
        double d0, s0, d1, s1;
        int32_t i0, i1;
        int f32 = e->traj_f32[ia];
        orc_frenet(xy, n, p0[0], p0[1], &d0, &s0, &i0, f32);
        orc_frenet(xy, n, p1[0], p1[1], &d1, &s1, &i1, f32);
        if (f32) { /* float32 s values: the subtraction and the division stay float32 */
            terms[0] = (double)((float)s1 - (float)s0) - (d1 - d0);
            terms[1] = (double)((float)s1 / (float)orc_polyline_length(xy, n, 1));
        } else {
            terms[0] = (s1 - s0) - (d1 - d0);
            terms[1] = s1 / orc_polyline_length(xy, n, 0);
        }
        double steer = orc_wrap_angle(p1[4] - p0[4]) / orc_norm2_vec(p1[0] - p0[0], p1[1] - p0[1]);
        steer = np_clip(steer, -0.3, 0.3);
        double sm = 1 / cur_v - fabs(steer);
        terms[2] = sm < 0 ? sm : 0.0; /* min(0, x): NaN -> 0 */
    }
    int collision = orc_collision_one(e, s, a);
    terms[3] = collision ? -1.0 : 0.0;
    int32_t ne;
    int off = orc_offroad_pos(&e->st, s, e->x[ia], e->y[ia], &ne);
    terms[4] = off ? -1.0 : 0.0;
    if (e->status[ia] == ORC_IDLE || e->status[ia] == ORC_FINISHED) {
        double dis = orc_norm2_vec(xy[2 * (n - 1)] - e->x[ia], xy[2 * (n - 1) + 1] - e->y[ia]);
        terms[5] = dis < 2.5 ? 10.0 : -5.0;
    }
    *terminated = collision || off;
}

/* reference: agent.py:462-466 set_ref_trajectory + engine.py:286-297 (arrays arrive as float32) */
void orc_set_ref_trajectory(orc_engine* e, int s, int a, int T, const float* xy, const float* vel, const float* h) {
    int64_t ia = IDX_SA(e, s, a);
    if (T > e->st.T) T = e->st.T;
    if (e->status[ia] == ORC_ACTIVE && !e->stale_valid[ia]) {
        e->stale_f32[ia] = (uint8_t)orc_policy_waypoint(e, s, a, e->stale_wp + 5 * ia);
        e->stale_valid[ia] = 1;
    }
    e->traj_len[ia] = T;
    e->traj_f32[ia] = 1;
    e->cursor[ia] = 0;
    for (int k = 0; k < T; k++) {
        e->traj_xy[2 * (TRAJ(e, s, a) + k)] = xy[2 * k];
        e->traj_xy[2 * (TRAJ(e, s, a) + k) + 1] = xy[2 * k + 1];
        e->traj_vel[2 * (TRAJ(e, s, a) + k)] = vel[2 * k];
        e->traj_vel[2 * (TRAJ(e, s, a) + k) + 1] = vel[2 * k + 1];
        e->traj_h[TRAJ(e, s, a) + k] = h[k];
    }
}

/* cumulative-length table used by the product for Trajectory.s: s_table[k] = np.sum(seg[:k]) */
void orc_s_table(const double* xy, int n, double* out /* [n] */, int f32) {
    for (int k = 0; k < n; k++) out[k] = orc_seg_sum(xy, n - 1, k, f32);
}

/* ------------------------------------------------------------------ views */

typedef struct {
    int32_t *tick, *n_active, *active, *status, *cursor, *ring_valid, *coll_count, *nearest_edge, *traj_len;
    double *x, *y, *h, *v, *lead_dis, *lead_v, *ring, *traj_xy, *traj_vel, *traj_h;
    uint8_t *lead_valid, *offroad;
    int64_t* agent_steps;
} orc_views_t;

void orc_views(orc_engine* e, orc_views_t* v) {
    v->tick = e->tick; v->n_active = e->n_active; v->active = e->active; v->status = e->status;
    v->cursor = e->cursor; v->ring_valid = e->ring_valid; v->coll_count = e->coll_count;
    v->nearest_edge = e->nearest_edge; v->traj_len = e->traj_len;
    v->x = e->x; v->y = e->y; v->h = e->h; v->v = e->v; v->lead_dis = e->lead_dis; v->lead_v = e->lead_v;
    v->ring = e->ring; v->traj_xy = e->traj_xy; v->traj_vel = e->traj_vel; v->traj_h = e->traj_h;
    v->lead_valid = e->lead_valid; v->offroad = e->offroad; v->agent_steps = e->agent_steps;
}

/* ------------------------------------------- multi-threaded rollout (baseline) */

typedef struct {
    orc_engine* e;
    int s0, s1, n_ticks, with_metrics;
} orc_job;

static void* orc_worker(void* p) {
    orc_job* j = (orc_job*)p;
    for (int s = j->s0; s < j->s1; s++)
        for (int t = 0; t < j->n_ticks; t++) {
            orc_step_scenario(j->e, s, NULL, NULL, NULL, 0);
            if (j->with_metrics) orc_metrics_scenario(j->e, s);
        }
    return NULL;
}

/* Steps every scenario n_ticks times on n_threads host threads (scenarios are independent,
 * SURVEY.md §8e); returns the number of agent updates performed. */
int64_t orc_rollout_mt(orc_engine* e, int n_ticks, int with_metrics, int n_threads) {
    int S = e->st.S;
    if (n_threads < 1) n_threads = 1;
    if (n_threads > S) n_threads = S;
    int64_t before = 0, after = 0;
    for (int s = 0; s < S; s++) before += e->agent_steps[s];
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * n_threads);
    orc_job* jobs = (orc_job*)malloc(sizeof(orc_job) * n_threads);
    for (int t = 0; t < n_threads; t++) {
        jobs[t].e = e;
        jobs[t].s0 = (int)((int64_t)S * t / n_threads);
        jobs[t].s1 = (int)((int64_t)S * (t + 1) / n_threads);
        jobs[t].n_ticks = n_ticks;
        jobs[t].with_metrics = with_metrics;
        pthread_create(&th[t], NULL, orc_worker, &jobs[t]);
    }
    for (int t = 0; t < n_threads; t++) pthread_join(th[t], NULL);
    for (int s = 0; s < S; s++) after += e->agent_steps[s];
    free(th);
    free(jobs);
    return after - before;
}
