// lcsim_loader.hpp — scenario loader / packer in C++: map.pb + agents.pb of WOMD-shaped scenario
// directories -> the packed arrays of lcsim_b200_static_host, threaded over the scenarios.  Host code only; included by
// lcsim_b200.cu.
//
// What it restates (and must reproduce bit for bit, tests/test_loader.py compares with lcsim_b200/packer.py, which is
// pinned to the real reference):
//   * the protobuf wire format of the two files (message / field numbers: lcsim/protos/{geo,agent,map}/*.proto as
//     listed in lcsim_b200/proto.py) — a varint / length-delimited reader, unknown fields skipped;
//   * Engine.__init__'s map / agent construction (engine.py:84-130): one junction owning lanes, road lines and road
//     edges; agents with an xy home, one schedule x one trip (agent.py:106-171);
//   * the 0.5 m arc-length resampling of the road edges (junction.py:59-78 -> utils/polygon.py:90-188: numpy
//     linspace / cumsum / digitize semantics, pairwise summation of np.sum) and get_polyline_dir (polygon.py:78-86);
//   * departure ticks from the float64 clock accumulation (engine.py:157), the stable waiting order (manager.py:33-35),
//     Agent.dynamic_flag (agent.py:153-156);
//   * the polyline centres of the observation builder (lane/lane.py:108-124 -> polygon.py:11-75).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace lcsload {

// ---------------------------------------------------------------- protobuf wire reader
struct Reader {
    const uint8_t *p, *end;
    bool ok = true;
    Reader(const uint8_t* b, const uint8_t* e) : p(b), end(e) {}
    bool more() const { return ok && p < end; }
    uint64_t varint() {
        uint64_t v = 0;
        for (int shift = 0; shift < 64; shift += 7) {
            if (p >= end) { ok = false; return 0; }
            const uint8_t b = *p++;
            v |= (uint64_t)(b & 0x7f) << shift;
            if (!(b & 0x80)) return v;
        }
        ok = false;
        return 0;
    }
    // next field: number and wire type
    bool tag(int& field, int& wt) {
        if (!more()) return false;
        const uint64_t t = varint();
        field = (int)(t >> 3);
        wt = (int)(t & 7);
        return ok;
    }
    double f64() {
        if (end - p < 8) { ok = false; return 0; }
        double d;
        memcpy(&d, p, 8);
        p += 8;
        return d;
    }
    Reader sub() { // length-delimited payload
        const uint64_t n = varint();
        if (!ok || (uint64_t)(end - p) < n) { // truncated: this reader and the returned one are both failed
            ok = false;
            Reader bad(end, end);
            bad.ok = false;
            return bad;
        }
        Reader r(p, p + n);
        p += n;
        return r;
    }
    void skip(int wt) {
        switch (wt) {
            case 0: varint(); break;
            case 1: if (end - p < 8) ok = false; else p += 8; break;
            case 2: sub(); break;
            case 5: if (end - p < 4) ok = false; else p += 4; break;
            default: ok = false;
        }
    }
    // repeated int32: packed (wire type 2) or one varint per tag (wire type 0)
    void ints(int wt, std::vector<int32_t>& out) {
        if (wt == 2) {
            Reader r = sub();
            while (r.more()) out.push_back((int32_t)r.varint());
            ok = ok && r.ok;
        } else if (wt == 0) {
            out.push_back((int32_t)varint());
        } else {
            skip(wt);
        }
    }
};

struct XY { double x = 0, y = 0; };
static XY read_xy(Reader r) { // geo.XYPosition: x = 1, y = 2 (double)
    XY v;
    int f, wt;
    while (r.tag(f, wt)) {
        if (f == 1 && wt == 1) v.x = r.f64();
        else if (f == 2 && wt == 1) v.y = r.f64();
        else r.skip(wt);
    }
    return v;
}
typedef std::vector<XY> Poly;

struct Scenario {
    // agents
    std::vector<int32_t> id, type, traj_len;
    std::vector<double> length, width, departure;
    std::vector<XY> home;
    std::vector<std::vector<double>> txy, tvel, th; // per agent: [n][2], [n][2], [n]
    int ads_index = 0;
    // map
    int junction_id = 0, n_junctions = 0;
    std::vector<Poly> lanes;
    std::vector<int32_t> lane_id;
    std::vector<Poly> edges; // raw nodes
    std::vector<int32_t> edge_type;
    // derived
    std::vector<double> edge_xy, edge_dir; // resampled, concatenated [E][2]
    std::vector<int32_t> edge_poly;
    std::vector<double> centres; // [P][3]
    std::string err;
};

static bool read_file(const std::string& path, std::vector<uint8_t>& buf) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    const long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    buf.resize(n > 0 ? (size_t)n : 0);
    const size_t got = n > 0 ? fread(buf.data(), 1, (size_t)n, f) : 0;
    fclose(f);
    return got == buf.size();
}

// map.Polyline / RoadLine / RoadEdge nodes
static void read_nodes(Reader r, int node_field, Poly& out, int32_t* type) {
    int f, wt;
    while (r.tag(f, wt)) {
        if (f == node_field && wt == 2) out.push_back(read_xy(r.sub()));
        else if (type && f == 1 && wt == 0) *type = (int32_t)r.varint();
        else r.skip(wt);
    }
}

static bool parse_map(const std::vector<uint8_t>& buf, Scenario& sc) {
    Reader m(buf.data(), buf.data() + buf.size());
    int f, wt;
    while (m.tag(f, wt)) {
        if (f == 2 && wt == 2) { // map.Lane: id = 1, center_line = 7 (Polyline: nodes = 1)
            Reader l = m.sub();
            int32_t id = 0;
            Poly pts;
            int lf, lwt;
            while (l.tag(lf, lwt)) {
                if (lf == 1 && lwt == 0) id = (int32_t)l.varint();
                else if (lf == 7 && lwt == 2) read_nodes(l.sub(), 1, pts, nullptr);
                else l.skip(lwt);
            }
            if (!l.ok) return false;
            sc.lanes.push_back(std::move(pts));
            sc.lane_id.push_back(id);
        } else if (f == 4 && wt == 2) { // map.Junction: id = 1, road_edges = 7 (RoadEdge: type = 1, nodes = 2)
            Reader j = m.sub();
            sc.n_junctions++;
            int jf, jwt;
            while (j.tag(jf, jwt)) {
                if (jf == 1 && jwt == 0) sc.junction_id = (int32_t)j.varint();
                else if (jf == 7 && jwt == 2) {
                    Poly pts;
                    int32_t type = 0;
                    read_nodes(j.sub(), 2, pts, &type);
                    sc.edges.push_back(std::move(pts));
                    sc.edge_type.push_back(type);
                } else j.skip(jwt);
            }
            if (!j.ok) return false;
        } else {
            m.skip(wt);
        }
    }
    return m.ok;
}

static bool parse_agents(const std::vector<uint8_t>& buf, Scenario& sc) {
    Reader top(buf.data(), buf.data() + buf.size());
    int f, wt;
    while (top.tag(f, wt)) {
        if (f == 2 && wt == 0) {
            sc.ads_index = (int)top.varint();
        } else if (f == 1 && wt == 2) { // agent.Agent
            Reader a = top.sub();
            int32_t id = 0, type = 0;
            double length = 0, width = 0;
            XY home;
            bool has_xy_home = false, has_lane_home = false;
            int n_sched = 0, n_trips = 0, loop_count = 0;
            bool sched_dep = false, trip_dep = false, trip_wait = false;
            double sched_dep_t = 0, trip_dep_t = 0, trip_wait_t = 0;
            std::vector<double> xy, vel, hd;
            int af, awt;
            while (a.tag(af, awt)) {
                if (af == 1 && awt == 0) id = (int32_t)a.varint();
                else if (af == 2 && awt == 2) { // AgentAttribute: type = 1, length = 2, width = 3
                    Reader at = a.sub();
                    int tf, twt;
                    while (at.tag(tf, twt)) {
                        if (tf == 1 && twt == 0) type = (int32_t)at.varint();
                        else if (tf == 2 && twt == 1) length = at.f64();
                        else if (tf == 3 && twt == 1) width = at.f64();
                        else at.skip(twt);
                    }
                } else if (af == 3 && awt == 2) { // home: geo.Position (lane_position = 1, xy_position = 3)
                    Reader h = a.sub();
                    int hf, hwt;
                    while (h.tag(hf, hwt)) {
                        if (hf == 3 && hwt == 2) { Reader x = h.sub(); has_xy_home = x.end > x.p; home = read_xy(x); }
                        else if (hf == 1 && hwt == 2) { Reader x = h.sub(); has_lane_home = x.end > x.p; }
                        else h.skip(hwt);
                    }
                } else if (af == 4 && awt == 2) { // Schedule: trips = 1, loop_count = 2, departure_time = 3
                    Reader s = a.sub();
                    n_sched++;
                    int sf, swt;
                    while (s.tag(sf, swt)) {
                        if (sf == 2 && swt == 0) loop_count = (int)s.varint();
                        else if (sf == 3 && swt == 1) { sched_dep = true; sched_dep_t = s.f64(); }
                        else if (sf == 1 && swt == 2) { // Trip: departure_time = 3, wait_time = 4, agent_states = 8
                            Reader t = s.sub();
                            n_trips++;
                            int tf, twt;
                            while (t.tag(tf, twt)) {
                                if (tf == 3 && twt == 1) { trip_dep = true; trip_dep_t = t.f64(); }
                                else if (tf == 4 && twt == 1) { trip_wait = true; trip_wait_t = t.f64(); }
                                else if (tf == 8 && twt == 2) { // AgentState: position = 1, velocity_x = 2, velocity_y = 3, heading = 4
                                    Reader st = t.sub();
                                    XY pos;
                                    double vx = 0, vy = 0, h = 0;
                                    int qf, qwt;
                                    while (st.tag(qf, qwt)) {
                                        if (qf == 1 && qwt == 2) pos = read_xy(st.sub());
                                        else if (qf == 2 && qwt == 1) vx = st.f64();
                                        else if (qf == 3 && qwt == 1) vy = st.f64();
                                        else if (qf == 4 && qwt == 1) h = st.f64();
                                        else st.skip(qwt);
                                    }
                                    xy.push_back(pos.x); xy.push_back(pos.y);
                                    vel.push_back(vx); vel.push_back(vy);
                                    hd.push_back(h);
                                } else t.skip(twt);
                            }
                        } else s.skip(swt);
                    }
                } else a.skip(awt);
            }
            if (!a.ok) return false;
            char msg[160];
            if (has_lane_home || !has_xy_home) { sc.err = "lane-position homes (LaneIDM agents) are not packed by this path"; return false; }
            if (n_sched != 1 || n_trips != 1 || loop_count != 1) {
                sc.err = "trajectory agents are expected to carry one schedule x one trip, loop_count=1";
                return false;
            }
            if (hd.empty()) {
                snprintf(msg, sizeof msg, "agent %d: empty agent_states (reference asserts, agent.py:142)", id);
                sc.err = msg;
                return false;
            }
            // packer.scenario_from_pb: trip.departure_time, else schedule.departure_time, else 0 + trip.wait_time, else 0
            const double dep = trip_dep ? trip_dep_t : sched_dep ? sched_dep_t : trip_wait ? 0.0 + trip_wait_t : 0.0;
            sc.id.push_back(id); sc.type.push_back(type); sc.length.push_back(length); sc.width.push_back(width);
            sc.departure.push_back(dep); sc.home.push_back(home);
            sc.traj_len.push_back((int32_t)hd.size());
            sc.txy.push_back(std::move(xy)); sc.tvel.push_back(std::move(vel)); sc.th.push_back(std::move(hd));
        } else {
            top.skip(wt);
        }
    }
    return top.ok;
}

// ---------------------------------------------------------------- numpy-faithful numerics
// np.add.reduce of a contiguous float64 vector (pairwise, 8-way unrolled blocks of <= 128)
static double np_sum(const double* a, int64_t n) {
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
    }
    int64_t n2 = n / 2;
    n2 -= n2 % 8;
    return np_sum(a, n2) + np_sum(a + n2, n - n2);
}

// packer.resample_polyline (polygon.py:90-188): returns [n][2]
static std::vector<double> resample(const Poly& pts, double interval) {
    const int n = (int)pts.size();
    std::vector<double> out;
    if (n < 2) {
        for (const XY& p : pts) { out.push_back(p.x); out.push_back(p.y); }
        return out;
    }
    std::vector<double> seg(n - 1), chord(n - 1), cum(n, 0.0);
    for (int k = 0; k + 1 < n; k++) {
        const double dx = pts[k + 1].x - pts[k].x, dy = pts[k + 1].y - pts[k].y;
        const double dx2 = dx * dx, dy2 = dy * dy;
        seg[k] = std::sqrt(dx2 + dy2);
    }
    const double total = np_sum(seg.data(), n - 1);
    const int n_out = (int)std::floor(total / interval) + 1;
    for (int k = 0; k + 1 < n; k++) chord[k] = seg[k] / total;
    for (int k = 0; k + 1 < n; k++) cum[k + 1] = cum[k] + chord[k]; // np.cumsum: running sum
    out.resize((size_t)n_out * 2);
    const double step = n_out > 1 ? 1.0 / (double)(n_out - 1) : 0.0;
    for (int i = 0; i < n_out; i++) {
        double u = (double)i * step;
        if (n_out > 1 && i == n_out - 1) u = 1.0; // np.linspace pins the end point
        // np.digitize(u, cum) with increasing bins = number of bins <= u
        int b = (int)(std::upper_bound(cum.begin(), cum.end(), u) - cum.begin());
        if (b <= 0 || u <= 0) b = 1;
        if (b >= n || u >= 1) b = n - 1;
        const double frac = (u - cum[b - 1]) / chord[b - 1];
        const double ddx = pts[b].x - pts[b - 1].x, ddy = pts[b].y - pts[b - 1].y;
        const double mx = ddx * frac, my = ddy * frac;
        out[2 * i] = pts[b - 1].x + mx;
        out[2 * i + 1] = pts[b - 1].y + my;
    }
    return out;
}

// packer.polyline_dir (polygon.py:78-86) of [n][2] points
static std::vector<double> poly_dir(const std::vector<double>& p) {
    const int n = (int)p.size() / 2;
    std::vector<double> d(p.size());
    for (int k = 0; k < n; k++) {
        const int prev = k == 0 ? n - 1 : k - 1; // np.roll
        d[2 * k] = p[2 * k] - p[2 * prev];
        d[2 * k + 1] = p[2 * k + 1] - p[2 * prev + 1];
    }
    if (n > 1) { d[0] = d[2]; d[1] = d[3]; }
    for (int k = 0; k < n; k++) {
        const double x2 = d[2 * k] * d[2 * k], y2 = d[2 * k + 1] * d[2 * k + 1];
        double nm = std::sqrt(x2 + y2);
        nm = nm < 1e-6 ? 1e-6 : (nm > 1e9 ? 1e9 : nm);
        d[2 * k] /= nm;
        d[2 * k + 1] /= nm;
    }
    return d;
}

// packer.depart_tick
static int32_t depart_tick(double departure, double start, double interval, int max_ticks) {
    double t = start;
    for (int j = 0; j <= max_ticks; j++) {
        if (departure <= t) return j;
        t += interval;
    }
    return 0x7fffffff;
}

static bool load_one(const std::string& dir, Scenario& sc) {
    std::vector<uint8_t> buf;
    if (!read_file(dir + "/map.pb", buf)) { sc.err = dir + "/map.pb: cannot read"; return false; }
    if (!parse_map(buf, sc)) { sc.err = dir + "/map.pb: malformed protobuf"; return false; }
    if (sc.n_junctions != 1) {
        sc.err = "trajectory-agent scenarios need exactly one junction (reference: junction/manager.py get_unique_junction)";
        return false;
    }
    if (!read_file(dir + "/agents.pb", buf)) { sc.err = dir + "/agents.pb: cannot read"; return false; }
    if (!parse_agents(buf, sc)) {
        if (sc.err.empty()) sc.err = dir + "/agents.pb: malformed protobuf";
        return false;
    }
    for (size_t a = 0; a < sc.id.size(); a++)
        for (size_t b = a + 1; b < sc.id.size(); b++)
            if (sc.id[a] == sc.id[b]) { sc.err = "duplicate agent ids"; return false; }
    // road edges: resampled points + directions, concatenated in pb order
    for (size_t k = 0; k < sc.edges.size(); k++) {
        const std::vector<double> r = resample(sc.edges[k], 0.5), d = poly_dir(r);
        sc.edge_xy.insert(sc.edge_xy.end(), r.begin(), r.end());
        sc.edge_dir.insert(sc.edge_dir.end(), d.begin(), d.end());
        sc.edge_poly.insert(sc.edge_poly.end(), r.size() / 2, (int32_t)k);
    }
    // polyline centres (packer.polyline_centres): 10-point pieces of the resampled lane centre-lines
    for (const Poly& lane : sc.lanes) {
        const std::vector<double> r = resample(lane, 0.5);
        const int n = (int)r.size() / 2;
        if (n == 0) continue;
        const std::vector<double> d = poly_dir(r);
        for (int i = 0; i < n; i += 10) {
            const int m = std::min(10, n - i);
            double acc[4] = {0, 0, 0, 0}; // x, y, dir_x, dir_y (z and dir_z are zero)
            for (int k = i; k < i + m; k++) {
                acc[0] += r[2 * k]; acc[1] += r[2 * k + 1]; acc[2] += d[2 * k]; acc[3] += d[2 * k + 1];
            }
            const float cx = (float)(acc[0] / m), cy = (float)(acc[1] / m), dx = (float)(acc[2] / m), dy = (float)(acc[3] / m);
            sc.centres.push_back((double)cx);
            sc.centres.push_back((double)cy);
            sc.centres.push_back((double)atan2f(dy, dx));
        }
    }
    return true;
}

} // namespace lcsload

struct lcsim_b200_scenarios {
    int S = 0, A = 0, T = 0, E = 0, P = 0;
    std::vector<int32_t> n_agents, ads_index, agent_id, agent_type, traj_len, depart_tick, wait_order, n_edge, edge_poly_id, n_poly,
        junction_id;
    std::vector<uint8_t> dynamic;
    std::vector<double> length, width, home_xy, traj_xy, traj_vel, traj_h, edge_xy, edge_dir, origin, poly;
    std::string err;
};

static int lcs_load_dirs(const char* const* dirs, int n_dirs, double start, double interval, int32_t total_steps, int n_threads,
                         lcsim_b200_scenarios* out) {
    using namespace lcsload;
    const int S = n_dirs;
    std::vector<Scenario> scs(S);
    std::vector<char> okv(S, 0);
    int nt = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    nt = std::max(1, std::min(std::min(nt, 64), S));
    auto run = [&](const std::function<void(int)>& fn) {
        std::vector<std::thread> th;
        for (int t = 0; t < nt; t++)
            th.emplace_back([&, t]() {
                for (int s = t; s < S; s += nt) fn(s);
            });
        for (auto& t : th) t.join();
    };
    run([&](int s) { okv[s] = load_one(dirs[s], scs[s]) ? 1 : 0; });
    for (int s = 0; s < S; s++)
        if (!okv[s]) {
            out->err = std::string(dirs[s]) + ": " + scs[s].err;
            return LCSIM_B200_ERR_ARG;
        }
    int A = 1, T = 1, E = 1, P = 1;
    for (const Scenario& sc : scs) {
        A = std::max(A, (int)sc.id.size());
        for (int32_t n : sc.traj_len) T = std::max(T, (int)n);
        E = std::max(E, (int)sc.edge_xy.size() / 2);
        P = std::max(P, (int)sc.centres.size() / 3);
    }
    E = (E + 31) / 32 * 32; // packer.pack pads E to whole 32-point blocks
    out->S = S; out->A = A; out->T = T; out->E = E; out->P = P;
    const size_t SA = (size_t)S * A, SAT = SA * T, SE = (size_t)S * E;
    out->n_agents.assign(S, 0); out->ads_index.assign(S, 0); out->n_edge.assign(S, 0); out->n_poly.assign(S, 0);
    out->junction_id.assign(S, 0);
    out->agent_id.assign(SA, -1); out->agent_type.assign(SA, 0); out->traj_len.assign(SA, 0);
    out->depart_tick.assign(SA, 0x7fffffff); out->wait_order.assign(SA, 0); out->dynamic.assign(SA, 0);
    out->length.assign(SA, 0.0); out->width.assign(SA, 0.0); out->home_xy.assign(SA * 2, 0.0);
    out->traj_xy.assign(SAT * 2, 0.0); out->traj_vel.assign(SAT * 2, 0.0); out->traj_h.assign(SAT, 0.0);
    out->edge_xy.assign(SE * 2, 0.0); out->edge_dir.assign(SE * 2, 0.0); out->edge_poly_id.assign(SE, 0);
    out->origin.assign((size_t)S * 2, 0.0); out->poly.assign((size_t)S * P * 3, 0.0);
    const int max_ticks = std::max(4 * total_steps, 1024);
    run([&](int s) {
        const Scenario& sc = scs[s];
        const int n = (int)sc.id.size();
        out->n_agents[s] = n;
        out->ads_index[s] = sc.ads_index;
        out->junction_id[s] = sc.junction_id;
        double mn[2] = {INFINITY, INFINITY}, mx[2] = {-INFINITY, -INFINITY};
        bool any = false;
        for (int a = 0; a < n; a++) {
            const size_t ia = (size_t)s * A + a;
            out->agent_id[ia] = sc.id[a];
            out->agent_type[ia] = sc.type[a];
            out->length[ia] = sc.length[a];
            out->width[ia] = sc.width[a];
            out->home_xy[2 * ia] = sc.home[a].x;
            out->home_xy[2 * ia + 1] = sc.home[a].y;
            const int t = sc.traj_len[a];
            out->traj_len[ia] = t;
            memcpy(&out->traj_xy[ia * T * 2], sc.txy[a].data(), sizeof(double) * 2 * t);
            memcpy(&out->traj_vel[ia * T * 2], sc.tvel[a].data(), sizeof(double) * 2 * t);
            memcpy(&out->traj_h[ia * T], sc.th[a].data(), sizeof(double) * t);
            double amn[2] = {INFINITY, INFINITY}, amx[2] = {-INFINITY, -INFINITY};
            for (int k = 0; k < t; k++)
                for (int c = 0; c < 2; c++) {
                    amn[c] = std::min(amn[c], sc.txy[a][2 * k + c]);
                    amx[c] = std::max(amx[c], sc.txy[a][2 * k + c]);
                }
            for (int c = 0; c < 2; c++) { mn[c] = std::min(mn[c], amn[c]); mx[c] = std::max(mx[c], amx[c]); }
            any = true;
            // agent.py:153-156: np.linalg.norm of the bbox extent (1-D vector: BLAS dot, x*x then fma) > 1 m and a vehicle
            const double ex = std::fabs(amx[0] - amn[0]), ey = std::fabs(amx[1] - amn[1]);
            const double diag = std::sqrt(std::fma(ey, ey, ex * ex));
            out->dynamic[ia] = (diag > 1.0 && sc.type[a] == 1) ? 1 : 0;
            out->depart_tick[ia] = depart_tick(sc.departure[a], start, interval, max_ticks);
        }
        std::vector<int> order(n);
        for (int a = 0; a < n; a++) order[a] = a;
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return sc.departure[a] < sc.departure[b]; });
        for (int a = 0; a < n; a++) out->wait_order[(size_t)s * A + a] = order[a];
        const int ne = (int)sc.edge_xy.size() / 2;
        out->n_edge[s] = ne;
        memcpy(&out->edge_xy[(size_t)s * E * 2], sc.edge_xy.data(), sizeof(double) * 2 * ne);
        memcpy(&out->edge_dir[(size_t)s * E * 2], sc.edge_dir.data(), sizeof(double) * 2 * ne);
        memcpy(&out->edge_poly_id[(size_t)s * E], sc.edge_poly.data(), sizeof(int32_t) * ne);
        for (int k = 0; k < ne; k++)
            for (int c = 0; c < 2; c++) {
                mn[c] = std::min(mn[c], sc.edge_xy[2 * k + c]);
                mx[c] = std::max(mx[c], sc.edge_xy[2 * k + c]);
                any = true;
            }
        // origin of the fp32 filter frame: np.round (half to even) of the middle of everything
        for (int c = 0; c < 2; c++) out->origin[2 * s + c] = any ? std::nearbyint(0.5 * (mn[c] + mx[c])) : 0.0;
        const int np_ = (int)sc.centres.size() / 3;
        out->n_poly[s] = np_;
        memcpy(&out->poly[(size_t)s * P * 3], sc.centres.data(), sizeof(double) * 3 * np_);
    });
    return 0;
}
