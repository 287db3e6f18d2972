// lcsim_encoder.cuh — encoder-input builder for ALL listed agents of every scenario: what
// AgentEncoder.forward (motion_diff/modules/agent_encoder.py:98-256) derives from the HeteroData before its attention
// layers, written straight out of the engine's state so that the model never calls torch_cluster:
//   x_a   (N,11,6)  |motion|, angle(head, motion), |vel|, angle(head, vel), length, width           (:153-166)
//   r_t   temporal edges (a,i) -> (a,10), i < 10, both steps valid, 10 - i <= time_span, with
//         |dpos|, angle(head_dst, dpos), wrap(dhead), i - 10                                          (:174-199)
//   pl2a  radius(x = agents, y = polyline centres, r): for every polyline the agents with d^2 < r^2 in ascending agent
//         order (cap 300), kept when the agent's last step is valid; |d|, angle(head_a, d), wrap(head_pl - head_a),
//         d = pos_pl - pos_a                                                                          (:202-231)
//   a2a   radius_graph(x = agents, r, loop = False): for every target agent the sources with d^2 < r^2 in ascending
//         order (cap 300), both valid; |d|, angle(head_dst, d), wrap(head_src - head_dst), d = pos_src - pos_dst (:233-255)
// Agent rows are the slots of the active list (the reference takes `set(ids)` order; the graph is permutation
// equivariant).  Edge lists are CSR: offsets per polyline / per target agent, sources ascending.  float32 like the
// reference's tensors; the neighbour arithmetic restates the torch_cluster call-site contract (parity unpinned at
// that boundary, see tests/obs_oracle.py), the feature formulas are those of motion_diff/modules/utils.py:106-119.
#pragma once

struct LceArgs {
    lcsim_b200_encoder_desc d;
    float r2_a2a, r2_pl2a;
    const float* poly; // [S][P][3]
    const int32_t* n_poly;
    int P;
};
#define LCE_MAX_NBR 300 // max_num_neighbors of both radius calls

LCS_D float lce_angle_between(float cx, float cy, float nx, float ny) { // utils.py:106-113
    // the dot product is a torch .sum(): it accumulates from +0, so (-0) + (-0) comes out as +0 and the angle of a zero
    // neighbour vector is 0, not pi
    return atan2f(lcs_fmul(cx, ny) - lcs_fmul(cy, nx), lcs_fadd(lcs_fadd(0.0f, lcs_fmul(cx, nx)), lcs_fmul(cy, ny)));
}
LCS_D float lce_norm2(float x, float y) { return sqrtf(lcs_fadd(lcs_fmul(x, x), lcs_fmul(y, y))); }

// history entry k (shift-register order, 0 = oldest) of agent ia as float32 (junction.py:319-325 casts the float64
// history to float32): x, y, vx, vy, h; returns validity
LCS_D bool lce_hist(const LcsParams& p, size_t ia, int k, float* x, float* y, float* vx, float* vy, float* h) {
    const int head = p.ring_head[ia], cnt = p.ring_count[ia];
    const double* r = p.ring + (ia * LCS_H + (head + k) % LCS_H) * 5;
    *x = (float)r[0]; *y = (float)r[1]; *vx = (float)r[2]; *vy = (float)r[3]; *h = (float)r[4];
    return k >= LCS_H - cnt;
}

// per listed agent: x_a rows and temporal edges
LCS_D void lce_agent_rows(const LcsParams& p, const LceArgs& o, int s, int slot) {
    const int n = p.n_active[s];
    const size_t row = (size_t)s * p.Ap + slot;
    if (slot >= n) {
        if (o.d.x_a)
            for (int k = 0; k < LCS_H * 6; k++) o.d.x_a[row * LCS_H * 6 + k] = 0.f;
        if (o.d.r_t_valid)
            for (int k = 0; k < LCS_H - 1; k++) o.d.r_t_valid[row * (LCS_H - 1) + k] = 0;
        if (o.d.agent) o.d.agent[row] = -1;
        return;
    }
    const int a = p.active[(size_t)s * p.Ap + slot];
    const size_t ia = (size_t)s * p.Ap + a;
    if (o.d.agent) o.d.agent[row] = a;
    const float L = (float)p.length[ia], W = (float)p.width[ia];
    float x[LCS_H], y[LCS_H], vx[LCS_H], vy[LCS_H], h[LCS_H];
    bool valid[LCS_H];
    for (int k = 0; k < LCS_H; k++) valid[k] = lce_hist(p, ia, k, &x[k], &y[k], &vx[k], &vy[k], &h[k]);
    if (o.d.x_a)
        for (int k = 0; k < LCS_H; k++) {
            float* f = o.d.x_a + (row * LCS_H + k) * 6;
            const float mx = k == 0 ? 0.f : x[k] - x[k - 1], my = k == 0 ? 0.f : y[k] - y[k - 1]; // motion vector (:109-115)
            const float hx = cosf(h[k]), hy = sinf(h[k]);
            f[0] = lce_norm2(mx, my);
            f[1] = lce_angle_between(hx, hy, mx, my);
            f[2] = lce_norm2(vx[k], vy[k]);
            f[3] = lce_angle_between(hx, hy, vx[k], vy[k]);
            f[4] = valid[k] ? L : 0.f; // history_states.shape is filled for valid steps only (agent.py:75-84)
            f[5] = valid[k] ? W : 0.f;
        }
    if (o.d.r_t || o.d.r_t_valid) {
        const int last = LCS_H - 1;
        const float hx = cosf(h[last]), hy = sinf(h[last]);
        for (int i = 0; i < last; i++) {
            const bool on = valid[i] && valid[last] && (last - i) <= o.d.time_span;
            if (o.d.r_t_valid) o.d.r_t_valid[row * last + i] = on ? 1 : 0;
            if (o.d.r_t) {
                float* f = o.d.r_t + (row * last + i) * 4;
                const float rx = x[last] - x[i], ry = y[last] - y[i];
                f[0] = on ? lce_norm2(rx, ry) : 0.f;
                f[1] = on ? lce_angle_between(hx, hy, rx, ry) : 0.f;
                f[2] = on ? lcs_wrap_f32(h[last] - h[i]) : 0.f;
                f[3] = on ? (float)(i - last) : 0.f;
            }
        }
    }
}

// serial statement of one scenario's edge lists (host emulation; the device kernel does the same with block scans)
#ifdef LCSIM_HOSTEMU
static void lce_scenario_serial(const LcsParams& p, const LceArgs& o, int s) {
    const int n = p.n_active[s];
    for (int slot = 0; slot < p.Ap; slot++) lce_agent_rows(p, o, s, slot);
    std::vector<float> ax(n), ay(n), ah(n);
    for (int i = 0; i < n; i++) lcs_last_pose_f32(p, (size_t)s * p.Ap + p.active[(size_t)s * p.Ap + i], &ax[i], &ay[i], &ah[i]);
    // a2a: CSR over target slots
    {
        int32_t* off = o.d.a2a_offset ? o.d.a2a_offset + (size_t)s * (p.Ap + 1) : nullptr;
        int e = 0;
        for (int i = 0; i <= p.Ap; i++) {
            if (off) off[i] = e;
            if (i >= n) continue;
            int kept = 0;
            for (int j = 0; j < n && kept < LCE_MAX_NBR; j++) {
                if (j == i) continue;
                const float dx = ax[j] - ax[i], dy = ay[j] - ay[i];
                if (!(lcs_fadd(lcs_fmul(dx, dx), lcs_fmul(dy, dy)) < o.r2_a2a)) continue;
                kept++;
                if (e < o.d.cap_a2a) {
                    if (o.d.a2a_src) o.d.a2a_src[(size_t)s * o.d.cap_a2a + e] = j;
                    if (o.d.a2a_feat) {
                        float* f = o.d.a2a_feat + ((size_t)s * o.d.cap_a2a + e) * 3;
                        f[0] = lce_norm2(dx, dy);
                        f[1] = lce_angle_between(cosf(ah[i]), sinf(ah[i]), dx, dy);
                        f[2] = lcs_wrap_f32(ah[j] - ah[i]);
                    }
                }
                e++;
            }
        }
        if (o.d.n_a2a) o.d.n_a2a[s] = e;
    }
    // pl2a: CSR over polylines
    {
        const int np_ = o.poly ? o.n_poly[s] : 0;
        int32_t* off = o.d.pl2a_offset ? o.d.pl2a_offset + (size_t)s * (o.P + 1) : nullptr;
        int e = 0;
        for (int q = 0; q <= o.P; q++) {
            if (off) off[q] = e;
            if (q >= np_) continue;
            const float* c = o.poly + ((size_t)s * o.P + q) * 3;
            int kept = 0;
            for (int i = 0; i < n && kept < LCE_MAX_NBR; i++) {
                const float dx = c[0] - ax[i], dy = c[1] - ay[i];
                if (!(lcs_fadd(lcs_fmul(dx, dx), lcs_fmul(dy, dy)) < o.r2_pl2a)) continue;
                kept++;
                if (e < o.d.cap_pl2a) {
                    if (o.d.pl2a_agent) o.d.pl2a_agent[(size_t)s * o.d.cap_pl2a + e] = i;
                    if (o.d.pl2a_feat) {
                        float* f = o.d.pl2a_feat + ((size_t)s * o.d.cap_pl2a + e) * 3;
                        f[0] = lce_norm2(dx, dy);
                        f[1] = lce_angle_between(cosf(ah[i]), sinf(ah[i]), dx, dy);
                        f[2] = lcs_wrap_f32(c[2] - ah[i]);
                    }
                }
                e++;
            }
        }
        if (o.d.n_pl2a) o.d.n_pl2a[s] = e;
    }
}
#endif

#ifndef LCSIM_HOSTEMU
// inclusive block scan of one int per thread (blockDim.x <= 1024); returns the exclusive prefix, *total = block sum
LCS_D int lce_block_scan(int v, int* warp_sums, int* total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += o;
    }
    __syncthreads(); // warp_sums may still be read from the previous call
    if (lane == 31) warp_sums[w] = inc;
    __syncthreads();
    int base = 0, tot = 0;
    for (int k = 0; k < nw; k++) {
        const int ws = warp_sums[k];
        if (k < w) base += ws;
        tot += ws;
    }
    *total = tot;
    return base + inc - v;
}

// One CTA per scenario.  Counting pass, block scan, fill pass for the two edge lists; the agents' last poses live in
// shared memory (each polyline / target walks them in ascending order, which is also the output order).
__global__ void __launch_bounds__(128) lce_encoder_kernel(const LcsParams p, const LceArgs o) {
    extern __shared__ float lce_sm[]; // [5][Ap]: x, y, h, cos h, sin h of the listed agents
    __shared__ int warp_sums[32];
    const int s = blockIdx.x, tid = threadIdx.x, n = p.n_active[s], Ap = p.Ap;
    float *ax = lce_sm, *ay = ax + Ap, *ah = ay + Ap, *ac = ah + Ap, *as = ac + Ap;
    for (int slot = tid; slot < Ap; slot += blockDim.x) {
        lce_agent_rows(p, o, s, slot);
        if (slot < n) {
            float x, y, h;
            lcs_last_pose_f32(p, (size_t)s * Ap + p.active[(size_t)s * Ap + slot], &x, &y, &h);
            ax[slot] = x; ay[slot] = y; ah[slot] = h; ac[slot] = cosf(h); as[slot] = sinf(h);
        }
    }
    __syncthreads();
    // ---- a2a: thread = target slot (chunks of blockDim)
    int base = 0;
    for (int i0 = 0; i0 < Ap; i0 += blockDim.x) {
        const int i = i0 + tid;
        int cnt = 0;
        if (i < n) {
            const float xi = ax[i], yi = ay[i];
            for (int j = 0; j < n && cnt < LCE_MAX_NBR; j++) {
                const float dx = ax[j] - xi, dy = ay[j] - yi;
                cnt += (j != i && lcs_fadd(lcs_fmul(dx, dx), lcs_fmul(dy, dy)) < o.r2_a2a) ? 1 : 0;
            }
        }
        int total;
        const int at = base + lce_block_scan(cnt, warp_sums, &total);
        if (i < Ap && o.d.a2a_offset) o.d.a2a_offset[(size_t)s * (Ap + 1) + i] = at;
        if (i < n && cnt) {
            const float xi = ax[i], yi = ay[i], hi = ah[i], ci = ac[i], si = as[i];
            int e = at, kept = 0;
            for (int j = 0; j < n && kept < LCE_MAX_NBR; j++) {
                const float dx = ax[j] - xi, dy = ay[j] - yi;
                if (j == i || !(lcs_fadd(lcs_fmul(dx, dx), lcs_fmul(dy, dy)) < o.r2_a2a)) continue;
                kept++;
                if (e < o.d.cap_a2a) {
                    if (o.d.a2a_src) o.d.a2a_src[(size_t)s * o.d.cap_a2a + e] = j;
                    if (o.d.a2a_feat) {
                        float* f = o.d.a2a_feat + ((size_t)s * o.d.cap_a2a + e) * 3;
                        f[0] = lce_norm2(dx, dy);
                        f[1] = lce_angle_between(ci, si, dx, dy);
                        f[2] = lcs_wrap_f32(ah[j] - hi);
                    }
                }
                e++;
            }
        }
        base += total;
    }
    if (tid == 0) {
        if (o.d.a2a_offset) o.d.a2a_offset[(size_t)s * (Ap + 1) + Ap] = base;
        if (o.d.n_a2a) o.d.n_a2a[s] = base;
    }
    // ---- pl2a: thread = polyline (chunks of blockDim)
    const int np_ = o.poly ? o.n_poly[s] : 0;
    base = 0;
    for (int q0 = 0; q0 < o.P; q0 += blockDim.x) {
        const int q = q0 + tid;
        int cnt = 0;
        float cx = 0.f, cy = 0.f, ch = 0.f;
        if (q < np_) {
            const float* c = o.poly + ((size_t)s * o.P + q) * 3;
            cx = c[0]; cy = c[1]; ch = c[2];
            for (int i = 0; i < n && cnt < LCE_MAX_NBR; i++) {
                const float dx = cx - ax[i], dy = cy - ay[i];
                cnt += lcs_fadd(lcs_fmul(dx, dx), lcs_fmul(dy, dy)) < o.r2_pl2a ? 1 : 0;
            }
        }
        int total;
        const int at = base + lce_block_scan(cnt, warp_sums, &total);
        if (q < o.P && o.d.pl2a_offset) o.d.pl2a_offset[(size_t)s * (o.P + 1) + q] = at;
        if (cnt) {
            int e = at, kept = 0;
            for (int i = 0; i < n && kept < LCE_MAX_NBR; i++) {
                const float dx = cx - ax[i], dy = cy - ay[i];
                if (!(lcs_fadd(lcs_fmul(dx, dx), lcs_fmul(dy, dy)) < o.r2_pl2a)) continue;
                kept++;
                if (e < o.d.cap_pl2a) {
                    if (o.d.pl2a_agent) o.d.pl2a_agent[(size_t)s * o.d.cap_pl2a + e] = i;
                    if (o.d.pl2a_feat) {
                        float* f = o.d.pl2a_feat + ((size_t)s * o.d.cap_pl2a + e) * 3;
                        f[0] = lce_norm2(dx, dy);
                        f[1] = lce_angle_between(ac[i], as[i], dx, dy);
                        f[2] = lcs_wrap_f32(ch - ah[i]);
                    }
                }
                e++;
            }
        }
        base += total;
    }
    if (tid == 0) {
        if (o.d.pl2a_offset) o.d.pl2a_offset[(size_t)s * (o.P + 1) + o.P] = base;
        if (o.d.n_pl2a) o.d.n_pl2a[s] = base;
    }
}
#endif
