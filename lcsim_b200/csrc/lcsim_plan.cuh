// lcsim_plan.cuh — guidance costs with analytic gradients and batched plan metrics over
// planned trajectories, float32 like the reference's tensors.  Included by lcsim_b200.cu.
//
//   Repeller.loss        motion_diff/modules/guidance.py:17-49     pairwise distance hinge between the agents' plans
//   NoOffRoad.loss       motion_diff/modules/guidance.py:52-85  -> metrics/roadgraph.py:160-232 signed distance to the
//                        nearest road-edge point with the prior-point rule
//   compute_overlap_rate metrics/overlap.py:7-43 -> metrics/geometry.py:8-101 (oriented-box overlap, float32)
//   compute_offroad_rate metrics/roadgraph.py:36-88
//
// The reference evaluates the losses on the WHOLE Batch (all agents of all graphs in one pairwise tensor); an Engine
// holds one junction = one graph, so here every scenario is its own graph.  The gradients are those torch.autograd
// returns for the loss expressions (the counts in the denominators and the argmin / sign selections are constants of
// the differentiation; norm has gradient 0 at 0, like torch's norm backward).
#pragma once

struct LcpArgs {
    int S, A, T;
    const float* xy; // [S][A][T][2] (guidance) — or traj5 [S][A][T][5] x, y, length, width, yaw (metrics)
    const float* traj5;
    const uint8_t* mask; // [S][A] or null (all of n_agents[s])
    const int32_t* n_agents; // [S]
    float radius, eps;
    // road-edge points of the roadgraph (float32 world frame): x, y, dir_x, dir_y and the polyline id of every point
    const float4* edge; // [S][E]
    const int32_t* edge_id; // [S][E]
    const int32_t* n_edge; // [S]
    int E;
    // outputs
    float* grad; // [S][A][T][2] or null
    float* sd; // [S][A][T] signed distances or null
    float* part_sum; // [S][P] per-block partial sums
    int32_t* part_cnt; // [S][P]
    int P;
    float* loss; // [S]
    int32_t* count; // [S]
    float* rate; // [S]
};

// relu(1 - d / R) pair term of Repeller.loss for p_i against p_j; accumulates value, count and the unnormalised
// gradient direction (p_i - p_j) / d
LCS_D void lcp_repel_pair(float xi, float yi, float xj, float yj, float inv_r, float& sum, int& cnt, float& gx, float& gy) {
    const float dx = xi - xj, dy = yi - yj;
    const float d = sqrtf(lcs_fadd(lcs_fmul(dx, dx), lcs_fmul(dy, dy)));
    const float a = 1.0f - d * inv_r;
    if (a > 0.0f) {
        sum += a;
        cnt += 1;
        if (d > 0.0f) {
            gx += dx / d;
            gy += dy / d;
        }
    }
}

// compute_signed_distance_to_nearest_road_edge_point for one query given its nearest point index k (first argmin of
// the float32 squared distance): prior point k-1 (or k), half-plane rule, norm * sign.  *ux, *uy = (q - nearest) / dist.
LCS_D float lcp_signed_distance(const float4* edge, const int32_t* edge_id, int k, float qx, float qy, float* ux, float* uy) {
    const int kp = k > 0 ? k - 1 : 0;
    const float4 e = edge[k], ep = edge[kp];
    const float px = qx - e.x, py = qy - e.y; // points_to_edge
    const float cross = lcs_fmul(px, e.w) - lcs_fmul(py, e.z); // z of (px, py, 0) x (dx, dy, 0)
    const float cross_prior = lcs_fmul(px, ep.w) - lcs_fmul(py, ep.z);
    const bool same = edge_id[k] == edge_id[kp];
    const float c = (same && cross_prior < cross) ? cross_prior : cross;
    const float sign = c > 0.0f ? 1.0f : (c < 0.0f ? -1.0f : 0.0f);
    const float dist = sqrtf(lcs_fadd(lcs_fmul(px, px), lcs_fmul(py, py)));
    *ux = dist > 0.0f ? px / dist : 0.0f;
    *uy = dist > 0.0f ? py / dist : 0.0f;
    return dist * sign;
}

// metrics/geometry.py:8-90 in float32: corners, projections on `first`'s two axes, strict > 0
struct LcpBox {
    float cx[4], cy[4], c, s;
};
LCS_D void lcp_make_box(float x, float y, float length, float width, float yaw, LcpBox& b) {
    const float c = cosf(yaw), s = sinf(yaw);
    const float hl = length / 2.0f, hw = width / 2.0f;
    b.c = c; b.s = s;
    b.cx[0] = x + hl * c + hw * s; b.cx[1] = x - hl * c + hw * s; b.cx[2] = x - hl * c - hw * s; b.cx[3] = x + hl * c - hw * s;
    b.cy[0] = y + hl * s - hw * c; b.cy[1] = y - hl * s - hw * c; b.cy[2] = y - hl * s + hw * c; b.cy[3] = y + hl * s + hw * c;
}
LCS_D bool lcp_overlap_a_over_b(const LcpBox& f, const LcpBox& o) {
    for (int axis = 0; axis < 2; axis++) {
        // normals_t columns: (c, s) and (-s, c)
        const float n0 = axis == 0 ? f.c : -f.s, n1 = axis == 0 ? f.s : f.c;
        float min_a = lcs_inf_f(), max_a = -lcs_inf_f(), min_b = lcs_inf_f(), max_b = -lcs_inf_f();
        for (int k = 0; k < 4; k++) {
            const float pa = f.cx[k] * n0 + f.cy[k] * n1, pb = o.cx[k] * n0 + o.cy[k] * n1;
            min_a = fminf(min_a, pa); max_a = fmaxf(max_a, pa);
            min_b = fminf(min_b, pb); max_b = fmaxf(max_b, pb);
        }
        if (!(fminf(max_a, max_b) - fmaxf(min_a, min_b) > 0.0f)) return false;
    }
    return true;
}
LCS_D bool lcp_has_overlap(const LcpBox& a, const LcpBox& b) { return lcp_overlap_a_over_b(a, b) && lcp_overlap_a_over_b(b, a); }

#ifndef LCSIM_HOSTEMU
#define LCP_TILE 2048

// block reduction of (float, int) to thread 0
LCS_D void lcp_block_reduce(float& v, int& c, float* sh_f, int* sh_i) {
    for (int d = 16; d > 0; d >>= 1) {
        v += __shfl_xor_sync(0xffffffffu, v, d);
        c += __shfl_xor_sync(0xffffffffu, c, d);
    }
    const int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    if ((threadIdx.x & 31) == 0) { sh_f[w] = v; sh_i[w] = c; }
    __syncthreads();
    if (threadIdx.x == 0) {
        v = 0.f; c = 0;
        for (int k = 0; k < nw; k++) { v += sh_f[k]; c += sh_i[k]; }
    }
}

// Repeller: grid (T, S), one thread per agent; the step's positions live in shared memory
__global__ void __launch_bounds__(1024) lcp_repeller_kernel(const LcpArgs o) {
    extern __shared__ float lcp_sm[];
    __shared__ float sh_f[32];
    __shared__ int sh_i[32];
    const int t = blockIdx.x, s = blockIdx.y, i = threadIdx.x;
    const int n = o.n_agents[s];
    float* sx = lcp_sm;
    float* sy = lcp_sm + blockDim.x;
    const size_t q = (((size_t)s * o.A + i) * o.T + t) * 2;
    float xi = 0.f, yi = 0.f;
    if (i < n) { xi = o.xy[q]; yi = o.xy[q + 1]; }
    sx[i] = xi; sy[i] = yi;
    __syncthreads();
    float sum = 0.f, gx = 0.f, gy = 0.f;
    int cnt = 0;
    if (i < n) {
        const float inv_r = 1.0f / o.radius;
        for (int j = 0; j < n; j++)
            if (j != i) lcp_repel_pair(xi, yi, sx[j], sy[j], inv_r, sum, cnt, gx, gy);
        if (o.grad) { o.grad[q] = gx; o.grad[q + 1] = gy; } // unnormalised: scaled by lcp_finalize_kernel
    }
    lcp_block_reduce(sum, cnt, sh_f, sh_i);
    if (i == 0) {
        o.part_sum[(size_t)s * o.P + t] = sum;
        o.part_cnt[(size_t)s * o.P + t] = cnt;
    }
}

// NoOffRoad / off-road rate: grid (ceil(A T / 256), S), one thread per query (agent, step); the scenario's road-edge
// points stream through shared memory in tiles, every thread keeps the first argmin of the float32 squared distance
__global__ void __launch_bounds__(256) lcp_offroad_kernel(const LcpArgs o, int metric) {
    __shared__ float2 tile[LCP_TILE];
    __shared__ float sh_f[32];
    __shared__ int sh_i[32];
    const int s = blockIdx.y, qi = blockIdx.x * blockDim.x + threadIdx.x;
    const int a = qi / o.T, t = qi - a * o.T;
    const int n = o.n_agents[s], ne = o.n_edge[s];
    const float4* edge = o.edge + (size_t)s * o.E;
    bool active = a < o.A && a < n;
    if (active && o.mask) active = o.mask[(size_t)s * o.A + a] != 0;
    float qx = 0.f, qy = 0.f;
    if (active) {
        if (metric) {
            const float* r = o.traj5 + (((size_t)s * o.A + a) * o.T + t) * 5;
            qx = r[0]; qy = r[1];
        } else {
            const float* r = o.xy + (((size_t)s * o.A + a) * o.T + t) * 2;
            qx = r[0]; qy = r[1];
        }
    }
    float best = lcs_inf_f();
    int bi = -1;
    for (int t0 = 0; t0 < ne; t0 += LCP_TILE) {
        __syncthreads();
        for (int k = threadIdx.x; k < LCP_TILE && t0 + k < ne; k += blockDim.x) {
            const float4 e = edge[t0 + k];
            tile[k] = make_float2(e.x, e.y);
        }
        __syncthreads();
        if (!active) continue;
        const int m = min(LCP_TILE, ne - t0);
        for (int k = 0; k < m; k++) {
            const float dx = tile[k].x - qx, dy = tile[k].y - qy;
            const float d2 = lcs_fadd(lcs_fmul(dx, dx), lcs_fmul(dy, dy)); // torch.sum(differences**2, -1)
            if (d2 < best) { best = d2; bi = t0 + k; }
        }
    }
    float sum = 0.f;
    int cnt = 0;
    if (active && bi >= 0) {
        float ux, uy;
        const float sd = lcp_signed_distance(edge, o.edge_id + (size_t)s * o.E, bi, qx, qy, &ux, &uy);
        const size_t q = ((size_t)s * o.A + a) * o.T + t;
        if (o.sd) o.sd[q] = sd;
        if (metric) {
            cnt = sd > 0.0f ? 1 : 0;
        } else {
            const float v = o.radius + sd;
            const bool on = v > 0.0f;
            if (on) { sum = v; cnt = 1; }
            // d/dq relu(R + |q - n| sign) = sign (q - n) / |q - n|
            const float sign = sd > 0.0f ? 1.0f : (sd < 0.0f ? -1.0f : 0.0f);
            if (o.grad) { o.grad[2 * q] = on ? sign * ux : 0.0f; o.grad[2 * q + 1] = on ? sign * uy : 0.0f; }
        }
    } else if (qi < o.A * o.T) { // masked out / padding / no edge points: defined zeros
        const size_t q = ((size_t)s * o.A + a) * o.T + t;
        if (!metric && o.grad) { o.grad[2 * q] = 0.f; o.grad[2 * q + 1] = 0.f; }
        if (o.sd) o.sd[q] = 0.f;
    }
    lcp_block_reduce(sum, cnt, sh_f, sh_i);
    if (threadIdx.x == 0) {
        o.part_sum[(size_t)s * o.P + blockIdx.x] = sum;
        o.part_cnt[(size_t)s * o.P + blockIdx.x] = cnt;
    }
}

// overlap rate: grid (T, S), one thread per agent; boxes of the step in shared memory
__global__ void __launch_bounds__(1024) lcp_overlap_kernel(const LcpArgs o) {
    extern __shared__ float lcp_sm[];
    __shared__ float sh_f[32];
    __shared__ int sh_i[32];
    LcpBox* boxes = (LcpBox*)lcp_sm;
    float* rad = (float*)(boxes + blockDim.x);
    float* ctr = rad + blockDim.x; // [2][blockDim]
    const int t = blockIdx.x, s = blockIdx.y, i = threadIdx.x;
    const int n = o.n_agents[s];
    if (i < n) {
        const float* r = o.traj5 + (((size_t)s * o.A + i) * o.T + t) * 5;
        lcp_make_box(r[0], r[1], r[2], r[3], r[4], boxes[i]);
        rad[i] = 0.5f * sqrtf(r[2] * r[2] + r[3] * r[3]) * 1.001f + 1e-3f;
        ctr[i] = r[0];
        ctr[blockDim.x + i] = r[1];
    }
    __syncthreads();
    int cnt = 0;
    float dummy = 0.f;
    if (i < n && (!o.mask || o.mask[(size_t)s * o.A + i])) {
        bool any = false;
        const LcpBox me = boxes[i];
        for (int j = 0; j < n && !any; j++) {
            if (j == i) continue;
            const float dx = ctr[j] - ctr[i], dy = ctr[blockDim.x + j] - ctr[blockDim.x + i], rr = rad[i] + rad[j];
            if (dx * dx + dy * dy > rr * rr) continue; // bounding circles apart: the projections cannot overlap
            any = lcp_has_overlap(me, boxes[j]);
        }
        cnt = any ? 1 : 0;
    }
    lcp_block_reduce(dummy, cnt, sh_f, sh_i);
    if (i == 0) {
        o.part_sum[(size_t)s * o.P + t] = 0.f;
        o.part_cnt[(size_t)s * o.P + t] = cnt;
    }
}

// per scenario: partial sums -> loss / count / rate, gradients scaled in place.  mode 0: Repeller (grad *= -2 / (R N)),
// 1: NoOffRoad (grad *= 1 / N), 2: rate = count / (number of masked agents) / T
__global__ void __launch_bounds__(256) lcp_finalize_kernel(const LcpArgs o, int mode, int n_parts) {
    __shared__ float sh_f[32];
    __shared__ int sh_i[32];
    __shared__ float s_scale;
    const int s = blockIdx.x;
    float sum = 0.f;
    int cnt = 0;
    for (int k = threadIdx.x; k < n_parts; k += blockDim.x) {
        sum += o.part_sum[(size_t)s * o.P + k];
        cnt += o.part_cnt[(size_t)s * o.P + k];
    }
    lcp_block_reduce(sum, cnt, sh_f, sh_i);
    if (threadIdx.x == 0) {
        const float denom = (float)cnt + o.eps;
        if (o.count) o.count[s] = cnt;
        if (mode < 2) {
            if (o.loss) o.loss[s] = sum / denom;
            s_scale = mode == 0 ? -2.0f / (o.radius * denom) : 1.0f / denom;
        } else {
            int nm = 0;
            for (int a = 0; a < o.n_agents[s]; a++) {
                nm += (!o.mask || o.mask[(size_t)s * o.A + a]) ? 1 : 0;
            }
            if (o.rate) o.rate[s] = (float)cnt / (float)nm / (float)o.T; // 0 / 0 = nan like the reference
        }
    }
    __syncthreads();
    if (mode < 2 && o.grad) {
        const float sc = s_scale;
        float* g = o.grad + (size_t)s * o.A * o.T * 2;
        for (size_t k = threadIdx.x; k < (size_t)o.A * o.T * 2; k += blockDim.x) g[k] *= sc;
    }
}
#endif

// serial statements (host emulation; also the definition the kernels are tested against on the CPU tier)
#ifdef LCSIM_HOSTEMU
static void lcp_repeller_serial(const LcpArgs& o) {
    for (int s = 0; s < o.S; s++) {
        const int n = o.n_agents[s];
        double total = 0.0;
        long cnt_all = 0;
        for (int t = 0; t < o.T; t++)
            for (int i = 0; i < o.A; i++) {
                const size_t q = (((size_t)s * o.A + i) * o.T + t) * 2;
                float sum = 0.f, gx = 0.f, gy = 0.f;
                int cnt = 0;
                if (i < n)
                    for (int j = 0; j < n; j++)
                        if (j != i) {
                            const size_t qj = (((size_t)s * o.A + j) * o.T + t) * 2;
                            lcp_repel_pair(o.xy[q], o.xy[q + 1], o.xy[qj], o.xy[qj + 1], 1.0f / o.radius, sum, cnt, gx, gy);
                        }
                if (o.grad) { o.grad[q] = gx; o.grad[q + 1] = gy; }
                total += sum;
                cnt_all += cnt;
            }
        const float denom = (float)cnt_all + o.eps;
        if (o.count) o.count[s] = (int32_t)cnt_all;
        if (o.loss) o.loss[s] = (float)total / denom;
        if (o.grad)
            for (size_t k = 0; k < (size_t)o.A * o.T * 2; k++) o.grad[(size_t)s * o.A * o.T * 2 + k] *= -2.0f / (o.radius * denom);
    }
}
static void lcp_offroad_serial(const LcpArgs& o, int metric) {
    for (int s = 0; s < o.S; s++) {
        const int n = o.n_agents[s], ne = o.n_edge[s];
        const float4* edge = o.edge + (size_t)s * o.E;
        double total = 0.0;
        long cnt_all = 0, nm = 0;
        for (int a = 0; a < o.A; a++) {
            bool active = a < n && (!o.mask || o.mask[(size_t)s * o.A + a]);
                    nm += active ? 1 : 0;
            for (int t = 0; t < o.T; t++) {
                const size_t q = ((size_t)s * o.A + a) * o.T + t;
                if (!metric && o.grad) { o.grad[2 * q] = 0.f; o.grad[2 * q + 1] = 0.f; }
                if (o.sd) o.sd[q] = 0.f;
                if (!active || ne == 0) continue;
                const float qx = metric ? o.traj5[q * 5] : o.xy[q * 2], qy = metric ? o.traj5[q * 5 + 1] : o.xy[q * 2 + 1];
                float best = lcs_inf_f();
                int bi = -1;
                for (int k = 0; k < ne; k++) {
                    const float dx = edge[k].x - qx, dy = edge[k].y - qy;
                    const float d2 = lcs_fadd(lcs_fmul(dx, dx), lcs_fmul(dy, dy));
                    if (d2 < best) { best = d2; bi = k; }
                }
                float ux, uy;
                const float sd = lcp_signed_distance(edge, o.edge_id + (size_t)s * o.E, bi, qx, qy, &ux, &uy);
                if (o.sd) o.sd[q] = sd;
                if (metric) {
                    cnt_all += sd > 0.0f ? 1 : 0;
                } else {
                    const float v = o.radius + sd;
                    if (v > 0.0f) {
                        total += v;
                        cnt_all += 1;
                        const float sign = sd > 0.0f ? 1.0f : (sd < 0.0f ? -1.0f : 0.0f);
                        if (o.grad) { o.grad[2 * q] = sign * ux; o.grad[2 * q + 1] = sign * uy; }
                    }
                }
            }
        }
        if (o.count) o.count[s] = (int32_t)cnt_all;
        if (metric) {
            if (o.rate) o.rate[s] = (float)cnt_all / (float)nm / (float)o.T;
        } else {
            const float denom = (float)cnt_all + o.eps;
            if (o.loss) o.loss[s] = (float)total / denom;
            if (o.grad)
                for (size_t k = 0; k < (size_t)o.A * o.T * 2; k++) o.grad[(size_t)s * o.A * o.T * 2 + k] *= 1.0f / denom;
        }
    }
}
static void lcp_overlap_serial(const LcpArgs& o) {
    std::vector<LcpBox> boxes(o.A);
    for (int s = 0; s < o.S; s++) {
        const int n = o.n_agents[s];
        long cnt_all = 0, nm = 0;
        for (int a = 0; a < n; a++) nm += (!o.mask || o.mask[(size_t)s * o.A + a]) ? 1 : 0;
        for (int t = 0; t < o.T; t++) {
            for (int i = 0; i < n; i++) {
                const float* r = o.traj5 + (((size_t)s * o.A + i) * o.T + t) * 5;
                lcp_make_box(r[0], r[1], r[2], r[3], r[4], boxes[i]);
            }
            for (int i = 0; i < n; i++) {
                if (o.mask && !o.mask[(size_t)s * o.A + i]) continue;
                bool any = false;
                for (int j = 0; j < n && !any; j++)
                    if (j != i) any = lcp_has_overlap(boxes[i], boxes[j]);
                cnt_all += any ? 1 : 0;
            }
        }
        if (o.count) o.count[s] = (int32_t)cnt_all;
        if (o.rate) o.rate[s] = (float)cnt_all / (float)nm / (float)o.T;
    }
}
#endif
