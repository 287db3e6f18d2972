// lcsim_core.cuh — device code of the per-tick agent-stepping path (one CTA per scenario,
// one thread per agent).  Included by lcsim_b200.cu (nvcc, sm_100a).  The same text also
// compiles as plain C++ with -DLCSIM_HOSTEMU, where a "kernel launch" is a loop over
// (scenario, phase, thread); that build exists only under tests/hostemu to debug the
// control flow on machines without a GPU and is never loaded by the package.
//
// Numerics contract: every decision the reference takes in float64 is taken
// here in float64, operation for operation, with the roundings pinned by explicit
// __dmul_rn/__dadd_rn/__fma_rn (no compiler contraction).  float32 is used only as a
// conservative FILTER in front of those decisions: a float32 scan may discard a candidate
// only when a rigorous error bound proves it cannot win; survivors are re-evaluated in
// float64.  Integer outputs are therefore those of the float64 algorithm.
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef LCSIM_HOSTEMU
#include <string.h>
#define LCS_D static inline
#define LCS_HD static inline
struct lcs_f2 { float x, y; };
struct alignas(16) lcs_f4 { float x, y, z, w; };
LCS_D double lcs_dmul(double a, double b) { return a * b; } // built with -ffp-contract=off
LCS_D double lcs_dadd(double a, double b) { return a + b; }
LCS_D double lcs_dsub(double a, double b) { return a - b; }
LCS_D double lcs_dfma(double a, double b, double c) { return fma(a, b, c); }
#define LCS_DM static inline
#define LCS_ST_STREAM(ptr, v) (*(ptr) = (v))
#define LCS_LD_STREAM(ptr) (*(ptr))
LCS_D void lcs_sincos(double a, double* s, double* c) { *s = sin(a); *c = cos(a); }
LCS_D float lcs_fmul(float a, float b) { return a * b; }
LCS_D float lcs_fadd(float a, float b) { return a + b; }
LCS_D int lcs_popc(uint32_t x) { return __builtin_popcount(x); }
LCS_D int lcs_ffs0(uint32_t x) { return __builtin_ctz(x); } // index of the lowest set bit (x != 0)
LCS_D int lcs_atomic_add(int* p, int v) { int o = *p; *p = o + v; return o; }
LCS_D void lcs_atomic_min_u64(unsigned long long* p, unsigned long long v) { if (v < *p) *p = v; }
LCS_D unsigned long long lcs_d2u(double d) { unsigned long long u; memcpy(&u, &d, 8); return u; }
LCS_D int lcs_f2i(float f) { int i; memcpy(&i, &f, 4); return i; }
LCS_D float lcs_i2f(int i) { float f; memcpy(&f, &i, 4); return f; }
LCS_D void lcs_atomic_max(int* p, int v) { if (v > *p) *p = v; }
#else
#define LCS_D __device__ __forceinline__
#define LCS_HD __host__ __device__ __forceinline__
typedef float2 lcs_f2;
typedef float4 lcs_f4;
LCS_D double lcs_dmul(double a, double b) { return __dmul_rn(a, b); }
LCS_D double lcs_dadd(double a, double b) { return __dadd_rn(a, b); }
LCS_D double lcs_dsub(double a, double b) { return __dsub_rn(a, b); }
LCS_D double lcs_dfma(double a, double b, double c) { return __fma_rn(a, b, c); }
// streaming hints for data that is written and not read again by the rollout (history ring, float32 pose copy) or read
// exactly once (the snapshot a metric item consumes): evict-first in L2, so that the per-tick state stays resident
#ifdef LCS_CS_HINTS
#define LCS_ST_STREAM(ptr, v) __stcs((ptr), (v))
#define LCS_LD_STREAM(ptr) __ldcs(ptr)
#else
#define LCS_ST_STREAM(ptr, v) (*(ptr) = (v))
#define LCS_LD_STREAM(ptr) (*(ptr))
#endif
// LCS_NOINLINE_MATH (experiment build): the long float64 sequences (sincos, wrap_angle, sqrt) exist once in the kernel
// instead of at every call site (the engine kernel is 200 KB of code against a 32 KB instruction cache per SM)
#ifdef LCS_NOINLINE_MATH
#define LCS_DM __device__ __noinline__
#else
#define LCS_DM __device__ __forceinline__
#endif
LCS_DM void lcs_sincos(double a, double* s, double* c) { sincos(a, s, c); }
LCS_D float lcs_fmul(float a, float b) { return __fmul_rn(a, b); }
LCS_D float lcs_fadd(float a, float b) { return __fadd_rn(a, b); }
LCS_D int lcs_popc(uint32_t x) { return __popc(x); }
LCS_D int lcs_ffs0(uint32_t x) { return __ffs((int)x) - 1; }
LCS_D int lcs_atomic_add(int* p, int v) { return atomicAdd(p, v); }
LCS_D void lcs_atomic_min_u64(unsigned long long* p, unsigned long long v) { atomicMin(p, v); }
LCS_D unsigned long long lcs_d2u(double d) { return (unsigned long long)__double_as_longlong(d); }
LCS_D int lcs_f2i(float f) { return __float_as_int(f); }
LCS_D float lcs_i2f(int i) { return __int_as_float(i); }
LCS_D void lcs_atomic_max(int* p, int v) { atomicMax(p, v); }
#endif

// host emulation only: counters of the fast / fallback paths (tests read them to see how often a shortcut applies)
#ifdef LCSIM_HOSTEMU
static long long lcs_dbg_count[8]; // 0 window ok, 1 window rejected, 2 blob (no window), 3 lead decided, 4 lead fallback, 5 lead none, 6 cursor tie-break pass
#define LCS_COUNT(i) (lcs_dbg_count[i]++)
#else
#define LCS_COUNT(i) ((void)0)
#endif

#define LCS_SOA_Y(Ap) (2 * (Ap) + 4) // offset of the y array inside LcsSmem::soa
#define LCS_WAITING 0
#define LCS_ACTIVE 1
#define LCS_FINISHED 2
#define LCS_IDLE 3
#define LCS_ACT_NONE 0
#define LCS_ACT_BICYCLE 1
#define LCS_ACT_DELTA 2
#define LCS_ACT_STATE 3
#define LCS_H 11
#define LCS_ROUTE 10
#define LCS_NSTAT 8
#define LCS_PI 3.14159265358979323846
#define LCS_FAR 3.0e18f // filter coordinate of "no point here" (its square stays finite in float32)

LCS_D float lcs_inf_f() {
#ifdef LCSIM_HOSTEMU
    return INFINITY;
#else
    return __int_as_float(0x7f800000);
#endif
}
LCS_D double lcs_inf_d() {
#ifdef LCSIM_HOSTEMU
    return (double)INFINITY;
#else
    return __longlong_as_double(0x7ff0000000000000LL);
#endif
}

// Replay record of waypoint k of a pristine (float64) reference trajectory: everything the policy's own STATE
// action (StateExpert, expert_policy.py:18-33 -> agent.py:377-388) produces when it lands on that waypoint, so the
// log-replay update is one 64-byte load.  canon = first index j with sqrt((x_j-x_k)^2 + (y_j-y_k)^2) == 0, i.e. the
// value Trajectory.next (type.py:42-48) returns for a query that IS waypoint k: the distance to k is exactly 0, the
// global minimum, and np.argmin takes the first index that attains it.
struct alignas(64) LcsWpRec {
    double x, y, h;
    double v; // np.sqrt(vx**2 + vy**2)                (agent.py:383, unfused)
    double ch, sh; // cos / sin of h
    double vn; // np.linalg.norm(vel) = sqrt(ddot)      (idm_policy.py:200, BLAS order: vx*vx then fma)
    int32_t canon, pad;
};

// Everything a tick needs, as raw device pointers.  Layouts: DESIGN.md §3.
struct LcsParams {
    int32_t S, A, Ap, T, Tp, E; // Tp = T rounded up to 4 points (32-byte filter entries)
    double dt;
    int32_t total_steps, reactive, no_static, allow_new_obj, with_metrics;
    // ---- static, per scenario
    const int32_t *n_agents, *ads_index;
    const double* origin; // [S][2]
    const float* eta_pts; // [S] bound on |fl32(p - origin) - (p - origin)| over all static points
    // ---- static, per agent [S][Ap]
    const int32_t *depart_tick, *wait_order;
    const int32_t* depart_sorted; // [S][Ap] depart_tick along wait_order (one coalesced load in the departure test)
    const uint8_t* dynamic;
    const double *length, *width; // [S][Ap]
    const double* home_xy; // [S][Ap][2]
    // pristine trajectories (reset restores re-planned ones from here)
    const int32_t* traj_len0;
    const double *traj_xy0, *traj_vel0, *traj_h0; // [S][Ap][T][2], [S][Ap][T][2], [S][Ap][T]
    const LcsWpRec* wp_rec; // [S][Ap][T] replay records of the pristine trajectories (null: not built)
    const lcs_f2* tf_xy0; // [S][Ap][Tp] float32 filter copy, contiguous per agent, relative to the origin;
                          // exact duplicates and the padding past traj_len -> LCS_FAR
    // ---- current trajectories (set_ref_trajectory overwrites)
    int32_t* traj_len;
    double *traj_xy, *traj_vel, *traj_h;
    lcs_f2* tf_xy;
    // safe radius of every live filter point: half the distance to the nearest live point of the same trajectory more
    // than LCS_SCAN_W indices away (+inf if none), rounded down.  A query closer than that to point j has its nearest
    // point within LCS_SCAN_W indices of j (triangle inequality), so a scan of a window around the previous cursor is
    // enough (lcs_scan_window); traj_blob = 1 marks trajectories whose radii are too small to ever help.
    const float* tf_rho0; // [S][Ap][Tp]
    float* tf_rho;
    const uint8_t* traj_blob0; // [S][Ap]
    uint8_t* traj_blob;
    uint8_t* traj_f32; // 1: arrays came from a float32 re-plan (engine.py:290-297)
    uint8_t *stale_valid, *stale_f32; // obs.ref_trajectory still points at the old Trajectory
    double* stale_wp; // [S][Ap][5]
    // ---- road edges
    const int32_t* n_edge; // [S]
    const double *edge_xy, *edge_dir; // [S][E][2] original order (exact phase)
    // hash grid over the de-duplicated edge points of a scenario
    const float* grid_geom; // [S][4] gx0, gy0, inv_cell, cell
    const int32_t* grid_dim; // [S][2] W, H
    const int32_t* grid_start; // [S][GC] CSR offsets (W*H+1 used), cells row-major
    int32_t GC;
    const lcs_f2* ge_xy; // [S][E] points sorted by cell, float32 relative to origin
    const int32_t* ge_idx; // [S][E] original index of each sorted point
    // candidate lists: for every cell of a fine grid, ALL edge points that can be nearest to some position in the cell
    const float* cl_geom; // [S][4] x0, y0, 1 / cell, cell (float32 local frame)
    const int32_t* cl_dim; // [S][2] W, H (0, 0: no lists for this scenario)
    const int32_t* cl_start; // [S][cl_nc + 1] offsets into the scenario's run of cl_xy / cl_idx, multiples of 4
    const long long* cl_base; // [S] first entry of the scenario in cl_xy / cl_idx
    const lcs_f2* cl_xy; // filter coordinates, 32-byte blocks, padded with LCS_FAR
    const int32_t* cl_idx; // original index
    int32_t cl_nc;
    // ---- state, per scenario
    int32_t *tick, *n_active, *wait_ptr;
    int64_t* stats; // [S][LCS_NSTAT]
    // ---- state, per agent
    int32_t *active; // [S][Ap]
    int32_t *status, *cursor, *ring_count, *ring_head;
    double* pose64; // [S][Ap][4]
    lcs_f4* pose32; // [S][Ap]
    double* ring; // [S][Ap][H][5]
    uint8_t* lead_valid;
    double *lead_dis, *lead_v;
    int32_t *coll_count, *nearest_edge;
    int32_t* edge_hint; // [S][Ap] last known nearest edge point (reset: the one of the home pose)
    const int32_t* home_edge; // [S][Ap] nearest edge point of the home pose (computed at upload)
    uint8_t* offroad;
    // ---- external actions of this launch
    const int32_t *act_agent, *act_type; // [S][K]
    const double* act_data; // [S][K][5]
    int32_t K;
    const double* ads_action; // [S][2] normalised (acc, steer) of every scenario's ADS, or null (env_step)
    // ---- per-tick snapshot handed from the tick kernel to the metric kernels (ring of snap_depth ticks): collision
    // counts and nearest-edge / off-road are OUTPUTS that no later tick reads, so they are computed by their own
    // kernels, which overlap with the following ticks instead of lengthening every scenario's serial chain
    int32_t* snap_hdr; // [D][S][4] n_old, n_new, n_fin, -
    int32_t* snap_agent; // [D][S][Ap] agent of every new-list slot | (needs an edge search) << 30
    double* snap_rec; // [D][S][Ap][6] x, y, cos h, sin h, length, width of that agent
    int32_t* snap_fin; // [D][S][Ap] agents that left the list in this tick (dense)
    int32_t snap_depth; // D
    // ---- work queue of the engine kernel: tickets are claimed at run time, in an order in which every item depends
    // only on items with smaller tickets (tick k of scenario s on its tick k-1 and on the metric item of tick k-D;
    // the metric item on its tick item) -> a running CTA can only wait for work that is already running or done
    int32_t* q_ticket; // [1] next ticket
    int32_t* q_seq; // [S][2] tick items / metric items of this launch completed for the scenario
    int32_t n_ticks; // ticks of this launch
    // fused launch: every tick item is followed by the scenario's metric item in the same CTA (no second ticket, no
    // dependency wait; the CTA re-reads the snapshot it wrote).  Used for single-tick launches, where there is no chain
    // of ticks to keep short; LCSIM_B200_FUSE=1 / 0 forces it on / off for every launch.
    int32_t fuse;
    // ticket order of a multi-tick launch with metric items: 0 = T(k, .) then M(k, .); 1 = interleaved, M one tick behind:
    // T(k, 0), M(k-1, 0), T(k, 1), M(k-1, 1), ... (every dependency of an item lies two slot generations back)
    int32_t order;
    // scenarios per ticket group of a multi-tick launch (0 = one group): see lcs_engine_kernel
    int32_t group;
    // position in the ticket order -> scenario (heaviest scenarios first: their tick items are the long ones of a wave and
    // their chain is the critical path of the rollout), or null = identity.  Experiment, LCSIM_B200_PERM=1 (measured: no gain).
    const int32_t* perm;
    // single-tick launch split in two (env_step: the observation build overlaps the metric items): 1 = tick items only
    // (the snapshot is still published), 2 = metric items only (of the tick the previous launch ran); no tickets wait
    int32_t only;
    // ---- launch mode
    long long* dbg; // [S][16] phase clocks (instrumented instantiation of the tick kernel only)
    int32_t mode; // 0 = step, 1 = reset
    int32_t* tick_log; // [n_ticks][S][4] agents updated, active after, collision sum, off-road sum (or null)
    const uint8_t* reset_mask; // [S] or null: scenarios this launch touches (reset launch: which are reset; step: which are stepped)
};

// Shared memory of one CTA (Ap threads).  Carved from one dynamic allocation.
#define LCS_CAND 8 // private broadphase list per thread
#define LCS_QPER 8 // pair-queue capacity per thread (queue = LCS_QPER * Ap candidate pairs)
struct LcsSmem {
    double* ex; // [Ap][8]  x, y, v, cos h, sin h, L, W, h          (agent index)
    unsigned long long* lead_key; // [Ap] packed (distance, slot) minimum per ego   (slot index)
    lcs_f4* crec; // [Ap]  x32, y32, filter radius, coordinate slack   (slot index, new list)
    lcs_f4* crec2; // [Ap] cos h, sin h, half length, half width in float32 (slot index)
    int32_t* act; // [Ap] active list at tick start (slot index)
    int32_t* act2; // [Ap] active list after departures / arrivals
    int32_t* dlist; // [Ap] agents activated this tick, in waiting-list order
    int32_t* cnt; // [Ap] collision counts (slot index)
    uint32_t* queue; // [LCS_QPER * Ap] candidate pairs (slot_i << 16 | slot_j)
    uint16_t* cand; // [LCS_CAND][Ap] per-thread broadphase survivors
    uint32_t* mask; // [3][32] ballot words: finished, popped, activated
    int32_t* misc; // [16] n_old, n_new, collision sum, off-road sum, tick, wait_ptr, queue length, -, then [8..11]:
                   // number of set flags of the ballot rows finished / popped / activated / moved
    float* soa; // [2][2 Ap + 4] x and y of the list slots once more as float arrays, each written twice (slot i at i and
                // i + n) so that the partner range i+1 .. i+n/2 of the pair loops is contiguous: four partners per load
    uint8_t* moved; // [Ap] 0: the agent's position is bit-identical to the one its edge outputs were computed for (agent index)
};

#ifdef LCSIM_HOSTEMU
static inline
#else
__host__ __device__ inline
#endif
size_t lcs_smem_bytes(int Ap) {
    return (size_t)Ap * (64 + 8 + 32 + 4 * 4 + 4 * LCS_QPER + 2 * LCS_CAND + 16 + 1) + 32 + 3 * 32 * 4 + 16 * 4 + 64;
}

LCS_D void lcs_smem_carve(LcsSmem& sm, unsigned char* base, int Ap) {
    unsigned char* p = base;
    sm.ex = (double*)p; p += (size_t)Ap * 64;
    sm.lead_key = (unsigned long long*)p; p += (size_t)Ap * 8;
    sm.crec = (lcs_f4*)p; p += (size_t)Ap * 16;
    sm.crec2 = (lcs_f4*)p; p += (size_t)Ap * 16;
    sm.act = (int32_t*)p; p += (size_t)Ap * 4;
    sm.act2 = (int32_t*)p; p += (size_t)Ap * 4;
    sm.dlist = (int32_t*)p; p += (size_t)Ap * 4;
    sm.cnt = (int32_t*)p; p += (size_t)Ap * 4;
    sm.queue = (uint32_t*)p; p += (size_t)Ap * 4 * LCS_QPER;
    sm.mask = (uint32_t*)p; p += 3 * 32 * 4;
    sm.misc = (int32_t*)p; p += 16 * 4;
    sm.cand = (uint16_t*)p; p += (size_t)Ap * 2 * LCS_CAND;
    sm.soa = (float*)p; p += ((size_t)Ap * 4 + 8) * 4;
    sm.moved = (uint8_t*)p;
}

// ------------------------------------------------------------------ small float64 helpers

// reference: utils/polygon.py:191-192 wrap_angle = (a + pi) % (2*pi) - pi with the floored
// modulo of Python/NumPy.  fmod is exact in IEEE arithmetic on both sides.
LCS_DM double lcs_wrap_angle(double a) {
    const double two_pi = 2 * LCS_PI;
    const double t = lcs_dadd(a, LCS_PI);
    double r;
    // fmod is exact; for |t| < 4 pi it is t itself or ONE exact subtraction (Sterbenz: 2 pi <= t < 4 pi), and Python's
    // floored modulo then adds 2 pi (rounded) to a negative remainder.  Headings live in this range; the library
    // fmod (a long division loop on the device) is kept for everything else.
    if (t >= 0.0 && t < two_pi) {
        r = t;
    } else if (t >= two_pi && t < 2 * two_pi) {
        r = lcs_dsub(t, two_pi);
    } else if (t < 0.0 && t > -two_pi) {
        r = lcs_dadd(t, two_pi);
    } else {
        r = fmod(t, two_pi);
        if (r != 0.0) {
            if (r < 0.0) r = lcs_dadd(r, two_pi);
        } else {
            r = 0.0;
        }
    }
    return lcs_dsub(r, LCS_PI);
}
// two-term dot product the way BLAS evaluates np.dot / np.matmul / 1-D np.linalg.norm: a0*b0 then fma
LCS_D double lcs_dot2(double a0, double b0, double a1, double b1) { return lcs_dfma(a1, b1, lcs_dmul(a0, b0)); }
// np.linalg.norm over a trailing axis of length 2 (unfused)
LCS_DM double lcs_norm2_axis(double dx, double dy) { return sqrt(lcs_dadd(lcs_dmul(dx, dx), lcs_dmul(dy, dy))); }
// np.linalg.norm of a 1-D 2-vector (BLAS ddot)
LCS_D double lcs_norm2_vec(double dx, double dy) { return sqrt(lcs_dot2(dx, dx, dy, dy)); }
// x ** 4 (Python float power = C pow(x, 4.0), glibc: below 1 ulp, correctly rounded but for rare cases): x^2 and its
// square as exact double-double products, one final rounding — tighter than (x*x)*(x*x) (three roundings) and than the
// CUDA pow (2 ulp), and a handful of instructions instead of ~200
LCS_D double lcs_pow4(double x) {
    const double h = lcs_dmul(x, x), l = lcs_dfma(x, x, -h); // x^2 = h + l exactly
    const double h2 = lcs_dmul(h, h), e = lcs_dfma(h, h, -h2); // h^2 = h2 + e exactly
    return lcs_dadd(h2, lcs_dadd(e, lcs_dmul(lcs_dmul(2.0, h), l))); // + 2 h l (l^2 lies below 2^-104 of the result)
}
LCS_D double lcs_pymax(double a, double b) { return b > a ? b : a; }
LCS_D double lcs_pymin(double a, double b) { return b < a ? b : a; }
LCS_D double lcs_npclip(double x, double lo, double hi) {
    if (x != x) return x;
    double m = x < lo ? lo : x;
    return m > hi ? hi : m;
}

// float32 candidate threshold: every point whose float64 distance can equal the minimum has a
// filter value <= thr when the best filter value is m.  eta bounds the error of
// the float32 coordinates (points + query).
// eta: static points carry at most eta_pts (measured at upload); the query and any re-planned
// (float32-sourced) point near it carry at most half an ulp of their local coordinate.
LCS_D float lcs_filter_eta(float eta_pts, float fqx, float fqy, float m) {
    return eta_pts + 1.2e-7f * (fabsf(fqx) + fabsf(fqy)) + 1.2e-7f * (sqrtf(m) + 1.0f);
}
LCS_D float lcs_filter_thr(float m, float eta) {
    float r = sqrtf(m) * 1.000001f + 2.1f * eta + 1e-7f;
    r *= 1.000001f;
    return r * r * 1.000001f;
}

// ------------------------------------------------------------------ ballot-word scans

// flag of "thread" tid -> bit tid of words[]
LCS_D void lcs_set_flag(uint32_t* words, int tid, bool f) {
#ifdef LCSIM_HOSTEMU
    if (f) words[tid >> 5] |= 1u << (tid & 31);
#else
    uint32_t b = __ballot_sync(0xffffffffu, f);
    if ((tid & 31) == 0) words[tid >> 5] = b;
#endif
}
// same, and the number of set flags of the whole row accumulates in *total (zeroed at tick start): every thread
// needs the totals afterwards, and popcounting all words again in every thread was 9 % of the kernel's instructions
LCS_D void lcs_set_flag_count(uint32_t* words, int tid, bool f, int32_t* total) {
#ifdef LCSIM_HOSTEMU
    if (f) {
        words[tid >> 5] |= 1u << (tid & 31);
        *total += 1;
    }
#else
    uint32_t b = __ballot_sync(0xffffffffu, f);
    if ((tid & 31) == 0) {
        words[tid >> 5] = b;
        if (b) atomicAdd(total, __popc(b));
    }
#endif
}
LCS_D int lcs_prefix(const uint32_t* words, int tid) { // number of set flags below tid
    int w = tid >> 5, r = lcs_popc(words[w] & ((1u << (tid & 31)) - 1u));
    for (int k = 0; k < w; k++) r += lcs_popc(words[k]);
    return r;
}
// index of the k-th set flag (k = 0 .. total-1)
LCS_D int lcs_select(const uint32_t* words, int nwords, int k) {
    int w = 0;
    for (; w < nwords - 1; w++) {
        const int c = lcs_popc(words[w]);
        if (k < c) break;
        k -= c;
    }
#ifdef LCSIM_HOSTEMU
    uint32_t x = words[w];
    for (int b = 0; b < 32; b++)
        if ((x >> b) & 1u) {
            if (k == 0) return 32 * w + b;
            k--;
        }
    return 32 * w + 31;
#else
    return 32 * w + (int)__fns(words[w], 0, k + 1);
#endif
}
LCS_D int lcs_total(const uint32_t* words, int nwords) {
    int r = 0;
    for (int k = 0; k < nwords; k++) r += lcs_popc(words[k]);
    return r;
}

// ------------------------------------------------------------------ cursor: Trajectory.next

// reference: agent/policy/type.py:42-48 Trajectory.next -> utils/polygon.py:271-293
// frenet_projection: index = first argmin_k sqrt((x_k-px)^2 + (y_k-py)^2) over ALL len points.
// Filter scan over the float32 copy (coalesced: [T][Ap]), exact float64 only on near-ties.
// float2 index of filter point k of agent a in scenario s
LCS_HD size_t lcs_tf_index(const LcsParams& p, int s, int k, int a) { return ((size_t)s * p.Ap + a) * p.Tp + k; }

// 32-byte load (4 filter points).  Thread-per-agent scans are uncoalesced by nature; one 32-byte sector per
// lane per request keeps the L1 line-lookup rate (one 128-byte line per cycle) from becoming the limit.
struct alignas(32) lcs_f8 { float v[8]; };
LCS_D lcs_f8 lcs_ld32(const lcs_f8* q) {
#ifdef LCSIM_HOSTEMU
    return *q;
#else
    lcs_f8 r;
    // not volatile: the scheduler must be free to issue several of these before the first use
    asm("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7])
                 : "l"(q));
    return r;
#endif
}

// The same load for data that is read once per search and never again soon (candidate lists: 3.5 GB of them at
// S=1024): L2 evict-first, so that these lines do not push the per-tick state (13 MB, re-read every tick) out of L2.
// Measured at S=1024 against a build with -DLCS_NO_STREAM_EF: log-replay 2.97e9 vs 2.91e9, reactive 1.137e9 vs 1.107e9.
LCS_D lcs_f8 lcs_ld32_stream(const lcs_f8* q) {
#if defined(LCSIM_HOSTEMU) || defined(LCS_NO_STREAM_EF)
    return lcs_ld32(q);
#else
    lcs_f8 r;
    asm("ld.global.L2::evict_first.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7])
                 : "l"(q));
    return r;
#endif
}

#define LCS_SCAN_POINT(PX, PY, K)            \
    {                                        \
        const float dx = (PX)-fqx, dy = (PY)-fqy; \
        const float d = dx * dx + dy * dy;   \
        const bool c = d < m;                \
        m2 = c ? m : fminf(m2, d);           \
        j = c ? (K) : j;                     \
        m = c ? d : m;                       \
    }

// Running state of one filter scan: best and second-best float32 squared distance, index of the best.
struct LcsScan {
    float fqx, fqy, m, m2;
    int j;
};
LCS_D void lcs_scan_begin(const LcsParams& p, int s, double qx, double qy, LcsScan& sc) {
    sc.fqx = (float)(qx - p.origin[2 * s]);
    sc.fqy = (float)(qy - p.origin[2 * s + 1]);
    sc.m = sc.m2 = lcs_inf_f();
    sc.j = 0;
}
// one 32-byte filter entry = points k .. k+3
LCS_D void lcs_scan_quad(LcsScan& sc, const lcs_f8 v, int k) {
    const float fqx = sc.fqx, fqy = sc.fqy;
    float m = sc.m, m2 = sc.m2;
    int j = sc.j;
    LCS_SCAN_POINT(v.v[0], v.v[1], k)
    LCS_SCAN_POINT(v.v[2], v.v[3], k + 1)
    LCS_SCAN_POINT(v.v[4], v.v[5], k + 2)
    LCS_SCAN_POINT(v.v[6], v.v[7], k + 3)
    sc.m = m; sc.m2 = m2; sc.j = j;
}
// the agent's own contiguous filter row; only ceil(len/4) sectors are touched (entries past len hold LCS_FAR)
LCS_D void lcs_scan_global(const LcsParams& p, int s, int a, int len, LcsScan& sc) {
    const lcs_f8* tf = (const lcs_f8*)(p.tf_xy + lcs_tf_index(p, s, 0, a));
    const int n4 = (len + 3) >> 2;
    for (int k4 = 0; k4 < n4; k4 += 4) { // four 32-byte loads in flight, then their 16 points
        lcs_f8 v[4];
#pragma unroll
        for (int r = 0; r < 4; r++)
            if (k4 + r < n4) v[r] = lcs_ld32(tf + k4 + r);
#pragma unroll
        for (int r = 0; r < 4; r++)
            if (k4 + r < n4) lcs_scan_quad(sc, v[r], 4 * (k4 + r));
    }
}
// Windowed scan: 16 points around the previous cursor (4 sectors in flight), accepted when the safe radius of the
// window's best point proves that no point outside the window can be nearer; otherwise the caller
// scans the whole row.  Agents move a few waypoints per tick at most, so this replaces ~23 sector loads by 5.
#define LCS_SCAN_W 4
LCS_D bool lcs_scan_window(const LcsParams& p, int s, int a, int len, int cursor, LcsScan& sc) {
    const size_t row = lcs_tf_index(p, s, 0, a);
    const int q0 = (cursor >> 2) - 1, w0 = q0 > 0 ? 4 * q0 : 0, n4 = (len + 3) >> 2;
    const lcs_f8* tf = (const lcs_f8*)(p.tf_xy + row) + (w0 >> 2);
    lcs_f8 v[4];
#pragma unroll
    for (int r = 0; r < 4; r++)
        if ((w0 >> 2) + r < n4) v[r] = lcs_ld32(tf + r);
#pragma unroll
    for (int r = 0; r < 4; r++)
        if ((w0 >> 2) + r < n4) lcs_scan_quad(sc, v[r], w0 + 4 * r);
    const int j = sc.j;
    if (!(sc.m < 1.0e30f)) return false; // no live point in the window
    if (!((j - LCS_SCAN_W >= w0 || w0 == 0) && (j + LCS_SCAN_W <= w0 + 15 || w0 + 16 >= len))) return false;
    const float rho = p.tf_rho[row + j];
    const float eta = lcs_filter_eta(p.eta_pts[s], sc.fqx, sc.fqy, sc.m);
    return sqrtf(sc.m) * 1.00001f + 1.5f * eta + 1e-6f < rho;
}
// merge of two lanes' filter results (best, second best, index of best)
LCS_D void lcs_edge_merge(float& m, float& m2, int& j, float om, float om2, int oj) {
    if (om < m) {
        m2 = fminf(m, om2);
        m = om;
        j = oj;
    } else {
        if (om == m && oj < j) j = oj;
        m2 = fminf(m2, om); // om == m counts as a tie: m2 == m sends the query to the exact phase
    }
}

// first stage of the cursor scan: true when the window settled it, false when the whole row has to be scanned
LCS_D bool lcs_scan_first(const LcsParams& p, int s, int a, int len, int cursor, double qx, double qy, LcsScan& sc) {
    lcs_scan_begin(p, s, qx, qy, sc);
    if (p.traj_blob[(size_t)s * p.Ap + a]) {
        LCS_COUNT(2);
        return false;
    }
    if (lcs_scan_window(p, s, a, len, cursor, sc)) {
        LCS_COUNT(0);
        return true;
    }
    LCS_COUNT(1);
    lcs_scan_begin(p, s, qx, qy, sc);
    return false;
}
#ifndef LCSIM_HOSTEMU
// Second stage on the device: the lanes of a warp that need a whole-row scan are served four at a time, each row by
// a group of 8 lanes (lane i of the group takes the 32-byte entries i, i + 8, i + 16, ... : the group's loads are
// contiguous 256-byte pieces, a quarter of the L1 wavefronts of private row walks, and three loads per lane are in
// flight); a butterfly inside the group merges best / second best / index (ties keep the lower index and force the
// exact pass, like the serial scan).  Every lane of the warp must call this.
LCS_D void lcs_scan_rows_warp(const LcsParams& p, int s, bool need, int a, int len, LcsScan& sc) {
    unsigned todo = __ballot_sync(0xffffffffu, need);
    const int lane = threadIdx.x & 31, grp = lane >> 3, sub = lane & 7;
    while (todo) {
        // the next (up to) four flagged lanes; group g serves the g-th of them
        int src = -1, mine = -1;
#pragma unroll
        for (int g = 0; g < 4; g++) {
            const int l = todo ? __ffs(todo) - 1 : -1;
            if (todo) todo &= todo - 1;
            if (g == grp) src = l;
            if (l == lane) mine = g;
        }
        const int srcl = src < 0 ? 0 : src;
        LcsScan t;
        t.fqx = __shfl_sync(0xffffffffu, sc.fqx, srcl);
        t.fqy = __shfl_sync(0xffffffffu, sc.fqy, srcl);
        t.m = t.m2 = lcs_inf_f();
        t.j = 0x7fffffff;
        const int ar = __shfl_sync(0xffffffffu, a, srcl);
        const int len_r = __shfl_sync(0xffffffffu, len, srcl); // every lane takes part in every shuffle
        const int n4 = src < 0 ? 0 : (len_r + 3) >> 2;
        const lcs_f8* row = (const lcs_f8*)(p.tf_xy + lcs_tf_index(p, s, 0, ar < 0 ? 0 : ar));
        for (int k4 = sub; k4 < n4; k4 += 24) {
            lcs_f8 v[3];
#pragma unroll
            for (int r = 0; r < 3; r++)
                if (k4 + 8 * r < n4) v[r] = lcs_ld32(row + k4 + 8 * r);
#pragma unroll
            for (int r = 0; r < 3; r++)
                if (k4 + 8 * r < n4) lcs_scan_quad(t, v[r], 4 * (k4 + 8 * r));
        }
#pragma unroll
        for (int d = 4; d > 0; d >>= 1) {
            const float om = __shfl_xor_sync(0xffffffffu, t.m, d), om2 = __shfl_xor_sync(0xffffffffu, t.m2, d);
            const int oj = __shfl_xor_sync(0xffffffffu, t.j, d);
            lcs_edge_merge(t.m, t.m2, t.j, om, om2, oj);
        }
        // hand the group's result to the lane it belongs to
        const int from = mine < 0 ? 0 : 8 * mine;
        const float rm = __shfl_sync(0xffffffffu, t.m, from), rm2 = __shfl_sync(0xffffffffu, t.m2, from);
        const int rj = __shfl_sync(0xffffffffu, t.j, from);
        if (mine >= 0) { sc.m = rm; sc.m2 = rm2; sc.j = rj; }
    }
}
#endif
// the cursor from a finished scan: unique candidate -> done; near-tie -> float64 on every candidate,
// first minimum wins (np.argmin)
LCS_D int lcs_scan_finish(const LcsParams& p, int s, int a, int len, double qx, double qy, const LcsScan& sc) {
    const float fqx = sc.fqx, fqy = sc.fqy;
    const float eta = lcs_filter_eta(p.eta_pts[s], fqx, fqy, sc.m);
    const float thr = lcs_filter_thr(sc.m, eta);
    if (sc.m2 > thr) return sc.j;
    LCS_COUNT(6);
    const double* t64 = p.traj_xy + ((size_t)s * p.Ap + a) * p.T * 2;
    const lcs_f8* tf = (const lcs_f8*)(p.tf_xy + lcs_tf_index(p, s, 0, a));
    double bd = lcs_inf_d();
    int best = sc.j;
    // near-tie (common for parked agents whose logged points jitter inside a few centimetres): the row once more, four
    // points per load (it is in L1 from the scan), float64 on the points inside the band, first minimum wins
    const int n4 = (len + 3) >> 2;
    for (int k4 = 0; k4 < n4; k4++) {
        const lcs_f8 v = lcs_ld32(tf + k4);
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const int k = 4 * k4 + r;
            const float dx = v.v[2 * r] - fqx, dy = v.v[2 * r + 1] - fqy;
            const float d = dx * dx + dy * dy;
            if (d <= thr && k < len) {
                const double e = lcs_norm2_axis(lcs_dsub(t64[2 * k], qx), lcs_dsub(t64[2 * k + 1], qy));
                if (e < bd) { bd = e; best = k; }
            }
        }
    }
    return best;
}

// ------------------------------------------------------------------ nearest road-edge point

// reference: junction/junction.py:365-385 check_offroad: k* = first argmin over the concatenated,
// resampled edge points; offroad = cross(pts[k*] - pos, dir[k*]) < 0.
// The points live in a uniform hash grid (cells row-major, CSR); a (2r+1)^2 window around the
// query is scanned row by row (one contiguous span per row), r grows until the window provably
// contains every point that can be the float64 argmin.
// The search is written per LANE so that one thread (nlanes = 1: rewards kernel) or a whole warp
// (nlanes = 32: tick kernel, coalesced loads + warp-shuffle argmin) can run it.
struct LcsEdgeQ {
    float fqx, fqy, gx0, gy0, inv, eta0;
    int W, H, x0, x1, y0, y1;
    const int32_t* cs;
    const lcs_f2* gp;
};
LCS_D void lcs_edge_begin(const LcsParams& p, int s, double qx, double qy, LcsEdgeQ& q) {
    q.fqx = (float)(qx - p.origin[2 * s]);
    q.fqy = (float)(qy - p.origin[2 * s + 1]);
    q.gx0 = p.grid_geom[4 * s];
    q.gy0 = p.grid_geom[4 * s + 1];
    q.inv = p.grid_geom[4 * s + 2];
    q.eta0 = p.eta_pts[s];
    q.W = p.grid_dim[2 * s];
    q.H = p.grid_dim[2 * s + 1];
    q.cs = p.grid_start + (size_t)s * p.GC;
    q.gp = p.ge_xy + (size_t)s * p.E;
    // cell of the query; clamped far outside the grid so the integer math below cannot overflow
    float fcx = floorf((q.fqx - q.gx0) * q.inv), fcy = floorf((q.fqy - q.gy0) * q.inv);
    fcx = fminf(fmaxf(fcx, -1.0e6f), 1.0e6f);
    fcy = fminf(fmaxf(fcy, -1.0e6f), 1.0e6f);
    const int cx = (int)fcx, cy = (int)fcy;
    q.x0 = cx - 1; q.x1 = cx + 1; q.y0 = cy - 1; q.y1 = cy + 1;
}
// this lane's share of the window: points lane, lane + nlanes, ... of every row span
LCS_D void lcs_edge_lane_scan(const LcsEdgeQ& q, int lane, int nlanes, float& m, float& m2, int& j) {
    const float fqx = q.fqx, fqy = q.fqy;
    m = lcs_inf_f();
    m2 = lcs_inf_f();
    j = -1;
    const int wx0 = q.x0 < 0 ? 0 : q.x0, wx1 = q.x1 > q.W - 1 ? q.W - 1 : q.x1;
    const int wy0 = q.y0 < 0 ? 0 : q.y0, wy1 = q.y1 > q.H - 1 ? q.H - 1 : q.y1;
    if (wx0 > wx1) return;
    for (int y = wy0; y <= wy1; y++) {
        const int b = q.cs[y * q.W + wx0], e = q.cs[y * q.W + wx1 + 1];
        for (int k = b + lane; k < e; k += nlanes) {
            const lcs_f2 pt = q.gp[k];
            LCS_SCAN_POINT(pt.x, pt.y, k)
        }
    }
}
// after the lanes' results were merged into (M, J): 1 = window final (thr valid), 0 = window grown, scan again,
// -1 = the scenario has no edge point at all
LCS_D int lcs_edge_step(LcsEdgeQ& q, float M, int J, float& thr) {
    const bool covers_all = q.x0 <= 0 && q.y0 <= 0 && q.x1 >= q.W - 1 && q.y1 >= q.H - 1;
    if (J >= 0) {
        thr = lcs_filter_thr(M, lcs_filter_eta(q.eta0, q.fqx, q.fqy, M));
        // every candidate lies within D of the query in each axis; the extra 1e-3 m absorbs the float32
        // rounding of the cell arithmetic (grid points were binned with the same formula)
        const float D = sqrtf(thr) * 1.000001f + 1e-3f + 4e-7f * (fabsf(q.fqx) + fabsf(q.fqy) + fabsf(q.gx0) + fabsf(q.gy0));
        const int nx0 = (int)fmaxf(floorf((q.fqx - D - q.gx0) * q.inv), -1.0e6f);
        const int nx1 = (int)fminf(floorf((q.fqx + D - q.gx0) * q.inv), 1.0e6f);
        const int ny0 = (int)fmaxf(floorf((q.fqy - D - q.gy0) * q.inv), -1.0e6f);
        const int ny1 = (int)fminf(floorf((q.fqy + D - q.gy0) * q.inv), 1.0e6f);
        if (covers_all || (nx0 >= q.x0 && nx1 <= q.x1 && ny0 >= q.y0 && ny1 <= q.y1)) return 1;
        q.x0 = nx0 < q.x0 ? nx0 : q.x0;
        q.x1 = nx1 > q.x1 ? nx1 : q.x1;
        q.y0 = ny0 < q.y0 ? ny0 : q.y0;
        q.y1 = ny1 > q.y1 ? ny1 : q.y1;
        return 0;
    }
    if (covers_all) return -1;
    const int r = q.x1 - q.x0; // grow geometrically while the window is empty
    q.x0 -= r; q.x1 += r; q.y0 -= r; q.y1 += r;
    return 0;
}
// near-tie: float64 evaluation of this lane's candidates; (distance, original index) lexicographic minimum
LCS_D void lcs_edge_lane_exact(const LcsParams& p, int s, const LcsEdgeQ& q, double qx, double qy, float thr, int lane,
                               int nlanes, double& bd, int& bi) {
    const double* e64 = p.edge_xy + (size_t)s * p.E * 2;
    const int32_t* gi = p.ge_idx + (size_t)s * p.E;
    bd = lcs_inf_d();
    bi = 0x7fffffff;
    const int wx0 = q.x0 < 0 ? 0 : q.x0, wx1 = q.x1 > q.W - 1 ? q.W - 1 : q.x1;
    const int wy0 = q.y0 < 0 ? 0 : q.y0, wy1 = q.y1 > q.H - 1 ? q.H - 1 : q.y1;
    for (int y = wy0; y <= wy1; y++) {
        const int b = q.cs[y * q.W + wx0], e = q.cs[y * q.W + wx1 + 1];
        for (int k = b + lane; k < e; k += nlanes) {
            const lcs_f2 pt = q.gp[k];
            const float dx = pt.x - q.fqx, dy = pt.y - q.fqy;
            if (dx * dx + dy * dy <= thr) {
                const int idx = gi[k];
                const double ee = lcs_norm2_axis(lcs_dsub(e64[2 * idx], qx), lcs_dsub(e64[2 * idx + 1], qy));
                if (ee < bd || (ee == bd && idx < bi)) { bd = ee; bi = idx; }
            }
        }
    }
}
// off-road side test at the nearest point (junction.py:383-385)
LCS_D bool lcs_edge_side(const LcsParams& p, int s, double qx, double qy, int best) {
    const double* e64 = p.edge_xy + (size_t)s * p.E * 2;
    const double* d64 = p.edge_dir + (size_t)s * p.E * 2;
    const double ax = lcs_dsub(e64[2 * best], qx), ay = lcs_dsub(e64[2 * best + 1], qy);
    const double cross = lcs_dsub(lcs_dmul(ax, d64[2 * best + 1]), lcs_dmul(ay, d64[2 * best])); // np.cross, 2-vectors
    return cross < 0;
}
// single-thread driver (rewards kernel)
LCS_D int lcs_nearest_edge(const LcsParams& p, int s, double qx, double qy, bool* offroad) {
    *offroad = false;
    if (p.n_edge[s] == 0) return -1;
    LcsEdgeQ q;
    lcs_edge_begin(p, s, qx, qy, q);
    float m, m2, thr = 0.f;
    int j;
    for (;;) {
        lcs_edge_lane_scan(q, 0, 1, m, m2, j);
        const int st = lcs_edge_step(q, m, j, thr);
        if (st < 0) return -1;
        if (st > 0) break;
    }
    int best = p.ge_idx[(size_t)s * p.E + j];
    if (m2 <= thr) {
        double bd;
        lcs_edge_lane_exact(p, s, q, qx, qy, thr, 0, 1, bd, best);
    }
    *offroad = lcs_edge_side(p, s, qx, qy, best);
    return best;
}

// Nearest edge point through the candidate lists.  List K(C) of cell C holds every edge point p with
// mindist(p, C) <= U(C) + 1 cm, U(C) = min over all points of maxdist(point, C): for a query anywhere in C the nearest
// point is no farther than U(C), so K(C) contains every point that can attain (or tie within the filter's error band)
// the minimum.  The search is one header load and a contiguous run of 32-byte blocks, four in flight; the float32
// filter / float64 tie-break logic is the one of the other searches.  Returns -2 when the query lies outside the grid.
// begin: the cell's span in the scenario's list run (b, e); status -2 = no list for this query
struct LcsCellQ {
    float fqx, fqy;
    int b, e; // e < 0: not available
};
LCS_D void lcs_edge_cell_begin(const LcsParams& p, int s, double qx, double qy, LcsCellQ& q) {
    q.b = 0;
    q.e = -1;
    const int W = p.cl_dim[2 * s], H = p.cl_dim[2 * s + 1];
    q.fqx = (float)(qx - p.origin[2 * s]);
    q.fqy = (float)(qy - p.origin[2 * s + 1]);
    if (W <= 0) return;
    const float gx0 = p.cl_geom[4 * s], gy0 = p.cl_geom[4 * s + 1], inv = p.cl_geom[4 * s + 2];
    const float fx = floorf((q.fqx - gx0) * inv), fy = floorf((q.fqy - gy0) * inv);
    if (!(fx >= 0.0f && fy >= 0.0f && fx < (float)W && fy < (float)H)) return;
    const int cell = (int)fy * W + (int)fx;
    const int32_t* cs = p.cl_start + (size_t)s * (p.cl_nc + 1);
    q.b = cs[cell];
    q.e = cs[cell + 1];
}
LCS_D int lcs_edge_cell_finish(const LcsParams& p, int s, double qx, double qy, const LcsCellQ& q, bool* offroad) {
    const float fqx = q.fqx, fqy = q.fqy;
    const int b = q.b, e = q.e;
    if (e <= b) return -2; // outside the grid, or no edge point at all in this scenario
    const size_t base = (size_t)p.cl_base[s];
    const lcs_f2* gp = p.cl_xy + base;
    float m = lcs_inf_f(), m2 = lcs_inf_f();
    int j = -1;
    for (int k0 = b; k0 < e; k0 += 16) {
        lcs_f8 v[4];
#pragma unroll
        for (int r = 0; r < 4; r++)
            if (k0 + 4 * r < e) v[r] = lcs_ld32_stream((const lcs_f8*)(gp + k0 + 4 * r));
#pragma unroll
        for (int r = 0; r < 4; r++)
            if (k0 + 4 * r < e) {
                const int k = k0 + 4 * r;
                LCS_SCAN_POINT(v[r].v[0], v[r].v[1], k)
                LCS_SCAN_POINT(v[r].v[2], v[r].v[3], k + 1)
                LCS_SCAN_POINT(v[r].v[4], v[r].v[5], k + 2)
                LCS_SCAN_POINT(v[r].v[6], v[r].v[7], k + 3)
            }
    }
    const int32_t* gi = p.cl_idx + base;
    const float thr = lcs_filter_thr(m, lcs_filter_eta(p.eta_pts[s], fqx, fqy, m));
    int best = gi[j];
    if (m2 <= thr) { // near-tie: float64 on every candidate, (distance, original index) minimum
        const double* e64 = p.edge_xy + (size_t)s * p.E * 2;
        double bd = lcs_inf_d();
        best = 0x7fffffff;
        for (int k = b; k < e; k++) {
            const lcs_f2 pt = gp[k];
            const float dx = pt.x - fqx, dy = pt.y - fqy;
            if (dx * dx + dy * dy <= thr) {
                const int idx = gi[k];
                const double ee = lcs_norm2_axis(lcs_dsub(e64[2 * idx], qx), lcs_dsub(e64[2 * idx + 1], qy));
                if (ee < bd || (ee == bd && idx < best)) { bd = ee; best = idx; }
            }
        }
    }
    *offroad = lcs_edge_side(p, s, qx, qy, best);
    return best;
}

LCS_D int lcs_nearest_edge_cell(const LcsParams& p, int s, double qx, double qy, bool* offroad) {
    LcsCellQ q;
    lcs_edge_cell_begin(p, s, qx, qy, q);
    return lcs_edge_cell_finish(p, s, qx, qy, q, offroad);
}

// Nearest edge point with a HINT: `hint` is the original index of some edge point (the agent's nearest point of
// the previous tick, or of its home pose).  Its distance bounds the answer, so only the grid rows / chords that
// intersect the disk of that radius are scanned, in 32-byte blocks (4 points per load), in a single pass.
LCS_D int lcs_nearest_edge_hint(const LcsParams& p, int s, double qx, double qy, int hint, bool* offroad) {
    if (hint < 0) return lcs_nearest_edge(p, s, qx, qy, offroad);
    *offroad = false;
    LcsEdgeQ q;
    lcs_edge_begin(p, s, qx, qy, q);
    const double* e64 = p.edge_xy + (size_t)s * p.E * 2;
    const float fqx = q.fqx, fqy = q.fqy;
    // upper bound of the hint's filter value, then of the candidate radius D0 (monotone in the filter minimum)
    const float hx = (float)(e64[2 * hint] - qx), hy = (float)(e64[2 * hint + 1] - qy);
    const float eta = lcs_filter_eta(q.eta0, fqx, fqy, hx * hx + hy * hy);
    const float b0 = sqrtf(hx * hx + hy * hy) * 1.000001f + 2.1f * eta + 1e-6f;
    const float thr0 = lcs_filter_thr(b0 * b0 * 1.000001f, eta);
    const float slack = 1e-3f + 4e-7f * (fabsf(fqx) + fabsf(fqy) + fabsf(q.gx0) + fabsf(q.gy0));
    const float D0 = sqrtf(thr0) * 1.000001f + slack;
    const float cell = p.grid_geom[4 * s + 3];
    int y0 = (int)fmaxf(floorf((fqy - D0 - q.gy0) * q.inv), 0.0f);
    int y1 = (int)fminf(floorf((fqy + D0 - q.gy0) * q.inv), (float)(q.H - 1));
    float m = lcs_inf_f(), m2 = lcs_inf_f();
    int j = -1;
    // rows are handled four at a time: first every row's span lookup (8 independent loads), then block i of
    // every row together (4 independent 32-byte loads) -> few dependent round trips instead of one per block
    for (int yb = y0; yb <= y1; yb += 4) {
        int bs[4], nb[4], nbmax = 0;
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const int y = yb + r;
            bs[r] = 0;
            nb[r] = 0;
            if (y <= y1) {
                // distance from the query to this row's band, shrunk by the rounding slack -> half chord
                const float lo = q.gy0 + (float)y * cell, hi = lo + cell;
                const float dy = fmaxf(fmaxf(lo - fqy, fqy - hi) - slack, 0.0f);
                const float h2 = D0 * D0 - dy * dy;
                if (h2 >= 0.0f) {
                    const float hc = sqrtf(h2) * 1.000001f + slack;
                    const int x0 = (int)fmaxf(floorf((fqx - hc - q.gx0) * q.inv), 0.0f);
                    const int x1 = (int)fminf(floorf((fqx + hc - q.gx0) * q.inv), (float)(q.W - 1));
                    if (x0 <= x1) {
                        bs[r] = q.cs[y * q.W + x0];
                        nb[r] = q.cs[y * q.W + x1 + 1];
                    }
                }
            }
        }
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const int b = bs[r] & ~3; // whole 32-byte blocks: the extra points are real points too
            nb[r] = nb[r] > b ? (nb[r] - b + 3) >> 2 : 0;
            bs[r] = b;
            nbmax = nb[r] > nbmax ? nb[r] : nbmax;
        }
        for (int i = 0; i < nbmax; i++) {
            lcs_f8 v[4];
#pragma unroll
            for (int r = 0; r < 4; r++)
                if (i < nb[r]) v[r] = lcs_ld32((const lcs_f8*)(q.gp + bs[r]) + i);
#pragma unroll
            for (int r = 0; r < 4; r++)
                if (i < nb[r]) {
                    const int k = bs[r] + 4 * i;
                    LCS_SCAN_POINT(v[r].v[0], v[r].v[1], k)
                    LCS_SCAN_POINT(v[r].v[2], v[r].v[3], k + 1)
                    LCS_SCAN_POINT(v[r].v[4], v[r].v[5], k + 2)
                    LCS_SCAN_POINT(v[r].v[6], v[r].v[7], k + 3)
                }
        }
    }
    if (j < 0) return lcs_nearest_edge(p, s, qx, qy, offroad); // cannot happen: the hint itself lies in the disk
    const float thr = lcs_filter_thr(m, lcs_filter_eta(q.eta0, fqx, fqy, m));
    if (thr > thr0) return lcs_nearest_edge(p, s, qx, qy, offroad); // idem (monotonicity); kept as a guard
    int best = p.ge_idx[(size_t)s * p.E + j];
    if (m2 <= thr) {
        const int32_t* gi = p.ge_idx + (size_t)s * p.E;
        double bd = lcs_inf_d();
        best = 0x7fffffff;
        for (int y = y0; y <= y1; y++) {
            const float lo = q.gy0 + (float)y * cell, hi = lo + cell;
            const float dy = fmaxf(fmaxf(lo - fqy, fqy - hi) - slack, 0.0f);
            const float h2 = D0 * D0 - dy * dy;
            if (h2 < 0.0f) continue;
            const float hc = sqrtf(h2) * 1.000001f + slack;
            const int x0 = (int)fmaxf(floorf((fqx - hc - q.gx0) * q.inv), 0.0f);
            const int x1 = (int)fminf(floorf((fqx + hc - q.gx0) * q.inv), (float)(q.W - 1));
            if (x0 > x1) continue;
            const int b = q.cs[y * q.W + x0] & ~3, e = q.cs[y * q.W + x1 + 1];
            for (int k = b; k < e; k++) {
                const lcs_f2 pt = q.gp[k];
                const float dx = pt.x - fqx, dyy = pt.y - fqy;
                if (dx * dx + dyy * dyy <= thr) {
                    const int idx = gi[k];
                    const double ee = lcs_norm2_axis(lcs_dsub(e64[2 * idx], qx), lcs_dsub(e64[2 * idx + 1], qy));
                    if (ee < bd || (ee == bd && idx < best)) { bd = ee; best = idx; }
                }
            }
        }
    }
    *offroad = lcs_edge_side(p, s, qx, qy, best);
    return best;
}

// ------------------------------------------------------------------ oriented-box SAT (exact)

// reference: utils/polygon.py:238-268 corners_from_bboxes
LCS_D void lcs_corners(const double* ex /* x,y,v,c,s,L,W,h */, double* X, double* Y) {
    const double x = ex[0], y = ex[1], c = ex[3], s = ex[4];
    const double hl = ex[5] / 2, hw = ex[6] / 2;
    const double lc = lcs_dmul(hl, c), ls = lcs_dmul(hl, s), wc = lcs_dmul(hw, c), ws = lcs_dmul(hw, s);
    X[0] = lcs_dadd(lcs_dadd(x, lc), ws);
    X[1] = lcs_dadd(lcs_dsub(x, lc), ws);
    X[2] = lcs_dsub(lcs_dsub(x, lc), ws);
    X[3] = lcs_dsub(lcs_dadd(x, lc), ws);
    Y[0] = lcs_dsub(lcs_dadd(y, ls), wc);
    Y[1] = lcs_dsub(lcs_dsub(y, ls), wc);
    Y[2] = lcs_dadd(lcs_dsub(y, ls), wc);
    Y[3] = lcs_dadd(lcs_dadd(y, ls), wc);
}
// reference: utils/polygon.py:206-230 _overlap_a_over_b (projection on `first`'s two axes, strict > 0)
LCS_D bool lcs_overlap_a_over_b(const double* fX, const double* fY, double c, double s, const double* sX, const double* sY) {
    for (int axis = 0; axis < 2; axis++) {
        const double n0 = axis == 0 ? c : -s, n1 = axis == 0 ? s : c;
        double min_a = lcs_inf_d(), max_a = -lcs_inf_d(), min_b = lcs_inf_d(), max_b = -lcs_inf_d();
        for (int k = 0; k < 4; k++) {
            const double pa = lcs_dot2(fX[k], n0, fY[k], n1);
            const double pb = lcs_dot2(sX[k], n0, sY[k], n1);
            min_a = pa < min_a ? pa : min_a;
            max_a = pa > max_a ? pa : max_a;
            min_b = pb < min_b ? pb : min_b;
            max_b = pb > max_b ? pb : max_b;
        }
        const double hi = max_a < max_b ? max_a : max_b, lo = min_a > min_b ? min_a : min_b;
        if (!(lcs_dsub(hi, lo) > 0)) return false;
    }
    return true;
}
LCS_D bool lcs_collide(const double* ea, const double* eb) {
    double aX[4], aY[4], bX[4], bY[4];
    lcs_corners(ea, aX, aY);
    lcs_corners(eb, bX, bY);
    return lcs_overlap_a_over_b(aX, aY, ea[3], ea[4], bX, bY) && lcs_overlap_a_over_b(bX, bY, eb[3], eb[4], aX, aY);
}

// ------------------------------------------------------------------ lead vehicle (exact part)

// reference: agent/agent.py:249-299, one (ego, other) pair.  cos/sin of the relative heading come
// from the per-agent cos/sin by the angle-difference identity instead of
// cos(wrap(h_o - h)); the two agree to ~2 ulp, which moves no integer output.
LCS_D bool lcs_lead_pair(const double* eg, const double* eo, double* dis, double* vlead) {
    const double ch = eg[3], sh = eg[4], L = eg[5], W = eg[6];
    const double dx = lcs_dsub(eo[0], eg[0]), dy = lcs_dsub(eo[1], eg[1]);
    const double rx = lcs_dot2(dx, ch, dy, sh);
    const double ry = lcs_dot2(dx, -sh, dy, ch);
    const double la = eo[5], wa = eo[6];
    const double c = lcs_dfma(eo[4], sh, lcs_dmul(eo[3], ch)); // cos(ho - h)
    const double sn = lcs_dsub(lcs_dmul(eo[4], ch), lcs_dmul(eo[3], sh)); // sin(ho - h)
    const double ls = lcs_dmul(la / 2.0, sn), wc = lcs_dmul(wa / 2.0, c);
    const double y0 = lcs_dsub(lcs_dadd(ry, ls), wc), y1 = lcs_dadd(lcs_dadd(ry, ls), wc);
    const double y2 = lcs_dadd(lcs_dsub(ry, ls), wc), y3 = lcs_dsub(lcs_dsub(ry, ls), wc);
    const double ymax = fmax(fmax(y0, y1), fmax(y2, y3)), ymin = fmin(fmin(y0, y1), fmin(y2, y3));
    if (ymax < -W / 2.0 || ymin > W / 2.0) return false;
    const double lc = lcs_dmul(la / 2.0, c), ws = lcs_dmul(wa / 2.0, sn);
    const double x0 = lcs_dadd(lcs_dadd(rx, lc), ws), x1 = lcs_dsub(lcs_dadd(rx, lc), ws);
    const double x2 = lcs_dsub(lcs_dsub(rx, lc), ws), x3 = lcs_dadd(lcs_dsub(rx, lc), ws);
    const double xmin = fmin(fmin(x0, x1), fmin(x2, x3));
    if (xmin < L / 2.0) return false;
    *dis = lcs_dsub(xmin, L / 2.0);
    *vlead = lcs_dmul(eo[2], c);
    return true;
}

// ------------------------------------------------------------------ policies

// reference: policy/idm_policy.py:226-245 TrajIDM._get_acc with the class constants
LCS_D double lcs_idm_acc(double self_v, double target_v, double ahead_v, double distance) {
    const double min_gap = 10.0, headway = 2.0, max_a = 2.0, max_brake = -2.0;
    double acc;
    if (lcs_dsub(distance, min_gap) < 0) {
        acc = -lcs_inf_d();
    } else {
        // 2*sqrt(-usual_brake*max_a) = 2*sqrt(4) = 4 exactly
        const double t = lcs_dadd(lcs_dmul(self_v, headway), lcs_dmul(self_v, lcs_dsub(self_v, ahead_v)) / 4.0);
        const double s_star = lcs_dadd(min_gap, lcs_pymax(t, 0.0));
        const double r1 = self_v / target_v, r2 = s_star / distance;
        acc = lcs_dmul(max_a, lcs_dsub(lcs_dsub(1.0, lcs_pow4(r1)), lcs_dmul(r2, r2)));
    }
    return lcs_pymax(max_brake, lcs_pymin(max_a, acc));
}

// obs.ref_trajectory.next_waypoint() as the policy sees it (type.py:32-40); after a re-plan the
// observation still holds the OLD trajectory until the next prepare_obs (agent.py:247,250 vs :462-466)
LCS_D bool lcs_policy_waypoint(const LcsParams& p, size_t ia, int cursor, int len, double* wp) {
    if (p.stale_valid[ia]) {
        for (int k = 0; k < 5; k++) wp[k] = p.stale_wp[5 * ia + k];
        return p.stale_f32[ia] != 0;
    }
    const int w = len == 1 ? cursor : cursor + 1;
    const size_t t0 = ia * p.T + w;
    wp[0] = p.traj_xy[2 * t0];
    wp[1] = p.traj_xy[2 * t0 + 1];
    wp[2] = p.traj_vel[2 * t0];
    wp[3] = p.traj_vel[2 * t0 + 1];
    wp[4] = p.traj_h[t0];
    return p.traj_f32[ia] != 0;
}

// reference: agent/agent.py:67-84 HistoryState.update + :419-426; circular instead of shifting
LCS_D void lcs_ring_push_at(const LcsParams& p, size_t ia, int head, int c, double x, double y, double h, double v, double ch,
                            double sh, bool f32) {
    double* r = p.ring + (ia * LCS_H + head) * 5;
    LCS_ST_STREAM(r + 0, x);
    LCS_ST_STREAM(r + 1, y);
    if (f32) { // runtime.v / heading are np.float32 after a STATE action read from a float32 re-plan
        LCS_ST_STREAM(r + 2, (double)((float)v * cosf((float)h)));
        LCS_ST_STREAM(r + 3, (double)((float)v * sinf((float)h)));
    } else {
        LCS_ST_STREAM(r + 2, lcs_dmul(v, ch));
        LCS_ST_STREAM(r + 3, lcs_dmul(v, sh));
    }
    LCS_ST_STREAM(r + 4, h);
    p.ring_head[ia] = head + 1 == LCS_H ? 0 : head + 1;
    if (c < LCS_H) p.ring_count[ia] = c + 1;
}
LCS_D void lcs_ring_push(const LcsParams& p, size_t ia, double x, double y, double h, double v, double ch, double sh, bool f32) {
    lcs_ring_push_at(p, ia, p.ring_head[ia], p.ring_count[ia], x, y, h, v, ch, sh, f32);
}


// ------------------------------------------------------------------ float32 SAT pre-test

// Classic OBB separating-axis test in float32 on coordinates relative to box a (all magnitudes
// of a few metres, absolute error ~1e-6 m).  Equivalent to the reference's corner-projection test
// (polygon.py:206-230: min(max_a,max_b) - max(min_a,min_b) > 0 on the four box axes) whenever the
// largest signed axis gap is farther than LCS_SAT_MARGIN from zero; returns -1 otherwise and the
// caller falls back to the float64 restatement (lcs_collide).
#define LCS_SAT_MARGIN 1.0e-4f
LCS_D int lcs_sat_f32(const double* ea, const double* eb) {
    const float dx = (float)(eb[0] - ea[0]), dy = (float)(eb[1] - ea[1]);
    const float ca = (float)ea[3], sa = (float)ea[4], cb = (float)eb[3], sb = (float)eb[4];
    const float hla = 0.5f * (float)ea[5], hwa = 0.5f * (float)ea[6], hlb = 0.5f * (float)eb[5], hwb = 0.5f * (float)eb[6];
    const float ac = fabsf(ca * cb + sa * sb), as = fabsf(sb * ca - cb * sa);
    const float t1 = fabsf(dx * ca + dy * sa) - (hla + hlb * ac + hwb * as);
    const float t2 = fabsf(dy * ca - dx * sa) - (hwa + hlb * as + hwb * ac);
    const float t3 = fabsf(dx * cb + dy * sb) - (hlb + hla * ac + hwa * as);
    const float t4 = fabsf(dy * cb - dx * sb) - (hwb + hla * as + hwa * ac);
    const float tmax = fmaxf(fmaxf(t1, t2), fmaxf(t3, t4));
    if (tmax > LCS_SAT_MARGIN) return 0;
    if (tmax < -LCS_SAT_MARGIN) return 1;
    return -1;
}

// ------------------------------------------------------------------ the tick, phase by phase
// Phases are separated by __syncthreads() in the kernel (lcsim_b200.cu) and by loop boundaries in
// the host emulation.  Threads map to LIST SLOTS (dense lanes: slot i < n_active), except in reset
// mode where phase A maps thread -> agent index.  Per-agent float64 records travel between phases
// through shared memory indexed by agent (sm.ex); the new list gets a slot-indexed float32 filter
// record (sm.crec).

struct LcsThread { // per-thread values that live across phases (registers on the device)
    bool fin, pop, act;
    int aw;
    int snap; // ring slot of the snapshot this item writes / reads
    int k; // tick index inside the launch (row of tick_log)
};

LCS_D void lcs_store_ex(LcsSmem& sm, int a, double x, double y, double v, double ch, double sh, double L, double W, double h) {
    double* ex = sm.ex + 8 * a;
    ex[0] = x; ex[1] = y; ex[2] = v; ex[3] = ch; ex[4] = sh; ex[5] = L; ex[6] = W; ex[7] = h;
}
LCS_D void lcs_store_pose(const LcsParams& p, int s, size_t ia, double x, double y, double h, double v) {
    double* ps = p.pose64 + 4 * ia;
    ps[0] = x; ps[1] = y; ps[2] = h; ps[3] = v;
    lcs_f4 p32;
    p32.x = (float)(x - p.origin[2 * s]);
    p32.y = (float)(y - p.origin[2 * s + 1]);
    p32.z = (float)h;
    p32.w = (float)v;
#if defined(LCS_CS_HINTS) && !defined(LCSIM_HOSTEMU)
    __stcs((float4*)&p.pose32[ia], make_float4(p32.x, p32.y, p32.z, p32.w));
#else
    p.pose32[ia] = p32;
#endif
}

#ifndef LCSIM_HOSTEMU
// reset launch: the history rings of a scenario (Ap x 11 x 5 doubles, contiguous) cleared by the whole CTA with 16-byte
// stores — per-thread loops over the agent's own row were 55 stores of 8 bytes at a stride of 440 bytes per lane,
// i.e. partial-sector writes: ~40 us of the 87 us reset launch at S=1024.  The caller puts a barrier behind it.
LCS_D void lcs_ring_clear(const LcsParams& p, int s) {
    double2* r = (double2*)(p.ring + (size_t)s * p.Ap * LCS_H * 5); // Ap is a multiple of 32: the block is 16-byte aligned
    const int n2 = p.Ap * LCS_H * 5 / 2;
    for (int k = threadIdx.x; k < n2; k += blockDim.x) r[k] = make_double2(0.0, 0.0);
}
#endif
// Agent.__init__ / Engine.reset for agent a (agent.py:106-171, engine.py:132-143)
LCS_D void lcs_reset_agent(const LcsParams& p, int s, int a) {
    const size_t ia = (size_t)s * p.Ap + a;
    p.status[ia] = LCS_WAITING;
    p.cursor[ia] = 0;
    p.ring_head[ia] = 0;
    p.ring_count[ia] = 0;
    p.lead_valid[ia] = 0;
    p.lead_dis[ia] = lcs_inf_d();
    p.lead_v[ia] = 0.0;
    p.stale_valid[ia] = 0;
    p.coll_count[ia] = 0;
    p.nearest_edge[ia] = -1;
    p.edge_hint[ia] = p.home_edge[ia];
    p.offroad[ia] = 0;
#ifdef LCSIM_HOSTEMU // (device: the CTA clears the scenario's whole ring with coalesced stores first, lcs_ring_clear)
    for (int k = 0; k < LCS_H * 5; k++) p.ring[ia * LCS_H * 5 + k] = 0.0;
#endif
    if (a >= p.n_agents[s]) { // padding agents: inert but defined
        lcs_store_pose(p, s, ia, 0.0, 0.0, 0.0, 0.0);
        return;
    }
    if (p.traj_f32[ia]) { // restore a re-planned trajectory
        p.traj_len[ia] = p.traj_len0[ia];
        for (int k = 0; k < p.T; k++) {
            const size_t t = ia * p.T + k;
            p.traj_xy[2 * t] = p.traj_xy0[2 * t];
            p.traj_xy[2 * t + 1] = p.traj_xy0[2 * t + 1];
            p.traj_vel[2 * t] = p.traj_vel0[2 * t];
            p.traj_vel[2 * t + 1] = p.traj_vel0[2 * t + 1];
            p.traj_h[t] = p.traj_h0[t];
        }
        for (int k = 0; k < p.Tp; k++) {
            const size_t f = lcs_tf_index(p, s, k, a);
            p.tf_xy[f] = p.tf_xy0[f];
            p.tf_rho[f] = p.tf_rho0[f];
        }
        p.traj_blob[ia] = p.traj_blob0[ia];
        p.traj_f32[ia] = 0;
    }
    const double vx = p.traj_vel[2 * ia * p.T], vy = p.traj_vel[2 * ia * p.T + 1];
    const double x = p.home_xy[2 * ia], y = p.home_xy[2 * ia + 1];
    const double v = sqrt(lcs_dadd(lcs_dmul(vx, vx), lcs_dmul(vy, vy))); // agent.py:144-146
    const double h = p.traj_h[ia * p.T];
    double sh, ch;
    lcs_sincos(h, &sh, &ch);
    lcs_ring_push(p, ia, x, y, h, v, ch, sh, false); // agent.py:159-165
    lcs_store_pose(p, s, ia, x, y, h, v);
}

// Engine.update for one agent (engine.py:221-225 -> manager.py:105-115 -> agent.py:310-426), in three
// parts so the Trajectory.next scan in the middle can run off the shared-memory ring:
//   pre  : action selection + state update, leaves the new state and the scan query in LcsUpd
//   scan : float32 filter scan (lcs_scan_*)
//   post : cursor, history ring, stores; returns true when the agent reached the end of its trajectory
struct LcsUpd {
    double x, y, h, v, ch, sh, qx, qy;
    int len, cursor;
    int ring_head, ring_count; // loaded with the rest of the state (the post part would otherwise wait for them)
    double L, W;
    int direct; // >= 0: the cursor is known without a scan (replay record)
    bool ring_f32;
    bool moved; // new position differs from the old one
};
LCS_D void lcs_update_pre(const LcsParams& p, int s, int a, LcsUpd& u) {
    const size_t ia = (size_t)s * p.Ap + a;
    const double* ps = p.pose64 + 4 * ia;
    double x = ps[0], y = ps[1], h = ps[2], v = ps[3], ch, sh;
    const int len = p.traj_len[ia];
    const int cursor = p.cursor[ia];
    // ---- which action (manager.py:111-115)
    int type = LCS_ACT_NONE;
    double d0 = 0, d1 = 0, d2 = 0, d3 = 0, d4 = 0;
    bool f32 = false;
    for (int k = 0; k < p.K; k++) {
        const size_t q = (size_t)s * p.K + k;
        if (p.act_agent[q] == a && p.act_type[q] != LCS_ACT_NONE) {
            type = p.act_type[q];
            d0 = p.act_data[5 * q];
            d1 = p.act_data[5 * q + 1];
            d2 = p.act_data[5 * q + 2];
            d3 = p.act_data[5 * q + 3];
            d4 = p.act_data[5 * q + 4];
        }
    }
    if (p.ads_action && a == p.ads_index[s]) {
        // BicycleAction.init_with_normalized (type.py:100-102) as BicycleSingleEnv.step applies it
        type = LCS_ACT_BICYCLE;
        d0 = lcs_dmul(p.ads_action[2 * s], 6.0);
        d1 = lcs_dmul(p.ads_action[2 * s + 1], 0.3);
    }
    u.direct = -1;
    u.cursor = cursor;
    u.ring_head = p.ring_head[ia];
    u.ring_count = p.ring_count[ia];
    u.L = p.length[ia];
    u.W = p.width[ia];
    const double x_old = x, y_old = y;
    if (type == LCS_ACT_NONE) {
        double wp[5];
        bool wp_f32 = false;
        double vn;
        // pristine float64 trajectory and a fresh observation: the waypoint comes as one replay record
        const bool use_rec = p.wp_rec && !p.stale_valid[ia] && !p.traj_f32[ia];
        if (use_rec) {
#if defined(LCSIM_HOSTEMU) || defined(LCS_NO_STREAM_EF)
            const LcsWpRec r = p.wp_rec[ia * p.T + (len == 1 ? cursor : cursor + 1)];
#else
            LcsWpRec r; // read once per rollout: L2 evict-first
            {
                const LcsWpRec* rp = p.wp_rec + ia * p.T + (len == 1 ? cursor : cursor + 1);
                long long w7;
                asm("ld.global.L2::evict_first.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.h), "=d"(r.v) : "l"(rp));
                asm("ld.global.L2::evict_first.v4.b64 {%0,%1,%2,%3}, [%4];"
                    : "=l"(*(long long*)&r.ch), "=l"(*(long long*)&r.sh), "=l"(*(long long*)&r.vn), "=l"(w7)
                    : "l"((const char*)rp + 32));
                r.canon = (int32_t)(w7 & 0xffffffffll);
            }
#endif
            if (!p.reactive) {
                // StateExpert -> STATE action -> Trajectory.next(new xy): all of it is in the record
                u.x = r.x; u.y = r.y; u.h = r.h; u.v = r.v; u.ch = r.ch; u.sh = r.sh; u.qx = r.x; u.qy = r.y;
                u.len = len;
                u.direct = r.canon;
                u.ring_f32 = false;
                u.moved = !(r.x == x_old && r.y == y_old);
                return;
            }
            wp[4] = r.h;
            vn = r.vn;
        } else {
            wp_f32 = lcs_policy_waypoint(p, ia, cursor, len, wp);
            vn = lcs_norm2_vec(wp[2], wp[3]);
            if (wp_f32) {
                const float fvx = (float)wp[2], fvy = (float)wp[3];
                vn = (double)sqrtf(fmaf(fvy, fvy, lcs_fmul(fvx, fvx))); // float32 BLAS-style dot
            }
        }
        if (p.reactive) {
            // ---- TrajIDM.get_action (idm_policy.py:196-224), dt = 0.1 default argument
            const double pdt = 0.1;
            double acc = lcs_npclip(lcs_dsub(vn, v) / pdt, -3.0, 3.0);
            const double dh = lcs_wrap_angle(lcs_dsub(wp[4], h));
            double steer;
            if (v < 0.6 || vn < 0.6)
                steer = 0.0;
            else
                steer = dh / lcs_dadd(lcs_dmul(v, pdt), lcs_dmul(lcs_dmul(0.5, acc), lcs_dmul(pdt, pdt)));
            steer = lcs_npclip(steer, -0.3, 0.3);
            if (p.lead_valid[ia]) {
                const double idm = lcs_idm_acc(v, 40.0, p.lead_v[ia], p.lead_dis[ia]);
                acc = lcs_pymax(lcs_pymin(acc, idm), -3.0);
            }
            type = LCS_ACT_BICYCLE;
            d0 = acc;
            d1 = steer;
        } else {
            // ---- StateExpert.get_action (expert_policy.py:18-33)
            type = LCS_ACT_STATE;
            d0 = wp[0]; d1 = wp[1]; d2 = wp[2]; d3 = wp[3]; d4 = wp[4];
            f32 = wp_f32;
        }
    }
    // ---- update_runtime_by_action (agent.py:377-416)
    double qx = x, qy = y; // position fed to Trajectory.next
    bool ring_f32 = false;
    const double dt = p.dt;
    if (type == LCS_ACT_STATE) {
        x = d0;
        y = d1;
        if (f32) {
            const float fvx = (float)d2, fvy = (float)d3;
            v = (double)sqrtf(lcs_fadd(lcs_fmul(fvx, fvx), lcs_fmul(fvy, fvy))); // np.sqrt(vx**2+vy**2) in float32
            ring_f32 = true;
        } else {
            v = sqrt(lcs_dadd(lcs_dmul(d2, d2), lcs_dmul(d3, d3)));
        }
        h = d4;
        qx = x;
        qy = y;
        lcs_sincos(h, &sh, &ch);
    } else if (type == LCS_ACT_BICYCLE) {
        const double acc = d0, steer = d1;
        const double delta = lcs_dadd(lcs_dmul(v, dt), lcs_dmul(lcs_dmul(0.5, acc), lcs_dmul(dt, dt)));
        const double nh = lcs_wrap_angle(lcs_dadd(h, lcs_dmul(steer, delta)));
        lcs_sincos(nh, &sh, &ch);
        x = lcs_dadd(x, lcs_dmul(delta, ch));
        y = lcs_dadd(y, lcs_dmul(delta, sh));
        v = lcs_dadd(v, lcs_dmul(acc, dt));
        h = nh;
    } else { // DELTA
        x = lcs_dadd(x, d0);
        y = lcs_dadd(y, d1);
        h = lcs_wrap_angle(lcs_dadd(h, d2));
        lcs_sincos(h, &sh, &ch);
    }
    u.x = x; u.y = y; u.h = h; u.v = v; u.ch = ch; u.sh = sh; u.qx = qx; u.qy = qy;
    u.len = len;
    u.ring_f32 = ring_f32;
    u.moved = !(x == x_old && y == y_old);
}
LCS_D bool lcs_update_post(const LcsParams& p, LcsSmem& sm, int s, int a, const LcsUpd& u, const LcsScan& sc) {
    const size_t ia = (size_t)s * p.Ap + a;
    const int cursor = u.direct >= 0 ? u.direct : lcs_scan_finish(p, s, a, u.len, u.qx, u.qy, sc);
    p.cursor[ia] = cursor;
    sm.moved[a] = u.moved ? 1 : 0;
    lcs_ring_push_at(p, ia, u.ring_head, u.ring_count, u.x, u.y, u.h, u.v, u.ch, u.sh, u.ring_f32);
    lcs_store_pose(p, s, ia, u.x, u.y, u.h, u.v);
    lcs_store_ex(sm, a, u.x, u.y, u.v, u.ch, u.sh, u.L, u.W, u.h);
    const bool fin = cursor >= u.len - 1; // type.py:50-51 reach_end
    if (fin) {
        // FINISHED now, IDLE after this tick's check_departure_and_arrival (single-trip schedules:
        // schedule.py:22-40 next_trip() is False); the agent leaves the list, its metric outputs reset
        p.status[ia] = LCS_IDLE; // (the metric kernels reset its collision / edge outputs: lcs_mc_load, lcs_me_run)
    }
    return fin;
}

// Phase A: shared-memory init + update (step mode: thread = old list slot) or reset (thread = agent).
// lcs_phase_init returns the agent this thread updates (-1: none).
LCS_D int lcs_phase_init(const LcsParams& p, LcsSmem& sm, int s, int tid, LcsThread& t) {
    for (int k = tid; k < 96; k += p.Ap) sm.mask[k] = 0;
    const int n_old = p.mode == 1 ? 0 : p.n_active[s];
    if (tid == 0) {
        sm.misc[0] = n_old;
        sm.misc[1] = 0; // n_new
        sm.misc[2] = 0; // collision sum
        sm.misc[3] = 0; // off-road sum
        sm.misc[4] = p.mode == 1 ? 0 : p.tick[s];
        sm.misc[5] = p.mode == 1 ? 0 : p.wait_ptr[s];
        sm.misc[6] = 0; // pair-queue length
        sm.misc[8] = sm.misc[9] = sm.misc[10] = sm.misc[11] = 0; // flag totals
        sm.misc[14] = sm.misc[15] = 0; // largest bounding radius / coordinate slack of the new list (float bits)
    }
    sm.cnt[tid] = 0;
    sm.moved[tid] = 1; // agents that enter the list this tick have no edge outputs yet
    sm.lead_key[tid] = ~0ull;
    t.fin = false;
    int a = -1;
    if (p.mode == 1)
        lcs_reset_agent(p, s, tid);
    else
        a = p.active[(size_t)s * p.Ap + tid]; // -1 beyond the list (lcs_phase_publish): no need to wait for n_active
    sm.act[tid] = a;
    return a;
}
LCS_D void lcs_phase_update(const LcsParams& p, LcsSmem& sm, int s, int tid, LcsThread& t) {
    const int a = lcs_phase_init(p, sm, s, tid, t);
    LcsUpd u;
    LcsScan sc;
    bool whole_row = false;
    u.len = 0;
    if (a >= 0) {
        lcs_update_pre(p, s, a, u);
        if (u.direct < 0) whole_row = !lcs_scan_first(p, s, a, u.len, u.cursor, u.qx, u.qy, sc);
    }
#ifndef LCSIM_HOSTEMU
    lcs_scan_rows_warp(p, s, whole_row, a, u.len, sc);
#else
    if (whole_row) lcs_scan_global(p, s, a, u.len, sc);
#endif
    if (a >= 0) t.fin = lcs_update_post(p, sm, s, a, u, sc);
}

// snapshot ring: entry of (this launch's ring slot, scenario s)
LCS_HD size_t lcs_snap_base(const LcsParams& p, int snap, int s) { return ((size_t)snap * p.S + s) * p.Ap; }
LCS_HD size_t lcs_snap_hdr(const LcsParams& p, int snap, int s) { return ((size_t)snap * p.S + s) * 4; }
#define LCS_SNAP_MOVED 0x40000000

// Phase B1: AgentManager.check_departure_and_arrival (manager.py:50-96), flags.  tid = list slot and,
// independently, offset into the waiting list (sorted by departure time, manager.py:33-35).
LCS_D void lcs_phase_depart_flags(const LcsParams& p, LcsSmem& sm, int s, int tid, LcsThread& t) {
    const int tick = sm.misc[4], wp = sm.misc[5];
    const bool allow = p.mode == 1 ? true : (p.allow_new_obj != 0);
    t.pop = t.act = false;
    t.aw = -1;
    if (allow && wp + tid < p.n_agents[s]) {
        t.pop = p.depart_sorted[(size_t)s * p.Ap + wp + tid] <= tick;
        if (t.pop) {
            const int aw = p.wait_order[(size_t)s * p.Ap + wp + tid];
            t.aw = aw;
            t.act = !(p.no_static && !p.dynamic[(size_t)s * p.Ap + aw] && aw != p.ads_index[s]);
        }
    }
    lcs_set_flag_count(sm.mask, tid, t.fin, &sm.misc[8]);
    lcs_set_flag_count(sm.mask + 32, tid, t.pop, &sm.misc[9]);
    lcs_set_flag_count(sm.mask + 64, tid, t.act, &sm.misc[10]);
}

// Phase B2: new active list — departures reuse the finished slots in order, else append; leftover
// finished slots are dropped preserving order.  Departing agents publish their record.
LCS_D void lcs_phase_depart_place(const LcsParams& p, LcsSmem& sm, int s, int tid, LcsThread& t) {
    const int n = sm.misc[0];
    const int n_arr = sm.misc[8], m = sm.misc[10];
    if (tid < n) {
        const int r = lcs_prefix(sm.mask, tid);
        if (!t.fin)
            sm.act2[tid - (r > m ? r - m : 0)] = sm.act[tid];
        else if (p.with_metrics)
            p.snap_fin[lcs_snap_base(p, t.snap, s) + r] = sm.act[tid]; // the metric kernels reset this agent's outputs
    }
    if (t.act) {
        const int q = lcs_prefix(sm.mask + 64, tid), aw = t.aw;
        const size_t ia = (size_t)s * p.Ap + aw;
        p.status[ia] = LCS_ACTIVE;
        sm.dlist[q] = aw;
        if (q >= n_arr) sm.act2[n + (q - n_arr)] = aw;
        const double* ps = p.pose64 + 4 * ia;
        double sh, ch;
        lcs_sincos(ps[2], &sh, &ch);
        lcs_store_ex(sm, aw, ps[0], ps[1], ps[3], ch, sh, p.length[ia], p.width[ia], ps[2]);
    }
}
LCS_D void lcs_phase_depart_fill(const LcsParams& p, LcsSmem& sm, int s, int tid, LcsThread& t) {
    const int n = sm.misc[0];
    const int n_arr = sm.misc[8], n_pop = sm.misc[9], m = sm.misc[10];
    if (tid < n && t.fin) {
        const int r = lcs_prefix(sm.mask, tid);
        if (r < m) sm.act2[tid] = sm.dlist[r]; // no slot before this one was dropped when r < m
    }
    if (tid == 0) {
        const int n_new = n - n_arr + m;
        sm.misc[1] = n_new;
        p.wait_ptr[s] = sm.misc[5] + n_pop;
        p.n_active[s] = n_new;
        int64_t* st = p.stats + (size_t)s * LCS_NSTAT;
        if (p.mode == 1) {
            for (int k = 0; k < LCS_NSTAT; k++) st[k] = 0;
            st[5] = m;
            p.tick[s] = 0;
        } else {
            st[0] += n;
            st[1] += n_new;
            st[4] += n_arr;
            st[5] += m;
            st[6] += 1;
            p.tick[s] = sm.misc[4] + 1;
            if (p.tick_log) {
                int32_t* lg = p.tick_log + ((size_t)t.k * p.S + s) * 4;
                lg[0] = n;
                lg[1] = n_new;
                if (!p.with_metrics) lg[2] = lg[3] = 0;
            }
        }
        if (p.with_metrics) {
            int32_t* h = p.snap_hdr + lcs_snap_hdr(p, t.snap, s);
            h[0] = n;
            h[1] = n_new;
            h[2] = n_arr;
        }
    }
}

// float32 filter records of a list slot from its float64 state (bounding circle with the coordinate slack, axes)
LCS_D void lcs_make_crec(const LcsParams& p, int s, double x, double y, double ch, double sh, double L64, double W64, lcs_f4* c1,
                         lcs_f4* c2) {
    lcs_f4 r;
    r.x = (float)(x - p.origin[2 * s]);
    r.y = (float)(y - p.origin[2 * s + 1]);
    const float L = (float)L64, W = (float)W64;
    r.w = 5e-5f + 1.2e-7f * (fabsf(r.x) + fabsf(r.y)); // float32 coordinate error
    r.z = 0.5f * sqrtf(L * L + W * W) * 1.00001f + r.w; // half diagonal, rounded up
    *c1 = r;
    lcs_f4 r2;
    r2.x = (float)ch;
    r2.y = (float)sh;
    r2.z = 0.5f * L;
    r2.w = 0.5f * W;
    *c2 = r2;
}

// Phase C: publish the list; snapshot of the new list for the metric kernels; slot-indexed float32 filter records
// for the lead search.
LCS_D void lcs_phase_publish(const LcsParams& p, LcsSmem& sm, int s, int tid, const LcsThread& t) {
    const size_t is = (size_t)s * p.Ap + tid;
    const int n_new = sm.misc[1];
    p.stale_valid[is] = 0; // prepare_obs refreshed every observation (manager.py:98-103)
    if (tid >= n_new) {
        p.active[is] = -1;
        if (p.reactive) sm.soa[tid] = sm.soa[LCS_SOA_Y(p.Ap) + tid] = LCS_FAR; // never inside a corridor
        return;
    }
    const int a = sm.act2[tid];
    p.active[is] = a;
    const double* ex = sm.ex + 8 * a;
    if (p.with_metrics) {
        // an agent whose position is bit-identical to the one its edge outputs were computed for keeps them
        const size_t q = lcs_snap_base(p, t.snap, s) + tid;
        p.snap_agent[q] = a | (sm.moved[a] ? LCS_SNAP_MOVED : 0);
        double* r = p.snap_rec + 6 * q;
        r[0] = ex[0]; r[1] = ex[1]; r[2] = ex[3]; r[3] = ex[4]; r[4] = ex[5]; r[5] = ex[6];
    }
    if (p.reactive) {
        lcs_make_crec(p, s, ex[0], ex[1], ex[3], ex[4], ex[5], ex[6], &sm.crec[tid], &sm.crec2[tid]);
        // coordinates once more as two float arrays for the lead search's first test (four partners per load), and the
        // largest bounding radius / coordinate slack of the list (positive floats order like their bit patterns)
        sm.soa[tid] = sm.crec[tid].x;
        sm.soa[LCS_SOA_Y(p.Ap) + tid] = sm.crec[tid].y;
        lcs_atomic_max(&sm.misc[14], lcs_f2i(sm.crec[tid].z));
        lcs_atomic_max(&sm.misc[15], lcs_f2i(sm.crec[tid].w));
    }
}

LCS_D void lcs_push_pair(LcsSmem& sm, int cap, int i, int j, bool* overflow) {
    const int q = lcs_atomic_add(&sm.misc[6], 1);
    *overflow = q >= cap;
    if (q < cap) sm.queue[q] = ((uint32_t)i << 16) | (uint32_t)j;
}

// TrajIDM branch of Agent.prepare_obs (agent.py:249-299): the nearest agent ahead in the ego's corridor.  The
// reference walks the active list and keeps the strict minimum of dis = min(corner x) - length / 2 among the agents
// that pass its two tests (corner y range meets [-w/2, w/2]; min corner x >= length / 2).
// One pass in float32 over the slot records RANKS the partners: min corner x = rx - (|hl cos| + |hw sin|), corner
// y range = ry -+ (|hl sin| + |hw cos|) with rx, ry the partner's centre in the ego frame (the four sign combinations
// of the reference's corner formulas).  e bounds the float32 error of those quantities, so a partner is certainly
// out / certainly in / undecided; the float64 restatement (lcs_lead_pair) is evaluated for the unique winner only:
// it must be certainly in and its error interval must lie below every other possible partner's.  Otherwise
// (rare: ties, partners on the edge of the corridor) the thread falls back to the float64 evaluation of every partner
// of the list, in list order with the reference's strict <.
#define LCS_LEAD_LIST 16 // private survivor list per thread (uint16 entries in the pair-queue area)
// full float32 classification of partner slot j for ego slot tid: certainly out / possible (interval [d - e, d + e] of its
// true distance, `sure` when certainly in); keeps the partner with the smallest upper end and the smallest lower end
// among the others
LCS_D void lcs_lead_classify(const LcsSmem& sm, int tid, int j, const lcs_f4 me, float c32, float s32, float hl, float hw, float& h1,
                             float& l1, float& lo2, int& j1, bool& sure1) {
    const lcs_f4 ot = sm.crec[j];
    const float dx = ot.x - me.x, dy = ot.y - me.y;
    const float rxx = dx * c32 + dy * s32, ryy = dy * c32 - dx * s32;
    // error of rx, ry and of the corner extents: coordinate rounding (slack) + float32 trig / arithmetic
    const float e = ot.w + me.w + 3e-6f * (fabsf(dx) + fabsf(dy)) + 2e-5f;
    // every corner lies within the other's half diagonal (ot.z) of its centre; a partner whose interval lies above both
    // h1 and lo2 changes neither
    if (fabsf(ryy) > hw + ot.z + e || rxx < hl - e || rxx - ot.z - hl - e > fmaxf(h1, lo2)) return;
    const lcs_f4 o2 = sm.crec2[j];
    const float cr = o2.x * c32 + o2.y * s32, sr = o2.y * c32 - o2.x * s32; // cos / sin of the relative heading
    const float ex = fabsf(o2.z * cr) + fabsf(o2.w * sr), ey = fabsf(o2.z * sr) + fabsf(o2.w * cr);
    const float xmin = rxx - ex, ymax = ryy + ey, ymin = ryy - ey;
    if (ymax < -hw - e || ymin > hw + e || xmin < hl - e) return; // certainly not a lead vehicle
    const bool sure = ymax >= -hw + e && ymin <= hw - e && xmin >= hl + e;
    const float d = xmin - hl, lo = d - e, hi = d + e;
    if (hi < h1) {
        if (j1 >= 0) lo2 = fminf(lo2, l1);
        h1 = hi; l1 = lo; j1 = j; sure1 = sure;
    } else {
        lo2 = fminf(lo2, lo);
    }
}

LCS_D void lcs_phase_lead(const LcsParams& p, LcsSmem& sm, int s, int tid) {
    const int n_new = sm.misc[1];
    if (tid >= n_new) return;
    const double* eg = sm.ex + 8 * sm.act2[tid];
    const lcs_f4 me = sm.crec[tid], me2 = sm.crec2[tid];
    const float c32 = me2.x, s32 = me2.y, hl = me2.z, hw = me2.w;
    // possible partners as intervals [d - e, d + e] of their true distance: the one with the smallest upper end (h1)
    // wins outright when it is certainly in and h1 lies below the lower ends of all others (lo2 = their minimum)
    float h1 = lcs_inf_f(), lo2 = lcs_inf_f();
    float l1 = 0.f;
    int j1 = -1;
    bool sure1 = false;
    // First test, four partners per step on the coordinate arrays: the centre must lie in the corridor widened by the
    // largest bounding radius of the list and not behind the ego (conservative: rmax, emax bound every partner's radius
    // and error).  A partner whose CENTRE lies well inside the corridor and whose whole circle is ahead of the ego is
    // certainly a lead candidate no farther than rx - hl: `bound` (the smallest such value + emax so far) prunes everybody
    // whose circle starts beyond it — they can neither win nor touch the winner's error interval.
    const float* sx = sm.soa;
    const float* sy = sm.soa + LCS_SOA_Y(p.Ap);
    const float rmax = lcs_i2f(sm.misc[14]), wmax = lcs_i2f(sm.misc[15]);
    // e <= 2 wmax + 3e-6 (|dx| + |dy|) + 2e-5 with |dx| + |dy| <= 2 (wmax - 5e-5) / 1.2e-7; 1e-4 more for a differently
    // contracted evaluation of rx, ry here and below
    const float emax = 52.0f * wmax + 1.2e-4f;
    const float cy = hw + rmax + emax, cx = hl - emax, cy_in = hw - emax, cx_in = hl + rmax + emax;
    const float off = rmax + hl + emax;
    // pass 1 (no divergence): the running bound, and the partners that can matter under it go to the thread's private
    // list (a column of shared memory) — classifying them inside this loop would make the warp walk the classification
    // for the union of its lanes' survivors
    float bound = lcs_inf_f();
    uint16_t* list = (uint16_t*)sm.queue; // [LCS_LEAD_LIST][Ap]
    int nl = 0;
    for (int j4 = 0; j4 < n_new; j4 += 4) {
        const lcs_f4 X = *(const lcs_f4*)(sx + j4), Y = *(const lcs_f4*)(sy + j4);
        float rx[4], ry[4];
        {
            const float dx0 = X.x - me.x, dx1 = X.y - me.x, dx2 = X.z - me.x, dx3 = X.w - me.x;
            const float dy0 = Y.x - me.y, dy1 = Y.y - me.y, dy2 = Y.z - me.y, dy3 = Y.w - me.y;
            rx[0] = dx0 * c32 + dy0 * s32; rx[1] = dx1 * c32 + dy1 * s32; rx[2] = dx2 * c32 + dy2 * s32; rx[3] = dx3 * c32 + dy3 * s32;
            ry[0] = fabsf(dy0 * c32 - dx0 * s32); ry[1] = fabsf(dy1 * c32 - dx1 * s32);
            ry[2] = fabsf(dy2 * c32 - dx2 * s32); ry[3] = fabsf(dy3 * c32 - dx3 * s32);
        }
        unsigned m = 0;
#pragma unroll
        for (int r = 0; r < 4; r++) {
            if (ry[r] <= cy_in && rx[r] >= cx_in) bound = fminf(bound, rx[r] - hl + emax); // (the ego itself: rx = 0 < cx_in)
            if (ry[r] <= cy && rx[r] >= cx && rx[r] - off <= bound) m |= 1u << r;
        }
        while (m) {
            const int r = lcs_ffs0(m);
            m &= m - 1;
            const int j = j4 + r;
            if (j >= n_new || j == tid) continue;
            if (nl < LCS_LEAD_LIST)
                list[nl++ * p.Ap + tid] = (uint16_t)j;
            else
                lcs_lead_classify(sm, tid, j, me, c32, s32, hl, hw, h1, l1, lo2, j1, sure1); // (list full: rare)
        }
    }
    // pass 2: the listed partners, every lane on its own list
    for (int k = 0; k < nl; k++) {
        const int j = list[k * p.Ap + tid];
        const float dx = sm.crec[j].x - me.x, dy = sm.crec[j].y - me.y;
        if (dx * c32 + dy * s32 - off > bound) continue; // listed before the bound tightened
        lcs_lead_classify(sm, tid, j, me, c32, s32, hl, hw, h1, l1, lo2, j1, sure1);
    }
    const size_t ia = (size_t)s * p.Ap + sm.act2[tid];
    double dis = lcs_inf_d(), vl = 0.0;
    bool valid = false;
    if (j1 >= 0) {
        bool decided = false;
        if (sure1 && h1 < lo2) {
            decided = lcs_lead_pair(eg, sm.ex + 8 * sm.act2[j1], &dis, &vl);
            valid = decided;
        }
        LCS_COUNT(decided ? 3 : 4);
        if (!decided) { // exact fallback
            dis = lcs_inf_d();
            vl = 0.0;
            valid = false;
            for (int j = 0; j < n_new; j++) {
                if (j == tid) continue;
                double d64, v64;
                if (lcs_lead_pair(eg, sm.ex + 8 * sm.act2[j], &d64, &v64) && d64 < dis) {
                    dis = d64; vl = v64; valid = true;
                }
            }
        }
    }
    if (j1 < 0) LCS_COUNT(5);
    p.lead_valid[ia] = valid ? 1 : 0;
    p.lead_dis[ia] = valid ? dis : lcs_inf_d();
    p.lead_v[ia] = valid ? vl : 0.0;
}

// float64 narrowphase of one queued pair (the float32 test could not decide it)
LCS_D void lcs_collide_eval(LcsSmem& sm, int i, int j) {
    if (lcs_collide(sm.ex + 8 * sm.act2[i], sm.ex + 8 * sm.act2[j])) {
        lcs_atomic_add(&sm.cnt[i], 1);
        lcs_atomic_add(&sm.cnt[j], 1);
    }
}
// float32 SAT on the slot records; dx, dy carry the coordinate slack of both boxes -> wider margin
LCS_D int lcs_sat_rec(const lcs_f4 a, const lcs_f4 a2, const lcs_f4 b, const lcs_f4 b2) {
    const float dx = b.x - a.x, dy = b.y - a.y;
    const float ca = a2.x, sa = a2.y, cb = b2.x, sb = b2.y;
    const float ac = fabsf(ca * cb + sa * sb), as = fabsf(sb * ca - cb * sa);
    const float t1 = fabsf(dx * ca + dy * sa) - (a2.z + b2.z * ac + b2.w * as);
    const float t2 = fabsf(dy * ca - dx * sa) - (a2.w + b2.z * as + b2.w * ac);
    const float t3 = fabsf(dx * cb + dy * sb) - (b2.z + a2.z * ac + a2.w * as);
    const float t4 = fabsf(dy * cb - dx * sb) - (b2.w + a2.z * as + a2.w * ac);
    const float tmax = fmaxf(fmaxf(t1, t2), fmaxf(t3, t4));
    const float mrg = LCS_SAT_MARGIN + 1.5f * (a.w + b.w);
    if (tmax > mrg) return 0;
    if (tmax < -mrg) return 1;
    return -1;
}

// Phase E1: check_collision_all (manager.py:313-327 -> polygon.py:195-235): bounding-circle broadphase, then
// a float32 SAT on the survivors.  Every unordered pair of slots is visited once: slot i tests i+1 .. i+n/2 (mod n).
LCS_D void lcs_collision_narrow(LcsSmem& sm, int cap, int i, int j, const lcs_f4 me, const lcs_f4 me2) {
    const int r = lcs_sat_rec(me, me2, sm.crec[j], sm.crec2[j]);
    if (r > 0) {
        lcs_atomic_add(&sm.cnt[i], 1);
        lcs_atomic_add(&sm.cnt[j], 1);
    } else if (r < 0) { // within the float32 error of touching: float64 decides (rare)
        bool overflow;
        lcs_push_pair(sm, cap, i, j, &overflow);
        if (overflow) lcs_collide_eval(sm, i, j);
    }
}
LCS_D void lcs_phase_collision_filter(const LcsParams& p, LcsSmem& sm, int s, int tid) {
    const int n = sm.misc[1], cap = LCS_QPER * p.Ap;
    if (tid >= n) return;
    const lcs_f4 me = sm.crec[tid], me2 = sm.crec2[tid];
    const int half = n >> 1;
    // partners tid+1 .. tid+half of the doubled coordinate arrays (no wrap); for even n the pair at distance n/2 is
    // taken by its lower slot only
    const int lo = tid + 1, hi = tid + half - ((2 * half == n && tid >= half) ? 1 : 0);
    const float* sx = sm.soa;
    const float* sy = sm.soa + LCS_SOA_Y(p.Ap);
    // first test, four partners per step: squared centre distance against (own radius + largest radius of the list)^2;
    // survivors get the pair's own radii
    const float rq = me.z + lcs_i2f(sm.misc[14]), thr2 = rq * rq * 1.00001f;
    int nc = 0;
    // the circle test runs over all partners first; the survivors (a handful per thread) are kept in a private
    // list so that the warp does not serialise a SAT evaluation into almost every iteration of this loop
    for (int g = lo & ~3; g <= hi; g += 4) {
        const lcs_f4 X = *(const lcs_f4*)(sx + g), Y = *(const lcs_f4*)(sy + g);
        const float dx0 = X.x - me.x, dx1 = X.y - me.x, dx2 = X.z - me.x, dx3 = X.w - me.x;
        const float dy0 = Y.x - me.y, dy1 = Y.y - me.y, dy2 = Y.z - me.y, dy3 = Y.w - me.y;
        unsigned m = (dx0 * dx0 + dy0 * dy0 <= thr2 ? 1u : 0u) | (dx1 * dx1 + dy1 * dy1 <= thr2 ? 2u : 0u) |
                     (dx2 * dx2 + dy2 * dy2 <= thr2 ? 4u : 0u) | (dx3 * dx3 + dy3 * dy3 <= thr2 ? 8u : 0u);
        if (g < lo) m &= 0xFu << (lo - g);
        if (g + 3 > hi) m &= 0xFu >> (g + 3 - hi);
        while (m) {
            const int r = lcs_ffs0(m);
            m &= m - 1;
            int j = g + r;
            if (j >= n) j -= n;
            const lcs_f4 ot = sm.crec[j];
            const float dx = ot.x - me.x, dy = ot.y - me.y;
            const float rr = ot.z + me.z; // both radii already include their coordinate slack
            if (dx * dx + dy * dy <= rr * rr * 1.00001f) {
                if (nc < LCS_CAND)
                    sm.cand[nc++ * p.Ap + tid] = (uint16_t)j;
                else
                    lcs_collision_narrow(sm, cap, tid, j, me, me2);
            }
        }
    }
    for (int k = 0; k < nc; k++) lcs_collision_narrow(sm, cap, tid, sm.cand[k * p.Ap + tid], me, me2);
}
// Phase E2: float64 narrowphase over the queued (undecided) pairs
LCS_D void lcs_phase_collision_drain(const LcsParams& p, LcsSmem& sm, int s, int tid) {
    const int cap = LCS_QPER * p.Ap;
    const int n = sm.misc[6] < cap ? sm.misc[6] : cap;
    for (int q = tid; q < n; q += p.Ap) {
        const uint32_t e = sm.queue[q];
        lcs_collide_eval(sm, (int)(e >> 16), (int)(e & 0xffff));
    }
}

// ------------------------------------------------------------------ metric kernels (collision / nearest edge)
// They read the snapshot the tick kernel wrote for its ring slot (p.snap): the new list, and per slot the float64
// state x, y, cos h, sin h, length, width.  Nothing they write is read by a later tick (the outputs are metrics,
// manager.py:313-345), so the host lets them run beside the following ticks.

// Collision kernel, load phase: thread = list slot; shared-memory records are SLOT-indexed here (act2 = identity).
// Returns the agent of this thread's slot (-1: none).
LCS_D int lcs_mc_load(const LcsParams& p, LcsSmem& sm, int s, int tid, const LcsThread& t) {
    const int32_t* h = p.snap_hdr + lcs_snap_hdr(p, t.snap, s);
    const int n_new = h[1], n_fin = h[2];
    const size_t base = lcs_snap_base(p, t.snap, s);
    if (tid == 0) {
        sm.misc[0] = h[0];
        sm.misc[1] = n_new;
        sm.misc[2] = 0; // collision sum
        sm.misc[6] = 0; // pair-queue length
    }
    sm.cnt[tid] = 0;
    sm.act2[tid] = tid;
    // the slot's record is loaded whether or not the slot is listed (stale but in bounds beyond the list): the loads
    // go out together with the header's instead of after it
    const int a = p.snap_agent[base + tid] & (LCS_SNAP_MOVED - 1);
    const double* rp = p.snap_rec + 6 * (base + tid);
    const double r[6] = {LCS_LD_STREAM(rp), LCS_LD_STREAM(rp + 1), LCS_LD_STREAM(rp + 2), LCS_LD_STREAM(rp + 3), LCS_LD_STREAM(rp + 4), LCS_LD_STREAM(rp + 5)};
    if (tid < n_fin) p.coll_count[(size_t)s * p.Ap + p.snap_fin[base + tid]] = 0; // left the list: outputs reset
    if (tid >= n_new) return -1;
    lcs_store_ex(sm, tid, r[0], r[1], 0.0, r[2], r[3], r[4], r[5], 0.0);
    lcs_make_crec(p, s, r[0], r[1], r[2], r[3], r[4], r[5], &sm.crec[tid], &sm.crec2[tid]);
    // coordinates once more as doubled float arrays (collision filter), largest bounding radius of the list
    sm.soa[tid] = sm.soa[tid + n_new] = sm.crec[tid].x;
    sm.soa[LCS_SOA_Y(p.Ap) + tid] = sm.soa[LCS_SOA_Y(p.Ap) + tid + n_new] = sm.crec[tid].y;
    lcs_atomic_max(&sm.misc[14], lcs_f2i(sm.crec[tid].z));
    return a;
}
LCS_D void lcs_mc_out(const LcsParams& p, LcsSmem& sm, int s, int tid, int a) {
    if (a < 0) return;
    const int c = sm.cnt[tid];
    p.coll_count[(size_t)s * p.Ap + a] = c;
    if (c) lcs_atomic_add(&sm.misc[2], c);
}
LCS_D void lcs_mc_stats(const LcsParams& p, LcsSmem& sm, int s, int tid, const LcsThread& t) {
    if (tid == 0 && p.mode == 0) {
        p.stats[(size_t)s * LCS_NSTAT + 2] += sm.misc[2];
        if (p.tick_log) p.tick_log[((size_t)t.k * p.S + s) * 4 + 2] = sm.misc[2];
    }
}

// Edge kernel (manager.py:337-345 -> junction.py:387-409): nearest road-edge point and side flag of every listed
// agent whose position changed.  words: [32] ballot words, cnt: [2] = {searches, off-road agents} (shared memory).
// Part 1: flags; the agents that keep last tick's outputs contribute their stored flag to the off-road sum.
LCS_D void lcs_me_flags(const LcsParams& p, int s, int tid, uint32_t* words, int32_t* cnt, const LcsThread& t) {
    const int32_t* h = p.snap_hdr + lcs_snap_hdr(p, t.snap, s);
    const int n_new = h[1], n_fin = h[2];
    const size_t base = lcs_snap_base(p, t.snap, s);
    bool need = false;
    if (tid < n_new) {
        const int v = p.snap_agent[base + tid];
        need = (v & LCS_SNAP_MOVED) != 0;
        // same query as the tick before (parked / waiting agents, exact-duplicate waypoints): the stored nearest
        // point and side flag were computed for exactly this position and stay as they are
        if (!need && p.offroad[(size_t)s * p.Ap + (v & (LCS_SNAP_MOVED - 1))]) lcs_atomic_add(&cnt[1], 1);
    }
    lcs_set_flag_count(words, tid, need, &cnt[0]);
    if (tid < n_fin) {
        const size_t ia = (size_t)s * p.Ap + p.snap_fin[base + tid];
        p.nearest_edge[ia] = -1;
        p.offroad[ia] = 0;
    }
}
// Part 2: the searches run on DENSE lanes: thread k takes the k-th slot that needs one (a warp pays for the longest
// search among its lanes whatever their number)
struct LcsEdgeJob { // one edge search of thread k = the k-th slot that needs one
    double qx, qy;
    size_t ia;
    LcsCellQ q;
    bool on;
};
// begin: query and the header loads of its candidate list
LCS_D void lcs_me_search_begin(const LcsParams& p, int s, int tid, const uint32_t* words, const int32_t* cnt, const LcsThread& t,
                               LcsEdgeJob& job) {
    job.on = tid < cnt[0];
    if (!job.on) return;
    const size_t q = lcs_snap_base(p, t.snap, s) + lcs_select(words, (p.Ap + 31) >> 5, tid);
    const int a = p.snap_agent[q] & (LCS_SNAP_MOVED - 1);
    job.ia = (size_t)s * p.Ap + a;
    job.qx = p.snap_rec[6 * q];
    job.qy = p.snap_rec[6 * q + 1];
    lcs_edge_cell_begin(p, s, job.qx, job.qy, job.q);
}
LCS_D void lcs_me_search_finish(const LcsParams& p, int s, int32_t* cnt, const LcsEdgeJob& job) {
    if (!job.on) return;
    bool off = false;
    int k = lcs_edge_cell_finish(p, s, job.qx, job.qy, job.q, &off);
    if (k == -2) k = lcs_nearest_edge_hint(p, s, job.qx, job.qy, p.edge_hint[job.ia], &off); // outside the lists' grid
    p.nearest_edge[job.ia] = k;
    p.edge_hint[job.ia] = k;
    p.offroad[job.ia] = off ? 1 : 0;
    if (off) lcs_atomic_add(&cnt[1], 1);
}
LCS_D void lcs_me_stats(const LcsParams& p, int s, int tid, const int32_t* cnt, const LcsThread& t) {
    if (tid == 0 && p.mode == 0) {
        p.stats[(size_t)s * LCS_NSTAT + 3] += cnt[1];
        if (p.tick_log) p.tick_log[((size_t)t.k * p.S + s) * 4 + 3] = cnt[1];
    }
}
