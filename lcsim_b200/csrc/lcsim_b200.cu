// lcsim_b200.cu — kernels and C ABI (include/lcsim_b200.h) of the B200-native agent-stepping path.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC
// (lcsim_b200/native.py:build()).  With -DLCSIM_HOSTEMU the same file compiles as plain C++ in which
// "device memory" is host memory and a launch is a loop; tests/hostemu uses that to exercise the
// ABI and the kernel control flow on a machine without a GPU.  The package never loads that build.
#include "../../include/lcsim_b200.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <string>
#include <thread>
#include <vector>

#ifdef LCSIM_HOSTEMU
typedef int cudaError_t;
typedef void* cudaStream_t;
#define cudaSuccess 0
static const char* cudaGetErrorString(cudaError_t) { return "hostemu"; }
static cudaError_t cudaMalloc(void** p, size_t n) { *p = calloc(1, n ? n : 1); return *p ? 0 : 2; }
static cudaError_t cudaFree(void* p) { free(p); return 0; }
enum { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
static cudaError_t cudaMemcpy(void* d, const void* s, size_t n, int) { memcpy(d, s, n); return 0; }
static cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, int, cudaStream_t) { memcpy(d, s, n); return 0; }
static cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return 0; }
static cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { memset(d, v, n); return 0; }
static cudaError_t cudaSetDevice(int) { return 0; }
static cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return 0; }
static cudaError_t cudaDeviceSynchronize() { return 0; }
static cudaError_t cudaGetLastError() { return 0; }
#define __global__
struct alignas(16) float4 { float x, y, z, w; };
#else
#include <cuda_runtime.h>
#endif

#include "lcsim_core.cuh"
#include "lcsim_plan.cuh"
#include "lcsim_loader.hpp"

#ifndef LCS_MC_CTAS
#define LCS_MC_CTAS 8 // resident CTAs per SM the 128-thread collision / edge kernels are compiled for
#endif
#ifndef LCS_ME_CTAS
#define LCS_ME_CTAS 8
#endif
#define LCS_SNAP_MAX 8 // largest depth of the snapshot ring
#ifndef LCS_MIN_CTAS
#define LCS_MIN_CTAS 7 // resident CTAs per SM the 128-thread tick kernel is compiled for (register cap)
#endif

// ------------------------------------------------------------------------------------ kernels

// ---- the engine kernel ------------------------------------------------------------------------------------------
// A launch covers n_ticks ticks of all S scenarios as a queue of work items, processed by a grid of persistent CTAs
// (one thread per active-list slot):
//   tick item   T(k, s): Engine.step for scenario s (engine.py:145-158: update -> check_departure_and_arrival ->
//                        prepare_obs -> clock); leaves a snapshot of the new list in ring slot k % D
//   metric item M(k, s): check_collision_all / check_offroads of that snapshot (manager.py:313-345); nothing a later
//                        tick reads, so it is off the scenario's serial chain and runs whenever a CTA is free
// Ticket order: T(0, 0..S-1), M(0, 0..S-1), T(1, ...), ...  T(k, s) needs T(k-1, s) and M(k-D, s) (its ring slot is
// free again), M(k, s) needs T(k, s) and M(k-1, s): always SMALLER tickets.  Tickets are claimed by running CTAs (atomic counter),
// so the item a CTA waits for was claimed by a CTA that already runs — forward progress holds for any grid size, with
// or without other work on the GPU; nothing depends on the order in which the hardware dispatches blocks.
#define LCS_TICK_PHASES(SYNC, FOR_T)                                                \
    FOR_T lcs_phase_update(p, sm, s, tid, TH);                                      \
    SYNC;                                                                           \
    FOR_T lcs_phase_depart_flags(p, sm, s, tid, TH);                                \
    SYNC;                                                                           \
    FOR_T lcs_phase_depart_place(p, sm, s, tid, TH);                                \
    SYNC;                                                                           \
    FOR_T lcs_phase_depart_fill(p, sm, s, tid, TH);                                 \
    SYNC;                                                                           \
    FOR_T lcs_phase_publish(p, sm, s, tid, TH);                                     \
    if (p.reactive) {                                                               \
        SYNC;                                                                       \
        FOR_T lcs_phase_lead(p, sm, s, tid);                                        \
    }
// AG: per-thread agent of the slot (device: a register; host emulation: an array)
#define LCS_METRIC_PHASES(SYNC, FOR_T, AG)                                          \
    FOR_T { if (tid < 3) sm.misc[12 + tid] = 0; }                                   \
    SYNC;                                                                           \
    FOR_T { AG = lcs_mc_load(p, sm, s, tid, TH); lcs_me_flags(p, s, tid, sm.mask, sm.misc + 12, TH); } \
    SYNC;                                                                           \
    FOR_T { lcs_phase_collision_filter(p, sm, s, tid); LcsEdgeJob job;                                    \
            lcs_me_search_begin(p, s, tid, sm.mask, sm.misc + 12, TH, job); lcs_me_search_finish(p, s, sm.misc + 12, job); } \
    SYNC;                                                                           \
    FOR_T lcs_phase_collision_drain(p, sm, s, tid);                                 \
    SYNC;                                                                           \
    FOR_T lcs_mc_out(p, sm, s, tid, AG);                                            \
    SYNC;                                                                           \
    FOR_T { lcs_mc_stats(p, sm, s, tid, TH); lcs_me_stats(p, s, tid, sm.misc + 12, TH); }

#ifndef LCSIM_HOSTEMU
// waits until scenario's pair of counters (tick items done, metric items done) reached (want_t, want_m): one 8-byte load per poll
LCS_D void lcs_wait_seq(const int32_t* pair, int want_t, int want_m) {
    int vt, vm;
    for (;;) {
        asm volatile("ld.relaxed.gpu.global.v2.s32 {%0, %1}, [%2];" : "=r"(vt), "=r"(vm) : "l"(pair) : "memory");
        if (vt >= want_t && vm >= want_m) break;
        __nanosleep(100);
    }
}

// The two item bodies are separate functions (not inlined into the queue loop: the loop's live state would push
// the 72-register kernel into kilobytes of spills).  CLK: instrumented instantiation (LCSIM_B200_CLOCKS=1), thread 0
// stamps clock64() at phase boundaries into p.dbg[s][*] (tick item: 0..6, metric item: 8..12); used only by
// profiles/phase_clocks.py.
#ifndef LCS_ITEM_INLINE
#define LCS_ITEM_INLINE __forceinline__
#endif
#define STAMP() do { if (CLK) { if (tid == 0) dbg[kk] = clock64(); kk++; } } while (0)
template <int MAXT, bool CLK>
__device__ LCS_ITEM_INLINE void lcs_item_tick(const LcsParams& p, unsigned char* smem_raw, int s, int k) {
    const int tid = threadIdx.x;
    LcsSmem sm;
    lcs_smem_carve(sm, smem_raw, p.Ap);
    LcsThread th;
    th.k = k;
    th.snap = k % p.snap_depth;
    long long* dbg = CLK ? p.dbg + (size_t)s * 16 : nullptr;
    int kk = 0;
    STAMP();
    if (CLK) {
        const int a = lcs_phase_init(p, sm, s, tid, th);
        LcsUpd u;
        LcsScan sc;
        bool whole_row = false;
        u.len = 0;
        if (a >= 0) lcs_update_pre(p, s, a, u);
        __syncthreads();
        STAMP(); // 1: pre done
        if (a >= 0 && u.direct < 0) whole_row = !lcs_scan_first(p, s, a, u.len, u.cursor, u.qx, u.qy, sc);
        lcs_scan_rows_warp(p, s, whole_row, a, u.len, sc);
        __syncthreads();
        STAMP(); // 2: scan done
        if (a >= 0) th.fin = lcs_update_post(p, sm, s, a, u, sc);
    } else {
        if (p.mode == 1) { // reset launch
            lcs_ring_clear(p, s);
            __syncthreads();
        }
        lcs_phase_update(p, sm, s, tid, th);
    }
    __syncthreads();
    STAMP(); // 3: post done
    lcs_phase_depart_flags(p, sm, s, tid, th);
    __syncthreads();
    lcs_phase_depart_place(p, sm, s, tid, th);
    __syncthreads();
    lcs_phase_depart_fill(p, sm, s, tid, th);
    __syncthreads();
    lcs_phase_publish(p, sm, s, tid, th);
    STAMP(); // 4: departures + publish
    if (p.reactive) {
        __syncthreads();
        lcs_phase_lead(p, sm, s, tid);
        if (CLK) __syncthreads();
        STAMP(); // 5: lead search
    }
}
template <int MAXT, bool CLK>
__device__ LCS_ITEM_INLINE void lcs_item_metric(const LcsParams& p, unsigned char* smem_raw, int s, int k) {
    const int tid = threadIdx.x;
    LcsSmem sm;
    lcs_smem_carve(sm, smem_raw, p.Ap);
    LcsThread th;
    th.k = k;
    th.snap = k % p.snap_depth;
    long long* dbg = CLK ? p.dbg + (size_t)s * 16 + 8 : nullptr;
    int kk = 0;
    STAMP(); // 8
    if (tid < 3) sm.misc[12 + tid] = 0; // edge-search counters, largest radius (float bits)
    __syncthreads();
    const int ag = lcs_mc_load(p, sm, s, tid, th);
    lcs_me_flags(p, s, tid, sm.mask, sm.misc + 12, th);
    __syncthreads();
    STAMP(); // 9: load
    lcs_phase_collision_filter(p, sm, s, tid);
    if (CLK) {
        __syncthreads();
        STAMP(); // 10: collision filter
    }
    // (issuing the list-header loads in front of the collision loop was measured: the registers they hold across the
    // loop cost more than the latency they hide, 2.77e9 vs 2.89e9)
    LcsEdgeJob job;
    lcs_me_search_begin(p, s, tid, sm.mask, sm.misc + 12, th, job);
    lcs_me_search_finish(p, s, sm.misc + 12, job);
    __syncthreads();
    STAMP(); // 11 (10 without CLK): edge search
    lcs_phase_collision_drain(p, sm, s, tid);
    __syncthreads();
    lcs_mc_out(p, sm, s, tid, ag);
    __syncthreads();
    lcs_mc_stats(p, sm, s, tid, th);
    lcs_me_stats(p, s, tid, sm.misc + 12, th);
    STAMP(); // 12: drain + out
}
#undef STAMP

template <int MAXT, bool CLK>
__global__ void __launch_bounds__(MAXT, (MAXT <= 128 ? LCS_MIN_CTAS : 1)) lcs_engine_kernel(const LcsParams p) {
    extern __shared__ __align__(128) unsigned char lcs_smem_raw[];
    __shared__ int s_item;
    const int tid = threadIdx.x;
    const bool fused = p.with_metrics && p.fuse; // tick items followed by their metric items
#ifdef LCS_NO_ONLY // experiment build: the split single-tick launches compiled out
    const int only = 0;
#else
    const int only = p.only;
#endif
    const int D = p.snap_depth;
    // one item per CTA; the ticket is claimed when the CTA RUNS (not blockIdx): see the forward-progress argument above
    if (tid == 0) s_item = atomicAdd(p.q_ticket, 1);
    __syncthreads();
    int item = s_item;
    // Scenario groups (multi-tick launches of more scenarios than the device holds CTAs): tickets run group-major —
    // the whole rollout of scenarios [0, G), then of [G, 2G), ... — so that the state a tick touches stays the size of one
    // group (L2-resident) instead of growing with S.  Groups are independent, inside a group the order is the usual one.
    int s0 = 0, Sg = p.S;
    const int mult = p.with_metrics && !fused && !only ? 2 : 1;
    if (p.group > 0 && p.group < p.S && p.n_ticks > 1) {
        const int per_group = p.n_ticks * mult * p.group, n_groups = (p.S + p.group - 1) / p.group;
        int g = item / per_group;
        if (g > n_groups - 1) g = n_groups - 1; // the last group may be smaller: it takes what is left
        item -= g * per_group;
        s0 = g * p.group;
        Sg = p.S - s0 < p.group ? p.S - s0 : p.group;
    }
    const int per_tick = mult * Sg;
    int k = item / per_tick, r = item - k * per_tick;
    bool metric = r >= Sg || only == 2;
    int s = r >= Sg ? r - Sg : r;
    if (p.order && mult == 2 && item >= Sg) {
        // experiment orders (LCSIM_B200_ORDER), metric items one tick behind their tick items: T(0, .), then rounds
        // q = 1 .. n-1 holding T(q, .) and M(q-1, .), then M(n-1, .).  order 1 interleaves a round as T(q, s), M(q-1, s)
        // pairs, order 2 keeps the round as two blocks T(q, .) | M(q-1, .).  In both, an item's dependencies were
        // claimed about 2 S tickets earlier (needs snap_depth >= 2; order 2 wants 3).  Measured at S=1024: order 1 is
        // SLOWER than the plain order (replay 2.76e9 vs 2.90e9, reactive 0.82e9 vs 1.08e9) - see profiles/NOTES_r2.md.
        const int u = item - Sg, q = u / per_tick + 1, v = u - (q - 1) * per_tick;
        if (q < p.n_ticks) {
            if (p.order == 1) {
                s = v >> 1;
                metric = v & 1;
            } else {
                metric = v >= Sg;
                s = metric ? v - Sg : v;
            }
            k = metric ? q - 1 : q;
        } else {
            s = v;
            metric = true;
            k = p.n_ticks - 1;
        }
    }
    s += s0;
    if (p.perm && p.n_ticks > 1) s = p.perm[s];
    long long t_wait = 0;
    if (CLK && tid == 0) t_wait = clock64();
    if (tid == 0 && !only) {
        // the scenario's previous items (claimed earlier, hence running or done): wait for their release, then one
        // acquire fence (it also drops this SM's stale L1 lines); the barrier hands the data to the whole CTA
        // (metric items of a scenario run in order: they update the same per-agent outputs, and an agent that did not
        // move keeps the flag its previous metric item stored)
        if (!metric) {
            if (k > 0) lcs_wait_seq(p.q_seq + 2 * s, k, p.with_metrics ? (fused ? k : k - D + 1) : 0);
        } else {
            lcs_wait_seq(p.q_seq + 2 * s, k + 1, k);
        }
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
        if (CLK) p.dbg[(size_t)s * 16 + (metric ? 15 : 7)] = clock64() - t_wait; // cycles spent waiting for the dependencies
    }
    __syncthreads();
    if (!(p.reset_mask && !p.reset_mask[s])) { // a masked-out scenario: nothing to do, but its items still count as done
        if (!metric) lcs_item_tick<MAXT, CLK>(p, lcs_smem_raw, s, k);
        if (fused) __syncthreads(); // the snapshot this CTA wrote is read back by other threads of it
        if (metric || fused) lcs_item_metric<MAXT, CLK>(p, lcs_smem_raw, s, k);
    }
    __syncthreads(); // every thread's stores precede the release (cumulativity through the barrier)
    if (tid == 0 && !only) {
        if (fused)
            asm volatile("st.release.gpu.global.v2.s32 [%0], {%1, %1};" ::"l"(p.q_seq + 2 * s), "r"(k + 1) : "memory");
        else
            asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p.q_seq + 2 * s + (metric ? 1 : 0)), "r"(k + 1) : "memory");
    }
}
#endif

// Runs p.n_ticks ticks (or the reset) of all scenarios.  grid_cap: CTAs the device holds at once (create()).
static void lcs_launch(const LcsParams& p, int grid_cap, cudaStream_t stream) {
#ifndef LCSIM_HOSTEMU
    const int Ap = p.Ap;
    const size_t smem = lcs_smem_bytes(Ap);
    const long long total = (long long)p.n_ticks * (p.with_metrics && !p.fuse && !p.only ? 2 : 1) * p.S;
    const int nblk = (int)total; // one CTA per item (n_ticks * 2 * S <= 2^31 - 1 is checked by the callers)
    (void)grid_cap;
    cudaMemsetAsync(p.q_ticket, 0, sizeof(int32_t) * (2 + 2 * (size_t)p.S), stream); // ticket and the counter pairs are one block
    if (p.dbg && Ap == 128 && p.mode == 0)
        lcs_engine_kernel<128, true><<<nblk, Ap, smem, stream>>>(p);
    else if (Ap <= 128)
        lcs_engine_kernel<128, false><<<nblk, Ap, smem, stream>>>(p);
    else if (Ap <= 256)
        lcs_engine_kernel<256, false><<<nblk, Ap, smem, stream>>>(p);
    else if (Ap <= 512)
        lcs_engine_kernel<512, false><<<nblk, Ap, smem, stream>>>(p);
    else
        lcs_engine_kernel<1024, false><<<nblk, Ap, smem, stream>>>(p);
#else
    (void)grid_cap;
    const int Ap = p.Ap;
    std::vector<unsigned char> raw(lcs_smem_bytes(Ap) + 16);
    std::vector<LcsThread> ths(Ap);
    std::vector<int> agent_of(Ap);
    for (int k = 0; k < p.n_ticks; k++)
        for (int s = 0; s < p.S; s++) {
            if (p.reset_mask && !p.reset_mask[s]) continue;
            LcsSmem sm;
            lcs_smem_carve(sm, (unsigned char*)(((uintptr_t)raw.data() + 15) & ~(uintptr_t)15), Ap);
#define TH ths[tid]
#define LCS_FOR_T for (int tid = 0; tid < Ap; tid++)
            LCS_FOR_T { ths[tid].k = k; ths[tid].snap = k % p.snap_depth; }
            LCS_TICK_PHASES((void)0, LCS_FOR_T)
            if (p.with_metrics) {
                for (int w = 0; w < 32; w++) sm.mask[w] = 0; // the emulated ballot only sets bits
                LCS_METRIC_PHASES((void)0, LCS_FOR_T, agent_of[tid])
            }
#undef LCS_FOR_T
#undef TH
        }
#endif
}

// ---- set_ref_trajectory: one CTA (64 threads) per (scenario, agent) entry
struct LcsReplanArgs {
    const int32_t *scenario, *agent;
    const float *xy, *vel, *heading;
    int n, Tn;
};
LCS_D void lcs_replan_stale(const LcsParams& p, const LcsReplanArgs& r, int e) {
    const int s = r.scenario[e], a = r.agent[e];
    const size_t ia = (size_t)s * p.Ap + a;
    if (p.status[ia] == LCS_ACTIVE && !p.stale_valid[ia]) {
        double wp[5];
        const bool f = lcs_policy_waypoint(p, ia, p.cursor[ia], p.traj_len[ia], wp);
        for (int k = 0; k < 5; k++) p.stale_wp[5 * ia + k] = wp[k];
        p.stale_f32[ia] = f ? 1 : 0;
        p.stale_valid[ia] = 1;
    }
}
LCS_D void lcs_replan_copy(const LcsParams& p, const LcsReplanArgs& r, int e, int k) {
    const int s = r.scenario[e], a = r.agent[e];
    const size_t ia = (size_t)s * p.Ap + a;
    const int Tn = r.Tn < p.T ? r.Tn : p.T;
    const size_t f = lcs_tf_index(p, s, k, a); // k < Tp
    if (k < Tn) {
        const size_t t = ia * p.T + k, q = (size_t)e * r.Tn + k;
        const double x = (double)r.xy[2 * q], y = (double)r.xy[2 * q + 1];
        p.traj_xy[2 * t] = x;
        p.traj_xy[2 * t + 1] = y;
        p.traj_vel[2 * t] = (double)r.vel[2 * q];
        p.traj_vel[2 * t + 1] = (double)r.vel[2 * q + 1];
        p.traj_h[t] = (double)r.heading[q];
        lcs_f2 v;
        v.x = (float)(x - p.origin[2 * s]);
        v.y = (float)(y - p.origin[2 * s + 1]);
        p.tf_xy[f] = v;
    } else {
        lcs_f2 v;
        v.x = LCS_FAR;
        v.y = LCS_FAR;
        p.tf_xy[f] = v;
    }
    // safe radius of the new point (lcs_scan_window): every point of a re-plan is live
    float rho = 0.f;
    if (k < Tn) {
        const size_t q0 = (size_t)e * r.Tn;
        const double x = (double)r.xy[2 * (q0 + k)], y = (double)r.xy[2 * (q0 + k) + 1];
        double d2 = lcs_inf_d();
        for (int m = 0; m < Tn; m++)
            if (m < k - LCS_SCAN_W || m > k + LCS_SCAN_W) {
                const double dx = (double)r.xy[2 * (q0 + m)] - x, dy = (double)r.xy[2 * (q0 + m) + 1] - y;
                const double dd = dx * dx + dy * dy;
                d2 = dd < d2 ? dd : d2;
            }
        if (d2 == lcs_inf_d()) {
            rho = lcs_inf_f();
        } else {
            const double h = 0.5 * sqrt(d2) * (1.0 - 1e-9);
            rho = (float)h;
            if ((double)rho > h) rho = nextafterf(rho, 0.0f); // round down
        }
    }
    p.tf_rho[f] = rho;
    if (k == 0) {
        p.traj_len[ia] = Tn;
        p.traj_f32[ia] = 1;
        p.traj_blob[ia] = 0;
        p.cursor[ia] = 0;
    }
}
#ifndef LCSIM_HOSTEMU
__global__ void lcs_replan_kernel(const LcsParams p, const LcsReplanArgs r) {
    const int e = blockIdx.x;
    if (threadIdx.x == 0) lcs_replan_stale(p, r, e);
    __syncthreads();
    for (int k = threadIdx.x; k < p.Tp; k += blockDim.x) lcs_replan_copy(p, r, e, k);
}
#endif

// ---- rewards / route for one agent per scenario (thread per scenario)
struct LcsRewardArgs {
    const int32_t* agent; // [S] or null (ADS)
    double coef[LCSIM_B200_N_REWARD_TERMS];
    int has_coef;
    double *reward, *terms, *route;
    uint8_t *terminated, *truncated;
    lcsim_b200_env_out* env; // [S] packed output of env_step, or null
    const double* s_tab0; // [S][Ap][T] np.sum(seg[:k]) of the pristine trajectories
};

// numpy's pairwise summation for n < 128 (np.add.reduce on a contiguous 1-D array): 8 running
// partial sums, combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), then the tail added in order.
template <typename F>
LCS_D F lcs_np_sum_small(const F* a, int n) {
    if (n < 8) {
        F r = 0;
        for (int i = 0; i < n; i++) r = r + a[i]; // additions only: nothing to contract
        return r;
    }
    F r[8];
    int i;
    for (i = 0; i < 8; i++) r[i] = a[i];
    for (i = 8; i < n - (n % 8); i += 8)
        for (int j = 0; j < 8; j++) r[j] = r[j] + a[i + j];
    F res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; i++) res = res + a[i];
    return res;
}

// np.sum(seg[:k]) of a re-planned (float32) trajectory, evaluated in float32 like the reference does
// (polygon.py:287-292 on float32 arrays).  T <= 128 is enforced at create() for this path.
LCS_D double lcs_seg_sum_f32(const double* xy, int n_seg, int k) {
    if (k > n_seg) k = n_seg;
    if (k <= 0) return 0.0;
    float sg[128];
    for (int i = 0; i < k; i++) {
        const float dx = (float)xy[2 * (i + 1)] - (float)xy[2 * i], dy = (float)xy[2 * (i + 1) + 1] - (float)xy[2 * i + 1];
        sg[i] = sqrtf(lcs_fadd(lcs_fmul(dx, dx), lcs_fmul(dy, dy)));
    }
    return (double)lcs_np_sum_small<float>(sg, k);
}

// first-index argmin of sqrt(dx^2+dy^2) in float64 (polygon.py:285)
LCS_D int lcs_argmin_exact(const double* pts, int n, double px, double py) {
    int best = 0;
    double bd = lcs_inf_d();
    for (int k = 0; k < n; k++) {
        const double d = lcs_norm2_axis(lcs_dsub(pts[2 * k], px), lcs_dsub(pts[2 * k + 1], py));
        if (d < bd) { bd = d; best = k; }
    }
    return best;
}

// reference: engine.py:301-357 _compute_rewards, agent.py:468-480 get_route, manager.py:289-311
// check_collision (lane=None agents: all active), agent.py:482-494 is_offroad
// One WARP per scenario on the device (lane = 0 .. 31; the host emulation calls it with nlanes = 1): the two frenet
// scans over the reference trajectory are split over the lanes and merged with shuffles, everything else is
// computed redundantly by all lanes (same addresses: broadcast loads) and stored by lane 0.
LCS_D void lcs_scan_lanes(const LcsParams& p, int s, int a, int len, LcsScan& sc, int lane, int nlanes) {
    const lcs_f8* tf = (const lcs_f8*)(p.tf_xy + lcs_tf_index(p, s, 0, a));
    const int n4 = (len + 3) >> 2;
    for (int k4 = lane; k4 < n4; k4 += nlanes) lcs_scan_quad(sc, lcs_ld32(tf + k4), 4 * k4);
#ifndef LCSIM_HOSTEMU
    if (nlanes > 1) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const float om = __shfl_xor_sync(0xffffffffu, sc.m, d), om2 = __shfl_xor_sync(0xffffffffu, sc.m2, d);
            const int oj = __shfl_xor_sync(0xffffffffu, sc.j, d);
            lcs_edge_merge(sc.m, sc.m2, sc.j, om, om2, oj);
        }
    }
#endif
}
LCS_D void lcs_rewards_one(const LcsParams& p, const LcsRewardArgs& r, int s, int lane, int nlanes, double* stage = nullptr) {
    const int a = r.agent ? r.agent[s] : p.ads_index[s];
    const size_t ia = (size_t)s * p.Ap + a;
    double terms[6] = {0, 0, 0, 0, 0, 0};
    const double* ps = p.pose64 + 4 * ia;
    const double x = ps[0], y = ps[1], h = ps[2];
    const double* xy = p.traj_xy + ia * p.T * 2;
    const int n = p.traj_len[ia];
    const int status = p.status[ia];
    const bool f32 = p.traj_f32[ia] != 0;
    if (p.ring_count[ia] > 1) {
        const int head = p.ring_head[ia];
        const double* p1 = p.ring + (ia * LCS_H + (head + LCS_H - 1) % LCS_H) * 5;
        const double* p0 = p.ring + (ia * LCS_H + (head + LCS_H - 2) % LCS_H) * 5;
        const double cur_v = lcs_norm2_vec(p1[2], p1[3]);
        // frenet_projection index = the same first-index argmin the tick uses for the cursor (filter + exact ties)
        LcsScan sc0, sc1;
        lcs_scan_begin(p, s, p0[0], p0[1], sc0);
        lcs_scan_lanes(p, s, a, n, sc0, lane, nlanes);
        lcs_scan_begin(p, s, p1[0], p1[1], sc1);
        lcs_scan_lanes(p, s, a, n, sc1, lane, nlanes);
        const int i0 = lcs_scan_finish(p, s, a, n, p0[0], p0[1], sc0), i1 = lcs_scan_finish(p, s, a, n, p1[0], p1[1], sc1);
        const double d0 = lcs_norm2_vec(lcs_dsub(p0[0], xy[2 * i0]), lcs_dsub(p0[1], xy[2 * i0 + 1]));
        const double d1 = lcs_norm2_vec(lcs_dsub(p1[0], xy[2 * i1]), lcs_dsub(p1[1], xy[2 * i1 + 1]));
        if (f32) {
            const double s0 = lcs_seg_sum_f32(xy, n - 1, i0), s1 = lcs_seg_sum_f32(xy, n - 1, i1);
            terms[0] = lcs_dsub((double)((float)s1 - (float)s0), lcs_dsub(d1, d0));
            terms[1] = (double)((float)s1 / (float)lcs_seg_sum_f32(xy, n - 1, n - 1));
        } else {
            const double* st = r.s_tab0 + ia * p.T;
            const double s0 = st[i0], s1 = st[i1];
            terms[0] = lcs_dsub(lcs_dsub(s1, s0), lcs_dsub(d1, d0));
            terms[1] = s1 / st[n - 1];
        }
        double steer = lcs_wrap_angle(lcs_dsub(p1[4], p0[4])) / lcs_norm2_vec(lcs_dsub(p1[0], p0[0]), lcs_dsub(p1[1], p0[1]));
        steer = lcs_npclip(steer, -0.3, 0.3);
        const double sm = lcs_dsub(1.0 / cur_v, fabs(steer));
        terms[2] = sm < 0 ? sm : 0.0; // min(0, x): NaN -> 0
    }
    // collision of this agent against every other active agent
    bool collision = false;
    const bool have_metrics = p.with_metrics && status == LCS_ACTIVE; // the tick already evaluated this agent
    if (have_metrics) {
        collision = p.coll_count[ia] > 0;
    } else if (status == LCS_ACTIVE) {
        double ea[8], eo[8];
        ea[0] = x; ea[1] = y; ea[2] = ps[3];
        lcs_sincos(h, &ea[4], &ea[3]);
        ea[5] = p.length[ia]; ea[6] = p.width[ia]; ea[7] = h;
        const float rad_a = 0.5f * sqrtf((float)(ea[5] * ea[5] + ea[6] * ea[6])) * 1.0001f + 1e-3f;
        const int na = p.n_active[s];
        for (int j = 0; j < na && !collision; j++) {
            const int o = p.active[(size_t)s * p.Ap + j];
            if (o == a) continue;
            const size_t io = (size_t)s * p.Ap + o;
            const double* po = p.pose64 + 4 * io;
            const double L = p.length[io], W = p.width[io];
            const float rad_o = 0.5f * sqrtf((float)(L * L + W * W)) * 1.0001f + 1e-3f;
            const double dx = po[0] - x, dy = po[1] - y;
            const double rr = (double)(rad_a + rad_o);
            if (dx * dx + dy * dy > rr * rr * 1.0001) continue; // bounding circles apart: SAT cannot overlap
            eo[0] = po[0]; eo[1] = po[1]; eo[2] = po[3];
            lcs_sincos(po[2], &eo[4], &eo[3]);
            eo[5] = L; eo[6] = W; eo[7] = po[2];
            collision = lcs_collide(ea, eo);
        }
    }
    terms[3] = collision ? -1.0 : 0.0;
    bool off;
    if (have_metrics)
        off = p.offroad[ia] != 0;
    else if (lcs_nearest_edge_cell(p, s, x, y, &off) == -2)
        lcs_nearest_edge_hint(p, s, x, y, p.edge_hint[ia], &off);
    terms[4] = off ? -1.0 : 0.0;
    if (status == LCS_IDLE || status == LCS_FINISHED) {
        const double dis = lcs_norm2_vec(lcs_dsub(xy[2 * (n - 1)], x), lcs_dsub(xy[2 * (n - 1) + 1], y));
        terms[5] = dis < 2.5 ? 10.0 : -5.0;
    }
    const bool finished = status == LCS_IDLE || status == LCS_FINISHED;
    const bool truncated = p.tick[s] >= p.total_steps || finished; // engine.py:197-200 (clock already advanced)
    double rew = 0.0;
    if (r.has_coef)
        for (int k = 0; k < 6; k++)
            if (r.coef[k] != 0.0) rew = lcs_dadd(rew, lcs_dmul(terms[k], r.coef[k]));
    double route[LCS_ROUTE * 2];
    {
        double sn, c;
        lcs_sincos(h, &sn, &c);
        const int cur = p.cursor[ia];
        for (int q = 0; q < LCS_ROUTE; q++) {
            const int k = cur + 1 + q;
            if (k < n) {
                const double dx = lcs_dsub(xy[2 * k], x), dy = lcs_dsub(xy[2 * k + 1], y);
                route[2 * q] = lcs_dot2(dx, c, dy, sn);
                route[2 * q + 1] = lcs_dot2(dx, -sn, dy, c);
            } else {
                route[2 * q] = 0.0;
                route[2 * q + 1] = 0.0;
            }
        }
    }
#ifndef LCSIM_HOSTEMU
    if (r.env && stage) {
        // the 224-byte record leaves the warp as ONE coalesced store (28 lanes x 8 bytes): the record may live in the
        // caller's pinned host memory, where 30 scalar stores of one lane would be 30 small PCIe writes
        static_assert(sizeof(lcsim_b200_env_out) == 28 * sizeof(double), "env_out layout");
        if (lane == 0) {
            lcsim_b200_env_out* o = (lcsim_b200_env_out*)stage;
            o->reward = rew;
            for (int k = 0; k < 6; k++) o->terms[k] = terms[k];
            for (int k = 0; k < LCS_ROUTE * 2; k++) o->route[k / 2][k % 2] = finished ? 0.0 : route[k]; // engine.py:189-192
            o->terminated = (collision || off) ? 1 : 0;
            o->truncated = truncated ? 1 : 0;
            o->finished = finished ? 1 : 0;
            for (int k = 0; k < 5; k++) o->pad[k] = 0;
        }
        __syncwarp();
        if (lane < 28) ((double*)(r.env + s))[lane] = stage[lane];
        __syncwarp();
    }
#endif
    if (lane != 0) return;
    if (r.env && !stage) {
        lcsim_b200_env_out* o = r.env + s;
        o->reward = rew;
        for (int k = 0; k < 6; k++) o->terms[k] = terms[k];
        for (int k = 0; k < LCS_ROUTE * 2; k++) o->route[k / 2][k % 2] = finished ? 0.0 : route[k]; // engine.py:189-192
        o->terminated = (collision || off) ? 1 : 0;
        o->truncated = truncated ? 1 : 0;
        o->finished = finished ? 1 : 0;
    }
    if (r.terms)
        for (int k = 0; k < 6; k++) r.terms[6 * (size_t)s + k] = terms[k];
    if (r.terminated) r.terminated[s] = (collision || off) ? 1 : 0;
    if (r.truncated) r.truncated[s] = truncated ? 1 : 0;
    if (r.reward) r.reward[s] = rew;
    if (r.route)
        for (int k = 0; k < LCS_ROUTE * 2; k++) r.route[(size_t)s * LCS_ROUTE * 2 + k] = route[k];
}
#ifndef LCSIM_HOSTEMU
__global__ void lcs_rewards_kernel(const LcsParams p, const LcsRewardArgs r) {
    __shared__ double stage[4][28]; // one env_out record per warp
    const int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (s < p.S) lcs_rewards_one(p, r, s, threadIdx.x & 31, 32, stage[threadIdx.x >> 5]);
}
// S x 8 int64 -> 8 sums: one block of 1024 threads, a thread reads whole 64-byte rows, warp shuffles, one barrier
__global__ void __launch_bounds__(1024) lcs_stats_kernel(const int64_t* stats, int S, int64_t* out) {
    __shared__ long long part[32][LCS_NSTAT];
    long long loc[LCS_NSTAT];
    for (int k = 0; k < LCS_NSTAT; k++) loc[k] = 0;
    for (int s = threadIdx.x; s < S; s += blockDim.x) {
        const longlong2* row = (const longlong2*)(stats + (size_t)s * LCS_NSTAT);
#pragma unroll
        for (int k = 0; k < LCS_NSTAT / 2; k++) {
            const longlong2 v = row[k];
            loc[2 * k] += v.x;
            loc[2 * k + 1] += v.y;
        }
    }
#pragma unroll
    for (int k = 0; k < LCS_NSTAT; k++)
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) loc[k] += __shfl_xor_sync(0xffffffffu, loc[k], d);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0)
        for (int k = 0; k < LCS_NSTAT; k++) part[w][k] = loc[k];
    __syncthreads();
    if (threadIdx.x < LCS_NSTAT) {
        long long t = 0;
        for (int q = 0; q < (int)(blockDim.x >> 5); q++) t += part[q][threadIdx.x];
        out[threadIdx.x] = t;
    }
}
#endif

// ------------------------------------------------------------------------------------ engine

struct lcsim_b200_engine {
    lcsim_b200_batch_desc desc;
    LcsParams p;
    std::vector<void*> allocs;
    double* s_tab0 = nullptr;
    bool uploaded = false;
    int64_t launches = 0;
    float* poly = nullptr; // [S][P][3] polyline centres (upload_polylines)
    int32_t* n_poly = nullptr;
    int P = 0;
    double* env_act = nullptr; // [S][2] staging of env_step actions
    lcsim_b200_env_out* env_out = nullptr; // [S] staging of env_step outputs
    LcsWpRec* wp_rec = nullptr;
    int64_t cand_entries = 0; // total length of the candidate lists
    uint8_t* host_mask = nullptr; // rollout_host staging
    int32_t* tick_log = nullptr; // per-tick log of the last rollout (a stable address for the captured graphs)
    size_t log_ticks = 0;
    float* lane_pts = nullptr; // [S][N][4] centreline points x, y, theta, 0 (upload_centrelines)
    int32_t *lane_ids = nullptr, *lane_n = nullptr;
    int lane_N = 0;
    float4* plan_edge = nullptr; // [S][E] road-edge points x, y, dir_x, dir_y in float32 (guidance / plan metrics)
    int32_t* plan_edge_id = nullptr; // [S][E]
    float* plan_sum = nullptr; // [S][plan_parts] partial sums of the guidance / metric kernels
    int32_t* plan_cnt = nullptr;
    int plan_parts = 0;
    int device = 0;
    int grid_cap = 1; // CTAs of the engine kernel the device holds at once (occupancy x SMs)
    int env_zerocopy = 1; // LCSIM_B200_ENV_ZEROCOPY
    int env_split = 0; // LCSIM_B200_ENV_SPLIT=1: env_step splits its launch into tick items | metric items when an observation is registered
    bool env_obs_on = false; // set_env_obs: env_step builds this observation behind its tick
    lcsim_b200_obs_desc env_obs;
    void *obs_stream = nullptr, *ev_tick = nullptr, *ev_obs = nullptr; // side stream + fork / join events of that build
    int fuse_mode = 2; // LCSIM_B200_FUSE: 0 = never, 1 = every launch, default = single-tick launches (LcsParams::fuse)
    std::string err;
};

static int lcs_fail(lcsim_b200_engine* e, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (e) e->err = buf;
    return code;
}
#define LCS_CUDA(e, call)                                                                                   \
    do {                                                                                                    \
        cudaError_t _st = (call);                                                                           \
        if (_st != cudaSuccess) return lcs_fail(e, LCSIM_B200_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(_st)); \
    } while (0)

template <typename Tp>
static int lcs_alloc(lcsim_b200_engine* e, Tp** out, size_t count) {
    void* ptr = nullptr;
    cudaError_t st = cudaMalloc(&ptr, std::max<size_t>(count, 1) * sizeof(Tp));
    if (st != cudaSuccess) return lcs_fail(e, LCSIM_B200_ERR_NOMEM, "cudaMalloc(%zu bytes): %s", count * sizeof(Tp), cudaGetErrorString(st));
    st = cudaMemset(ptr, 0, std::max<size_t>(count, 1) * sizeof(Tp));
    if (st != cudaSuccess) return lcs_fail(e, LCSIM_B200_ERR_CUDA, "cudaMemset: %s", cudaGetErrorString(st));
    e->allocs.push_back(ptr);
    *out = (Tp*)ptr;
    return 0;
}
#define LCS_ALLOC(e, field, count)                                            \
    do {                                                                      \
        int _rc = lcs_alloc(e, (decltype(field)*)&(field), (size_t)(count)); \
        if (_rc) return _rc;                                                  \
    } while (0)

// Every entry point runs with the engine's device current and restores the caller's on return (two engines on
// different GPUs in one process, or a caller whose current device is another one).
struct LcsDeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit LcsDeviceGuard(int dev) {
#ifndef LCSIM_HOSTEMU
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) ok = cudaSetDevice(dev) == cudaSuccess;
        else prev = -1;
#endif
    }
    ~LcsDeviceGuard() {
#ifndef LCSIM_HOSTEMU
        if (prev >= 0) cudaSetDevice(prev);
#endif
    }
};
#define LCS_GUARD(e) LcsDeviceGuard _guard((e)->device)

// frees one of the engine's device blocks now (re-uploads replace their buffers instead of piling them up)
template <typename Tp>
static void lcs_release(lcsim_b200_engine* e, Tp*& ptr) {
    if (!ptr) return;
    void* raw = const_cast<void*>(static_cast<const void*>(ptr));
    auto it = std::find(e->allocs.begin(), e->allocs.end(), raw);
    if (it != e->allocs.end()) e->allocs.erase(it);
    cudaFree(raw);
    ptr = nullptr;
}

extern "C" int lcsim_b200_abi_version(void) { return LCSIM_B200_ABI_VERSION; }

extern "C" const char* lcsim_b200_last_error(const lcsim_b200_engine* e) { return e ? e->err.c_str() : "null engine"; }

extern "C" int64_t lcsim_b200_launch_count(const lcsim_b200_engine* e) { return e ? e->launches : 0; }

extern "C" void lcsim_b200_destroy(lcsim_b200_engine* e) {
    if (!e) return;
    for (void* q : e->allocs) cudaFree(q);
#ifndef LCSIM_HOSTEMU
    if (e->ev_tick) cudaEventDestroy((cudaEvent_t)e->ev_tick);
    if (e->ev_obs) cudaEventDestroy((cudaEvent_t)e->ev_obs);
    if (e->obs_stream) cudaStreamDestroy((cudaStream_t)e->obs_stream);
#endif
    delete e;
}

extern "C" int lcsim_b200_create(const lcsim_b200_batch_desc* d, lcsim_b200_engine** out) {
    if (!d || !out) return LCSIM_B200_ERR_ARG;
    *out = nullptr;
    if (d->abi_version != LCSIM_B200_ABI_VERSION) return LCSIM_B200_ERR_ARG;
    if (d->S < 1 || d->A < 1 || d->A > 1024 || d->T < 1 || d->E < 0) return LCSIM_B200_ERR_ARG;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1 || d->device < 0 || d->device >= ndev) return LCSIM_B200_ERR_CUDA;
    LcsDeviceGuard guard(d->device); // the caller's current device is restored on return
    if (!guard.ok) return LCSIM_B200_ERR_CUDA;

    lcsim_b200_engine* e = new lcsim_b200_engine();
    e->desc = *d;
    LcsParams& p = e->p;
    memset(&p, 0, sizeof p);
    p.S = d->S;
    p.A = d->A;
    p.Ap = (d->A + 31) / 32 * 32;
    p.T = d->T;
    p.Tp = (d->T + 3) / 4 * 4;
    p.E = (std::max(d->E, 1) + 3) / 4 * 4; // 32-byte blocks of 4 points
    p.dt = d->interval;
    p.total_steps = d->total_steps;
    p.reactive = d->reactive;
    p.no_static = d->no_static;
    p.allow_new_obj = d->allow_new_obj;
    p.with_metrics = d->with_metrics;
    *out = e; // returned even on failure below so the caller can read last_error and destroy
    e->device = d->device;
    {
        // reactive mode: metric items run one tick behind (order 2 below) over a ring of 3 snapshots: measured 1.110e9
        // vs 1.073e9 agent-steps/s at S=1024 (log-replay: 2.85e9 vs 2.90e9, so it keeps the plain order and 2 snapshots)
        const char* env = getenv("LCSIM_B200_SNAP_DEPTH");
        p.snap_depth = env ? std::min(LCS_SNAP_MAX, std::max(1, atoi(env))) : (p.reactive ? 3 : 2);
        if (!p.with_metrics) p.snap_depth = 1;
    }
    p.n_ticks = 1;
#ifndef LCSIM_HOSTEMU
    {
        int per_sm = 0, n_sm = 0;
        const size_t smem = lcs_smem_bytes(p.Ap);
        cudaError_t st = cudaSuccess;
        if (smem > 48 * 1024) {
            st = cudaFuncSetAttribute(lcs_engine_kernel<1024, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (st == cudaSuccess) st = cudaFuncSetAttribute(lcs_engine_kernel<512, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        }
        if (st == cudaSuccess) {
            if (p.Ap <= 128) st = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lcs_engine_kernel<128, false>, p.Ap, smem);
            else if (p.Ap <= 256) st = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lcs_engine_kernel<256, false>, p.Ap, smem);
            else if (p.Ap <= 512) st = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lcs_engine_kernel<512, false>, p.Ap, smem);
            else st = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lcs_engine_kernel<1024, false>, p.Ap, smem);
        }
        if (st == cudaSuccess) st = cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, d->device);
        if (st != cudaSuccess || per_sm < 1) return lcs_fail(e, LCSIM_B200_ERR_CUDA, "engine kernel cannot be resident (A=%d): %s", d->A, cudaGetErrorString(st));
        e->grid_cap = per_sm * n_sm;
        const char* env = getenv("LCSIM_B200_GRID"); // experiments: CTAs of the persistent grid
        if (env && atoi(env) > 0) e->grid_cap = atoi(env);
    }
#endif
    const size_t S = p.S, SA = S * p.Ap, SAT = SA * p.T, SE = S * p.E;
    LCS_ALLOC(e, p.n_agents, S);
    LCS_ALLOC(e, p.ads_index, S);
    LCS_ALLOC(e, p.origin, S * 2);
    LCS_ALLOC(e, p.eta_pts, S);
    LCS_ALLOC(e, p.depart_tick, SA);
    LCS_ALLOC(e, p.wait_order, SA);
    LCS_ALLOC(e, p.depart_sorted, SA);
    LCS_ALLOC(e, p.dynamic, SA);
    LCS_ALLOC(e, p.length, SA);
    LCS_ALLOC(e, p.width, SA);
    LCS_ALLOC(e, p.home_xy, SA * 2);
    LCS_ALLOC(e, p.traj_len0, SA);
    LCS_ALLOC(e, p.traj_xy0, SAT * 2);
    LCS_ALLOC(e, p.traj_vel0, SAT * 2);
    LCS_ALLOC(e, p.traj_h0, SAT);
    LCS_ALLOC(e, p.tf_xy0, SA * p.Tp);
    LCS_ALLOC(e, p.tf_rho0, SA * p.Tp);
    LCS_ALLOC(e, p.traj_blob0, SA);
    {
        const char* env = getenv("LCSIM_B200_RECORDS"); // =0: experiments without the replay records
        if (!(env && env[0] == '0')) {
            LCS_ALLOC(e, e->wp_rec, SAT);
            p.wp_rec = e->wp_rec;
        }
    }
    LCS_ALLOC(e, p.traj_len, SA);
    LCS_ALLOC(e, p.traj_xy, SAT * 2);
    LCS_ALLOC(e, p.traj_vel, SAT * 2);
    LCS_ALLOC(e, p.traj_h, SAT);
    LCS_ALLOC(e, p.tf_xy, SA * p.Tp);
    LCS_ALLOC(e, p.tf_rho, SA * p.Tp);
    LCS_ALLOC(e, p.traj_blob, SA);
    LCS_ALLOC(e, p.traj_f32, SA);
    LCS_ALLOC(e, p.stale_valid, SA);
    LCS_ALLOC(e, p.stale_f32, SA);
    LCS_ALLOC(e, p.stale_wp, SA * 5);
    LCS_ALLOC(e, p.n_edge, S);
    LCS_ALLOC(e, p.edge_xy, SE * 2);
    LCS_ALLOC(e, p.edge_dir, SE * 2);
    LCS_ALLOC(e, p.grid_geom, S * 4);
    LCS_ALLOC(e, p.grid_dim, S * 2);
    LCS_ALLOC(e, p.ge_xy, SE);
    LCS_ALLOC(e, p.ge_idx, SE);
    LCS_ALLOC(e, p.q_ticket, 2 + 2 * S);
    p.q_seq = p.q_ticket + 2; // 8-byte aligned pairs
    p.fuse = 0;
    {
        const char* om = getenv("LCSIM_B200_ORDER"); // experiments: 0 = plain ticket order of the rollout launch
        p.order = om ? std::min(2, std::max(0, atoi(om))) : (p.reactive ? 2 : 0);
        if (p.snap_depth < 2) p.order = 0; // T(k, s) would wait for M(k-1, s), whose ticket follows it
        // ticket groups: log-replay runs batches beyond one generation of CTAs group by group (S=4096: 2.55e9 -> see
        // profiles/NOTES_r2.md); the reactive mode gains from MORE scenarios per wave (1.14e9 at 1024, 1.30e9 at 4096)
        const char* gm = getenv("LCSIM_B200_GROUP");
        p.group = gm ? std::max(0, atoi(gm)) : (p.reactive ? 0 : e->grid_cap);
    }
    {
        // measured at 512 envs: the split costs a launch more than the overlap returns (device tick 90.9 us split vs
        // 86.4 us fused, e2e 117 vs 112 us), so it is off unless LCSIM_B200_ENV_SPLIT=1
        const char* zc = getenv("LCSIM_B200_ENV_ZEROCOPY");
        if (zc) e->env_zerocopy = atoi(zc);
        const char* sp = getenv("LCSIM_B200_ENV_SPLIT");
        e->env_split = sp && sp[0] == '1';
    }
    {
        const char* fm = getenv("LCSIM_B200_FUSE");
        e->fuse_mode = !fm ? 2 : atoi(fm) ? 1 : 0;
    }
    if (p.with_metrics) {
        const size_t D = p.snap_depth;
        LCS_ALLOC(e, p.snap_hdr, D * S * 4);
        LCS_ALLOC(e, p.snap_agent, D * SA);
        LCS_ALLOC(e, p.snap_rec, D * SA * 6);
        LCS_ALLOC(e, p.snap_fin, D * SA);
    }
    LCS_ALLOC(e, p.tick, S);
    LCS_ALLOC(e, p.n_active, S);
    LCS_ALLOC(e, p.wait_ptr, S);
    LCS_ALLOC(e, p.stats, S * LCS_NSTAT);
    LCS_ALLOC(e, p.active, SA);
    LCS_ALLOC(e, p.status, SA);
    LCS_ALLOC(e, p.cursor, SA);
    LCS_ALLOC(e, p.ring_count, SA);
    LCS_ALLOC(e, p.ring_head, SA);
    LCS_ALLOC(e, p.pose64, SA * 4);
    LCS_ALLOC(e, p.pose32, SA);
    LCS_ALLOC(e, p.ring, SA * LCS_H * 5);
    LCS_ALLOC(e, p.lead_valid, SA);
    LCS_ALLOC(e, p.lead_dis, SA);
    LCS_ALLOC(e, p.lead_v, SA);
    LCS_ALLOC(e, p.coll_count, SA);
    LCS_ALLOC(e, p.nearest_edge, SA);
    LCS_ALLOC(e, p.edge_hint, SA);
    LCS_ALLOC(e, p.home_edge, SA);
    LCS_ALLOC(e, p.offroad, SA);
    LCS_ALLOC(e, e->s_tab0, SAT);
    LCS_ALLOC(e, e->env_act, S * 2);
    LCS_ALLOC(e, e->env_out, S);
    {
        const char* env = getenv("LCSIM_B200_CLOCKS");
        if (env && env[0] == '1') LCS_ALLOC(e, p.dbg, S * 16);
    }
    return 0;
}

// ---- host-side preparation of the static arrays ------------------------------------------------

namespace {

// largest float not above x (x >= 0)
float lcs_round_down_f32(double x) {
    float f = (float)x;
    if ((double)f > x) f = std::nextafterf(f, 0.0f);
    return f;
}

// Allocator of the large zero-initialised staging arrays: calloc hands out untouched zero pages and the vector's
// value-initialisation is a no-op, so a gigabyte of staging costs nothing until the (threaded) per-scenario
// preparation writes it.  Only for trivially copyable element types whose zero value is all-zero bytes.
template <typename Tp>
struct LcsZeroAlloc {
    using value_type = Tp;
    LcsZeroAlloc() = default;
    template <typename U>
    LcsZeroAlloc(const LcsZeroAlloc<U>&) {}
    Tp* allocate(size_t n) {
        void* q = calloc(n ? n : 1, sizeof(Tp));
        if (!q) throw std::bad_alloc();
        return (Tp*)q;
    }
    void deallocate(Tp* q, size_t) { free(q); }
    template <typename U>
    void construct(U*) {} // the bytes are zero already
    template <typename U, typename... Args>
    void construct(U* q, Args&&... args) { ::new ((void*)q) U(std::forward<Args>(args)...); }
    template <typename U>
    bool operator==(const LcsZeroAlloc<U>&) const { return true; }
    template <typename U>
    bool operator!=(const LcsZeroAlloc<U>&) const { return false; }
};
template <typename Tp>
using LcsZeroVec = std::vector<Tp, LcsZeroAlloc<Tp>>;

struct Staging {
    std::vector<int32_t> depart_sorted, depart_tick, wait_order, traj_len, ge_idx, grid_dim, grid_start, home_edge;
    std::vector<uint8_t> dynamic;
    std::vector<double> length, width, home_xy;
    LcsZeroVec<double> traj_xy, traj_vel, traj_h, s_tab, edge_xy, edge_dir;
    std::vector<lcs_f2> tf_xy, ge_xy;
    LcsZeroVec<float> tf_rho;
    std::vector<uint8_t> traj_blob;
    std::vector<float> cl_geom;
    std::vector<int32_t> cl_dim;
    std::vector<std::vector<int32_t>> cl_start, cl_idx; // per scenario
    std::vector<std::vector<lcs_f2>> cl_xy;
    LcsZeroVec<LcsWpRec> wp_rec;
    std::vector<float> grid_geom, eta;
};

// np.sum(seg[:k]) for every k of one trajectory, numpy pairwise order (n < 128 path is exact for any T
// <= 128; longer prefixes fall back to the recursive split below).
double np_sum_rec(const double* a, int64_t n) {
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
    return np_sum_rec(a, n2) + np_sum_rec(a + n2, n - n2);
}

} // namespace

// LCSIM_B200_UPLOAD_TIMES=1: phase times of upload_static on stderr
struct LcsStopwatch {
    bool on;
    std::chrono::steady_clock::time_point t0;
    LcsStopwatch() : on(getenv("LCSIM_B200_UPLOAD_TIMES") != nullptr), t0(std::chrono::steady_clock::now()) {}
    void lap(const char* what) {
        if (!on) return;
        const auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[lcsim_b200 upload] %-28s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};

static void lcs_parallel_range(int n, const std::function<void(int, int)>& fn) {
    int nt = (int)std::thread::hardware_concurrency();
    if (nt < 1) nt = 1;
    if (nt > 32) nt = 32;
    if (nt > n) nt = n;
    if (nt <= 1) {
        fn(0, n);
        return;
    }
    std::vector<std::thread> th;
    for (int t = 0; t < nt; t++) th.emplace_back(fn, (int)((int64_t)n * t / nt), (int)((int64_t)n * (t + 1) / nt));
    for (auto& t : th) t.join();
}

extern "C" int lcsim_b200_upload_static(lcsim_b200_engine* e, const lcsim_b200_static_host* st) {
    if (!e || !st) return LCSIM_B200_ERR_ARG;
    LCS_GUARD(e);
    LcsParams& p = e->p;
    const int S = p.S, A = p.A, Ap = p.Ap, T = p.T, E = e->desc.E;
    const size_t SA = (size_t)S * Ap, SAT = SA * T;
    for (int s = 0; s < S; s++) {
        if (st->n_agents[s] < 0 || st->n_agents[s] > A) return lcs_fail(e, LCSIM_B200_ERR_ARG, "scenario %d: n_agents %d out of [0,%d]", s, st->n_agents[s], A);
        if (st->n_edge[s] < 0 || st->n_edge[s] > E) return lcs_fail(e, LCSIM_B200_ERR_ARG, "scenario %d: n_edge %d out of [0,%d]", s, st->n_edge[s], E);
        if (st->n_agents[s] > 0 && (st->ads_index[s] < 0 || st->ads_index[s] >= st->n_agents[s]))
            return lcs_fail(e, LCSIM_B200_ERR_ARG, "scenario %d: ads_index %d out of range", s, st->ads_index[s]);
        int prev = -1;
        for (int i = 0; i < st->n_agents[s]; i++) {
            const int a = st->wait_order[(size_t)s * A + i];
            if (a < 0 || a >= st->n_agents[s]) return lcs_fail(e, LCSIM_B200_ERR_ARG, "scenario %d: wait_order[%d]=%d out of range", s, i, a);
            const int tl = st->traj_len[(size_t)s * A + a];
            if (tl < 1 || tl > T) return lcs_fail(e, LCSIM_B200_ERR_ARG, "scenario %d agent %d: traj_len %d out of [1,%d]", s, a, tl, T);
            const int dtk = st->depart_tick[(size_t)s * A + a];
            if (dtk < prev) return lcs_fail(e, LCSIM_B200_ERR_ARG, "scenario %d: depart_tick not sorted along wait_order", s);
            prev = dtk;
        }
    }
    LcsStopwatch sw;
    Staging g;
    g.depart_tick.assign(SA, 0x7fffffff);
    g.wait_order.assign(SA, 0);
    g.depart_sorted.assign(SA, 0x7fffffff);
    g.traj_len.assign(SA, 1);
    g.dynamic.assign(SA, 0);
    g.length.assign(SA, 0.0);
    g.width.assign(SA, 0.0);
    g.home_xy.assign(SA * 2, 0.0);
    g.traj_xy.resize(SAT * 2); // zero pages (LcsZeroAlloc), first touched by the threads below
    g.traj_vel.resize(SAT * 2);
    g.traj_h.resize(SAT);
    g.s_tab.resize(SAT);
    if (e->wp_rec) g.wp_rec.resize(SAT);
    lcs_f2 far;
    far.x = LCS_FAR;
    far.y = LCS_FAR;
    g.tf_xy.assign(SA * p.Tp, far);
    g.tf_rho.resize(SA * p.Tp);
    g.traj_blob.assign(SA, 0);
    const size_t Ed = (size_t)p.E;
    g.edge_xy.resize((size_t)S * Ed * 2);
    g.edge_dir.resize((size_t)S * Ed * 2);
    g.ge_xy.assign((size_t)S * Ed, far);
    g.ge_idx.assign((size_t)S * Ed, 0);
    g.grid_geom.assign((size_t)S * 4, 0.f);
    g.grid_dim.assign((size_t)S * 2, 1);
    g.eta.assign(S, 0.f);
    g.home_edge.assign(SA, -1);
    g.cl_geom.assign((size_t)S * 4, 0.f);
    g.cl_dim.assign((size_t)S * 2, 0);
    g.cl_start.resize(S);
    g.cl_idx.resize(S);
    g.cl_xy.resize(S);
    // LCSIM_B200_CAND_CELL: cell of the candidate lists in metres, 0 = no lists (hinted search only).  Measured at S=1024
    // (log-replay / reactive, profiles/NOTES_r2.md): 3 m 2.81e9 / 1.08e9, 2 m 2.95e9 / 1.12e9, 1.5 m 3.02e9 / 1.136e9, 1.25 m
    // 3.04e9 / 1.148e9 (upload 2.8 s instead of 2.2), 1 m 2.32e9 / 1.03e9 (the 192 x 192 cap no longer covers the maps)
    float cl_cell = 1.5f;
    {
        const char* env = getenv("LCSIM_B200_CAND_CELL");
        if (env) cl_cell = (float)atof(env);
    }
    const float cell = e->desc.grid_cell > 0 ? e->desc.grid_cell : 4.0f;
    std::vector<std::vector<int32_t>> cell_starts(S);

    sw.lap("staging alloc");
    std::atomic<long long> dbg_ns[4]; // LCSIM_B200_UPLOAD_TIMES: thread time of the list builder's levels (coarse, 2 x 2, cells)
    for (auto& v : dbg_ns) v = 0;
    lcs_parallel_range(S, [&](int s0, int s1) {
        std::vector<double> seg(T);
        std::vector<char> live;
        std::vector<std::pair<std::pair<double, double>, int>> keyed;
        for (int s = s0; s < s1; s++) {
            const double ox = st->origin[2 * s], oy = st->origin[2 * s + 1];
            double eta = 0.0;
            auto to_f32 = [&](double x, double y) {
                lcs_f2 v;
                v.x = (float)(x - ox);
                v.y = (float)(y - oy);
                const double ex = (double)v.x - (x - ox), ey = (double)v.y - (y - oy);
                eta = std::max(eta, std::sqrt(ex * ex + ey * ey));
                return v;
            };
            for (int a = 0; a < st->n_agents[s]; a++) {
                const size_t src = (size_t)s * A + a, dst = (size_t)s * Ap + a;
                g.depart_tick[dst] = st->depart_tick[src];
                g.dynamic[dst] = st->dynamic[src];
                g.length[dst] = st->length[src];
                g.width[dst] = st->width[src];
                g.home_xy[2 * dst] = st->home_xy[2 * src];
                g.home_xy[2 * dst + 1] = st->home_xy[2 * src + 1];
                const int n = st->traj_len[src];
                g.traj_len[dst] = n;
                const double* xy = st->traj_xy + src * T * 2;
                memcpy(&g.traj_xy[dst * T * 2], xy, sizeof(double) * 2 * n);
                memcpy(&g.traj_vel[dst * T * 2], st->traj_vel + src * T * 2, sizeof(double) * 2 * n);
                memcpy(&g.traj_h[dst * T], st->traj_h + src * T, sizeof(double) * n);
                // a point that repeats an earlier one exactly can never be the FIRST argmin: drop it from
                // the filter copy (the float64 arrays keep it)
                keyed.clear();
                for (int k = 0; k < n; k++) keyed.push_back({{xy[2 * k], xy[2 * k + 1]}, k});
                std::sort(keyed.begin(), keyed.end());
                live.assign(n, 0);
                for (int q = 0; q < n; q++) {
                    const bool dup = q > 0 && keyed[q].first == keyed[q - 1].first;
                    const int k = keyed[q].second;
                    if (!dup) {
                        g.tf_xy[lcs_tf_index(p, s, k, a)] = to_f32(xy[2 * k], xy[2 * k + 1]);
                        live[k] = 1;
                    }
                }
                // safe radii of the live points (lcs_scan_window)
                {
                    float best = 0.f;
                    for (int k = 0; k < n; k++) {
                        if (!live[k]) continue;
                        double d2 = 1e300;
                        for (int m = 0; m < n; m++)
                            if (live[m] && (m < k - LCS_SCAN_W || m > k + LCS_SCAN_W)) {
                                const double dx = xy[2 * m] - xy[2 * k], dy = xy[2 * m + 1] - xy[2 * k + 1];
                                d2 = std::min(d2, dx * dx + dy * dy);
                            }
                        const float rho = d2 > 1e290 ? INFINITY : lcs_round_down_f32(0.5 * std::sqrt(d2) * (1.0 - 1e-9));
                        g.tf_rho[lcs_tf_index(p, s, k, a)] = rho;
                        best = std::max(best, rho);
                    }
                    g.traj_blob[dst] = best < 0.05f ? 1 : 0; // radii of a few cm: a moving query never fits, skip the window
                }
                // Trajectory.s table: s_tab[k] = np.sum(seg[:k]) (polygon.py:287-292)
                for (int k = 0; k + 1 < n; k++) {
                    const double dx = xy[2 * (k + 1)] - xy[2 * k], dy = xy[2 * (k + 1) + 1] - xy[2 * k + 1];
                    seg[k] = std::sqrt(dx * dx + dy * dy);
                }
                for (int k = 0; k < n; k++) g.s_tab[dst * T + k] = np_sum_rec(seg.data(), k);
                if (e->wp_rec) {
                    const double* vel = st->traj_vel + src * T * 2;
                    const double* hd = st->traj_h + src * T;
                    for (int k = 0; k < n; k++) {
                        LcsWpRec& r = g.wp_rec[dst * T + k];
                        r.x = xy[2 * k];
                        r.y = xy[2 * k + 1];
                        r.h = hd[k];
                        const double vx = vel[2 * k], vy = vel[2 * k + 1];
                        const double vx2 = vx * vx, vy2 = vy * vy; // rounded products: no contraction possible below
                        r.v = std::sqrt(vx2 + vy2);
                        r.vn = std::sqrt(std::fma(vy, vy, vx2));
                        r.ch = std::cos(r.h);
                        r.sh = std::sin(r.h);
                        int canon = k;
                        for (int j = 0; j < k; j++) {
                            const double dx = xy[2 * j] - xy[2 * k], dy = xy[2 * j + 1] - xy[2 * k + 1];
                            const double dx2 = dx * dx, dy2 = dy * dy;
                            if (std::sqrt(dx2 + dy2) == 0.0) { canon = j; break; }
                        }
                        r.canon = canon;
                    }
                }
            }
            for (int i = 0; i < A; i++) g.wait_order[(size_t)s * Ap + i] = st->wait_order[(size_t)s * A + i];
            for (int i = 0; i < st->n_agents[s]; i++)
                g.depart_sorted[(size_t)s * Ap + i] = st->depart_tick[(size_t)s * A + st->wait_order[(size_t)s * A + i]];
            // ---- road edges: exact arrays + hash grid over the de-duplicated float32 points
            const int ne = st->n_edge[s];
            const double* exy = st->edge_xy + (size_t)s * E * 2;
            memcpy(&g.edge_xy[(size_t)s * Ed * 2], exy, sizeof(double) * 2 * ne);
            memcpy(&g.edge_dir[(size_t)s * Ed * 2], st->edge_dir + (size_t)s * E * 2, sizeof(double) * 2 * ne);
            keyed.clear();
            for (int k = 0; k < ne; k++) keyed.push_back({{exy[2 * k], exy[2 * k + 1]}, k});
            std::sort(keyed.begin(), keyed.end());
            std::vector<int> keep;
            for (int q = 0; q < ne; q++)
                if (!(q > 0 && keyed[q].first == keyed[q - 1].first)) keep.push_back(keyed[q].second);
            std::sort(keep.begin(), keep.end());
            std::vector<lcs_f2> pts(keep.size());
            float minx = 0, miny = 0, maxx = 0, maxy = 0;
            for (size_t q = 0; q < keep.size(); q++) {
                pts[q] = to_f32(exy[2 * keep[q]], exy[2 * keep[q] + 1]);
                if (q == 0) {
                    minx = maxx = pts[q].x;
                    miny = maxy = pts[q].y;
                } else {
                    minx = std::min(minx, pts[q].x); maxx = std::max(maxx, pts[q].x);
                    miny = std::min(miny, pts[q].y); maxy = std::max(maxy, pts[q].y);
                }
            }
            const float inv = 1.0f / cell;
            const float gx0 = minx, gy0 = miny;
            int W = 1, H = 1;
            if (!keep.empty()) {
                W = (int)floorf((maxx - gx0) * inv) + 1;
                H = (int)floorf((maxy - gy0) * inv) + 1;
            }
            g.grid_geom[4 * s] = gx0; g.grid_geom[4 * s + 1] = gy0; g.grid_geom[4 * s + 2] = inv; g.grid_geom[4 * s + 3] = cell;
            g.grid_dim[2 * s] = W; g.grid_dim[2 * s + 1] = H;
            std::vector<int32_t>& cs = cell_starts[s];
            cs.assign((size_t)W * H + 1, 0);
            std::vector<int> cell_of(keep.size());
            for (size_t q = 0; q < keep.size(); q++) {
                // the same float32 expression the kernel uses to locate a query (lcs_nearest_edge)
                int cx = (int)floorf((pts[q].x - gx0) * inv), cy = (int)floorf((pts[q].y - gy0) * inv);
                cx = std::min(std::max(cx, 0), W - 1);
                cy = std::min(std::max(cy, 0), H - 1);
                cell_of[q] = cy * W + cx;
                cs[cell_of[q] + 1]++;
            }
            for (size_t c = 0; c < (size_t)W * H; c++) cs[c + 1] += cs[c];
            std::vector<int32_t> fill(cs.begin(), cs.end() - 1);
            for (size_t q = 0; q < keep.size(); q++) { // ascending original index inside each cell
                const int pos = fill[cell_of[q]]++;
                g.ge_xy[(size_t)s * Ed + pos] = pts[q];
                g.ge_idx[(size_t)s * Ed + pos] = keep[q];
            }
            g.eta[s] = (float)(eta * 1.0001) + 1e-12f;
            // hint for the first off-road query of every agent: the nearest edge point of its home pose
            // (first index among exact ties; any edge point would be a valid hint, the nearest is the tightest)
            for (int a = 0; a < st->n_agents[s]; a++) {
                const double hx = st->home_xy[2 * ((size_t)s * A + a)], hy = st->home_xy[2 * ((size_t)s * A + a) + 1];
                double bd = 1e300;
                int bi = -1;
                for (int k : keep) {
                    const double dx = exy[2 * k] - hx, dy = exy[2 * k + 1] - hy, d = dx * dx + dy * dy;
                    if (d < bd) { bd = d; bi = k; }
                }
                g.home_edge[(size_t)s * Ap + a] = bi;
            }
            // ---- candidate lists (lcs_nearest_edge_cell): fine grid over the edge points' box + 12 m, at most 192 x 192
            if (cl_cell > 0.f && !keep.empty()) {
                const float mg = 12.0f;
                const float cx0 = minx - mg, cy0 = miny - mg;
                int cw = (int)std::ceil((maxx + mg - cx0) / cl_cell), ch = (int)std::ceil((maxy + mg - cy0) / cl_cell);
                cw = std::min(std::max(cw, 1), 192);
                ch = std::min(std::max(ch, 1), 192);
                const float cinv = 1.0f / cl_cell;
                g.cl_geom[4 * s] = cx0; g.cl_geom[4 * s + 1] = cy0; g.cl_geom[4 * s + 2] = cinv; g.cl_geom[4 * s + 3] = cl_cell;
                g.cl_dim[2 * s] = cw; g.cl_dim[2 * s + 1] = ch;
                const size_t np_ = keep.size();
                std::vector<double> px(np_), py(np_);
                for (size_t q = 0; q < np_; q++) { // local float64 coordinates of the kept points
                    px[q] = exy[2 * keep[q]] - ox;
                    py[q] = exy[2 * keep[q] + 1] - oy;
                }
                std::vector<int32_t>& cst = g.cl_start[s];
                std::vector<int32_t>& cidx = g.cl_idx[s];
                std::vector<lcs_f2>& cxy = g.cl_xy[s];
                cst.assign((size_t)cw * ch + 1, 0);
                // the kernel finds the cell with float32 arithmetic: widen every cell by 2 cm against its rounding
                const double infl = 0.02;
                // K(box) = points with mindist(point, box) <= U(box) + 1 cm, U(box) = min over points of maxdist(point, box).
                // For a box inside a larger one, K(box) is a subset of K(larger) and so is the point that attains
                // U(box): the fine cells are therefore built from the list of their 8 x 8 block, not from all points.
                // Four levels (32 x 32 blocks, 8 x 8 blocks, 2 x 2 blocks, cells); every level keeps its points' coordinates in
                // contiguous arrays, so both passes of a selection are plain streaming loops.
                struct Lvl {
                    std::vector<int> idx;
                    std::vector<double> x, y;
                };
                auto select = [&](double x0, double x1, double y0, double y1, const Lvl& from, Lvl& out, bool want_xy) {
                    const size_t n = from.idx.size();
                    const double *fxp = from.x.data(), *fyp = from.y.data();
                    double U2 = 1e300; // squared distance to the farthest corner, minimised over the points
                    for (size_t k = 0; k < n; k++) {
                        const double fx = std::max(fxp[k] - x0, x1 - fxp[k]), fy = std::max(fyp[k] - y0, y1 - fyp[k]);
                        U2 = std::min(U2, fx * fx + fy * fy);
                    }
                    const double U = std::sqrt(U2) + 0.01, lim = U * U;
                    out.idx.clear();
                    out.x.clear();
                    out.y.clear();
                    for (size_t k = 0; k < n; k++) {
                        const double nx = std::max(std::max(x0 - fxp[k], fxp[k] - x1), 0.0), ny = std::max(std::max(y0 - fyp[k], fyp[k] - y1), 0.0);
                        if (nx * nx + ny * ny <= lim) {
                            out.idx.push_back(from.idx[k]);
                            if (want_xy) { // (the cells' own lists are only read as indices)
                                out.x.push_back(fxp[k]);
                                out.y.push_back(fyp[k]);
                            }
                        }
                    }
                };
                auto box = [&](int bx0, int bx1, int by0, int by1, const Lvl& from, Lvl& out, bool want_xy = true) { // cells [bx0, bx1) x [by0, by1)
                    select((double)cx0 + (double)bx0 * cl_cell - infl, (double)cx0 + (double)bx1 * cl_cell + infl,
                           (double)cy0 + (double)by0 * cl_cell - infl, (double)cy0 + (double)by1 * cl_cell + infl, from, out, want_xy);
                };
                Lvl all, fine;
                all.idx.resize(np_);
                for (size_t q = 0; q < np_; q++) all.idx[q] = (int)q;
                all.x = px;
                all.y = py;
                const int TB = 32, CB = 8, MB = 2; // four levels: 32 x 32 blocks, 8 x 8, 2 x 2, cells (each nested in the one before)
                std::vector<Lvl> top_lists((cw + TB - 1) / TB), row_lists((cw + CB - 1) / CB), mid_lists((cw + MB - 1) / MB);
                cxy.reserve(np_ * 96); // ~5 MB per scenario at 1.5 m cells: no re-allocation on the way
                cidx.reserve(np_ * 96);
                auto tnow = [] { return std::chrono::steady_clock::now(); };
                auto t0 = tnow();
                for (int cyy = 0; cyy < ch && cxy.size() <= ((size_t)4 << 20); cyy++) {
                    if (cyy % CB == 0) { // the coarse blocks of this band of rows
                        if (sw.on) t0 = tnow();
                        if (cyy % TB == 0)
                            for (int tx = 0; tx * TB < cw; tx++) box(tx * TB, std::min(cw, (tx + 1) * TB), cyy, std::min(ch, cyy + TB), all, top_lists[tx]);
                        for (int bx = 0; bx * CB < cw; bx++)
                            box(bx * CB, std::min(cw, (bx + 1) * CB), cyy, std::min(ch, cyy + CB), top_lists[bx * CB / TB], row_lists[bx]);
                        if (sw.on) dbg_ns[0] += std::chrono::duration_cast<std::chrono::nanoseconds>(tnow() - t0).count();
                    }
                    if (sw.on) t0 = tnow();
                    if (cyy % MB == 0) // the 2 x 2 blocks of this pair of rows, each from its coarse block
                        for (int mx = 0; mx * MB < cw; mx++)
                            box(mx * MB, std::min(cw, (mx + 1) * MB), cyy, std::min(ch, cyy + MB), row_lists[mx * MB / CB], mid_lists[mx]);
                    if (sw.on) {
                        dbg_ns[1] += std::chrono::duration_cast<std::chrono::nanoseconds>(tnow() - t0).count();
                        t0 = tnow();
                    }
                    for (int cxx = 0; cxx < cw; cxx++) {
                        box(cxx, cxx + 1, cyy, cyy + 1, mid_lists[cxx / MB], fine, false);
                        cst[(size_t)cyy * cw + cxx] = (int32_t)cxy.size();
                        for (int q : fine.idx) { // ascending original index (keep[] is sorted)
                            cxy.push_back(pts[q]);
                            cidx.push_back(keep[q]);
                        }
                        while (cxy.size() % 4) { // whole 32-byte blocks
                            cxy.push_back(far);
                            cidx.push_back(0);
                        }
                    }
                    if (sw.on) dbg_ns[2] += std::chrono::duration_cast<std::chrono::nanoseconds>(tnow() - t0).count();
                }
                cst[(size_t)cw * ch] = (int32_t)cxy.size();
                if (cxy.size() > (size_t)4 << 20) {
                    // degenerate geometry (e.g. many points equidistant from a region): lists of this scenario would
                    // take > 48 MB — drop them, its searches use the hinted hash grid
                    g.cl_dim[2 * s] = g.cl_dim[2 * s + 1] = 0;
                    std::vector<int32_t>().swap(cst);
                    std::vector<int32_t>().swap(cidx);
                    std::vector<lcs_f2>().swap(cxy);
                }
            }
        }
    });
    if (sw.on)
        fprintf(stderr, "[lcsim_b200 upload] list builder, thread time: 8x8 blocks %.0f ms, 2x2 blocks %.0f ms, cells %.0f ms\n",
                dbg_ns[0] / 1e6, dbg_ns[1] / 1e6, dbg_ns[2] / 1e6);
    sw.lap("per-scenario preparation");
    int GC = 2;
    for (int s = 0; s < S; s++) GC = std::max<int>(GC, (int)cell_starts[s].size());
    g.grid_start.assign((size_t)S * GC, 0);
    for (int s = 0; s < S; s++) {
        std::copy(cell_starts[s].begin(), cell_starts[s].end(), g.grid_start.begin() + (size_t)s * GC);
        for (size_t c = cell_starts[s].size(); c < (size_t)GC; c++) g.grid_start[(size_t)s * GC + c] = cell_starts[s].back();
    }
    // batch-wide bound of the candidate lists (12 bytes per entry): at most half of the free device memory (and
    // LCSIM_B200_CAND_MAX_MB when set); beyond it the largest scenarios fall back to the hinted hash-grid search
    {
        size_t budget = (size_t)64 << 30;
#ifndef LCSIM_HOSTEMU
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) budget = free_b / 2;
#endif
        const char* env = getenv("LCSIM_B200_CAND_MAX_MB");
        if (env) budget = std::min<size_t>(budget, (size_t)atoll(env) << 20);
        size_t total = 0;
        for (int s = 0; s < S; s++) total += g.cl_xy[s].size();
        if (total * 12 > budget) {
            std::vector<int> order(S);
            for (int s = 0; s < S; s++) order[s] = s;
            std::sort(order.begin(), order.end(), [&](int a, int b) { return g.cl_xy[a].size() > g.cl_xy[b].size(); });
            for (int s : order) {
                if (total * 12 <= budget) break;
                total -= g.cl_xy[s].size();
                g.cl_dim[2 * s] = g.cl_dim[2 * s + 1] = 0;
                std::vector<int32_t>().swap(g.cl_start[s]);
                std::vector<int32_t>().swap(g.cl_idx[s]);
                std::vector<lcs_f2>().swap(g.cl_xy[s]);
            }
        }
    }
    // candidate lists: the scenarios' runs go to the device one after the other, straight from the vectors the
    // threads filled (no host-side concatenation: that was a single-threaded 3.5 GB copy at S=1024)
    {
        int nc = 1;
        for (int s = 0; s < S; s++) nc = std::max<int>(nc, (int)g.cl_start[s].size() - 1);
        LcsZeroVec<int32_t> cstart((size_t)S * (nc + 1));
        std::vector<long long> cbase(S, 0);
        size_t total = 0;
        for (int s = 0; s < S; s++) {
            cbase[s] = (long long)total;
            total += g.cl_xy[s].size();
        }
        lcs_parallel_range(S, [&](int s0, int s1) {
            for (int s = s0; s < s1; s++) {
                const std::vector<int32_t>& c = g.cl_start[s];
                int32_t* dst = cstart.data() + (size_t)s * (nc + 1);
                if (c.empty()) continue;
                std::copy(c.begin(), c.end(), dst);
                for (size_t k = c.size(); k < (size_t)nc + 1; k++) dst[k] = c.back();
            }
        });
        sw.lap("candidate lists: offsets");
        p.cl_nc = nc;
        lcs_release(e, p.cl_geom); // a second upload replaces the first one's lists
        lcs_release(e, p.cl_dim);
        lcs_release(e, p.cl_start);
        lcs_release(e, p.cl_base);
        lcs_release(e, p.cl_xy);
        lcs_release(e, p.cl_idx);
        lcs_release(e, p.grid_start);
        LCS_ALLOC(e, p.cl_geom, (size_t)S * 4);
        LCS_ALLOC(e, p.cl_dim, (size_t)S * 2);
        LCS_ALLOC(e, p.cl_start, cstart.size());
        LCS_ALLOC(e, p.cl_base, S);
        {
            // (no memset of the two big arrays: every entry is written below; total == 0 keeps 4 far points)
            void *qx = nullptr, *qi = nullptr;
            const size_t n = std::max<size_t>(total, 4);
            if (cudaMalloc(&qx, n * sizeof(lcs_f2)) != cudaSuccess || cudaMalloc(&qi, n * sizeof(int32_t)) != cudaSuccess) {
                if (qx) cudaFree(qx);
                return lcs_fail(e, LCSIM_B200_ERR_NOMEM, "cudaMalloc(%zu bytes) for the candidate lists failed", n * 12);
            }
            e->allocs.push_back(qx);
            e->allocs.push_back(qi);
            p.cl_xy = (const lcs_f2*)qx;
            p.cl_idx = (const int32_t*)qi;
            if (total == 0) {
                const lcs_f2 f4[4] = {far, far, far, far};
                LCS_CUDA(e, cudaMemcpy(qx, f4, sizeof f4, cudaMemcpyHostToDevice));
                LCS_CUDA(e, cudaMemset(qi, 0, 4 * sizeof(int32_t)));
            }
        }
        LCS_CUDA(e, cudaMemcpy((void*)p.cl_geom, g.cl_geom.data(), g.cl_geom.size() * sizeof(float), cudaMemcpyHostToDevice));
        LCS_CUDA(e, cudaMemcpy((void*)p.cl_dim, g.cl_dim.data(), g.cl_dim.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
        LCS_CUDA(e, cudaMemcpy((void*)p.cl_start, cstart.data(), cstart.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
        LCS_CUDA(e, cudaMemcpy((void*)p.cl_base, cbase.data(), cbase.size() * sizeof(long long), cudaMemcpyHostToDevice));
        for (int s = 0; s < S; s++) {
            const size_t n = g.cl_xy[s].size();
            if (!n) continue;
            LCS_CUDA(e, cudaMemcpyAsync((void*)(p.cl_xy + cbase[s]), g.cl_xy[s].data(), n * sizeof(lcs_f2), cudaMemcpyHostToDevice, 0));
            LCS_CUDA(e, cudaMemcpyAsync((void*)(p.cl_idx + cbase[s]), g.cl_idx[s].data(), n * sizeof(int32_t), cudaMemcpyHostToDevice, 0));
        }
        LCS_CUDA(e, cudaDeviceSynchronize());
        for (int s = 0; s < S; s++) {
            std::vector<lcs_f2>().swap(g.cl_xy[s]);
            std::vector<int32_t>().swap(g.cl_idx[s]);
        }
        e->cand_entries = (int64_t)total;
        if (sw.on) fprintf(stderr, "[lcsim_b200 upload] candidate lists: %.1f MB (%lld entries, cell %.2f m)\n", total * 12.0 / 1e6, (long long)total, cl_cell);
        sw.lap("candidate lists: device");
    }
    p.GC = GC;
    LCS_ALLOC(e, p.grid_start, (size_t)S * GC);

#define LCS_UP(dst, vec) LCS_CUDA(e, cudaMemcpy((void*)(dst), (vec).data(), (vec).size() * sizeof((vec)[0]), cudaMemcpyHostToDevice))
    LCS_CUDA(e, cudaMemcpy((void*)p.n_agents, st->n_agents, sizeof(int32_t) * S, cudaMemcpyHostToDevice));
    LCS_CUDA(e, cudaMemcpy((void*)p.ads_index, st->ads_index, sizeof(int32_t) * S, cudaMemcpyHostToDevice));
    LCS_CUDA(e, cudaMemcpy((void*)p.origin, st->origin, sizeof(double) * 2 * S, cudaMemcpyHostToDevice));
    LCS_CUDA(e, cudaMemcpy((void*)p.n_edge, st->n_edge, sizeof(int32_t) * S, cudaMemcpyHostToDevice));
    LCS_UP(p.eta_pts, g.eta);
    LCS_UP(p.depart_tick, g.depart_tick);
    LCS_UP(p.wait_order, g.wait_order);
    LCS_UP(p.depart_sorted, g.depart_sorted);
    LCS_UP(p.dynamic, g.dynamic);
    LCS_UP(p.length, g.length);
    LCS_UP(p.width, g.width);
    LCS_UP(p.home_xy, g.home_xy);
    LCS_UP(p.traj_len0, g.traj_len);
    LCS_UP(p.traj_xy0, g.traj_xy);
    LCS_UP(p.traj_vel0, g.traj_vel);
    LCS_UP(p.traj_h0, g.traj_h);
    LCS_UP(p.tf_xy0, g.tf_xy);
    LCS_UP(p.tf_rho0, g.tf_rho);
    LCS_UP(p.tf_rho, g.tf_rho);
    LCS_UP(p.traj_blob0, g.traj_blob);
    LCS_UP(p.traj_blob, g.traj_blob);
    LCS_UP(p.traj_len, g.traj_len);
    LCS_UP(p.traj_xy, g.traj_xy);
    LCS_UP(p.traj_vel, g.traj_vel);
    LCS_UP(p.traj_h, g.traj_h);
    LCS_UP(p.tf_xy, g.tf_xy);
    LCS_UP(e->s_tab0, g.s_tab);
    if (e->wp_rec) LCS_UP(e->wp_rec, g.wp_rec);
    LCS_UP(p.edge_xy, g.edge_xy);
    LCS_UP(p.edge_dir, g.edge_dir);
    LCS_UP(p.grid_geom, g.grid_geom);
    LCS_UP(p.grid_dim, g.grid_dim);
    LCS_UP(p.grid_start, g.grid_start);
    LCS_UP(p.ge_xy, g.ge_xy);
    LCS_UP(p.ge_idx, g.ge_idx);
    LCS_UP(p.home_edge, g.home_edge);
#undef LCS_UP
    {
        // the road-edge points as the reference's float32 roadgraph tensors hold them (junction.py:133-150: float64
        // resampled points and get_polyline_dir directions, cast to float32), for the guidance / plan-metric kernels
        std::vector<float4> pe((size_t)S * Ed);
        for (int s = 0; s < S; s++)
            for (int k = 0; k < st->n_edge[s]; k++) {
                const size_t src = ((size_t)s * E + k) * 2, dst = (size_t)s * Ed + k;
                pe[dst].x = (float)st->edge_xy[src]; pe[dst].y = (float)st->edge_xy[src + 1];
                pe[dst].z = (float)st->edge_dir[src]; pe[dst].w = (float)st->edge_dir[src + 1];
            }
        lcs_release(e, e->plan_edge);
        lcs_release(e, e->plan_edge_id);
        LCS_ALLOC(e, e->plan_edge, pe.size());
        LCS_CUDA(e, cudaMemcpy(e->plan_edge, pe.data(), pe.size() * sizeof(float4), cudaMemcpyHostToDevice));
        if (st->edge_poly_id) {
            std::vector<int32_t> ids((size_t)S * Ed, 0);
            for (int s = 0; s < S; s++)
                for (int k = 0; k < st->n_edge[s]; k++) ids[(size_t)s * Ed + k] = st->edge_poly_id[(size_t)s * E + k];
            LCS_ALLOC(e, e->plan_edge_id, ids.size());
            LCS_CUDA(e, cudaMemcpy(e->plan_edge_id, ids.data(), ids.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
        }
    }
    LCS_CUDA(e, cudaMemset(p.traj_f32, 0, SA));
    LCS_CUDA(e, cudaMemset(p.stale_valid, 0, SA));
    {
        // experiment (LCSIM_B200_PERM=1): ticket order of the rollout launches by decreasing weight of the scenarios
        // (agent-steps of their trajectories).  Measured at S=1024: log-replay 2.88e9 vs 2.95e9 without, reactive 1.136e9
        // vs 1.129e9 — the heavy items of a wave start together anyway; off by default
        const char* pm = getenv("LCSIM_B200_PERM");
        lcs_release(e, p.perm);
        if (pm && pm[0] == '1' && S > 1) {
            std::vector<long long> w(S, 0);
            for (int s = 0; s < S; s++)
                for (int a = 0; a < st->n_agents[s]; a++) w[s] += st->traj_len[(size_t)s * A + a];
            std::vector<int32_t> perm(S);
            for (int s = 0; s < S; s++) perm[s] = s;
            std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) { return w[a] > w[b]; });
            LCS_ALLOC(e, p.perm, S);
            LCS_CUDA(e, cudaMemcpy((void*)p.perm, perm.data(), sizeof(int32_t) * S, cudaMemcpyHostToDevice));
        }
    }
    sw.lap("other arrays: device");
    e->uploaded = true;
    int rc = lcsim_b200_reset(e, nullptr, nullptr);
    if (rc) return rc;
    LCS_CUDA(e, cudaDeviceSynchronize());
    return 0;
}

static void lcs_launch_ticks(lcsim_b200_engine* e, LcsParams& q, int n_ticks, cudaStream_t st, int only = 0) {
    q.n_ticks = n_ticks;
    q.only = q.with_metrics ? only : 0;
    // (the phase-clock runs launch tick by tick to time the items of a rollout: they stay split)
    q.fuse = !q.only && q.with_metrics && (e->fuse_mode == 1 || (e->fuse_mode == 2 && n_ticks == 1 && !q.dbg)) ? 1 : 0;
    lcs_launch(q, e->grid_cap, st);
    e->launches += 1;
}

extern "C" int lcsim_b200_reset(lcsim_b200_engine* e, const uint8_t* reset_mask, void* stream) {
    if (!e) return LCSIM_B200_ERR_ARG;
    if (!e->uploaded) return lcs_fail(e, LCSIM_B200_ERR_STATE, "reset before upload_static");
    LCS_GUARD(e);
    LcsParams q = e->p;
    q.mode = 1;
    q.reset_mask = reset_mask;
    q.K = 0;
    lcs_launch_ticks(e, q, 1, (cudaStream_t)stream);
    LCS_CUDA(e, cudaGetLastError());
    return 0;
}

extern "C" int lcsim_b200_step(lcsim_b200_engine* e, const int32_t* act_agent, const int32_t* act_type,
                               const double* act_data, int K, void* stream) {
    return lcsim_b200_step_masked(e, nullptr, act_agent, act_type, act_data, K, stream);
}

extern "C" int lcsim_b200_step_masked(lcsim_b200_engine* e, const uint8_t* step_mask, const int32_t* act_agent,
                                      const int32_t* act_type, const double* act_data, int K, void* stream) {
    if (!e) return LCSIM_B200_ERR_ARG;
    if (!e->uploaded) return lcs_fail(e, LCSIM_B200_ERR_STATE, "step before upload_static");
    if (K < 0 || (K > 0 && (!act_agent || !act_type || !act_data))) return lcs_fail(e, LCSIM_B200_ERR_ARG, "K=%d with null action arrays", K);
    LCS_GUARD(e);
    LcsParams q = e->p;
    q.mode = 0;
    q.reset_mask = step_mask;
    q.act_agent = act_agent;
    q.act_type = act_type;
    q.act_data = act_data;
    q.K = K;
    lcs_launch_ticks(e, q, 1, (cudaStream_t)stream);
    LCS_CUDA(e, cudaGetLastError());
    return 0;
}

// n_ticks steps without external actions in ONE launch.  want_log: per tick and scenario the engine's own buffer
// e->tick_log ([n_ticks][S][4]) receives agents updated, active agents after the tick, collision-count sum, off-road
// count (the last two with_metrics only); log_out (device, may be null) gets a copy.
static int lcs_rollout(lcsim_b200_engine* e, int n_ticks, bool want_log, int32_t* log_out, void* stream) {
    if (!e->uploaded) return lcs_fail(e, LCSIM_B200_ERR_STATE, "rollout before upload_static");
    if (n_ticks == 0) return 0;
    if ((long long)n_ticks * 2 * e->p.S > 0x7fffffffLL) return lcs_fail(e, LCSIM_B200_ERR_ARG, "rollout: n_ticks * 2 * S exceeds the grid limit");
    LcsParams q = e->p;
    q.mode = 0;
    q.K = 0;
    if (want_log && (size_t)n_ticks > e->log_ticks) {
        lcs_release(e, e->tick_log);
        LCS_ALLOC(e, e->tick_log, (size_t)n_ticks * q.S * 4);
        e->log_ticks = n_ticks;
    }
    q.tick_log = want_log ? e->tick_log : nullptr;
    cudaStream_t us = (cudaStream_t)stream;
    if (q.dbg) { // phase clocks: one tick per launch, so that the stamps are those of the last tick
        for (int k = 0; k < n_ticks; k++) lcs_launch_ticks(e, q, 1, us);
    } else {
        lcs_launch_ticks(e, q, n_ticks, us);
    }
    LCS_CUDA(e, cudaGetLastError());
    if (log_out)
        LCS_CUDA(e, cudaMemcpyAsync(log_out, e->tick_log, sizeof(int32_t) * (size_t)n_ticks * q.S * 4, cudaMemcpyDeviceToDevice, us));
    return 0;
}

extern "C" int lcsim_b200_rollout(lcsim_b200_engine* e, int n_ticks, void* stream) {
    if (!e || n_ticks < 0) return LCSIM_B200_ERR_ARG;
    LCS_GUARD(e);
    return lcs_rollout(e, n_ticks, false, nullptr, stream);
}

extern "C" int lcsim_b200_rollout_log(lcsim_b200_engine* e, int n_ticks, int32_t* tick_log, void* stream) {
    if (!e || n_ticks < 0) return LCSIM_B200_ERR_ARG;
    LCS_GUARD(e);
    return lcs_rollout(e, n_ticks, tick_log != nullptr, tick_log, stream);
}

// Host-buffer form of a whole rollout: what a caller of the reference does with `engine.reset(); for _ in range(n):
// engine.step()` and then reads from the engine, for every scenario of the batch, in ONE call.
extern "C" int lcsim_b200_rollout_host(lcsim_b200_engine* e, const uint8_t* reset_mask_host, int n_ticks,
                                       const lcsim_b200_rollout_out* out, int sync, void* stream) {
    if (!e || n_ticks < 0 || !out) return LCSIM_B200_ERR_ARG;
    if (!e->uploaded) return lcs_fail(e, LCSIM_B200_ERR_STATE, "rollout_host before upload_static");
    LCS_GUARD(e);
    cudaStream_t st = (cudaStream_t)stream;
    const LcsParams& p = e->p;
    const size_t S = p.S, SA = S * p.Ap;
    if (!e->host_mask) LCS_ALLOC(e, e->host_mask, S);
    if (reset_mask_host) LCS_CUDA(e, cudaMemcpyAsync(e->host_mask, reset_mask_host, S, cudaMemcpyHostToDevice, st));
    int rc = lcsim_b200_reset(e, reset_mask_host ? e->host_mask : nullptr, stream);
    if (rc) return rc;
    rc = lcs_rollout(e, n_ticks, out->tick_log != nullptr, nullptr, stream);
    if (rc) return rc;
#ifndef LCSIM_HOSTEMU
    if (out->stats) lcs_stats_kernel<<<1, 1024, 0, st>>>(p.stats, p.S, (int64_t*)e->env_act); // env_act doubles as 64-byte scratch
#else
    if (out->stats)
        for (int k = 0; k < LCS_NSTAT; k++) {
            int64_t acc = 0;
            for (size_t s = 0; s < S; s++) acc += p.stats[s * LCS_NSTAT + k];
            ((int64_t*)e->env_act)[k] = acc;
        }
#endif
#define LCS_D2H(dst, src, bytes) \
    if (dst) LCS_CUDA(e, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st))
    LCS_D2H(out->tick_log, e->tick_log, sizeof(int32_t) * (size_t)n_ticks * S * 4);
    LCS_D2H(out->pose32, p.pose32, sizeof(float) * 4 * SA);
    LCS_D2H(out->status, p.status, sizeof(int32_t) * SA);
    LCS_D2H(out->coll_count, p.coll_count, sizeof(int32_t) * SA);
    LCS_D2H(out->nearest_edge, p.nearest_edge, sizeof(int32_t) * SA);
    LCS_D2H(out->offroad, p.offroad, SA);
    LCS_D2H(out->stats, e->env_act, sizeof(int64_t) * LCS_NSTAT);
#undef LCS_D2H
#ifndef LCSIM_HOSTEMU
    if (sync) LCS_CUDA(e, cudaStreamSynchronize(st));
#endif
    return 0;
}

extern "C" int lcsim_b200_set_ref_trajectory(lcsim_b200_engine* e, const int32_t* scenario, const int32_t* agent,
                                             const float* xy, const float* vel, const float* heading, int n, int T,
                                             void* stream) {
    if (!e) return LCSIM_B200_ERR_ARG;
    if (!e->uploaded) return lcs_fail(e, LCSIM_B200_ERR_STATE, "set_ref_trajectory before upload_static");
    LCS_GUARD(e);
    if (n < 0 || T < 1 || (n > 0 && (!scenario || !agent || !xy || !vel || !heading))) return lcs_fail(e, LCSIM_B200_ERR_ARG, "bad re-plan arguments");
    if (T > 128) return lcs_fail(e, LCSIM_B200_ERR_ARG, "re-plan length %d > 128", T);
    if (n == 0) return 0;
    LcsReplanArgs r{scenario, agent, xy, vel, heading, n, T};
#ifndef LCSIM_HOSTEMU
    lcs_replan_kernel<<<n, 64, 0, (cudaStream_t)stream>>>(e->p, r);
    LCS_CUDA(e, cudaGetLastError());
#else
    for (int i = 0; i < n; i++) {
        lcs_replan_stale(e->p, r, i);
        for (int k = 0; k < e->p.Tp; k++) lcs_replan_copy(e->p, r, i, k);
    }
#endif
    return 0;
}

extern "C" int lcsim_b200_rewards(lcsim_b200_engine* e, const int32_t* agent, const double* coef6, double* reward,
                                  double* terms, uint8_t* terminated, uint8_t* truncated, double* route, void* stream) {
    if (!e) return LCSIM_B200_ERR_ARG;
    if (!e->uploaded) return lcs_fail(e, LCSIM_B200_ERR_STATE, "rewards before upload_static");
    LCS_GUARD(e);
    LcsRewardArgs r;
    memset(&r, 0, sizeof r);
    r.agent = agent;
    r.has_coef = coef6 != nullptr;
    if (coef6) memcpy(r.coef, coef6, sizeof r.coef);
    r.reward = reward;
    r.terms = terms;
    r.route = route;
    r.terminated = terminated;
    r.truncated = truncated;
    r.s_tab0 = e->s_tab0;
#ifndef LCSIM_HOSTEMU
    lcs_rewards_kernel<<<(e->p.S + 3) / 4, 128, 0, (cudaStream_t)stream>>>(e->p, r);
    LCS_CUDA(e, cudaGetLastError());
#else
    for (int s = 0; s < e->p.S; s++) lcs_rewards_one(e->p, r, s, 0, 1);
#endif
    return 0;
}

static int lcs_build_obs(lcsim_b200_engine* e, const lcsim_b200_obs_desc* obs, void* stream);

extern "C" int lcsim_b200_set_env_obs(lcsim_b200_engine* e, const lcsim_b200_obs_desc* obs) {
    if (!e) return LCSIM_B200_ERR_ARG;
    if (!e->uploaded) return lcs_fail(e, LCSIM_B200_ERR_STATE, "set_env_obs before upload_static");
    LCS_GUARD(e);
    e->env_obs_on = false;
    if (!obs) return 0;
    if (obs->k_agents < 0 || obs->k_polys < 0 || obs->k_agents > 1024 || obs->k_polys > 1024 || !(obs->radius > 0))
        return lcs_fail(e, LCSIM_B200_ERR_ARG, "set_env_obs: k_agents/k_polys must be in [0,1024], radius > 0");
    if ((obs->nbr_poly || obs->nbr_poly_feat) && !e->poly)
        return lcs_fail(e, LCSIM_B200_ERR_STATE, "set_env_obs: polyline outputs requested before upload_polylines");
#ifndef LCSIM_HOSTEMU
    if (!e->obs_stream) {
        cudaStream_t st2;
        cudaEvent_t a, b;
        LCS_CUDA(e, cudaStreamCreateWithFlags(&st2, cudaStreamNonBlocking));
        e->obs_stream = st2;
        LCS_CUDA(e, cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
        e->ev_tick = a;
        LCS_CUDA(e, cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
        e->ev_obs = b;
    }
#endif
    e->env_obs = *obs;
    e->env_obs_on = true;
    return 0;
}

static int lcs_env_step(lcsim_b200_engine* e, const double* actions, const double* coef6, lcsim_b200_env_out* out, bool host,
                        int sync, void* stream);
extern "C" int lcsim_b200_env_step(lcsim_b200_engine* e, const double* actions_host, const double* coef6,
                                   lcsim_b200_env_out* out_host, int sync, void* stream) {
    if (!e || !out_host) return LCSIM_B200_ERR_ARG;
    if (!e->uploaded) return lcs_fail(e, LCSIM_B200_ERR_STATE, "env_step before upload_static");
    LCS_GUARD(e);
    return lcs_env_step(e, actions_host, coef6, out_host, true, sync, stream);
}
extern "C" int lcsim_b200_env_step_device(lcsim_b200_engine* e, const double* actions_dev, const double* coef6,
                                          lcsim_b200_env_out* out_dev, void* stream) {
    if (!e || !out_dev) return LCSIM_B200_ERR_ARG;
    if (!e->uploaded) return lcs_fail(e, LCSIM_B200_ERR_STATE, "env_step_device before upload_static");
    LCS_GUARD(e);
    return lcs_env_step(e, actions_dev, coef6, out_dev, false, 0, stream);
}
// device alias of a pinned (page-locked, mapped) host buffer, or null: kernels can then read / write the caller's
// buffer across PCIe directly and the staging copy is not needed
static void* lcs_pinned_alias(const void* host) {
#ifndef LCSIM_HOSTEMU
    cudaPointerAttributes at;
    if (host && cudaPointerGetAttributes(&at, host) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer) return at.devicePointer;
    cudaGetLastError();
#endif
    (void)host;
    return nullptr;
}
// host = true: actions / out are host buffers (copied through the engine's staging blocks); false: device buffers, used in place
static int lcs_env_step(lcsim_b200_engine* e, const double* actions_host, const double* coef6, lcsim_b200_env_out* out_host, bool host,
                        int sync, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    LcsParams q = e->p;
    q.mode = 0;
    q.K = 0;
    // pinned host buffers (LCSIM_B200_ENV_ZEROCOPY: 1 = results, 2 = results and actions, 0 = never): the reward kernel
    // writes the records into the caller's buffer / the tick reads the actions from it, no staging copy on the stream
    lcsim_b200_env_out* out_alias = host && e->env_zerocopy >= 1 ? (lcsim_b200_env_out*)lcs_pinned_alias(out_host) : nullptr;
    const double* act_alias = host && actions_host && e->env_zerocopy >= 2 ? (const double*)lcs_pinned_alias(actions_host) : nullptr;
    if (actions_host && host && !act_alias) {
        LCS_CUDA(e, cudaMemcpyAsync(e->env_act, actions_host, sizeof(double) * 2 * q.S, cudaMemcpyHostToDevice, st));
        q.ads_action = e->env_act;
    } else if (actions_host) {
        q.ads_action = act_alias ? act_alias : actions_host;
    }
    LcsRewardArgs r;
    memset(&r, 0, sizeof r);
    r.has_coef = coef6 != nullptr;
    if (coef6) memcpy(r.coef, coef6, sizeof r.coef);
    r.env = !host ? out_host : out_alias ? out_alias : e->env_out;
    r.s_tab0 = e->s_tab0;
#ifndef LCSIM_HOSTEMU
    if (e->env_obs_on && q.with_metrics && !q.dbg && e->env_split) {
        // the observation needs the tick items only (history ring, list, poses): the launch is split into its tick
        // items and its metric items, and the observation build runs beside the metric items and the reward kernel
        lcs_launch_ticks(e, q, 1, st, 1);
        LCS_CUDA(e, cudaEventRecord((cudaEvent_t)e->ev_tick, st));
        LCS_CUDA(e, cudaStreamWaitEvent((cudaStream_t)e->obs_stream, (cudaEvent_t)e->ev_tick, 0));
        int rc = lcs_build_obs(e, &e->env_obs, e->obs_stream);
        if (rc) return rc;
        LCS_CUDA(e, cudaEventRecord((cudaEvent_t)e->ev_obs, (cudaStream_t)e->obs_stream));
        lcs_launch_ticks(e, q, 1, st, 2);
    } else {
        lcs_launch_ticks(e, q, 1, st);
        if (e->env_obs_on) { // beside the reward kernel only
            LCS_CUDA(e, cudaEventRecord((cudaEvent_t)e->ev_tick, st));
            LCS_CUDA(e, cudaStreamWaitEvent((cudaStream_t)e->obs_stream, (cudaEvent_t)e->ev_tick, 0));
            int rc = lcs_build_obs(e, &e->env_obs, e->obs_stream);
            if (rc) return rc;
            LCS_CUDA(e, cudaEventRecord((cudaEvent_t)e->ev_obs, (cudaStream_t)e->obs_stream));
        }
    }
    lcs_rewards_kernel<<<(q.S + 3) / 4, 128, 0, st>>>(e->p, r);
    LCS_CUDA(e, cudaGetLastError());
    if (e->env_obs_on) LCS_CUDA(e, cudaStreamWaitEvent(st, (cudaEvent_t)e->ev_obs, 0));
    if (host && !out_alias) LCS_CUDA(e, cudaMemcpyAsync(out_host, e->env_out, sizeof(lcsim_b200_env_out) * q.S, cudaMemcpyDeviceToHost, st));
    if (sync) LCS_CUDA(e, cudaStreamSynchronize(st));
#else
    lcs_launch_ticks(e, q, 1, st);
    if (e->env_obs_on) {
        int rc = lcs_build_obs(e, &e->env_obs, stream);
        if (rc) return rc;
    }
    for (int s = 0; s < q.S; s++) lcs_rewards_one(e->p, r, s, 0, 1);
    if (host) memcpy(out_host, e->env_out, sizeof(lcsim_b200_env_out) * q.S);
#endif
    return 0;
}

// ---- observation builder -------------------------------------------------------------------
struct LcsObsArgs {
    lcsim_b200_obs_desc d;
    float r2;
    const float* poly; // [S][P][3] x, y, heading (float32, world frame like the reference's tensors)
    const int32_t* n_poly;
    int P;
};
#define LCS_OBS_CAP 1024 // candidates kept per list before the top-k / first-k cut

// motion_diff/modules/utils.py:116-119 wrap_angle on float32 tensors (Python-style modulo)
LCS_D float lcs_wrap_f32(float a) {
    const float pi = 3.14159265358979323846f, two_pi = 6.28318530717958647692f;
    float r = fmodf(a + pi, two_pi);
    if (r != 0.0f && r < 0.0f) r += two_pi;
    return -pi + r;
}
// relation feature of the encoder (agent_encoder.py:217-231,243-255): |d|, angle_between(head_ego, d), wrap(dh)
LCS_D void lcs_rel_feature(float ex, float ey, float eh, float nx, float ny, float nh, float* f) {
    const float rx = nx - ex, ry = ny - ey;
    f[0] = sqrtf(rx * rx + ry * ry);
    const float hx = cosf(eh), hy = sinf(eh);
    f[1] = atan2f(hx * ry - hy * rx, (0.0f + hx * rx) + hy * ry); // the dot is a torch .sum(): accumulated from +0
    f[2] = lcs_wrap_f32(nh - eh);
}
// newest history entry of agent a as the float32 tensors the reference builds (junction.py:319-325)
LCS_D void lcs_last_pose_f32(const LcsParams& p, size_t ia, float* x, float* y, float* h) {
    const int head = p.ring_head[ia];
    const double* r = p.ring + (ia * LCS_H + (head + LCS_H - 1) % LCS_H) * 5;
    *x = (float)r[0];
    *y = (float)r[1];
    *h = (float)r[4];
}
// one cell of the (N,11,6) history stack: list slot `slot`, shift-register position k (thread per cell: consecutive
// threads write consecutive 24-byte cells; one thread per row was measured 1.6x slower)
LCS_D void lcs_history_cell(const LcsParams& p, const LcsObsArgs& o, int s, int slot, int k, float* stage = nullptr) {
    // stage: where the six values go instead of their place in the tensor (the kernel stages a tile in shared memory)
    float* out = stage ? stage : o.d.history + (((size_t)s * p.Ap + slot) * LCS_H + k) * 6;
    const int a = slot < p.n_active[s] ? p.active[(size_t)s * p.Ap + slot] : -1;
    if (k == 0 && o.d.history_agent) o.d.history_agent[(size_t)s * p.Ap + slot] = a;
    if (a < 0) {
        for (int c = 0; c < 6; c++) out[c] = 0.f;
        return;
    }
    const size_t ia = (size_t)s * p.Ap + a;
    const int head = p.ring_head[ia], cnt = p.ring_count[ia];
    const double* r = p.ring + (ia * LCS_H + (head + k) % LCS_H) * 5;
    for (int c = 0; c < 5; c++) out[c] = (float)r[c];
    out[5] = k >= LCS_H - cnt ? 1.f : 0.f;
}

#ifdef LCSIM_HOSTEMU
// Serial statement of one scenario's gather (host emulation only).
static void lcs_obs_scenario_serial(const LcsParams& p, const LcsObsArgs& o, int s) {
    const int e = o.d.ego ? o.d.ego[s] : p.ads_index[s];
    const size_t ie = (size_t)s * p.Ap + e;
    for (int which = 0; which < 2; which++) {
        const int cap = which == 0 ? o.d.k_agents : o.d.k_polys;
        int32_t* idx_out = which == 0 ? o.d.nbr_agent : o.d.nbr_poly;
        float* feat_out = which == 0 ? o.d.nbr_agent_feat : o.d.nbr_poly_feat;
        int32_t* n_out = which == 0 ? o.d.n_nbr_agent : o.d.n_nbr_poly;
        std::vector<std::pair<float, int>> cand;
        std::vector<float> nx, ny, nh;
        float ex = 0, ey = 0, eh = 0;
        if (p.status[ie] == LCS_ACTIVE) {
            lcs_last_pose_f32(p, ie, &ex, &ey, &eh);
            const int n = which == 0 ? p.n_active[s] : (o.poly ? o.n_poly[s] : 0);
            for (int j = 0; j < n; j++) {
                float x, y, h;
                int id;
                if (which == 0) {
                    id = p.active[(size_t)s * p.Ap + j];
                    if (id == e) continue;
                    lcs_last_pose_f32(p, (size_t)s * p.Ap + id, &x, &y, &h);
                } else {
                    id = j;
                    const float* q = o.poly + ((size_t)s * o.P + j) * 3;
                    x = q[0]; y = q[1]; h = q[2];
                }
                const float dx = x - ex, dy = y - ey, d2 = lcs_fadd(lcs_fmul(dx, dx), lcs_fmul(dy, dy)); // unfused, like torch
                if (d2 < o.r2) {
                    cand.push_back({d2, id});
                    nx.push_back(x); ny.push_back(y); nh.push_back(h);
                }
            }
        }
        std::vector<int> order(cand.size());
        for (size_t k = 0; k < order.size(); k++) order[k] = (int)k;
        if (o.d.sort_by_distance)
            // by distance, ties in source order (the device sorts keys of (distance bits, candidate index))
            std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cand[a].first < cand[b].first; });
        if (n_out) n_out[s] = (int)cand.size();
        for (int k = 0; k < cap; k++) {
            const bool have = k < (int)order.size();
            if (idx_out) idx_out[(size_t)s * cap + k] = have ? cand[order[k]].second : -1;
            if (feat_out) {
                float f[3] = {0.f, 0.f, 0.f};
                if (have) lcs_rel_feature(ex, ey, eh, nx[order[k]], ny[order[k]], nh[order[k]], f);
                for (int c = 0; c < 3; c++) feat_out[((size_t)s * cap + k) * 3 + c] = f[c];
            }
        }
    }
    if (o.d.history)
        for (int slot = 0; slot < p.Ap; slot++)
            for (int k = 0; k < LCS_H; k++) lcs_history_cell(p, o, s, slot, k);
}
#endif

#ifndef LCSIM_HOSTEMU
// grid (S, 2 + LCS_OBS_HIST): the agent list (y = 0), the polyline list (y = 1) and LCS_OBS_HIST slices of the history
// stack (y >= 2) of a scenario are independent pieces of work, one CTA each.  Candidates are compacted in index order
// with ballot-word prefix sums (one barrier per 128 candidates: the flag words are double-buffered and every thread
// carries the running count itself).  The k nearest (k <= 128) are SELECTED, not sorted out of everything: a four-pass
// radix select over the distance bits finds the k-th smallest distance (256-bin histograms in shared memory), the
// candidates below it plus the first ties in index order are compacted, and only those k keys (distance bits, candidate
// index) go through a bitonic sort — one key per thread, strides below 32 in registers with warp shuffles.  Sorting all
// m candidates (203 polyline centres on average, 512 keys padded) made the kernel issue-bound: 21 us of its 27.
// Larger k sort everything (the general path).
#define LCS_OBS_HIST 4
__global__ void __launch_bounds__(128) lcs_obs_kernel(const LcsParams p, const LcsObsArgs o) {
    __shared__ float c_d[LCS_OBS_CAP];
    __shared__ int c_id[LCS_OBS_CAP];
    __shared__ unsigned long long keys[LCS_OBS_CAP];
    __shared__ uint32_t words[2][2][4];
    __shared__ int hist[256];
    __shared__ int s_sel[2]; // radix select: chosen bin, elements still wanted inside it
    const int s = blockIdx.x, tid = threadIdx.x;
    if (blockIdx.y >= 2) {
        if (o.d.history) {
            // 128 cells at a time: the values go to a tile in shared memory and leave as whole 16-byte pieces of the
            // (contiguous) output — six 4-byte stores per thread at a stride of 24 bytes were 144 partial-sector writes
            // per warp instead of 24 sectors, and the L2 write rate was the kernel's limit (20 us of 27)
            float* tile = (float*)keys; // [128][6]
            const int cells = p.Ap * LCS_H, part = (cells + LCS_OBS_HIST - 1) / LCS_OBS_HIST, c0 = (blockIdx.y - 2) * part;
            const int c1 = c0 + part < cells ? c0 + part : cells;
            float* base = o.d.history + (size_t)s * cells * 6; // 16-byte aligned: cells * 6 floats is a multiple of 4
            for (int b0 = c0; b0 < c1; b0 += 128) {
                const int c = b0 + tid, nb = c1 - b0 < 128 ? c1 - b0 : 128;
                if (c < c1) lcs_history_cell(p, o, s, c / LCS_H, c % LCS_H, tile + 6 * tid);
                __syncthreads();
                const int f0 = b0 * 6, f1 = f0 + nb * 6; // floats [f0, f1) of the scenario's block
                // head and tail up to the 16-byte boundaries as scalars, the middle as float4
                const int a0 = (f0 + 3) & ~3, a1 = f1 & ~3;
                for (int f = f0 + tid; f < (a0 < f1 ? a0 : f1); f += 128) base[f] = tile[f - f0];
                for (int f = a0 + 4 * tid; f + 3 < a1 + 0 && f < a1; f += 4 * 128) {
                    const float4 v = make_float4(tile[f - f0], tile[f - f0 + 1], tile[f - f0 + 2], tile[f - f0 + 3]);
                    *(float4*)(base + f) = v;
                }
                for (int f = (a1 > a0 ? a1 : a0) + tid; f < f1; f += 128) base[f] = tile[f - f0];
                __syncthreads();
            }
        }
        return;
    }
    const int e = o.d.ego ? o.d.ego[s] : p.ads_index[s];
    const size_t ie = (size_t)s * p.Ap + e;
    const bool ego_active = p.status[ie] == LCS_ACTIVE;
    float ex = 0, ey = 0, eh = 0;
    if (ego_active) lcs_last_pose_f32(p, ie, &ex, &ey, &eh);
    const int which = blockIdx.y;
    const int cap = which == 0 ? o.d.k_agents : o.d.k_polys;
    int32_t* idx_out = which == 0 ? o.d.nbr_agent : o.d.nbr_poly;
    float* feat_out = which == 0 ? o.d.nbr_agent_feat : o.d.nbr_poly_feat;
    int32_t* n_out = which == 0 ? o.d.n_nbr_agent : o.d.n_nbr_poly;
    if (!idx_out && !feat_out && !n_out) return; // nothing asked of this list
    const int n = !ego_active ? 0 : (which == 0 ? p.n_active[s] : (o.poly ? o.n_poly[s] : 0));
    int found = 0; // running count, the same in every thread
    for (int j0 = 0, it = 0; j0 < n; j0 += 128, it ^= 1) {
        const int j = j0 + tid;
        bool hit = false;
        float d2 = 0;
        int id = -1;
        if (j < n) {
            float x = 0, y = 0, h = 0;
            if (which == 0) {
                id = p.active[(size_t)s * p.Ap + j];
                if (id != e) lcs_last_pose_f32(p, (size_t)s * p.Ap + id, &x, &y, &h);
            } else {
                id = j;
                const float* q = o.poly + ((size_t)s * o.P + j) * 3;
                x = q[0]; y = q[1];
            }
            const float dx = x - ex, dy = y - ey;
            d2 = lcs_fadd(lcs_fmul(dx, dx), lcs_fmul(dy, dy)); // unfused, like torch
            hit = !(which == 0 && id == e) && d2 < o.r2;
        }
        lcs_set_flag(words[it][0], tid, hit);
        __syncthreads(); // (the next iteration writes the other buffer; nobody is more than one barrier ahead)
        const int pos = found + lcs_prefix(words[it][0], tid);
        if (hit && pos < LCS_OBS_CAP) {
            c_d[pos] = d2; c_id[pos] = id;
        }
        found += lcs_total(words[it][0], 4);
    }
    const int m = found < LCS_OBS_CAP ? found : LCS_OBS_CAP;
    if (tid == 0 && n_out) n_out[s] = found;
    __syncthreads();
    // output position of candidate c: its index (ascending order), or its rank by (distance, candidate index)
    const bool sorted = o.d.sort_by_distance && m > 1;
    int n_sort = m; // keys to sort (in keys[0 .. n_sort))
    if (sorted && cap <= 128 && m > cap) {
        // ---- radix select of the cap-th smallest distance (d2 >= 0: the bits order like the values)
        uint32_t prefix = 0;
        int want = cap;
        for (int pass = 3; pass >= 0; pass--) {
            for (int b = tid; b < 256; b += 128) hist[b] = 0;
            __syncthreads();
            for (int c = tid; c < m; c += 128) {
                const uint32_t bits = __float_as_uint(c_d[c]);
                if (pass == 3 || (bits >> (8 * (pass + 1))) == prefix) atomicAdd(&hist[(bits >> (8 * pass)) & 255u], 1);
            }
            __syncthreads();
            if (tid < 32) { // lane l owns bins 8 l .. 8 l + 7: the bin in which the running count reaches `want`
                int cnt[8], sum = 0;
#pragma unroll
                for (int q = 0; q < 8; q++) { cnt[q] = hist[8 * tid + q]; sum += cnt[q]; }
                int incl = sum;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int o2 = __shfl_up_sync(0xffffffffu, incl, d);
                    if (tid >= d) incl += o2;
                }
                int before = incl - sum;
                if (before < want && want <= incl) {
#pragma unroll
                    for (int q = 0; q < 8; q++) {
                        if (before < want && want <= before + cnt[q]) { s_sel[0] = 8 * tid + q; s_sel[1] = want - before; }
                        before += cnt[q];
                    }
                }
            }
            __syncthreads();
            prefix = (prefix << 8) | (uint32_t)s_sel[0];
            want = s_sel[1];
        }
        // ---- compaction in candidate (= source index) order: everything below the threshold, and the first `want` ties
        int n_lt = 0, n_eq = 0;
        for (int c0 = 0, it = 0; c0 < m; c0 += 128, it ^= 1) {
            const int c = c0 + tid;
            const uint32_t bits = c < m ? __float_as_uint(c_d[c]) : 0xffffffffu;
            const bool lt = c < m && bits < prefix, eq = c < m && bits == prefix;
            lcs_set_flag(words[it][0], tid, lt);
            lcs_set_flag(words[it][1], tid, eq);
            __syncthreads();
            const int pl = n_lt + lcs_prefix(words[it][0], tid), pe = n_eq + lcs_prefix(words[it][1], tid);
            if (lt || (eq && pe < want)) keys[pl + (pe < want ? pe : want)] = ((unsigned long long)bits << 32) | (unsigned)c;
            n_lt += lcs_total(words[it][0], 4);
            n_eq += lcs_total(words[it][1], 4);
        }
        n_sort = cap;
        __syncthreads();
    } else if (sorted) {
        for (int c = tid; c < m; c += 128) keys[c] = ((unsigned long long)__float_as_uint(c_d[c]) << 32) | (unsigned)c;
        __syncthreads();
    }
    if (sorted) {
        int n2 = 128;
        while (n2 < n_sort) n2 <<= 1; // 128 .. LCS_OBS_CAP = 1024: 1 .. 8 keys per thread, element c = r * 128 + tid
        for (int c = n_sort + tid; c < n2; c += 128) keys[c] = ~0ull;
        __syncthreads();
        for (int k = 2; k <= n2; k <<= 1) {
            int j = k >> 1;
            for (; j >= 32; j >>= 1) { // partner in another warp's range: through shared memory
                for (int c = tid; c < n2; c += 128) {
                    const int q = c ^ j;
                    if (q > c) {
                        const unsigned long long a = keys[c], b = keys[q];
                        const bool up = (c & k) == 0;
                        if ((a > b) == up) { keys[c] = b; keys[q] = a; }
                    }
                }
                __syncthreads();
            }
            for (int c = tid; c < n2; c += 128) { // the remaining strides stay inside the warp: registers + shuffles
                unsigned long long a = keys[c];
                const bool up = (c & k) == 0;
                for (int jj = j; jj > 0; jj >>= 1) {
                    const unsigned long long b = __shfl_xor_sync(0xffffffffu, a, jj);
                    const bool take_min = ((c & jj) == 0) == up;
                    a = take_min ? (a < b ? a : b) : (a > b ? a : b);
                }
                keys[c] = a;
            }
            __syncthreads();
        }
    }
    for (int r = tid; r < m && r < cap; r += 128) {
        const int c = sorted ? (int)(keys[r] & 0xffffffffu) : r; // candidate at output position r
        const int id = c_id[c];
        if (idx_out) idx_out[(size_t)s * cap + r] = id;
        if (feat_out) {
            float x, y, h, f[3];
            if (which == 0) {
                lcs_last_pose_f32(p, (size_t)s * p.Ap + id, &x, &y, &h);
            } else {
                const float* q = o.poly + ((size_t)s * o.P + id) * 3;
                x = q[0]; y = q[1]; h = q[2];
            }
            lcs_rel_feature(ex, ey, eh, x, y, h, f);
            for (int k = 0; k < 3; k++) feat_out[((size_t)s * cap + r) * 3 + k] = f[k];
        }
    }
    for (int k = m + tid; k < cap; k += 128) {
        if (idx_out) idx_out[(size_t)s * cap + k] = -1;
        if (feat_out)
            for (int c = 0; c < 3; c++) feat_out[((size_t)s * cap + k) * 3 + c] = 0.f;
    }
}
#endif

extern "C" int lcsim_b200_upload_polylines(lcsim_b200_engine* e, const int32_t* n_poly, const double* poly_xyh, int P) {
    if (!e || !n_poly || !poly_xyh || P < 1) return LCSIM_B200_ERR_ARG;
    LCS_GUARD(e);
    const int S = e->p.S;
    std::vector<float> f((size_t)S * P * 3, 0.f);
    for (int s = 0; s < S; s++) {
        if (n_poly[s] < 0 || n_poly[s] > P) return lcs_fail(e, LCSIM_B200_ERR_ARG, "scenario %d: n_poly %d out of [0,%d]", s, n_poly[s], P);
        for (size_t k = 0; k < (size_t)n_poly[s] * 3; k++) f[(size_t)s * P * 3 + k] = (float)poly_xyh[(size_t)s * P * 3 + k];
    }
    lcs_release(e, e->poly); // a second upload replaces the first
    lcs_release(e, e->n_poly);
    LCS_ALLOC(e, e->poly, (size_t)S * P * 3);
    LCS_ALLOC(e, e->n_poly, S);
    LCS_CUDA(e, cudaMemcpy(e->poly, f.data(), f.size() * sizeof(float), cudaMemcpyHostToDevice));
    LCS_CUDA(e, cudaMemcpy(e->n_poly, n_poly, sizeof(int32_t) * S, cudaMemcpyHostToDevice));
    e->P = P;
    return 0;
}

extern "C" int lcsim_b200_build_obs(lcsim_b200_engine* e, const lcsim_b200_obs_desc* obs, void* stream) {
    if (!e || !obs) return LCSIM_B200_ERR_ARG;
    if (!e->uploaded) return lcs_fail(e, LCSIM_B200_ERR_STATE, "build_obs before upload_static");
    LCS_GUARD(e);
    return lcs_build_obs(e, obs, stream);
}
static int lcs_build_obs(lcsim_b200_engine* e, const lcsim_b200_obs_desc* obs, void* stream) {
    if (obs->k_agents < 0 || obs->k_polys < 0 || obs->k_agents > LCS_OBS_CAP || obs->k_polys > LCS_OBS_CAP || !(obs->radius > 0))
        return lcs_fail(e, LCSIM_B200_ERR_ARG, "build_obs: k_agents/k_polys must be in [0,%d], radius > 0", LCS_OBS_CAP);
    if ((obs->nbr_poly || obs->nbr_poly_feat) && !e->poly)
        return lcs_fail(e, LCSIM_B200_ERR_STATE, "build_obs: polyline outputs requested before upload_polylines");
    LcsObsArgs o;
    o.d = *obs;
    o.r2 = (float)obs->radius * (float)obs->radius;
    o.poly = e->poly;
    o.n_poly = e->n_poly;
    o.P = e->P;
#ifndef LCSIM_HOSTEMU
    lcs_obs_kernel<<<dim3(e->p.S, 2 + LCS_OBS_HIST), 128, 0, (cudaStream_t)stream>>>(e->p, o);
    LCS_CUDA(e, cudaGetLastError());
#else
    for (int s = 0; s < e->p.S; s++) lcs_obs_scenario_serial(e->p, o, s);
#endif
    return 0;
}

// ---- encoder inputs for all agents: lcsim_encoder.cuh -------------------------------------------------
#include "lcsim_encoder.cuh"
extern "C" int lcsim_b200_build_encoder_inputs(lcsim_b200_engine* e, const lcsim_b200_encoder_desc* d, void* stream) {
    if (!e || !d) return LCSIM_B200_ERR_ARG;
    if (!e->uploaded) return lcs_fail(e, LCSIM_B200_ERR_STATE, "build_encoder_inputs before upload_static");
    LCS_GUARD(e);
    if (d->cap_a2a < 0 || d->cap_pl2a < 0 || !(d->a2a_radius > 0) || !(d->pl2a_radius > 0) || d->time_span < 0)
        return lcs_fail(e, LCSIM_B200_ERR_ARG, "build_encoder_inputs: capacities >= 0, radii > 0, time_span >= 0");
    if ((d->pl2a_offset || d->pl2a_agent || d->pl2a_feat || d->n_pl2a) && !e->poly)
        return lcs_fail(e, LCSIM_B200_ERR_STATE, "build_encoder_inputs: polyline outputs requested before upload_polylines");
    LceArgs o;
    o.d = *d;
    o.r2_a2a = (float)d->a2a_radius * (float)d->a2a_radius;
    o.r2_pl2a = (float)d->pl2a_radius * (float)d->pl2a_radius;
    o.poly = e->poly;
    o.n_poly = e->n_poly;
    o.P = e->poly ? e->P : 0;
#ifndef LCSIM_HOSTEMU
    lce_encoder_kernel<<<e->p.S, 128, sizeof(float) * 5 * e->p.Ap, (cudaStream_t)stream>>>(e->p, o);
    LCS_CUDA(e, cudaGetLastError());
#else
    for (int s = 0; s < e->p.S; s++) lce_scenario_serial(e->p, o, s);
#endif
    return 0;
}

// ---- lane-centreline projection -----------------------------------------------------------
// The reference never projects a WOMD agent onto a lane inside the engine (agents carry lane=None, agent.py:133-139);
// the lane-centreline projections live in motion_diff: OnLane.loss (modules/guidance.py:100-133: nearest centreline
// point with |wrap(dheading)| < 0.2 rad and distance < 5 m) and compute_lane_heading_diff (metrics/roadgraph.py:
// 91-157: among the 64 nearest centreline points by squared distance, the smallest |wrap(dheading)|).  Both work on
// the float32 tensors Junction.roadgraph_points (junction.py:186-253) restricted to centreline types 1|2, i.e. the
// 0.5 m-resampled lane centrelines in lane order.  One kernel produces, for every ACTIVE agent (query = its float32
// world position and heading, like data["agent"]["xyz"/"heading"]): the first-index argmin point, that point's
// polyline id (= the agent's lane), the distance, the OnLane gated distance and the top-K heading difference.
struct LcsLaneArgs {
    lcsim_b200_lane_desc d;
    const float4* pts; // [S][N] x, y, theta = atan2(dir_y, dir_x) (float32, computed at upload), unused
    const int32_t* ids; // [S][N]
    const int32_t* n_pts; // [S]
    int N, K;
};
#define LCS_LANE_TILE 1024
#define LCS_LANE_KMAX 64

LCS_D float lcs_lane_d2(float qx, float qy, float px, float py) {
    const float dx = qx - px, dy = qy - py;
    return lcs_fadd(lcs_fmul(dx, dx), lcs_fmul(dy, dy)); // torch: sum((q - c) ** 2, -1), unfused
}
// serial statement of one query (host emulation, and the definition the kernel is tested against)
LCS_D void lcs_lane_query_serial(const LcsLaneArgs& o, int s, float qx, float qy, float qh, int* best_i, float* best_d2,
                                 float* on_d2, float* hdiff, float* heap_d, int* heap_i) {
    const float4* P = o.pts + (size_t)s * o.N;
    const int n = o.n_pts[s], K = o.K;
    *best_i = -1;
    *best_d2 = lcs_inf_f();
    *on_d2 = lcs_inf_f();
    int hn = 0;
    for (int k = 0; k < n; k++) {
        const float d2 = lcs_lane_d2(qx, qy, P[k].x, P[k].y);
        if (d2 < *best_d2) { *best_d2 = d2; *best_i = k; }
        if (d2 < *on_d2 && sqrtf(d2) < 5.0f && fabsf(lcs_wrap_f32(qh - P[k].z)) < 0.2f) *on_d2 = d2;
        if (K > 0) { // bounded set of the K smallest (d2, index): replace the current maximum
            if (hn < K) {
                heap_d[hn] = d2; heap_i[hn] = k; hn++;
            } else {
                int m = 0;
                for (int q = 1; q < K; q++)
                    if (heap_d[q] > heap_d[m] || (heap_d[q] == heap_d[m] && heap_i[q] > heap_i[m])) m = q;
                if (d2 < heap_d[m]) { heap_d[m] = d2; heap_i[m] = k; }
            }
        }
    }
    *hdiff = -1.0f; // fewer than K centreline points: the reference skips the batch (roadgraph.py:144-145)
    if (K > 0 && hn == K) {
        float best = lcs_inf_f();
        for (int q = 0; q < K; q++) best = fminf(best, fabsf(lcs_wrap_f32(qh - P[heap_i[q]].z)));
        *hdiff = best;
    }
}
LCS_D void lcs_lane_store(const LcsParams& p, const LcsLaneArgs& o, int s, int a, int bi, float bd2, float on_d2, float hd) {
    const size_t ia = (size_t)s * p.Ap + a;
    if (o.d.lane_pt) o.d.lane_pt[ia] = bi;
    if (o.d.lane_id) o.d.lane_id[ia] = bi >= 0 ? o.ids[(size_t)s * o.N + bi] : -1;
    if (o.d.lane_dist) o.d.lane_dist[ia] = bi >= 0 ? sqrtf(bd2) : lcs_inf_f();
    if (o.d.onlane_dist) o.d.onlane_dist[ia] = sqrtf(on_d2);
    if (o.d.heading_diff) o.d.heading_diff[ia] = hd;
}
LCS_D void lcs_lane_store_inactive(const LcsParams& p, const LcsLaneArgs& o, int s, int a) {
    const size_t ia = (size_t)s * p.Ap + a;
    if (o.d.lane_pt) o.d.lane_pt[ia] = -1;
    if (o.d.lane_id) o.d.lane_id[ia] = -1;
    if (o.d.lane_dist) o.d.lane_dist[ia] = lcs_inf_f();
    if (o.d.onlane_dist) o.d.onlane_dist[ia] = lcs_inf_f();
    if (o.d.heading_diff) o.d.heading_diff[ia] = -1.0f;
}

#ifndef LCSIM_HOSTEMU
// mbarrier / bulk-copy primitives (sm_90+): one thread arms the barrier with the byte count and issues a 1-D
// cp.async.bulk global -> shared; the copy engine completes the transaction on the barrier; everybody waits on the parity
LCS_D uint32_t lcs_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
LCS_D void lcs_mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(lcs_smem_u32(bar)), "r"(count));
}
LCS_D void lcs_bulk_load(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(lcs_smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(lcs_smem_u32(dst_smem)),
                 "l"(src), "r"(bytes), "r"(lcs_smem_u32(bar))
                 : "memory");
}
LCS_D void lcs_mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n.reg .pred p;\nLCS_WAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@!p bra LCS_WAIT_%=;\n}\n" ::"r"(lcs_smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// sift-down of entry r in the max-heap of column `tid` ([K][128] layout), ordered by (distance, index)
LCS_D void lcs_heap_sift(float* hd, int* hi, int tid, int K, int r) {
    const float d = hd[r * 128 + tid];
    const int ix = hi[r * 128 + tid];
    for (;;) {
        int c = 2 * r + 1;
        if (c >= K) break;
        float dc = hd[c * 128 + tid];
        int ic = hi[c * 128 + tid];
        if (c + 1 < K) {
            const float d2 = hd[(c + 1) * 128 + tid];
            const int i2 = hi[(c + 1) * 128 + tid];
            if (d2 > dc || (d2 == dc && i2 > ic)) { c++; dc = d2; ic = i2; }
        }
        if (!(dc > d || (dc == d && ic > ix))) break;
        hd[r * 128 + tid] = dc; hi[r * 128 + tid] = ic;
        r = c;
    }
    hd[r * 128 + tid] = d; hi[r * 128 + tid] = ix;
}

// One CTA per scenario, thread = active-list slot; the centreline points stream through shared memory in tiles of
// LCS_LANE_TILE, every thread tests its query against the tile.  BULK: the tiles are moved by the copy engine
// (cp.async.bulk + mbarrier, two buffers: tile i+1 lands while tile i is scanned, no thread touches the data on its
// way in); otherwise by coalesced float4 loads and stores of the CTA (LCSIM_B200_LANE_BULK=0).  The K smallest
// distances of a query live in a shared-memory column [K][threads] (conflict-free) organised as a max-heap.
template <bool BULK>
__global__ void __launch_bounds__(128) lcs_lane_kernel(const LcsParams p, const LcsLaneArgs o) {
    extern __shared__ __align__(128) unsigned char lane_smem[];
    __shared__ __align__(8) uint64_t full[2];
    float4* tiles = (float4*)lane_smem; // [BULK ? 2 : 1][LCS_LANE_TILE]
    float* hd = (float*)(tiles + (BULK ? 2 : 1) * LCS_LANE_TILE); // [K][128]
    int* hi = (int*)(hd + (size_t)o.K * 128);
    const int s = blockIdx.x, tid = threadIdx.x, K = o.K;
    const int n = o.n_pts[s], na = p.n_active[s];
    const float4* P = o.pts + (size_t)s * o.N;
    const int ntiles = (n + LCS_LANE_TILE - 1) / LCS_LANE_TILE;
    if (BULK) {
        if (tid == 0) {
            lcs_mbar_init(&full[0], 1);
            lcs_mbar_init(&full[1], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
    }
    unsigned used[2] = {0, 0}; // how often each buffer has been filled (parity of its barrier)
    for (int a = tid; a < p.Ap; a += 128)
        if (a < p.A && p.status[(size_t)s * p.Ap + a] != LCS_ACTIVE) lcs_lane_store_inactive(p, o, s, a);
    for (int base = 0; base < na; base += 128) {
        const int slot = base + tid;
        const int a = slot < na ? p.active[(size_t)s * p.Ap + slot] : -1;
        float qx = 0, qy = 0, qh = 0;
        if (a >= 0) {
            const double* ps = p.pose64 + 4 * ((size_t)s * p.Ap + a);
            qx = (float)ps[0]; qy = (float)ps[1]; qh = (float)ps[2];
        }
        int bi = -1, hn = 0;
        float bd2 = lcs_inf_f(), on_d2 = lcs_inf_f(), hmax = lcs_inf_f();
        if (BULK && tid == 0)
            for (int it = 0; it < 2 && it < ntiles; it++)
                lcs_bulk_load(tiles + it * LCS_LANE_TILE, P + it * LCS_LANE_TILE,
                              16u * (unsigned)min(LCS_LANE_TILE, n - it * LCS_LANE_TILE), &full[it]);
        for (int it = 0; it < ntiles; it++) {
            const int t0 = it * LCS_LANE_TILE, buf = BULK ? (it & 1) : 0;
            const float4* tile = tiles + buf * LCS_LANE_TILE;
            if (BULK) {
                lcs_mbar_wait(&full[buf], used[buf] & 1);
                used[buf]++;
            } else {
                __syncthreads();
                for (int k = tid; k < LCS_LANE_TILE && t0 + k < n; k += 128) tiles[k] = P[t0 + k];
                __syncthreads();
            }
            if (a >= 0) {
                const int m = min(LCS_LANE_TILE, n - t0);
                for (int k = 0; k < m; k++) {
                    const float4 c = tile[k];
                    const float d2 = lcs_lane_d2(qx, qy, c.x, c.y);
                    if (d2 < bd2) { bd2 = d2; bi = t0 + k; }
                    if (d2 < on_d2 && d2 < 25.001f) // cheap pre-test, then the reference's exact gates
                        if (sqrtf(d2) < 5.0f && fabsf(lcs_wrap_f32(qh - c.z)) < 0.2f) on_d2 = d2;
                    if (K > 0) {
                        // the K smallest (distance, index) pairs as a binary MAX-heap in the thread's shared-memory column:
                        // a point enters only when it beats the root, and then costs one sift-down (log2 K levels) — points
                        // along a lane that approaches the query beat the current maximum one after the other
                        if (hn < K) {
                            hd[hn * 128 + tid] = d2; hi[hn * 128 + tid] = t0 + k; hn++;
                            if (hn == K) { // first full column: heapify
                                for (int r0 = K / 2 - 1; r0 >= 0; r0--) lcs_heap_sift(hd, hi, tid, K, r0);
                                hmax = hd[tid];
                            }
                        } else if (d2 < hmax) {
                            hd[tid] = d2; hi[tid] = t0 + k;
                            lcs_heap_sift(hd, hi, tid, K, 0);
                            hmax = hd[tid];
                        }
                    }
                }
            }
            if (BULK) {
                __syncthreads(); // everybody is done with this buffer: the copy engine may refill it
                if (tid == 0 && it + 2 < ntiles)
                    lcs_bulk_load(tiles + buf * LCS_LANE_TILE, P + (it + 2) * LCS_LANE_TILE,
                                  16u * (unsigned)min(LCS_LANE_TILE, n - (it + 2) * LCS_LANE_TILE), &full[buf]);
            }
        }
        if (a >= 0) {
            float hdiff = -1.0f;
            if (K > 0 && hn == K) {
                hdiff = lcs_inf_f();
                for (int q = 0; q < K; q++) hdiff = fminf(hdiff, fabsf(lcs_wrap_f32(qh - P[hi[q * 128 + tid]].z)));
            }
            lcs_lane_store(p, o, s, a, bi, bd2, on_d2, hdiff);
        }
    }
}
#endif

extern "C" int lcsim_b200_upload_centrelines(lcsim_b200_engine* e, const int32_t* n_pts, const float* xy_theta,
                                             const int32_t* ids, int N) {
    if (!e || !n_pts || !xy_theta || !ids || N < 1) return LCSIM_B200_ERR_ARG;
    LCS_GUARD(e);
    const int S = e->p.S;
    std::vector<float> f((size_t)S * N * 4, 0.f);
    for (int s = 0; s < S; s++) {
        if (n_pts[s] < 0 || n_pts[s] > N) return lcs_fail(e, LCSIM_B200_ERR_ARG, "scenario %d: n_pts %d out of [0,%d]", s, n_pts[s], N);
        for (int k = 0; k < n_pts[s]; k++)
            for (int c = 0; c < 3; c++) f[((size_t)s * N + k) * 4 + c] = xy_theta[((size_t)s * N + k) * 3 + c];
    }
    lcs_release(e, e->lane_pts); // a second upload replaces the first
    lcs_release(e, e->lane_ids);
    lcs_release(e, e->lane_n);
    LCS_ALLOC(e, e->lane_pts, (size_t)S * N * 4);
    LCS_ALLOC(e, e->lane_ids, (size_t)S * N);
    LCS_ALLOC(e, e->lane_n, S);
    LCS_CUDA(e, cudaMemcpy(e->lane_pts, f.data(), f.size() * sizeof(float), cudaMemcpyHostToDevice));
    LCS_CUDA(e, cudaMemcpy(e->lane_ids, ids, sizeof(int32_t) * (size_t)S * N, cudaMemcpyHostToDevice));
    LCS_CUDA(e, cudaMemcpy(e->lane_n, n_pts, sizeof(int32_t) * S, cudaMemcpyHostToDevice));
    e->lane_N = N;
    return 0;
}

extern "C" int lcsim_b200_project_lanes(lcsim_b200_engine* e, const lcsim_b200_lane_desc* d, void* stream) {
    if (!e || !d) return LCSIM_B200_ERR_ARG;
    if (!e->uploaded) return lcs_fail(e, LCSIM_B200_ERR_STATE, "project_lanes before upload_static");
    LCS_GUARD(e);
    if (!e->lane_pts) return lcs_fail(e, LCSIM_B200_ERR_STATE, "project_lanes before upload_centrelines");
    if (d->top_k < 0 || d->top_k > LCS_LANE_KMAX) return lcs_fail(e, LCSIM_B200_ERR_ARG, "top_k must be in [0,%d]", LCS_LANE_KMAX);
    LcsLaneArgs o;
    o.d = *d;
    o.pts = (const float4*)e->lane_pts;
    o.ids = e->lane_ids;
    o.n_pts = e->lane_n;
    o.N = e->lane_N;
    o.K = d->heading_diff ? d->top_k : 0;
#ifndef LCSIM_HOSTEMU
    static const bool bulk = []() { // LCSIM_B200_LANE_BULK=0: tiles through the CTA's own loads / stores
        const char* env = getenv("LCSIM_B200_LANE_BULK");
        return !(env && env[0] == '0');
    }();
    const size_t smem = sizeof(float4) * LCS_LANE_TILE * (bulk ? 2 : 1) + (size_t)o.K * 128 * 8;
    if (bulk) {
        if (smem > 48 * 1024) LCS_CUDA(e, cudaFuncSetAttribute(lcs_lane_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        lcs_lane_kernel<true><<<e->p.S, 128, smem, (cudaStream_t)stream>>>(e->p, o);
    } else {
        if (smem > 48 * 1024) LCS_CUDA(e, cudaFuncSetAttribute(lcs_lane_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        lcs_lane_kernel<false><<<e->p.S, 128, smem, (cudaStream_t)stream>>>(e->p, o);
    }
    LCS_CUDA(e, cudaGetLastError());
#else
    const LcsParams& p = e->p;
    std::vector<float> hd(LCS_LANE_KMAX);
    std::vector<int> hi(LCS_LANE_KMAX);
    for (int s = 0; s < p.S; s++)
        for (int a = 0; a < p.A; a++) {
            const size_t ia = (size_t)s * p.Ap + a;
            if (p.status[ia] != LCS_ACTIVE) {
                lcs_lane_store_inactive(p, o, s, a);
                continue;
            }
            int bi;
            float bd2, on_d2, hdiff;
            lcs_lane_query_serial(o, s, (float)p.pose64[4 * ia], (float)p.pose64[4 * ia + 1], (float)p.pose64[4 * ia + 2], &bi, &bd2,
                                  &on_d2, &hdiff, hd.data(), hi.data());
            lcs_lane_store(p, o, s, a, bi, bd2, on_d2, hdiff);
        }
#endif
    return 0;
}

// ---- scenario loader: lcsim_loader.hpp -------------------------------------------------------------
extern "C" int lcsim_b200_load_dirs(const char* const* dirs, int n_dirs, double start, double interval, int32_t total_steps,
                                    int n_threads, lcsim_b200_scenarios** out) {
    if (!dirs || n_dirs < 1 || !out) return LCSIM_B200_ERR_ARG;
    lcsim_b200_scenarios* sc = new lcsim_b200_scenarios();
    *out = sc; // returned on failure too: the caller reads the message and frees it
    return lcs_load_dirs(dirs, n_dirs, start, interval, total_steps, n_threads, sc);
}
extern "C" int lcsim_b200_scenarios_dims(const lcsim_b200_scenarios* sc, int32_t* dims) {
    if (!sc || !dims) return LCSIM_B200_ERR_ARG;
    dims[0] = sc->S; dims[1] = sc->A; dims[2] = sc->T; dims[3] = sc->E; dims[4] = sc->P;
    return 0;
}
extern "C" int lcsim_b200_scenarios_static(const lcsim_b200_scenarios* sc, lcsim_b200_static_host* o) {
    if (!sc || !o) return LCSIM_B200_ERR_ARG;
    o->n_agents = sc->n_agents.data(); o->ads_index = sc->ads_index.data(); o->agent_type = sc->agent_type.data();
    o->length = sc->length.data(); o->width = sc->width.data(); o->dynamic = sc->dynamic.data(); o->traj_len = sc->traj_len.data();
    o->depart_tick = sc->depart_tick.data(); o->wait_order = sc->wait_order.data(); o->home_xy = sc->home_xy.data();
    o->traj_xy = sc->traj_xy.data(); o->traj_vel = sc->traj_vel.data(); o->traj_h = sc->traj_h.data(); o->n_edge = sc->n_edge.data();
    o->edge_xy = sc->edge_xy.data(); o->edge_dir = sc->edge_dir.data(); o->origin = sc->origin.data();
    o->edge_poly_id = sc->edge_poly_id.data();
    return 0;
}
extern "C" const int32_t* lcsim_b200_scenarios_agent_id(const lcsim_b200_scenarios* sc) { return sc ? sc->agent_id.data() : nullptr; }
extern "C" const int32_t* lcsim_b200_scenarios_junction_id(const lcsim_b200_scenarios* sc) { return sc ? sc->junction_id.data() : nullptr; }
extern "C" const double* lcsim_b200_scenarios_polylines(const lcsim_b200_scenarios* sc, const int32_t** n_poly) {
    if (!sc) return nullptr;
    if (n_poly) *n_poly = sc->n_poly.data();
    return sc->poly.data();
}
extern "C" const char* lcsim_b200_scenarios_error(const lcsim_b200_scenarios* sc) { return sc ? sc->err.c_str() : "null"; }
extern "C" void lcsim_b200_scenarios_free(lcsim_b200_scenarios* sc) { delete sc; }

// ---- guidance costs / plan metrics: lcsim_plan.cuh ---------------------------------------------
static int lcs_plan_parts(lcsim_b200_engine* e, int parts) {
    if (parts <= e->plan_parts) return 0;
    lcs_release(e, e->plan_sum);
    lcs_release(e, e->plan_cnt);
    LCS_ALLOC(e, e->plan_sum, (size_t)e->p.S * parts);
    LCS_ALLOC(e, e->plan_cnt, (size_t)e->p.S * parts);
    e->plan_parts = parts;
    return 0;
}
static void lcs_plan_args(lcsim_b200_engine* e, LcpArgs& o, int T) {
    memset(&o, 0, sizeof o);
    o.S = e->p.S; o.A = e->p.A; o.T = T;
    o.n_agents = e->p.n_agents;
    o.edge = e->plan_edge; o.edge_id = e->plan_edge_id; o.n_edge = e->p.n_edge; o.E = e->p.E;
    o.part_sum = e->plan_sum; o.part_cnt = e->plan_cnt; o.P = e->plan_parts;
}

extern "C" int lcsim_b200_guidance_cost(lcsim_b200_engine* e, int which, const float* xy, int T, float radius, float eps,
                                        float* loss, float* grad, float* sd, void* stream) {
    if (!e || !xy || T < 1 || !(radius > 0) || (which != 0 && which != 1)) return LCSIM_B200_ERR_ARG;
    if (!e->uploaded) return lcs_fail(e, LCSIM_B200_ERR_STATE, "guidance_cost before upload_static");
    if (which == 1 && !e->plan_edge_id) return lcs_fail(e, LCSIM_B200_ERR_STATE, "guidance_cost: NoOffRoad needs static.edge_poly_id");
    LCS_GUARD(e);
    const int A = e->p.A, qblocks = (A * T + 255) / 256;
    int rc = lcs_plan_parts(e, std::max(T, qblocks));
    if (rc) return rc;
    LcpArgs o;
    lcs_plan_args(e, o, T);
    o.xy = xy; o.radius = radius; o.eps = eps; o.loss = loss; o.grad = grad; o.sd = sd;
#ifndef LCSIM_HOSTEMU
    cudaStream_t st = (cudaStream_t)stream;
    if (which == 0) {
        const int nthr = (A + 31) / 32 * 32;
        lcp_repeller_kernel<<<dim3(T, o.S), nthr, sizeof(float) * 2 * nthr, st>>>(o);
        lcp_finalize_kernel<<<o.S, 256, 0, st>>>(o, 0, T);
    } else {
        lcp_offroad_kernel<<<dim3(qblocks, o.S), 256, 0, st>>>(o, 0);
        lcp_finalize_kernel<<<o.S, 256, 0, st>>>(o, 1, qblocks);
    }
    LCS_CUDA(e, cudaGetLastError());
#else
    if (which == 0) lcp_repeller_serial(o);
    else lcp_offroad_serial(o, 0);
#endif
    return 0;
}

extern "C" int lcsim_b200_plan_metrics(lcsim_b200_engine* e, const float* traj5, int T, const uint8_t* overlap_mask,
                                       const uint8_t* offroad_mask, float* overlap_rate, int32_t* overlap_count,
                                       float* offroad_rate, int32_t* offroad_count, void* stream) {
    if (!e || !traj5 || T < 1) return LCSIM_B200_ERR_ARG;
    if (!e->uploaded) return lcs_fail(e, LCSIM_B200_ERR_STATE, "plan_metrics before upload_static");
    const bool want_off = offroad_rate || offroad_count, want_ovl = overlap_rate || overlap_count;
    if (want_off && !e->plan_edge_id) return lcs_fail(e, LCSIM_B200_ERR_STATE, "plan_metrics: the off-road rate needs static.edge_poly_id");
    LCS_GUARD(e);
    const int A = e->p.A, qblocks = (A * T + 255) / 256;
    int rc = lcs_plan_parts(e, std::max(T, qblocks));
    if (rc) return rc;
    LcpArgs o;
    lcs_plan_args(e, o, T);
    o.traj5 = traj5;
#ifndef LCSIM_HOSTEMU
    cudaStream_t st = (cudaStream_t)stream;
    if (want_ovl) {
        const int nthr = (A + 31) / 32 * 32;
        const size_t smem = (sizeof(LcpBox) + 3 * sizeof(float)) * nthr;
        if (smem > 48 * 1024) LCS_CUDA(e, cudaFuncSetAttribute(lcp_overlap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        o.mask = overlap_mask; o.rate = overlap_rate; o.count = overlap_count;
        lcp_overlap_kernel<<<dim3(T, o.S), nthr, smem, st>>>(o);
        lcp_finalize_kernel<<<o.S, 256, 0, st>>>(o, 2, T);
    }
    if (want_off) {
        o.mask = offroad_mask; o.rate = offroad_rate; o.count = offroad_count;
        lcp_offroad_kernel<<<dim3(qblocks, o.S), 256, 0, st>>>(o, 1);
        lcp_finalize_kernel<<<o.S, 256, 0, st>>>(o, 2, qblocks);
    }
    LCS_CUDA(e, cudaGetLastError());
#else
    if (want_ovl) {
        o.mask = overlap_mask; o.rate = overlap_rate; o.count = overlap_count;
        lcp_overlap_serial(o);
    }
    if (want_off) {
        o.mask = offroad_mask; o.rate = offroad_rate; o.count = offroad_count;
        lcp_offroad_serial(o, 1);
    }
#endif
    return 0;
}

extern "C" int lcsim_b200_views(lcsim_b200_engine* e, lcsim_b200_state_views* v) {
    if (!e || !v) return LCSIM_B200_ERR_ARG;
    const LcsParams& p = e->p;
    v->S = p.S; v->A = p.A; v->A_stride = p.Ap; v->T = p.T; v->E = p.E; v->H = LCS_H;
    v->tick = p.tick; v->n_active = p.n_active; v->active = p.active; v->status = p.status; v->cursor = p.cursor;
    v->ring_count = p.ring_count; v->ring_head = p.ring_head; v->pose64 = p.pose64; v->pose32 = (float*)p.pose32;
    v->ring = p.ring; v->lead_valid = p.lead_valid; v->lead_dis = p.lead_dis; v->lead_v = p.lead_v;
    v->coll_count = p.coll_count; v->nearest_edge = p.nearest_edge; v->offroad = p.offroad;
    v->traj_len = p.traj_len; v->traj_xy = p.traj_xy; v->traj_vel = p.traj_vel; v->traj_h = p.traj_h;
    v->stats = p.stats; v->origin = (double*)p.origin;
    return 0;
}

extern "C" int lcsim_b200_episode_stats(lcsim_b200_engine* e, int64_t* stats_device, void* stream) {
    if (!e || !stats_device) return LCSIM_B200_ERR_ARG;
    LCS_GUARD(e);
#ifndef LCSIM_HOSTEMU
    lcs_stats_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(e->p.stats, e->p.S, stats_device);
    LCS_CUDA(e, cudaGetLastError());
#else
    for (int k = 0; k < LCS_NSTAT; k++) {
        int64_t acc = 0;
        for (int s = 0; s < e->p.S; s++) acc += e->p.stats[(size_t)s * LCS_NSTAT + k];
        stats_device[k] = acc;
    }
#endif
    return 0;
}

#ifdef LCSIM_HOSTEMU
extern "C" long long* lcsim_hostemu_counters(void) { return lcs_dbg_count; }
extern "C" double lcsim_hostemu_pow4(double x) { return lcs_pow4(x); } // test hook (tests/test_numerics.py)
extern "C" double lcsim_hostemu_wrap_angle(double a) { return lcs_wrap_angle(a); }
#endif
// not part of the public ABI: device pointer of the phase-clock buffer ([S][16] int64) or NULL
extern "C" void* lcsim_b200_debug_clocks(lcsim_b200_engine* e) { return e ? (void*)e->p.dbg : nullptr; }

// ====================================================================================================
// City engine: LaneIDM on a lane graph.  Kernels: lcsim_city.cuh.
#include "lcsim_city.cuh"

// ---- exclusive prefix sum of int32 (in place), total written to *total (device).  Device: block-local scan of
// 2048-element tiles, scan of the tile sums by one block, offset add.  n <= 2048 * 2048.
#ifndef LCSIM_HOSTEMU
#define LCC_SCAN_TILE 2048
__global__ void __launch_bounds__(1024) lcc_scan_tiles(const int32_t* in, int32_t* data, int n, int32_t* tile_sum) {
    __shared__ int32_t warp_sum[32];
    const int tid = threadIdx.x, base = blockIdx.x * LCC_SCAN_TILE + 2 * tid;
    const int a = base < n ? in[base] : 0, b = base + 1 < n ? in[base + 1] : 0;
    int v = a + b;
    const int lane = tid & 31, w = tid >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += o;
    }
    if (lane == 31) warp_sum[w] = v;
    __syncthreads();
    if (w == 0) {
        int s = warp_sum[lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_up_sync(0xffffffffu, s, d);
            if (lane >= d) s += o;
        }
        warp_sum[lane] = s;
    }
    __syncthreads();
    const int excl = v - (a + b) + (w > 0 ? warp_sum[w - 1] : 0);
    if (base < n) data[base] = excl;
    if (base + 1 < n) data[base + 1] = excl + a;
    if (tid == 1023) tile_sum[blockIdx.x] = excl + a + b;
}
// second pass: element i gets the sum of the tiles in front of its own; every block adds those tile sums up itself
// (a city has ~100 tiles, a block of 256 elements lies inside one tile), so there is no kernel for the scan of the
// tile sums between the two passes.  Block 0 also writes the grand total.
__global__ void __launch_bounds__(256) lcc_scan_add(int32_t* data, int n, const int32_t* tile_sum, int n_tiles, int32_t* copy, int32_t* total) {
    __shared__ int32_t part[8];
    __shared__ int32_t s_off;
    const int tile = blockIdx.x * 256 / LCC_SCAN_TILE, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int upto = (blockIdx.x == 0 && total) ? n_tiles : tile; // block 0 sums everything (tile == 0 contributes nothing to it)
    int32_t acc = 0, acc_all = 0;
    for (int j = threadIdx.x; j < upto; j += 256) {
        const int32_t v = tile_sum[j];
        acc_all += v;
        if (j < tile) acc += v;
    }
    const bool want_total = blockIdx.x == 0 && total;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        acc += __shfl_xor_sync(0xffffffffu, acc, d);
        acc_all += __shfl_xor_sync(0xffffffffu, acc_all, d);
    }
    if (lane == 0) part[w] = want_total ? acc_all : acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        int32_t t = 0;
        for (int k = 0; k < 8; k++) t += part[k];
        if (want_total) *total = t;
        s_off = want_total ? 0 : t; // block 0 lies in tile 0: nothing in front of it
    }
    __syncthreads();
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i < n) {
        const int32_t v = data[i] + s_off;
        data[i] = v;
        if (copy) copy[i] = v; // the scatter's cursor array, written in the same pass
    }
}
#endif

#define LCC_READER_ROUNDS 2
struct lcsim_b200_city {
    LccP c;
    std::vector<void*> allocs;
    int32_t* tile_sum = nullptr; // [3][tile_cap]
    size_t tile_cap = 0;
    int with_metrics = 1;
    int device = 0;
    bool fresh_reset = false;
    int64_t launches = 0;
    int launches_per_tick = 0;
#ifndef LCSIM_HOSTEMU
    cudaGraphExec_t tick_graph = nullptr; // one tick's launch sequence, captured once (LCSIM_B200_CITY_GRAPH=0: off)
    int use_graph = 1;
    // the part of a tick behind the new active list forks into independent branches (lane buckets + observations |
    // collision cells + SAT | off-road): side streams and fork / join events (LCSIM_B200_CITY_FORK=0: one stream)
    cudaStream_t side[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev_upd = nullptr, ev_list = nullptr, ev_side[3] = {nullptr, nullptr, nullptr};
    int use_fork = 1;
#endif
    std::string err;
};

// which: the scan's own tile-sum scratch (0 = list ranks, 1 = lane buckets, 2 = collision cells; the last two run on
// concurrent branches of the tick)
// in: the values (may be `data` itself: in place)
static void lcc_scan(lcsim_b200_city* e, int which, const int32_t* in, int32_t* data, int n, int32_t* total, int32_t* copy, cudaStream_t st) {
#ifndef LCSIM_HOSTEMU
    const int tiles = (n + LCC_SCAN_TILE - 1) / LCC_SCAN_TILE;
    int32_t* ts = e->tile_sum + (size_t)which * e->tile_cap;
    lcc_scan_tiles<<<tiles, 1024, 0, st>>>(in, data, n, ts);
    lcc_scan_add<<<(n + 255) / 256, 256, 0, st>>>(data, n, ts, tiles, copy, total);
    e->launches += 2;
#else
    (void)which;
    int32_t acc = 0;
    for (int i = 0; i < n; i++) {
        const int32_t v = in[i];
        data[i] = acc;
        if (copy) copy[i] = acc;
        acc += v;
    }
    if (total) *total = acc;
#endif
}

static int lcc_fail(lcsim_b200_city* e, int code, const char* msg) {
    if (e) e->err = msg;
    return code;
}
template <typename Tp>
static int lcc_upload(lcsim_b200_city* e, const Tp** dst, const Tp* src, size_t count) {
    void* ptr = nullptr;
    if (cudaMalloc(&ptr, std::max<size_t>(count, 1) * sizeof(Tp)) != cudaSuccess) return lcc_fail(e, LCSIM_B200_ERR_NOMEM, "cudaMalloc failed");
    e->allocs.push_back(ptr);
    if (count && cudaMemcpy(ptr, src, count * sizeof(Tp), cudaMemcpyHostToDevice) != cudaSuccess) return lcc_fail(e, LCSIM_B200_ERR_CUDA, "cudaMemcpy failed");
    *dst = (const Tp*)ptr;
    return 0;
}
template <typename Tp>
static int lcc_alloc(lcsim_b200_city* e, Tp** dst, size_t count) {
    void* ptr = nullptr;
    if (cudaMalloc(&ptr, std::max<size_t>(count, 1) * sizeof(Tp)) != cudaSuccess) return lcc_fail(e, LCSIM_B200_ERR_NOMEM, "cudaMalloc failed");
    e->allocs.push_back(ptr);
    if (cudaMemset(ptr, 0, std::max<size_t>(count, 1) * sizeof(Tp)) != cudaSuccess) return lcc_fail(e, LCSIM_B200_ERR_CUDA, "cudaMemset failed");
    *dst = (Tp*)ptr;
    return 0;
}
#define LCC_TRY(x)          \
    do {                    \
        int _rc = (x);      \
        if (_rc) return _rc; \
    } while (0)

extern "C" const char* lcsim_b200_city_last_error(const lcsim_b200_city* e) { return e ? e->err.c_str() : "null city"; }
extern "C" int64_t lcsim_b200_city_launch_count(const lcsim_b200_city* e) { return e ? e->launches : 0; }
extern "C" void lcsim_b200_city_destroy(lcsim_b200_city* e) {
    if (!e) return;
#ifndef LCSIM_HOSTEMU
    if (e->tick_graph) cudaGraphExecDestroy(e->tick_graph);
#endif
    for (void* q : e->allocs) cudaFree(q);
#ifndef LCSIM_HOSTEMU
    for (int k = 0; k < 3; k++) {
        if (e->ev_side[k]) cudaEventDestroy(e->ev_side[k]);
        if (e->side[k]) cudaStreamDestroy(e->side[k]);
    }
    if (e->ev_upd) cudaEventDestroy(e->ev_upd);
    if (e->ev_list) cudaEventDestroy(e->ev_list);
#endif
    delete e;
}

extern "C" int lcsim_b200_city_create(const lcsim_b200_city_static* st, int device, int with_metrics, lcsim_b200_city** out) {
    if (!st || !out) return LCSIM_B200_ERR_ARG;
    *out = nullptr;
    if (st->n_agents < 1 || st->n_lanes < 1 || st->n_agents >= (1 << 24)) return LCSIM_B200_ERR_ARG;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1 || device < 0 || device >= ndev) return LCSIM_B200_ERR_CUDA;
    LcsDeviceGuard guard(device); // the caller's current device is restored on return
    if (!guard.ok) return LCSIM_B200_ERR_CUDA;
    lcsim_b200_city* e = new lcsim_b200_city();
    *out = e;
    e->device = device;
    LccP& c = e->c;
    memset(&c, 0, sizeof c);
    const int N = st->n_agents, NL = st->n_lanes;
    c.N = N;
    c.NL = NL;
    c.dt = st->interval;
    c.with_metrics = e->with_metrics = with_metrics;
#ifndef LCSIM_HOSTEMU
    {
        const char* env = getenv("LCSIM_B200_CITY_GRAPH");
        e->use_graph = !(env && env[0] == '0');
    }
#endif
    const int NN = st->lane_node_start[NL];
    // ---- validation of the graph indices the kernels follow
    for (int l = 0; l < NL; l++) {
        if (st->lane_node_start[l + 1] - st->lane_node_start[l] < 2) return lcc_fail(e, LCSIM_B200_ERR_ARG, "lane with fewer than 2 centre-line nodes");
        if ((st->lane_road[l] >= 0) == (st->lane_junction[l] >= 0)) return lcc_fail(e, LCSIM_B200_ERR_ARG, "lane must belong to exactly one road or junction");
        if (st->lane_junction[l] >= 0 && (st->lane_succ0[l] < 0 || st->lane_succ0[l] >= NL)) return lcc_fail(e, LCSIM_B200_ERR_ARG, "junction lane without successor");
        if (st->lane_eset[l] < 0 || st->lane_eset[l] >= st->n_esets) return lcc_fail(e, LCSIM_B200_ERR_ARG, "lane_eset out of range");
        if (st->lane_left[l] >= NL || st->lane_right[l] >= NL) return lcc_fail(e, LCSIM_B200_ERR_ARG, "lane neighbour out of range");
    }
    const int n_route = st->route_start[N];
    for (int a = 0; a < N; a++) {
        if (st->home_lane[a] < 0 || st->home_lane[a] >= NL) return lcc_fail(e, LCSIM_B200_ERR_ARG, "home_lane out of range");
        if (st->route_start[a + 1] - st->route_start[a] < 1) return lcc_fail(e, LCSIM_B200_ERR_ARG, "agent with an empty route (reference: Route.valid is False)");
        if (st->wait_order[a] < 0 || st->wait_order[a] >= N) return lcc_fail(e, LCSIM_B200_ERR_ARG, "wait_order out of range");
        if (a > 0 && st->depart_tick[st->wait_order[a]] < st->depart_tick[st->wait_order[a - 1]]) return lcc_fail(e, LCSIM_B200_ERR_ARG, "depart_tick not sorted along wait_order");
    }
    // ---- spatial hash geometry: bounding box of all lane nodes, cell >= largest bounding-circle diameter
    double x0 = 1e300, y0 = 1e300, x1 = -1e300, y1 = -1e300, diag = 0;
    for (int k = 0; k < NN; k++) {
        x0 = std::min(x0, st->node_xy[2 * k]); x1 = std::max(x1, st->node_xy[2 * k]);
        y0 = std::min(y0, st->node_xy[2 * k + 1]); y1 = std::max(y1, st->node_xy[2 * k + 1]);
    }
    for (int a = 0; a < N; a++) diag = std::max(diag, std::sqrt(st->length[a] * st->length[a] + st->width[a] * st->width[a]));
    double cell = std::max(8.0, diag * 1.01 + 1e-3);
    for (;;) { // bound the table
        c.cell_w = (int)std::floor((x1 - x0) / cell) + 1;
        c.cell_h = (int)std::floor((y1 - y0) / cell) + 1;
        if ((int64_t)c.cell_w * c.cell_h <= 4000000) break;
        cell *= 1.5;
    }
    c.n_cells = c.cell_w * c.cell_h;
    c.cell_x0 = x0; c.cell_y0 = y0; c.cell_inv = 1.0 / cell;
    // ---- static
    LCC_TRY(lcc_upload(e, &c.lane_length, st->lane_length, NL));
    LCC_TRY(lcc_upload(e, &c.lane_max_speed, st->lane_max_speed, NL));
    LCC_TRY(lcc_upload(e, &c.lane_width, st->lane_width, NL));
    LCC_TRY(lcc_upload(e, &c.lane_left, st->lane_left, NL));
    LCC_TRY(lcc_upload(e, &c.lane_right, st->lane_right, NL));
    LCC_TRY(lcc_upload(e, &c.lane_road, st->lane_road, NL));
    LCC_TRY(lcc_upload(e, &c.lane_junction, st->lane_junction, NL));
    LCC_TRY(lcc_upload(e, &c.lane_offset, st->lane_offset, NL));
    LCC_TRY(lcc_upload(e, &c.lane_succ0, st->lane_succ0, NL));
    LCC_TRY(lcc_upload(e, &c.lane_node_start, st->lane_node_start, NL + 1));
    LCC_TRY(lcc_upload(e, &c.lane_eset, st->lane_eset, NL));
    LCC_TRY(lcc_upload(e, &c.node_xy, st->node_xy, (size_t)NN * 2));
    LCC_TRY(lcc_upload(e, &c.node_cum, st->node_cum, NN));
    LCC_TRY(lcc_upload(e, &c.node_dir, st->node_dir, NN));
    LCC_TRY(lcc_upload(e, &c.eset_start, st->eset_start, st->n_esets + 1));
    LCC_TRY(lcc_upload(e, &c.edge_xy, st->edge_xy, (size_t)st->eset_start[st->n_esets] * 2));
    LCC_TRY(lcc_upload(e, &c.edge_dir, st->edge_dir, (size_t)st->eset_start[st->n_esets] * 2));
    LCC_TRY(lcc_upload(e, &c.grp_start, st->grp_start, st->n_groups + 1));
    LCC_TRY(lcc_upload(e, &c.grp_lane, st->grp_lane, st->grp_start[st->n_groups]));
    LCC_TRY(lcc_upload(e, &c.grp_pre_offset, st->grp_pre_offset, st->grp_start[st->n_groups]));
    LCC_TRY(lcc_upload(e, &c.length, st->length, N));
    LCC_TRY(lcc_upload(e, &c.width, st->width, N));
    LCC_TRY(lcc_upload(e, &c.idm, st->idm, (size_t)N * 5));
    LCC_TRY(lcc_upload(e, &c.home_s, st->home_s, N));
    LCC_TRY(lcc_upload(e, &c.end_s, st->end_s, N));
    LCC_TRY(lcc_upload(e, &c.home_lane, st->home_lane, N));
    LCC_TRY(lcc_upload(e, &c.end_road, st->end_road, N));
    LCC_TRY(lcc_upload(e, &c.end_offset, st->end_offset, N));
    LCC_TRY(lcc_upload(e, &c.depart_tick, st->depart_tick, N));
    LCC_TRY(lcc_upload(e, &c.wait_order, st->wait_order, N));
    LCC_TRY(lcc_upload(e, &c.route_start, st->route_start, N + 1));
    LCC_TRY(lcc_upload(e, &c.route_grp, st->route_grp, std::max(n_route - N, 0)));
    // ---- state
    LCC_TRY(lcc_alloc(e, &c.status, N)); LCC_TRY(lcc_alloc(e, &c.lane, N)); LCC_TRY(lcc_alloc(e, &c.r_pop, N));
    LCC_TRY(lcc_alloc(e, &c.g_pop, N)); LCC_TRY(lcc_alloc(e, &c.at_road, N)); LCC_TRY(lcc_alloc(e, &c.ring_head, N));
    LCC_TRY(lcc_alloc(e, &c.ring_count, N)); LCC_TRY(lcc_alloc(e, &c.coll_count, N));
    LCC_TRY(lcc_alloc(e, &c.s, N)); LCC_TRY(lcc_alloc(e, &c.v, N)); LCC_TRY(lcc_alloc(e, &c.xy, (size_t)N * 2));
    LCC_TRY(lcc_alloc(e, &c.h, N)); LCC_TRY(lcc_alloc(e, &c.lead_dis, N)); LCC_TRY(lcc_alloc(e, &c.lead_v, N));
    LCC_TRY(lcc_alloc(e, &c.ring, (size_t)N * LCS_H * 5));
    LCC_TRY(lcc_alloc(e, &c.lead_valid, N)); LCC_TRY(lcc_alloc(e, &c.offroad, N));
    LCC_TRY(lcc_alloc(e, &c.is_lc, N)); LCC_TRY(lcc_alloc(e, &c.lc_target, N)); LCC_TRY(lcc_alloc(e, &c.slot_of, N));
    LCC_TRY(lcc_alloc(e, &c.rd_ahead, N)); LCC_TRY(lcc_alloc(e, &c.rd_lane, N)); LCC_TRY(lcc_alloc(e, &c.done_round, N));
    LCC_TRY(lcc_alloc(e, &c.lc_total, N)); LCC_TRY(lcc_alloc(e, &c.lc_completed, N)); LCC_TRY(lcc_alloc(e, &c.lc_target_s, N));
    LCC_TRY(lcc_alloc(e, &c.v2, N)); LCC_TRY(lcc_alloc(e, &c.acc, N)); LCC_TRY(lcc_alloc(e, &c.rd_dist, N));
    LCC_TRY(lcc_alloc(e, &c.active, N)); LCC_TRY(lcc_alloc(e, &c.active2, N)); LCC_TRY(lcc_alloc(e, &c.fin, N));
    LCC_TRY(lcc_alloc(e, &c.rank, N)); LCC_TRY(lcc_alloc(e, &c.scal, 16)); LCC_TRY(lcc_alloc(e, &c.stats, LCS_NSTAT));
    LCC_TRY(lcc_alloc(e, &c.ent_lane, (size_t)2 * N)); LCC_TRY(lcc_alloc(e, &c.ent_agent, (size_t)2 * N));
    LCC_TRY(lcc_alloc(e, &c.ent_order, (size_t)2 * N)); LCC_TRY(lcc_alloc(e, &c.ent_s, (size_t)2 * N));
    LCC_TRY(lcc_alloc(e, &c.lane_start, NL + 1)); LCC_TRY(lcc_alloc(e, &c.lane_cur, NL + 1));
    LCC_TRY(lcc_alloc(e, &c.bkt_agent, (size_t)2 * N)); LCC_TRY(lcc_alloc(e, &c.bkt_order, (size_t)2 * N));
    LCC_TRY(lcc_alloc(e, &c.bkt_s, (size_t)2 * N));
    LCC_TRY(lcc_alloc(e, &c.cell_start, c.n_cells + 1)); LCC_TRY(lcc_alloc(e, &c.cell_cur, c.n_cells + 1));
    LCC_TRY(lcc_alloc(e, &c.cell_agent, N));
    LCC_TRY(lcc_alloc(e, &c.cell_rec, (size_t)N * 8));
    const int max_scan = std::max(std::max(N, NL + 1), c.n_cells + 1);
    e->tile_cap = (size_t)max_scan / 2048 + 2;
    LCC_TRY(lcc_alloc(e, &e->tile_sum, 3 * e->tile_cap));
    {
        const char* ord = getenv("LCSIM_B200_CITY_ORDER");
        c.by_agent = !(ord && ord[0] == 's');
    }
#ifndef LCSIM_HOSTEMU
    {
        const char* env = getenv("LCSIM_B200_CITY_FORK");
        e->use_fork = !(env && env[0] == '0');
        bool ok = true;
        if (e->use_fork) {
            for (int k = 0; k < 3 && ok; k++)
                ok = cudaStreamCreateWithFlags(&e->side[k], cudaStreamNonBlocking) == cudaSuccess &&
                     cudaEventCreateWithFlags(&e->ev_side[k], cudaEventDisableTiming) == cudaSuccess;
            ok = ok && cudaEventCreateWithFlags(&e->ev_upd, cudaEventDisableTiming) == cudaSuccess &&
                 cudaEventCreateWithFlags(&e->ev_list, cudaEventDisableTiming) == cudaSuccess;
            if (!ok) return lcc_fail(e, LCSIM_B200_ERR_CUDA, "city: side streams / events could not be created");
        }
    }
#endif
    int rc = lcsim_b200_city_reset(e, nullptr);
    if (rc) return rc;
    if (cudaDeviceSynchronize() != cudaSuccess) return lcc_fail(e, LCSIM_B200_ERR_CUDA, "reset failed");
    return 0;
}

// arrivals/departures -> (lane buckets) -> observations -> metrics: the part of a tick after the update, which is
// also everything Engine.reset does after re-initialising the agents
// `after_update`: the caller recorded e->ev_upd on st behind the last kernel that READS the lane buckets / cells of the
// previous tick, so the branches may clear their counters from there on (tick); otherwise they start behind the list.
static void lcc_prepare(lcsim_b200_city* e, LccP& c, bool swap_lanes, bool after_update, cudaStream_t st) {
    const int N = c.N;
#ifndef LCSIM_HOSTEMU
    const bool fork = e->use_fork;
    cudaStream_t sa = fork ? e->side[0] : st, sb = fork ? e->side[1] : st, sc = fork ? e->side[2] : st;
    if (fork) {
        if (!after_update) cudaEventRecord(e->ev_upd, st);
        if (swap_lanes) cudaStreamWaitEvent(sa, e->ev_upd, 0);
        if (c.with_metrics) cudaStreamWaitEvent(sb, e->ev_upd, 0);
    }
#else
    cudaStream_t sa = st, sb = st, sc = st;
    (void)after_update;
#endif
    if (swap_lanes) cudaMemsetAsync(c.lane_start, 0, sizeof(int32_t) * (c.NL + 1), sa);
    if (c.with_metrics) cudaMemsetAsync(c.cell_start, 0, sizeof(int32_t) * (c.n_cells + 1), sb);
    // ---- the new active list (main stream); in a tick the pop test ran with the speed commit (lcc_commit_pop)
    if (!after_update) {
        LCC_RUN(lcc_pop_count, c, N, st);
        e->launches += 1;
    }
    lcc_scan(e, 0, c.fin, c.rank, N, c.scal + LCC_N_ARR, nullptr, st);
#ifndef LCSIM_HOSTEMU
    lcc_place_all_k<<<(N + 255) / 256, 256, 0, st>>>(c, N); // + lcc_place_end by its last block
#else
    LCC_RUN(lcc_place, c, N, st);
    LCC_RUN(lcc_place_end, c, 1, st);
#endif
    e->launches += 1;
#ifndef LCSIM_HOSTEMU
    if (fork) {
        cudaEventRecord(e->ev_list, st);
        cudaStreamWaitEvent(sa, e->ev_list, 0);
        if (c.with_metrics) {
            cudaStreamWaitEvent(sb, e->ev_list, 0);
            cudaStreamWaitEvent(sc, e->ev_list, 0);
        }
    }
#endif
    // ---- branch A: Lane.prepare (agents <- new_agents, lane.py:171-175) and the observations
    if (swap_lanes) {
        LCC_RUN(lcc_bucket_count, c, 2 * N, sa);
        lcc_scan(e, 1, c.lane_start, c.lane_start, c.NL + 1, nullptr, c.lane_cur, sa);
        LCC_RUN(lcc_bucket_scatter, c, 2 * N, sa);
        e->launches += 2; // (lcc_bucket_end is the first line of lcc_observe)
    }
    LCC_RUN(lcc_observe, c, N, sa);
    e->launches += 1;
    if (c.with_metrics) {
        // ---- branch B: collision cells + SAT; branch C: off-road
        LCC_RUN(lcc_cell_count, c, N, sb);
        lcc_scan(e, 2, c.cell_start, c.cell_start, c.n_cells + 1, nullptr, c.cell_cur, sb);
        LCC_RUN(lcc_cell_scatter, c, N, sb);
        LCC_RUN(lcc_collide, c, N, sb);
        LCC_RUN(lcc_offroad, c, N, sc);
        e->launches += 4;
    }
    LCC_RUN(lcc_finish, c, N, st); // list <- new list; a tick also advances the clock here
    e->launches += 1;
#ifndef LCSIM_HOSTEMU
    if (fork) { // join
        cudaEventRecord(e->ev_side[0], sa);
        cudaStreamWaitEvent(st, e->ev_side[0], 0);
        if (c.with_metrics) {
            cudaEventRecord(e->ev_side[1], sb);
            cudaStreamWaitEvent(st, e->ev_side[1], 0);
            cudaEventRecord(e->ev_side[2], sc);
            cudaStreamWaitEvent(st, e->ev_side[2], 0);
        }
    }
#endif
}

extern "C" int lcsim_b200_city_reset(lcsim_b200_city* e, void* stream) {
    if (!e) return LCSIM_B200_ERR_ARG;
    LCS_GUARD(e);
    cudaStream_t st = (cudaStream_t)stream;
    LccP c = e->c;
    c.mode = 1;
    cudaMemsetAsync(c.scal, 0, sizeof(int32_t) * 16, st);
    cudaMemsetAsync(c.stats, 0, sizeof(int64_t) * LCS_NSTAT, st);
    cudaMemsetAsync(c.lane_start, 0, sizeof(int32_t) * (c.NL + 1), st); // lane_manager.reset(): every lane list empty
    LCC_RUN(lcc_reset_agent, c, c.N, st);
    e->launches += 1;
    // Engine.reset (engine.py:139-143) never calls lane_manager.prepare(): the departed agents stay pending
    lcc_prepare(e, c, false, false, st);
    if (cudaGetLastError() != cudaSuccess) return lcc_fail(e, LCSIM_B200_ERR_CUDA, "city reset launch failed");
    return 0;
}

// one Engine.step(): update (+ the readers' rounds) -> prepare -> clock
static void lcc_tick(lcsim_b200_city* e, LccP& c, cudaStream_t st) {
    LCC_RUN(lcc_update, c, c.N, st);
    for (int r = 1; r <= LCC_READER_ROUNDS; r++) { // readers of live speeds (lane changes), by dependency depth
        LccP cr = c;
        cr.mode = r;
        LCC_RUN(lcc_update_round, cr, c.N, st);
    }
    LCC_RUN(lcc_update_serial, c, 1, st);
    LCC_RUN(lcc_commit_pop, c, c.N, st);
#ifndef LCSIM_HOSTEMU
    if (e->use_fork) cudaEventRecord(e->ev_upd, st);
#endif
    lcc_prepare(e, c, true, true, st);
    e->launches += 3 + LCC_READER_ROUNDS;
}

extern "C" int lcsim_b200_city_step(lcsim_b200_city* e, int n_ticks, void* stream) {
    if (!e || n_ticks < 0) return LCSIM_B200_ERR_ARG;
    LCS_GUARD(e);
    cudaStream_t st = (cudaStream_t)stream;
    LccP c = e->c;
    c.mode = 0;
#ifndef LCSIM_HOSTEMU
    // The launch sequence of a tick is static (all state lives in device memory), and at ~25 short launches per
    // tick the host's launch rate is the limit: capture it once into a CUDA graph and replay that.
    if (e->use_graph && !e->tick_graph && n_ticks > 0) {
        cudaStream_t cs = nullptr;
        cudaGraph_t g = nullptr;
        const int64_t before = e->launches;
        bool ok = cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking) == cudaSuccess &&
                  cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
        if (ok) {
            lcc_tick(e, c, cs);
            ok = cudaStreamEndCapture(cs, &g) == cudaSuccess && g && cudaGraphInstantiate(&e->tick_graph, g, 0) == cudaSuccess;
        }
        e->launches_per_tick = (int)(e->launches - before);
        e->launches = before;
        if (g) cudaGraphDestroy(g);
        if (cs) cudaStreamDestroy(cs);
        if (!ok) {
            cudaGetLastError();
            e->tick_graph = nullptr;
            e->use_graph = 0;
        }
    }
    if (e->tick_graph) {
        for (int k = 0; k < n_ticks; k++)
            if (cudaGraphLaunch(e->tick_graph, st) != cudaSuccess) return lcc_fail(e, LCSIM_B200_ERR_CUDA, "city tick graph launch failed");
        e->launches += (int64_t)n_ticks * e->launches_per_tick;
        return 0;
    }
#endif
    for (int k = 0; k < n_ticks; k++) lcc_tick(e, c, st);
    if (cudaGetLastError() != cudaSuccess) return lcc_fail(e, LCSIM_B200_ERR_CUDA, "city step launch failed");
    return 0;
}

extern "C" int lcsim_b200_city_views(lcsim_b200_city* e, lcsim_b200_city_state_views* v) {
    if (!e || !v) return LCSIM_B200_ERR_ARG;
    const LccP& c = e->c;
    v->N = c.N; v->n_lanes = c.NL; v->H = LCS_H;
    v->scalars = c.scal; v->is_lc = c.is_lc; v->active = c.active; v->status = c.status; v->lane = c.lane; v->s = c.s; v->v = c.v; v->xy = c.xy;
    v->heading = c.h; v->lead_valid = c.lead_valid; v->lead_dis = c.lead_dis; v->lead_v = c.lead_v;
    v->coll_count = c.coll_count; v->offroad = c.offroad; v->ring = c.ring; v->ring_head = c.ring_head;
    v->ring_count = c.ring_count; v->stats = c.stats;
    return 0;
}
