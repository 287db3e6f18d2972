/*
 * lcsim_b200.h — C ABI of the B200-native per-tick agent-stepping path of LCSim.
 *
 * The reference (tsinghua-fib-lab/LCSim) is pure Python and has no FFI of its own; its callers
 * reach the hot path through Python methods.  Every entry point below names the reference
 * method(s) it replaces (paths relative to the reference root).  INTEGRATION.md shows the
 * ctypes stub a maintainer of the reference would add to route those methods here.
 *
 * Conventions
 *  - every function returns 0 on success or a negative lcsim_b200_status; the message is read
 *    with lcsim_b200_last_error().  No exception crosses this boundary.
 *  - pointers documented "device" are CUDA device pointers of the engine's device; pointers
 *    documented "host" are ordinary host pointers.  `stream` is a cudaStream_t passed as void*.
 *  - the engine owns every persistent device buffer; lcsim_b200_views() lends them out.  Outputs
 *    of build_obs / rewards are caller-owned buffers borrowed for the call.
 *  - no host synchronisation and no allocation happen inside step / rollout / rewards / build_obs /
 *    set_ref_trajectory / episode_stats: the work is only enqueued on `stream` (graph-capturable).
 *  - a handle is not thread-safe: one host thread drives one handle.
 *  - there is no CPU fallback: on a machine without a usable CUDA device create() fails.
 */
#ifndef LCSIM_B200_H
#define LCSIM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LCSIM_B200_ABI_VERSION 2

/* reference: lcsim/entity/agent/agent.py:89-90 */
#define LCSIM_B200_H_STEPS 11 /* NUM_HISTORICAL_STEPS */
#define LCSIM_B200_ROUTE_LEN 10 /* ROUTE_LENGTH */
#define LCSIM_B200_N_REWARD_TERMS 6 /* forward, route_progress, smooth, collision, offroad, destination */
#define LCSIM_B200_N_STATS 8

typedef enum {
    LCSIM_B200_OK = 0,
    LCSIM_B200_ERR_ARG = -1, /* bad argument / shape */
    LCSIM_B200_ERR_CUDA = -2, /* a CUDA runtime call failed (message holds cudaGetErrorString) */
    LCSIM_B200_ERR_STATE = -3, /* call order (e.g. step before upload_static) */
    LCSIM_B200_ERR_NOMEM = -4
} lcsim_b200_status;

/* reference: lcsim/entity/agent/agent.py:29-33 Agent.Status */
enum { LCSIM_B200_WAITING = 0, LCSIM_B200_ACTIVE = 1, LCSIM_B200_FINISHED = 2, LCSIM_B200_IDLE = 3 };
/* reference: lcsim/entity/agent/policy/type.py:79-84 Action.ActionType (0 = "no external action") */
enum { LCSIM_B200_ACT_NONE = 0, LCSIM_B200_ACT_BICYCLE = 1, LCSIM_B200_ACT_DELTA = 2, LCSIM_B200_ACT_STATE = 3 };

/* per-scenario episode statistics (int64), summed over ranks by the caller (one NCCL all-reduce);
 * the distributed form of AgentManager.get_collision_rate / get_offroad_rate, manager.py:329-345 */
enum {
    LCSIM_B200_STAT_AGENT_STEPS = 0, /* sum over ticks of len(active_agents) at update time */
    LCSIM_B200_STAT_ACTIVE_SUM = 1, /* sum over ticks of len(active_agents) after the tick */
    LCSIM_B200_STAT_COLLISION_SUM = 2, /* sum over ticks of check_collision_all().sum() */
    LCSIM_B200_STAT_OFFROAD_SUM = 3, /* sum over ticks of check_offroads().sum() */
    LCSIM_B200_STAT_ARRIVALS = 4, /* agents that went FINISHED -> IDLE */
    LCSIM_B200_STAT_DEPARTURES = 5, /* agents that went WAITING -> ACTIVE */
    LCSIM_B200_STAT_TICKS = 6,
    LCSIM_B200_STAT_RESERVED = 7
};

typedef struct lcsim_b200_engine lcsim_b200_engine;

/* Shape and control of one batch of independent scenarios.
 * reference: the `control` block of lcsim/config/example.yml read at lcsim/engine/engine.py:59-86 */
typedef struct {
    int32_t abi_version; /* LCSIM_B200_ABI_VERSION */
    int32_t S; /* scenarios in this batch (this rank's shard) */
    int32_t A; /* max agents per scenario (<= 1024) */
    int32_t T; /* max reference-trajectory length */
    int32_t E; /* max resampled road-edge points per scenario */
    double start, interval; /* control.step.start / interval */
    int32_t total_steps; /* control.step.total */
    int32_t reactive; /* control.reactive: TrajIDM instead of StateExpert (agent.py:167-171) */
    int32_t no_static; /* control.no_static (manager.py:76-77) */
    int32_t allow_new_obj; /* control.allow_new_obj (engine.py:208-210) */
    int32_t device; /* CUDA device ordinal */
    int32_t with_metrics; /* 1: every tick also fills collision counts / nearest edge / off-road */
    float grid_cell; /* road-edge hash-grid cell in metres; <= 0 selects the default (4 m) */
} lcsim_b200_batch_desc;

/* Packed static inputs, host pointers, C-contiguous, exactly the arrays of
 * lcsim_b200.packer.StaticBatch.  What Engine.__init__ derives from map.pb / agents.pb
 * (reference: engine.py:84-130, agent/agent.py:106-171, junction/junction.py:59-78,374-380). */
typedef struct {
    const int32_t* n_agents; /* [S] */
    const int32_t* ads_index; /* [S] */
    const int32_t* agent_type; /* [S][A] */
    const double* length; /* [S][A] */
    const double* width; /* [S][A] */
    const uint8_t* dynamic; /* [S][A] Agent.dynamic_flag */
    const int32_t* traj_len; /* [S][A] */
    const int32_t* depart_tick; /* [S][A] first tick j with departure_time <= t_j (float64 clock) */
    const int32_t* wait_order; /* [S][A] agent indices sorted by departure time (stable) */
    const double* home_xy; /* [S][A][2] */
    const double* traj_xy; /* [S][A][T][2] */
    const double* traj_vel; /* [S][A][T][2] */
    const double* traj_h; /* [S][A][T] */
    const int32_t* n_edge; /* [S] */
    const double* edge_xy; /* [S][E][2] resampled road-edge points, pb order */
    const double* edge_dir; /* [S][E][2] unit direction of the segment ending at the point */
    const double* origin; /* [S][2] origin of the scenario-local fp32 frame */
    const int32_t* edge_poly_id; /* [S][E] ordinal of the road-edge polyline every point belongs to (the `ids` of
                                  * Junction.roadgraph_points, junction.py:133-144), or NULL: guidance / plan metrics
                                  * that need the prior-point rule are then unavailable */
} lcsim_b200_static_host;

/* Borrowed device pointers into the engine's state (valid until destroy).  A_stride >= A is the
 * padded agent pitch of every per-agent array.
 * reference: Agent.runtime / Agent.history_states / Trajectory cursor (agent.py:35-84, type.py:10-71),
 * AgentManager.active_agents (manager.py:22), and the metric aggregates (manager.py:313-345). */
typedef struct {
    int32_t S, A, A_stride, T, E, H;
    int32_t* tick; /* [S] Engine.current_step */
    int32_t* n_active; /* [S] */
    int32_t* active; /* [S][A_stride] active_agents, list order */
    int32_t* status; /* [S][A_stride] */
    int32_t* cursor; /* [S][A_stride] Trajectory.index */
    int32_t* ring_count; /* [S][A_stride] valid history entries (<= H) */
    int32_t* ring_head; /* [S][A_stride] next slot to write (circular) */
    double* pose64; /* [S][A_stride][4] x, y, heading, v (world frame, master state) */
    float* pose32; /* [S][A_stride][4] x - origin_x, y - origin_y, heading, v */
    double* ring; /* [S][A_stride][H][5] x, y, vx, vy, heading; circular, oldest at ring_head */
    uint8_t* lead_valid; /* [S][A_stride] Observation.ahead_vehicle is not None */
    double* lead_dis; /* [S][A_stride] */
    double* lead_v; /* [S][A_stride] */
    int32_t* coll_count; /* [S][A_stride] check_collision_all() count, 0 for inactive */
    int32_t* nearest_edge; /* [S][A_stride] index into the concatenated edge points, -1 inactive */
    uint8_t* offroad; /* [S][A_stride] */
    int32_t* traj_len; /* [S][A_stride] current (after re-plans) */
    double* traj_xy; /* [S][A_stride][T][2] */
    double* traj_vel; /* [S][A_stride][T][2] */
    double* traj_h; /* [S][A_stride][T] */
    int64_t* stats; /* [S][LCSIM_B200_N_STATS] */
    double* origin; /* [S][2] */
} lcsim_b200_state_views;

int lcsim_b200_abi_version(void);

/* Allocates the device state for a batch.  Replaces constructing S `Engine(config)` objects
 * (engine.py:53-130).  Fails with LCSIM_B200_ERR_CUDA when no CUDA device is usable. */
int lcsim_b200_create(const lcsim_b200_batch_desc* desc, lcsim_b200_engine** out);

/* Uploads the packed scenarios, builds the fp32 filter copies and the road-edge hash grid, and
 * resets every scenario.  Replaces the map / agent construction in Engine.__init__
 * (engine.py:90-130).  Synchronous (init-time). */
int lcsim_b200_upload_static(lcsim_b200_engine* e, const lcsim_b200_static_host* st);

/* Engine.reset (engine.py:132-143) for the scenarios whose byte in reset_mask (device, [S]) is
 * non-zero; NULL resets all. */
int lcsim_b200_reset(lcsim_b200_engine* e, const uint8_t* reset_mask, void* stream);

/* Engine.step(actions) (engine.py:145-158) for every scenario: update -> departures/arrivals ->
 * prepare_obs -> clock, plus (desc.with_metrics) check_collision_all / check_offroads
 * (manager.py:313-345).  External actions (device pointers, may be NULL when K == 0):
 *   act_agent [S][K] agent index inside the scenario (-1 = unused entry)
 *   act_type  [S][K] LCSIM_B200_ACT_*
 *   act_data  [S][K][5] BICYCLE: acc, steer | DELTA: dx, dy, dheading | STATE: x, y, vx, vy, heading
 * Agents without an entry use their own policy (manager.py:111-115). */
int lcsim_b200_step(lcsim_b200_engine* e, const int32_t* act_agent, const int32_t* act_type, const double* act_data,
                    int K, void* stream);

/* Same for the scenarios whose byte in step_mask (device, [S]) is non-zero only; the others keep their state and
 * clock.  This is `engine_list[i].step(actions)` of lcsim/envs/gym_warpper.py:78-87, where every scenario of a worker
 * is an Engine of its own and only the current one advances.  NULL steps all. */
int lcsim_b200_step_masked(lcsim_b200_engine* e, const uint8_t* step_mask, const int32_t* act_agent, const int32_t* act_type,
                           const double* act_data, int K, void* stream);

/* n_ticks consecutive Engine.step() calls without external actions (engine.py:145-158), i.e. the reference's
 * `for _ in range(n): engine.step()` loop for every scenario of the batch; state and outputs afterwards are those
 * of the last tick.  One launch per tick.  When the whole batch is co-resident on the device, consecutive ticks go to
 * two alternating internal streams and the CTA of a scenario starts as soon as that scenario's previous tick is done
 * (a release/acquire flag per scenario), so the next tick fills the slots idle behind the slowest scenarios; larger
 * batches run as two sub-batches on separate streams.  The caller's stream waits for all of it.  Environment of
 * create(): LCSIM_B200_PIPELINE=0 / LCSIM_B200_GROUPS=1 switch these off, LCSIM_B200_FUSED=1 selects the fused
 * kernel in which the CTA of a scenario runs all n_ticks ticks back to back (measured slower on B200). */
int lcsim_b200_rollout(lcsim_b200_engine* e, int n_ticks, void* stream);

/* Same, and additionally records per tick and scenario what the reference's caller could have read between the
 * steps: tick_log (device, int32 [n_ticks][S][4]) = agents updated by the tick, len(active_agents) after it,
 * sum(check_collision_all()) and the number of off-road agents (manager.py:313-345); needs desc.with_metrics. */
int lcsim_b200_rollout_log(lcsim_b200_engine* e, int n_ticks, int32_t* tick_log, void* stream);

/* A whole rollout with HOST buffers, one call: Engine.reset() for the scenarios whose byte in reset_mask_host is
 * non-zero (NULL = all; engine.py:132-143), n_ticks x Engine.step() (engine.py:145-158), then everything a caller of
 * the reference reads back from the engines, copied to host memory (any pointer may be NULL; [S][A_stride] pitch):
 * the per-tick log of lcsim_b200_rollout_log, float32 poses, status, check_collision_all() counts, nearest edge
 * point, off-road flags and the batch statistics.  H2D: the mask; D2H: the outputs; sync != 0 waits for the stream.
 * Use pinned host memory for asynchronous copies. */
typedef struct {
    int32_t* tick_log; /* [n_ticks][S][4] */
    float* pose32; /* [S][A_stride][4] */
    int32_t* status; /* [S][A_stride] */
    int32_t* coll_count;
    int32_t* nearest_edge;
    uint8_t* offroad;
    int64_t* stats; /* [LCSIM_B200_N_STATS] */
} lcsim_b200_rollout_out;
int lcsim_b200_rollout_host(lcsim_b200_engine* e, const uint8_t* reset_mask_host, int n_ticks,
                            const lcsim_b200_rollout_out* out, int sync, void* stream);

/* Agent.set_ref_trajectory (agent.py:462-466) as called from Engine._plan_trajectory
 * (engine.py:286-297): n (scenario, agent) pairs receive float32 (T,2) xy, (T,2) vel, (T,) heading
 * straight from device memory; cursor <- 0.  All pointers device. */
int lcsim_b200_set_ref_trajectory(lcsim_b200_engine* e, const int32_t* scenario, const int32_t* agent, const float* xy,
                                  const float* vel, const float* heading, int n, int T, void* stream);

/* Engine._compute_rewards (engine.py:301-357) + Agent.get_route (agent.py:468-480) +
 * AgentManager.check_collision / Agent.is_offroad for one agent per scenario (agent [S] device,
 * NULL = the ADS).  Outputs (device, any may be NULL):
 *   terms [S][6] float64, terminated [S] u8, truncated [S] u8 (engine.py:197-200),
 *   route [S][10][2] float64, reward [S] float64 = sum coef6[k] * terms[k] (coef6 host, may be NULL). */
int lcsim_b200_rewards(lcsim_b200_engine* e, const int32_t* agent, const double* coef6, double* reward, double* terms,
                       uint8_t* terminated, uint8_t* truncated, double* route, void* stream);

/* Ego-centric neighbour gather feeding lcsim.envs observations.  For one ego per scenario:
 *  - a2a: the other ACTIVE agents whose last history position lies within `radius` of the ego's
 *    (motion_diff/modules/agent_encoder.py:235-242 torch_cluster.radius_graph call-site contract: float32
 *    positions, strict d^2 < r^2, no self loop, ascending source order = active-list order here, at most
 *    k_agents kept), each with the reference's relation feature r_a2a = (|d|, angle_between(head_ego, d),
 *    wrap(h_nbr - h_ego)) (:243-255), d = pos_nbr - pos_ego;
 *  - pl2a: the polyline centres within `radius` of the ego (:208-216 torch_cluster.radius), same feature
 *    with the polyline heading atan2(dir_y, dir_x) (:217-231);
 *  - sort_by_distance = 1 returns the k NEAREST instead of the first k in index order (top-k select);
 *  - history: the (N,11,.) float32 stack of junction/junction.py:301-334 for the active list, row i = list slot i.
 * The neighbour arithmetic of the reference lives in torch_cluster (not vendored): parity at that boundary is
 * pinned only against tests/obs_oracle.py, a numpy restatement of the call-site contract. */
typedef struct {
    const int32_t* ego; /* [S] device agent index, NULL = the ADS */
    double radius; /* 50.0 in the reference */
    int32_t k_agents; /* capacity of the agent-neighbour list per ego (<= 1024) */
    int32_t k_polys; /* capacity of the polyline-neighbour list per ego (<= 1024) */
    int32_t sort_by_distance; /* 0: ascending index (torch_cluster order); 1: nearest first */
    /* outputs, device, caller-owned; any may be NULL */
    int32_t* nbr_agent; /* [S][k_agents] agent index, -1 padded */
    float* nbr_agent_feat; /* [S][k_agents][3] */
    int32_t* n_nbr_agent; /* [S] neighbours found (before the cap) */
    int32_t* nbr_poly; /* [S][k_polys] polyline-centre index, -1 padded */
    float* nbr_poly_feat; /* [S][k_polys][3] */
    int32_t* n_nbr_poly; /* [S] */
    float* history; /* [S][A_stride][H][6] x, y, vx, vy, heading, valid; shift-register order, rows = list slots */
    int32_t* history_agent; /* [S][A_stride] agent index of every history row, -1 beyond n_active */
} lcsim_b200_obs_desc;

/* polyline centres of the roadgraph (lane/lane.py:108-124, junction/junction.py:117-184 ->
 * polygon.py:11-75): host [S][P][3] = x, y, heading; packer.polyline_centres() computes them */
int lcsim_b200_upload_polylines(lcsim_b200_engine* e, const int32_t* n_poly, const double* poly_xyh, int P);
int lcsim_b200_build_obs(lcsim_b200_engine* e, const lcsim_b200_obs_desc* obs, void* stream);

/* Encoder inputs for ALL listed agents of every scenario: what AgentEncoder.forward derives from the
 * HeteroData in front of its attention layers (motion_diff/modules/agent_encoder.py:98-256), written from the engine's
 * state so that the model needs no torch_cluster call.  Agent rows = slots of the active list.  Outputs (device,
 * caller-owned, any may be NULL):
 *   agent       [S][A_stride]         agent index of every row, -1 beyond n_active
 *   x_a         [S][A_stride][H][6]   |motion|, angle(head, motion), |vel|, angle(head, vel), length, width (:153-166)
 *   r_t         [S][A_stride][H-1][4] temporal edge (row, i) -> (row, H-1): |dpos|, angle(head_dst, dpos), wrap(dhead),
 *               i - (H-1); r_t_valid [S][A_stride][H-1] says which exist (both steps valid, H-1-i <= time_span; :174-199)
 *   a2a         CSR over target rows: a2a_offset [S][A_stride+1], a2a_src [S][cap_a2a] source rows ascending (radius_graph,
 *               loop=False, d^2 < r^2, at most 300 per target; :233-242), a2a_feat [S][cap_a2a][3] = |d|,
 *               angle(head_dst, d), wrap(head_src - head_dst) with d = pos_src - pos_dst (:243-255); n_a2a [S] edges found
 *   pl2a        CSR over polyline centres: pl2a_offset [S][P+1], pl2a_agent [S][cap_pl2a] agent rows ascending (radius,
 *               at most 300 per polyline; :202-216), pl2a_feat [S][cap_pl2a][3] = |d|, angle(head_a, d),
 *               wrap(head_pl - head_a) with d = pos_pl - pos_a (:217-231); n_pl2a [S]
 * Edges beyond a capacity are counted but not stored.  The neighbour arithmetic of the reference lives in torch_cluster
 * (not vendored): restated from its call-site contract, see lcsim_b200_obs_desc. */
typedef struct {
    double a2a_radius, pl2a_radius; /* 50, 50 in motion_diff/configs/config.yml */
    int32_t time_span; /* 10 */
    int32_t cap_a2a, cap_pl2a; /* edge capacities per scenario */
    int32_t* agent;
    float* x_a;
    float* r_t;
    uint8_t* r_t_valid;
    int32_t* a2a_offset;
    int32_t* a2a_src;
    float* a2a_feat;
    int32_t* n_a2a;
    int32_t* pl2a_offset;
    int32_t* pl2a_agent;
    float* pl2a_feat;
    int32_t* n_pl2a;
} lcsim_b200_encoder_desc;
int lcsim_b200_build_encoder_inputs(lcsim_b200_engine* e, const lcsim_b200_encoder_desc* d, void* stream);

/* One environment step for ALL S scenarios with HOST buffers: the batched form of
 * BicycleSingleEnv.step (lcsim/envs/gym_warpper.py:78-87).  actions_host [S][2] are the normalised
 * (acc, steer) in [-1, 1] of every scenario's ADS -> BicycleAction.init_with_normalized (x6.0, x0.3;
 * policy/type.py:100-102) -> Engine.step({ads: action}) -> Engine.get_step_return's reward, terminated,
 * truncated, reward terms and route (engine.py:160-201).  NULL actions let the ADS follow its own policy.
 * The call enqueues H2D copy, tick, reward kernel and D2H copy on `stream`; with sync != 0 it also waits for
 * the stream, so out_host is readable on return.  Use pinned host memory for truly asynchronous copies. */
typedef struct {
    double reward; /* sum coef6[k] * terms[k] (0 when coef6 is NULL) */
    double terms[LCSIM_B200_N_REWARD_TERMS];
    double route[LCSIM_B200_ROUTE_LEN][2];
    uint8_t terminated, truncated, finished, pad[5]; /* finished: the ADS is IDLE/FINISHED (obs are zeros then) */
} lcsim_b200_env_out;
int lcsim_b200_env_step(lcsim_b200_engine* e, const double* actions_host, const double* coef6,
                        lcsim_b200_env_out* out_host, int sync, void* stream);
/* The same step with DEVICE buffers (a policy that lives on the GPU): actions_dev [S][2] float64 or NULL, out_dev [S]
 * records; nothing is copied and the call never waits for the stream. */
int lcsim_b200_env_step_device(lcsim_b200_engine* e, const double* actions_dev, const double* coef6,
                               lcsim_b200_env_out* out_dev, void* stream);
/* Registers the observation build of env_step: BicycleSingleEnv.step returns the observation with the reward
 * (gym_warpper.py:78-92 -> Engine.get_step_return, engine.py:160-201), so after this call every lcsim_b200_env_step
 * also runs lcsim_b200_build_obs(obs) behind its tick (on an engine-owned side stream, concurrently with the reward
 * kernel; joined before the D2H copy, so a sync != 0 covers it).  The descriptor is copied; its device output buffers
 * stay caller-owned and must outlive the registration.  obs == NULL removes it. */
int lcsim_b200_set_env_obs(lcsim_b200_engine* e, const lcsim_b200_obs_desc* obs);

/* Lane-centreline projection.  The centreline points are Junction.roadgraph_points restricted to
 * the centreline types 1|2 (junction/junction.py:186-253 <- lane/lane.py:108-124: the 0.5 m-resampled lane
 * centrelines, float32, lane order); upload them once: xy_theta [S][N][3] = x, y, atan2(dir_y, dir_x) (float32,
 * world frame like the reference's tensors), ids [S][N] = polyline (lane) id of every point.
 * packer.centreline_points() builds both from a scenario. */
int lcsim_b200_upload_centrelines(lcsim_b200_engine* e, const int32_t* n_pts, const float* xy_theta, const int32_t* ids,
                                  int N);

/* For every ACTIVE agent (query = float32 world position and heading of its pose), outputs [S][A_stride], device,
 * caller-owned, any may be NULL; inactive agents get -1 / +inf / -1:
 *   lane_pt      first-index argmin over the centreline points of the float32 squared distance
 *   lane_id      ids[lane_pt]: the lane the agent is on (the integer lane id a WOMD agent lacks in the reference)
 *   lane_dist    sqrt of that minimum
 *   onlane_dist  OnLane.loss's per-query term (motion_diff/modules/guidance.py:100-133): min distance over the
 *                points with |wrap(heading - theta)| < 0.2 and distance < 5.0, +inf if none
 *   heading_diff compute_lane_heading_diff's per-query term (motion_diff/metrics/roadgraph.py:146-155): min
 *                |wrap(heading - theta)| over the top_k (64 in the reference, <= 64) nearest points by squared
 *                distance; -1 when the scenario has fewer than top_k centreline points (the reference skips it). */
typedef struct {
    int32_t top_k;
    int32_t* lane_pt;
    int32_t* lane_id;
    float* lane_dist;
    float* onlane_dist;
    float* heading_diff;
} lcsim_b200_lane_desc;
int lcsim_b200_project_lanes(lcsim_b200_engine* e, const lcsim_b200_lane_desc* d, void* stream);

/* Guidance costs over planned trajectories with their analytic gradients, float32 like the reference's
 * tensors; every scenario is one graph.  xy (device) [S][A][T][2] world-frame positions of the plans (what
 * diffuser._reconstruct_traj returns), rows a >= n_agents[s] ignored.  Outputs (device, may be NULL): loss [S], grad
 * [S][A][T][2] = d loss / d xy as torch.autograd gives it for the loss expression.
 *   which = 0  Repeller.loss  (motion_diff/modules/guidance.py:17-49; radius 6, eps 1e-6 in the reference)
 *   which = 1  NoOffRoad.loss (guidance.py:52-85 -> metrics/roadgraph.py:160-232; radius 1, eps 1e-6): needs
 *              static.edge_poly_id; sd (device [S][A][T], may be NULL) receives the signed distances */
int lcsim_b200_guidance_cost(lcsim_b200_engine* e, int which, const float* xy, int T, float radius, float eps, float* loss,
                             float* grad, float* sd, void* stream);

/* Batched plan metrics: traj5 (device) [S][A][T][5] = x, y, length, width, yaw; masks (device, [S][A]
 * u8, NULL = every agent) select the agents that are counted.  Outputs (device, any may be NULL), per scenario:
 *   overlap_rate / overlap_count  compute_overlap_rate (motion_diff/metrics/overlap.py:7-43: boxes of ALL agents of a
 *                                 step are tested pairwise, metrics/geometry.py:8-101; an agent counts when it
 *                                 overlaps anybody), rate = count / sum(mask) / T
 *   offroad_rate / offroad_count  compute_offroad_rate (metrics/roadgraph.py:36-88; the caller ANDs the vehicle mask
 *                                 into offroad_mask), signed distance > 0 */
int lcsim_b200_plan_metrics(lcsim_b200_engine* e, const float* traj5, int T, const uint8_t* overlap_mask,
                            const uint8_t* offroad_mask, float* overlap_rate, int32_t* overlap_count, float* offroad_rate,
                            int32_t* offroad_count, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Scenario loader / packer: reads `map.pb` + `agents.pb` of n_dirs WOMD-shaped scenario directories and
 * packs them for lcsim_b200_upload_static — the native form of what Engine.__init__ does per scenario (engine.py:84-130:
 * protobuf parse; junction/junction.py:59-78 -> utils/polygon.py:90-188 0.5 m resampling of the road edges;
 * polygon.py:78-86 direction vectors; agent/agent.py:106-171 homes, trajectories, dynamic flags; manager.py:33-35
 * waiting order; lane/lane.py:108-124 polyline centres), threaded over the scenarios (n_threads <= 0: all cores).
 * Bit-identical to lcsim_b200/packer.py, which is pinned to the real reference.  Host memory only; no GPU needed. */
typedef struct lcsim_b200_scenarios lcsim_b200_scenarios;
int lcsim_b200_load_dirs(const char* const* dirs, int n_dirs, double start, double interval, int32_t total_steps, int n_threads,
                         lcsim_b200_scenarios** out);
/* dims[5] = S, A, T, E, P (largest number of polyline centres) */
int lcsim_b200_scenarios_dims(const lcsim_b200_scenarios* sc, int32_t* dims);
/* the packed arrays (owned by sc, valid until lcsim_b200_scenarios_free), ready for lcsim_b200_upload_static */
int lcsim_b200_scenarios_static(const lcsim_b200_scenarios* sc, lcsim_b200_static_host* out);
const int32_t* lcsim_b200_scenarios_agent_id(const lcsim_b200_scenarios* sc); /* [S][A] pb ids, -1 padded */
const int32_t* lcsim_b200_scenarios_junction_id(const lcsim_b200_scenarios* sc); /* [S] */
/* polyline centres for lcsim_b200_upload_polylines: *n_poly [S], return [S][P][3] x, y, heading */
const double* lcsim_b200_scenarios_polylines(const lcsim_b200_scenarios* sc, const int32_t** n_poly);
const char* lcsim_b200_scenarios_error(const lcsim_b200_scenarios* sc);
void lcsim_b200_scenarios_free(lcsim_b200_scenarios* sc);

int lcsim_b200_views(lcsim_b200_engine* e, lcsim_b200_state_views* out);

/* Sums the per-scenario statistics into stats_device[LCSIM_B200_N_STATS] (int64, device). */
int lcsim_b200_episode_stats(lcsim_b200_engine* e, int64_t* stats_device, void* stream);

/* number of tick-kernel launches since create (bench.py's gpu_launches) */
int64_t lcsim_b200_launch_count(const lcsim_b200_engine* e);

const char* lcsim_b200_last_error(const lcsim_b200_engine* e);
void lcsim_b200_destroy(lcsim_b200_engine* e);

/* ---------------------------------------------------------------------------------------------------------------
 * City engine: ONE scenario with lane-position agents driven by LaneIDM on the lane graph (SURVEY.md A13,
 * BASELINE configs[3]: a MOSS-shaped city with up to 200k vehicles).  Replaces, for such a scenario, Engine.reset /
 * Engine.step (engine.py:132-158,203-225) with AgentManager.update / check_departure_and_arrival / prepare_obs
 * (manager.py:50-115), the LaneIDM branch of Agent.prepare_obs (agent.py:198-243), LaneIDM.get_action
 * (policy/idm_policy.py:53-116), the LANE branch of update_runtime_by_action (agent.py:314-376), Lane.* (lane/lane.py:
 * 171-300), Route.* (agent/route.py:102-203), check_collision_all (manager.py:313-327, via a spatial hash) and
 * Agent.is_offroad (agent.py:482-494), including the forced-lane-change branch (idm_policy.py:64-84,118-167,
 * agent.py:331-368) with its order-dependent reads of already-updated neighbours.  scalars[6] tells that a lane
 * change was started, scalars[7] that the reference would have raised (target lane missing / not on the same road).
 * Static arrays: lcsim_b200.city.CityStatic / pack_city (dense indices in pb order), host pointers. */
typedef struct {
    double start, interval;
    int32_t n_lanes, n_agents, n_esets, n_groups;
    const double* lane_length; /* [NL] Lane.length */
    const double* lane_max_speed;
    const int32_t* lane_road; /* [NL] parent road index or -1 */
    const int32_t* lane_junction; /* [NL] parent junction index or -1 */
    const int32_t* lane_offset; /* [NL] offset_in_road */
    const int32_t* lane_succ0; /* [NL] first successor (unique_successor of junction lanes) or -1 */
    const int32_t* lane_left; /* [NL] Lane.left_neighbor() or -1 */
    const int32_t* lane_right; /* [NL] Lane.right_neighbor() or -1 */
    const double* lane_width; /* [NL] */
    const int32_t* lane_node_start; /* [NL+1] */
    const double* node_xy; /* [NN][2] raw centre-line nodes */
    const double* node_cum; /* [NN] Lane.line_lengths */
    const double* node_dir; /* [NN] Lane.line_directions of the segment ending at the node */
    const int32_t* lane_eset; /* [NL] edge set Agent.is_offroad checks on this lane */
    const int32_t* eset_start; /* [n_esets+1] */
    const double* edge_xy; /* [NE][2] */
    const double* edge_dir; /* [NE][2] */
    const int32_t* grp_start; /* [n_groups+1] junction lane groups (junction.py:90-103) */
    const int32_t* grp_lane; /* junction lane index */
    const int32_t* grp_pre_offset; /* offset_in_road of the lane's first predecessor (route.py:93-99) */
    const double* length; /* [N] */
    const double* width;
    const double* idm; /* [N][5] usual_braking_a, max_braking_a, max_a, max_v, min_gap (idm_policy.py:42-51) */
    const int32_t* home_lane;
    const double* home_s;
    const int32_t* end_road; /* road index of trip.end's lane */
    const int32_t* end_offset; /* its offset_in_road */
    const double* end_s;
    const int32_t* depart_tick; /* first tick j with departure_time <= t_j (float64 clock) */
    const int32_t* wait_order; /* agents sorted by departure time (stable) */
    const int32_t* route_start; /* [N+1] into the concatenated road lists */
    const int32_t* route_grp; /* group index of every consecutive road pair: hop h of agent a = route_grp[route_start[a] - a + h] */
} lcsim_b200_city_static;

typedef struct {
    int32_t N, n_lanes, H;
    int32_t* scalars; /* [16] 0: n_active, 1: waiting-list pointer, 2: current_step, 6: a lane change was started, 7: lane-change error */
    int32_t* is_lc; /* [N] LaneIDM.lc_runtime.is_lc */
    int32_t* active; /* [N] active_agents, list order */
    int32_t* status; /* [N] */
    int32_t* lane; /* [N] runtime.lane (dense index) */
    double* s; /* [N] runtime.s */
    double* v;
    double* xy; /* [N][2] */
    double* heading;
    uint8_t* lead_valid; /* obs.ahead_vehicle is not None */
    double* lead_dis;
    double* lead_v;
    int32_t* coll_count; /* check_collision_all() count, 0 for inactive */
    uint8_t* offroad; /* Agent.is_offroad() */
    double* ring; /* [N][H][5] circular history */
    int32_t* ring_head;
    int32_t* ring_count;
    int64_t* stats; /* [LCSIM_B200_N_STATS] */
} lcsim_b200_city_state_views;

typedef struct lcsim_b200_city lcsim_b200_city;
int lcsim_b200_city_create(const lcsim_b200_city_static* st, int device, int with_metrics, lcsim_b200_city** out);
int lcsim_b200_city_reset(lcsim_b200_city* e, void* stream);
int lcsim_b200_city_step(lcsim_b200_city* e, int n_ticks, void* stream);
int lcsim_b200_city_views(lcsim_b200_city* e, lcsim_b200_city_state_views* out);
int64_t lcsim_b200_city_launch_count(const lcsim_b200_city* e);
const char* lcsim_b200_city_last_error(const lcsim_b200_city* e);
void lcsim_b200_city_destroy(lcsim_b200_city* e);

#ifdef __cplusplus
}
#endif
#endif /* LCSIM_B200_H */
