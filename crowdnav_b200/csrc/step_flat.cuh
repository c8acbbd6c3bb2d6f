// step_flat.cuh -- CrowdSim step for small crowds (N <= 5 humans), register-resident ORCA solver.
//
// Same contract as step_kernel (crowd_sim/envs/crowd_sim.py:317-420 + orca.py:82-132 + explorer.py:41-72).
// Mapping: one thread per (env, agent) solve, L = N + 1 lanes per env, floor(32 / L) whole envs per warp so an env
// never straddles a warp: all intra-env exchange (candidate positions/velocities/radii for the neighbour scan, the
// robot's position and action for the swept-segment test, the per-human clearances for the min / any reduction) is
// done with warp shuffles -- no shared-memory staging, no block barrier on the common path.
// The <= N ORCA lines of a solve live in REGISTERS: every loop over lines is fully unrolled (template on N), so
// the LP code has static register indexing, no local/shared memory traffic and instruction-level parallelism
// across the independent (i, j) line pairs.
// linearProgram3 (needed by ~4.6 % of the solves, i.e. by some lane of ~3 of 4 warps) is NOT run in place: the
// solves that need it are compacted into a shared-memory queue. Inside linearProgram3 the sub-problem of each line i
// (linearProgram2 over the lines projected onto i, started from optVelocity * radius) depends only on the lines, not
// on the running result, so the <= N-1 sub-problems of a queued solve run on N-1 LANES IN PARALLEL (the sequential
// shared-memory LP code of orca_device.cuh), followed by a 4-step scan. Before this, the pass was a ~1500-instruction
// serial chain on a handful of lanes with the whole block waiting: ~4.5 of 14 us per launch.
// The queue is per BLOCK (WARPQ = false: one warp runs the pass for the whole block, the others wait at a barrier;
// fewest instructions, best when the launch fills the chip) or per WARP (WARPQ = true: no block barrier, every warp runs
// the pass for its own 1-2 solves; best for launches that leave the SMs mostly empty). cs::launch() picks by grid size.
//
// Evidence that motivated this design (profiles/r01_*): one-thread-per-agent with shared-memory lines ran at 13.6/32
// active lanes and 3480 instructions per warp; the warp-per-env cooperative variant needed 1750 warp instructions
// per env. This kernel needs 483 per env (2.4 k per warp of 5 envs) at 23/32 active lanes.
#pragma once
#include "crowdsim_common.cuh"
#include "orca_spec.cuh"

namespace cs {

#define CS_FULL 0xffffffffu

// EPW = 32 / (N + 1) whole envs per warp (dense packing; sparser packings were measured and are never faster,
// profiles/r01_tune_epw_n5.txt). STAGE is a profiling aid (scripts/latency_probe.cu instantiates cut-down variants to
// attribute latency); the library only instantiates the full kernel (STAGE = 99).
// Register budget: asking for 6 resident blocks per SM (<= 80 registers, a few bytes of spill) is neutral at 4096 envs and
// 9 % faster at 65 k .. 1 M envs than the unconstrained 90-register build (scripts/latency_probe.cu, -DCS_FLAT_MINBLOCKS=1/6/8).
// CS_FLAT_WPB = warps per block, CS_FLAT_MINBLOCKS = resident-blocks hint (per 128 threads), CS_FLAT_WARP_LP3 = default of the
// WARPQ template parameter: build-time knobs for A/B runs (scripts/gpu_variants.sh); 4 / 6 / size-dependent were kept.
// EXPERIMENT: -DCS_FLAT_NO_ROT compiles the unicycle-robot trigonometry out of the small-crowd kernel
// (~12 % of its SASS, never executed by holonomic robots) to measure what the instruction-fetch stalls cost; such a build
// only serves holonomic / ORCA robots. Default: runtime test, as the reference's agent.py:115-135.
#ifdef CS_FLAT_NO_ROT
#define CS_FLAT_IS_ROT(k) false
#else
#define CS_FLAT_IS_ROT(k) ((k).robot_policy == CROWDSIM_ROBOT_EXTERNAL_ROT)
#endif
#ifndef CS_FLAT_WPB
#define CS_FLAT_WPB 4
#endif
#ifndef CS_FLAT_WARP_LP3
#define CS_FLAT_WARP_LP3 0
#endif
#ifndef CS_FLAT_MINBLOCKS
#define CS_FLAT_MINBLOCKS 6
#endif
template <int N, int STAGE = 99, bool WARPQ = (CS_FLAT_WARP_LP3 != 0)>
__global__ void __launch_bounds__(32 * CS_FLAT_WPB, CS_FLAT_MINBLOCKS * 4 / CS_FLAT_WPB) step_flat_kernel(const __grid_constant__ StepArgs A)
{
    if constexpr (STAGE == 0) return;
    using namespace orca;
    constexpr int L = N + 1, M = N, EPW = 32 / L, WPB = CS_FLAT_WPB, T = 32 * WPB;
    constexpr int SUB = (M > 1) ? M - 1 : 1;                // lanes per queued lp3 item (sub-problems i = 1 .. M-1)
    constexpr int QF = 4 * M + 5;                           // floats per queued lp3 work item
    __shared__ float s_q[QF][T];                            // [field][slot]: lines of an item = orca::Lines(base = &s_q[0][slot], stride = T)
    __shared__ float s_p[4 * SUB][T];                       // per-thread projected lines of the sub-problem
    __shared__ float s_r2[3][T];                            // per-thread sub-problem result (x, y, ok)
    __shared__ float s_res[2][T];
    __shared__ int s_qcount;

    const KParams &k = A.k;
    const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
    const int le = lane / L, a = lane - le * L;             // env within the warp, agent within the env
    const int ebase = le * L;                               // first lane of my env
    const int e = (blockIdx.x * WPB + wib) * EPW + le;
    const bool is_robot = (a == N);
    const bool env_ok = (le < EPW) && (e < A.B);
    if (tid == 0) s_qcount = 0;

    // ---- all global loads of the step are issued up front, unconditionally for valid envs, so that they overlap into ONE
    // DRAM round trip (active flag -> state -> episode accumulators / slot state used to be dependent ones) ----
    // (idle lanes get a goal 5 m away: a zero goal vector would drag the warp through the f64 sqrt / division slow paths)
    double2 pos = make_double2(0, 0), vel = pos, goal = make_double2(3, 4), attr = make_double2(0.3, 1.0);
    double theta = 0, gtime = 0; double2 ext = make_double2(0, 0);
    uint8_t act_flag = 1, slot_state = 0, want_flag = 0;
    int ep_t = 0, ep_tc = 0, ep_c = -1; double ep_ret = 0, ep_mds = 0;
    if (env_ok) {
        if (A.st.active) act_flag = A.st.active[e];
        if (!is_robot) {
            const size_t i = (size_t)e * N + a;
            pos = ld2(A.st.h_pos, i); vel = ld2(A.st.h_vel, i); goal = ld2(A.st.h_goal, i); attr = ld2(A.st.h_attr, i);
        } else {
            pos = ld2(A.st.r_pos, e); vel = ld2(A.st.r_vel, e); goal = ld2(A.st.r_goal, e); attr = ld2(A.st.r_attr, e);
            gtime = A.st.g_time[e];
            if (CS_FLAT_IS_ROT(k)) theta = A.st.r_theta[e];
            if (k.robot_policy != CROWDSIM_ROBOT_ORCA) ext = ld2(A.io.action, e);
            if (A.has_ep) { ep_t = A.ep.ep_steps[e]; ep_ret = A.ep.ep_return[e]; ep_tc = A.ep.ep_too_close[e]; ep_mds = A.ep.ep_min_dist_sum[e]; ep_c = A.ep.ep_case[e]; }
            if (A.has_ar) { slot_state = *reinterpret_cast<volatile uint8_t *>(A.ar.n_state + e); want_flag = A.ar.want[e]; }
        }
    }
    const bool live = env_ok && (act_flag != 0);
    if constexpr (STAGE == 1) {            // loads + stores only
        if (live && !is_robot) { const size_t i = (size_t)e * N + a; st2(A.st.h_pos, i, pos); st2(A.st.h_vel, i, make_double2(vel.x + goal.x * 0, vel.y + attr.x * 0)); }
        if (live && is_robot) { st2(A.st.r_pos, e, pos); A.st.g_time[e] = gtime + ext.x * 0 + theta * 0; }
        return;
    }
    // float32 view of myself for the other lanes of my env (rvo2 boundary casts, orca.py:100-110)
    const float fpx = (float)pos.x, fpy = (float)pos.y, fvx = (float)vel.x, fvy = (float)vel.y;
    const float frh = (float)(attr.x + 0.01 + k.human_safety_space);     // my radius as seen by a human observer
    const float frr = (float)(attr.x + 0.01 + k.robot_safety_space);     // ... by the robot
    const bool solves = live && (!is_robot || k.robot_policy == CROWDSIM_ROBOT_ORCA);

    // ---- orca.py:113-115 preferred velocity (float64) ----
    const double gvx = goal.x - pos.x, gvy = goal.y - pos.y;
    const double speed = norm2(gvx, gvy);
    const V2 pref = mk((float)((speed > 1) ? gvx / speed : gvx), (float)((speed > 1) ? gvy / speed : gvy));
    const V2 p = mk(fpx, fpy), v = mk(fvx, fvy);
    const float r = is_robot ? frr : frh;
    const float max_speed = (float)attr.y;

    // ---- neighbour scan: candidate slot c -> agent j (reference order: other humans, then the robot iff visible) ----
    float dsq[M]; bool inr[M]; int jj[M];
    #pragma unroll
    for (int c = 0; c < M; ++c) {
        int j; bool cv;
        if (is_robot) { j = c; cv = true; }
        else if (c < N - 1) { j = (c < a) ? c : c + 1; cv = true; }
        else { j = N; cv = (k.robot_visible != 0); }
        jj[c] = j;
        const float qx = __shfl_sync(CS_FULL, fpx, ebase + j), qy = __shfl_sync(CS_FULL, fpy, ebase + j);
        dsq[c] = abssq(p - mk(qx, qy));
        inr[c] = solves && cv && (k.max_neighbors > 0) && dsq[c] < sqr(k.neighbor_dist);
    }
    // rank of each candidate = position RVO2's insertion sort (strict <, ties in scan order) would give it: for cc < c,
    // cc precedes c iff dsq[cc] <= dsq[c] -- one comparison per unordered pair. The agent index of the kk-th nearest is
    // then read from a packed word (3 bits per position, N <= 5) instead of an M x M select cascade; together 7 % fewer
    // instructions per warp than the 2 M^2 compare-and-select form (ncu source view, step_flat.cuh:130/134 before).
    int rank[M];
    #pragma unroll
    for (int c = 0; c < M; ++c) rank[c] = 0;
    #pragma unroll
    for (int c = 1; c < M; ++c) {
        #pragma unroll
        for (int cc = 0; cc < c; ++cc) {
            const bool le = dsq[cc] <= dsq[c];
            rank[c] += (inr[cc] && le) ? 1 : 0;
            rank[cc] += (inr[c] && !le) ? 1 : 0;
        }
    }
    int nl = 0; unsigned packed = 0u;
    #pragma unroll
    for (int c = 0; c < M; ++c) if (inr[c]) { packed |= (unsigned)jj[c] << (3 * rank[c]); ++nl; }
    int src[M];                             // src[kk] = agent index of the kk-th nearest (0 beyond nl)
    #pragma unroll
    for (int kk = 0; kk < M; ++kk) src[kk] = (int)((packed >> (3 * kk)) & 7u);
    nl = nl < k.max_neighbors ? nl : k.max_neighbors;

    // ---- ORCA lines in rank order, in registers ----
    RegLines<M> R; bool valid[M];
    #pragma unroll
    for (int kk = 0; kk < M; ++kk) {
        const int sl = ebase + src[kk];
        const float qx = __shfl_sync(CS_FULL, fpx, sl), qy = __shfl_sync(CS_FULL, fpy, sl);
        const float wx = __shfl_sync(CS_FULL, fvx, sl), wy = __shfl_sync(CS_FULL, fvy, sl);
        const float rh = __shfl_sync(CS_FULL, frh, sl), rr = __shfl_sync(CS_FULL, frr, sl);
        valid[kk] = kk < nl;
        R.p[kk] = mk(0.f, 0.f); R.d[kk] = mk(0.f, 0.f);
        if (valid[kk]) make_line_sel(p, v, r, mk(qx, qy), mk(wx, wy), is_robot ? rr : rh, k.inv_time_horizon, k.inv_time_step, R.p[kk], R.d[kk]);
    }

    // ---- linear programs: lp2 in place, lp3 deferred to the block-compacted pass ----
    if constexpr (STAGE == 2) {            // + preferred velocity, neighbour scan, ORCA lines
        float acc = pref.x + pref.y;
        #pragma unroll
        for (int kk = 0; kk < M; ++kk) acc += R.p[kk].x + R.p[kk].y + R.d[kk].x + R.d[kk].y;
        if (live && !is_robot) { const size_t i = (size_t)e * N + a; st2(A.st.h_vel, i, make_double2(vel.x, vel.y + (double)acc * 0)); }
        return;
    }
    // speculative lp1 candidates for every line (orca_spec.cuh), then linearProgram2 as a scan
    V2 cand[M]; bool feas[M];
    lp1_all<M, M>(R, valid, max_speed, pref, false, cand, feas);
    V2 nv = mk(0.f, 0.f);
    const int fail = lp2_scan<M, M>(R, valid, nl, cand, feas, lp2_init(pref, max_speed), nv);
    if constexpr (STAGE == 3) {            // + lp1 candidates and the lp2 scan
        if (live && !is_robot) { const size_t i = (size_t)e * N + a; st2(A.st.h_vel, i, make_double2(vel.x + (double)nv.x * 0, vel.y + (double)(nv.y + fail) * 0)); }
        return;
    }
    const bool need3 = solves && fail < nl;
    if constexpr (WARPQ) {
        // warp-level queue: no block barrier; every warp runs the sub-problems of its own solves
        const unsigned m3 = __ballot_sync(CS_FULL, need3);
        if (m3) {
            const int wbase = wib * 32;
            const int cnt = __popc(m3);
            const int slot = wbase + __popc(m3 & ((1u << lane) - 1u));
            if (need3) {
                #pragma unroll
                for (int kk = 0; kk < M; ++kk) {
                    s_q[4 * kk + 0][slot] = R.p[kk].x; s_q[4 * kk + 1][slot] = R.p[kk].y;
                    s_q[4 * kk + 2][slot] = R.d[kk].x; s_q[4 * kk + 3][slot] = R.d[kk].y;
                }
                s_q[4 * M + 0][slot] = __int_as_float(nl); s_q[4 * M + 1][slot] = __int_as_float(fail);
                s_q[4 * M + 2][slot] = max_speed; s_q[4 * M + 3][slot] = nv.x; s_q[4 * M + 4][slot] = nv.y;
            }
            __syncwarp();
            constexpr int IPP = 32 / SUB;                        // items per pass
            for (int base = 0; base < cnt; base += IPP) {
                const int item = wbase + base + lane / SUB, i = lane % SUB + 1;
                const bool mine = (lane < IPP * SUB) && (base + lane / SUB) < cnt;
                if (mine) {
                    const Lines Lq = { &s_q[0][item], T };
                    const int qn = __float_as_int(s_q[4 * M + 0][item]);
                    bool ok = false; V2 r2 = mk(0.f, 0.f);
                    if (M > 1 && i < qn) {
                        const Lines Pq = { &s_p[0][tid], T };
                        ok = lp3_subproblem(Lq, i, s_q[4 * M + 2][item], Pq, r2);
                    }
                    s_r2[0][tid] = r2.x; s_r2[1][tid] = r2.y; s_r2[2][tid] = ok ? 1.0f : 0.0f;
                }
                __syncwarp();
                if (mine && i == 1) {
                    const Lines Lq = { &s_q[0][item], T };
                    const int qn = __float_as_int(s_q[4 * M + 0][item]), qf = __float_as_int(s_q[4 * M + 1][item]);
                    const float qr = s_q[4 * M + 2][item];
                    V2 res = mk(s_q[4 * M + 3][item], s_q[4 * M + 4][item]);
                    lp3_outer_scan(Lq, qn, qf, qr, res, [&](int ii, V2 &r2) {
                        const int src = tid + (ii - 1);
                        r2 = mk(s_r2[0][src], s_r2[1][src]);
                        return s_r2[2][src] != 0.0f;
                    });
                    s_res[0][item] = res.x; s_res[1][item] = res.y;
                }
                __syncwarp();
            }
            if (need3) nv = mk(s_res[0][slot], s_res[1][slot]);
        }
    } else {
        __syncthreads();                                         // s_qcount = 0 visible
        int slot = -1;
        if (need3) {
            slot = atomicAdd(&s_qcount, 1);
            #pragma unroll
            for (int kk = 0; kk < M; ++kk) {
                s_q[4 * kk + 0][slot] = R.p[kk].x; s_q[4 * kk + 1][slot] = R.p[kk].y;
                s_q[4 * kk + 2][slot] = R.d[kk].x; s_q[4 * kk + 3][slot] = R.d[kk].y;
            }
            s_q[4 * M + 0][slot] = __int_as_float(nl); s_q[4 * M + 1][slot] = __int_as_float(fail);
            s_q[4 * M + 2][slot] = max_speed; s_q[4 * M + 3][slot] = nv.x; s_q[4 * M + 4][slot] = nv.y;
        }
        if (__syncthreads_or(need3 ? 1 : 0)) {
            const int cnt = s_qcount;
            constexpr int IPP = T / SUB;                         // items per pass
            for (int base = 0; base < cnt; base += IPP) {
                const int item = base + tid / SUB, i = tid % SUB + 1;
                const bool mine = (tid < IPP * SUB) && item < cnt;
                if (mine) {
                    const Lines Lq = { &s_q[0][item], T };
                    const int qn = __float_as_int(s_q[4 * M + 0][item]);
                    bool ok = false; V2 r2 = mk(0.f, 0.f);
                    if (M > 1 && i < qn) {
                        // sequential shared-memory LP code (early exits): measured faster here than a register-resident
                        // speculative version of the sub-problem (LP3 stage 3.5 vs 4.5 us at 4096 envs)
                        const Lines Pq = { &s_p[0][tid], T };
                        ok = lp3_subproblem(Lq, i, s_q[4 * M + 2][item], Pq, r2);
                    }
                    s_r2[0][tid] = r2.x; s_r2[1][tid] = r2.y; s_r2[2][tid] = ok ? 1.0f : 0.0f;
                }
                __syncthreads();
                if (mine && i == 1) {                            // the item's first lane runs linearProgram3's outer scan
                    const Lines Lq = { &s_q[0][item], T };
                    const int qn = __float_as_int(s_q[4 * M + 0][item]), qf = __float_as_int(s_q[4 * M + 1][item]);
                    const float qr = s_q[4 * M + 2][item];
                    V2 res = mk(s_q[4 * M + 3][item], s_q[4 * M + 4][item]);
                    lp3_outer_scan(Lq, qn, qf, qr, res, [&](int ii, V2 &r2) {
                        const int src = tid + (ii - 1);               // lane of sub-problem ii of this item
                        r2 = mk(s_r2[0][src], s_r2[1][src]);
                        return s_r2[2][src] != 0.0f;
                    });
                    s_res[0][item] = res.x; s_res[1][item] = res.y;
                }
                __syncthreads();
            }
            if (need3) nv = mk(s_res[0][slot], s_res[1][slot]);
        }
    }

    if (A.act_only) {                      // crowdsim_orca_act: the robot's ORCA decision only, nothing is mutated
        if (live && is_robot) st2(A.io.action_out, e, make_double2((double)nv.x, (double)nv.y));
        return;
    }
    if constexpr (STAGE == 4) {            // + lp3
        if (live && !is_robot) { const size_t i = (size_t)e * N + a; st2(A.st.h_vel, i, make_double2(vel.x + (double)nv.x * 0, vel.y + (double)nv.y * 0)); }
        return;
    }
    // ---- robot velocity of this step, broadcast inside the env ----
    double ax = 0, ay = 0, rvx = 0, rvy = 0;
    if (is_robot) {
        if (k.robot_policy == CROWDSIM_ROBOT_ORCA) { ax = (double)nv.x; ay = (double)nv.y; rvx = ax; rvy = ay; }
        else if (CS_FLAT_IS_ROT(k)) { ax = ext.x; ay = ext.y; rvx = ax * cos(ay + theta); rvy = ax * sin(ay + theta); }
        else { ax = ext.x; ay = ext.y; rvx = ax; rvy = ay; }
    }
    const int rl = ebase + N;                                // my env's robot lane
    const double Rvx = __shfl_sync(CS_FULL, rvx, rl), Rvy = __shfl_sync(CS_FULL, rvy, rl);
    const double Rpx = __shfl_sync(CS_FULL, pos.x, rl), Rpy = __shfl_sync(CS_FULL, pos.y, rl);
    const double Rrad = __shfl_sync(CS_FULL, attr.x, rl);

    // ---- human lanes: swept-segment clearance (crowd_sim.py:333-345) + Euler step (agent.py:122-135) ----
    const double dt = k.time_step;
    double closest = 0.0;
    if (live && !is_robot) {
        const double px = pos.x - Rpx, py = pos.y - Rpy;
        const double vx = vel.x - Rvx, vy = vel.y - Rvy;
        const double ex = px + vx * dt, ey = py + vy * dt;
        closest = point_to_segment_dist0(px, py, ex, ey) - attr.x - Rrad;
    }
    // ordered fold over the env's humans (first collision breaks, crowd_sim.py:346-351); consumed by the robot lane
    double dmin = __longlong_as_double(0x7ff0000000000000LL); bool collision = false;
    #pragma unroll
    for (int i = 0; i < N; ++i) {
        const double ci = __shfl_sync(CS_FULL, closest, ebase + i);
        if (!collision) { if (ci < 0) collision = true; else if (ci < dmin) dmin = ci; }
    }

    // ---- robot lane: ladder, update, bookkeeping; decides about auto-reset ----
    int install = 0;
    if (is_robot && env_ok) {
        bool done = false;
        if (live) {
            double npx, npy, nvx, nvy;
            if (!CS_FLAT_IS_ROT(k)) { npx = pos.x + ax * dt; npy = pos.y + ay * dt; nvx = ax; nvy = ay; }
            else { const double th = theta + ay; npx = pos.x + cos(th) * ax * dt; npy = pos.y + sin(th) * ax * dt; nvx = nvy = 0; }
            const bool reaching_goal = norm2(npx - goal.x, npy - goal.y) < attr.x;
            double reward; int info;
            if (gtime >= k.time_limit - 1) { reward = 0; done = true; info = CROWDSIM_INFO_TIMEOUT; }
            else if (collision) { reward = k.collision_penalty; done = true; info = CROWDSIM_INFO_COLLISION; }
            else if (reaching_goal) { reward = k.success_reward; done = true; info = CROWDSIM_INFO_REACHGOAL; }
            else if (dmin < k.discomfort_dist) { reward = (dmin - k.discomfort_dist) * k.discomfort_penalty_factor * dt; done = false; info = CROWDSIM_INFO_DANGER; }
            else { reward = 0; done = false; info = CROWDSIM_INFO_NOTHING; }
            if (CS_FLAT_IS_ROT(k)) {
                double nth = fmod(theta + ay, 2 * CS_PI); if (nth < 0) nth += 2 * CS_PI;
                A.st.r_theta[e] = nth; nvx = ax * cos(nth); nvy = ax * sin(nth);
            }
            st2(A.st.r_pos, e, make_double2(npx, npy));
            st2(A.st.r_vel, e, make_double2(nvx, nvy));
            const double ntime = gtime + dt;
            A.st.g_time[e] = ntime;
            if (A.io.action_out) st2(A.io.action_out, e, make_double2(nvx, nvy));
            A.io.reward[e] = reward; A.io.dmin[e] = dmin; A.io.done[e] = done ? 1 : 0; A.io.info[e] = (uint8_t)info;
            if (A.has_ep) {
                const crowdsim_episodes &ep = A.ep;
                const int t = ep_t;
                const double disc = (t < ep.discount_len) ? ep.discount[t] : 0.0;
                const double ret = ep_ret + disc * reward;
                int tc = ep_tc; double mds = ep_mds;
                if (info == CROWDSIM_INFO_DANGER) { tc += 1; mds += dmin; ep.ep_too_close[e] = tc; ep.ep_min_dist_sum[e] = mds; }
                ep.ep_return[e] = ret; ep.ep_steps[e] = t + 1;
                if (done) {
                    const int cs_ = ep_c;
                    if (cs_ >= 0) {
                        ep.res_info[cs_] = (uint8_t)info; ep.res_steps[cs_] = t + 1;
                        ep.res_time[cs_] = (info == CROWDSIM_INFO_TIMEOUT) ? k.time_limit : ntime;
                        ep.res_return[cs_] = ret; ep.res_too_close[cs_] = tc; ep.res_min_dist_sum[cs_] = mds;
                        if (ep.res_final_rpos) st2(ep.res_final_rpos, cs_, make_double2(npx, npy));
                    }
                    if (A.st.active && !A.has_ar) A.st.active[e] = 0;
                }
            }
        }
        if (A.has_ar) install = ar_decide(A, e, slot_state, live && done, !live && want_flag != 0);
    }
    if (A.has_ar) {                                          // warp-uniform
        install = __shfl_sync(CS_FULL, install, rl) && env_ok;
        if (install) { if (is_robot) ar_install_robot(A, e); else ar_install_human(A, e, N, a); }
        __syncwarp();
        if (install && is_robot) *reinterpret_cast<volatile uint8_t *>(A.ar.n_state + e) = CROWDSIM_SLOT_EMPTY;
    }
    if (live && !is_robot && !install) {
        const double hx = (double)nv.x, hy = (double)nv.y;
        const size_t i = (size_t)e * N + a;
        st2(A.st.h_pos, i, make_double2(pos.x + hx * dt, pos.y + hy * dt));
        st2(A.st.h_vel, i, make_double2(hx, hy));
    }
}

}  // namespace cs
