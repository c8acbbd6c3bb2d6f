/*
 * oracle/rvo2_f32.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C99, IEEE float32, one rounding per operation) of the
 * agent step of the RVO2 library v2.0.x as bundled by sybrenstuvel/Python-RVO2,
 * the un-vendored, un-pinned native dependency behind the reference call sites
 *   /root/reference/crowd_sim/envs/policy/orca.py:95-129
 *   /root/reference/crowd_sim/envs/crowd_sim.py:221-245
 * The RVO2 sources are NOT under /root/reference; this file follows the
 * behavioural spec in SURVEY.md Appendix A (A.2 neighbours, A.3 ORCA lines,
 * A.4 linear programs) and is written fresh, array-based.
 *
 * PARITY STATUS: "parity unpinned" against a real rvo2 binary (absent here, no
 * network). Pinned instead against (a) the independent restatement digests of
 * SURVEY.md Appendix B (213/284/3 on the 500 test cases, timeout cases
 * 118/168/224, 15 190 env-steps, collision-list sha256 prefix ab25dfb557239780)
 * and (b) analytic known-answer tests (tests/test_oracle_rvo2.py).
 *
 * Build rule: gcc -O2 -ffp-contract=off (no -ffast-math, no -march=native) so
 * that no multiply-add is contracted; x86-64 SSE2 scalar float == IEEE binary32.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may use anything under oracle/.
 */
#ifndef ORACLE_RVO2_F32_H
#define ORACLE_RVO2_F32_H

#include <math.h>
#include <stddef.h>

#define ORC_EPS 0.00001f          /* RVO_EPSILON */
#define ORC_MAX_LINES 32           /* hard cap on maxNeighbors in the oracle */

typedef struct { float x, y; } orc_v2;
typedef struct { orc_v2 point, dir; } orc_line;

static inline orc_v2 orc_mk(float x, float y) { orc_v2 r; r.x = x; r.y = y; return r; }
static inline orc_v2 orc_add(orc_v2 a, orc_v2 b) { return orc_mk(a.x + b.x, a.y + b.y); }
static inline orc_v2 orc_sub(orc_v2 a, orc_v2 b) { return orc_mk(a.x - b.x, a.y - b.y); }
static inline orc_v2 orc_neg(orc_v2 a) { return orc_mk(-a.x, -a.y); }
static inline orc_v2 orc_scale(float s, orc_v2 a) { return orc_mk(s * a.x, s * a.y); }
/* Vector2::operator/ : multiply by the reciprocal (SURVEY App. A preamble). */
static inline orc_v2 orc_div(orc_v2 a, float s) { const float inv = 1.0f / s; return orc_mk(a.x * inv, a.y * inv); }
static inline float orc_dot(orc_v2 a, orc_v2 b) { return a.x * b.x + a.y * b.y; }
static inline float orc_det(orc_v2 a, orc_v2 b) { return a.x * b.y - a.y * b.x; }
static inline float orc_abssq(orc_v2 a) { return orc_dot(a, a); }
static inline float orc_abs(orc_v2 a) { return sqrtf(orc_abssq(a)); }
static inline orc_v2 orc_normalize(orc_v2 a) { return orc_div(a, orc_abs(a)); }
static inline float orc_sqr(float a) { return a * a; }

/* A.2: insertAgentNeighbor. List (dist_sq[], idx[]) of current size *cnt, capacity max_nb. */
static inline void orc_insert_neighbor(float dist_sq, int other, float *nd, int *ni, int *cnt,
                                       int max_nb, float *range_sq)
{
    if (dist_sq < *range_sq) {
        if (*cnt < max_nb) { nd[*cnt] = dist_sq; ni[*cnt] = other; ++*cnt; }
        int i = *cnt - 1;
        while (i != 0 && dist_sq < nd[i - 1]) { nd[i] = nd[i - 1]; ni[i] = ni[i - 1]; --i; }
        nd[i] = dist_sq; ni[i] = other;
        if (*cnt == max_nb) *range_sq = nd[*cnt - 1];
    }
}

/* A.3: one ORCA half-plane of `self` induced by `other`. */
static inline orc_line orc_make_line(orc_v2 p, orc_v2 v, float r, orc_v2 po, orc_v2 vo, float ro,
                                     float inv_time_horizon, float time_step)
{
    const orc_v2 rel_pos = orc_sub(po, p);
    const orc_v2 rel_vel = orc_sub(v, vo);
    const float dist_sq = orc_abssq(rel_pos);
    const float comb_r = r + ro;
    const float comb_r_sq = orc_sqr(comb_r);
    orc_line line; orc_v2 u;

    if (dist_sq > comb_r_sq) {
        const orc_v2 w = orc_sub(rel_vel, orc_scale(inv_time_horizon, rel_pos));
        const float w_len_sq = orc_abssq(w);
        const float dot1 = orc_dot(w, rel_pos);
        if (dot1 < 0.0f && orc_sqr(dot1) > comb_r_sq * w_len_sq) {
            /* cut-off circle */
            const float w_len = sqrtf(w_len_sq);
            const orc_v2 unit_w = orc_div(w, w_len);
            line.dir = orc_mk(unit_w.y, -unit_w.x);
            u = orc_scale(comb_r * inv_time_horizon - w_len, unit_w);
        } else {
            /* legs */
            const float leg = sqrtf(dist_sq - comb_r_sq);
            if (orc_det(rel_pos, w) > 0.0f) {
                line.dir = orc_div(orc_mk(rel_pos.x * leg - rel_pos.y * comb_r,
                                          rel_pos.x * comb_r + rel_pos.y * leg), dist_sq);
            } else {
                line.dir = orc_neg(orc_div(orc_mk(rel_pos.x * leg + rel_pos.y * comb_r,
                                                  -rel_pos.x * comb_r + rel_pos.y * leg), dist_sq));
            }
            const float dot2 = orc_dot(rel_vel, line.dir);
            u = orc_sub(orc_scale(dot2, line.dir), rel_vel);
        }
    } else {
        /* already overlapping: cut-off circle of one time step */
        const float inv_dt = 1.0f / time_step;
        const orc_v2 w = orc_sub(rel_vel, orc_scale(inv_dt, rel_pos));
        const float w_len = orc_abs(w);
        const orc_v2 unit_w = orc_div(w, w_len);
        line.dir = orc_mk(unit_w.y, -unit_w.x);
        u = orc_scale(comb_r * inv_dt - w_len, unit_w);
    }
    line.point = orc_add(v, orc_scale(0.5f, u));
    return line;
}

/* A.4 lp1 */
static inline int orc_lp1(const orc_line *lines, int line_no, float radius, orc_v2 opt, int dir_opt,
                          orc_v2 *result)
{
    const orc_line L = lines[line_no];
    const float dp = orc_dot(L.point, L.dir);
    const float disc = orc_sqr(dp) + orc_sqr(radius) - orc_abssq(L.point);
    if (disc < 0.0f) return 0;
    const float sq = sqrtf(disc);
    float t_left = -dp - sq;
    float t_right = -dp + sq;
    for (int i = 0; i < line_no; ++i) {
        const float den = orc_det(L.dir, lines[i].dir);
        const float num = orc_det(lines[i].dir, orc_sub(L.point, lines[i].point));
        if (fabsf(den) <= ORC_EPS) {
            if (num < 0.0f) return 0;
            continue;
        }
        const float t = num / den;
        if (den >= 0.0f) { t_right = (t < t_right) ? t : t_right; }   /* std::min(tRight, t) */
        else             { t_left  = (t_left < t) ? t : t_left; }     /* std::max(tLeft, t)  */
        if (t_left > t_right) return 0;
    }
    if (dir_opt) {
        if (orc_dot(opt, L.dir) > 0.0f) *result = orc_add(L.point, orc_scale(t_right, L.dir));
        else                            *result = orc_add(L.point, orc_scale(t_left, L.dir));
    } else {
        const float t = orc_dot(L.dir, orc_sub(opt, L.point));
        if (t < t_left)       *result = orc_add(L.point, orc_scale(t_left, L.dir));
        else if (t > t_right) *result = orc_add(L.point, orc_scale(t_right, L.dir));
        else                  *result = orc_add(L.point, orc_scale(t, L.dir));
    }
    return 1;
}

/* A.4 lp2; returns index of failing line or n. */
static inline int orc_lp2(const orc_line *lines, int n, float radius, orc_v2 opt, int dir_opt,
                          orc_v2 *result)
{
    if (dir_opt)                                   *result = orc_mk(opt.x * radius, opt.y * radius);
    else if (orc_abssq(opt) > orc_sqr(radius)) { const orc_v2 nv = orc_normalize(opt);
                                                   *result = orc_mk(nv.x * radius, nv.y * radius); }
    else                                           *result = opt;
    for (int i = 0; i < n; ++i) {
        if (orc_det(lines[i].dir, orc_sub(lines[i].point, *result)) > 0.0f) {
            const orc_v2 tmp = *result;
            if (!orc_lp1(lines, i, radius, opt, dir_opt, result)) { *result = tmp; return i; }
        }
    }
    return n;
}

/* A.4 lp3 (no obstacle lines: numObstLines == 0 always, crowd_sim never adds obstacles). */
static inline void orc_lp3(const orc_line *lines, int n, int begin, float radius, orc_v2 *result)
{
    float distance = 0.0f;
    orc_line proj[ORC_MAX_LINES];
    for (int i = begin; i < n; ++i) {
        if (orc_det(lines[i].dir, orc_sub(lines[i].point, *result)) > distance) {
            int np = 0;
            for (int j = 0; j < i; ++j) {
                orc_line pl;
                const float d = orc_det(lines[i].dir, lines[j].dir);
                if (fabsf(d) <= ORC_EPS) {
                    if (orc_dot(lines[i].dir, lines[j].dir) > 0.0f) continue;
                    pl.point = orc_scale(0.5f, orc_add(lines[i].point, lines[j].point));
                } else {
                    const float t = orc_det(lines[j].dir, orc_sub(lines[i].point, lines[j].point)) / d;
                    pl.point = orc_add(lines[i].point, orc_scale(t, lines[i].dir));
                }
                pl.dir = orc_normalize(orc_sub(lines[j].dir, lines[i].dir));
                proj[np++] = pl;
            }
            const orc_v2 tmp = *result;
            if (orc_lp2(proj, np, radius, orc_mk(-lines[i].dir.y, lines[i].dir.x), 1, result) < np)
                *result = tmp;
            distance = orc_det(lines[i].dir, orc_sub(lines[i].point, *result));
        }
    }
}

/* Statistics the oracle can report (used by tests / DESIGN.md workload characterisation). */
typedef struct { long solves, lines, lp1_calls, lp3_calls; } orc_stats;

/*
 * One agent's computeNeighbors + computeNewVelocity, brute-force neighbour scan in
 * candidate order 0..m-1 (== sim index order when the kd-tree is a single leaf; for
 * >10 agents the kd traversal differs from this only in the order of exact distSq ties).
 * Candidates: positions/velocities/radii of the OTHER agents (self excluded by caller).
 */
static inline orc_v2 orc_solve(orc_v2 p, orc_v2 v, float r, float max_speed, orc_v2 pref,
                               const orc_v2 *op, const orc_v2 *ov, const float *orad, int m,
                               float neighbor_dist, int max_nb, float time_horizon, float time_step,
                               orc_stats *st)
{
    float nd[ORC_MAX_LINES]; int ni[ORC_MAX_LINES]; int cnt = 0;
    orc_line lines[ORC_MAX_LINES];
    if (max_nb > ORC_MAX_LINES) max_nb = ORC_MAX_LINES;
    float range_sq = orc_sqr(neighbor_dist);
    if (max_nb > 0)
        for (int j = 0; j < m; ++j)
            orc_insert_neighbor(orc_abssq(orc_sub(p, op[j])), j, nd, ni, &cnt, max_nb, &range_sq);
    const float inv_th = 1.0f / time_horizon;
    for (int k = 0; k < cnt; ++k)
        lines[k] = orc_make_line(p, v, r, op[ni[k]], ov[ni[k]], orad[ni[k]], inv_th, time_step);
    orc_v2 nv;
    const int fail = orc_lp2(lines, cnt, max_speed, pref, 0, &nv);
    if (fail < cnt) orc_lp3(lines, cnt, fail, max_speed, &nv);
    if (st) { st->solves++; st->lines += cnt; st->lp3_calls += (fail < cnt); }
    return nv;
}

#endif /* ORACLE_RVO2_F32_H */
