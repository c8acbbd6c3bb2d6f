/*
 * oracle/rvo2_sim.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A whole-simulator restatement (RVOSimulator + Agent + KdTree, no obstacles) behind a
 * C ABI that oracle/shims/rvo2/__init__.py wraps as `rvo2.PyRVOSimulator`, so that the
 * reference's own Python (/root/reference/crowd_sim/envs/policy/orca.py:95-129,
 * /root/reference/crowd_sim/envs/crowd_sim.py:221-245) runs UNMODIFIED on top of it.
 * Follows SURVEY.md Appendix A.1 (doStep), A.2 (kd-tree neighbour query incl. leaf size 10,
 * midpoint split on the longer bbox side, nearer child first, persistent agent permutation),
 * A.3/A.4 via oracle/rvo2_f32.h, A.5 (double -> float casts at the boundary).
 * "parity unpinned" vs a real rvo2 binary -- see the header of oracle/rvo2_f32.h.
 */
#include <stdlib.h>
#include <string.h>
#include "rvo2_f32.h"

#define KD_MAX_LEAF 10

typedef struct {
    orc_v2 pos, vel, pref, newvel;
    float radius, max_speed, neighbor_dist, time_horizon, time_horizon_obst;
    int max_nb;
} sim_agent;

typedef struct { int begin, end, left, right; float minx, maxx, miny, maxy; } kd_node;

typedef struct {
    float time_step, global_time;
    /* defaults */
    float d_neighbor_dist, d_time_horizon, d_time_horizon_obst, d_radius, d_max_speed; int d_max_nb;
    orc_v2 d_vel;
    int n, cap;
    sim_agent *agents;
    /* kd-tree: persistent permutation of agent ids + nodes */
    int kd_n; int *kd_agents; kd_node *kd_nodes;
    orc_stats stats;
} rvo_sim;

rvo_sim *rvo_sim_create(double time_step, double neighbor_dist, long max_nb, double time_horizon,
                        double time_horizon_obst, double radius, double max_speed, double vx, double vy)
{
    rvo_sim *s = (rvo_sim *)calloc(1, sizeof(rvo_sim));
    s->time_step = (float)time_step; s->d_neighbor_dist = (float)neighbor_dist; s->d_max_nb = (int)max_nb;
    s->d_time_horizon = (float)time_horizon; s->d_time_horizon_obst = (float)time_horizon_obst;
    s->d_radius = (float)radius; s->d_max_speed = (float)max_speed; s->d_vel = orc_mk((float)vx, (float)vy);
    return s;
}

void rvo_sim_destroy(rvo_sim *s) { if (!s) return; free(s->agents); free(s->kd_agents); free(s->kd_nodes); free(s); }

long rvo_sim_add_agent(rvo_sim *s, double px, double py, double neighbor_dist, long max_nb, double time_horizon,
                       double time_horizon_obst, double radius, double max_speed, double vx, double vy)
{
    if (s->n == s->cap) { s->cap = s->cap ? 2 * s->cap : 8; s->agents = (sim_agent *)realloc(s->agents, s->cap * sizeof(sim_agent)); }
    sim_agent *a = &s->agents[s->n];
    memset(a, 0, sizeof(*a));
    a->pos = orc_mk((float)px, (float)py); a->vel = orc_mk((float)vx, (float)vy);
    a->neighbor_dist = (float)neighbor_dist; a->max_nb = (int)max_nb; a->time_horizon = (float)time_horizon;
    a->time_horizon_obst = (float)time_horizon_obst; a->radius = (float)radius; a->max_speed = (float)max_speed;
    return s->n++;
}

long rvo_sim_add_agent_default(rvo_sim *s, double px, double py)
{
    return rvo_sim_add_agent(s, px, py, s->d_neighbor_dist, s->d_max_nb, s->d_time_horizon, s->d_time_horizon_obst,
                             s->d_radius, s->d_max_speed, s->d_vel.x, s->d_vel.y);
}

long rvo_sim_num_agents(const rvo_sim *s) { return s->n; }
double rvo_sim_global_time(const rvo_sim *s) { return s->global_time; }
int rvo_sim_set_position(rvo_sim *s, long i, double x, double y) { if (i < 0 || i >= s->n) return -1; s->agents[i].pos = orc_mk((float)x, (float)y); return 0; }
int rvo_sim_set_velocity(rvo_sim *s, long i, double x, double y) { if (i < 0 || i >= s->n) return -1; s->agents[i].vel = orc_mk((float)x, (float)y); return 0; }
int rvo_sim_set_pref_velocity(rvo_sim *s, long i, double x, double y) { if (i < 0 || i >= s->n) return -1; s->agents[i].pref = orc_mk((float)x, (float)y); return 0; }
int rvo_sim_get_position(const rvo_sim *s, long i, double *x, double *y) { if (i < 0 || i >= s->n) return -1; *x = s->agents[i].pos.x; *y = s->agents[i].pos.y; return 0; }
int rvo_sim_get_velocity(const rvo_sim *s, long i, double *x, double *y) { if (i < 0 || i >= s->n) return -1; *x = s->agents[i].vel.x; *y = s->agents[i].vel.y; return 0; }
void rvo_sim_get_stats(const rvo_sim *s, long *out4) { out4[0] = s->stats.solves; out4[1] = s->stats.lines; out4[2] = s->stats.lp1_calls; out4[3] = s->stats.lp3_calls; }

/* ---- kd-tree (A.2) ---- */
static float kd_coord(const rvo_sim *s, int slot, int vertical) { const orc_v2 p = s->agents[s->kd_agents[slot]].pos; return vertical ? p.x : p.y; }

static void kd_build_rec(rvo_sim *s, int begin, int end, int node)
{
    kd_node *nd = &s->kd_nodes[node];
    nd->begin = begin; nd->end = end;
    orc_v2 p0 = s->agents[s->kd_agents[begin]].pos;
    nd->minx = nd->maxx = p0.x; nd->miny = nd->maxy = p0.y;
    for (int i = begin + 1; i < end; ++i) {
        const orc_v2 p = s->agents[s->kd_agents[i]].pos;
        nd->maxx = (nd->maxx < p.x) ? p.x : nd->maxx; nd->minx = (p.x < nd->minx) ? p.x : nd->minx;
        nd->maxy = (nd->maxy < p.y) ? p.y : nd->maxy; nd->miny = (p.y < nd->miny) ? p.y : nd->miny;
    }
    if (end - begin > KD_MAX_LEAF) {
        const int vertical = (nd->maxx - nd->minx > nd->maxy - nd->miny);
        const float split = vertical ? 0.5f * (nd->maxx + nd->minx) : 0.5f * (nd->maxy + nd->miny);
        int left = begin, right = end;
        while (left < right) {
            while (left < right && kd_coord(s, left, vertical) < split) ++left;
            while (right > left && kd_coord(s, right - 1, vertical) >= split) --right;
            if (left < right) { int t = s->kd_agents[left]; s->kd_agents[left] = s->kd_agents[right - 1]; s->kd_agents[right - 1] = t; ++left; --right; }
        }
        if (left == begin) { ++left; ++right; }
        nd->left = node + 1;
        nd->right = node + 2 * (left - begin);
        kd_build_rec(s, begin, left, nd->left);
        kd_build_rec(s, left, end, s->kd_nodes[node].right);
    }
}

static void kd_build(rvo_sim *s)
{
    if (s->kd_n < s->n) {
        s->kd_agents = (int *)realloc(s->kd_agents, s->n * sizeof(int));
        for (int i = s->kd_n; i < s->n; ++i) s->kd_agents[i] = i;
        s->kd_n = s->n;
        s->kd_nodes = (kd_node *)realloc(s->kd_nodes, (2 * s->n - 1) * sizeof(kd_node));
    }
    if (s->n > 0) kd_build_rec(s, 0, s->n, 0);
}

typedef struct { float nd[ORC_MAX_LINES]; int ni[ORC_MAX_LINES]; int cnt; } nb_list;

static float kd_box_dist_sq(const kd_node *b, orc_v2 p)
{
    const float a = b->minx - p.x, c = p.x - b->maxx, d = b->miny - p.y, e = p.y - b->maxy;
    const float z = 0.0f;
    return orc_sqr(z < a ? a : z) + orc_sqr(z < c ? c : z) + orc_sqr(z < d ? d : z) + orc_sqr(z < e ? e : z);
}

static void kd_query_rec(const rvo_sim *s, int self, float *range_sq, int node, nb_list *nl, int max_nb)
{
    const kd_node *nd = &s->kd_nodes[node];
    if (nd->end - nd->begin <= KD_MAX_LEAF) {
        for (int i = nd->begin; i < nd->end; ++i) {
            const int other = s->kd_agents[i];
            if (other != self)
                orc_insert_neighbor(orc_abssq(orc_sub(s->agents[self].pos, s->agents[other].pos)), other,
                                    nl->nd, nl->ni, &nl->cnt, max_nb, range_sq);
        }
    } else {
        const float dl = kd_box_dist_sq(&s->kd_nodes[nd->left], s->agents[self].pos);
        const float dr = kd_box_dist_sq(&s->kd_nodes[nd->right], s->agents[self].pos);
        if (dl < dr) {
            if (dl < *range_sq) { kd_query_rec(s, self, range_sq, nd->left, nl, max_nb);
                                  if (dr < *range_sq) kd_query_rec(s, self, range_sq, nd->right, nl, max_nb); }
        } else {
            if (dr < *range_sq) { kd_query_rec(s, self, range_sq, nd->right, nl, max_nb);
                                  if (dl < *range_sq) kd_query_rec(s, self, range_sq, nd->left, nl, max_nb); }
        }
    }
}

/* A.1 doStep: all agents solve from the same pre-step state, then all update. */
void rvo_sim_do_step(rvo_sim *s)
{
    kd_build(s);
    for (int a = 0; a < s->n; ++a) {
        sim_agent *ag = &s->agents[a];
        nb_list nl; nl.cnt = 0;
        int max_nb = ag->max_nb > ORC_MAX_LINES ? ORC_MAX_LINES : ag->max_nb;
        float range_sq = orc_sqr(ag->neighbor_dist);
        if (max_nb > 0) kd_query_rec(s, a, &range_sq, 0, &nl, max_nb);
        orc_line lines[ORC_MAX_LINES];
        const float inv_th = 1.0f / ag->time_horizon;
        for (int k = 0; k < nl.cnt; ++k) {
            const sim_agent *o = &s->agents[nl.ni[k]];
            lines[k] = orc_make_line(ag->pos, ag->vel, ag->radius, o->pos, o->vel, o->radius, inv_th, s->time_step);
        }
        orc_v2 nv;
        const int fail = orc_lp2(lines, nl.cnt, ag->max_speed, ag->pref, 0, &nv);
        if (fail < nl.cnt) orc_lp3(lines, nl.cnt, fail, ag->max_speed, &nv);
        ag->newvel = nv;
        if (a == 0) { s->stats.solves++; s->stats.lines += nl.cnt; s->stats.lp3_calls += (fail < nl.cnt); }
    }
    for (int a = 0; a < s->n; ++a) {
        sim_agent *ag = &s->agents[a];
        ag->vel = ag->newvel;
        ag->pos = orc_add(ag->pos, orc_scale(s->time_step, ag->vel));   /* position_ += velocity_ * timeStep_ */
    }
    s->global_time += s->time_step;
}
