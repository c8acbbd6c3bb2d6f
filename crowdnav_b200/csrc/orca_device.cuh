// orca_device.cuh -- float32 ORCA velocity solver as sm_100a device code.
//
// What it computes: for ONE agent, the new velocity RVO2 would give it in doStep():
// neighbour selection (<= max_neighbors nearest within neighbor_dist, ascending distSq, ties in scan
// order), one ORCA half-plane per neighbour, then the 2-D linear programs (lp2 incremental, lp3 fallback
// that minimises the maximum penetration). It replaces, per (env, agent), the reference call chain
//   crowd_sim/envs/utils/human.py:9-17 -> crowd_sim/envs/policy/orca.py:82-132 -> rvo2 doStep (external C++),
// and follows SURVEY.md Appendix A.2-A.4.
//
// Numerics contract: every operation is an individually rounded IEEE binary32 op in RVO2's expression
// order ("a / s" on vectors multiplies by the reciprocal). This translation unit MUST be compiled with
// --fmad=false (no FFMA contraction) and without -use_fast_math (IEEE sqrt / div). Flags are enforced in
// crowdnav_b200/build.py and checked by tests (cuobjdump: no FFMA in the solver kernels).
//
// Storage: the <= 10 ORCA lines of a thread (and the <= 9 projected lines lp3 builds) are indexed
// dynamically by the LPs, so they live in shared memory, one column per thread: element (k, c) of thread t
// is at base[(k * 4 + c) * stride + t] -- consecutive threads hit consecutive banks, no conflicts.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

// The solver is plain arithmetic: it also compiles for the host, which lets a CPU test fuzz it against the C oracle
// with millions of random line sets (tests/native/lp_fuzz.cu) without spending GPU time.
#define ORCA_HD __host__ __device__

namespace orca {

constexpr float kEps = 0.00001f;   // RVO_EPSILON

struct V2 { float x, y; };
ORCA_HD __forceinline__ V2 mk(float x, float y) { V2 r; r.x = x; r.y = y; return r; }
// (Packed add/mul.rn.f32x2 for the (x, y) arithmetic was tried and is NOT usable under the one-rounding-per-operation contract:
// ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even under --fmad=false -- profiles/r02_f32x2_parity_fail.txt.)
ORCA_HD __forceinline__ V2 operator+(V2 a, V2 b) { return mk(a.x + b.x, a.y + b.y); }
ORCA_HD __forceinline__ V2 operator-(V2 a, V2 b) { return mk(a.x - b.x, a.y - b.y); }
ORCA_HD __forceinline__ V2 operator-(V2 a) { return mk(-a.x, -a.y); }
ORCA_HD __forceinline__ V2 operator*(float s, V2 a) { return mk(s * a.x, s * a.y); }
ORCA_HD __forceinline__ V2 vdiv(V2 a, float s) { const float inv = 1.0f / s; return mk(a.x * inv, a.y * inv); }
ORCA_HD __forceinline__ float dot(V2 a, V2 b) { return a.x * b.x + a.y * b.y; }
ORCA_HD __forceinline__ float det(V2 a, V2 b) { return a.x * b.y - a.y * b.x; }
ORCA_HD __forceinline__ float abssq(V2 a) { return dot(a, a); }
ORCA_HD __forceinline__ float sqr(float a) { return a * a; }
ORCA_HD __forceinline__ V2 normalize(V2 a) { return vdiv(a, sqrtf(abssq(a))); }

// Per-thread column view of a line array in shared memory.
struct Lines {
    float *base;   // already offset by the thread's column
    int stride;    // threads per block (column count)
    ORCA_HD __forceinline__ V2 point(int k) const { return mk(base[(k * 4 + 0) * stride], base[(k * 4 + 1) * stride]); }
    ORCA_HD __forceinline__ V2 dir(int k) const { return mk(base[(k * 4 + 2) * stride], base[(k * 4 + 3) * stride]); }
    ORCA_HD __forceinline__ void set(int k, V2 p, V2 d) const {
        base[(k * 4 + 0) * stride] = p.x; base[(k * 4 + 1) * stride] = p.y;
        base[(k * 4 + 2) * stride] = d.x; base[(k * 4 + 3) * stride] = d.y;
    }
};

// A.2 insertAgentNeighbor on per-thread shared-memory columns nd[k*stride], ni[k*stride].
ORCA_HD __forceinline__ void insert_neighbor(float dist_sq, int other, float *nd, int *ni, int stride,
                                                int &cnt, int max_nb, float &range_sq)
{
    if (dist_sq < range_sq) {
        if (cnt < max_nb) { nd[cnt * stride] = dist_sq; ni[cnt * stride] = other; ++cnt; }
        int i = cnt - 1;
        while (i != 0 && dist_sq < nd[(i - 1) * stride]) { nd[i * stride] = nd[(i - 1) * stride]; ni[i * stride] = ni[(i - 1) * stride]; --i; }
        nd[i * stride] = dist_sq; ni[i * stride] = other;
        if (cnt == max_nb) range_sq = nd[(cnt - 1) * stride];
    }
}

// A.3 one ORCA half-plane.
ORCA_HD __forceinline__ void make_line(V2 p, V2 v, float r, V2 po, V2 vo, float ro, float inv_th, float inv_dt,
                                          V2 &point, V2 &dir)
{
    const V2 rel_pos = po - p;
    const V2 rel_vel = v - vo;
    const float dist_sq = abssq(rel_pos);
    const float comb_r = r + ro;
    const float comb_r_sq = sqr(comb_r);
    V2 u;
    if (dist_sq > comb_r_sq) {
        const V2 w = rel_vel - inv_th * rel_pos;
        const float w_len_sq = abssq(w);
        const float dot1 = dot(w, rel_pos);
        if (dot1 < 0.0f && sqr(dot1) > comb_r_sq * w_len_sq) {
            const float w_len = sqrtf(w_len_sq);
            const V2 unit_w = vdiv(w, w_len);
            dir = mk(unit_w.y, -unit_w.x);
            u = (comb_r * inv_th - w_len) * unit_w;
        } else {
            const float leg = sqrtf(dist_sq - comb_r_sq);
            if (det(rel_pos, w) > 0.0f)
                dir = vdiv(mk(rel_pos.x * leg - rel_pos.y * comb_r, rel_pos.x * comb_r + rel_pos.y * leg), dist_sq);
            else
                dir = -vdiv(mk(rel_pos.x * leg + rel_pos.y * comb_r, -rel_pos.x * comb_r + rel_pos.y * leg), dist_sq);
            const float dot2 = dot(rel_vel, dir);
            u = dot2 * dir - rel_vel;
        }
    } else {
        const V2 w = rel_vel - inv_dt * rel_pos;
        const float w_len = sqrtf(abssq(w));
        const V2 unit_w = vdiv(w, w_len);
        dir = mk(unit_w.y, -unit_w.x);
        u = (comb_r * inv_dt - w_len) * unit_w;
    }
    point = v + 0.5f * u;
}

// A.4 lp1
ORCA_HD __forceinline__ bool lp1(const Lines &L, int line_no, float radius, V2 opt, bool dir_opt, V2 &result)
{
    const V2 lp = L.point(line_no), ld = L.dir(line_no);
    const float dp = dot(lp, ld);
    const float disc = sqr(dp) + sqr(radius) - abssq(lp);
    if (disc < 0.0f) return false;
    const float sq = sqrtf(disc);
    float t_left = -dp - sq;
    float t_right = -dp + sq;
    for (int i = 0; i < line_no; ++i) {
        const V2 ip = L.point(i), id = L.dir(i);
        const float den = det(ld, id);
        const float num = det(id, lp - ip);
        if (fabsf(den) <= kEps) {
            if (num < 0.0f) return false;
            continue;
        }
        const float t = num / den;
        if (den >= 0.0f) t_right = (t < t_right) ? t : t_right;
        else             t_left = (t_left < t) ? t : t_left;
        if (t_left > t_right) return false;
    }
    if (dir_opt) {
        if (dot(opt, ld) > 0.0f) result = lp + t_right * ld;
        else                     result = lp + t_left * ld;
    } else {
        const float t = dot(ld, opt - lp);
        if (t < t_left)       result = lp + t_left * ld;
        else if (t > t_right) result = lp + t_right * ld;
        else                  result = lp + t * ld;
    }
    return true;
}

// A.4 lp2
ORCA_HD __forceinline__ int lp2(const Lines &L, int n, float radius, V2 opt, bool dir_opt, V2 &result)
{
    if (dir_opt)                          result = mk(opt.x * radius, opt.y * radius);
    else if (abssq(opt) > sqr(radius)) { const V2 nv = normalize(opt); result = mk(nv.x * radius, nv.y * radius); }
    else                                  result = opt;
    for (int i = 0; i < n; ++i) {
        if (det(L.dir(i), L.point(i) - result) > 0.0f) {
            const V2 tmp = result;
            if (!lp1(L, i, radius, opt, dir_opt, result)) { result = tmp; return i; }
        }
    }
    return n;
}

// Projected lines of line i onto lines j < i (the loop body of linearProgram3): written to P in j order, parallel
// same-direction lines skipped. Returns their number.
ORCA_HD __forceinline__ int lp3_project(const Lines &L, int i, const Lines &P)
{
    const V2 li_p = L.point(i), li_d = L.dir(i);
    int np = 0;
    for (int j = 0; j < i; ++j) {
        const V2 lj_p = L.point(j), lj_d = L.dir(j);
        V2 pp;
        const float d = det(li_d, lj_d);
        if (fabsf(d) <= kEps) {
            if (dot(li_d, lj_d) > 0.0f) continue;
            pp = 0.5f * (li_p + lj_p);
        } else {
            const float t = det(lj_d, li_p - lj_p) / d;
            pp = li_p + t * li_d;
        }
        P.set(np++, pp, normalize(lj_d - li_d));
    }
    return np;
}

// The sub-problem of line i inside linearProgram3: linearProgram2 over the projected lines, direction optimisation,
// STARTING FROM optVelocity * radius -- it does not depend on the running result, only on the lines. Returns false if
// it fails (RVO2 then keeps the current result).
ORCA_HD __forceinline__ bool lp3_subproblem(const Lines &L, int i, float radius, const Lines &P, V2 &r2)
{
    const V2 li_d = L.dir(i);
    const int np = lp3_project(L, i, P);
    return !(lp2(P, np, radius, mk(-li_d.y, li_d.x), true, r2) < np);
}

// linearProgram3's outer loop given the (independent) sub-problem results: `sub(ii, r2)` returns whether sub-problem ii
// succeeded and its point. Line 0 has no projected lines: linearProgram2 over the empty set returns optVelocity * radius.
template <typename SubFn>
ORCA_HD __forceinline__ void lp3_outer_scan(const Lines &L, int n, int begin, float radius, V2 &result, SubFn sub)
{
    float distance = 0.0f;
    if (begin == 0 && n > 0) {
        const V2 d0 = L.dir(0), p0 = L.point(0);
        if (det(d0, p0 - result) > 0.0f) { result = mk(-d0.y * radius, d0.x * radius); distance = det(d0, p0 - result); }
    }
    for (int ii = (begin > 1 ? begin : 1); ii < n; ++ii) {
        const V2 di = L.dir(ii), pi = L.point(ii);
        if (det(di, pi - result) > distance) {
            V2 r2;
            if (sub(ii, r2)) result = r2;              // on failure the current result is kept
            distance = det(di, pi - result);
        }
    }
}

// A.4 lp3 (numObstLines == 0: crowd_sim never adds obstacles)
static ORCA_HD __noinline__ void lp3(const Lines &L, int n, int begin, float radius, const Lines &P, V2 &result)
{
    float distance = 0.0f;
    for (int i = begin; i < n; ++i) {
        const V2 li_p = L.point(i), li_d = L.dir(i);
        if (det(li_d, li_p - result) > distance) {
            V2 r2 = result;
            if (lp3_subproblem(L, i, radius, P, r2)) result = r2;      // on failure the current result is kept
            distance = det(li_d, li_p - result);
        }
    }
}

}  // namespace orca
