// orca_spec.cuh -- "speculative" formulation of RVO2's incremental 2-D linear programs for small line counts (M <= 5),
// written for SIMT execution: straight-line, select-based code with independent dependency chains.
//
// Observation that makes it possible (RVO2 linearProgram1/2/3):
//   linearProgram1(lines, i, radius, opt, dirOpt, result) never READS `result`: the feasible interval [tLeft, tRight]
//   on line i is determined by lines 0..i-1 and the speed disc, and the returned point is the point of that interval
//   closest to `opt` (or its extreme in direction `opt`). The running result only enters linearProgram2 through the
//   violation test det(dir_i, point_i - result) > 0 that decides WHETHER lp1 is called for line i.
// So for every line i we can compute, up front and independently,
//       feas_i  = "lp1(i) would succeed"            cand_i = "the point lp1(i) would return"
// with exactly the operations (and operation order) lp1 would execute, and linearProgram2 collapses into a scan
//       for i: if (violated_i(result)) { if (!feas_i) fail at i; else result = cand_i; }
// The early exits of lp1's loop do not change the outcome: tLeft only grows and tRight only shrinks, so "tLeft >
// tRight after some prefix" == "tLeft > tRight at the end"; a parallel line with negative numerator fails regardless
// of where it is met; min/max folds are done in the same j order. Results are therefore bit-identical to the
// sequential code (parity tests compare every velocity for equality), while a warp of 32 different solves executes
// one common instruction stream with M(M-1)/2 independent divisions in flight instead of the union of 32 divergent
// control paths.
// linearProgram3 has a similar structure one level up (its per-line sub-problems do not depend on the running result);
// step_flat.cuh exploits that by running the sub-problems of a queued solve on parallel lanes.
#pragma once
#include "orca_device.cuh"

namespace orca {

template <int M> struct RegLines { V2 p[M], d[M]; };

// Neighbour order of the small-crowd kernel (step_flat.cuh carries the same statements inline; this host-compilable copy
// is what tests/native/lp_fuzz.cu checks against RVO2's insertion sort -- folding the kernel onto it changes the
// generated code, so that waits for a GPU parity run). candidates c = 0..M-1 in RVO2's scan order with squared distance dsq[c],
// inr[c] = "within neighbour range" and agent index id[c] (< 8). Rank of an in-range candidate = the position RVO2's
// insertAgentNeighbor (strict <, so ties keep scan order) would give it: for cc < c, cc precedes c iff dsq[cc] <= dsq[c]
// -- one comparison per unordered pair. The agent index of the kk-th nearest is then read from a packed word (3 bits per
// position) instead of an M x M select cascade; together 7 % fewer instructions per warp than the 2 M^2
// compare-and-select form (ncu source view before / after). Returns the number of in-range candidates; src[kk] = 0 beyond.
template <int M>
ORCA_HD __forceinline__ int neighbour_order(const float (&dsq)[M], const bool (&inr)[M], const int (&id)[M], int (&src)[M])
{
    static_assert(M <= 10, "3 bits per position in a 32-bit word");
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
    for (int c = 0; c < M; ++c) if (inr[c]) { packed |= (unsigned)id[c] << (3 * rank[c]); ++nl; }
    #pragma unroll
    for (int kk = 0; kk < M; ++kk) src[kk] = (int)((packed >> (3 * kk)) & 7u);
    return nl;
}

// One ORCA half-plane with the two non-colliding variants (cut-off circle / legs) both evaluated and selected; the
// already-overlapping case (0.09 % of lines) stays a real branch. Operation order inside each variant is RVO2's
// (make_line in orca_device.cuh). Measured on B200 (scripts/latency_probe.cu, 4096 envs): this form 5.4 us for
// loads + neighbour scan + 5 lines per solve; a fully branch-free form (overlap folded in, no per-line valid branch)
// 7.0 us -- the extra arithmetic costs more than the removed divergence, the chains do not overlap in practice.
// Keeping it out of line (__noinline__, to shrink the 5 k-instruction kernel) was also measured: 9.6 vs 9.2 us per launch.
ORCA_HD __forceinline__ void make_line_sel(V2 p, V2 v, float r, V2 po, V2 vo, float ro, float inv_th, float inv_dt,
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
        // cut-off circle
        const float w_len = sqrtf(w_len_sq);
        const V2 unit_w = vdiv(w, w_len);
        const V2 dir_c = mk(unit_w.y, -unit_w.x);
        const V2 u_c = (comb_r * inv_th - w_len) * unit_w;
        // legs
        const float leg = sqrtf(dist_sq - comb_r_sq);
        const bool left = det(rel_pos, w) > 0.0f;
        const V2 num_l = mk(rel_pos.x * leg - rel_pos.y * comb_r, rel_pos.x * comb_r + rel_pos.y * leg);
        const V2 num_r = mk(rel_pos.x * leg + rel_pos.y * comb_r, -rel_pos.x * comb_r + rel_pos.y * leg);
        const V2 q = vdiv(left ? num_l : num_r, dist_sq);
        const V2 dir_l = left ? q : -q;
        const float dot2 = dot(rel_vel, dir_l);
        const V2 u_l = dot2 * dir_l - rel_vel;
        const bool cutoff = dot1 < 0.0f && sqr(dot1) > comb_r_sq * w_len_sq;
        dir = cutoff ? dir_c : dir_l;
        u = cutoff ? u_c : u_l;
    } else {
        const V2 w = rel_vel - inv_dt * rel_pos;
        const float w_len = sqrtf(abssq(w));
        const V2 unit_w = vdiv(w, w_len);
        dir = mk(unit_w.y, -unit_w.x);
        u = (comb_r * inv_dt - w_len) * unit_w;
    }
    point = v + 0.5f * u;
}

// The same half-plane split for straight-line issue: make_line_far is make_line_sel's non-overlapping branch on its own (both
// variants evaluated, selected; NO branch), make_line_overlap the already-overlapping branch. A caller evaluates make_line_far
// for all its lines unconditionally -- the constructions are independent, so their dependent chains (2 sqrt + 2 div each)
// interleave -- and repairs the rare overlapping lines (0.09 % of all) afterwards. For an overlapping pair make_line_far
// computes sqrtf of a negative number (NaN); the result is discarded. Same operations per line => same bits.
ORCA_HD __forceinline__ void make_line_far(V2 p, V2 v, float r, V2 po, V2 vo, float ro, float inv_th, V2 &point, V2 &dir, bool &overlap)
{
    const V2 rel_pos = po - p;
    const V2 rel_vel = v - vo;
    const float dist_sq = abssq(rel_pos);
    const float comb_r = r + ro;
    const float comb_r_sq = sqr(comb_r);
    overlap = !(dist_sq > comb_r_sq);
    const V2 w = rel_vel - inv_th * rel_pos;
    const float w_len_sq = abssq(w);
    const float dot1 = dot(w, rel_pos);
    const float w_len = sqrtf(w_len_sq);
    const V2 unit_w = vdiv(w, w_len);
    const V2 dir_c = mk(unit_w.y, -unit_w.x);
    const V2 u_c = (comb_r * inv_th - w_len) * unit_w;
    const float leg = sqrtf(dist_sq - comb_r_sq);
    const bool left = det(rel_pos, w) > 0.0f;
    const V2 num_l = mk(rel_pos.x * leg - rel_pos.y * comb_r, rel_pos.x * comb_r + rel_pos.y * leg);
    const V2 num_r = mk(rel_pos.x * leg + rel_pos.y * comb_r, -rel_pos.x * comb_r + rel_pos.y * leg);
    const V2 q = vdiv(left ? num_l : num_r, dist_sq);
    const V2 dir_l = left ? q : -q;
    const float dot2 = dot(rel_vel, dir_l);
    const V2 u_l = dot2 * dir_l - rel_vel;
    const bool cutoff = dot1 < 0.0f && sqr(dot1) > comb_r_sq * w_len_sq;
    dir = cutoff ? dir_c : dir_l;
    const V2 u = cutoff ? u_c : u_l;
    point = v + 0.5f * u;
}

ORCA_HD __forceinline__ void make_line_overlap(V2 p, V2 v, float r, V2 po, V2 vo, float ro, float inv_dt, V2 &point, V2 &dir)
{
    const V2 rel_pos = po - p;
    const V2 rel_vel = v - vo;
    const float comb_r = r + ro;
    const V2 w = rel_vel - inv_dt * rel_pos;
    const float w_len = sqrtf(abssq(w));
    const V2 unit_w = vdiv(w, w_len);
    dir = mk(unit_w.y, -unit_w.x);
    const V2 u = (comb_r * inv_dt - w_len) * unit_w;
    point = v + 0.5f * u;
}

// lp1 candidates of every position (speculative). valid[i]: position i holds a line (absent positions never constrain).
// CNT = number of leading positions to evaluate (compile time, <= M).
template <int M, int CNT>
ORCA_HD __forceinline__ void lp1_all(const RegLines<M> &R, const bool (&valid)[M], float radius, V2 opt, bool dir_opt,
                                        V2 (&cand)[M], bool (&feas)[M])
{
    #pragma unroll
    for (int i = 0; i < CNT; ++i) {
        const V2 lp = R.p[i], ld = R.d[i];
        const float dp = dot(lp, ld);
        const float disc = sqr(dp) + sqr(radius) - abssq(lp);
        const float sq = sqrtf(disc);
        float t_left = -dp - sq, t_right = -dp + sq;
        bool bad = disc < 0.0f;
        #pragma unroll
        for (int j = 0; j < i; ++j) {
            const float den = det(ld, R.d[j]);
            const float num = det(R.d[j], lp - R.p[j]);
            const bool use = valid[j];
            const bool par = fabsf(den) <= kEps;
            // absent positions hold zero lines (num = den = 0) and parallel lines have |den| <= eps: an IEEE division with a
            // zero numerator or denominator sends the whole warp through the division slow path (ncu: 10 slow-path calls
            // per warp = 8-9 % of all executed instructions). t is only consumed when (use && !par), so the other lanes
            // divide 1 by 1 instead.
            const bool live_pair = use && !par;
            const float t = (live_pair ? num : 1.0f) / (live_pair ? den : 1.0f);
            bad = bad || (use && par && num < 0.0f);
            const bool right = use && !par && den >= 0.0f, leftb = use && !par && den < 0.0f;
            t_right = (right && t < t_right) ? t : t_right;          // std::min(tRight, t)
            t_left = (leftb && t_left < t) ? t : t_left;             // std::max(tLeft, t)
        }
        feas[i] = !bad && !(t_left > t_right);
        if (dir_opt) {
            cand[i] = (dot(opt, ld) > 0.0f) ? (lp + t_right * ld) : (lp + t_left * ld);
        } else {
            const float t = dot(ld, opt - lp);
            const float tc = (t < t_left) ? t_left : ((t > t_right) ? t_right : t);
            cand[i] = lp + tc * ld;
        }
    }
}

// linearProgram2 as a scan over precomputed candidates. Returns the failing position or count.
template <int M, int CNT>
ORCA_HD __forceinline__ int lp2_scan(const RegLines<M> &R, const bool (&valid)[M], int count, const V2 (&cand)[M], const bool (&feas)[M],
                                        V2 init, V2 &result)
{
    result = init;
    int fail = count;
    #pragma unroll
    for (int i = 0; i < CNT; ++i) {
        const bool viol = (i < count) && (fail == count) && valid[i] && det(R.d[i], R.p[i] - result) > 0.0f;
        if (viol) { if (!feas[i]) fail = i; else result = cand[i]; }
    }
    return fail;
}

// Initial point of linearProgram2 (closest-point mode).
ORCA_HD __forceinline__ V2 lp2_init(V2 opt, float radius)
{
    if (abssq(opt) > sqr(radius)) { const V2 nv = normalize(opt); return mk(nv.x * radius, nv.y * radius); }
    return opt;
}

// Sorted-list form of RVO2's insertAgentNeighbor for the crowd kernel (step_mid.cuh): (td, tj) = the <= M nearest
// candidates seen so far, ascending; a candidate is inserted with an unrolled compare-and-shift network (static register
// indexing). Strict <: a candidate goes BEHIND equal distances (ties keep scan order) and one that is not nearer than the
// M-th entry is dropped (RVO2: rangeSq = list.back() once the list is full). Absent entries hold +inf; pass dd = +inf for
// a candidate that is out of range / not a candidate (never inserted).
template <int M>
ORCA_HD __forceinline__ void insert_sorted(float dd, int j, float (&td)[M], int (&tj)[M])
{
    #pragma unroll
    for (int kk = M - 1; kk >= 0; --kk) {                     // downwards: entry kk - 1 still holds its old value
        const int km = kk > 0 ? kk - 1 : 0;
        const bool lt = dd < td[kk];
        const bool ltp = (kk > 0) && (dd < td[km]);
        td[kk] = ltp ? td[km] : (lt ? dd : td[kk]);
        tj[kk] = ltp ? tj[km] : (lt ? j : tj[kk]);
    }
}

// ---- linearProgram3 spread over lanes (step_flat.cuh, step_mid.cuh) ----------------------------------------------------
// RVO2's linearProgram3 visits the lines i = begin .. n-1; for a line that is violated by more than the running `distance`
// it builds the lines j < i projected onto i and solves linearProgram2 over them in direction-optimisation mode, STARTING
// FROM optVelocity * radius -- so the sub-problem of line i depends only on the lines, never on the running result.
// That gives three levels of independent work per solve: the (i, j) projections (one division + one normalisation each),
// the per-i sub-problems (speculative lp1_all + lp2_scan over <= M-1 projected lines, all pair intersections in flight at
// once), and a short outer scan. The kernels put each level on its own set of lanes; these are the per-lane functions
// (host-compilable: tests/native/lp_fuzz.cu part F checks the composition against the oracle's sequential lp3).

// pair index q = i (i - 1) / 2 + j  <->  (i, j), 0 <= j < i <= 9
ORCA_HD __forceinline__ void lp3_pair_of(int q, int &i, int &j)
{
    i = 1 + (q >= 1) + (q >= 3) + (q >= 6) + (q >= 10) + (q >= 15) + (q >= 21) + (q >= 28) + (q >= 36);
    j = q - i * (i - 1) / 2;
}

// Line j projected onto line i (the loop body of linearProgram3, same operations as lp3_project). Returns false for the
// pairs RVO2 skips (parallel, same direction). Parallel lanes divide 1 by 1 (the quotient is unused there): a zero
// denominator would send the whole warp through the IEEE-division slow path.
ORCA_HD __forceinline__ bool lp3_project_pair(V2 pi, V2 di, V2 pj, V2 dj, V2 &pp, V2 &pd)
{
    const float d = det(di, dj);
    const bool par = fabsf(d) <= kEps;
    if (par && dot(di, dj) > 0.0f) { pp = mk(0.f, 0.f); pd = mk(0.f, 0.f); return false; }
    const float t = (par ? 1.0f : det(dj, pi - pj)) / (par ? 1.0f : d);
    pp = par ? 0.5f * (pi + pj) : pi + t * di;
    pd = normalize(dj - di);
    return true;
}

// Sub-problem of line i: linearProgram2 (direction optimisation, opt = perpendicular of dir_i) over its K projected-line
// positions (absent / skipped positions: valid = false, zero lines). Returns false when it fails (RVO2 keeps the running
// result then); r2 = the point it returns.
template <int K>
ORCA_HD __forceinline__ bool lp3_sub_spec(const RegLines<K> &P, const bool (&valid)[K], float radius, V2 di, V2 &r2)
{
    const V2 opt = mk(-di.y, di.x);
    V2 cand[K]; bool feas[K];
    lp1_all<K, K>(P, valid, radius, opt, true, cand, feas);
    return lp2_scan<K, K>(P, valid, K, cand, feas, mk(opt.x * radius, opt.y * radius), r2) == K;
}

}  // namespace orca
