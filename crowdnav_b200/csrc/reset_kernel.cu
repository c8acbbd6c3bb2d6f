// reset_kernel.cu -- on-device scenario generation (sm_100a), bit-compatible with numpy's legacy MT19937.
//
// Replaces crowd_sim/envs/crowd_sim.py:251-312 (CrowdSim.reset) and the generators :155-207
// (generate_circle_crossing_human / generate_square_crossing_human) incl. agent.py:39-45
// (sample_random_attributes): np.random.seed(seed) == init_genrand(seed), np.random.random() == genrand_res53.
//
// One thread per environment that needs a reset. The 624-word MT19937 state of a thread lives in a caller-owned
// global scratch laid out [624][B] (word-major), so the threads of a warp touch consecutive addresses. The twist is
// done lazily, in place and in order (word i of the next block needs old words i, i+1 and word i+397 mod 624, which
// is old for i < 227 and already-new afterwards -- exactly the dependency order of the classic in-place loop), so a
// reset only pays for the words it actually draws (~40 for 5 circle-crossing humans) on top of the 624-step seeding
// recurrence, which is inherently sequential.
//
// float64 arithmetic in the reference's expression order; cos/sin are CUDA's (<= 1-2 ulp from glibc's, which numpy
// uses): initial coordinates can differ from the CPU reference in the last bit (tests allow 4 ulp).
#include "crowdsim_common.cuh"

namespace cs {

struct MT {
    uint32_t *mt;      // &scratch[slot], stride B between words
    size_t stride;
    int pos;           // next word to produce, 0..623 (wraps)
    __device__ __forceinline__ uint32_t &w(int i) { return mt[(size_t)i * stride]; }
    __device__ void seed(uint32_t s) {
        for (int i = 0; i < 624; ++i) { w(i) = s; s = 1812433253u * (s ^ (s >> 30)) + (uint32_t)i + 1u; }
        pos = 0;
    }
    __device__ __forceinline__ uint32_t next() {
        const int i = pos;
        const int i1 = (i == 623) ? 0 : i + 1;
        const int im = (i < 227) ? i + 397 : i - 227;
        const uint32_t y0 = (w(i) & 0x80000000u) | (w(i1) & 0x7fffffffu);
        uint32_t y = w(im) ^ (y0 >> 1) ^ ((y0 & 1u) ? 0x9908b0dfu : 0u);
        w(i) = y;
        pos = i1;
        y ^= (y >> 11); y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= (y >> 18);
        return y;
    }
    __device__ __forceinline__ double next_double() {          // genrand_res53
        const uint32_t a = next() >> 5, b = next() >> 6;
        return (a * 67108864.0 + b) / 9007199254740992.0;
    }
};

// Scene of one env: N humans by rejection sampling (crowd_sim.py:155-207), written to hp/hg/ha ([N][2] each).
// The robot is fixed at (0, -R) -> (0, R) (crowd_sim.py:274) and takes part in the separation tests.
__device__ __forceinline__ void generate_scene(MT &rng, const crowdsim_reset_args &a, int N, double *hp, double *hg, double *ha)
{
    const double rpx = 0.0, rpy = -a.circle_radius, rgx = 0.0, rgy = a.circle_radius;
    for (int i = 0; i < N; ++i) {
        double radius = a.human_radius, v_pref = a.human_v_pref;
        if (a.randomize_attributes) {                      // agent.py:44-45
            v_pref = 0.5 + (1.5 - 0.5) * rng.next_double();
            radius = 0.3 + (0.5 - 0.3) * rng.next_double();
        }
        double px, py, gx, gy;
        if (a.rule == CROWDSIM_RULE_CIRCLE) {              // crowd_sim.py:155-176
            for (;;) {
                const double angle = rng.next_double() * CS_PI * 2;
                const double px_noise = (rng.next_double() - 0.5) * v_pref;
                const double py_noise = (rng.next_double() - 0.5) * v_pref;
                px = a.circle_radius * cos(angle) + px_noise;
                py = a.circle_radius * sin(angle) + py_noise;
                bool collide = false;
                for (int k = -1; k < i && !collide; ++k) {
                    const double ar = (k < 0) ? a.robot_radius : ha[2 * k];
                    const double apx = (k < 0) ? rpx : hp[2 * k], apy = (k < 0) ? rpy : hp[2 * k + 1];
                    const double agx = (k < 0) ? rgx : hg[2 * k], agy = (k < 0) ? rgy : hg[2 * k + 1];
                    const double min_dist = radius + ar + a.discomfort_dist;
                    if (norm2(px - apx, py - apy) < min_dist || norm2(px - agx, py - agy) < min_dist) collide = true;
                }
                if (!collide) break;
            }
            gx = -px; gy = -py;
        } else {                                           // crowd_sim.py:178-207
            const double sign = (rng.next_double() > 0.5) ? -1.0 : 1.0;
            for (;;) {
                px = rng.next_double() * a.square_width * 0.5 * sign;
                py = (rng.next_double() - 0.5) * a.square_width;
                bool collide = false;
                for (int k = -1; k < i && !collide; ++k) {
                    const double ar = (k < 0) ? a.robot_radius : ha[2 * k];
                    const double apx = (k < 0) ? rpx : hp[2 * k], apy = (k < 0) ? rpy : hp[2 * k + 1];
                    if (norm2(px - apx, py - apy) < radius + ar + a.discomfort_dist) collide = true;
                }
                if (!collide) break;
            }
            for (;;) {
                gx = rng.next_double() * a.square_width * 0.5 * -sign;
                gy = (rng.next_double() - 0.5) * a.square_width;
                bool collide = false;
                for (int k = -1; k < i && !collide; ++k) {
                    const double ar = (k < 0) ? a.robot_radius : ha[2 * k];
                    const double agx = (k < 0) ? rgx : hg[2 * k], agy = (k < 0) ? rgy : hg[2 * k + 1];
                    if (norm2(gx - agx, gy - agy) < radius + ar + a.discomfort_dist) collide = true;
                }
                if (!collide) break;
            }
        }
        hp[2 * i] = px; hp[2 * i + 1] = py; hg[2 * i] = gx; hg[2 * i + 1] = gy;
        ha[2 * i] = radius; ha[2 * i + 1] = v_pref;
    }
}

// Seed of the next scene of slot e: per-slot seed (+ stride) or the shared case queue. Returns false when the queue is empty.
__device__ __forceinline__ bool next_seed(const crowdsim_reset_args &a, int e, uint32_t &seed, int &case_id)
{
    if (a.case_counter) {
        const int c = atomicAdd(a.case_counter, 1);
        if (c >= a.case_total) return false;
        seed = a.seed_base + (uint32_t)c; case_id = c;
        return true;
    }
    seed = a.seed[e];
    if (a.seed_stride) a.seed[e] = seed + a.seed_stride;
    case_id = -1;
    return true;
}

struct ResetKArgs {
    crowdsim_reset_args a;
    crowdsim_state st;
    crowdsim_episodes ep;
    crowdsim_autoreset ar;
    int has_ep, B, N;
};

__global__ void __launch_bounds__(128) reset_kernel(const __grid_constant__ ResetKArgs A)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= A.B) return;
    const crowdsim_reset_args &a = A.a;
    if (a.mask && !a.mask[e]) return;
    const int N = A.N;
    uint32_t seed; int case_id;
    if (!next_seed(a, e, seed, case_id)) {                 // case queue exhausted: the env goes idle
        if (A.st.active) A.st.active[e] = 0;
        if (A.has_ep) A.ep.ep_case[e] = -1;
        return;
    }
    MT rng; rng.mt = a.mt_scratch + e; rng.stride = (size_t)A.B;
    rng.seed(seed);
    double *hp = A.st.h_pos + (size_t)e * N * 2, *hv = A.st.h_vel + (size_t)e * N * 2;
    double *hg = A.st.h_goal + (size_t)e * N * 2, *ha = A.st.h_attr + (size_t)e * N * 2;
    st2(A.st.r_pos, e, make_double2(0.0, -a.circle_radius)); st2(A.st.r_goal, e, make_double2(0.0, a.circle_radius));   // crowd_sim.py:274
    st2(A.st.r_vel, e, make_double2(0, 0)); st2(A.st.r_attr, e, make_double2(a.robot_radius, a.robot_v_pref));
    if (A.st.r_theta) A.st.r_theta[e] = CS_PI / 2;
    A.st.g_time[e] = 0.0;
    generate_scene(rng, a, N, hp, hg, ha);
    for (int i = 0; i < N; ++i) { hv[2 * i] = 0.0; hv[2 * i + 1] = 0.0; }
    if (A.st.active) A.st.active[e] = 1;
    if (A.has_ep) {
        A.ep.ep_steps[e] = 0; A.ep.ep_return[e] = 0.0; A.ep.ep_too_close[e] = 0; A.ep.ep_min_dist_sum[e] = 0.0;
        if (a.case_counter) A.ep.ep_case[e] = case_id;
    }
}

// Generator side of the auto-reset protocol (include/crowdsim_b200.h): fill EMPTY next-scene slots, mark them READY.
__global__ void __launch_bounds__(128) prefetch_kernel(const __grid_constant__ ResetKArgs A)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= A.B) return;
    const crowdsim_autoreset &ar = A.ar;
    if (*reinterpret_cast<volatile uint8_t *>(ar.n_state + e) != CROWDSIM_SLOT_EMPTY) return;
    const crowdsim_reset_args &a = A.a;
    const int N = A.N;
    uint32_t seed; int case_id;
    if (!next_seed(a, e, seed, case_id)) { ar.n_state[e] = CROWDSIM_SLOT_EXHAUSTED; return; }
    MT rng; rng.mt = a.mt_scratch + e; rng.stride = (size_t)A.B;
    rng.seed(seed);
    generate_scene(rng, a, N, ar.n_h_pos + (size_t)e * N * 2, ar.n_h_goal + (size_t)e * N * 2, ar.n_h_attr + (size_t)e * N * 2);
    ar.n_case[e] = case_id;
    __threadfence();                                       // scene visible before the flag
    *reinterpret_cast<volatile uint8_t *>(ar.n_state + e) = CROWDSIM_SLOT_READY;
}

}  // namespace cs

static int check_reset_args(const crowdsim_reset_args *args, int B, int N)
{
    if (!args || B < 0 || N < 0 || !args->mt_scratch) return CROWDSIM_EINVAL;
    if (!args->case_counter && !args->seed) return CROWDSIM_EINVAL;
    if (N > CROWDSIM_MAX_HUMANS) return CROWDSIM_EUNSUPPORTED;
    if (args->rule != CROWDSIM_RULE_CIRCLE && args->rule != CROWDSIM_RULE_SQUARE) return CROWDSIM_EUNSUPPORTED;
    return CROWDSIM_OK;
}

extern "C" int crowdsim_reset(const crowdsim_reset_args *args, int B, int N, crowdsim_state *st, crowdsim_episodes *ep,
                              void *stream)
{
    if (!st) return CROWDSIM_EINVAL;
    if (int rc = check_reset_args(args, B, N)) return rc;
    if (N > 0 && (!st->h_pos || !st->h_vel || !st->h_goal || !st->h_attr)) return CROWDSIM_EINVAL;
    if (!st->r_pos || !st->r_vel || !st->r_goal || !st->r_attr || !st->g_time) return CROWDSIM_EINVAL;
    if (ep && (!ep->ep_steps || !ep->ep_return || !ep->ep_too_close || !ep->ep_min_dist_sum || !ep->ep_case)) return CROWDSIM_EINVAL;
    if (B == 0) return CROWDSIM_OK;
    cs::ResetKArgs A; A.a = *args; A.st = *st; A.has_ep = ep != nullptr; A.B = B; A.N = N;
    if (ep) A.ep = *ep; else memset(&A.ep, 0, sizeof(A.ep));
    memset(&A.ar, 0, sizeof(A.ar));
    const int threads = 128, blocks = (B + threads - 1) / threads;
    cs::reset_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(A);
    ++cs::g_launches;
    return (int)cudaGetLastError();
}

extern "C" int crowdsim_prefetch_scenes(const crowdsim_reset_args *args, int B, int N, const crowdsim_autoreset *ar, void *stream)
{
    if (!ar) return CROWDSIM_EINVAL;
    if (int rc = check_reset_args(args, B, N)) return rc;
    if (!ar->n_state || !ar->n_case || !ar->want || (N > 0 && (!ar->n_h_pos || !ar->n_h_goal || !ar->n_h_attr))) return CROWDSIM_EINVAL;
    if (B == 0) return CROWDSIM_OK;
    cs::ResetKArgs A; A.a = *args; A.ar = *ar; A.has_ep = 0; A.B = B; A.N = N;
    memset(&A.st, 0, sizeof(A.st)); memset(&A.ep, 0, sizeof(A.ep));
    const int threads = 128, blocks = (B + threads - 1) / threads;
    cs::prefetch_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(A);
    ++cs::g_launches;
    return (int)cudaGetLastError();
}
