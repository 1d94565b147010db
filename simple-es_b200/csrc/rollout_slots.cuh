// rollout_slots.cuh -- the persistent "warp owns S offspring slots, lane owns one episode" rollout
// kernel, generic over the environment / policy pair (CartPole-v1 MLP, simple_spread MLP).
//
// Replaces, per generation: the mp.Pool fan-out (loop.py:66-78), RolloutWorker (loop.py:108-125),
// GymEnvModel.forward (networks/neural_network.py:20-36), the env wrappers' reset/step
// (envs/gym_wrapper.py:23-45,69-77, envs/pettingzoo_wrapper.py:22-58), the third-party physics
//  and the perturbation half of _gen_offsprings
// (offspring_strategies.py:53-60,169-176,312-326).
//
// Mapping: persistent warps, no inter-warp communication.
//   * a warp owns S "offspring slots" in shared memory; a slot holds the D perturbed weights of one
//     offspring as float4 quads, slot-interleaved ([quad][slot]) so that an LDS.128 of one quad by 32
//     lanes touches at most S*16 B = one conflict-free wavefront (S = 8).
//   * a lane runs ONE episode at a time: env state in registers, fp32 policy from the slot's weights.
//     When its episode ends it is handed the next pending (slot, episode) pair by a warp-synchronous
//     scheduler (ballot + prefix), so lanes stay busy although CartPole episodes last 8..500 steps.
//   * refill is demand driven: the warp takes just enough new offspring ids from a global atomic
//     counter to occupy its idle lanes, never more.  Only `lanes_used` = E*floor(32/E) lanes take
//     work, so when every episode has the same length (a converged CartPole population, or
//     simple_spread's fixed 25 cycles) whole slots start and finish together, no episode is left
//     waiting for a later round, and the last round of a generation is packed into few full warps
//     while the others exit.
//   * the weights of a new offspring are re-derived from Philox(generation, id) by the whole warp:
//     there is no noise table, and nothing but 16 B per offspring ever goes to HBM.
//
// An Env type provides:
//   D, NQ, STATE_DIM, N_AGENTS, UNIT_REWARD (reward == 1.0 per step: fitness = steps / E exactly)
//   struct State;  init(State&, p, id, ep);
//   store_quad<S>(w[NQ][S], q, s, float4)   place flat parameter quad q of slot s (an Env may permute the table)
//   bind<S>(State&, w[NQ][S], slot)          a lane was handed an episode of `slot` (an Env may cache weights in registers)
//   step<S>(State&, w[NQ][S], slot, p, int *actions) -> done   (also adds the step reward to State::ret)
//   store_trace(State&, double *row)
#pragma once
#include "ses_common.cuh"

namespace ses {

constexpr int MAX_PEERS = 8;     // ranks of one NVSwitch box

struct RolloutParams {
    const float *parents;        // [n_parents][D]
    const float *w_override;     // optional [n_local][D]
    const double *init_states;   // optional [E][state_dim]
    double *fitness;             // [P]
    long long *steps;            // [P]
    double *trace;               // optional [n_trace][200][state_dim]
    int *trace_actions;          // optional [n_trace][200][n_agents]
    int *work_counter;           // zeroed before launch
    unsigned long long *total_steps;   // optional: += env steps simulated by this launch
    float sigma;
    uint32_t seed;
    uint32_t gen;
    Layout layout;
    Shard shard;                 // this rank's slice: local index (work queue) <-> global offspring id
    int E;
    int max_step;
    int pomdp;
    int init_mode;
    int n_trace;
    int slots_cap;               // <= S: slots a warp may hold
    int lanes_used;              // lanes of a warp that take episodes (E*floor(32/E) by default)
    int n_agents;                // simple_spread only
    int n_peers;                 // fused fitness exchange: other ranks' exchange buffers (NVLink peer memory)
    double *peer_fitness[MAX_PEERS];
    int sparse_rank;             // >= 0: warps of the CTAs that arrive on their SM as number sparse_rank or later take at most
    int sparse_quota;            //       sparse_quota offspring in total ("fractional" warps of a launch that does not fill the SMs)
    int split_ok;                // Envs with step_split(): a warp left with few episodes and nothing to refill spreads each over 2 / 4 lanes
    int strict_tail;             // > 0: once fewer than this many offspring are left in the queue, a warp takes a new offspring
                                 // only if ALL its E episodes get a lane at once (see the scheduler)
};

constexpr int MAX_E = 32;

// Fitness all-gather fused into the rollout: the lane that retires an offspring stores its fitness into the
// exchange buffer of every peer GPU (plain st.global on NVLink-mapped peer pointers), so no collective has
// to move the vector afterwards; ses_peer_barrier() publishes the stores.
__device__ __forceinline__ void publish_fitness(const RolloutParams &p, int id, double f)
{
    p.fitness[id] = f;
    if (p.n_peers > 0) {
#pragma unroll 1
        for (int r = 0; r < p.n_peers; ++r) p.peer_fitness[r][id] = f;
        __threadfence_system();
    }
}

template <class Env, int S, bool NeedRet>
struct __align__(16) SlotSmem;

template <class Env, int S>
struct __align__(16) SlotSmem<Env, S, false> {
    float4 w[Env::NQ][S];
    int off_id[S];
    int ep_next[S];
    int ep_done[S];
    int steps[S];
};

template <class Env, int S>
struct __align__(16) SlotSmem<Env, S, true> {
    float4 w[Env::NQ][S];
    double ret[S][MAX_E];        // per-episode returns, summed in episode order when the slot retires
    int off_id[S];
    int ep_next[S];
    int ep_done[S];
    int steps[S];
};

template <class Env, class = void> struct EnvSplit { static constexpr bool value = false; };
template <class Env> struct EnvSplit<Env, decltype((void)Env::SPLIT)> { static constexpr bool value = Env::SPLIT; };

// Straggler phase of the slot kernel (Envs that provide step_split<K, S>()): the warp's queue is empty and it holds at most
// 32 / K running episodes, so every further warp-step would cost a full step's latency for a handful of lanes.  The running
// episodes are compacted -- episode g moves to lanes [K g, K g + K) with its state replicated -- and stepped by K lanes each
// until they end, or until so few are left that twice as many lanes per episode fit (K = 2 -> 4).  A finished episode is
// booked into its slot (steps, ep_done) exactly as the throughput loop does; the caller retires the slots afterwards.
template <class Env, int S, int K, class Smem>
__device__ __forceinline__ void run_split(Smem &sm, int pomdp, int max_step, int lane, bool &active, int &slot, int &nstep, double &x,
                                          double &xd, double &th, double &thd)
{
    static_assert(Env::UNIT_REWARD && Env::N_AGENTS == 1, "run_split: CartPole-like envs");
    const unsigned FULL = 0xffffffffu;
    const unsigned act_mask = __ballot_sync(FULL, active);
    const int n_act = __popc(act_mask);
    const int g = lane / K;
    const int src = g < n_act ? __fns(act_mask, 0, g + 1) : lane;         // lane that holds the g-th running episode
    x = __shfl_sync(FULL, x, src); xd = __shfl_sync(FULL, xd, src);
    th = __shfl_sync(FULL, th, src); thd = __shfl_sync(FULL, thd, src);
    slot = __shfl_sync(FULL, slot, src); nstep = __shfl_sync(FULL, nstep, src);
    active = g < n_act;
    if (!active) slot = 0;                                                // idle groups step a dummy episode (result unused)
    typename Env::template SplitW<K> wr;
    Env::template split_load<K, S>(wr, sm.w, slot, lane);
    constexpr int RESPLIT = K < 4 ? 32 / (2 * K) : 0;
    for (;;) {
        const int n = __popc(__ballot_sync(FULL, active)) / K;
        if (n == 0 || n <= RESPLIT) break;
        int action;
        bool done = Env::template step_split<K>(x, xd, th, thd, wr, pomdp, lane, action);
        if (active) {
            ++nstep;
            if (nstep >= max_step) done = true;
            if (done) {
                if ((lane & (K - 1)) == 0) {
                    atomicAdd(&sm.steps[slot], nstep);
                    atomicAdd(&sm.ep_done[slot], 1);
                }
                active = false;
                slot = 0;
            }
        }
    }
    active = active && (lane & (K - 1)) == 0;                             // the leader lane keeps the episode
    __syncwarp();
}

// Kept out of line (and fed scalars only) so that the throughput loop of k_rollout_slots keeps the register allocation and
// instruction schedule it has without this phase: the loop is bound by register-operand delivery, and its speed moved by
// 4-5 % with the allocation ptxas happened to choose when this code was inlined after it (profiles/r02_k1_experiments.md).
template <class Env, int S, class Smem>
__device__ __noinline__ void split_phase(Smem *smp, int pomdp, int max_step, int slot, int nstep, double x, double xd, double th, double thd)
{
    Smem &sm = *smp;
    const int lane = threadIdx.x & 31;
    bool active = slot >= 0;
    if (__popc(__ballot_sync(0xffffffffu, active)) > 8)
        run_split<Env, S, 2>(sm, pomdp, max_step, lane, active, slot, nstep, x, xd, th, thd);      // returns when <= 8 are left
    if (__ballot_sync(0xffffffffu, active))
        run_split<Env, S, 4>(sm, pomdp, max_step, lane, active, slot, nstep, x, xd, th, thd);      // runs them to their end
}

template <class Env, int S, int WARPS, bool TRACE>
__global__ void __launch_bounds__(WARPS * 32) k_rollout_slots(const RolloutParams p)
{
    using Smem = SlotSmem<Env, S, !Env::UNIT_REWARD>;
    constexpr int NQ = Env::NQ, D = Env::D;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem &sm = reinterpret_cast<Smem *>(smem_raw)[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    const unsigned FULL = 0xffffffffu;
    const unsigned lt = lanemask_lt();
    const bool usable = lane < p.lanes_used;

    if (lane < S) { sm.off_id[lane] = -1; sm.ep_next[lane] = 0; sm.ep_done[lane] = 0; sm.steps[lane] = 0; }
    __syncwarp();
    // Sparse warps.  A launch whose episodes do not fill every resident warp (one rank's share of an 8-GPU run: 2.3 warps
    // per SM sub-partition) would leave some sub-partitions with three full warps and others with two; the former set the
    // time.  Instead every SM gets the same CTAs: those that arrive first take full loads, the CTA that arrives as number
    // sparse_rank shares what is left -- a few episodes per warp, which the straggler phase then runs on 2 or 4 lanes each,
    // at a fraction of a full warp's cost per step.  work_counter[1 + smid] counts the CTAs arriving on an SM.
    int quota = 0x7fffffff;
    if (p.sparse_rank >= 0) {
        __shared__ int cta_rank_s;
        if (threadIdx.x == 0) {
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            cta_rank_s = atomicAdd(p.work_counter + 1 + (smid & 255u), 1);
        }
        __syncthreads();
        if (cta_rank_s >= p.sparse_rank) quota = p.sparse_quota;
    }

    // per-lane episode state
    int slot = -1, nstep = 0;
    [[maybe_unused]] int ep = 0;
    typename Env::State st;
    unsigned long long warp_steps = 0;   // lanes < S: steps of the offspring they retired
    bool more = true;          // warp-uniform: the global offspring queue may still hold work
    bool sched = true;         // warp-uniform: something changed that the scheduler must look at
    [[maybe_unused]] bool to_split = false;

    for (;;) {
        if (sched) {
            // ------------------------------------------------------------------ scheduler
            __syncwarp();
            int my_id = -1;
            if (lane < S) {
                my_id = sm.off_id[lane];
                if (my_id >= 0 && sm.ep_done[lane] == p.E) {       // offspring finished: emit fitness
                    const int stp = sm.steps[lane];
                    p.steps[my_id] = (long long)stp;
                    warp_steps += (unsigned long long)stp;
                    double total;
                    if constexpr (Env::UNIT_REWARD) {
                        total = (double)stp;                        // reward 1.0 per step
                    } else {
                        total = 0.0;
                        for (int e = 0; e < p.E; ++e) total = __dadd_rn(total, sm.ret[lane][e]);
                    }
                    publish_fitness(p, my_id, __ddiv_rn(total, (double)p.E));   // loop.py:124
                    sm.off_id[lane] = -1;
                    my_id = -1;
                }
            }
            const unsigned empty_mask = __ballot_sync(FULL, lane < p.slots_cap && my_id < 0);
            const int pend_mine = (lane < S && my_id >= 0) ? (p.E - sm.ep_next[lane]) : 0;
            int pending = pend_mine;
#pragma unroll
            for (int o = 16; o; o >>= 1) pending += __shfl_xor_sync(FULL, pending, o);
            const unsigned idle_mask = __ballot_sync(FULL, usable && slot < 0);
            const int n_idle = __popc(idle_mask);
            // demand-driven refill: just enough new offspring to occupy the idle lanes.  When lanes_used is not a multiple
            // of E (32 lanes, E = 5) the last offspring taken is split over two rounds of the warp, which keeps every lane
            // busy -- except at the end of the queue, where the left-over episodes would cost each warp one more, nearly
            // empty, round: there (fewer than strict_tail offspring left) an offspring is taken only if it fits entirely.
            int want = (n_idle - pending + p.E - 1) / p.E;
            if (p.strict_tail > 0 && more && want > 0) {
                int taken = 0;                                     // one lane reads: the value must be warp-uniform
                if (lane == 0) taken = *reinterpret_cast<volatile int *>(p.work_counter);
                taken = __shfl_sync(FULL, taken, 0);
                if (p.shard.n_local - taken <= p.strict_tail) {
                    const int whole = (n_idle - pending) / p.E;
                    // (a warp with fewer lanes than E, nothing running and nothing pending must still take one)
                    if (whole > 0 || pending > 0 || __ballot_sync(FULL, slot >= 0) != 0) want = whole;
                }
            }
            want = min(min(max(want, 0), __popc(empty_mask)), quota);
            if (quota == 0) more = false;                          // a sparse warp that has taken its share
            if (want > 0 && more) {
                int base = 0;
                if (lane == 0) base = atomicAdd(p.work_counter, want);
                base = __shfl_sync(FULL, base, 0);                 // local index of the first new offspring
                if (base + want >= p.shard.n_local) more = false;
                const int got = max(0, min(want, p.shard.n_local - base));
                if (quota != 0x7fffffff) { quota -= want; if (quota == 0) more = false; }
                const int my_rank = __popc(empty_mask & lt);       // rank of this lane's slot among the empty ones
                const bool fill = ((empty_mask >> lane) & 1u) && my_rank < got;
                if (fill) {
                    sm.off_id[lane] = p.shard.local_to_id(base + my_rank);
                    sm.ep_next[lane] = 0;
                    sm.ep_done[lane] = 0;
                    sm.steps[lane] = 0;
                }
                const unsigned fill_mask = __ballot_sync(FULL, fill);
                __syncwarp();
                // regenerate the weights of the newly filled slots: (slot, quad) tasks over 32 lanes
                const int ntask = got * NQ;
                for (int t = lane; t < ntask; t += 32) {
                    const int k = t / NQ, q = t - k * NQ;
                    const int s = __fns(fill_mask, 0, k + 1);      // k-th filled slot
                    const int id = sm.off_id[s];
                    float4 wq;
                    if (p.w_override) {
                        const float *row = p.w_override + (size_t)p.shard.id_to_local(id) * D;
                        const int d = 4 * q;
                        wq.x = row[d];
                        wq.y = d + 1 < D ? row[d + 1] : 0.0f;
                        wq.z = d + 2 < D ? row[d + 2] : 0.0f;
                        wq.w = d + 3 < D ? row[d + 3] : 0.0f;
                    } else {
                        float sg;
                        const uint32_t nid = p.layout.noise_id(id, sg);
                        wq = offspring_quad(p.parents + (size_t)p.layout.parent(id) * D, D, q,
                                            p.layout.perturbed(id), __fmul_rn(p.sigma, sg), p.seed, nid, p.gen);
                    }
                    Env::template store_quad<S>(sm.w, q, s, wq);
                }
                __syncwarp();
            }
            // hand pending (slot, episode) pairs to idle lanes, in slot order (the barrier orders the retire / refill
            // writes of lanes < S to off_id / ep_next before every lane reads them: votes alone do not order memory)
            __syncwarp();
            const int r = __popc(idle_mask & lt);
            int acc = 0, my_slot = -1, my_ep = 0, my_prefix = 0, my_avail = 0;
#pragma unroll
            for (int s = 0; s < S; ++s) {
                const int nx = sm.ep_next[s];
                const int av = (sm.off_id[s] >= 0) ? (p.E - nx) : 0;
                if (usable && slot < 0 && my_slot < 0 && r < acc + av) { my_slot = s; my_ep = nx + (r - acc); }
                if (lane == s) { my_prefix = acc; my_avail = av; }
                acc += av;
            }
            __syncwarp();
            if (lane < S) sm.ep_next[lane] += max(0, min(my_avail, n_idle - my_prefix));
            __syncwarp();
            if (my_slot >= 0) {
                slot = my_slot; ep = my_ep; nstep = 0;
                Env::init(st, p, sm.off_id[my_slot], my_ep);
                Env::template bind<S>(st, sm.w, my_slot);
            }
            const unsigned act_mask = __ballot_sync(FULL, slot >= 0);
            if (act_mask == 0) break;                              // queue empty and every lane idle
            if constexpr (EnvSplit<Env>::value && !TRACE) {
                // nothing left to hand out (queue empty, every pending pair has a lane) and few episodes running: leave the
                // throughput loop for the straggler phase below
                if (p.split_ok && !more && acc <= n_idle && __popc(act_mask) <= 16) { to_split = true; break; }
            }
        }

        bool just_done = false;
        if (slot >= 0) {
            // ------------------------------------------------------------------ one env step
            int actions[Env::N_AGENTS];
            bool done = Env::template step<S>(st, sm.w, slot, p, actions);
            ++nstep;                                               // gym_wrapper.py:33 / pettingzoo_wrapper.py:34
            if (nstep >= p.max_step) done = true;                  // gym_wrapper.py:37-39, TimeLimit / max_cycles
            if constexpr (TRACE) {
                const int local = p.shard.id_to_local(sm.off_id[slot]);
                if (ep == 0 && local < p.n_trace && nstep <= 200) {
                    Env::store_trace(st, p.trace + ((size_t)local * 200 + (nstep - 1)) * Env::STATE_DIM);
#pragma unroll
                    for (int a = 0; a < Env::N_AGENTS; ++a) p.trace_actions[((size_t)local * 200 + (nstep - 1)) * Env::N_AGENTS + a] = actions[a];
                }
            }
            if (done) {
                if constexpr (!Env::UNIT_REWARD) sm.ret[slot][ep] = st.ret;
                atomicAdd(&sm.steps[slot], nstep);
                atomicAdd(&sm.ep_done[slot], 1);
                slot = -1;
                just_done = true;
            }
        }
        // the scheduler has work only right after an episode ended (a lane to re-arm, maybe a slot
        // to retire and refill); otherwise idle lanes stay idle and the warp keeps stepping
        sched = __ballot_sync(FULL, just_done) != 0;
    }
    if constexpr (EnvSplit<Env>::value && !TRACE) {
        // ------------------------------------------------------------------ straggler phase (kept out of the loop above so
        // that its code does not touch the throughput loop's register allocation and schedule)
        if (to_split) {
            split_phase<Env, S, Smem>(&sm, p.pomdp, p.max_step, slot, nstep, st.x, st.xd, st.th, st.thd);
            __syncwarp();
            if (lane < S) {                                        // retire every slot the warp still holds
                const int my_id = sm.off_id[lane];
                if (my_id >= 0) {
                    const int stp = sm.steps[lane];
                    p.steps[my_id] = (long long)stp;
                    warp_steps += (unsigned long long)stp;
                    publish_fitness(p, my_id, __ddiv_rn((double)stp, (double)p.E));
                    sm.off_id[lane] = -1;
                }
            }
        }
    }
    if (p.total_steps) {
#pragma unroll
        for (int o = 16; o; o >>= 1) warp_steps += __shfl_xor_sync(FULL, warp_steps, o);
        if (lane == 0 && warp_steps) atomicAdd(p.total_steps, warp_steps);
    }
}

}  // namespace ses
