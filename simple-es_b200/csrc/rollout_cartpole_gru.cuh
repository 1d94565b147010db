// rollout_cartpole_gru.cuh -- K1, CartPole-v1 (optionally POMDP) with the GRU policy (D = 6562).
//
// Replaces the same reference code as rollout_cartpole_mlp.cuh, with GymEnvModel's GRU branch
// (networks/neural_network.py:15-16,25-27,38-40; torch nn.GRU cell).
//
// Mapping: 26 kB of weights per offspring do not fit per lane, so a WARP owns
// one offspring: its weights live in shared memory, lane j owns hidden unit j, and the warp steps up
// to EC episodes of that offspring in lockstep so that every weight read from shared memory feeds EC
// FMAs.  W_ih / W_hh are stored tiled as [k/4][row][4]: lane j reads its three gate rows with
// conflict-free LDS.128, the x / h vectors come back as broadcast LDS.128 from small per-warp
// buffers.  Lane e (e < EC) additionally owns the float64 cart-pole state of episode e and evaluates
// the two logits with the contract's sequential order, then the physics.  8 warps (offspring) are
// resident per SM -- shared-memory capacity is what bounds this variant.
#pragma once
#include <cstdio>
#include "rollout_cartpole_mlp.cuh"

namespace ses {

constexpr int GRU_D = param_count(4, 2, 1);            // 6562
constexpr int G3 = 3 * HID;                            // 96 gate rows
// flat parameter offsets (nn.Module.parameters() order, networks/neural_network.py:12-17)
constexpr int GO_W1 = 0, GO_B1 = 128, GO_WIH = 160, GO_WHH = GO_WIH + G3 * HID, GO_BIH = GO_WHH + G3 * HID,
              GO_BHH = GO_BIH + G3, GO_W2 = GO_BHH + G3, GO_B2 = GO_W2 + 2 * HID;
static_assert(GO_B2 + 2 == GRU_D, "GRU layout");
constexpr int GRU_NQ = (GRU_D + 3) / 4;                // 1641 quads (flat order)

template <int EC>
struct __align__(16) GruWarpSmem {
    float4 wih[HID / 4][G3];       // [k/4][row] -> W_ih[row][4*(k/4) .. +3]
    float4 whh[HID / 4][G3];
    float4 small[(GRU_D - 2 * G3 * HID + 3) / 4 + 1];   // W1, b1 | b_ih, b_hh, W2, b2 in flat order (gap removed)
    float xbuf[EC][HID];           // tanh(fc1) per episode
    float hbuf[EC][HID];           // GRU hidden state per episode
    float obuf[EC][HID + 1];       // tanh(h') per episode, padded: lane e walks row e
};

// index into `small` (floats) of flat parameter d outside the two big matrices
__device__ __forceinline__ int gru_small_index(int d) { return d < GO_WIH ? d : d - 2 * G3 * HID; }

template <int EC, int WARPS, bool TRACE>
__global__ void __launch_bounds__(WARPS * 32) k_rollout_cartpole_gru(const RolloutParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    GruWarpSmem<EC> &sm = reinterpret_cast<GruWarpSmem<EC> *>(smem_raw)[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    const unsigned FULL = 0xffffffffu;
    float *smallf = reinterpret_cast<float *>(sm.small);

    unsigned long long warp_steps = 0;
    for (;;) {
        // ---------------------------------------------------------------- next offspring
        int id = 0;
        if (lane == 0) id = atomicAdd(p.work_counter, 1);
        id = __shfl_sync(FULL, id, 0) + p.id_begin;
        if (id >= p.id_end) break;
        __syncwarp();
        // weights: flat quad q -> shared memory (big matrices tiled, the rest in flat order)
        {
            const float *prow = p.w_override ? p.w_override + (size_t)(id - p.id_begin) * GRU_D
                                             : p.parents + (size_t)p.layout.parent(id) * GRU_D;
            const bool pert = p.w_override ? false : p.layout.perturbed(id);
            for (int q = lane; q < GRU_NQ; q += 32) {
                // 6562 = 4*1640 + 2 and every block boundary is a multiple of 4 except the very end
                const float4 w = offspring_quad(prow, GRU_D, q, pert, p.sigma, p.seed, (uint32_t)id, p.gen);
                const int d = 4 * q;
                if (d >= GO_WIH && d < GO_WHH) {
                    const int o = d - GO_WIH; sm.wih[(o & 31) >> 2][o >> 5] = w;
                } else if (d >= GO_WHH && d < GO_BIH) {
                    const int o = d - GO_WHH; sm.whh[(o & 31) >> 2][o >> 5] = w;
                } else {
                    sm.small[gru_small_index(d) >> 2] = w;
                }
            }
        }
        __syncwarp();
        // per-lane constants of this offspring: fc1 row j, biases of gate rows j, 32+j, 64+j
        const float4 w1 = sm.small[lane];                              // W1[j][0..3]
        const float b1 = smallf[GO_B1 + lane];
        const float bir = smallf[gru_small_index(GO_BIH) + lane], biz = smallf[gru_small_index(GO_BIH) + HID + lane],
                    bin = smallf[gru_small_index(GO_BIH) + 2 * HID + lane];
        const float bhr = smallf[gru_small_index(GO_BHH) + lane], bhz = smallf[gru_small_index(GO_BHH) + HID + lane],
                    bhn = smallf[gru_small_index(GO_BHH) + 2 * HID + lane];
        const float *w2 = smallf + gru_small_index(GO_W2);
        const float b20 = smallf[gru_small_index(GO_B2)], b21 = smallf[gru_small_index(GO_B2) + 1];

        long long total_steps = 0;
        for (int e0 = 0; e0 < p.E; e0 += EC) {
            // ------------------------------------------------------------ a chunk of up to EC episodes
            const int ne = min(EC, p.E - e0);
            double x = 0.0, xd = 0.0, th = 0.0, thd = 0.0;             // lane e: state of episode e0 + e
            int nstep = 0;
            bool alive = lane < ne;
            if (alive) {
                if (p.init_states) {
                    const double *s0 = p.init_states + 4 * (e0 + lane);
                    x = s0[0]; xd = s0[1]; th = s0[2]; thd = s0[3];
                } else {
                    cartpole_init(p.seed, p.init_mode, p.gen, (uint32_t)id, (uint32_t)(e0 + lane), x, xd, th, thd);
                }
            }
            float h[EC];
#pragma unroll
            for (int e = 0; e < EC; ++e) { h[e] = 0.0f; sm.hbuf[e][lane] = 0.0f; }      // model.reset() (neural_network.py:38-40)
            unsigned alive_mask = __ballot_sync(FULL, alive);

            while (alive_mask) {
                // fc1 for every live episode: obs of episode e lives in lane e
                const float o0 = (float)x, o2 = (float)th;
                const float o1 = p.pomdp ? 0.0f : (float)xd;
                const float o3 = p.pomdp ? 0.0f : (float)thd;
#pragma unroll
                for (int e = 0; e < EC; ++e) {
                    const float q0 = __shfl_sync(FULL, o0, e), q1 = __shfl_sync(FULL, o1, e),
                                q2 = __shfl_sync(FULL, o2, e), q3 = __shfl_sync(FULL, o3, e);
                    float a = b1;
                    a = fmaf(w1.x, q0, a); a = fmaf(w1.y, q1, a); a = fmaf(w1.z, q2, a); a = fmaf(w1.w, q3, a);
                    sm.xbuf[e][lane] = tanh32_fast(a);
                }
                __syncwarp();
                // gate pre-activations: lane j accumulates rows j, 32+j, 64+j of W_ih x and W_hh h
                float gir[EC], giz[EC], gin[EC], ghr[EC], ghz[EC], ghn[EC];
#pragma unroll
                for (int e = 0; e < EC; ++e) { gir[e] = bir; giz[e] = biz; gin[e] = bin; ghr[e] = bhr; ghz[e] = bhz; ghn[e] = bhn; }
#pragma unroll 2
                for (int kq = 0; kq < HID / 4; ++kq) {
                    const float4 ar = sm.wih[kq][lane], az = sm.wih[kq][HID + lane], an = sm.wih[kq][2 * HID + lane];
                    const float4 cr = sm.whh[kq][lane], cz = sm.whh[kq][HID + lane], cn = sm.whh[kq][2 * HID + lane];
#pragma unroll
                    for (int e = 0; e < EC; ++e) {
                        const float4 xv = *reinterpret_cast<const float4 *>(&sm.xbuf[e][4 * kq]);
                        const float4 hv = *reinterpret_cast<const float4 *>(&sm.hbuf[e][4 * kq]);
                        gir[e] = fmaf(ar.x, xv.x, gir[e]); gir[e] = fmaf(ar.y, xv.y, gir[e]); gir[e] = fmaf(ar.z, xv.z, gir[e]); gir[e] = fmaf(ar.w, xv.w, gir[e]);
                        giz[e] = fmaf(az.x, xv.x, giz[e]); giz[e] = fmaf(az.y, xv.y, giz[e]); giz[e] = fmaf(az.z, xv.z, giz[e]); giz[e] = fmaf(az.w, xv.w, giz[e]);
                        gin[e] = fmaf(an.x, xv.x, gin[e]); gin[e] = fmaf(an.y, xv.y, gin[e]); gin[e] = fmaf(an.z, xv.z, gin[e]); gin[e] = fmaf(an.w, xv.w, gin[e]);
                        ghr[e] = fmaf(cr.x, hv.x, ghr[e]); ghr[e] = fmaf(cr.y, hv.y, ghr[e]); ghr[e] = fmaf(cr.z, hv.z, ghr[e]); ghr[e] = fmaf(cr.w, hv.w, ghr[e]);
                        ghz[e] = fmaf(cz.x, hv.x, ghz[e]); ghz[e] = fmaf(cz.y, hv.y, ghz[e]); ghz[e] = fmaf(cz.z, hv.z, ghz[e]); ghz[e] = fmaf(cz.w, hv.w, ghz[e]);
                        ghn[e] = fmaf(cn.x, hv.x, ghn[e]); ghn[e] = fmaf(cn.y, hv.y, ghn[e]); ghn[e] = fmaf(cn.z, hv.z, ghn[e]); ghn[e] = fmaf(cn.w, hv.w, ghn[e]);
                    }
                }
                __syncwarp();                                          // everyone has read hbuf before it is rewritten
                // GRU cell (torch gate order r, z, n) and the output non-linearity
#pragma unroll
                for (int e = 0; e < EC; ++e) {
                    const float rg = fmaf(0.5f, tanh32_fast(__fmul_rn(0.5f, __fadd_rn(gir[e], ghr[e]))), 0.5f);
                    const float zg = fmaf(0.5f, tanh32_fast(__fmul_rn(0.5f, __fadd_rn(giz[e], ghz[e]))), 0.5f);
                    const float ng = tanh32_fast(fmaf(rg, ghn[e], gin[e]));
                    const float hn = fmaf(zg, h[e], __fmul_rn(__fsub_rn(1.0f, zg), ng));
                    // a finished episode keeps its (unused) state; the values are simply never read again
                    h[e] = hn;
                    sm.hbuf[e][lane] = hn;
                    sm.obuf[e][lane] = tanh32_fast(hn);
                }
                __syncwarp();
                // lane e: logits in the contract's sequential order, action, physics
                bool done = false;
                int action = 0;
                if (alive) {
                    float z0 = b20, z1 = b21;
                    const float *orow = sm.obuf[lane];
#pragma unroll 8
                    for (int j = 0; j < HID; ++j) {
                        const float oj = orow[j];
                        z0 = fmaf(w2[j], oj, z0);
                        z1 = fmaf(w2[HID + j], oj, z1);
                    }
                    action = argmax_softmax2(z0, z1);
                    done = cartpole_step(x, xd, th, thd, action);
                    ++nstep;
                    if (nstep >= p.max_step) done = true;
                    if constexpr (TRACE) {
                        const int local = id - p.id_begin;
                        if (e0 + lane == 0 && local < p.n_trace && nstep <= 200) {
                            double *t = p.trace + ((size_t)local * 200 + (nstep - 1)) * 4;
                            t[0] = x; t[1] = xd; t[2] = th; t[3] = thd;
                            p.trace_actions[(size_t)local * 200 + (nstep - 1)] = action;
                        }
                    }
                    if (done) alive = false;
                }
                alive_mask = __ballot_sync(FULL, alive);
            }
            // chunk total: lanes 0..ne-1 hold their episode lengths
            int n = lane < ne ? nstep : 0;
#pragma unroll
            for (int o = 16; o; o >>= 1) n += __shfl_xor_sync(FULL, n, o);
            total_steps += n;
        }
        warp_steps += (unsigned long long)total_steps;
        if (lane == 0) {
            p.steps[id] = total_steps;
            publish_fitness(p, id, __ddiv_rn((double)total_steps, (double)p.E));
        }
    }
    if (p.total_steps && lane == 0 && warp_steps) atomicAdd(p.total_steps, warp_steps);
}

template <int EC>
static int launch_rollout_cartpole_gru_ec(int num_sms, int ctas_per_sm, const RolloutParams &rp, bool trace, cudaStream_t st,
                                          int64_t *launches, char *err, size_t errlen)
{
    constexpr int WARPS = 4;
    const size_t smem = WARPS * sizeof(GruWarpSmem<EC>);
    auto kern = trace ? k_rollout_cartpole_gru<EC, WARPS, true> : k_rollout_cartpole_gru<EC, WARPS, false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int per_sm = 0;
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, WARPS * 32, smem);
    if (e != cudaSuccess || per_sm < 1) {
        snprintf(err, errlen, "GRU rollout kernel cannot be launched (smem %zu B): %s", smem, cudaGetErrorString(e));
        return -1;
    }
    if (ctas_per_sm > 0 && ctas_per_sm < per_sm) per_sm = ctas_per_sm;
    const int n_local = rp.id_end - rp.id_begin;
    int grid = per_sm * num_sms;
    const int need = (n_local + WARPS - 1) / WARPS;
    if (grid > need) grid = need;
    kern<<<grid, WARPS * 32, smem, st>>>(rp);
    e = cudaGetLastError();
    if (e != cudaSuccess) { snprintf(err, errlen, "GRU rollout launch failed: %s", cudaGetErrorString(e)); return -1; }
    *launches += 1;
    return 0;
}

static int launch_rollout_cartpole_gru(int num_sms, int ctas_per_sm, const RolloutParams &rp, bool trace, cudaStream_t st,
                                       int64_t *launches, char *err, size_t errlen)
{
    // episodes of one offspring step in lockstep, EC at a time (E > 5: chunks of 5, the last one partial)
    switch (rp.E >= 5 ? 5 : rp.E) {
    case 1: return launch_rollout_cartpole_gru_ec<1>(num_sms, ctas_per_sm, rp, trace, st, launches, err, errlen);
    case 2: return launch_rollout_cartpole_gru_ec<2>(num_sms, ctas_per_sm, rp, trace, st, launches, err, errlen);
    case 3: return launch_rollout_cartpole_gru_ec<3>(num_sms, ctas_per_sm, rp, trace, st, launches, err, errlen);
    case 4: return launch_rollout_cartpole_gru_ec<4>(num_sms, ctas_per_sm, rp, trace, st, launches, err, errlen);
    default: return launch_rollout_cartpole_gru_ec<5>(num_sms, ctas_per_sm, rp, trace, st, launches, err, errlen);
    }
}

}  // namespace ses
