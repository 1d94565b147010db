/*
 * ses_b200_test.h -- test hooks of the engine's CUDA library.  NOT part of the product ABI: these entry points (and the
 * alternative rollout kernels selected by SES_K1_VARIANT / SES_GRU_VARIANT / SES_K2_FUSED=0 / SES_SPREAD_SLOTS8) exist only in
 * libses_b200_tests.so, the -DSES_BUILD_TESTS build of the same sources that tests/ loads; the shipped libses_b200.so holds one
 * rollout kernel per (environment, policy, slots per warp) and exports exactly what ses_b200.h declares.
 */
#ifndef SES_B200_TEST_H
#define SES_B200_TEST_H

#include "ses_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Test hook: the numerical-contract functions on the device, elementwise.
 * kind: 0 tanh32, 1 sigmoid32, 2 ln32, 3 sin2pi32, 4 cos2pi32 (in/out f32);
 *       5 sin64, 6 cos64 (in/out f64); 7 tanh32 with the division fast path written out (what K1 runs). */
int ses_test_math(int32_t kind, const void *in_dev, void *out_dev, int64_t n, void *stream);
int ses_test_normals(ses_handle *h, uint32_t generation, int32_t id, float *out_dev /* [D] */, void *stream);
/* launch geometry of the handle's last slot-kernel rollout: out[8] = { grid, lanes, tail_start, sparse_rank,
 * sparse_quota, resident CTAs per SM, resident warps, 0 } -- lets the tests assert which scheduler path a launch took */
int ses_test_k1_geometry(ses_handle *h, int32_t *out_host /* [8] */);
/* counts mismatches between K1's 3-instruction x/1.1 and IEEE division over n pseudo-random doubles */
int ses_test_div_total_mass(uint64_t n, uint64_t *mismatches_host);
/* counts mismatches between the branch-free double division of K1 variant 6 (ddiv_fast) and IEEE division over n pseudo-random
 * operand pairs of the CartPole step's ranges (|a| in [1, 64), b in [0.5, 1)) */
int ses_test_ddiv_fast(uint64_t n, uint64_t *mismatches_host);
/* counts float32 inputs x in [lo, hi] (and -x) for which K1's fast-path tanh differs from the contract's tanh32 */
int ses_test_tanh_fast_exhaustive(float lo, float hi, uint64_t *mismatches_host);
/* the same for the packed (FFMA2) tanh of K1's hidden-unit pairs, both halves; newton = 0 drops the Newton step on
 * the reciprocal seed (K1 variant 2), newton = 1 keeps it (variant 1) */
int ses_test_tanh_x2_exhaustive(int32_t newton, float lo, float hi, uint64_t *mismatches_host);

#ifdef __cplusplus
}
#endif
#endif /* SES_B200_TEST_H */
