// rollout_cartpole_gru.cuh -- K1, CartPole-v1 (POMDP) with the GRU policy (D = 6562).  [stub]
#pragma once
#include <cstdio>
#include "rollout_cartpole_mlp.cuh"
namespace ses {
static int launch_rollout_cartpole_gru(int, int, const RolloutParams &, bool, cudaStream_t, int64_t *, char *err, size_t errlen)
{
    snprintf(err, errlen, "GRU rollout kernel not built yet");
    return -1;
}
}  // namespace ses
