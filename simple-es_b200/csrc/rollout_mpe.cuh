// rollout_mpe.cuh -- K1, PettingZoo MPE simple_spread with the shared MLP policy.  [stub]
#pragma once
#include <cstdio>
#include "rollout_cartpole_mlp.cuh"
namespace ses {
static int launch_rollout_mpe(int, int, const RolloutParams &, bool, cudaStream_t, int64_t *, char *err, size_t errlen)
{
    snprintf(err, errlen, "simple_spread rollout kernel not built yet");
    return -1;
}
}  // namespace ses
