// opt_common.cuh — the per-element Adam update (optimizers.py:204-245) shared by the multi-tensor optimizer kernels
// (elementwise.cu) and the single-launch MLP train step (mlp_step.cu).
#pragma once
#include "common.cuh"

namespace b200nn {

// b^t for an integer t >= 0 by repeated squaring (t counts optimizer updates): ~2 log2(t) fp64 multiplies instead of pow()
__device__ __forceinline__ double powi(double b, long long t) {
    double r = 1.0;
    while (t > 0) {
        if (t & 1) r *= b;
        b *= b;
        t >>= 1;
    }
    return r;
}

struct OptScalars {
    float clip, lr, b1, b2, omb1, omb2, inv_bc1, inv_bc2, eps;
};

__device__ __forceinline__ void adam_elem(float& p, float g, float& m, float& v, const OptScalars& s) {
    g *= s.clip;
    m = s.b1 * m + s.omb1 * g;
    v = s.b2 * v + s.omb2 * g * g;
    const float mh = m * s.inv_bc1;
    const float vh = v * s.inv_bc2;
    const float denom = fmaxf(sqrtf(vh) + s.eps, 1e-16f);
    float upd = s.lr * mh / denom;
    if (!isfinite(upd)) upd = 0.0f;  // np.nan_to_num(update, nan=0, posinf=0, neginf=0), optimizers.py:222
    p -= upd;
}

}  // namespace b200nn
