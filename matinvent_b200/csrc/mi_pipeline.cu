// Per-crystal post-sampling work of the RL loop, on the device: the validity pre-filter and the composition-level
// rewards.  At 10 k crystals per RL iteration (BASELINE.json configs[4]) the reference's per-crystal Python loops and its
// mp.Pool (pipeline/filters/opt_filter.py:38-63, rewards/calculators/pymatgen/calc.py:57-73, rewards/reward.py:68-115)
// are what is left once the sampler is fast; here they are one warp per crystal.
#include "mi_common.cuh"

namespace {

// ---------------------------------------------------------------------------------------------------- validity pre-filter
// mask[b] bit 0: max(a, b, c) < max_len                                   (opt_filter.py:53-55, in-tree rule)
//         bit 1: min periodic interatomic distance >= min_dist and |det L| >= min_vol and max(a, b, c) <= hard_len
//                (mattergen's `structure_validity`, opt_filter.py:51 — un-vendored: RECALLED semantics, parity unpinned)
// The minimum-image distance is taken over the 27 neighbouring images of the cell as sampled (exact whenever the
// shortest lattice vector combination is among them, which holds for every cell whose angles lie in the sampler's
// clamp range and is the same search radius_graph_pbc uses, utils.py:417-430).
__global__ void validity_prefilter_kernel(const float* __restrict__ frac, const float* __restrict__ L,
                                          const float* __restrict__ lengths, const int* __restrict__ node_off, int B,
                                          float max_len, float min_dist, float min_vol, float hard_len,
                                          int* __restrict__ mask, float* __restrict__ dmin_out) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * (blockDim.x >> 5) + warp;
    if (b >= B) return;
    const float* l = L + 9 * b;
    float m[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) m[k] = __ldg(l + k);
    const int n0 = node_off[b], n1 = node_off[b + 1], n = n1 - n0;
    float best = 3.0e38f;
    // pairs (k, j), k <= j; k == j covers the 26 non-zero images of the atom itself (cells shorter than min_dist)
    const int pairs = n * (n + 1) / 2;
    for (int p = lane; p < pairs; p += 32) {
        // unrank p -> (k, j), k <= j, of the row-major lower triangle (float estimate, then fix-up)
        int j = (int)floorf((sqrtf(8.0f * (float)p + 1.0f) - 1.0f) * 0.5f);
        while ((j + 1) * (j + 2) / 2 <= p) ++j;
        while (j * (j + 1) / 2 > p) --j;
        const int k = p - j * (j + 1) / 2;
        const float dx = frac[3 * (n0 + j) + 0] - frac[3 * (n0 + k) + 0];
        const float dy = frac[3 * (n0 + j) + 1] - frac[3 * (n0 + k) + 1];
        const float dz = frac[3 * (n0 + j) + 2] - frac[3 * (n0 + k) + 2];
        // wrap the fractional difference into [-0.5, 0.5) first, then search the 27 images around it
        const float fx = dx - rintf(dx), fy = dy - rintf(dy), fz = dz - rintf(dz);
#pragma unroll
        for (int ia = -1; ia <= 1; ++ia)
#pragma unroll
            for (int ib = -1; ib <= 1; ++ib)
#pragma unroll
                for (int ic = -1; ic <= 1; ++ic) {
                    if (k == j && ia == 0 && ib == 0 && ic == 0) continue;
                    const float a = fx + (float)ia, bb = fy + (float)ib, c = fz + (float)ic;
                    const float cx = a * m[0] + bb * m[3] + c * m[6];
                    const float cy = a * m[1] + bb * m[4] + c * m[7];
                    const float cz = a * m[2] + bb * m[5] + c * m[8];
                    best = fminf(best, cx * cx + cy * cy + cz * cz);
                }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best = fminf(best, __shfl_xor_sync(0xffffffffu, best, o));
    if (lane == 0) {
        const float la = lengths[3 * b], lb = lengths[3 * b + 1], lc = lengths[3 * b + 2];
        const float lmax = fmaxf(la, fmaxf(lb, lc));
        const float det = m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) +
                          m[2] * (m[3] * m[7] - m[4] * m[6]);
        const float d = sqrtf(best);
        int bits = 0;
        if (lmax < max_len) bits |= 1;
        if (d >= min_dist && fabsf(det) >= min_vol && lmax <= hard_len && isfinite(det)) bits |= 2;
        mask[b] = bits;
        if (dmin_out) dmin_out[b] = d;
    }
}

// ---------------------------------------------------------------------------------------------------- composition rewards
// prop[p][b] = sum_el w_el(b) * table[p][el], w = mass fraction (mode 0) or atomic fraction (mode 1) of the element in
// crystal b — the form of pymatgen's HHIModel.get_hhi_reserve / CostAnalyzer.get_cost_per_kg / the crustal-abundance
// average of rewards/calculators/pymatgen/calc.py:24-45, 57-92.  A NaN table entry of a present element makes the
// property NaN ("failed", calc.py:63-70).  Then rewards/reward.py:51-115: nan_to_num for the reported property,
// linear_scaling per property (ascending / descending / target), reduce (mean / min / weight), failed -> 0.
// All arithmetic in double, like the reference's numpy float64.
struct PropCfg {
    int mode;        // 0 mass-fraction weights, 1 atomic-fraction weights
    int target;      // 0 ascending, 1 descending, 2 target value
    double minv, maxv, tval, weight;
};
constexpr int MAX_PROPS = 8;
struct RewardParams {
    int P, reduce;   // reduce: 0 mean, 1 min, 2 weight
    PropCfg cfg[MAX_PROPS];
};

__device__ __forceinline__ double linear_scaling(double v, double minv, double maxv) {
    double ss = (v - minv) / (maxv - minv);
    if (ss > 1.0) ss = 1.0;
    if (ss < 0.0) ss = 0.0;
    return ss;
}

__global__ void composition_reward_kernel(const int* __restrict__ Z, const int* __restrict__ node_off, int B,
                                          const double* __restrict__ tables /*[P][128]*/, const double* __restrict__ mass /*[128]*/,
                                          const RewardParams rp, double* __restrict__ props /*[P][B]*/,
                                          double* __restrict__ rewards, int* __restrict__ failed) {
    __shared__ int hist[8][128];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * 8 + warp;
    if (b >= B) return;
    int* h = hist[warp];
    for (int k = lane; k < 128; k += 32) h[k] = 0;
    __syncwarp();
    const int n0 = node_off[b], n1 = node_off[b + 1];
    for (int i = n0 + lane; i < n1; i += 32) {
        const int z = Z[i];
        if (z >= 0 && z < 128) atomicAdd(h + z, 1);
    }
    __syncwarp();
    if (lane != 0) return;
    // elements in increasing Z (a fixed summation order; the oracle uses the same)
    double mtot = 0.0;
    const int n = n1 - n0;
    for (int k = 0; k < 128; ++k)
        if (h[k]) mtot += (double)h[k] * mass[k];
    bool any_nan = false;
    double acc = 0.0, mn = 1.0e300;
    for (int p = 0; p < rp.P; ++p) {
        const PropCfg c = rp.cfg[p];
        double v = 0.0;
        for (int k = 0; k < 128; ++k)
            if (h[k]) {
                const double w = c.mode == 0 ? ((double)h[k] * mass[k]) / mtot : (double)h[k] / (double)n;
                v += w * tables[p * 128 + k];
            }
        const bool bad = isnan(v);
        any_nan |= bad;
        const double v0 = bad ? 0.0 : v;                          // np.nan_to_num(prop, nan=0.0)
        props[(long long)p * B + b] = v0;
        double s;
        if (c.target == 0) s = linear_scaling(v0, c.minv, c.maxv);
        else if (c.target == 1) s = linear_scaling(-v0, -c.maxv, -c.minv);
        else s = linear_scaling(-fabs(v0 - c.tval), -c.maxv, -c.minv);
        if (rp.reduce == 2) s = s * c.weight;
        acc += s;
        mn = fmin(mn, s);
    }
    double r = rp.reduce == 0 ? acc / (double)rp.P : (rp.reduce == 1 ? mn : acc);
    if (any_nan) r = 0.0;
    rewards[b] = r;
    failed[b] = any_nan ? 1 : 0;
}

}  // namespace

extern "C" int mi_validity_prefilter(const float* frac, const float* L, const float* lengths, const int* node_off, int B,
                                     float max_len, float min_dist, float min_vol, float hard_len, int* mask, float* dmin,
                                     mi_stream_t stream) {
    if (B <= 0) return MI_OK;
    MI_CHECK_ARG(frac && L && lengths && node_off && mask, "null pointer");
    validity_prefilter_kernel<<<mi_div_up(B, 4), 128, 0, (cudaStream_t)stream>>>(frac, L, lengths, node_off, B, max_len, min_dist,
                                                                                 min_vol, hard_len, mask, dmin);
    MI_CHECK_LAUNCH();
    return MI_OK;
}

extern "C" int mi_composition_reward(const int* Z, const int* node_off, int B, const double* tables, const double* mass, int P,
                                     const int* modes, const int* targets, const double* minv, const double* maxv,
                                     const double* tval, const double* weight, int reduce, double* props, double* rewards,
                                     int* failed, mi_stream_t stream) {
    if (B <= 0) return MI_OK;
    MI_CHECK_ARG(Z && node_off && tables && mass && props && rewards && failed, "null pointer");
    MI_CHECK_ARG(P >= 1 && P <= MAX_PROPS, "1..8 properties");
    MI_CHECK_ARG(reduce >= 0 && reduce <= 2, "reduce: 0 mean, 1 min, 2 weight");
    MI_CHECK_ARG(modes && targets && minv && maxv && tval && weight, "null property configuration (host arrays)");
    RewardParams rp;
    rp.P = P;
    rp.reduce = reduce;
    for (int p = 0; p < P; ++p) {
        MI_CHECK_ARG(targets[p] >= 0 && targets[p] <= 2 && (modes[p] == 0 || modes[p] == 1), "bad property mode / target");
        rp.cfg[p].mode = modes[p];
        rp.cfg[p].target = targets[p];
        rp.cfg[p].minv = minv[p];
        rp.cfg[p].maxv = maxv[p];
        rp.cfg[p].tval = tval[p];
        rp.cfg[p].weight = weight[p];
    }
    composition_reward_kernel<<<mi_div_up(B, 8), 256, 0, (cudaStream_t)stream>>>(Z, node_off, B, tables, mass, rp, props, rewards,
                                                                                 failed);
    MI_CHECK_LAUNCH();
    return MI_OK;
}

// ---------------------------------------------------------------------------------------------------- MatterGen adapter
// models/mattergen/loss.py:63-73: the per-sample loss is the weighted sum of the per-field per-sample losses
// (pos 0.1, cell 1.0, atomic_numbers 1.0 by default), accumulated in the order the fields are given.
namespace {
struct FieldSum {
    const float* f[4];
    float w[4];
    int F;
};
__global__ void weighted_field_sum_kernel(const FieldSum fs, int B, float* __restrict__ out) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    float acc = fs.w[0] * fs.f[0][b];
    for (int k = 1; k < fs.F; ++k) acc = acc + fs.w[k] * fs.f[k][b];
    out[b] = acc;
}
}  // namespace

extern "C" int mi_weighted_field_sum(int B, int F, const float* f0, const float* f1, const float* f2, const float* f3,
                                     const float* weights_host, float* out, mi_stream_t stream) {
    if (B <= 0) return MI_OK;
    MI_CHECK_ARG(F >= 1 && F <= 4 && weights_host && out && f0, "1..4 fields");
    FieldSum fs;
    const float* f[4] = {f0, f1, f2, f3};
    for (int k = 0; k < 4; ++k) {
        fs.f[k] = f[k];
        fs.w[k] = k < F ? weights_host[k] : 0.f;
        MI_CHECK_ARG(k >= F || f[k] != nullptr, "null field");
    }
    fs.F = F;
    weighted_field_sum_kernel<<<mi_div_up(B, 256), 256, 0, (cudaStream_t)stream>>>(fs, B, out);
    MI_CHECK_LAUNCH();
    return MI_OK;
}
