// hfa_dp.cu -- the stay / advance / skip log-max recurrence over T frames x S states.
//
// Reference semantics: tools/alignment_decoder.py:170-230 (forward_pass) and :245-257 (init).
// For every frame t >= 1 and state i (all f32 unless noted, ratio = f64(T)/f64(S)):
//     stay[i] = (dp[i] + e[t,i]) + not_edge[t]
//     adv[j]  = f32( f64( (dp[j] + e[t,j]) + edge[t] ) + f64(curr[j]) * ratio )     (source state j)
//     dp'[i]  = max(stay[i], adv[i-1], adv[i-2] if ids[i-1]==0)   strict '>' in that order (ties ->
//               the earlier candidate), backpointer = which one won
//     curr[i] = stayed ? max(curr[i], e[t,i]) : e[t,i];   curr[i] = 0 for id-0 (SP) states
// The recurrence is serial in t (curr depends on the path history, so there is no associative scan)
// and embarrassingly parallel across utterances, which is how it is mapped to the GPU:
//
//   hfa_dp_warp_kernel<K>  S <= 32*K <= 256: ONE WARP per utterance, K consecutive states per lane
//       held in registers (dp, curr, backpointer bits); per frame only the last two advance scores
//       of the left neighbour lane cross lanes (2 warp shuffles).  Emission rows are streamed
//       HBM -> smem by the TMA engine as contiguous 16-frame tiles (1-D bulk copy + mbarrier,
//       double buffered); backpointers leave as one bit-packed u32 per state per 16 frames.
//   hfa_dp_cta_kernel      S <= 8192: ONE CTA per utterance, 8 states per thread, the two boundary
//       scores cross warps through shared memory with one __syncthreads per frame.
//
// HBM traffic per DP cell: 4 B emission read + 2 bits backpointer write (+ 8 B per frame of edge
// logs) = the 4.25 B/cell "algorithmic bytes" of DESIGN.md.
#include "hfa_common.cuh"

namespace {

// one frame of the recurrence for the K states owned by this lane/thread.
// up1 / up2: advance scores of states (first-1) and (first-2), already -inf where they do not exist.
template <int K>
__device__ __forceinline__ void hfa_select(const float (&e)[K], const float (&stay)[K],
                                           const float (&adv)[K], float up1, float up2,
                                           uint32_t sp_mask, uint32_t skip_mask, uint32_t m1,
                                           uint32_t m2, float (&dp)[K], float (&cu)[K],
                                           uint32_t (&bits)[K])
{
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const float p2 = (k == 0) ? up1 : adv[k - 1];
        const float p3 = (k == 0) ? up2 : ((k == 1) ? up1 : adv[k - 2]);
        const bool g1 = p2 > stay[k];
        const float m = g1 ? p2 : stay[k];
        const bool g2 = ((skip_mask >> k) & 1u) && (p3 > m);
        dp[k] = g2 ? p3 : m;
        if (g1) bits[k] |= m1;
        if (g2) bits[k] |= m2;
        const float held = fmaxf(cu[k], e[k]);
        float c = (g1 || g2) ? e[k] : held;
        if ((sp_mask >> k) & 1u) c = 0.0f;
        cu[k] = c;
    }
}

template <int K> __device__ __forceinline__ void hfa_load_row(const float *row, float (&e)[K])
{
    // row is 8-byte (K even) / 16-byte (K % 4 == 0) aligned: Sp % 4 == 0 and first = lane * K
    if constexpr (K % 4 == 0) {
#pragma unroll
        for (int q = 0; q < K / 4; ++q) {
            const float4 v = reinterpret_cast<const float4 *>(row)[q];
            e[4 * q] = v.x; e[4 * q + 1] = v.y; e[4 * q + 2] = v.z; e[4 * q + 3] = v.w;
        }
    } else if constexpr (K % 2 == 0) {
#pragma unroll
        for (int q = 0; q < K / 2; ++q) {
            const float2 v = reinterpret_cast<const float2 *>(row)[q];
            e[2 * q] = v.x; e[2 * q + 1] = v.y;
        }
    } else {
#pragma unroll
        for (int k = 0; k < K; ++k) e[k] = row[k];
    }
}

template <int K>
__device__ __forceinline__ void hfa_store_bits(uint32_t *dst, const uint32_t (&bits)[K], int first,
                                               int Sp)
{
    // dst = &bp[word_row * Sp + first]; Sp % 4 == 0, so for K % 4 == 0 a lane is all-in or all-out
    if constexpr (K % 4 == 0) {
        if (first < Sp) {
#pragma unroll
            for (int q = 0; q < K / 4; ++q)
                if (first + 4 * q < Sp)
                    reinterpret_cast<uint4 *>(dst)[q] =
                        make_uint4(bits[4 * q], bits[4 * q + 1], bits[4 * q + 2], bits[4 * q + 3]);
        }
    } else if constexpr (K % 2 == 0) {
#pragma unroll
        for (int q = 0; q < K / 2; ++q)
            if (first + 2 * q < Sp)
                reinterpret_cast<uint2 *>(dst)[q] = make_uint2(bits[2 * q], bits[2 * q + 1]);
    } else {
#pragma unroll
        for (int k = 0; k < K; ++k)
            if (first + k < Sp) dst[k] = bits[k];
    }
}

// per-thread static state flags (alignment_decoder.py:194,227): bit k set when state first+k is an
// id-0 (SP) state / when state first+k may be entered by a two-state jump (i >= 2, ids[i-1] == 0)
template <int K>
__device__ __forceinline__ void hfa_state_masks(const int32_t *ids, int first, int S,
                                                uint32_t &sp_mask, uint32_t &skip_mask)
{
    sp_mask = 0;
    skip_mask = 0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int i = first + k;
        if (i < S) {
            if (ids[i] == 0) sp_mask |= 1u << k;
            if (i >= 2 && ids[i - 1] == 0) skip_mask |= 1u << k;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// warp per utterance
// ---------------------------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(32)
hfa_dp_warp_kernel(HfaWs ws, const int32_t *__restrict__ order, float *__restrict__ dp_dump)
{
    constexpr int ROW_MAX = 32 * K;                       // floats per smem tile row (upper bound)
    constexpr int TILE_FLOATS = HFA_TILE_T * ROW_MAX;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *tile0 = reinterpret_cast<float *>(smem_raw);
    float2 *edge0 = reinterpret_cast<float2 *>(smem_raw + 2 * TILE_FLOATS * sizeof(float));
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw + 2 * TILE_FLOATS * sizeof(float) +
                                                 2 * HFA_TILE_T * sizeof(float2));

    const int lane = threadIdx.x;
    const int u = order[blockIdx.x];
    const HfaUtt m = ws.utt[u];
    const int T = m.T, S = m.S, Sp = m.Sp;
    const int first = lane * K;
    const int n_tiles = (T + HFA_TILE_T - 1) / HFA_TILE_T;
    const float *g_emis = ws.emis + m.emis_off;
    const float2 *g_edge = ws.edge2 + m.edge_off;
    uint32_t *g_bp = ws.bp + m.bp_off;
    const double ratio = __ddiv_rn((double)T, (double)S);     // T / S (:186)

    auto issue = [&](int i) {                                  // lane 0 only
        const int st = i & 1;
        const int t0 = i * HFA_TILE_T;
        const int rows = min(HFA_TILE_T, T - t0);
        const uint32_t bytes = (uint32_t)rows * (uint32_t)Sp * 4u;
        hfa_mbar_expect_tx(&bar[st], bytes + HFA_TILE_T * (uint32_t)sizeof(float2));
        hfa_bulk_load(tile0 + st * TILE_FLOATS, g_emis + (int64_t)t0 * Sp, bytes, &bar[st]);
        hfa_bulk_load(edge0 + st * HFA_TILE_T, g_edge + t0, HFA_TILE_T * (uint32_t)sizeof(float2),
                      &bar[st]);
    };

    if (lane == 0) {
        hfa_mbar_init(&bar[0], 1);
        hfa_mbar_init(&bar[1], 1);
        hfa_fence_mbar_init();
        issue(0);
        if (n_tiles > 1) issue(1);
    }
    uint32_t sp_mask, skip_mask;
    hfa_state_masks<K>(ws.ids + m.seg_off, first, S, sp_mask, skip_mask);
    const bool lead_sp = (ws.ids[m.seg_off] == 0) && (S > 1);
    __syncwarp();

    float dp[K], cu[K];
    uint32_t bits[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        dp[k] = HFA_NEG_INF;
        cu[k] = HFA_NEG_INF;
    }

    for (int i = 0; i < n_tiles; ++i) {
        const int st = i & 1;
        hfa_mbar_wait(&bar[st], (uint32_t)((i >> 1) & 1));
        const float *tl = tile0 + st * TILE_FLOATS + first;
        const float2 *et = edge0 + st * HFA_TILE_T;
        const int rows = min(HFA_TILE_T, T - i * HFA_TILE_T);
        int tt = 0;
#pragma unroll
        for (int k = 0; k < K; ++k) bits[k] = 0;
        if (i == 0) {
            // t = 0 (:250-254): state 0 is seeded, and state 1 too when the sequence starts with SP
            float e[K];
            hfa_load_row<K>(tl, e);
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const int s = first + k;
                if (s == 0 || (s == 1 && lead_sp)) {
                    dp[k] = e[k];
                    cu[k] = e[k];
                }
                if (dp_dump != nullptr && s < S) dp_dump[m.cell_off + s] = dp[k];
            }
            tt = 1;
        }
#pragma unroll 2
        for (; tt < rows; ++tt) {
            float e[K], stay[K], adv[K];
            hfa_load_row<K>(tl + tt * Sp, e);
            const float2 ed = et[tt];
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const float base = __fadd_rn(dp[k], e[k]);
                stay[k] = __fadd_rn(base, ed.y);
                adv[k] = hfa_advance(__fadd_rn(base, ed.x), cu[k], ratio);
            }
            float up1 = __shfl_up_sync(0xffffffffu, adv[K - 1], 1);
            float up2;
            if constexpr (K >= 2) up2 = __shfl_up_sync(0xffffffffu, adv[K - 2], 1);
            else up2 = __shfl_up_sync(0xffffffffu, adv[0], 2);
            if (lane == 0) up1 = HFA_NEG_INF;      // state -1 does not exist; up2 is masked by skip
            hfa_select<K>(e, stay, adv, up1, up2, sp_mask, skip_mask, 1u << tt, 0x10000u << tt, dp,
                          cu, bits);
            if (dp_dump != nullptr) {
                const int64_t o = m.cell_off + (int64_t)(i * HFA_TILE_T + tt) * S;
#pragma unroll
                for (int k = 0; k < K; ++k)
                    if (first + k < S) dp_dump[o + first + k] = dp[k];
            }
        }
        hfa_store_bits<K>(g_bp + (int64_t)i * Sp + first, bits, first, Sp);
        __syncwarp();                               // every lane is done reading stage `st`
        if (lane == 0 && i + 2 < n_tiles) issue(i + 2);
    }

    // scores of the last two states at T-1 for the end-state rule (:269-272)
#pragma unroll
    for (int k = 0; k < K; ++k) {
        if (first + k == S - 1) ws.dp_last[2 * u] = dp[k];
        if (first + k == S - 2) ws.dp_last[2 * u + 1] = dp[k];
    }
}

// ---------------------------------------------------------------------------------------------
// CTA per utterance (256 < S <= 8192)
// ---------------------------------------------------------------------------------------------
constexpr int HFA_CTA_STAGES = 3;

template <int NT>
__global__ void __launch_bounds__(NT)
hfa_dp_cta_kernel(HfaWs ws, const int32_t *__restrict__ order, int tile_t, int stage_floats,
                  float *__restrict__ dp_dump)
{
    constexpr int K = HFA_CTA_K;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // [stages x stage_floats f32][stages x 16 float2][stages mbarriers][2 x 32 float2 exchange]
    float *tile0 = reinterpret_cast<float *>(smem_raw);
    float2 *edge0 = reinterpret_cast<float2 *>(tile0 + HFA_CTA_STAGES * stage_floats);
    uint64_t *bar = reinterpret_cast<uint64_t *>(edge0 + HFA_CTA_STAGES * HFA_TILE_T);
    float2 *xch = reinterpret_cast<float2 *>(bar + HFA_CTA_STAGES);   // [2][32]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int u = order[blockIdx.x];
    const HfaUtt m = ws.utt[u];
    const int T = m.T, S = m.S, Sp = m.Sp;
    const int first = tid * K;
    const int n_tiles = (T + tile_t - 1) / tile_t;
    const float *g_emis = ws.emis + m.emis_off;
    const float2 *g_edge = ws.edge2 + m.edge_off;
    uint32_t *g_bp = ws.bp + m.bp_off;
    const double ratio = __ddiv_rn((double)T, (double)S);
    // the edge pair always travels as the whole 16-frame (128 B) block that contains the tile
    constexpr uint32_t edge_bytes = HFA_TILE_T * (uint32_t)sizeof(float2);

    auto issue = [&](int i) {                                  // thread 0 only
        const int st = i % HFA_CTA_STAGES;
        const int t0 = i * tile_t;
        const int rows = min(tile_t, T - t0);
        const uint32_t bytes = (uint32_t)rows * (uint32_t)Sp * 4u;
        hfa_mbar_expect_tx(&bar[st], bytes + edge_bytes);
        hfa_bulk_load(tile0 + st * stage_floats, g_emis + (int64_t)t0 * Sp, bytes, &bar[st]);
        hfa_bulk_load(edge0 + st * HFA_TILE_T, g_edge + (t0 & ~(HFA_TILE_T - 1)), edge_bytes,
                      &bar[st]);
    };

    if (tid == 0) {
        for (int s = 0; s < HFA_CTA_STAGES; ++s) hfa_mbar_init(&bar[s], 1);
        hfa_fence_mbar_init();
        for (int i = 0; i < HFA_CTA_STAGES && i < n_tiles; ++i) issue(i);
    }
    uint32_t sp_mask, skip_mask;
    hfa_state_masks<K>(ws.ids + m.seg_off, first, S, sp_mask, skip_mask);
    const bool lead_sp = (ws.ids[m.seg_off] == 0) && (S > 1);
    __syncthreads();

    float dp[K], cu[K];
    uint32_t bits[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        dp[k] = HFA_NEG_INF;
        cu[k] = HFA_NEG_INF;
        bits[k] = 0;
    }

    for (int i = 0; i < n_tiles; ++i) {
        const int st = i % HFA_CTA_STAGES;
        hfa_mbar_wait(&bar[st], (uint32_t)((i / HFA_CTA_STAGES) & 1));
        // threads past the last padded state still run the loop (they take part in the barriers)
        // but read row 0 of the stage instead of running off its end
        const float *tl = tile0 + st * stage_floats + (first < Sp ? first : 0);
        const float2 *et = edge0 + st * HFA_TILE_T;
        const int rows = min(tile_t, T - i * tile_t);
        for (int tt = 0; tt < rows; ++tt) {
            const int t = i * tile_t + tt;
            float e[K];
            hfa_load_row<K>(tl + tt * Sp, e);
            if (t == 0) {
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const int s = first + k;
                    if (s == 0 || (s == 1 && lead_sp)) {
                        dp[k] = e[k];
                        cu[k] = e[k];
                    }
                    if (dp_dump != nullptr && s < S) dp_dump[m.cell_off + s] = dp[k];
                }
                continue;
            }
            const float2 ed = et[t & (HFA_TILE_T - 1)];
            float stay[K], adv[K];
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const float base = __fadd_rn(dp[k], e[k]);
                stay[k] = __fadd_rn(base, ed.y);
                adv[k] = hfa_advance(__fadd_rn(base, ed.x), cu[k], ratio);
            }
            float up1 = __shfl_up_sync(0xffffffffu, adv[K - 1], 1);
            float up2 = __shfl_up_sync(0xffffffffu, adv[K - 2], 1);
            float2 *slot = xch + (t & 1) * 32;
            if (lane == 31) slot[warp] = make_float2(adv[K - 1], adv[K - 2]);
            __syncthreads();
            if (lane == 0) {
                if (warp == 0) {
                    up1 = HFA_NEG_INF;
                    up2 = HFA_NEG_INF;
                } else {
                    const float2 v = slot[warp - 1];
                    up1 = v.x;
                    up2 = v.y;
                }
            }
            const int b = t & 15;
            hfa_select<K>(e, stay, adv, up1, up2, sp_mask, skip_mask, 1u << b, 0x10000u << b, dp, cu,
                          bits);
            if (dp_dump != nullptr) {
                const int64_t o = m.cell_off + (int64_t)t * S;
#pragma unroll
                for (int k = 0; k < K; ++k)
                    if (first + k < S) dp_dump[o + first + k] = dp[k];
            }
            if (b == 15 || t == T - 1) {
                hfa_store_bits<K>(g_bp + (int64_t)(t >> 4) * Sp + first, bits, first, Sp);
#pragma unroll
                for (int k = 0; k < K; ++k) bits[k] = 0;
            }
        }
        // Every thread has read its last row of stage `st` before the barrier of that frame (or,
        // for a tile that only holds frame 0, before this point): one more barrier frees the stage.
        __syncthreads();
        if (tid == 0 && i + HFA_CTA_STAGES < n_tiles) issue(i + HFA_CTA_STAGES);
    }
    if (T == 1) hfa_store_bits<K>(g_bp + first, bits, first, Sp);   // row 0 word (all zero)

#pragma unroll
    for (int k = 0; k < K; ++k) {
        if (first + k == S - 1) ws.dp_last[2 * u] = dp[k];
        if (first + k == S - 2) ws.dp_last[2 * u + 1] = dp[k];
    }
}

template <int K>
cudaError_t launch_warp(const HfaLaunchCtx &c, const int32_t *order, int n, float *dp_dump)
{
    if (n <= 0) return cudaSuccess;
    const size_t smem = 2 * HFA_TILE_T * 32 * K * sizeof(float) + 2 * HFA_TILE_T * sizeof(float2) +
                        2 * sizeof(uint64_t);
    cudaError_t e = cudaFuncSetAttribute(hfa_dp_warp_kernel<K>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    hfa_dp_warp_kernel<K><<<n, 32, smem, c.stream>>>(c.ws, order, dp_dump);
    return cudaGetLastError();
}

}  // namespace

// order: device pointer to the utterance indices of this class; n: how many
cudaError_t hfa_launch_dp_warp(const HfaLaunchCtx &c, int K, const int32_t *order, int n,
                               float *dp_dump)
{
    switch (K) {
        case 1: return launch_warp<1>(c, order, n, dp_dump);
        case 2: return launch_warp<2>(c, order, n, dp_dump);
        case 3: return launch_warp<3>(c, order, n, dp_dump);
        case 4: return launch_warp<4>(c, order, n, dp_dump);
        case 5: return launch_warp<5>(c, order, n, dp_dump);
        case 6: return launch_warp<6>(c, order, n, dp_dump);
        case 7: return launch_warp<7>(c, order, n, dp_dump);
        case 8: return launch_warp<8>(c, order, n, dp_dump);
        default: return cudaErrorInvalidValue;
    }
}

// all utterances of the CTA class share one launch; max_sp = largest padded S among them
cudaError_t hfa_launch_dp_cta(const HfaLaunchCtx &c, const int32_t *order, int n, int max_sp,
                              float *dp_dump)
{
    if (n <= 0) return cudaSuccess;
    int threads = (max_sp + HFA_CTA_K - 1) / HFA_CTA_K;
    threads = ((threads + 31) / 32) * 32;
    if (threads > 1024) return cudaErrorInvalidValue;
    // frames per stage: largest power of two <= 16 whose stage stays under ~64 KB
    int tile_t = HFA_TILE_T;
    while (tile_t > 1 && (size_t)tile_t * max_sp * sizeof(float) > 64 * 1024) tile_t >>= 1;
    const int stage_floats = tile_t * max_sp;
    const size_t smem = (size_t)HFA_CTA_STAGES * stage_floats * sizeof(float) +
                        HFA_CTA_STAGES * HFA_TILE_T * sizeof(float2) +
                        HFA_CTA_STAGES * sizeof(uint64_t) + 2 * 32 * sizeof(float2);
    cudaError_t e;
#define HFA_CTA_LAUNCH(NT)                                                                         \
    e = cudaFuncSetAttribute(hfa_dp_cta_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                             (int)smem);                                                           \
    if (e != cudaSuccess) return e;                                                                \
    hfa_dp_cta_kernel<NT><<<n, threads, smem, c.stream>>>(c.ws, order, tile_t, stage_floats, dp_dump)
    if (threads <= 256) { HFA_CTA_LAUNCH(256); }
    else if (threads <= 512) { HFA_CTA_LAUNCH(512); }
    else { HFA_CTA_LAUNCH(1024); }
#undef HFA_CTA_LAUNCH
    return cudaGetLastError();
}
