// hfa_emission.cu -- logits -> emission matrix and edge logs, one fused pass.
//
// Reference: tools/alignment_decoder.py:35-40 (mask = ids of ph_seq U {0}), :53 (others pushed down
// by 1e9 in f32), :62-65 (log_softmax), :239 (gather columns by ph_seq_id), :68-71 (edge sigmoid,
// rescale, clamp), :84 (edge_prob = clip(p[t] + p[t-1], 0, 1) in f64), :241-242 (f64 log -> f32).
//
// One CTA = 64 consecutive frames of one utterance.  Each warp owns 8 of those rows and handles
// them two at a time (two independent rows of loads in flight): coalesced, strided-view-aware row
// load -> masked max / sum(exp) by warp butterfly -> row parked in shared memory -> gather by the
// phoneme ids (also in shared memory) -> coalesced float4 store of the [Sp] emission row.
// Algorithmic HBM bytes per frame: V*sizeof(in) + 4 (edge logit) read, 4*Sp + 12 written.
#include "hfa_common.cuh"

#define HFA_EMIS_ROWS 64
#define HFA_EMIS_WARPS 8
#define HFA_EMIS_MAX_VPL 4        // register-staged path covers V <= 128; larger V streams via smem

namespace {

__device__ __forceinline__ float warp_max(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// utterance that owns row-block `blk` (row_blocks is an exclusive prefix, [n+1])
__device__ __forceinline__ int find_utt(const int32_t *row_blocks, int n, int blk)
{
    int lo = 0, hi = n;          // invariant: row_blocks[lo] <= blk < row_blocks[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (row_blocks[mid] <= blk) lo = mid;
        else hi = mid;
    }
    return lo;
}

// clamp((sigmoid(x) - 0.1) / 0.8, 0, 1), all f32 (:68-71)
__device__ __forceinline__ float edge_pred(float x)
{
    const float sg = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x)));
    const float v = __fdiv_rn(__fsub_rn(sg, 0.1f), 0.8f);
    return fminf(fmaxf(v, 0.0f), 1.0f);
}

__device__ __forceinline__ void edge_logs(float p, float p_prev, float2 &out)
{
    double ep = __dadd_rn((double)p, (double)p_prev);                 // :84 (f64 sum of f32 values)
    ep = fmin(fmax(ep, 0.0), 1.0);
    out.x = (float)log(__dadd_rn(ep, 1e-6));                          // :241
    out.y = (float)log(__dadd_rn(__dsub_rn(1.0, ep), 1e-6));          // :242  (1 - ep) + 1e-6
}

template <typename TIn>
__global__ void __launch_bounds__(HFA_EMIS_WARPS * 32)
hfa_emission_kernel(HfaWs ws, int n_utt, int V, int sp_cap)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int32_t *ids_sm = reinterpret_cast<int32_t *>(smem_raw);                 // [sp_cap]
    uint32_t *mask_sm = reinterpret_cast<uint32_t *>(ids_sm + sp_cap);       // [ceil(V/32)] keep bits
    const int mask_words = (V + 31) >> 5;
    float *rows_sm = reinterpret_cast<float *>(mask_sm + ((mask_words + 3) & ~3));  // [warps][V]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int u = find_utt(ws.row_blocks, n_utt, blockIdx.x);
    const HfaUtt m = ws.utt[u];
    const HfaInput in = ws.inputs[u];
    const int T = m.T, S = m.S, Sp = m.Sp;
    const int t_base = (blockIdx.x - ws.row_blocks[u]) * HFA_EMIS_ROWS;

    for (int w = tid; w < mask_words; w += blockDim.x) mask_sm[w] = (w == 0) ? 1u : 0u;  // id 0 (:39)
    __syncthreads();
    const int32_t *ids = ws.ids + m.seg_off;
    for (int s = tid; s < Sp; s += blockDim.x) {
        const int id = (s < S) ? ids[s] : 0;
        ids_sm[s] = id;
        atomicOr(&mask_sm[id >> 5], 1u << (id & 31));
    }
    __syncthreads();

    const TIn *frame = reinterpret_cast<const TIn *>(in.frame);
    float *rowbuf = rows_sm + warp * V;
    float *g_out = ws.emis + m.emis_off;

    if (V <= 32 * HFA_EMIS_MAX_VPL) {
        // two rows per iteration, values staged in registers
        for (int j = 0; j < HFA_EMIS_ROWS / HFA_EMIS_WARPS; j += 2) {
            float x[2][HFA_EMIS_MAX_VPL];
            int tr[2];
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                tr[r] = t_base + warp + HFA_EMIS_WARPS * (j + r);
                const TIn *src = frame + (int64_t)tr[r] * in.frame_st;
#pragma unroll
                for (int q = 0; q < HFA_EMIS_MAX_VPL; ++q) {
                    const int v = lane + 32 * q;
                    x[r][q] = HFA_NEG_INF;
                    if (tr[r] < T && v < V) {
                        float xv = hfa_to_float<TIn>(src[(int64_t)v * in.frame_sv]);
                        if (!((mask_sm[v >> 5] >> (v & 31)) & 1u)) xv = __fsub_rn(xv, 1e9f);  // :53
                        x[r][q] = xv;
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                if (tr[r] >= T) continue;                    // warp-uniform
                float mx = x[r][0];
#pragma unroll
                for (int q = 1; q < HFA_EMIS_MAX_VPL; ++q) mx = fmaxf(mx, x[r][q]);
                mx = warp_max(mx);
                float sum = 0.0f;
#pragma unroll
                for (int q = 0; q < HFA_EMIS_MAX_VPL; ++q) {
                    const int v = lane + 32 * q;
                    if (v < V) {
                        sum = __fadd_rn(sum, expf(__fsub_rn(x[r][q], mx)));
                        rowbuf[v] = x[r][q];
                    }
                }
                sum = warp_sum(sum);
                const float lse = logf(sum);
                __syncwarp();
                float *dst = g_out + (int64_t)tr[r] * Sp;
                for (int s4 = lane * 4; s4 < Sp; s4 += 128) {
                    const int4 id4 = *reinterpret_cast<const int4 *>(ids_sm + s4);
                    float4 o;
                    o.x = (s4 + 0 < S) ? __fsub_rn(__fsub_rn(rowbuf[id4.x], mx), lse) : HFA_NEG_INF;
                    o.y = (s4 + 1 < S) ? __fsub_rn(__fsub_rn(rowbuf[id4.y], mx), lse) : HFA_NEG_INF;
                    o.z = (s4 + 2 < S) ? __fsub_rn(__fsub_rn(rowbuf[id4.z], mx), lse) : HFA_NEG_INF;
                    o.w = (s4 + 3 < S) ? __fsub_rn(__fsub_rn(rowbuf[id4.w], mx), lse) : HFA_NEG_INF;
                    *reinterpret_cast<float4 *>(dst + s4) = o;
                }
                __syncwarp();
            }
        }
    } else {
        // wide vocabularies: one row at a time through the shared-memory row buffer
        for (int j = 0; j < HFA_EMIS_ROWS / HFA_EMIS_WARPS; ++j) {
            const int t = t_base + warp + HFA_EMIS_WARPS * j;
            if (t >= T) continue;
            const TIn *src = frame + (int64_t)t * in.frame_st;
            float mx = HFA_NEG_INF;
            for (int v = lane; v < V; v += 32) {
                float xv = hfa_to_float<TIn>(src[(int64_t)v * in.frame_sv]);
                if (!((mask_sm[v >> 5] >> (v & 31)) & 1u)) xv = __fsub_rn(xv, 1e9f);
                rowbuf[v] = xv;
                mx = fmaxf(mx, xv);
            }
            mx = warp_max(mx);
            float sum = 0.0f;
            for (int v = lane; v < V; v += 32) sum = __fadd_rn(sum, expf(__fsub_rn(rowbuf[v], mx)));
            sum = warp_sum(sum);
            const float lse = logf(sum);
            __syncwarp();
            float *dst = g_out + (int64_t)t * Sp;
            for (int s4 = lane * 4; s4 < Sp; s4 += 128) {
                const int4 id4 = *reinterpret_cast<const int4 *>(ids_sm + s4);
                float4 o;
                o.x = (s4 + 0 < S) ? __fsub_rn(__fsub_rn(rowbuf[id4.x], mx), lse) : HFA_NEG_INF;
                o.y = (s4 + 1 < S) ? __fsub_rn(__fsub_rn(rowbuf[id4.y], mx), lse) : HFA_NEG_INF;
                o.z = (s4 + 2 < S) ? __fsub_rn(__fsub_rn(rowbuf[id4.z], mx), lse) : HFA_NEG_INF;
                o.w = (s4 + 3 < S) ? __fsub_rn(__fsub_rn(rowbuf[id4.w], mx), lse) : HFA_NEG_INF;
                *reinterpret_cast<float4 *>(dst + s4) = o;
            }
            __syncwarp();
        }
    }

    // edge stream of this CTA's 64 frames, one thread per frame
    if (tid < HFA_EMIS_ROWS) {
        const int t = t_base + tid;
        if (t < T) {
            const TIn *edge = reinterpret_cast<const TIn *>(in.edge);
            const float p = edge_pred(hfa_to_float<TIn>(edge[(int64_t)t * in.edge_st]));
            const float pp =
                (t > 0) ? edge_pred(hfa_to_float<TIn>(edge[(int64_t)(t - 1) * in.edge_st])) : 0.0f;
            float2 lg;
            edge_logs(p, pp, lg);
            ws.edge2[m.edge_off + t] = lg;
            ws.edge_p[m.edge_off + t] = p;
        }
    }
}

// the reference's forward_pass inputs given directly (dense ragged), repacked into the workspace
__global__ void __launch_bounds__(256)
hfa_pack_kernel(HfaWs ws, int n_utt, const float *__restrict__ prob_log,
                const float *__restrict__ edge_log, const float *__restrict__ not_edge_log,
                const float *__restrict__ edge_pred_in)
{
    const int tid = threadIdx.x;
    const int u = find_utt(ws.row_blocks, n_utt, blockIdx.x);
    const HfaUtt m = ws.utt[u];
    const int T = m.T, S = m.S, Sp = m.Sp;
    const int t_base = (blockIdx.x - ws.row_blocks[u]) * HFA_EMIS_ROWS;
    const int rows = min(HFA_EMIS_ROWS, T - t_base);
    const float *src = prob_log + m.cell_off + (int64_t)t_base * S;
    float *dst = ws.emis + m.emis_off + (int64_t)t_base * Sp;
    for (int i = tid; i < rows * Sp; i += blockDim.x) {
        const int r = i / Sp, s = i - r * Sp;
        dst[i] = (s < S) ? src[(int64_t)r * S + s] : HFA_NEG_INF;
    }
    if (tid < rows) {
        const int t = t_base + tid;
        ws.edge2[m.edge_off + t] = make_float2(edge_log[m.frame_off + t], not_edge_log[m.frame_off + t]);
        ws.edge_p[m.edge_off + t] = edge_pred_in ? edge_pred_in[m.frame_off + t] : 0.0f;
    }
}

}  // namespace

cudaError_t hfa_launch_emission(const HfaLaunchCtx &c, int total_row_blocks, int max_sp, int dtype)
{
    if (total_row_blocks <= 0) return cudaSuccess;
    const int V = c.vocab;
    const int mask_words = (V + 31) >> 5;
    const size_t smem = (size_t)max_sp * 4 + (size_t)((mask_words + 3) & ~3) * 4 +
                        (size_t)HFA_EMIS_WARPS * V * 4;
    cudaError_t e;
#define HFA_EMIS_LAUNCH(TIn)                                                                       \
    e = cudaFuncSetAttribute(hfa_emission_kernel<TIn>,                                             \
                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);              \
    if (e != cudaSuccess) return e;                                                                \
    hfa_emission_kernel<TIn><<<total_row_blocks, HFA_EMIS_WARPS * 32, smem, c.stream>>>(           \
        c.ws, c.n_utt, V, max_sp)
    if (dtype == 0) { HFA_EMIS_LAUNCH(float); }
    else if (dtype == 1) { HFA_EMIS_LAUNCH(__half); }
    else if (dtype == 2) { HFA_EMIS_LAUNCH(__nv_bfloat16); }
    else return cudaErrorInvalidValue;
#undef HFA_EMIS_LAUNCH
    return cudaGetLastError();
}

cudaError_t hfa_launch_pack(const HfaLaunchCtx &c, int total_row_blocks, const float *prob_log,
                            const float *edge_log, const float *not_edge_log, const float *edge_pred)
{
    if (total_row_blocks <= 0) return cudaSuccess;
    hfa_pack_kernel<<<total_row_blocks, 256, 0, c.stream>>>(c.ws, c.n_utt, prob_log, edge_log,
                                                            not_edge_log, edge_pred);
    return cudaGetLastError();
}
