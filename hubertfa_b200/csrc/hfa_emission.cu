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

__device__ __forceinline__ float lds_f32(uint32_t addr)
{
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}

// VPL = vocabulary entries per lane (V <= 32 * VPL) for the register-staged path, 0 = any V.
// Shared memory: ids [sp_cap] (pads -> V), keep-mask words, row buffers [warps][rows][V + 1] whose
// extra slot V holds -inf: a pad column gathers that slot and comes out as -inf without a select.
template <typename TIn, int VPL>
__global__ void __launch_bounds__(HFA_EMIS_WARPS * 32)
hfa_emission_kernel(HfaWs ws, int n_utt, int V, int sp_cap)
{
    constexpr int RPW = HFA_EMIS_ROWS / HFA_EMIS_WARPS;                      // rows per warp
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int32_t *ids_sm = reinterpret_cast<int32_t *>(smem_raw);                 // [sp_cap]
    uint32_t *mask_sm = reinterpret_cast<uint32_t *>(ids_sm + sp_cap);       // [ceil(V/32)] keep bits
    const int mask_words = (V + 31) >> 5;
    float *rows_sm = reinterpret_cast<float *>(mask_sm + ((mask_words + 3) & ~3));
    const int VS = V + 1;                                                    // row buffer stride

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int u = find_utt(ws.row_blocks, n_utt, blockIdx.x);
    const HfaUtt m = ws.utt[u];
    const HfaInput in = ws.inputs[u];
    const int T = m.T, S = m.S, Sp = m.Sp;
    const int t_base = (blockIdx.x - ws.row_blocks[u]) * HFA_EMIS_ROWS;
    const TIn *frame = reinterpret_cast<const TIn *>(in.frame);
    float *g_out = ws.emis + m.emis_off;

    // (1) every warp puts the loads of ALL its rows in flight before anything else.  Branch-free:
    //     out-of-range lanes / rows read a clamped (valid) address and are replaced by -inf.
    float x[RPW][VPL > 0 ? VPL : 1];
    if constexpr (VPL > 0) {
        const int64_t row_step = (int64_t)HFA_EMIS_WARPS * in.frame_st;
#pragma unroll
        for (int q = 0; q < VPL; ++q) {
            const int v = lane + 32 * q;
            const TIn *col = frame + (int64_t)min(v, V - 1) * in.frame_sv;
#pragma unroll
            for (int r = 0; r < RPW; ++r) {
                const int t = min(t_base + warp + HFA_EMIS_WARPS * r, T - 1);
                const float xv = hfa_to_float<TIn>(col[(int64_t)t * in.frame_st]);
                x[r][q] = (v < V) ? xv : HFA_NEG_INF;
            }
        }
        (void)row_step;
    }

    // (2) phoneme ids and the keep-mask of this utterance -> shared memory
    for (int w = tid; w < mask_words; w += blockDim.x) mask_sm[w] = (w == 0) ? 1u : 0u;  // id 0 (:39)
    __syncthreads();
    const int32_t *ids = ws.ids + m.seg_off;
    for (int s = tid; s < Sp; s += blockDim.x) {
        const int id = (s < S) ? ids[s] : V;
        ids_sm[s] = id;
        if (s < S) atomicOr(&mask_sm[id >> 5], 1u << (id & 31));
    }
    __syncthreads();

    if constexpr (VPL > 0) {
        // row buffers: stride 32*VPL+1 so every lane stores unconditionally; slots V.. hold the -inf
        // of the out-of-range lanes, which is exactly the pad sentinel the gather needs at slot V
        constexpr int RS = 32 * VPL + 1;
        float *rowbuf = rows_sm + (size_t)warp * RPW * RS;
        const uint32_t rowbuf_sa = hfa_smem_u32(rowbuf);
        // per-lane constants: the 1e9 penalty of its vocabulary entries (:53) and the byte offsets
        // of the (up to 8) row-buffer slots it gathers
        float pen[VPL];
#pragma unroll
        for (int q = 0; q < VPL; ++q) {
            const int v = min(lane + 32 * q, V - 1);
            pen[q] = ((mask_sm[v >> 5] >> (v & 31)) & 1u) ? 0.0f : 1e9f;
        }
        uint32_t goff[2][4];
#pragma unroll
        for (int it = 0; it < 2; ++it) {
            const int s4 = min(lane * 4 + 128 * it, Sp - 4);
            const int4 id4 = *reinterpret_cast<const int4 *>(ids_sm + s4);
            goff[it][0] = (uint32_t)id4.x * 4u; goff[it][1] = (uint32_t)id4.y * 4u;
            goff[it][2] = (uint32_t)id4.z * 4u; goff[it][3] = (uint32_t)id4.w * 4u;
        }

        // (3) masked max / log-sum-exp of the 8 rows (independent chains), rows parked in smem
        float mx[RPW], lse[RPW];
#pragma unroll
        for (int r = 0; r < RPW; ++r) {
            float mr = HFA_NEG_INF;
#pragma unroll
            for (int q = 0; q < VPL; ++q) {
                x[r][q] = __fsub_rn(x[r][q], pen[q]);                        // :53 (x - 0 is exact)
                mr = fmaxf(mr, x[r][q]);
            }
            mx[r] = warp_max(mr);
        }
#pragma unroll
        for (int r = 0; r < RPW; ++r) {
            float sum = 0.0f;
#pragma unroll
            for (int q = 0; q < VPL; ++q) {
                sum = __fadd_rn(sum, expf(__fsub_rn(x[r][q], mx[r])));       // exp(-inf) = 0 for v >= V
                rowbuf[r * RS + lane + 32 * q] = x[r][q];
            }
            lse[r] = logf(warp_sum(sum));
        }
        __syncwarp();

        // (4) gather by phoneme id and store: out[t][s] = (x[id[s]] - max) - lse
        const int rows_here = min(RPW, (T - t_base - warp + HFA_EMIS_WARPS - 1) / HFA_EMIS_WARPS);
        float *dst = g_out + (int64_t)(t_base + warp) * Sp + lane * 4;
        const int dst_step = HFA_EMIS_WARPS * Sp;
        const bool st0 = lane * 4 < Sp, st1 = lane * 4 + 128 < Sp;
#pragma unroll
        for (int r = 0; r < RPW; ++r) {
            if (r >= rows_here) break;                                       // warp-uniform
            const uint32_t rb = rowbuf_sa + (uint32_t)(r * RS) * 4u;
            float4 o0, o1;
            o0.x = __fsub_rn(__fsub_rn(lds_f32(rb + goff[0][0]), mx[r]), lse[r]);
            o0.y = __fsub_rn(__fsub_rn(lds_f32(rb + goff[0][1]), mx[r]), lse[r]);
            o0.z = __fsub_rn(__fsub_rn(lds_f32(rb + goff[0][2]), mx[r]), lse[r]);
            o0.w = __fsub_rn(__fsub_rn(lds_f32(rb + goff[0][3]), mx[r]), lse[r]);
            if (st0) *reinterpret_cast<float4 *>(dst + r * dst_step) = o0;
            if (Sp > 128) {                                                  // warp-uniform
                o1.x = __fsub_rn(__fsub_rn(lds_f32(rb + goff[1][0]), mx[r]), lse[r]);
                o1.y = __fsub_rn(__fsub_rn(lds_f32(rb + goff[1][1]), mx[r]), lse[r]);
                o1.z = __fsub_rn(__fsub_rn(lds_f32(rb + goff[1][2]), mx[r]), lse[r]);
                o1.w = __fsub_rn(__fsub_rn(lds_f32(rb + goff[1][3]), mx[r]), lse[r]);
                if (st1) *reinterpret_cast<float4 *>(dst + r * dst_step + 128) = o1;
            }
        }
        if (Sp > 256) {                                                      // long phoneme sequences
            for (int r = 0; r < rows_here; ++r) {
                for (int s4 = lane * 4 + 256; s4 < Sp; s4 += 128) {
                    const int4 id4 = *reinterpret_cast<const int4 *>(ids_sm + s4);
                    float4 o;
                    o.x = __fsub_rn(__fsub_rn(rowbuf[r * RS + id4.x], mx[r]), lse[r]);
                    o.y = __fsub_rn(__fsub_rn(rowbuf[r * RS + id4.y], mx[r]), lse[r]);
                    o.z = __fsub_rn(__fsub_rn(rowbuf[r * RS + id4.z], mx[r]), lse[r]);
                    o.w = __fsub_rn(__fsub_rn(rowbuf[r * RS + id4.w], mx[r]), lse[r]);
                    *reinterpret_cast<float4 *>(dst + r * dst_step + (s4 - lane * 4)) = o;
                }
            }
        }
    } else {
        // wide vocabularies: one row at a time through the shared-memory row buffer
        float *rowbuf = rows_sm + (size_t)warp * VS;
        if (lane == 0) rowbuf[V] = HFA_NEG_INF;
        __syncwarp();
        for (int j = 0; j < RPW; ++j) {
            const int t = t_base + warp + HFA_EMIS_WARPS * j;
            if (t >= T) continue;
            const TIn *src = frame + (int64_t)t * in.frame_st;
            float mr = HFA_NEG_INF;
            for (int v = lane; v < V; v += 32) {
                float xv = hfa_to_float<TIn>(src[(int64_t)v * in.frame_sv]);
                if (!((mask_sm[v >> 5] >> (v & 31)) & 1u)) xv = __fsub_rn(xv, 1e9f);
                rowbuf[v] = xv;
                mr = fmaxf(mr, xv);
            }
            mr = warp_max(mr);
            float sum = 0.0f;
            for (int v = lane; v < V; v += 32) sum = __fadd_rn(sum, expf(__fsub_rn(rowbuf[v], mr)));
            const float l = logf(warp_sum(sum));
            __syncwarp();
            float *dst = g_out + (int64_t)t * Sp;
            for (int s4 = lane * 4; s4 < Sp; s4 += 128) {
                const int4 id4 = *reinterpret_cast<const int4 *>(ids_sm + s4);
                float4 o;
                o.x = __fsub_rn(__fsub_rn(rowbuf[id4.x], mr), l);
                o.y = __fsub_rn(__fsub_rn(rowbuf[id4.y], mr), l);
                o.z = __fsub_rn(__fsub_rn(rowbuf[id4.z], mr), l);
                o.w = __fsub_rn(__fsub_rn(rowbuf[id4.w], mr), l);
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

template <typename TIn>
static cudaError_t launch_emission_t(const HfaLaunchCtx &c, int blocks, int max_sp)
{
    const int V = c.vocab;
    const int vpl = (V + 31) / 32;
    const int mask_words = (V + 31) >> 5;
    const size_t rows = (vpl <= HFA_EMIS_MAX_VPL) ? (size_t)HFA_EMIS_ROWS : (size_t)HFA_EMIS_WARPS;
    const size_t stride = (vpl <= HFA_EMIS_MAX_VPL) ? (size_t)(32 * vpl + 1) : (size_t)(V + 1);
    const size_t smem = (size_t)max_sp * 4 + (size_t)((mask_words + 3) & ~3) * 4 + rows * stride * 4;
    cudaError_t e;
#define HFA_EMIS_LAUNCH(VPL)                                                                       \
    e = cudaFuncSetAttribute(hfa_emission_kernel<TIn, VPL>,                                        \
                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);              \
    if (e != cudaSuccess) return e;                                                                \
    hfa_emission_kernel<TIn, VPL><<<blocks, HFA_EMIS_WARPS * 32, smem, c.stream>>>(c.ws, c.n_utt,  \
                                                                                  V, max_sp)
    switch (vpl <= HFA_EMIS_MAX_VPL ? vpl : 0) {
        case 1: HFA_EMIS_LAUNCH(1); break;
        case 2: HFA_EMIS_LAUNCH(2); break;
        case 3: HFA_EMIS_LAUNCH(3); break;
        case 4: HFA_EMIS_LAUNCH(4); break;
        default: HFA_EMIS_LAUNCH(0); break;
    }
#undef HFA_EMIS_LAUNCH
    return cudaGetLastError();
}

cudaError_t hfa_launch_emission(const HfaLaunchCtx &c, int total_row_blocks, int max_sp, int dtype)
{
    if (total_row_blocks <= 0) return cudaSuccess;
    if (dtype == 0) return launch_emission_t<float>(c, total_row_blocks, max_sp);
    if (dtype == 1) return launch_emission_t<__half>(c, total_row_blocks, max_sp);
    if (dtype == 2) return launch_emission_t<__nv_bfloat16>(c, total_row_blocks, max_sp);
    return cudaErrorInvalidValue;
}

cudaError_t hfa_launch_pack(const HfaLaunchCtx &c, int total_row_blocks, const float *prob_log,
                            const float *edge_log, const float *not_edge_log, const float *edge_pred)
{
    if (total_row_blocks <= 0) return cudaSuccess;
    hfa_pack_kernel<<<total_row_blocks, 256, 0, c.stream>>>(c.ws, c.n_utt, prob_log, edge_log,
                                                            not_edge_log, edge_pred);
    return cudaGetLastError();
}
