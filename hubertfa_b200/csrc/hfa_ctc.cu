// hfa_ctc.cu -- greedy CTC decode of one utterance on the device.
//
// Reference: AlignmentDecoder.ctc(), tools/alignment_decoder.py:145-150 (used by validation_step,
// networks/task/forced_alignment.py:413):
//     ctc = argmax(ctc_logits, -1);  keep frame t iff ctc[t] != ctc[t-1] (ctc[-1] := 0) and ctc[t] != 0
// One CTA: every warp takes frames round-robin for the argmax (first maximum wins, like numpy), then
// the CTA compacts the kept ids in frame order with a running prefix sum.  Index work: bit-exact.
#include "hfa_common.cuh"

namespace {

constexpr int HFA_CTC_THREADS = 1024;

template <typename TIn>
__global__ void __launch_bounds__(HFA_CTC_THREADS)
hfa_ctc_greedy_kernel(const TIn *__restrict__ logits, int T, int V, int64_t st_t, int64_t st_v,
                      int32_t *__restrict__ arg, int32_t *__restrict__ out_ids, int32_t *__restrict__ out_len)
{
    __shared__ int warp_sum[HFA_CTC_THREADS / 32];
    __shared__ int carry_sm;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // 1. argmax per frame (f32 compare after .float(), :77; ties -> the lowest index)
    for (int t = warp; t < T; t += HFA_CTC_THREADS / 32) {
        const TIn *row = logits + (int64_t)t * st_t;
        float best = HFA_NEG_INF;
        int bi = 0x7fffffff;
        for (int v = lane; v < V; v += 32) {
            const float x = hfa_to_float<TIn>(row[(int64_t)v * st_v]);
            if (x > best || bi == 0x7fffffff) {       // the first element always enters (all -inf rows -> 0)
                best = x;
                bi = v;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (oi != 0x7fffffff && (bi == 0x7fffffff || ob > best || (ob == best && oi < bi))) {
                best = ob;
                bi = oi;
            }
        }
        if (lane == 0) arg[t] = bi;
    }
    if (tid == 0) carry_sm = 0;
    __syncthreads();
    // 2. keep / compact in frame order
    for (int base = 0; base < T; base += HFA_CTC_THREADS) {
        const int t = base + tid;
        int id = 0, keep = 0;
        if (t < T) {
            id = arg[t];
            const int prev = (t > 0) ? arg[t - 1] : 0;
            keep = (id != prev && id != 0) ? 1 : 0;
        }
        int incl = keep;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) warp_sum[warp] = incl;
        __syncthreads();
        int pos = carry_sm + incl;
        for (int q = 0; q < warp; ++q) pos += warp_sum[q];
        __syncthreads();
        if (keep) out_ids[pos - 1] = id;
        if (tid == HFA_CTC_THREADS - 1) carry_sm = pos;
        __syncthreads();
    }
    if (tid == 0) *out_len = carry_sm;
}

}  // namespace

cudaError_t hfa_launch_ctc_greedy(const void *logits, int dtype, int T, int V, int64_t st_t, int64_t st_v,
                                  int32_t *arg, int32_t *out_ids, int32_t *out_len, cudaStream_t stream)
{
    if (dtype == 0)
        hfa_ctc_greedy_kernel<float><<<1, HFA_CTC_THREADS, 0, stream>>>(static_cast<const float *>(logits), T, V,
                                                                        st_t, st_v, arg, out_ids, out_len);
    else if (dtype == 1)
        hfa_ctc_greedy_kernel<__half><<<1, HFA_CTC_THREADS, 0, stream>>>(static_cast<const __half *>(logits), T, V,
                                                                         st_t, st_v, arg, out_ids, out_len);
    else if (dtype == 2)
        hfa_ctc_greedy_kernel<__nv_bfloat16><<<1, HFA_CTC_THREADS, 0, stream>>>(
            static_cast<const __nv_bfloat16 *>(logits), T, V, st_t, st_v, arg, out_ids, out_len);
    else
        return cudaErrorInvalidValue;
    return cudaGetLastError();
}
