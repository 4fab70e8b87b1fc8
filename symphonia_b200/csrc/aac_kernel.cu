// AAC-LC synthesis for sm_100a: TNS (aac/ics/tns.rs:149-199) then the filterbank of Dsp::synth
// (aac/dsp.rs:57-158): 1024-point or 8 x 128-point IMDCT, sine / KBD windows, the four window
// sequences, overlap-add through the per-channel `delay` line.
//
// Work decomposition: `delay` is overwritten from the current frame only, so a
// channel's frames are cut into chunks of consecutive frames; one 64-thread CTA walks one chunk with
// the delay line in shared memory.  A chunk that does not start its run recomputes the previous
// frame's delay (one halo frame, output suppressed).  TNS is a serial recurrence along frequency:
// it runs in a pre-pass, one LANE per filter, into a scratch copy of the spectra it touches.
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/symgpu.h"
#include "codec_kernels.h"
#include "imdct.cuh"
#include "tables.h"

namespace symgpu {
namespace {

constexpr int kAacThreads = 64;
constexpr int P0 = 512 - 64, P1 = 512 + 64; // SHORT_WIN_POINT0/1, aac/dsp.rs:19-20

// ---- TNS ------------------------------------------------------------------------------------
// coeffs[i] -= coeffs[i -/+ (j+1)] * lpc[j], j ascending, in place (tns.rs:183-196).  The last ORDER
// outputs live in registers; the `j < m` guard reproduces `order.min(m)` for the first lines.
template <int ORDER>
__device__ __forceinline__ void tns_filter(float* c, int start, int end, bool down, const float* __restrict__ lpc_g) {
    float lpc[ORDER], h[ORDER];
#pragma unroll
    for (int j = 0; j < ORDER; ++j) {
        lpc[j] = lpc_g[j];
        h[j] = 0.0f;
    }
    const int len = end - start;
    for (int m = 0; m < len; ++m) {
        const int i = down ? end - 1 - m : start + m;
        float v = c[i];
#pragma unroll
        for (int j = 0; j < ORDER; ++j)
            if (j < m) v -= h[j] * lpc[j];
#pragma unroll
        for (int j = ORDER - 1; j > 0; --j) h[j] = h[j - 1];
        h[0] = v;
        c[i] = v;
    }
}

__device__ void tns_dispatch(float* c, const symgpu_aac_tns& f) {
    const int start = f.start, end = f.end;
    const bool down = f.direction != 0;
    switch (f.order) {
#define TNS_CASE(N) case N: tns_filter<N>(c, start, end, down, f.lpc); break;
        TNS_CASE(1) TNS_CASE(2) TNS_CASE(3) TNS_CASE(4) TNS_CASE(5) TNS_CASE(6) TNS_CASE(7) TNS_CASE(8) TNS_CASE(9)
        TNS_CASE(10) TNS_CASE(11) TNS_CASE(12) TNS_CASE(13) TNS_CASE(14) TNS_CASE(15) TNS_CASE(16) TNS_CASE(17)
        TNS_CASE(18) TNS_CASE(19) TNS_CASE(20)
#undef TNS_CASE
        default: break;
    }
}

// One warp per channel-frame; lanes = filters of that channel (their line ranges are disjoint).
__global__ void __launch_bounds__(256) aac_tns_kernel(const symgpu_aac_unit* __restrict__ units,
                                                      const symgpu_aac_tns* __restrict__ tns, const float* __restrict__ coeffs,
                                                      float* __restrict__ scratch, uint32_t n_units) {
    __shared__ float buf[8][1024];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t u = blockIdx.x * 8 + warp;
    if (u >= n_units) return;
    const symgpu_aac_unit unit = units[u];
    if (unit.n_tns == 0) return;
    float* c = buf[warp];
    const float4* src = reinterpret_cast<const float4*>(coeffs + (size_t)u * 1024);
    for (int i = lane; i < 256; i += 32) reinterpret_cast<float4*>(c)[i] = __ldg(src + i);
    __syncwarp();
    for (int f = lane; f < unit.n_tns; f += 32) tns_dispatch(c, tns[unit.tns_first + f]);
    __syncwarp();
    float4* dst = reinterpret_cast<float4*>(scratch + (size_t)u * 1024);
    for (int i = lane; i < 256; i += 32) dst[i] = reinterpret_cast<float4*>(c)[i];
}

// ---- filterbank --------------------------------------------------------------------------------
struct AacSmem {
    float spec[1024];
    float out[2048];
    float delay[1024];
    float2 z[zpad_len(512)];
};

__global__ void __launch_bounds__(kAacThreads) aac_synth_kernel(AacArgs a) {
    __shared__ AacSmem sm;
    const int tid = threadIdx.x;
    const CodecChunk ck = a.chunks[blockIdx.x];
    const int ch = ck.channel;
    const CodecTables* __restrict__ tab = a.tab;
    const FftTables* ft = reinterpret_cast<const FftTables*>(tab->fft_lit16);
    const uint32_t gen = a.gen[ck.stream];
    const float* st_in = a.states + (((size_t)ck.stream * 2 + (gen & 1)) * 2 + ch) * 1024;
    float* st_out = a.states + (((size_t)ck.stream * 2 + ((gen + 1) & 1)) * 2 + ch) * 1024;
    const bool load_state = ck.flags & kChunkLoadState;

    if (load_state)
        for (int i = tid; i < 1024; i += kAacThreads) sm.delay[i] = st_in[i];
    __syncthreads();

    const int f_begin = (int)ck.first - (load_state ? 0 : 1);
    const int f_end = (int)ck.first + ck.count;
    for (int f = f_begin; f < f_end; ++f) {
        const bool emit = f >= (int)ck.first;
        const size_t unit_idx = 2 * (size_t)f + ch;
        const symgpu_aac_unit u = a.units[unit_idx];
        const float* src = (u.n_tns ? a.tns_scratch : a.coeffs) + unit_idx * 1024;
        for (int i = tid; i < 256; i += kAacThreads)
            reinterpret_cast<float4*>(sm.spec)[i] = __ldg(reinterpret_cast<const float4*>(src) + i);
        __syncthreads();
        const int seq = u.window_sequence;
        if (seq != SYMGPU_AAC_EIGHT_SHORT)
            imdct_blocks<9>(sm.spec, sm.out, sm.z, 1, reinterpret_cast<const float2*>(tab->aac_tw_long), ft, tid, kAacThreads);
        else
            imdct_blocks<6>(sm.spec, sm.out, sm.z, 8, reinterpret_cast<const float2*>(tab->aac_tw_short), ft, tid, kAacThreads);

        const float* lw = u.window_shape ? tab->aac_kbd_long : tab->aac_sine_long;
        const float* sw = u.window_shape ? tab->aac_kbd_short : tab->aac_sine_short;
        const float* plw = u.prev_window_shape ? tab->aac_kbd_long : tab->aac_sine_long;
        const float* psw = u.prev_window_shape ? tab->aac_kbd_short : tab->aac_sine_short;
        float* dst = a.pcm + unit_idx * 1024;

        // pcm_short[x] of aac/dsp.rs:86-101, rebuilt per sample with the reference's operation order:
        // the second half of window w-1 is written first (assignment for w-1 = 0, "0.0 +=" otherwise),
        // then the first half of window w is added.
        auto pcm_short = [&](int x) -> float {
            const int w = x >> 7, i = x & 127;
            if (w == 0) return sm.out[i] * __ldg(psw + i);
            const float t2 = sm.out[256 * (w - 1) + 128 + i] * __ldg(sw + 127 - i);
            const float prev = (w == 1) ? t2 : 0.0f + t2;
            if (w == 8) return prev;
            return prev + sm.out[256 * w + i] * __ldg(sw + i);
        };

#pragma unroll 4
        for (int i = tid; i < 1024; i += kAacThreads) {
            const float d = sm.delay[i];
            float y, nd;
            switch (seq) {
                case SYMGPU_AAC_ONLY_LONG:
                    y = d + (sm.out[i] * __ldg(plw + i));
                    nd = sm.out[i + 1024] * __ldg(lw + 1023 - i);
                    break;
                case SYMGPU_AAC_LONG_START:
                    y = d + (sm.out[i] * __ldg(plw + i));
                    nd = i < P0 ? sm.out[i + 1024] : i < P1 ? sm.out[i + 1024] * __ldg(sw + 127 - (i - P0)) : 0.0f;
                    break;
                case SYMGPU_AAC_EIGHT_SHORT:
                    y = i < P0 ? d : d + pcm_short(i - P0);
                    nd = i < P1 ? pcm_short(i + 512 + 64) : 0.0f;
                    break;
                default: // LONG_STOP
                    y = i < P0 ? d : i < P1 ? d + sm.out[i] * __ldg(psw + i - P0) : d + sm.out[i];
                    nd = sm.out[i + 1024] * __ldg(lw + 1023 - i);
                    break;
            }
            if (emit) dst[i] = y;
            sm.delay[i] = nd;
        }
        __syncthreads();
    }

    if (ck.flags & kChunkStoreState)
        for (int i = tid; i < 1024; i += kAacThreads) st_out[i] = sm.delay[i];

    // launch epilogue: the last CTA publishes the new state generation (see mp3_kernel.cu)
    __shared__ bool is_last;
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        is_last = atomicAdd(a.done, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (is_last) {
        for (unsigned i = tid; i < gridDim.x; i += kAacThreads)
            if ((a.chunks[i].flags & kChunkStoreState) && a.chunks[i].channel == 0) a.gen[a.chunks[i].stream] += 1;
        if (tid == 0) *a.done = 0;
    }
}

} // namespace

cudaError_t aac_launch(const AacArgs& a, uint32_t n_units, bool any_tns, int n_chunks, cudaStream_t stream) {
    if (any_tns) {
        aac_tns_kernel<<<(n_units + 7) / 8, 256, 0, stream>>>(a.units, a.tns, a.coeffs, a.tns_scratch_rw, n_units);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    aac_synth_kernel<<<n_chunks, kAacThreads, 0, stream>>>(a);
    return cudaGetLastError();
}

} // namespace symgpu
