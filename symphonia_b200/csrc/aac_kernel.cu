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
// One CTA = one chunk of <= K consecutive frames of one channel, plus one slot for the frame before
// the chunk (its delay line).  Each frame has its own group of 64 threads and its own named barrier,
// so the K+1 IMDCTs proceed independently; the window / overlap step then reads the IMDCT output of
// frame f and of frame f-1 (the `delay` of the reference is a pure function of frame f-1's output).
constexpr int kAacK = kAacChunkFrames;
struct alignas(16) AacFrameSmem {
    float out[2048];          // spectrum (first 1024 floats) until the pre-twiddle has consumed it, then pcm_long
    float2 z[zpad_len(512)];
};
// Twiddle tables staged in shared memory once per CTA: the 200 KB of frame slots leave almost no L1,
// and the FFT passes read twiddles at lane-dependent indices (ncu: long_scoreboard 7.7 per issue with
// the tables in global memory).
struct alignas(16) AacTabSmem {
    float2 fft[8 + 16 + 480]; // FftTables prefix: lit16, lit32, merge tables of sizes 64..512
    float2 tw_long[512];
    float2 tw_short[64];
};

// delay[i] after a frame with IMDCT output `out` (aac/dsp.rs:131-157) -- what the NEXT frame overlaps with.
__device__ __forceinline__ float aac_new_delay(int seq, const float* out, const float* __restrict__ lw,
                                               const float* __restrict__ sw, const float* __restrict__ psw, int i);

// pcm_short[x] of aac/dsp.rs:86-101, rebuilt per sample with the reference's operation order: the
// second half of window w-1 is written first (assignment for w-1 = 0, "0.0 +=" otherwise), then the
// first half of window w is added.
__device__ __forceinline__ float aac_pcm_short(const float* out, const float* __restrict__ sw,
                                               const float* __restrict__ psw, int x) {
    const int w = x >> 7, i = x & 127;
    if (w == 0) return out[i] * __ldg(psw + i);
    const float t2 = out[256 * (w - 1) + 128 + i] * __ldg(sw + 127 - i);
    const float prev = (w == 1) ? t2 : 0.0f + t2;
    if (w == 8) return prev;
    return prev + out[256 * w + i] * __ldg(sw + i);
}

__device__ __forceinline__ float aac_new_delay(int seq, const float* out, const float* __restrict__ lw,
                                               const float* __restrict__ sw, const float* __restrict__ psw, int i) {
    switch (seq) {
        case SYMGPU_AAC_ONLY_LONG:
        case SYMGPU_AAC_LONG_STOP: return out[i + 1024] * __ldg(lw + 1023 - i);
        case SYMGPU_AAC_EIGHT_SHORT: return i < P1 ? aac_pcm_short(out, sw, psw, i + 512 + 64) : 0.0f;
        default: // LONG_START
            return i < P0 ? out[i + 1024] : i < P1 ? out[i + 1024] * __ldg(sw + 127 - (i - P0)) : 0.0f;
    }
}

__global__ void __launch_bounds__((kAacK + 1) * 64, 2) aac_synth_kernel(AacArgs a) {
    extern __shared__ __align__(16) unsigned char aac_raw[];
    AacFrameSmem* fs = reinterpret_cast<AacFrameSmem*>(aac_raw);
    AacTabSmem& ts = *reinterpret_cast<AacTabSmem*>(aac_raw + (kAacK + 1) * sizeof(AacFrameSmem));
    __shared__ bool is_last;
    const int tid = threadIdx.x;
    const int grp = tid >> 6, gt = tid & 63; // frame slot of this thread, thread within the slot's group
    const CodecChunk ck = a.chunks[blockIdx.x];
    const int ch = ck.channel;
    const CodecTables* __restrict__ tab = a.tab;
    {
        const float2* g_fft = reinterpret_cast<const float2*>(tab->fft_lit16);
        const float2* g_twl = reinterpret_cast<const float2*>(tab->aac_tw_long);
        const float2* g_tws = reinterpret_cast<const float2*>(tab->aac_tw_short);
        for (int i = tid; i < 504; i += blockDim.x) ts.fft[i] = __ldg(g_fft + i);
        for (int i = tid; i < 512; i += blockDim.x) ts.tw_long[i] = __ldg(g_twl + i);
        if (tid < 64) ts.tw_short[tid] = __ldg(g_tws + tid);
    }
    __syncthreads();
    const FftTables* ft = reinterpret_cast<const FftTables*>(ts.fft);
    const uint32_t gen = a.gen[ck.stream];
    const float* st_in = a.states + (((size_t)ck.stream * 2 + (gen & 1)) * 2 + ch) * 1024;
    float* st_out = a.states + (((size_t)ck.stream * 2 + ((gen + 1) & 1)) * 2 + ch) * 1024;
    const bool load_state = ck.flags & kChunkLoadState;
    const int count = ck.count;

    // slot 0 = the frame before the chunk (or the stream state), slot k = chunk frame k-1
    const int f = (int)ck.first - 1 + grp;
    const bool have_frame = grp <= count && (grp > 0 || !load_state);
    symgpu_aac_unit u = {};
    if (have_frame) {
        const size_t unit_idx = 2 * (size_t)f + ch;
        u = a.units[unit_idx];
        AacFrameSmem& me = fs[grp];
        const float* src = (u.n_tns ? a.tns_scratch : a.coeffs) + unit_idx * 1024;
        for (int i = gt; i < 256; i += 64) reinterpret_cast<float4*>(me.out)[i] = __ldg(reinterpret_cast<const float4*>(src) + i);
        NamedSync sync{1 + grp, 64};
        sync();
        // the spectrum sits in out[0..1024); the pre-twiddle reads all of it before anything is written back
        if (u.window_sequence != SYMGPU_AAC_EIGHT_SHORT)
            imdct_blocks<9>(me.out, me.out, me.z, 1, ts.tw_long, ft, gt, 64, sync);
        else
            imdct_blocks<6>(me.out, me.out, me.z, 8, ts.tw_short, ft, gt, 64, sync);
    } else if (grp == 0) {
        // run start: slot 0 holds the delay line itself (stored in out[1024..2048))
        for (int i = gt; i < 1024; i += 64) fs[0].out[1024 + i] = st_in[i];
    }
    __syncthreads();

    // window + overlap-add (aac/dsp.rs:103-129): thread = (frame slot, sample)
    if (grp >= 1 && grp <= count) {
        const float* out = fs[grp].out;
        const float* pout = fs[grp - 1].out;
        const symgpu_aac_unit pu = (grp > 1 || !load_state) ? a.units[2 * (size_t)(f - 1) + ch] : symgpu_aac_unit{};
        const bool prev_is_state = grp == 1 && load_state;
        const int seq = u.window_sequence, pseq = pu.window_sequence;
        const float* sw = u.window_shape ? tab->aac_kbd_short : tab->aac_sine_short;
        const float* plw = u.prev_window_shape ? tab->aac_kbd_long : tab->aac_sine_long;
        const float* psw = u.prev_window_shape ? tab->aac_kbd_short : tab->aac_sine_short;
        // windows of the PREVIOUS frame, for its delay line
        const float* q_lw = pu.window_shape ? tab->aac_kbd_long : tab->aac_sine_long;
        const float* q_sw = pu.window_shape ? tab->aac_kbd_short : tab->aac_sine_short;
        const float* q_psw = pu.prev_window_shape ? tab->aac_kbd_short : tab->aac_sine_short;
        float* dst = a.pcm + (2 * (size_t)f + ch) * 1024;
#pragma unroll 4
        for (int i = gt; i < 1024; i += 64) {
            const float d = prev_is_state ? pout[1024 + i] : aac_new_delay(pseq, pout, q_lw, q_sw, q_psw, i);
            float y;
            switch (seq) {
                case SYMGPU_AAC_ONLY_LONG:
                case SYMGPU_AAC_LONG_START: y = d + (out[i] * __ldg(plw + i)); break;
                case SYMGPU_AAC_EIGHT_SHORT: y = i < P0 ? d : d + aac_pcm_short(out, sw, psw, i - P0); break;
                default: y = i < P0 ? d : i < P1 ? d + out[i] * __ldg(psw + i - P0) : d + out[i]; break; // LONG_STOP
            }
            dst[i] = y;
        }
        if (grp == count && (ck.flags & kChunkStoreState)) { // the run's last frame leaves its delay line in the state
            const float* lw = u.window_shape ? tab->aac_kbd_long : tab->aac_sine_long;
            for (int i = gt; i < 1024; i += 64) st_out[i] = aac_new_delay(seq, out, lw, sw, psw, i);
        }
    }

    // launch epilogue: the last CTA publishes the new state generation (see mp3_kernel.cu)
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        is_last = atomicAdd(a.done, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (is_last) {
        for (unsigned i = tid; i < gridDim.x; i += blockDim.x)
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
    constexpr size_t smem = (kAacK + 1) * sizeof(AacFrameSmem) + sizeof(AacTabSmem);
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(aac_synth_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    aac_synth_kernel<<<n_chunks, (kAacK + 1) * 64, smem, stream>>>(a);
    return cudaGetLastError();
}

} // namespace symgpu
