// Power-of-two IMDCT shared by the AAC and Vorbis kernels: the GPU form of symphonia-core's
// `Imdct::imdct` (dsp/mdct.rs:67-146) over its in-tree radix-2 FFT (dsp/fft/no_simd.rs).
//
// The reference runs, per block: pre-twiddle -> bit reversal -> fft2/4/8/16/32 base cases ->
// breadth-first merges -> post-twiddle.  Here a group of threads owns one block (or a batch of equal
// blocks) in shared memory and executes the SAME butterfly network, three levels per pass with the
// eight points of a thread held in registers between levels.  Every butterfly performs exactly the
// reference's operations on the reference's operands:
//   level size <= 32: k = 0 -> q = o; k = N/4 -> q = (o.im, -o.re); k = N/8, 3N/8 -> the 1/sqrt(2)
//     forms of no_simd.rs:302-305/312/320; any other k -> literal * o  (num-complex Mul);
//   level size >= 64: q = o * W[k] from the f64-built tables (no_simd.rs:16-36, merge :222-238).
// Lanes that need different rules compute the candidate forms and SELECT (no divergent branches);
// the selected value is bit-identical to what the reference's specialised code computes.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace symgpu {

constexpr float kFrac1Sqrt2F = 0.707106781186547524400844362104849039f;

// Device-side tables (global memory, built on the host in tables.cpp).
struct FftTables {
    float2 lit16[8];       // (cos(pi k/8), -sin(pi k/8)), k = 0..7   -- level-16 literals
    float2 lit32[16];      // (cos(pi k/16), -sin(pi k/16)), k = 0..15 -- level-32 literals
    float2 merge[2048 - 32]; // level size s >= 64: W_s[k] at merge[s/2 - 32 + k], s = 64 .. 2048
};

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
// num-complex: (a+bi)(c+di) = (ac - bd) + (ad + bc)i
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// Shared-memory index of logical complex element p: one pad element per 8 keeps the stride-8
// accesses of the first pass on distinct banks.
__device__ __forceinline__ int zpad(int p) { return p + (p >> 3); }
__host__ __device__ constexpr int zpad_len(int n) { return n + (n >> 3) + 1; }

// q = o "times" the level-`size` twiddle k, with the reference's rules.  `size` is a compile-time
// constant at every call site; k is per-thread.
template <int SIZE>
__device__ __forceinline__ float2 twiddled(float2 o, int k, const FftTables* __restrict__ ft) {
    if constexpr (SIZE == 2) {
        return o;
    } else if constexpr (SIZE == 4) {
        return k == 0 ? o : make_float2(o.y, -o.x);
    } else if constexpr (SIZE <= 32) {
        constexpr int half = SIZE / 2;
        const float a = kFrac1Sqrt2F * o.x, b = kFrac1Sqrt2F * o.y;      // k = N/8
        const float na = -kFrac1Sqrt2F * o.x, nb = -kFrac1Sqrt2F * o.y;  // k = 3N/8
        float2 q;
        if constexpr (SIZE == 8) {
            q = o; // every k of an 8-point level is special; placeholder for k = 0
        } else {
            const float2 w = SIZE == 16 ? ft->lit16[k] : ft->lit32[k];
            q = cmul(w, o); // literal on the left (no_simd.rs:309 ...)
        }
        if (k == 0) q = o;
        if (4 * k == half) q = make_float2(a + b, b - a);
        if (2 * k == half) q = make_float2(o.y, -o.x);
        if (4 * k == 3 * half) q = make_float2(na - nb, na + nb);
        return q;
    } else {
        return cmul(o, ft->merge[SIZE / 2 - 32 + k]); // merge(): o * w
    }
}

// NL butterfly levels on the R = 2^NL register points of one thread.  The points are logical
// elements base + j * stride (stride = 2^(L0-1) = half of the first level's size).
template <int L0, int NL>
__device__ __forceinline__ void levels_in_registers(float2 (&v)[1 << NL], int t_in_stride, const FftTables* __restrict__ ft) {
    constexpr int R = 1 << NL;
    constexpr int stride = 1 << (L0 - 1);
#pragma unroll
    for (int l = 0; l < NL; ++l) {
        const int span = 1 << l; // register distance of the butterfly partners
#pragma unroll
        for (int j = 0; j < R; ++j) {
            if ((j & span) == 0) {
                // twiddle index within the half: (position of element j) mod (stride * span)
                const int k = t_in_stride + stride * (j & (span - 1));
                float2 q;
                // level size = 2 * stride * span
                if (l == 0) q = twiddled<2 * stride>(v[j + span], k, ft);
                else if (l == 1) q = twiddled<4 * stride>(v[j + span], k, ft);
                else q = twiddled<8 * stride>(v[j + span], k, ft);
                const float2 p = v[j];
                v[j] = cadd(p, q);
                v[j + span] = csub(p, q);
            }
        }
    }
}

// One pass over `n_pts` logical points (possibly several independent blocks laid end to end, each a
// multiple of 2^(L0+NL-1) long): levels L0 .. L0+NL-1.
template <int L0, int NL>
__device__ __forceinline__ void fft_pass(float2* z, int n_pts, int tid, int n_threads, const FftTables* __restrict__ ft) {
    constexpr int R = 1 << NL;
    constexpr int stride = 1 << (L0 - 1);
    for (int t = tid; t < n_pts / R; t += n_threads) {
        const int t_in = t & (stride - 1);
        const int base = (t >> (L0 - 1)) * (stride * R) + t_in;
        float2 v[R];
#pragma unroll
        for (int j = 0; j < R; ++j) v[j] = z[zpad(base + j * stride)];
        levels_in_registers<L0, NL>(v, t_in, ft);
#pragma unroll
        for (int j = 0; j < R; ++j) z[zpad(base + j * stride)] = v[j];
    }
}

// Barrier policies for the thread group that owns a block: the whole CTA, or a named barrier shared by
// `N` threads (several groups of one CTA each running their own IMDCT).
struct CtaSync {
    __device__ __forceinline__ void operator()() const { __syncthreads(); }
};
struct WarpSync { // a block owned by ONE warp: no hardware barrier at all
    __device__ __forceinline__ void operator()() const { __syncwarp(); }
};
struct NamedSync {
    int id, n;
    __device__ __forceinline__ void operator()() const { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
};

// Forward FFT of `batch` independent blocks of 2^LOG2 complex points stored consecutively in z
// (already in bit-reversed order).  Every thread of the owning group takes part.
template <int LOG2, typename Sync>
__device__ __forceinline__ void fft_levels(float2* z, int batch, int tid, int n_threads, const FftTables* __restrict__ ft,
                                           Sync sync) {
    const int n_pts = batch << LOG2;
    static_assert(LOG2 >= 4 && LOG2 <= 11, "FFT sizes 16..2048");
    fft_pass<1, 3>(z, n_pts, tid, n_threads, ft);
    sync();
    if constexpr (LOG2 == 4) {
        fft_pass<4, 1>(z, n_pts, tid, n_threads, ft);
    } else if constexpr (LOG2 == 5) {
        fft_pass<4, 2>(z, n_pts, tid, n_threads, ft);
    } else {
        fft_pass<4, 3>(z, n_pts, tid, n_threads, ft);
        if constexpr (LOG2 > 6) {
            sync();
            if constexpr (LOG2 == 7) fft_pass<7, 1>(z, n_pts, tid, n_threads, ft);
            else if constexpr (LOG2 == 8) fft_pass<7, 2>(z, n_pts, tid, n_threads, ft);
            else {
                fft_pass<7, 3>(z, n_pts, tid, n_threads, ft);
                if constexpr (LOG2 > 9) {
                    sync();
                    if constexpr (LOG2 == 10) fft_pass<10, 1>(z, n_pts, tid, n_threads, ft);
                    else fft_pass<10, 2>(z, n_pts, tid, n_threads, ft);
                }
            }
        }
    }
    sync();
}

// IMDCT of `batch` blocks: spec [batch][N] (shared) -> out [batch][2N] (shared), N = 2^(LOG2+1)
// spectral lines per block, FFT size n2 = 2^LOG2.  tw = Imdct.twiddle (n2 complex).  z is scratch of
// zpad_len(batch * n2) complex.  All `n_threads` threads of the group must call this.
template <int LOG2, typename Sync = CtaSync>
__device__ __forceinline__ void imdct_blocks(const float* spec, float* out, float2* z, int batch,
                                             const float2* __restrict__ tw, const FftTables* __restrict__ ft, int tid,
                                             int n_threads, Sync sync = Sync()) {
    constexpr int n2 = 1 << LOG2, n = 2 * n2, n4 = n2 / 2;
    // Pre-twiddle (mdct.rs:81-88) fused with the bit-reversal permutation (no_simd.rs:101-107).
    for (int e = tid; e < batch * n2; e += n_threads) {
        const int b = e >> LOG2, i = e & (n2 - 1);
        const float* s = spec + b * n;
        const float even = s[2 * i];
        const float odd = -s[n - 1 - 2 * i];
        const float2 w = tw[i];
        const float re = odd * w.y - even * w.x;
        const float im = odd * w.x + even * w.y;
        const int r = (int)(__brev((unsigned)i) >> (32 - LOG2));
        z[zpad((b << LOG2) + r)] = make_float2(re, im);
    }
    sync();
    fft_levels<LOG2>(z, batch, tid, n_threads, ft, sync);
    // Post-twiddle (mdct.rs:100-137): val = w * conj(x), scattered into the four quarters.
    for (int e = tid; e < batch * n2; e += n_threads) {
        const int b = e >> LOG2, k = e & (n2 - 1);
        const float2 x = z[zpad(e)];
        const float2 w = tw[k];
        const float2 val = cmul(w, make_float2(x.x, -x.y));
        float* o = out + b * 2 * n;
        if (k < n4) {
            const int fi = 2 * k, ri = n2 - 1 - 2 * k;
            o[ri] = -val.y;
            o[n2 + fi] = val.y;
            o[2 * n2 + ri] = val.x;
            o[3 * n2 + fi] = val.x;
        } else {
            const int i = k - n4;
            const int fi = 2 * i, ri = n2 - 1 - 2 * i;
            o[fi] = -val.x;
            o[n2 + ri] = val.x;
            o[2 * n2 + fi] = val.y;
            o[3 * n2 + ri] = val.y;
        }
    }
    sync();
}

// ---- IMDCT without an output array ---------------------------------------------------------------------------------------
// The 2N outputs of the post-twiddle (mdct.rs:100-137) are the N numbers val[k].re / val[k].im (k < N/2), each written twice,
// once negated in the first quarter: out is a permutation (with signs) of z.  `imdct_to_z` therefore stops at z[k] = val[k]
// (in place: element k is read and written by the same thread) and `imdct_out` returns out[j] from it -- 4.6 KB of shared
// memory per 1024-line block instead of 12.8 KB, a quarter of the shared-memory stores, and no staging of the spectrum: the
// pre-twiddle reads it from where it lies (global memory), two mirrored float2 per pair of FFT inputs, every float used.
//
// `pair(b, l)` returns the spectral lines (l, l + 1) of block b (l even): a plain load, or whatever produces the spectrum
// (the Vorbis kernel multiplies floor and residue right here).  z: zpad_len(batch * 2^LOG2) float2 of shared memory.
// UNROLL: pre-twiddle iterations whose loads are in flight together (2 when the pair functor is heavy, more for a plain load).
template <int LOG2, int UNROLL = 2, typename Pair, typename Sync>
__device__ __forceinline__ void imdct_to_z_from(Pair pair, float2* z, int batch, const float2* __restrict__ tw,
                                                const FftTables* __restrict__ ft, int tid, int n_threads, Sync sync) {
    constexpr int n2 = 1 << LOG2, n = 2 * n2, n4 = n2 / 2;
    // FFT input i needs spec[2i] and spec[n-1-2i]; input n2-1-i needs spec[n-2-2i] and spec[2i+1]: the same two pairs.
#pragma unroll UNROLL
    for (int e = tid; e < batch * n4; e += n_threads) {
        const int b = e >> (LOG2 - 1), i = e & (n4 - 1), i2 = n2 - 1 - i;
        const float2 lo = pair(b, 2 * i);
        const float2 hi = pair(b, n - 2 - 2 * i);
        {
            const float even = lo.x, odd = -hi.y;
            const float2 w = tw[i];
            const int r = (int)(__brev((unsigned)i) >> (32 - LOG2));
            z[zpad((b << LOG2) + r)] = make_float2(odd * w.y - even * w.x, odd * w.x + even * w.y);
        }
        {
            const float even = hi.x, odd = -lo.y;
            const float2 w = tw[i2];
            const int r = (int)(__brev((unsigned)i2) >> (32 - LOG2));
            z[zpad((b << LOG2) + r)] = make_float2(odd * w.y - even * w.x, odd * w.x + even * w.y);
        }
    }
    sync();
    fft_levels<LOG2>(z, batch, tid, n_threads, ft, sync);
    for (int e = tid; e < batch * n2; e += n_threads) {
        const float2 x = z[zpad(e)];
        z[zpad(e)] = cmul(tw[e & (n2 - 1)], make_float2(x.x, -x.y));
    }
    sync();
}

// spec: [batch][N] floats in global memory (8-byte aligned), N = 2^(LOG2+1).
template <int LOG2, typename Sync>
__device__ __forceinline__ void imdct_to_z(const float* __restrict__ spec, float2* z, int batch, const float2* __restrict__ tw,
                                           const FftTables* __restrict__ ft, int tid, int n_threads, Sync sync) {
    auto pair = [spec](int b, int l) { return __ldg(reinterpret_cast<const float2*>(spec + (b << (LOG2 + 1)) + l)); };
    imdct_to_z_from<LOG2>(pair, z, batch, tw, ft, tid, n_threads, sync);
}

// out[j] (0 <= j < 4 * n2) of block `b` after imdct_to_z, bit for bit what imdct_blocks stores.
template <int LOG2>
__device__ __forceinline__ float imdct_out(const float2* z, int b, int j) {
    constexpr int n2 = 1 << LOG2, n4 = n2 / 2;
    const int q = j >> LOG2, r = j & (n2 - 1);
    const bool odd = r & 1;
    // quarters 0 and 2 take k < n4 from odd r (mirrored), quarters 1 and 3 from even r
    const bool low = (q & 1) ? !odd : odd;
    const int half = odd ? (n2 - 1 - r) >> 1 : r >> 1;
    const int k = low ? half : n4 + half;
    const float2 v = z[zpad((b << LOG2) + k)];
    // low: quarters 0, 1 -> im, quarters 2, 3 -> re; high: quarters 0, 1 -> re, quarters 2, 3 -> im
    const float val = ((q >> 1) ^ (low ? 0 : 1)) ? v.x : v.y;
    return q == 0 ? -val : val;
}

// Window + overlap-add of two equally sized blocks straight from their z arrays:
//   dst[j] = prev_out[2*n2 + j] * fall[len - 1 - j] + cur_out[j] * rise[j],   0 <= j < len = 2 * n2
// (the reference's `overlap * win[len-1-j] + imdct[j] * win[j]`, dsp.rs:96-101; aac/dsp.rs:104-108 is the same expression).
// Iteration i (0 <= i < n2 / 2) produces the adjacent outputs (2i, 2i+1) and (n2+2i, n2+2i+1): by the post-twiddle's scatter
// (see imdct_out) they are, for the current block, (-z[n4+i].re, -z[n4-1-i].im) and (z[i].im, z[n2-1-i].re), and for the second
// half of the previous block (z[n4+i].im, z[n4-1-i].re) and (z[i].re, z[n2-1-i].im) -- four 8-byte loads per block for four
// outputs, no index arithmetic per sample, 8-byte stores.  fall / rise: `len` floats each, 8-byte aligned.
template <typename WinLoad>
__device__ __forceinline__ void overlap_add_equal(const float2* zc, const float2* zp, int log2, WinLoad win2, float* dst, int tid,
                                                  int n_threads) {
    const int n2 = 1 << log2, n4 = n2 >> 1, len = 2 * n2;
#pragma unroll 2
    for (int i = tid; i < n4; i += n_threads) {
        const float2 ca = zc[zpad(n4 + i)], cb = zc[zpad(n4 - 1 - i)], cc = zc[zpad(i)], cd = zc[zpad(n2 - 1 - i)];
        const float2 pa = zp[zpad(n4 + i)], pb = zp[zpad(n4 - 1 - i)], pc = zp[zpad(i)], pd = zp[zpad(n2 - 1 - i)];
        {
            const int j = 2 * i;
            const float2 r = win2(false, j), f = win2(true, len - 2 - j); // f = (fall for j + 1, fall for j)
            float2 y;
            y.x = pa.y * f.y + (-ca.x) * r.x;
            y.y = pb.x * f.x + (-cb.y) * r.y;
            *reinterpret_cast<float2*>(dst + j) = y;
        }
        {
            const int j = n2 + 2 * i;
            const float2 r = win2(false, j), f = win2(true, len - 2 - j);
            float2 y;
            y.x = pc.x * f.y + cc.y * r.x;
            y.y = pd.y * f.x + cd.x * r.y;
            *reinterpret_cast<float2*>(dst + j) = y;
        }
    }
}

// The same for a block size known only at run time (one block: b = 0).
__device__ __forceinline__ float imdct_out_rt(const float2* z, int log2, int j) {
    const int n2 = 1 << log2, n4 = n2 >> 1;
    const int q = j >> log2, r = j & (n2 - 1);
    const bool odd = r & 1;
    const bool low = (q & 1) ? !odd : odd;
    const int half = odd ? (n2 - 1 - r) >> 1 : r >> 1;
    const float2 v = z[zpad(low ? half : n4 + half)];
    const float val = ((q >> 1) ^ (low ? 0 : 1)) ? v.x : v.y;
    return q == 0 ? -val : val;
}

} // namespace symgpu
