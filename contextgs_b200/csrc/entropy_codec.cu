// GPU bitstream codec for the anchor attributes: replaces `encoder_gaussian` /
// `decoder_gaussian` + torchac (utils/encodings.py:83-144, driven chunk by chunk from
// scene/gaussian_model.py:1192-1232 and :1422-1477) and the factorised-prior / binary-mask streams
// (`latent_codec.compress`, `encoder`, `decoder`: gaussian_model.py:1088,1259; encodings.py:147-183).
//
// The reference builds a dense [symbols x alphabet] CDF table on the GPU, copies it to the host and runs
// a sequential C++ arithmetic coder per 1000-anchor chunk.  Here every chunk of every stream is coded by
// its own GPU thread with a byte-wise 32-bit range coder (carry propagation through a cached byte, 16-bit
// cumulative frequencies); the discretised-Gaussian CDF of a symbol is evaluated IN CLOSED FORM from
// (mean, scale, Q) -- two erf per encoded symbol, a binary search of ~log2(alphabet) evaluations per decoded
// symbol -- so no table ever exists.  The container is this library's own (torchac is not in the reference
// tree, its stream format cannot be pinned): parity = encode -> decode returns the quantised tensors bit
// for bit, and the stream length matches the estimated bits.
//
// Cumulative frequency of symbol index i in [0, L] (alphabet = [smin, smax] of the whole (level, attribute)
// stream, found by a parallel pre-pass; L = smax - smin + 1):
//   C(i) = min(rn(Phi(((smin + i) - 0.5) * Q; mean, scale) * (65536 - L)), 65536 - L) + i
// (the "+ i" keeps every symbol codable, as torchac's _convert_to_int_and_normalize does).
#include "entropy_math.cuh"

namespace cgs {
namespace codec {

constexpr uint32_t kTopValue = 1u << 24;
constexpr int kTotalBits = 16;

struct Encoder {
    uint64_t low;
    uint32_t range;
    uint32_t cache_size;
    uint8_t cache;
    uint32_t *out;      // 4-byte aligned
    uint32_t pos, cap;  // bytes written / capacity
    uint32_t word;
    bool overflow;

    __device__ void init(uint32_t *o, uint32_t capacity)
    {
        low = 0; range = 0xffffffffu; cache_size = 1; cache = 0; out = o; pos = 0; cap = capacity; word = 0; overflow = false;
    }
    __device__ __forceinline__ void put(uint8_t b)
    {
        word |= (uint32_t)b << (8 * (pos & 3));
        if ((pos & 3) == 3) {
            if (pos < cap) out[pos >> 2] = word;
            else overflow = true;
            word = 0;
        }
        ++pos;
    }
    __device__ __forceinline__ void shift_low()
    {
        if ((uint32_t)low < 0xff000000u || (low >> 32) != 0) {
            const uint8_t carry = (uint8_t)(low >> 32);
            uint8_t temp = cache;
            do {
                put((uint8_t)(temp + carry));
                temp = 0xff;
            } while (--cache_size);
            cache = (uint8_t)((low >> 24) & 0xff);
        }
        ++cache_size;
        low = (low & 0x00ffffffull) << 8;
    }
    __device__ __forceinline__ void encode(uint32_t lo, uint32_t hi)   // cumulative frequencies out of 2^16
    {
        const uint32_t r = range >> kTotalBits;
        low += (uint64_t)r * lo;
        range = r * (hi - lo);
        while (range < kTopValue) {
            range <<= 8;
            shift_low();
        }
    }
    __device__ uint32_t finish()
    {
        for (int i = 0; i < 5; ++i) shift_low();
        const uint32_t n = pos;
        while (pos & 3) put(0);   // flush the partial word (padding is not counted)
        return n;
    }
};

struct Decoder {
    const uint8_t *in;
    uint32_t pos, len;
    uint32_t code, range;
    __device__ __forceinline__ uint8_t next() { return pos < len ? in[pos++] : (uint8_t)0; }
    __device__ void init(const uint8_t *p, uint32_t n)
    {
        in = p; pos = 0; len = n; code = 0; range = 0xffffffffu;
        for (int i = 0; i < 5; ++i) code = (code << 8) | next();
    }
    __device__ __forceinline__ uint32_t target()
    {
        const uint32_t v = code / (range >> kTotalBits);
        return v > 0xffffu ? 0xffffu : v;
    }
    __device__ __forceinline__ void consume(uint32_t lo, uint32_t hi)
    {
        const uint32_t r = range >> kTotalBits;
        code -= r * lo;
        range = r * (hi - lo);
        while (range < kTopValue) {
            code = (code << 8) | next();
            range <<= 8;
        }
    }
};

// cumulative frequency of the boundary below symbol s (see the header); L = alphabet size
__device__ __forceinline__ uint32_t gauss_cum(int s, int smin, int L, float Q, float mean, float inv_scale)
{
    const float z = ((float)s - 0.5f) * Q;
    const float phi = 0.5f * (1.0f + erff((z - mean) * inv_scale * 0.70710678118654752f));
    const uint32_t M = 65536u - (uint32_t)L;
    uint32_t c = __float2uint_rn(phi * (float)M);
    c = c > M ? M : c;
    return c + (uint32_t)(s - smin);
}

struct GaussStream {
    const int32_t *orig_idx;   // [n_rows] level row -> original anchor
    const float *params;       // [n_rows][176]
    const float *mask;         // [N][10] (offsets stream only)
    int n_rows, chunk_rows, attr, dim, col0;   // attr 0 feat / 1 scaling / 2 offsets; dim 50 / 6 / 30; col0 0 / 50 / 56
};

__device__ __forceinline__ bool coded(const GaussStream &g, int o, int k)
{
    return g.attr != 2 || g.mask[(size_t)o * 10 + k / 3] != 0.0f;
}

// Alphabet of one (level, attribute) stream: min / max of rint(value / Q) over all its coded values.
// One thread per level row, warp-reduced, two atomics per warp.
__global__ void __launch_bounds__(256)
gauss_minmax_kernel(GaussStream g, const float *__restrict__ values, int32_t *__restrict__ minmax)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    int smin = 0x7fffffff, smax = -0x7fffffff;
    if (r < g.n_rows) {
        const int o = g.orig_idx[r];
        const float Q = g.params[(size_t)r * kLdG2 + 172 + g.attr];
        const float *x = values + (size_t)o * g.dim;
        for (int k = 0; k < g.dim; ++k) {
            if (!coded(g, o, k)) continue;
            const int s = (int)rintf(__fdiv_rn(x[k], Q));
            smin = min(smin, s);
            smax = max(smax, s);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        smin = min(smin, __shfl_xor_sync(0xffffffffu, smin, o));
        smax = max(smax, __shfl_xor_sync(0xffffffffu, smax, o));
    }
    if ((threadIdx.x & 31) == 0 && smin <= smax) {
        atomicMin(&minmax[0], smin);
        atomicMax(&minmax[1], smax);
    }
}

// One thread per chunk.  The next symbol's value / mean / scale are fetched before the current symbol is coded:
// the coder's carried state (low, range) is the only true dependency between symbols.
__global__ void __launch_bounds__(64)
gauss_encode_kernel(GaussStream g, const float *__restrict__ values /* [N][dim], quantised */,
                    const int32_t *__restrict__ minmax, uint32_t *__restrict__ out, uint32_t cap_bytes,
                    int32_t *__restrict__ stream_len, int32_t *__restrict__ err)
{
    const int chunk = blockIdx.x * blockDim.x + threadIdx.x;
    const int n_chunks = (g.n_rows + g.chunk_rows - 1) / g.chunk_rows;
    if (chunk >= n_chunks) return;
    const int r0 = chunk * g.chunk_rows, r1 = min(r0 + g.chunk_rows, g.n_rows);
    const int smin = minmax[0], smax = minmax[1];
    const int L = smax - smin + 1;
    if (smax < smin || L > 32768) {   // empty stream / cannot happen after the +-15000-step clamp of STE_multistep
        if (smax >= smin) atomicExch(err, 2);
        stream_len[chunk] = 0;
        return;
    }
    Encoder enc;
    enc.init(out + (size_t)chunk * (cap_bytes / 4), cap_bytes);
    for (int r = r0; r < r1; ++r) {
        const int o = g.orig_idx[r];
        const float *pr = g.params + (size_t)r * kLdG2;
        const float Q = pr[172 + g.attr];
        const float *x = values + (size_t)o * g.dim;
        float xv = x[0], mv = pr[g.col0], sv = pr[kCE + g.col0];
        for (int k = 0; k < g.dim; ++k) {
            const float xc = xv, mean = mv, sc = sv;
            if (k + 1 < g.dim) {
                xv = x[k + 1]; mv = pr[g.col0 + k + 1]; sv = pr[kCE + g.col0 + k + 1];
            }
            if (!coded(g, o, k)) continue;
            const int s = (int)rintf(__fdiv_rn(xc, Q));
            const float inv = __frcp_rn(fmaxf(sc, 1e-9f));
            const uint32_t lo = gauss_cum(s, smin, L, Q, mean, inv), hi = gauss_cum(s + 1, smin, L, Q, mean, inv);
            if (hi <= lo || s < smin || s > smax) {   // erf not monotone at rounding level: the symbol would be undecodable
                atomicExch(err, 1);
                continue;
            }
            enc.encode(lo, hi);
        }
    }
    stream_len[chunk] = (int32_t)enc.finish();
    if (enc.overflow) atomicExch(err, 3);
}

__global__ void __launch_bounds__(64)
gauss_decode_kernel(GaussStream g, const uint8_t *__restrict__ bytes, const int64_t *__restrict__ stream_off,
                    const int32_t *__restrict__ stream_len, const int32_t *__restrict__ minmax,
                    float *__restrict__ values /* [N][dim] out */)
{
    const int chunk = blockIdx.x * blockDim.x + threadIdx.x;
    const int n_chunks = (g.n_rows + g.chunk_rows - 1) / g.chunk_rows;
    if (chunk >= n_chunks) return;
    const int r0 = chunk * g.chunk_rows, r1 = min(r0 + g.chunk_rows, g.n_rows);
    const int smin = minmax[0], smax = minmax[1];
    const int L = smax - smin + 1;
    Decoder dec;
    dec.init(bytes + stream_off[chunk], (uint32_t)stream_len[chunk]);
    for (int r = r0; r < r1; ++r) {
        const int o = g.orig_idx[r];
        const float *pr = g.params + (size_t)r * kLdG2;
        const float Q = pr[172 + g.attr];
        float *x = values + (size_t)o * g.dim;
        for (int k = 0; k < g.dim; ++k) {
            if (!coded(g, o, k)) {
                x[k] = 0.0f;
                continue;
            }
            const float mean = pr[g.col0 + k];
            const float inv = __frcp_rn(fmaxf(pr[kCE + g.col0 + k], 1e-9f));
            const uint32_t v = dec.target();
            // largest s in [smin, smax] with C(s) <= v; the search starts around the predicted mean
            int lo_s = smin, hi_s = smax;
            const int guess = min(max((int)rintf(__fdiv_rn(mean, Q)), smin), smax);
            if (gauss_cum(guess, smin, L, Q, mean, inv) <= v) lo_s = guess; else hi_s = guess - 1;
            while (lo_s < hi_s) {
                const int mid = lo_s + (hi_s - lo_s + 1) / 2;
                if (gauss_cum(mid, smin, L, Q, mean, inv) <= v) lo_s = mid; else hi_s = mid - 1;
            }
            const uint32_t clo = gauss_cum(lo_s, smin, L, Q, mean, inv), chi = gauss_cum(lo_s + 1, smin, L, Q, mean, inv);
            dec.consume(clo, chi > clo ? chi : clo + 1);
            x[k] = (float)lo_s * Q;
        }
    }
}

// ---- static-table streams (hyper latents: one table per channel; offset masks: one table) ------------
// symbols[n][C] int16 (already relative to the table's first symbol); table of channel c = tables[(c % T)][.]
__global__ void __launch_bounds__(64)
table_encode_kernel(const int16_t *__restrict__ symbols, int n_rows, int C, int chunk_rows, const uint32_t *__restrict__ tables,
                    int T, int table_ld, uint32_t *__restrict__ out, uint32_t cap_bytes, int32_t *__restrict__ stream_len,
                    int32_t *__restrict__ err)
{
    const int chunk = blockIdx.x * blockDim.x + threadIdx.x;
    const int n_chunks = (n_rows + chunk_rows - 1) / chunk_rows;
    if (chunk >= n_chunks) return;
    const int r0 = chunk * chunk_rows, r1 = min(r0 + chunk_rows, n_rows);
    Encoder enc;
    enc.init(out + (size_t)chunk * (cap_bytes / 4), cap_bytes);
    for (int r = r0; r < r1; ++r)
        for (int c = 0; c < C; ++c) {
            const int s = symbols[(size_t)r * C + c];
            const uint32_t *tb = tables + (size_t)(c % T) * table_ld;
            if (s < 0 || s + 1 >= table_ld || tb[s + 1] <= tb[s]) {
                atomicExch(err, 1);
                continue;
            }
            enc.encode(tb[s], tb[s + 1]);
        }
    stream_len[chunk] = (int32_t)enc.finish();
    if (enc.overflow) atomicExch(err, 3);
}

__global__ void __launch_bounds__(64)
table_decode_kernel(const uint8_t *__restrict__ bytes, const int64_t *__restrict__ stream_off,
                    const int32_t *__restrict__ stream_len, int n_rows, int C, int chunk_rows,
                    const uint32_t *__restrict__ tables, const int32_t *__restrict__ table_len, int T, int table_ld,
                    int16_t *__restrict__ symbols)
{
    const int chunk = blockIdx.x * blockDim.x + threadIdx.x;
    const int n_chunks = (n_rows + chunk_rows - 1) / chunk_rows;
    if (chunk >= n_chunks) return;
    const int r0 = chunk * chunk_rows, r1 = min(r0 + chunk_rows, n_rows);
    Decoder dec;
    dec.init(bytes + stream_off[chunk], (uint32_t)stream_len[chunk]);
    for (int r = r0; r < r1; ++r)
        for (int c = 0; c < C; ++c) {
            const uint32_t *tb = tables + (size_t)(c % T) * table_ld;
            const int Lsym = table_len[c % T];   // symbols in this table: boundaries tb[0..Lsym]
            const uint32_t v = dec.target();
            int lo = 0, hi = Lsym - 1;
            while (lo < hi) {
                const int mid = lo + (hi - lo + 1) / 2;
                if (tb[mid] <= v) lo = mid; else hi = mid - 1;
            }
            dec.consume(tb[lo], tb[lo + 1]);
            symbols[(size_t)r * C + c] = (int16_t)lo;
        }
}

// streams written at a fixed stride -> one packed byte string (one warp per stream)
__global__ void __launch_bounds__(256)
pack_streams_kernel(const uint32_t *__restrict__ scratch, uint32_t cap_bytes, const int32_t *__restrict__ stream_len,
                    const int64_t *__restrict__ stream_off, int n_streams, uint8_t *__restrict__ packed)
{
    const int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (s >= n_streams) return;
    const uint8_t *src = reinterpret_cast<const uint8_t *>(scratch) + (size_t)s * cap_bytes;
    uint8_t *dst = packed + stream_off[s];
    const int n = stream_len[s];
    for (int i = lane; i < n; i += 32) dst[i] = src[i];
}

}  // namespace codec
}  // namespace cgs

using namespace cgs;

static int attr_layout(int attr, int *dim, int *col0)
{
    if (attr == 0) { *dim = kCF; *col0 = 0; return 0; }
    if (attr == 1) { *dim = kCS; *col0 = kCF; return 0; }
    if (attr == 2) { *dim = kCO; *col0 = kCF + kCS; return 0; }
    set_error("entropy codec: attribute must be 0 (feat), 1 (scaling) or 2 (offsets)");
    return -2;
}

extern "C" int64_t cgs_codec_gauss_stream_capacity(int attr, int chunk_rows)
{
    int dim, col0;
    if (attr_layout(attr, &dim, &col0)) return -1;
    return ((int64_t)2 * dim * chunk_rows + 16 + 3) / 4 * 4;   // <= 16 bits per symbol + flush
}

extern "C" int cgs_codec_gauss_minmax(int attr, const int32_t *orig_idx, int n_rows, const float *params,
                                      const float *mask, const float *values, int32_t *minmax, void *stream)
{
    codec::GaussStream g;
    if (int e = attr_layout(attr, &g.dim, &g.col0)) return e;
    CGS_CHECK_PTR(minmax);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaMemsetAsync(minmax, 0x7f, sizeof(int32_t), st);        // +2139062143
    cudaMemsetAsync(minmax + 1, 0x80, sizeof(int32_t), st);    // -2139062144: an empty stream keeps max < min
    if (n_rows <= 0) return check_launch(__func__);
    CGS_CHECK_PTR(orig_idx); CGS_CHECK_PTR(params); CGS_CHECK_PTR(values);
    if (attr == 2) CGS_CHECK_PTR(mask);
    g.orig_idx = orig_idx; g.params = params; g.mask = mask; g.n_rows = n_rows; g.chunk_rows = 1; g.attr = attr;
    StageScope sc(ST_CODEC, st, 1);
    codec::gauss_minmax_kernel<<<(n_rows + 255) / 256, 256, 0, st>>>(g, values, minmax);
    return check_launch(__func__);
}

extern "C" int cgs_codec_gauss_encode(int attr, const int32_t *orig_idx, int n_rows, int chunk_rows, const float *params,
                                      const float *mask, const float *values, const int32_t *minmax, uint32_t *scratch,
                                      int64_t cap_bytes, int32_t *stream_len, int32_t *err, void *stream)
{
    if (n_rows <= 0) return 0;
    codec::GaussStream g;
    if (int e = attr_layout(attr, &g.dim, &g.col0)) return e;
    CGS_CHECK_PTR(orig_idx); CGS_CHECK_PTR(params); CGS_CHECK_PTR(values); CGS_CHECK_PTR(scratch);
    CGS_CHECK_PTR(stream_len); CGS_CHECK_PTR(minmax); CGS_CHECK_PTR(err);
    if (attr == 2) CGS_CHECK_PTR(mask);
    if (chunk_rows <= 0 || cap_bytes < cgs_codec_gauss_stream_capacity(attr, chunk_rows) || (cap_bytes & 3)) {
        set_error("%s: invalid chunk size / stream capacity", __func__);
        return -2;
    }
    g.orig_idx = orig_idx; g.params = params; g.mask = mask; g.n_rows = n_rows; g.chunk_rows = chunk_rows; g.attr = attr;
    const int n_chunks = (n_rows + chunk_rows - 1) / chunk_rows;
    StageScope sc(ST_CODEC, static_cast<cudaStream_t>(stream), 1);
    codec::gauss_encode_kernel<<<(n_chunks + 63) / 64, 64, 0, static_cast<cudaStream_t>(stream)>>>(
        g, values, minmax, scratch, (uint32_t)cap_bytes, stream_len, err);
    return check_launch(__func__);
}

extern "C" int cgs_codec_gauss_decode(int attr, const int32_t *orig_idx, int n_rows, int chunk_rows, const float *params,
                                      const float *mask, const uint8_t *bytes, const int64_t *stream_off,
                                      const int32_t *stream_len, const int32_t *minmax, float *values, void *stream)
{
    if (n_rows <= 0) return 0;
    codec::GaussStream g;
    if (int e = attr_layout(attr, &g.dim, &g.col0)) return e;
    CGS_CHECK_PTR(orig_idx); CGS_CHECK_PTR(params); CGS_CHECK_PTR(bytes); CGS_CHECK_PTR(stream_off);
    CGS_CHECK_PTR(stream_len); CGS_CHECK_PTR(minmax); CGS_CHECK_PTR(values);
    if (attr == 2) CGS_CHECK_PTR(mask);
    if (chunk_rows <= 0) {
        set_error("%s: invalid chunk size", __func__);
        return -2;
    }
    g.orig_idx = orig_idx; g.params = params; g.mask = mask; g.n_rows = n_rows; g.chunk_rows = chunk_rows; g.attr = attr;
    const int n_chunks = (n_rows + chunk_rows - 1) / chunk_rows;
    StageScope sc(ST_CODEC, static_cast<cudaStream_t>(stream), 1);
    codec::gauss_decode_kernel<<<(n_chunks + 63) / 64, 64, 0, static_cast<cudaStream_t>(stream)>>>(
        g, bytes, stream_off, stream_len, minmax, values);
    return check_launch(__func__);
}

extern "C" int cgs_codec_table_encode(const int16_t *symbols, int n_rows, int C, int chunk_rows, const uint32_t *tables,
                                      int T, int table_ld, uint32_t *scratch, int64_t cap_bytes, int32_t *stream_len,
                                      int32_t *err, void *stream)
{
    if (n_rows <= 0) return 0;
    CGS_CHECK_PTR(symbols); CGS_CHECK_PTR(tables); CGS_CHECK_PTR(scratch); CGS_CHECK_PTR(stream_len); CGS_CHECK_PTR(err);
    if (C <= 0 || T <= 0 || table_ld < 2 || chunk_rows <= 0 || (cap_bytes & 3) ||
        cap_bytes < ((int64_t)2 * C * chunk_rows + 16 + 3) / 4 * 4) {
        set_error("%s: invalid shape / stream capacity", __func__);
        return -2;
    }
    const int n_chunks = (n_rows + chunk_rows - 1) / chunk_rows;
    StageScope sc(ST_CODEC, static_cast<cudaStream_t>(stream), 1);
    codec::table_encode_kernel<<<(n_chunks + 63) / 64, 64, 0, static_cast<cudaStream_t>(stream)>>>(
        symbols, n_rows, C, chunk_rows, tables, T, table_ld, scratch, (uint32_t)cap_bytes, stream_len, err);
    return check_launch(__func__);
}

extern "C" int cgs_codec_table_decode(const uint8_t *bytes, const int64_t *stream_off, const int32_t *stream_len,
                                      int n_rows, int C, int chunk_rows, const uint32_t *tables, const int32_t *table_len,
                                      int T, int table_ld, int16_t *symbols, void *stream)
{
    if (n_rows <= 0) return 0;
    CGS_CHECK_PTR(bytes); CGS_CHECK_PTR(stream_off); CGS_CHECK_PTR(stream_len); CGS_CHECK_PTR(tables);
    CGS_CHECK_PTR(table_len); CGS_CHECK_PTR(symbols);
    if (C <= 0 || T <= 0 || table_ld < 2 || chunk_rows <= 0) {
        set_error("%s: invalid shape", __func__);
        return -2;
    }
    const int n_chunks = (n_rows + chunk_rows - 1) / chunk_rows;
    StageScope sc(ST_CODEC, static_cast<cudaStream_t>(stream), 1);
    codec::table_decode_kernel<<<(n_chunks + 63) / 64, 64, 0, static_cast<cudaStream_t>(stream)>>>(
        bytes, stream_off, stream_len, n_rows, C, chunk_rows, tables, table_len, T, table_ld, symbols);
    return check_launch(__func__);
}

extern "C" int cgs_codec_pack_streams(const uint32_t *scratch, int64_t cap_bytes, const int32_t *stream_len,
                                      const int64_t *stream_off, int n_streams, uint8_t *packed, void *stream)
{
    if (n_streams <= 0) return 0;
    CGS_CHECK_PTR(scratch); CGS_CHECK_PTR(stream_len); CGS_CHECK_PTR(stream_off); CGS_CHECK_PTR(packed);
    StageScope sc(ST_CODEC, static_cast<cudaStream_t>(stream), 1);
    codec::pack_streams_kernel<<<(n_streams * 32 + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        scratch, (uint32_t)cap_bytes, stream_len, stream_off, n_streams, packed);
    return check_launch(__func__);
}
