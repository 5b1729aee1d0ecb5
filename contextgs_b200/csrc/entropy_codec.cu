// GPU bitstream codec for the anchor attributes: replaces `encoder_gaussian` /
// `decoder_gaussian` + torchac (utils/encodings.py:83-144, driven chunk by chunk from
// scene/gaussian_model.py:1192-1232 and :1422-1477) and the factorised-prior / binary-mask streams
// (`latent_codec.compress`, `encoder`, `decoder`: gaussian_model.py:1088,1259; encodings.py:147-183).
//
// The reference builds a dense [symbols x alphabet] CDF table on the GPU, copies it to the host and runs
// a sequential C++ arithmetic coder per 1000-anchor chunk.  Here every chunk of every stream is coded by
// its own GPU thread with a byte-wise 32-bit range coder (carry propagation through a cached byte, 16-bit
// cumulative frequencies); the discretised-Gaussian CDF of a symbol is evaluated IN CLOSED FORM from
// (mean, scale, Q), so no table ever exists.  Encoding is two passes per level: a fully parallel pass writes the
// coding interval of every value (two erf each, coalesced), then one thread per chunk runs only the carried
// (low, range) state over its intervals.  Decoding inverts the Gaussian CDF at the coder's target to land on the
// symbol directly and confirms it with ~2 CDF evaluations (bisection only in the tails).  The container is this library's own (torchac is not in the reference
// tree, its stream format cannot be pinned): parity = encode -> decode returns the quantised tensors bit
// for bit, and the stream length matches the estimated bits.
//
// Cumulative frequency of symbol index i in [0, L] (alphabet = [smin, smax] of the whole (level, attribute)
// stream, found while the level is quantised; L = smax - smin + 1):
//   C(i) = min(rn(Phi((((smin + i) - 0.5) * Q - mean) / scale) * (65536 - L)), 65536 - L) + i
// (the "+ i" keeps every symbol codable, as torchac's _convert_to_int_and_normalize does).  Phi is the coder's own
// tabulated normal CDF (phi_interp below, |error| < 2e-7).
#include <algorithm>

#include "entropy_math.cuh"

namespace cgs {
namespace codec {

constexpr uint32_t kTopValue = 1u << 24;
constexpr int kTotalBits = 16;

// Encoder.  The byte string is the classic carry-propagating range coder's digits of the final code value -- without the
// always-zero leading byte, and terminated by ONE byte: the coder ends on V = low rounded up to a multiple of 2^24 (inside
// [low, low + range) because range >= 2^24) and V's three low bytes are zeros the decoder supplies itself.  The same bytes
// come out of oracle/codec_ref.py's cached-byte formulation.  Here the 0 / 1 / 2 bytes a symbol shifts out of `low` are appended to a 64-bit register of pending bytes
// without a loop, and stored 32 bits at a time; a carry out of `low` is an increment of that register and only ripples into
// memory when every pending byte is 0xff.
struct Encoder {
    uint32_t low, range;
    uint32_t *out;             // 4-byte aligned
    uint32_t pos, stored, cap; // bytes emitted / of those, stored to memory (multiple of 4) / capacity (multiple of 4)
    uint64_t acc;              // the pos - stored (<= 3 between symbols) pending bytes, first emitted most significant
    bool overflow;

    __device__ void init(uint32_t *o, uint32_t capacity)
    {
        low = 0; range = 0xffffffffu; out = o; pos = 0; stored = 0; cap = capacity; acc = 0; overflow = false;
    }
    __device__ __forceinline__ void append(uint32_t bytes, uint32_t nb)   // nb <= 2
    {
        acc = (acc << (8 * nb)) | bytes;
        pos += nb;
        const uint32_t cnt = pos - stored;
        if (cnt >= 4u) {
            const uint32_t w = (uint32_t)(acc >> (8 * (cnt - 4u)));
            if (stored + 4u <= cap) out[stored >> 2] = __byte_perm(w, 0, 0x0123);   // memory order = emission order
            else overflow = true;
            stored += 4u;
        }
    }
    __device__ __forceinline__ void carry()
    {
        const uint32_t cnt = pos - stored;
        const uint64_t mask = (1ull << (8 * cnt)) - 1ull;
        if ((acc & mask) != mask) { acc += 1ull; return; }
        acc &= ~mask;
        for (int i = (int)(stored >> 2) - 1; i >= 0; --i) {
            if ((uint32_t)i >= cap / 4) continue;
            const uint32_t w = __byte_perm(out[i], 0, 0x0123) + 1u;
            out[i] = __byte_perm(w, 0, 0x0123);
            if (w != 0u) break;
        }
    }
    __device__ __forceinline__ void encode(uint32_t lo, uint32_t hi)   // cumulative frequencies out of 2^16
    {
        const uint32_t r = range >> kTotalBits;
        const uint32_t t = low + r * lo;
        if (t < low) carry();
        low = t;
        range = r * (hi - lo);                                             // >= 2^8: at most two bytes leave
        const uint32_t nb = range < kTopValue ? (range < (1u << 16) ? 2u : 1u) : 0u;
        append(__funnelshift_l(low, 0u, 8 * nb), nb);                      // the top nb bytes of low (0 for nb = 0)
        low <<= 8 * nb;
        range <<= 8 * nb;
    }
    __device__ uint32_t finish()
    {
        const uint32_t t = low + 0x00ffffffu;   // round up to a multiple of 2^24
        if (t < low) carry();
        append(t >> 24, 1);
        const uint32_t n = pos, cnt = pos - stored;
        if (cnt) {   // flush the partial word (padding is not counted)
            if (stored + 4u <= cap) out[stored >> 2] = __byte_perm((uint32_t)(acc << (8 * (4u - cnt))), 0, 0x0123);
            else overflow = true;
        }
        return n;
    }
};

// Decoder.  The chunk's bytes are read through 32-bit words (the aligned word holding the current byte and the next one,
// fetched a word ahead of its use): one global load per four bytes, off the coder's critical path.  No byte past the end
// of the chunk is touched (the word holding its last bytes is assembled from them); bytes past the chunk read as zero.
struct Decoder {
    const uint32_t *words;   // aligned word that holds byte 0 of the chunk
    uint32_t bpos, end;      // byte position / end of the chunk, both counted from words[0]
    uint32_t cur, nxt;       // words[bpos / 4] and words[bpos / 4 + 1]
    uint32_t code, range;
    __device__ __forceinline__ uint32_t fetch(uint32_t w) const
    {
        const uint32_t lo = 4u * w;
        if (lo + 4u <= end) return __ldg(words + w);
        if (lo >= end) return 0u;
        // the word that holds the chunk's last bytes: the byte string may END inside it, so only its own bytes are read
        const uint8_t *p = reinterpret_cast<const uint8_t *>(words + w);
        uint32_t v = 0;
        for (uint32_t i = 0; lo + i < end; ++i) v |= (uint32_t)p[i] << (8u * i);
        return v;
    }
    __device__ __forceinline__ uint32_t next()
    {
        const uint32_t b = bpos < end ? (cur >> (8u * (bpos & 3u))) & 0xffu : 0u;
        ++bpos;
        if ((bpos & 3u) == 0u) {
            cur = nxt;
            nxt = fetch((bpos >> 2) + 1u);
        }
        return b;
    }
    __device__ void init(const uint8_t *p, uint32_t n)
    {
        const uintptr_t a = reinterpret_cast<uintptr_t>(p);
        words = reinterpret_cast<const uint32_t *>(a & ~(uintptr_t)3);
        bpos = (uint32_t)(a & 3u);
        end = bpos + n;
        cur = fetch(0);
        nxt = fetch(1);
        code = 0; range = 0xffffffffu;
        for (int i = 0; i < 4; ++i) code = (code << 8) | next();
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

// Phi of the coder: the standard normal CDF tabulated at kPhiN + 1 points of [kPhiZ0, -kPhiZ0] (host, fp64 erfc, rounded to
// fp32; T[0] = 0, T[kPhiN] = 1) and interpolated linearly: |error| < 2e-7, far below the 2^-16 resolution of the coder,
// monotone, and an order of magnitude cheaper than erff inside the sequential decoder.  Every operation is a single
// correctly rounded fp32 operation in a fixed order (no contraction), so the function is reproducible outside this file
// (tests/test_codec_gpu.py rebuilds the dense CDF tables with torch ops from cgs_codec_phi_table).
constexpr int kPhiN = 4096;
constexpr float kPhiZ0 = -4.75f;
constexpr float kPhiInvH = (float)kPhiN / 9.5f;
__device__ float g_phi_table[kPhiN + 1];

__device__ __forceinline__ void load_phi_table(float *__restrict__ T)
{
    for (int i = threadIdx.x; i <= kPhiN; i += blockDim.x) T[i] = g_phi_table[i];
    __syncthreads();
}

__device__ __forceinline__ float phi_interp(const float *__restrict__ T, float z)
{
    float t = __fmul_rn(__fsub_rn(z, kPhiZ0), kPhiInvH);
    t = fminf(fmaxf(t, 0.0f), (float)kPhiN);   // NaN -> 0
    const int j = min((int)t, kPhiN - 1);
    const float f = __fsub_rn(t, (float)j);
    const float t0 = T[j];
    return __fadd_rn(t0, __fmul_rn(f, __fsub_rn(T[j + 1], t0)));
}

// cumulative frequency of the boundary below symbol s (see the header); M = 65536 - alphabet size
__device__ __forceinline__ uint32_t gauss_cum(const float *__restrict__ T, int s, int smin, uint32_t M, float Q, float mean,
                                              float inv_scale)
{
    const float z = __fmul_rn((float)s - 0.5f, Q);
    const float phi = phi_interp(T, __fmul_rn(__fsub_rn(z, mean), inv_scale));
    uint32_t c = __float2uint_rn(__fmul_rn(phi, (float)M));
    c = c > M ? M : c;
    return c + (uint32_t)(s - smin);
}

// All three Gaussian streams (feat / scaling / offsets) of ONE level; attr 0 / 1 / 2, dim 50 / 6 / 30, first params
// column 0 / 50 / 56.  Chunk ids run over the three streams back to back: [0, n_chunks[0]) feat, then scaling, then offsets.
struct LevelStreams {
    const int32_t *orig_idx;   // [n_rows] level row -> original anchor
    const float *params;       // [n_rows][176]
    const float *mask;         // [N][10] (offsets stream only)
    float *values[3];          // [N][dim] quantised values (read by the encoder, written by the decoder)
    int n_rows;
    int rows[3];               // level rows per chunk
    int n_chunks[3];
    uint32_t cap[3];           // bytes reserved per chunk in the encoder's scratch
    int64_t region[3];         // first scratch WORD of each stream
};

constexpr uint32_t kSkip = 0xffffffffu;   // interval slot of a value that is not coded (lo 65535 + width 65536 cannot occur)

__device__ __forceinline__ int attr_of_col(int col) { return col < kCF ? 0 : col < kCF + kCS ? 1 : 2; }
__device__ __forceinline__ int attr_dim(int a) { return a == 0 ? kCF : a == 1 ? kCS : kCO; }
__device__ __forceinline__ int attr_col0(int a) { return a == 0 ? 0 : a == 1 ? kCF : kCF + kCS; }
template <typename T> __device__ __forceinline__ T pick3(const T (&v)[3], int a) { return a == 0 ? v[0] : a == 1 ? v[1] : v[2]; }

// chunk id over the three streams -> (attr, chunk of that stream)
__device__ __forceinline__ bool locate_chunk(const LevelStreams &g, int c, int &attr, int &cl)
{
    if (c < g.n_chunks[0]) { attr = 0; cl = c; return true; }
    c -= g.n_chunks[0];
    if (c < g.n_chunks[1]) { attr = 1; cl = c; return true; }
    c -= g.n_chunks[1];
    attr = 2; cl = c;
    return c < g.n_chunks[2];
}

__global__ void minmax_init_kernel(int32_t *minmax)
{
    if (threadIdx.x < 6) minmax[threadIdx.x] = (threadIdx.x & 1) ? -2139062144 : 2139062143;   // empty stream: max < min
}

// Alphabets of the three streams of a level: min / max of rint(value / Q) over the coded values.  One warp per level
// row (lane -> columns lane, lane + 32, lane + 64 of the 86 coded values: the row's loads are coalesced and its anchor
// index / Q are fetched once), redux per warp, shared atomics per block, six global atomics per block.
constexpr int kRowsPerWarp = 4;

__global__ void __launch_bounds__(256)
gauss_level_minmax_kernel(LevelStreams g, int32_t *__restrict__ minmax)
{
    __shared__ int32_t sm[6];
    if (threadIdx.x < 6) sm[threadIdx.x] = (threadIdx.x & 1) ? -0x7fffffff : 0x7fffffff;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int row0 = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * kRowsPerWarp;
    int lo[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff}, hi[3] = {-0x7fffffff, -0x7fffffff, -0x7fffffff};
#pragma unroll
    for (int rr = 0; rr < kRowsPerWarp; ++rr) {
        const int row = row0 + rr;
        if (row >= g.n_rows) break;
        const int o = g.orig_idx[row];
        const float *pr = g.params + (size_t)row * kLdG2;
#pragma unroll
        for (int j = 0; j < 3; ++j) {   // j is also the stream of most of these columns; the exact one is attr
            const int col = lane + 32 * j;
            if (col >= kCE) continue;
            const int attr = attr_of_col(col), k = col - attr_col0(attr);
            if (attr == 2 && g.mask[(size_t)o * 10 + k / 3] == 0.0f) continue;
            const int s = (int)rintf(__fdiv_rn(pick3(g.values, attr)[(size_t)o * attr_dim(attr) + k], pr[172 + attr]));
            if (attr == 0) { lo[0] = min(lo[0], s); hi[0] = max(hi[0], s); }
            else if (attr == 1) { lo[1] = min(lo[1], s); hi[1] = max(hi[1], s); }
            else { lo[2] = min(lo[2], s); hi[2] = max(hi[2], s); }
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const int l = __reduce_min_sync(0xffffffffu, lo[a]), h = __reduce_max_sync(0xffffffffu, hi[a]);
        if (lane == 0 && l <= h) {
            atomicMin(&sm[2 * a], l);
            atomicMax(&sm[2 * a + 1], h);
        }
    }
    __syncthreads();
    if (threadIdx.x < 6 && sm[threadIdx.x & ~1] <= sm[threadIdx.x | 1]) {
        if (threadIdx.x & 1) atomicMax(&minmax[threadIdx.x], sm[threadIdx.x]);
        else atomicMin(&minmax[threadIdx.x], sm[threadIdx.x]);
    }
}

// Encoder pass 1: the coding interval [C(s), C(s + 1)) of every coded value, packed as lo | (hi - lo - 1) << 16 into
// iv[n_rows * col0 + row * dim + k] (kSkip for values that are not coded), so the CDF evaluations and the division are out
// of the sequential coder and all loads are coalesced.  One warp per level row, in four passes with compile-time stream
// (feat 0..31, feat 32..49, scaling, offsets): no per-value stream selection, the row's anchor index / steps loaded once.
struct CodedValue {
    float x, mean, scale;
    bool coded;
};

template <int ATTR>
__device__ __forceinline__ CodedValue load_value(const LevelStreams &g, const float *__restrict__ pr, int o, int k, bool active)
{
    constexpr int dim = ATTR == 0 ? kCF : ATTR == 1 ? kCS : kCO;
    constexpr int col0 = ATTR == 0 ? 0 : ATTR == 1 ? kCF : kCF + kCS;
    CodedValue v;
    v.coded = active && (ATTR != 2 || g.mask[(size_t)o * 10 + k / 3] != 0.0f);
    v.x = active ? g.values[ATTR][(size_t)o * dim + k] : 0.0f;
    v.mean = active ? pr[col0 + k] : 0.0f;
    v.scale = active ? pr[kCE + col0 + k] : 1.0f;
    return v;
}

template <int ATTR>
__device__ __forceinline__ void interval_of_value(const LevelStreams &g, const float *__restrict__ T, const CodedValue &v, float Q,
                                                  int row, int k, bool active, int smin, int smax, uint32_t *__restrict__ iv,
                                                  int32_t *__restrict__ err)
{
    constexpr int dim = ATTR == 0 ? kCF : ATTR == 1 ? kCS : kCO;
    constexpr int col0 = ATTR == 0 ? 0 : ATTR == 1 ? kCF : kCF + kCS;
    if (!active) return;
    uint32_t packed = kSkip;
    if (smax >= smin && v.coded) {
        const float inv = __frcp_rn(fmaxf(v.scale, 1e-9f));
        // x is an integer multiple of Q (|multiple| <= 15000): the approximate quotient rounds to the same integer
        const int s = __float2int_rn(__fdividef(v.x, Q));
        const uint32_t M = 65536u - (uint32_t)(smax - smin + 1);
        const uint32_t lo = gauss_cum(T, s, smin, M, Q, v.mean, inv), hi = gauss_cum(T, s + 1, smin, M, Q, v.mean, inv);
        if (hi <= lo || s < smin || s > smax) atomicExch(err, 1);   // CDF not monotone at rounding level: undecodable
        else packed = lo | ((hi - lo - 1u) << 16);
    }
    iv[(size_t)g.n_rows * col0 + (size_t)row * dim + k] = packed;
}

__global__ void __launch_bounds__(256)
gauss_level_intervals_kernel(LevelStreams g, const int32_t *__restrict__ minmax, uint32_t *__restrict__ iv,
                             int32_t *__restrict__ err)
{
    __shared__ float T[kPhiN + 1];
    load_phi_table(T);
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    int lo0 = minmax[0], hi0 = minmax[1], lo1 = minmax[2], hi1 = minmax[3], lo2 = minmax[4], hi2 = minmax[5];
    if (hi0 - lo0 + 1 > 32768 || hi1 - lo1 + 1 > 32768 || hi2 - lo2 + 1 > 32768) {
        // cannot happen after the +-15000-step clamp of STE_multistep; such a stream is skipped entirely
        if (threadIdx.x == 0) atomicExch(err, 2);
        if (hi0 - lo0 + 1 > 32768) hi0 = lo0 - 1;
        if (hi1 - lo1 + 1 > 32768) hi1 = lo1 - 1;
        if (hi2 - lo2 + 1 > 32768) hi2 = lo2 - 1;
    }
    // persistent blocks (the table is staged once per block), rows dealt to the warps round-robin; every load of a row is
    // issued before the first value is evaluated, the next row's anchor index one row ahead
    int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int o_next = row < g.n_rows ? g.orig_idx[row] : 0;
    for (; row < g.n_rows; row += warps) {
        const int o = o_next;
        if (row + warps < g.n_rows) o_next = g.orig_idx[row + warps];
        const float *pr = g.params + (size_t)row * kLdG2;
        const bool a1 = lane < kCF - 32, a2 = lane < kCS, a3 = lane < kCO;
        const CodedValue v0 = load_value<0>(g, pr, o, lane, true);
        const CodedValue v1 = load_value<0>(g, pr, o, 32 + lane, a1);
        const CodedValue v2 = load_value<1>(g, pr, o, lane, a2);
        const CodedValue v3 = load_value<2>(g, pr, o, lane, a3);
        const float Qf = pr[172], Qs = pr[173], Qo = pr[174];
        interval_of_value<0>(g, T, v0, Qf, row, lane, true, lo0, hi0, iv, err);
        interval_of_value<0>(g, T, v1, Qf, row, 32 + lane, a1, lo0, hi0, iv, err);
        interval_of_value<1>(g, T, v2, Qs, row, lane, a2, lo1, hi1, iv, err);
        interval_of_value<2>(g, T, v3, Qo, row, lane, a3, lo2, hi2, iv, err);
    }
}

// Encoder pass 2, one thread per chunk of any of the three streams: only the range coder's carried state is sequential.
// Intervals are read two at a time (every stream has an even number of values per row) one pair ahead of their use.
__global__ void __launch_bounds__(64)
gauss_level_encode_kernel(LevelStreams g, const int32_t *__restrict__ minmax, const uint32_t *__restrict__ iv,
                          uint32_t *__restrict__ out, int32_t *__restrict__ stream_len, int32_t *__restrict__ err)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    int attr, cl;
    if (!locate_chunk(g, c, attr, cl)) return;
    if (minmax[2 * attr + 1] < minmax[2 * attr] || minmax[2 * attr + 1] - minmax[2 * attr] + 1 > 32768) {
        stream_len[c] = 0;   // empty stream / alphabet error (reported by pass 1)
        return;
    }
    const int dim = attr_dim(attr), rows = pick3(g.rows, attr);
    const int r0 = cl * rows, r1 = min(r0 + rows, g.n_rows);
    const uint32_t cap = pick3(g.cap, attr);
    Encoder enc;
    enc.init(out + pick3(g.region, attr) + (size_t)cl * (cap / 4), cap);
    const uint2 *src = reinterpret_cast<const uint2 *>(iv + (size_t)g.n_rows * attr_col0(attr) + (size_t)r0 * dim);
    const int pairs = (r1 - r0) * dim / 2;
    const uint2 none = make_uint2(kSkip, kSkip);
    uint2 n1 = pairs > 0 ? src[0] : none, n2 = pairs > 1 ? src[1] : none, n3 = pairs > 2 ? src[2] : none;
    for (int i = 0; i < pairs; ++i) {   // three pairs in flight
        const uint2 cur = n1;
        n1 = n2; n2 = n3;
        if (i + 3 < pairs) n3 = src[i + 3];
        if (cur.x != kSkip) enc.encode(cur.x & 0xffffu, (cur.x & 0xffffu) + (cur.x >> 16) + 1u);
        if (cur.y != kSkip) enc.encode(cur.y & 0xffffu, (cur.y & 0xffffu) + (cur.y >> 16) + 1u);
    }
    stream_len[c] = (int32_t)enc.finish();
    if (enc.overflow) atomicExch(err, 3);
}

// One warp per chunk: scratch slot -> its place in the level's packed byte string.
__global__ void __launch_bounds__(256)
gauss_level_pack_kernel(LevelStreams g, const uint32_t *__restrict__ scratch, const int32_t *__restrict__ stream_len,
                        const int64_t *__restrict__ stream_off, uint8_t *__restrict__ packed)
{
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    int attr, cl;
    if (!locate_chunk(g, c, attr, cl)) return;
    const uint8_t *src = reinterpret_cast<const uint8_t *>(scratch + pick3(g.region, attr)) + (size_t)cl * pick3(g.cap, attr);
    uint8_t *dst = packed + stream_off[c];
    const int n = stream_len[c];
    for (int i = lane; i < n; i += 32) dst[i] = src[i];
}

// One thread per chunk of any of the three streams.  Symbol search: the inverse Gaussian CDF of the coder's target gives
// the symbol to within a step; it is confirmed / corrected by evaluating C next to it -- normally two evaluations --
// galloping further and bisecting only where the estimate is off (far tails, where C is flat).
__global__ void __launch_bounds__(128)
gauss_level_decode_kernel(LevelStreams g, const uint8_t *__restrict__ b0, const uint8_t *__restrict__ b1,
                          const uint8_t *__restrict__ b2, const int64_t *__restrict__ stream_off,
                          const int32_t *__restrict__ stream_len, const int32_t *__restrict__ minmax)
{
    __shared__ float T[kPhiN + 1];
    load_phi_table(T);
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    int attr, cl;
    if (!locate_chunk(g, c, attr, cl)) return;
    const int dim = attr_dim(attr), col0 = attr_col0(attr), rows = pick3(g.rows, attr);
    const int r0 = cl * rows, r1 = min(r0 + rows, g.n_rows);
    const int smin = minmax[2 * attr], smax = minmax[2 * attr + 1];
    const int L = smax - smin + 1;
    const uint32_t M = 65536u - (uint32_t)L;
    const float inv_M = 1.0f / (float)M;
    // byte offsets are cumulative over the whole level; each stream's bytes start at its first chunk's offset
    const uint8_t *bytes = attr == 0 ? b0 : attr == 1 ? b1 : b2;
    const int64_t first = stream_off[c - cl];
    Decoder dec;
    dec.init(bytes + (stream_off[c] - first), (uint32_t)stream_len[c]);
    float *values = pick3(g.values, attr);
    // Values are decoded two at a time (every stream has an even number per row): their (mean, scale) arrive as 8-byte
    // loads issued one pair ahead, the row's offset masks as one bit field, the decoded pair leaves as one 8-byte store.
    int o_next = g.orig_idx[r0];
    for (int r = r0; r < r1; ++r) {
        const int o = o_next;
        if (r + 1 < r1) o_next = g.orig_idx[r + 1];
        const float *pr = g.params + (size_t)r * kLdG2;
        const float Q = pr[172 + attr];
        const float inv_Q = __frcp_rn(Q);
        float *x = values + (size_t)o * dim;
        uint32_t mkbits = 0xffffffffu;   // bit (k / 3): value k is coded (feat / scaling: always)
        if (attr == 2) {
            mkbits = 0;
            const float2 *mk = reinterpret_cast<const float2 *>(g.mask + (size_t)o * 10);
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                const float2 m2 = __ldg(mk + i);
                mkbits |= (m2.x != 0.0f ? 1u : 0u) << (2 * i) | (m2.y != 0.0f ? 2u : 0u) << (2 * i);
            }
        }
        const float2 *pm = reinterpret_cast<const float2 *>(pr + col0), *ps = reinterpret_cast<const float2 *>(pr + kCE + col0);
        float2 m_next = __ldg(pm), s_next = __ldg(ps);
        for (int k = 0; k < dim; k += 2) {
            const float2 m2 = m_next, s2 = s_next;
            if (k + 2 < dim) {
                m_next = __ldg(pm + (k >> 1) + 1);
                s_next = __ldg(ps + (k >> 1) + 1);
            }
            float out0 = 0.0f, out1 = 0.0f;
#pragma unroll 1
            for (int h = 0; h < 2; ++h) {
            if (!((mkbits >> ((k + h) / 3)) & 1u)) continue;   // not coded: decodes to 0
            const float mean = h ? m2.y : m2.x;
            const float sc = fmaxf(h ? s2.y : s2.x, 1e-9f);
            const float inv = __frcp_rn(sc);
            const uint32_t v = dec.target();
            // z-score of the lower boundary of symbol s
            auto zscore = [&](int s) { return (((float)s - 0.5f) * Q - mean) * inv; };
            int lo_s, hi_s;
            uint32_t clo = 0, chi = 0;
            bool have_lo = false, have_hi = false;  // clo == C(lo_s) / chi == C(hi_s + 1) already evaluated
            // Flat tails first.  Below z = -kFlat the rounded Gaussian term of C is exactly 0 (Phi(-4.7) M < 0.09) and above
            // +kFlat it is exactly M, so C(s) = s - smin resp. M + s - smin there and the symbol follows from v alone.  A
            // badly predicted value (most of an untrained model's) costs ~16 bits but no search at all.
            constexpr float kFlat = 4.7f;
            const int s_a = smin + (int)v, s_b = s_a - (int)M;
            const float z_a = zscore(s_a);
            if (s_a <= smax && z_a < -kFlat) {
                lo_s = hi_s = s_a; clo = v; have_lo = true;
                if (z_a + Q * inv < -kFlat) { chi = v + 1u; have_hi = true; }
            } else if (s_b >= smin && s_b <= smax && zscore(s_b) > kFlat) {
                lo_s = hi_s = s_b; clo = v; chi = v + 1u; have_lo = have_hi = true;
            } else {
                // The symbol lies within (or one step outside) mean +- kFlat sigma: were it deeper in a tail, one of the
                // two cases above would have recognised it.
                // invariant: C(lo_s) <= v (or lo_s == smin), C(hi_s + 1) > v (or hi_s == smax)
                lo_s = (int)fminf(fmaxf(floorf((mean - kFlat * sc) * inv_Q) - 1.0f, (float)smin), (float)smax);
                hi_s = (int)fminf(fmaxf(ceilf((mean + kFlat * sc) * inv_Q) + 2.0f, (float)lo_s), (float)smax);
                // C(s) = rn(Phi_s M) + (s - smin) <= v: invert Phi with the ramp term taken at the current estimate of s
                // (first the mean's symbol)
                int m = min(max(__float2int_rn(mean * inv_Q), lo_s), hi_s);
                const int iters = L > 2048 ? 2 : 1;   // the second pass only pays where the ramp is a large part of C
                for (int it = 0; it < iters; ++it) {
                    float u = ((float)v - (float)(m - smin) + 0.5f) * inv_M;
                    u = fminf(fmaxf(u, 1e-7f), 1.0f - 1e-7f);
                    const float xs = fmaf(normcdfinvf(u), sc, mean);
                    m = (int)fminf(fmaxf(floorf(fmaf(xs, inv_Q, 0.5f)), (float)lo_s), (float)hi_s);
                }
                // confirm: probe the estimate, then gallop away from it until the symbol is bracketed, then bisect
                bool up = true, bracketed = false;
                for (int step = 0; lo_s < hi_s && !bracketed; step = 2 * step + 1) {
                    if (step) m = up ? min(lo_s + step, hi_s) : max(hi_s - step + 1, lo_s + 1);
                    const uint32_t cw = gauss_cum(T, m, smin, M, Q, mean, inv);
                    const bool le = cw <= v;
                    if (le) { lo_s = m; clo = cw; have_lo = true; } else { hi_s = m - 1; chi = cw; have_hi = true; }
                    if (step) bracketed = le != up; else up = le;
                }
                while (lo_s < hi_s) {
                    const int mid = lo_s + (hi_s - lo_s + 1) / 2;
                    const uint32_t cw = gauss_cum(T, mid, smin, M, Q, mean, inv);
                    if (cw <= v) { lo_s = mid; clo = cw; have_lo = true; } else { hi_s = mid - 1; chi = cw; have_hi = true; }
                }
            }
            if (!have_lo) clo = gauss_cum(T, lo_s, smin, M, Q, mean, inv);
            if (!have_hi || hi_s < lo_s) chi = gauss_cum(T, lo_s + 1, smin, M, Q, mean, inv);
            dec.consume(clo, chi > clo ? chi : clo + 1);
            if (h) out1 = (float)lo_s * Q; else out0 = (float)lo_s * Q;
            }
            *reinterpret_cast<float2 *>(x + k) = make_float2(out0, out1);
        }
    }
}

// ---- static-table streams (hyper latents: one table per channel; offset masks: one table) ------------
// symbols[n][C] int16 (already relative to the table's first symbol); table of channel c = tables[(c % T)][.]
// kSmem: the tables (T * table_ld words) are staged in shared memory -- the coder's table look-ups (encode: two per
// symbol, decode: a binary search per symbol) are then off the global-memory latency path.
template <bool kSmem>
__device__ __forceinline__ const uint32_t *stage_tables(const uint32_t *__restrict__ tables, int words)
{
    extern __shared__ uint32_t s_tables[];
    if (!kSmem) return tables;
    for (int i = threadIdx.x; i < words; i += blockDim.x) s_tables[i] = tables[i];
    __syncthreads();
    return s_tables;
}

template <bool kSmem>
__global__ void __launch_bounds__(64)
table_encode_kernel(const int16_t *__restrict__ symbols, int n_rows, int C, int chunk_rows, const uint32_t *__restrict__ tables,
                    int T, int table_ld, uint32_t *__restrict__ out, uint32_t cap_bytes, int32_t *__restrict__ stream_len,
                    int32_t *__restrict__ err)
{
    const uint32_t *tabs = stage_tables<kSmem>(tables, T * table_ld);
    const int chunk = blockIdx.x * blockDim.x + threadIdx.x;
    const int n_chunks = (n_rows + chunk_rows - 1) / chunk_rows;
    if (chunk >= n_chunks) return;
    const int r0 = chunk * chunk_rows, r1 = min(r0 + chunk_rows, n_rows);
    Encoder enc;
    enc.init(out + (size_t)chunk * (cap_bytes / 4), cap_bytes);
    const int16_t *sym = symbols + (size_t)r0 * C;
    const int total = (r1 - r0) * C;
    int s_next = total > 0 ? sym[0] : 0;
    for (int i = 0, c = 0; i < total; ++i) {
        const int s = s_next;
        if (i + 1 < total) s_next = sym[i + 1];
        const uint32_t *tb = tabs + (size_t)(c % T) * table_ld;
        c = c + 1 == C ? 0 : c + 1;
        if (s < 0 || s + 1 >= table_ld || tb[s + 1] <= tb[s]) {
            atomicExch(err, 1);
            continue;
        }
        enc.encode(tb[s], tb[s + 1]);
    }
    stream_len[chunk] = (int32_t)enc.finish();
    if (enc.overflow) atomicExch(err, 3);
}

template <bool kSmem>
__global__ void __launch_bounds__(64)
table_decode_kernel(const uint8_t *__restrict__ bytes, const int64_t *__restrict__ stream_off,
                    const int32_t *__restrict__ stream_len, int n_rows, int C, int chunk_rows,
                    const uint32_t *__restrict__ tables, const int32_t *__restrict__ table_len, int T, int table_ld,
                    int16_t *__restrict__ symbols)
{
    const uint32_t *tabs = stage_tables<kSmem>(tables, T * table_ld);
    const int chunk = blockIdx.x * blockDim.x + threadIdx.x;
    const int n_chunks = (n_rows + chunk_rows - 1) / chunk_rows;
    if (chunk >= n_chunks) return;
    const int r0 = chunk * chunk_rows, r1 = min(r0 + chunk_rows, n_rows);
    Decoder dec;
    dec.init(bytes + stream_off[chunk], (uint32_t)stream_len[chunk]);
    for (int r = r0; r < r1; ++r)
        for (int c = 0; c < C; ++c) {
            const uint32_t *tb = tabs + (size_t)(c % T) * table_ld;
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

// host copy of the Phi table; uploaded to each device on first use
static const float *phi_table_host()
{
    static float table[codec::kPhiN + 1];
    static bool ready = false;
    if (!ready) {
        for (int j = 0; j <= codec::kPhiN; ++j) {
            const double z = -4.75 + 9.5 * (double)j / (double)codec::kPhiN;   // exact: 9.5 / 4096 is a binary fraction
            table[j] = (float)(0.5 * erfc(-z * 0.70710678118654752440));
        }
        table[0] = 0.0f;
        table[codec::kPhiN] = 1.0f;
        ready = true;
    }
    return table;
}

static int ensure_phi_table()
{
    static bool uploaded[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (!uploaded[dev]) {
        if (cudaMemcpyToSymbol(codec::g_phi_table, phi_table_host(), sizeof(float) * (codec::kPhiN + 1)) != cudaSuccess) {
            set_error("entropy codec: uploading the Phi table failed: %s", cudaGetErrorString(cudaGetLastError()));
            return -101;
        }
        uploaded[dev] = true;
    }
    return 0;
}

extern "C" int cgs_codec_phi_table(float *table, int n, float *z0, float *inv_h)
{
    if (!table || n != codec::kPhiN + 1) {
        set_error("%s: the table has %d entries", __func__, codec::kPhiN + 1);
        return -2;
    }
    const float *t = phi_table_host();
    for (int j = 0; j < n; ++j) table[j] = t[j];
    if (z0) *z0 = codec::kPhiZ0;
    if (inv_h) *inv_h = codec::kPhiInvH;
    return 0;
}

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

// fills the launch descriptor shared by the level-wide entry points; returns 0 or an error code
static int level_streams(codec::LevelStreams *g, const char *fn, const int32_t *orig_idx, int n_rows, const int *chunk_rows,
                         const float *params, const float *mask, float *feat_q, float *scaling_q, float *offsets_q)
{
    if (!orig_idx || !params || !mask || !feat_q || !scaling_q || !offsets_q || !chunk_rows) {
        set_error("%s: null pointer argument", fn);
        return -2;
    }
    g->orig_idx = orig_idx; g->params = params; g->mask = mask; g->n_rows = n_rows;
    g->values[0] = feat_q; g->values[1] = scaling_q; g->values[2] = offsets_q;
    int64_t word = 0;
    for (int a = 0; a < 3; ++a) {
        if (chunk_rows[a] <= 0) {
            set_error("%s: invalid chunk size", fn);
            return -2;
        }
        g->rows[a] = chunk_rows[a];
        g->n_chunks[a] = (n_rows + chunk_rows[a] - 1) / chunk_rows[a];
        g->cap[a] = (uint32_t)cgs_codec_gauss_stream_capacity(a, chunk_rows[a]);
        g->region[a] = word;
        word += (int64_t)g->n_chunks[a] * (g->cap[a] / 4);
    }
    return 0;
}

extern "C" int64_t cgs_codec_gauss_level_chunks(int n_rows, const int *chunk_rows, int32_t *n_chunks)
{
    if (!chunk_rows || n_rows < 0) return -1;
    int64_t words = 0;
    for (int a = 0; a < 3; ++a) {
        if (chunk_rows[a] <= 0) return -1;
        const int n = (n_rows + chunk_rows[a] - 1) / chunk_rows[a];
        if (n_chunks) n_chunks[a] = n;
        words += (int64_t)n * (cgs_codec_gauss_stream_capacity(a, chunk_rows[a]) / 4);
    }
    return words;
}

extern "C" int cgs_codec_gauss_level_minmax(const int32_t *orig_idx, int n_rows, const float *params, const float *mask,
                                            const float *feat_q, const float *scaling_q, const float *offsets_q,
                                            int32_t *minmax, void *stream)
{
    CGS_CHECK_PTR(minmax);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    StageScope sc(ST_CODEC, st, n_rows > 0 ? 2 : 1);
    codec::minmax_init_kernel<<<1, 32, 0, st>>>(minmax);
    if (n_rows <= 0) return check_launch(__func__);
    codec::LevelStreams g;
    const int one[3] = {1, 1, 1};
    if (int e = level_streams(&g, __func__, orig_idx, n_rows, one, params, mask, const_cast<float *>(feat_q),
                              const_cast<float *>(scaling_q), const_cast<float *>(offsets_q)))
        return e;
    const int rows_per_block = 8 * codec::kRowsPerWarp;
    codec::gauss_level_minmax_kernel<<<(n_rows + rows_per_block - 1) / rows_per_block, 256, 0, st>>>(g, minmax);
    return check_launch(__func__);
}

extern "C" int cgs_codec_gauss_level_encode(const int32_t *orig_idx, int n_rows, const int *chunk_rows, const float *params,
                                            const float *mask, const float *feat_q, const float *scaling_q,
                                            const float *offsets_q, const int32_t *minmax, uint32_t *intervals,
                                            uint32_t *scratch, int32_t *stream_len, int32_t *err, void *stream)
{
    if (n_rows <= 0) return 0;
    CGS_CHECK_PTR(minmax); CGS_CHECK_PTR(intervals); CGS_CHECK_PTR(scratch); CGS_CHECK_PTR(stream_len); CGS_CHECK_PTR(err);
    codec::LevelStreams g;
    if (int e = level_streams(&g, __func__, orig_idx, n_rows, chunk_rows, params, mask, const_cast<float *>(feat_q),
                              const_cast<float *>(scaling_q), const_cast<float *>(offsets_q)))
        return e;
    if (int e = ensure_phi_table()) return e;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    StageScope sc(ST_CODEC, st, 2);
    codec::gauss_level_intervals_kernel<<<std::min((n_rows + 7) / 8, 148 * 8), 256, 0, st>>>(g, minmax, intervals, err);
    const int chunks = g.n_chunks[0] + g.n_chunks[1] + g.n_chunks[2];
    codec::gauss_level_encode_kernel<<<(chunks + 63) / 64, 64, 0, st>>>(g, minmax, intervals, scratch, stream_len, err);
    return check_launch(__func__);
}

extern "C" int cgs_codec_gauss_level_pack(int n_rows, const int *chunk_rows, const uint32_t *scratch,
                                          const int32_t *stream_len, const int64_t *stream_off, uint8_t *packed, void *stream)
{
    if (n_rows <= 0) return 0;
    CGS_CHECK_PTR(scratch); CGS_CHECK_PTR(stream_len); CGS_CHECK_PTR(stream_off); CGS_CHECK_PTR(packed);
    codec::LevelStreams g;
    float dummy;   // the descriptor's tensor pointers are not used by the pack kernel
    if (int e = level_streams(&g, __func__, reinterpret_cast<const int32_t *>(&dummy), n_rows, chunk_rows, &dummy, &dummy,
                              &dummy, &dummy, &dummy))
        return e;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    StageScope sc(ST_CODEC, st, 1);
    const int64_t chunks = (int64_t)g.n_chunks[0] + g.n_chunks[1] + g.n_chunks[2];
    codec::gauss_level_pack_kernel<<<(unsigned)((chunks * 32 + 255) / 256), 256, 0, st>>>(g, scratch, stream_len, stream_off,
                                                                                         packed);
    return check_launch(__func__);
}

extern "C" int cgs_codec_gauss_level_decode(const int32_t *orig_idx, int n_rows, const int *chunk_rows, const float *params,
                                            const float *mask, const uint8_t *feat_bytes, const uint8_t *scaling_bytes,
                                            const uint8_t *offsets_bytes, const int64_t *stream_off,
                                            const int32_t *stream_len, const int32_t *minmax, float *feat_q,
                                            float *scaling_q, float *offsets_q, void *stream)
{
    if (n_rows <= 0) return 0;
    CGS_CHECK_PTR(feat_bytes); CGS_CHECK_PTR(scaling_bytes); CGS_CHECK_PTR(offsets_bytes); CGS_CHECK_PTR(stream_off);
    CGS_CHECK_PTR(stream_len); CGS_CHECK_PTR(minmax);
    codec::LevelStreams g;
    if (int e = level_streams(&g, __func__, orig_idx, n_rows, chunk_rows, params, mask, feat_q, scaling_q, offsets_q)) return e;
    if (int e = ensure_phi_table()) return e;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    StageScope sc(ST_CODEC, st, 1);
    const int chunks = g.n_chunks[0] + g.n_chunks[1] + g.n_chunks[2];
    codec::gauss_level_decode_kernel<<<(chunks + 127) / 128, 128, 0, st>>>(g, feat_bytes, scaling_bytes, offsets_bytes, stream_off,
                                                                       stream_len, minmax);
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
    const size_t table_bytes = (size_t)T * table_ld * sizeof(uint32_t);
    if (table_bytes <= 40 * 1024)
        codec::table_encode_kernel<true><<<(n_chunks + 63) / 64, 64, table_bytes, static_cast<cudaStream_t>(stream)>>>(
            symbols, n_rows, C, chunk_rows, tables, T, table_ld, scratch, (uint32_t)cap_bytes, stream_len, err);
    else
        codec::table_encode_kernel<false><<<(n_chunks + 63) / 64, 64, 0, static_cast<cudaStream_t>(stream)>>>(
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
    const size_t table_bytes = (size_t)T * table_ld * sizeof(uint32_t);
    if (table_bytes <= 40 * 1024)
        codec::table_decode_kernel<true><<<(n_chunks + 63) / 64, 64, table_bytes, static_cast<cudaStream_t>(stream)>>>(
            bytes, stream_off, stream_len, n_rows, C, chunk_rows, tables, table_len, T, table_ld, symbols);
    else
        codec::table_decode_kernel<false><<<(n_chunks + 63) / 64, 64, 0, static_cast<cudaStream_t>(stream)>>>(
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
