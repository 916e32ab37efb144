// ntt.cu -- multi-pass radix-2^r NTT (r <= 8 per pass): register-resident radix-8 butterfly units, shared memory only
// for the exchanges between them, precomputed twiddles.
//
// n = 2^k is split into P = ceil(k/8) digits N_1..N_P (balanced).  With i = (i_1,..,i_P) (i_1 most significant)
// and j = j_1 + N_1 j_2 + N_1 N_2 j_3 + ..:
//   pass p < P : for every (j_1..j_{p-1}) and every inner index i' (C adjacent ones per CTA, so global accesses
//                are C*32-byte runs): N_p-point DFT over digit p, times omega_n^(i' * j_p * N_1..N_{p-1}),
//                written back to the same positions of the other buffer;
//   pass P     : N_P-point DFTs over contiguous runs (C runs per CTA, adjacent in j_1), written to the digit-
//                reversed positions, which are C*32-byte runs again.
// Each pass reads one buffer and writes the other, so the result lands in d_dst iff P is odd (fft.cu:193-211).
//
// Inside a pass the r radix-2 DIF stages are done in groups of up to three: a thread holds the 8 elements of a
// radix-8 unit in registers (12 butterflies = 12 Montgomery products), the first group loads straight from global
// memory and the last one stores straight to it, so an 8-stage pass crosses shared memory twice instead of 8 times.
// IMAD-bound: (n/2) log2 n butterfly products + 1 (direct table) or 2 (two-level table) products per element and
// pass boundary; HBM traffic 64 B per element and pass.
#include "ntt.cuh"
#include "field.cuh"

#include <algorithm>
#include <array>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <vector>

namespace pb {

// Tile shape, measured at 2^24 on B200 (profiles/r2_ntt_variants.md): 256 threads x radix-8 units (124 registers, 2 CTAs of 72 KB per SM)
// 4.10 ms; 128 threads x radix-4 units capped at 5 CTAs per SM (92 registers, 36 KB tiles) 3.86 ms -- more, smaller CTAs overlap one
// tile's global loads and barriers with the others' butterflies; 128 x radix-8 4.14, 256 x radix-4 4.26, 6 CTAs per SM (80 registers) 3.93.
#ifndef PANDA_NTT_THREADS
#define PANDA_NTT_THREADS 128
#endif
static constexpr int NTT_THREADS = PANDA_NTT_THREADS;  // threads per CTA; a tile holds 8 elements per thread (one radix-8 unit each)
#ifndef PANDA_NTT_MIN_CTAS
#define PANDA_NTT_MIN_CTAS 5
#endif
static constexpr int NTT_MIN_CTAS = PANDA_NTT_MIN_CTAS;  // __launch_bounds__ minimum of resident CTAs per SM (caps the registers)
static constexpr unsigned NTT_TILE_LOG = NTT_THREADS == 128 ? 10 : NTT_THREADS == 512 ? 12 : 11;
static constexpr unsigned NTT_MIN_LOGC = NTT_TILE_LOG - 8;       // columns per tile of a full radix-256 pass: 32-byte elements in runs of 2^NTT_MIN_LOGC
#ifndef PANDA_NTT_GROUP
#define PANDA_NTT_GROUP 2
#endif
static constexpr unsigned NTT_GROUP = PANDA_NTT_GROUP;  // radix-2 stages a thread does on register-resident elements between two shared-memory exchanges
static constexpr unsigned NTT_MAX_RADIX_LOG = 8;      // fft.cu:10 MAX_LOG2_RADIX
static constexpr unsigned NTT_MAX_PARTS = 16;         // destination buffers of the exchange step (GPUs of one box)
static constexpr unsigned NTT_DIRECT_LOG = 26;        // pass boundaries with at most 2^26 distinct twiddles get a direct table (one product per element
                                                      // instead of two through the two-level table): 32 B of extra HBM traffic per element of that pass,
                                                      // 512 MiB per cached (omega, 2^24) pair.  With the small tiles above 3.68 vs 3.86 ms at 2^24, 16.5 vs
                                                      // 17.1 ms at 2^26; a table that does not fit falls back to 2^20 (ntt_get_tables)

struct NttShape {
    unsigned log_n, passes;
    unsigned r[4];
};

static NttShape ntt_shape(unsigned log_n) {
    NttShape s{};
    s.log_n = log_n;
    s.passes = (log_n + NTT_MAX_RADIX_LOG - 1) / NTT_MAX_RADIX_LOG;
    for (unsigned p = 0; p < s.passes; p++) s.r[p] = log_n / s.passes + (p < log_n % s.passes ? 1 : 0);
    return s;
}

// ---- twiddle tables --------------------------------------------------------------------------------
// One device array per (device, log n, omega, direction):
//   [omega][scale = 2^-log_n][t_lo: omega^e, e < 2^lo_bits][t_hi: omega^(e << lo_bits)]      two-level table, any exponent < n
//   [stage table of r_a: omega^(e << (k - r_a)), e < 2^(r_a-1)][stage table of r_b]           butterfly twiddles of a pass
//   [direct table of boundary p: omega^(e << log_O_p), e < n >> log_O_p]                      where that is <= 2^20 entries

struct NttSegment { size_t off; uint32_t count; unsigned shift; };

struct NttTableLayout {
    unsigned lo_bits, hi_bits, r_a, r_b;
    size_t off_omega, off_scale, off_lo, off_hi, off_sa, off_sb, words;
    size_t off_direct[4];           // per pass boundary p (twiddle applied at the end of pass p); 0 = none
    unsigned nseg;
    NttSegment seg[8];
};

static unsigned ntt_direct_log() {       // PANDA_NTT_DIRECT_LOG overrides (tuning)
    static const unsigned v = [] { const char *e = getenv("PANDA_NTT_DIRECT_LOG"); return e ? (unsigned)atoi(e) : NTT_DIRECT_LOG; }();
    return v;
}

static NttTableLayout ntt_table_layout(const NttShape &s, unsigned direct_log = ntt_direct_log()) {
    NttTableLayout t{};
    t.lo_bits = (s.log_n + 1) / 2;
    t.hi_bits = s.log_n - t.lo_bits;
    t.r_a = s.passes ? s.r[0] : 0;
    t.r_b = s.passes ? s.r[s.passes - 1] : 0;
    size_t off = 0;
    auto add = [&](uint32_t count, unsigned shift) {
        const size_t at = off;
        t.seg[t.nseg++] = NttSegment{at, count, shift};
        off += (size_t)8 * count;
        return at;
    };
    t.off_omega = off; off += 8;
    t.off_scale = off; off += 8;
    t.off_lo = add(1u << t.lo_bits, 0);
    t.off_hi = add(1u << t.hi_bits, t.lo_bits);
    t.off_sa = add(t.r_a ? 1u << (t.r_a - 1) : 1, t.r_a ? s.log_n - t.r_a : 0);
    t.off_sb = add(t.r_b ? 1u << (t.r_b - 1) : 1, t.r_b ? s.log_n - t.r_b : 0);
    unsigned log_O = 0;
    for (unsigned p = 0; p + 1 < s.passes; p++) {
        const unsigned size_log = s.log_n - log_O;
        if (size_log <= direct_log) t.off_direct[p] = add(1u << size_log, log_O);
        log_O += s.r[p];
    }
    t.words = off;
    return t;
}

template <class F>
PB_DEV F fe_pow_u32(const F &base, uint32_t e) {
    F acc = F::one(), b = base;
#pragma unroll 1
    while (e) {
        if (e & 1) acc = acc * b;
        b = b.sqr();
        e >>= 1;
    }
    return acc;
}

struct NttTableArgs { unsigned log_n, nseg; NttSegment seg[8]; size_t off_scale; int inverse; };

template <class P>
__global__ void k_ntt_tables(uint32_t *tab, const NttTableArgs a) {
    using F = Fe<P>;
    uint32_t total = 0;
    for (unsigned s = 0; s < a.nseg; s++) total += a.seg[s].count;
    const uint32_t nmask = a.log_n >= 32 ? 0xffffffffu : ((1u << a.log_n) - 1);
    const F omega = F::load_plain(tab);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i <= total; i += gridDim.x * blockDim.x) {
        if (i == total) {
            // scale = 2^-log_n (Montgomery): halve ONE log_n times
            F x = F::one().canon();
            for (unsigned k = 0; k < a.log_n; k++) {
                uint32_t odd = x.l[0] & 1, carry = 0;
                if (odd) {   // x += p (may carry out of the top limb)
                    x.l[0] = ptx::add_cc(x.l[0], P::mod(0));
#pragma unroll
                    for (int q = 1; q < F::N; q++) x.l[q] = ptx::addc_cc(x.l[q], P::mod(q));
                    carry = ptx::addc(0, 0);
                }
#pragma unroll
                for (int q = 0; q < F::N - 1; q++) x.l[q] = __funnelshift_r(x.l[q], x.l[q + 1], 1);
                x.l[F::N - 1] = (x.l[F::N - 1] >> 1) | (carry << 31);
            }
            x.store(tab + a.off_scale);
            continue;
        }
        uint32_t t = i;
        unsigned s = 0;
        while (t >= a.seg[s].count) { t -= a.seg[s].count; s++; }
        uint32_t e = (t << a.seg[s].shift) & nmask;
        if (a.inverse) e = (0u - e) & nmask;
        fe_pow_u32(omega, e).canon().store(tab + a.seg[s].off + (size_t)t * 8);
    }
}

// ---- shared-memory tile helpers --------------------------------------------------------------------
// tile element (row, col) limb l lives at sm[l * LS + row * CP + col]

template <class F>
PB_DEV F sm_load(const uint32_t *sm, uint32_t LS, uint32_t idx) {
    F x;
#pragma unroll
    for (int l = 0; l < F::N; l++) x.l[l] = sm[l * LS + idx];
    return x;
}
template <class F>
PB_DEV void sm_store(uint32_t *sm, uint32_t LS, uint32_t idx, const F &x) {
#pragma unroll
    for (int l = 0; l < F::N; l++) sm[l * LS + idx] = x.l[l];
}

// DIF stages s .. s+G-1 of an N = 2^r point transform on the 2^G elements a thread holds:
// x[k] is row hi * 2^(r-s) + k * 2^(r-G-s) + low.  Butterfly (u, v) -> (u + v, (u - v) * w^(j << stage)), j = lower row mod half.
template <class F, int G>
PB_DEV void radix_unit(F (&x)[1 << G], unsigned s, unsigned r, uint32_t low, const uint32_t *__restrict__ stage_tw) {
#pragma unroll
    for (int t = 0; t < G; t++) {
        constexpr int size = 1 << G;
        const int dist = size >> (t + 1);
#pragma unroll
        for (int k0 = 0; k0 < size; k0++) {
            if (k0 & dist) continue;
            const uint32_t j = ((uint32_t)(k0 & (dist - 1)) << (r - G - s)) + low;
            const uint32_t e = j << (s + t);
            const F u = x[k0], v = x[k0 + dist];
            x[k0] = u + v;
            F d = u - v;
            if (e) d = d * F::load(stage_tw + (size_t)e * F::N);
            x[k0 + dist] = d;
        }
    }
}

// stage groups of a pass: ceil(r / 3) groups, as even as possible
struct NttGroups { unsigned n, g[8]; };
PB_DEV NttGroups ntt_groups(unsigned r) {
    NttGroups q{};
    q.n = (r + NTT_GROUP - 1) / NTT_GROUP;
    for (unsigned i = 0; i < q.n; i++) q.g[i] = r / q.n + (i < r % q.n ? 1 : 0);
    return q;
}

// One group of stages over the whole N x C tile.  ld(row, col) / st(row, col, x) reach global memory for the first / last
// group and shared memory otherwise (decided by the caller's functors).  row_fastest: consecutive threads take consecutive rows
// (contiguous runs of the last pass), otherwise consecutive columns.
template <class F, int G, class Ld, class St>
PB_DEV void run_group(unsigned s, unsigned r, unsigned logC, bool row_fastest, const uint32_t *__restrict__ stage_tw, Ld &&ld, St &&st) {
    const uint32_t units = (1u << (r - G)) << logC;
    const unsigned lowbits = r - G - s;
#pragma unroll 1
    for (uint32_t u = threadIdx.x; u < units; u += blockDim.x) {
        uint32_t col, rest;
        if (row_fastest) { rest = u & ((1u << (r - G)) - 1); col = u >> (r - G); }
        else { col = u & ((1u << logC) - 1); rest = u >> logC; }
        const uint32_t low = rest & ((1u << lowbits) - 1), hi = rest >> lowbits;
        const uint32_t row0 = (hi << (r - s)) + low;
        F x[1 << G];
#pragma unroll
        for (int k = 0; k < (1 << G); k++) x[k] = ld(row0 + ((uint32_t)k << lowbits), col);
        radix_unit<F, G>(x, s, r, low, stage_tw);
#pragma unroll
        for (int k = 0; k < (1 << G); k++) st(row0 + ((uint32_t)k << lowbits), col, x[k]);
    }
}

// all r stages of the tile: gl/gs are the global load / store functors, shared memory carries the exchanges in between.
// After the r DIF stages row `pos` holds X[bitrev_r(pos)]; gs receives (pos, col, value).
template <class F, class GLd, class GSt>
PB_DEV void tile_transform(uint32_t *sm, uint32_t LS, uint32_t CP, unsigned r, unsigned logC, bool first_row_fastest,
                           const uint32_t *__restrict__ stage_tw, GLd &&gl, GSt &&gs) {
    const NttGroups q = ntt_groups(r);
    unsigned s = 0;
#pragma unroll 1
    for (unsigned gi = 0; gi < q.n; gi++) {
        const bool first = gi == 0, last = gi + 1 == q.n;
        auto ld = [&](uint32_t row, uint32_t col) -> F { return first ? gl(row, col) : sm_load<F>(sm, LS, row * CP + col); };
        auto st = [&](uint32_t row, uint32_t col, const F &x) { if (last) gs(row, col, x); else sm_store(sm, LS, row * CP + col, x); };
        const bool rf = first && !last && first_row_fastest;
        if (NTT_GROUP >= 3 && q.g[gi] == 3) { if constexpr (NTT_GROUP >= 3) run_group<F, 3>(s, r, logC, rf, stage_tw, ld, st); }
        else if (q.g[gi] == 2) run_group<F, 2>(s, r, logC, rf, stage_tw, ld, st);
        else run_group<F, 1>(s, r, logC, rf, stage_tw, ld, st);
        if (!last) __syncthreads();
        s += q.g[gi];
    }
}

struct NttPassArgs {
    unsigned log_n, r, log_M, log_O, logC, lo_bits;
    unsigned log_bpt;                 // last pass: log2(CTAs per transform); the CTA index above that is the batch row
    unsigned passes, rad[4];          // all radices (last pass: digit reversal)
    const uint32_t *stage_tw, *t_lo, *t_hi, *t_direct, *scale;
};

// pass p < P
template <class P>
__global__ void __launch_bounds__(NTT_THREADS, NTT_MIN_CTAS) k_ntt_cols(const uint32_t *__restrict__ src, uint32_t *__restrict__ dst, const NttPassArgs a) {
    using F = Fe<P>;
    extern __shared__ uint32_t sm[];
    const uint32_t N = 1u << a.r, C = 1u << a.logC, CP = C > 1 ? C + 1 : 1, LS = N * CP;
    const uint32_t blocks_per_o = 1u << (a.log_M - a.logC);
    const uint32_t o = blockIdx.x / blocks_per_o, ib = blockIdx.x % blocks_per_o;     // o also carries the batch row
    const uint32_t i0 = ib << a.logC;
    const size_t base = ((size_t)o << (a.r + a.log_M)) + i0;
    const uint32_t lo_mask = (1u << a.lo_bits) - 1;
    auto gl = [&](uint32_t row, uint32_t col) -> F { return F::load(src + (base + ((size_t)row << a.log_M) + col) * F::N); };
    auto gs = [&](uint32_t pos, uint32_t col, F x) {
        const uint32_t jp = __brev(pos) >> (32 - a.r);
        const uint32_t prod = (i0 + col) * jp;                    // < n >> log_O
        if (prod) {
            F tw;
            if (a.t_direct) tw = F::load(a.t_direct + (size_t)prod * F::N);
            else {
                const uint32_t ex = prod << a.log_O;
                tw = F::load(a.t_hi + (size_t)(ex >> a.lo_bits) * F::N) * F::load(a.t_lo + (size_t)(ex & lo_mask) * F::N);
            }
            x = x * tw;
        }
        x.canon().store(dst + (base + ((size_t)jp << a.log_M) + col) * F::N);
    };
    tile_transform<F>(sm, LS, CP, a.r, a.logC, false, a.stage_tw, gl, gs);
}

// pass P (last): contiguous runs in, digit-reversed positions out
template <class P>
__global__ void __launch_bounds__(NTT_THREADS, NTT_MIN_CTAS) k_ntt_last(const uint32_t *__restrict__ src, uint32_t *__restrict__ dst, const NttPassArgs a) {
    using F = Fe<P>;
    extern __shared__ uint32_t sm[];
    const uint32_t N = 1u << a.r, C = 1u << a.logC, CP = C > 1 ? C + 1 : 1, LS = N * CP;
    const unsigned log_OP = a.log_O;                                  // log2(N_1 .. N_{P-1})
    const unsigned r1 = a.passes > 1 ? a.rad[0] : 0;
    const uint32_t rest_count = 1u << (log_OP - r1);
    const uint32_t bid = blockIdx.x & ((1u << a.log_bpt) - 1);
    src += ((size_t)(blockIdx.x >> a.log_bpt) << a.log_n) * F::N;      // batch row (contiguous transforms)
    dst += ((size_t)(blockIdx.x >> a.log_bpt) << a.log_n) * F::N;
    const uint32_t j1_0 = (bid / rest_count) << a.logC, rest = bid % rest_count;
    // digit reversal of (j_2 .. j_{P-1})
    uint32_t revp = 0;
    {
        uint32_t tmp = rest;
        unsigned shift_total = 0;
        for (unsigned d = 1; d + 1 < a.passes; d++) shift_total += a.rad[d];
        for (unsigned d = a.passes - 1; d-- > 1;) {       // d = P-2 .. 1  (0-based digit index of j_{d+1})
            shift_total -= a.rad[d];
            revp += (tmp & ((1u << a.rad[d]) - 1)) << shift_total;
            tmp >>= a.rad[d];
        }
    }
    F scale;
    if (a.scale) scale = F::load(a.scale);
    const size_t out_base = (size_t)j1_0 + ((size_t)revp << r1);
    auto gl = [&](uint32_t ip, uint32_t cidx) -> F {
        const size_t o = (size_t)(j1_0 + cidx) * rest_count + rest;
        return F::load(src + ((o << a.r) + ip) * F::N);
    };
    auto gs = [&](uint32_t pos, uint32_t cidx, F x) {
        const uint32_t jp = __brev(pos) >> (32 - a.r);
        if (a.scale) x = x * scale;
        x.canon().store(dst + (out_base + cidx + ((size_t)jp << log_OP)) * F::N);
    };
    tile_transform<F>(sm, LS, CP, a.r, a.logC, true, a.stage_tw, gl, gs);
}

// ---- exchange step of the four-step transform -----------------------------------------------------
// dst[h][(c mod Cp) * ld + col_off + r] = src[r][c] * omega^((row0 + r) * c)      h = c / Cp, Cp = cols / parts
// i.e. a tiled transpose of the rows x cols matrix `src` fused with the four-step twiddle, whose column blocks land in
// `parts` destination buffers: per-rank staging chunks for an NCCL all-to-all, or peer-mapped buffers of the other GPUs
// (NVLink stores straight into the consumer's row layout).  With parts = 1 and no twiddle it is a plain transpose.
// Plain transpose: HBM bound, 32 B read + 32 B written per element in 512-byte runs (6.6 TB/s measured).  With the twiddle: 2.25 products per
// element, integer-pipe bound.
struct NttExchangeArgs {
    const uint32_t *src;
    unsigned log_rows, log_cols, log_part_cols, row0;
    uint32_t *dst[NTT_MAX_PARTS];
    size_t ld, col_off;
    const uint32_t *t_lo, *t_hi;      // nullptr: no twiddle
    unsigned lo_bits, log_n;
};

// A CTA takes NTT_XCHG_TILES 16 x 16 tiles along a row of tiles: thread (tr, tc) keeps its row and steps 16 columns at a time, so its twiddle
// omega^(R * c) advances by the constant omega^(16 R) -- one product for the update instead of the two-level table's product and its two gathers:
// (2 + (K - 1) + K) / K = 2.25 products per element for K = 4 where every element on its own took 3 (ncu on that form: FMA-heavy pipe 82 % of
// elapsed, 10 stall cycles per issue on the math-pipe throttle -- product-bound).  All K elements of a thread are loaded up front and the tiles
// leave through shared memory behind ONE barrier.
static constexpr uint32_t NTT_XCHG_TILES = 4;

template <class P>
__global__ void __launch_bounds__(256) k_ntt_exchange(const NttExchangeArgs a) {
    using F = Fe<P>;
    // limb planes of T * TP words, padded so that the two 4-limb halves of an element sit 16 banks apart (the store phase reads them side by side)
    constexpr uint32_t T = 16, TP = T + 1, LS = T * TP + 4, K = NTT_XCHG_TILES, TILE_WORDS = F::N * LS;
    static_assert(F::N == 8 && (4 * LS) % 32 == 16, "store phase: 16-byte halves of 32-byte elements");
    __shared__ uint32_t sm[K * TILE_WORDS];
    const uint32_t rows = 1u << a.log_rows, cols = 1u << a.log_cols;
    const uint32_t tiles_c = (cols + T * K - 1) / (T * K);
    const uint32_t r_base = (blockIdx.x / tiles_c) * T, c_base0 = (blockIdx.x % tiles_c) * T * K;
    const uint32_t tr = threadIdx.x / T, tc = threadIdx.x % T;
    const uint32_t r_in = r_base + tr;
    if (r_in < rows) {
        F x[K];
#pragma unroll
        for (uint32_t j = 0; j < K; j++) {
            const uint32_t c = c_base0 + j * T + tc;
            if (c < cols) x[j] = F::load(a.src + ((size_t)r_in * cols + c) * F::N);
        }
        if (a.t_lo) {
            const uint32_t lo_mask = (1u << a.lo_bits) - 1;
            const uint64_t n_mask = ((uint64_t)1 << a.log_n) - 1, R = (uint64_t)a.row0 + r_in;
            auto power = [&](uint32_t ex) { return F::load(a.t_hi + (size_t)(ex >> a.lo_bits) * F::N) * F::load(a.t_lo + (size_t)(ex & lo_mask) * F::N); };
            F tw = power((uint32_t)((R * (c_base0 + tc)) & n_mask));
            const F step = power((uint32_t)((R * T) & n_mask));
#pragma unroll
            for (uint32_t j = 0; j < K; j++) {
                if (c_base0 + j * T + tc < cols) x[j] = (x[j] * tw).canon();
                if (j + 1 < K) tw = tw * step;
            }
        }
#pragma unroll
        for (uint32_t j = 0; j < K; j++)
            if (c_base0 + j * T + tc < cols) sm_store(sm + j * TILE_WORDS, LS, tc * TP + tr, x[j]);
    }
    __syncthreads();
    // Store phase: a lane writes ONE 16-byte half of an element, so a warp's store instruction covers 512 contiguous bytes (16 elements of
    // one output row).  With a whole 32-byte element per lane every instruction left 16-byte holes in its sectors -- harmless in local HBM
    // (L2 merges the two instructions), but peer stores leave the GPU as they are issued and NVLink carried twice the packets, half empty
    // (2^26 on two GPUs: exchange step 1.85 -> 1.25 ms = the kernel's local time; plain local transpose 5.1 -> 5.7 TB/s).
#pragma unroll 1
    for (uint32_t j = 0; j < K; j++) {
        const uint32_t c_base = c_base0 + j * T;
        if (c_base >= cols) break;
        const uint32_t *tile = sm + j * TILE_WORDS;
#pragma unroll
        for (uint32_t pass = 0; pass < 2; pass++) {
            const uint32_t e = (threadIdx.x >> 1) + pass * (T * T / 2), half = threadIdx.x & 1;
            const uint32_t oc = e / T, orow = e % T;
            const uint32_t r = r_base + orow, c = c_base + oc;
            if (r < rows && c < cols) {
                const uint32_t *q = tile + (half * 4) * LS + oc * TP + orow;
                const uint4 v = make_uint4(q[0], q[LS], q[2 * LS], q[3 * LS]);
                const uint32_t h = c >> a.log_part_cols, cl = c & ((1u << a.log_part_cols) - 1);
                reinterpret_cast<uint4 *>(a.dst[h] + ((size_t)cl * a.ld + a.col_off + r) * F::N)[half] = v;
            }
        }
    }
}

// ---- coset transforms ---------------------------------------------------------------------------------
template <class P>
__global__ void k_fe_invert(uint32_t *x) {
    using F = Fe<P>;
    fe_inverse(F::load_plain(x)).canon().store(x);
}

// data[i] *= g^i (two-level table of the powers of g): the pre-scaling of a coset NTT / the post-scaling of its inverse.
// HBM-bound pass (64 B per element) with 2 products per element.
template <class P>
__global__ void __launch_bounds__(256) k_ntt_coset_scale(uint32_t *__restrict__ data, uint32_t n, const uint32_t *__restrict__ t_lo,
                                                         const uint32_t *__restrict__ t_hi, unsigned lo_bits) {
    using F = Fe<P>;
    const uint32_t lo_mask = (1u << lo_bits) - 1;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        F x = F::load_plain(data + (size_t)i * F::N);
        if (i) x = x * (F::load(t_hi + (size_t)(i >> lo_bits) * F::N) * F::load(t_lo + (size_t)(i & lo_mask) * F::N));
        x.canon().store(data + (size_t)i * F::N);
    }
}

// ---- bit-reversal permutation (natural <-> bit-reversed order variants of the transforms) ------------------------
// dst[bitrev_k(i)] = src[i].  Tiles of 32 x 32 elements: i = (a, m, b) with a and b the top and bottom 5 bits; a tile fixes m, reads rows
// (a fixed, b running: 1 KiB runs) and writes rows of the reversed index (rev(b) fixed, rev(a) running: 1 KiB runs) through
// shared memory.  HBM-bound: 32 B read + 32 B written per element.
template <class P>
__global__ void __launch_bounds__(256) k_ntt_bit_reverse(const uint32_t *__restrict__ src, uint32_t *__restrict__ dst, unsigned log_n) {
    using F = Fe<P>;
    __shared__ uint32_t sm[F::N * 32 * 33];
    if (log_n < 10) {                                        // small sizes: one element per thread, no tiling
        const uint32_t n = 1u << log_n;
        for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
            const uint32_t j = log_n ? __brev(i) >> (32 - log_n) : 0;
            F::load(src + (size_t)i * F::N).store(dst + (size_t)j * F::N);
        }
        return;
    }
    const unsigned mid_bits = log_n - 10;
    const uint32_t m = blockIdx.x;                           // middle bits
    const uint32_t rm = mid_bits ? __brev(m) >> (32 - mid_bits) : 0;
    for (uint32_t e = threadIdx.x; e < 1024; e += blockDim.x) {
        const uint32_t a = e >> 5, b = e & 31;
        const size_t i = ((size_t)a << (log_n - 5)) | ((size_t)m << 5) | b;
        sm_store(sm, 32 * 33, a * 33 + b, F::load(src + i * F::N));
    }
    __syncthreads();
    for (uint32_t e = threadIdx.x; e < 1024; e += blockDim.x) {
        const uint32_t rb = e >> 5, ra = e & 31;              // output row = rev(b), column = rev(a)
        const uint32_t a = __brev(ra) >> 27, b = __brev(rb) >> 27;
        const size_t j = ((size_t)rb << (log_n - 5)) | ((size_t)rm << 5) | ra;
        sm_load<F>(sm, 32 * 33, a * 33 + b).store(dst + j * F::N);
    }
}

// ---- host side ---------------------------------------------------------------------------------------

// Cache entries are reference counted: a caller holds its entry until its kernels are queued, so a concurrent eviction or
// panda_ntt_tear_down on another thread cannot free the table underneath it (cudaFree waits for work that is already queued).
struct NttCacheEntry {
    int device = 0;
    int kind = 0;           // 0: powers of a root of unity (transform tables); 1: powers of a coset generator (two-level table only)
    unsigned log_n = 0;
    bool inverse = false;
    std::array<uint32_t, 8> omega{};
    uint32_t *d_tab = nullptr;
    cudaEvent_t ready = nullptr;     // recorded after the build on the builder's stream; every other stream waits on it
    unsigned long long last_use = 0;
    NttTableLayout layout{};
    ~NttCacheEntry() {
        int cur = 0;
        cudaGetDevice(&cur);
        if (cur != device) cudaSetDevice(device);
        if (d_tab) cudaFree(d_tab);
        if (ready) cudaEventDestroy(ready);
        if (cur != device) cudaSetDevice(cur);
    }
};
using NttTables = std::shared_ptr<NttCacheEntry>;

static std::mutex g_ntt_mutex;
static std::vector<NttTables> g_ntt_cache;
static unsigned long long g_ntt_clock = 0;
static constexpr size_t NTT_CACHE_MAX = 16;

#define PB_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "[panda-b200] CUDA error %d (%s) at %s:%d\n", (int)e_, cudaGetErrorString(e_), __FILE__, __LINE__); return e_; } } while (0)

static cudaError_t ntt_get_tables(const NttShape &shape, const void *omega_host, bool inverse, cudaStream_t stream, NttTables *out, int kind = 0) {
    int dev = 0;
    PB_CUDA(cudaGetDevice(&dev));
    std::array<uint32_t, 8> om;
    memcpy(om.data(), omega_host, 32);
    std::unique_lock<std::mutex> lock(g_ntt_mutex);
    for (auto &e : g_ntt_cache)
        if (e->device == dev && e->kind == kind && e->log_n == shape.log_n && e->inverse == inverse && e->omega == om) {
            e->last_use = ++g_ntt_clock;
            *out = e;
            lock.unlock();
            // the table may still be being built on another stream
            PB_CUDA(cudaStreamWaitEvent(stream, (*out)->ready, 0));
            return cudaSuccess;
        }
    auto e = std::make_shared<NttCacheEntry>();
    e->device = dev; e->kind = kind; e->log_n = shape.log_n; e->inverse = inverse; e->omega = om;
    e->layout = ntt_table_layout(shape);
    if (kind == 1) {           // coset generator: only the two-level table (segments 0 and 1)
        e->layout.nseg = 2;
        e->layout.words = e->layout.off_sa;
    }
    cudaError_t err = cudaMalloc((void **)&e->d_tab, e->layout.words * 4);
    if (err == cudaErrorMemoryAllocation && kind == 0 && ntt_direct_log() > 20) {      // the big direct tables are an optimisation, not a requirement
        cudaGetLastError();
        e->layout = ntt_table_layout(shape, 20);
        err = cudaMalloc((void **)&e->d_tab, e->layout.words * 4);
    }
    if (err != cudaSuccess) { e->d_tab = nullptr; return err; }
    // e->omega lives as long as the cache entry, so the async copy's source stays valid
    err = cudaMemcpyAsync(e->d_tab, e->omega.data(), 32, cudaMemcpyHostToDevice, stream);
    if (err == cudaSuccess) {
        const NttTableLayout &t = e->layout;
        NttTableArgs ta{};
        ta.log_n = shape.log_n; ta.nseg = t.nseg; ta.off_scale = t.off_scale; ta.inverse = inverse ? 1 : 0;
        if (kind == 1) {           // g has no small order: g^-e is a power of g^-1, not g^(n-e)
            if (inverse) k_fe_invert<Bn254Fr><<<1, 1, 0, stream>>>(e->d_tab);
            ta.inverse = 0;
        }
        uint32_t total = 0;
        for (unsigned s = 0; s < t.nseg; s++) { ta.seg[s] = t.seg[s]; total += t.seg[s].count; }
        const unsigned blocks = std::min<uint32_t>((total + 128) / 128, 148 * 16);
        k_ntt_tables<Bn254Fr><<<blocks, 128, 0, stream>>>(e->d_tab, ta);
        err = cudaGetLastError();
    }
    if (err == cudaSuccess) err = cudaEventCreateWithFlags(&e->ready, cudaEventDisableTiming);
    if (err == cudaSuccess) err = cudaEventRecord(e->ready, stream);
    if (err != cudaSuccess) return err;          // e's destructor frees the table
    e->last_use = ++g_ntt_clock;
    NttTables victim;                            // freed after the lock is dropped (cudaFree synchronises the device)
    if (g_ntt_cache.size() >= NTT_CACHE_MAX) {   // bounded: drop the least recently used entry
        size_t v = 0;
        for (size_t i = 1; i < g_ntt_cache.size(); i++) if (g_ntt_cache[i]->last_use < g_ntt_cache[v]->last_use) v = i;
        victim = g_ntt_cache[v];
        g_ntt_cache.erase(g_ntt_cache.begin() + v);
    }
    g_ntt_cache.push_back(e);
    *out = e;
    lock.unlock();
    victim.reset();
    return cudaSuccess;
}

cudaError_t ntt_release_tables() {               // the current device's tables (one NTT unit per device, like panda_msm_tear_down)
    std::vector<NttTables> dropped;
    {
        std::lock_guard<std::mutex> lock(g_ntt_mutex);
        int cur = 0;
        cudaGetDevice(&cur);
        for (size_t i = 0; i < g_ntt_cache.size();) {
            if (g_ntt_cache[i]->device == cur) { dropped.push_back(g_ntt_cache[i]); g_ntt_cache.erase(g_ntt_cache.begin() + i); }
            else i++;
        }
    }
    dropped.clear();                             // entries still referenced by a caller that is queueing kernels die when it lets go
    return cudaSuccess;
}

cudaError_t ntt_run(NttField field, void *d_src, void *d_dst, unsigned log_n, const void *omega_host, bool inverse,
                    cudaStream_t stream, unsigned *result_in_dst, unsigned batch, float *pass_ms) {
    (void)field;
    if (log_n > 28 || !omega_host) return cudaErrorInvalidValue;      // 2-adicity of BN254 Fr (paramter.cuh:241)
    if (batch == 0 || ((uint64_t)batch << log_n) > ((uint64_t)1 << 31)) return cudaErrorInvalidValue;
    const NttShape shape = ntt_shape(log_n);
    if (result_in_dst) *result_in_dst = shape.passes & 1;
    if (shape.passes == 0) return cudaSuccess;                         // n = 1: the transform is the identity
    NttTables tab;                          // held until the kernels below are queued
    PB_CUDA(ntt_get_tables(shape, omega_host, inverse, stream, &tab));
    const NttTableLayout &t = tab->layout;

    static bool attr_done[64] = {};
    int dev = 0;
    PB_CUDA(cudaGetDevice(&dev));
    const size_t max_smem = (size_t)8 * (256 * ((1u << NTT_MIN_LOGC) + 1)) * 4;   // r = 8 with the fewest columns; shorter transforms with more columns stay below it
    if (dev < 64 && !attr_done[dev]) {
        PB_CUDA(cudaFuncSetAttribute(k_ntt_cols<Bn254Fr>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem));
        PB_CUDA(cudaFuncSetAttribute(k_ntt_last<Bn254Fr>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem));
        attr_done[dev] = true;
    }

    uint32_t *src = (uint32_t *)d_src, *dst = (uint32_t *)d_dst;
    unsigned log_O = 0;
    cudaEvent_t marks[5] = {};                                         // pass_ms (diagnostics only): events between the passes
    if (pass_ms) for (unsigned p = 0; p <= shape.passes; p++) PB_CUDA(cudaEventCreate(&marks[p]));
    for (unsigned p = 0; p < shape.passes; p++) {
        if (pass_ms) PB_CUDA(cudaEventRecord(marks[p], stream));
        NttPassArgs a{};
        a.log_n = log_n; a.r = shape.r[p]; a.log_O = log_O; a.lo_bits = t.lo_bits; a.passes = shape.passes;
        for (unsigned d = 0; d < 4; d++) a.rad[d] = shape.r[d];
        a.log_M = log_n - log_O - a.r;
        a.stage_tw = tab->d_tab + (a.r == t.r_a ? t.off_sa : t.off_sb);
        a.t_lo = tab->d_tab + t.off_lo; a.t_hi = tab->d_tab + t.off_hi;
        a.t_direct = nullptr;
        a.scale = nullptr;
        const bool last = p + 1 == shape.passes;
        if (!last) {
            if (t.off_direct[p]) a.t_direct = tab->d_tab + t.off_direct[p];
            // C adjacent columns per CTA: at least 8 (256-byte runs), more for short transforms so that a tile holds 2048 elements
            // (one radix-8 unit per thread)
            a.logC = std::min<unsigned>(a.log_M, std::max<unsigned>(NTT_MIN_LOGC, NTT_TILE_LOG - std::min<unsigned>(a.r, NTT_TILE_LOG)));
            const uint32_t C = 1u << a.logC, CP = C > 1 ? C + 1 : 1;
            const size_t smem = a.r > NTT_GROUP ? (size_t)8 * ((size_t)(1u << a.r) * CP) * 4 : 0;   // a single stage group never touches smem
            const uint32_t blocks = batch << (log_O + a.log_M - a.logC);   // batch rows extend the outer index: (t * 2^log_O + o)
            k_ntt_cols<Bn254Fr><<<blocks, NTT_THREADS, smem, stream>>>(src, dst, a);
        } else {
            const unsigned r1 = shape.passes > 1 ? shape.r[0] : 0;
            a.logC = std::min<unsigned>(r1, std::max<unsigned>(NTT_MIN_LOGC, NTT_TILE_LOG - std::min<unsigned>(a.r, NTT_TILE_LOG)));
            if (inverse) a.scale = tab->d_tab + t.off_scale;
            const uint32_t C = 1u << a.logC, CP = C > 1 ? C + 1 : 1;
            const size_t smem = a.r > NTT_GROUP ? (size_t)8 * ((size_t)(1u << a.r) * CP) * 4 : 0;
            a.log_bpt = log_O - a.logC;
            const uint32_t blocks = batch << a.log_bpt;
            k_ntt_last<Bn254Fr><<<blocks, NTT_THREADS, smem, stream>>>(src, dst, a);
        }
        PB_CUDA(cudaGetLastError());
        log_O += a.r;
        std::swap(src, dst);
    }
    if (pass_ms) {
        PB_CUDA(cudaEventRecord(marks[shape.passes], stream));
        PB_CUDA(cudaStreamSynchronize(stream));
        for (unsigned p = 0; p < 4; p++) pass_ms[p] = 0.f;
        for (unsigned p = 0; p < shape.passes; p++) cudaEventElapsedTime(&pass_ms[p], marks[p], marks[p + 1]);
        for (unsigned p = 0; p <= shape.passes; p++) cudaEventDestroy(marks[p]);
    }
    return cudaSuccess;
}

cudaError_t ntt_exchange(NttField field, const void *d_src, unsigned log_rows, unsigned log_cols, unsigned row_offset, const void *omega_host,
                         unsigned log_n, bool inverse, unsigned parts, void *const *dst, size_t ld, size_t col_offset, cudaStream_t stream) {
    (void)field;
    if (!d_src || !dst || parts == 0 || parts > NTT_MAX_PARTS || (parts & (parts - 1)) || log_rows + log_cols > 31) return cudaErrorInvalidValue;
    unsigned log_parts = 0; while ((1u << log_parts) < parts) log_parts++;
    if (log_parts > log_cols) return cudaErrorInvalidValue;
    NttExchangeArgs a{};
    a.src = (const uint32_t *)d_src;
    a.log_rows = log_rows; a.log_cols = log_cols; a.log_part_cols = log_cols - log_parts; a.row0 = row_offset;
    for (unsigned h = 0; h < parts; h++) { if (!dst[h]) return cudaErrorInvalidValue; a.dst[h] = (uint32_t *)dst[h]; }
    a.ld = ld; a.col_off = col_offset;
    NttTables tab;
    if (omega_host) {
        if (log_n > 28) return cudaErrorInvalidValue;
        PB_CUDA(ntt_get_tables(ntt_shape(log_n), omega_host, inverse, stream, &tab));
        a.t_lo = tab->d_tab + tab->layout.off_lo; a.t_hi = tab->d_tab + tab->layout.off_hi;
        a.lo_bits = tab->layout.lo_bits; a.log_n = log_n;
    }
    const uint32_t tiles_r = ((1u << log_rows) + 15) / 16, tiles_c = ((1u << log_cols) + 16 * NTT_XCHG_TILES - 1) / (16 * NTT_XCHG_TILES);
    k_ntt_exchange<Bn254Fr><<<tiles_r * tiles_c, 256, 0, stream>>>(a);
    return cudaGetLastError();
}

cudaError_t ntt_coset_scale(NttField field, void *d_data, unsigned log_n, const void *gen_host, bool inverse, cudaStream_t stream) {
    (void)field;
    if (!d_data || !gen_host || log_n > 28) return cudaErrorInvalidValue;
    NttTables tab;
    PB_CUDA(ntt_get_tables(ntt_shape(log_n), gen_host, inverse, stream, &tab, 1));
    const uint32_t n = 1u << log_n;
    k_ntt_coset_scale<Bn254Fr><<<std::min<uint32_t>((n + 255) / 256, 148 * 8), 256, 0, stream>>>((uint32_t *)d_data, n, tab->d_tab + tab->layout.off_lo,
                                                                                                tab->d_tab + tab->layout.off_hi, tab->layout.lo_bits);
    return cudaGetLastError();
}

// omega^(2^k) on the host (Montgomery form in and out, BN254 Fr): the sub-roots of the multi-GPU four-step transform.
// Plain CIOS Montgomery squaring on 4 x 64-bit limbs; a handful of calls per transform, nothing hot.
void ntt_pow2k_host(const void *omega_host, unsigned k, void *out_host) {
    typedef unsigned __int128 u128;
    static const uint64_t P[4] = {0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull};
    static const uint64_t NINV = 0xc2e1f593efffffffull;            // -r^-1 mod 2^64
    uint64_t a[4];
    memcpy(a, omega_host, 32);
    for (unsigned it = 0; it < k; it++) {
        uint64_t t[6] = {0, 0, 0, 0, 0, 0};
        for (int i = 0; i < 4; i++) {
            u128 c = 0;
            for (int j = 0; j < 4; j++) { c += (u128)a[j] * a[i] + t[j]; t[j] = (uint64_t)c; c >>= 64; }
            c += t[4]; t[4] = (uint64_t)c; t[5] = (uint64_t)(c >> 64);
            const uint64_t m = t[0] * NINV;
            c = (u128)m * P[0] + t[0]; c >>= 64;
            for (int j = 1; j < 4; j++) { c += (u128)m * P[j] + t[j]; t[j - 1] = (uint64_t)c; c >>= 64; }
            c += t[4]; t[3] = (uint64_t)c; t[4] = t[5] + (uint64_t)(c >> 64);
        }
        // conditional subtraction to [0, r)
        uint64_t d[4]; unsigned __int128 br = 0;
        for (int j = 0; j < 4; j++) { u128 v = (u128)t[j] - P[j] - (uint64_t)br; d[j] = (uint64_t)v; br = (v >> 64) & 1; }
        const bool ge = t[4] != 0 || br == 0;
        for (int j = 0; j < 4; j++) a[j] = ge ? d[j] : t[j];
    }
    memcpy(out_host, a, 32);
}

cudaError_t ntt_bit_reverse(NttField field, const void *d_src, void *d_dst, unsigned log_n, cudaStream_t stream) {
    (void)field;
    if (!d_src || !d_dst || d_src == d_dst || log_n > 30) return cudaErrorInvalidValue;
    const uint32_t blocks = log_n < 10 ? std::max<uint32_t>(1, (1u << log_n) / 256) : (1u << (log_n - 10));
    k_ntt_bit_reverse<Bn254Fr><<<blocks, 256, 0, stream>>>((const uint32_t *)d_src, (uint32_t *)d_dst, log_n);
    return cudaGetLastError();
}

}  // namespace pb
