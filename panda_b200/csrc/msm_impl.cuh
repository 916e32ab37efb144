// msm_impl.cuh -- kernels and templated driver of the Pippenger MSM (see msm.cu for the overview).
// Instantiated once per curve in msm_bn254.cu / msm_bls12_377.cu so the two compile in parallel.
#pragma once
#include "msm.cuh"
#include "ec.cuh"

#include <algorithm>
#include <cstdio>
#include <type_traits>
#include <vector>

namespace pb {

struct Bn254 {
    using Fq = Fe<Bn254Fq>;
    using Fr = Fe<Bn254Fr>;
    static constexpr int SCALAR_BITS = 254;
};
struct Bls377 {
    using Fq = Fe<Bls377Fq>;
    using Fr = Fe<Bls377Fr>;
    static constexpr int SCALAR_BITS = 253;
};

struct Bls381 {
    using Fq = Fe<Bls381Fq>;
    using Fr = Fe<Bls381Fr>;
    static constexpr int SCALAR_BITS = 255;
};

// -DPANDA_BOUNDS_CHECK=1 (the pool's GPUs refuse compute-sanitizer): every index the sort / accumulation kernels compute into the workspace is
// checked against its array's capacity and a violation traps; the GPU suite is run once per round against a library built this way
// (profiles/scripts/r2_bounds_check.sh).  Compiled out of the product.
#ifndef PANDA_BOUNDS_CHECK
#define PANDA_BOUNDS_CHECK 0
#endif
#if PANDA_BOUNDS_CHECK
#define PB_CHECK(cond, what) do { if (!(cond)) { printf("[panda-b200] bounds check failed: %s (block %u thread %u)\n", what, blockIdx.x, threadIdx.x); __trap(); } } while (0)
#else
#define PB_CHECK(cond, what) do { } while (0)
#endif
static constexpr int ACC_THREADS = 128;
#ifndef PANDA_ACC_MAXNREG
#define PANDA_ACC_MAXNREG 128     // k_accumulate_range: 4 CTAs per SM (16 warps); 136 (3 CTAs, room for a side-stream CTA beside them) measured 29.15 vs 28.38 ms at 2^24
#endif
static constexpr int RED_THREADS = 128;
static constexpr int WIN_THREADS = 256;
static constexpr int BIG_THREADS = 256;
static constexpr uint32_t BIG_SPAN = 8;             // buckets with more partial slots than this are pre-reduced by a whole CTA

// ----------------------------------------------------------------------------------------------------
// Signed-digit recoding of one canonical scalar, shared by the histogram and the scatter kernels.
// Windows 0 .. wide-1 are c bits wide, the others c-1 (the table plan balances its windows this way so that no window is short;
// wide = W gives uniform windows).  Calls f(window, magnitude (1 .. 2^(width-1)), negative) for every non-zero digit.
template <class Fr, class F>
PB_DEV void for_each_digit(const Fr &s, uint32_t c, uint32_t W, uint32_t wide, F &&f) {
    // The limbs are fed one by one (static register indices) into a 64-bit bit buffer from which the windows are cut: ~15 instructions per
    // window where shifting the whole 256-bit value down after every window took ~35 (the class-shard passes were instruction-bound:
    // ncu, 658 warp instructions per scalar).  The inner loop's trip count depends on c only, so the warp stays converged.
    uint64_t buf = 0;
    uint32_t have = 0, w = 0, carry = 0;
    uint32_t cw = w < wide ? c : c - 1;
    auto emit = [&](uint32_t bits) {
        const uint32_t v = bits + carry;
        uint32_t mag = v, neg = 0;
        carry = 0;
        if (w + 1 < W && v > (1u << (cw - 1))) { mag = (1u << cw) - v; neg = 1; carry = 1; }   // the top window is never recoded
        if (mag) f(w, mag, neg);
        w++;
        cw = w < wide ? c : c - 1;
    };
#pragma unroll
    for (int k = 0; k < Fr::N; k++) {
        if (have <= 32) buf |= (uint64_t)s.l[k] << have;      // (beyond that only the top window is left and the remaining limbs are zero)
        have += 32;
        while (have >= cw && w + 1 < W) {
            const uint32_t bits = (uint32_t)buf & ((1u << cw) - 1);
            buf >>= cw;
            have -= cw;
            emit(bits);
        }
    }
    if (w < W) emit((uint32_t)buf);        // the top window takes whatever is left (W * c covers the scalar width)
}

// K1 (windowed plan): scalars -> per-bucket counts and the digit codes, window-major (code[w*n + i]): |d|-1 with the sign in bit 31,
// 0xFFFFFFFF = zero digit.  (32-bit codes since the windows may be wider than 16 bits: at 2^24 c = 17 saves a window.)
// L2-atomic bound: 32 B read + 4W B written per scalar, W reductions into an L2-resident histogram.
static constexpr uint32_t CODE_SKIP32 = 0xFFFFFFFFu;
template <class C>
__global__ void __launch_bounds__(256) k_digits(const uint32_t *__restrict__ scalars, uint32_t n, uint32_t c, uint32_t W, uint32_t wide, uint32_t nb,
                                                uint32_t class_log2, uint32_t class_index, uint32_t *__restrict__ codes, uint32_t *__restrict__ counts) {
    using Fr = typename C::Fr;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        Fr s = Fr::load(scalars + (size_t)i * Fr::N).from_mont();     // canonical integer; the input is left untouched
        uint32_t next_w = 0;
        const uint32_t class_mask = (1u << class_log2) - 1;
        for_each_digit(s, c, W, wide, [&](uint32_t w, uint32_t mag, uint32_t neg) {
            // bucket-class shard: only the buckets congruent to class_index survive, renumbered 0 .. nb-1 (bucket = local * 2^class_log2 + class_index)
            if (((mag - 1) & class_mask) != class_index) return;
            const uint32_t local = (mag - 1) >> class_log2;
            for (; next_w < w; next_w++) codes[(size_t)next_w * n + i] = CODE_SKIP32;
            PB_CHECK(w < W && local < nb, "k_digits: digit");
            codes[(size_t)w * n + i] = local | (neg << 31);
            next_w = w + 1;
            atomicAdd(&counts[(size_t)w * nb + local], 1u);
        });
        for (; next_w < W; next_w++) codes[(size_t)next_w * n + i] = CODE_SKIP32;
    }
}

// K2a/b/c: exclusive scan of the bucket counts of every bucket set -> offsets[set][0..nb], cursor[set][0..nb-1].
// Three small launches (tile sums, scan of the tile sums, apply) so that a 2^21-bucket set is scanned by 512 CTAs.
// Buckets whose entries span more than BIG_SPAN accumulation segments are appended to big_list.
// 256-thread CTAs: in a streamed MSM these kernels start while an earlier chunk's accumulation fills every SM's register file, and a CTA
// of 1024 threads had to wait for two accumulation CTAs of one SM to retire (traced: 0.5 ms for a 30-microsecond scan).
static constexpr uint32_t SCAN_TILE = 4096, SCAN_THREADS = 256, SCAN_PER = SCAN_TILE / SCAN_THREADS;

// block-wide inclusive scan of one value per thread (SCAN_THREADS threads); returns the inclusive prefix, *total the block sum
PB_DEV uint32_t scan_block_incl(uint32_t v, uint32_t *warp_tot, uint32_t *total) {
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if ((int)lane >= o) incl += t; }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t t = lane < SCAN_THREADS / 32 ? warp_tot[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint32_t u = __shfl_up_sync(0xffffffffu, t, o); if ((int)lane >= o) t += u; }
        warp_tot[lane] = t;
    }
    __syncthreads();
    *total = warp_tot[SCAN_THREADS / 32 - 1];
    return incl + (warp ? warp_tot[warp - 1] : 0);
}

static __global__ void __launch_bounds__(SCAN_THREADS) k_scan_tiles(const uint32_t *__restrict__ counts, uint32_t nb, uint32_t tiles_ps, uint32_t *__restrict__ tile_sums) {
    __shared__ uint32_t warp_tot[32];
    const uint32_t set = blockIdx.y, tile = blockIdx.x, tid = threadIdx.x;
    const uint32_t *cw = counts + (size_t)set * nb;
    uint32_t v = 0;
#pragma unroll
    for (uint32_t k = 0; k < SCAN_PER; k++) { const uint32_t idx = tile * SCAN_TILE + k * SCAN_THREADS + tid; if (idx < nb) v += cw[idx]; }
    uint32_t total;
    scan_block_incl(v, warp_tot, &total);
    if (tid == 0) tile_sums[(size_t)set * tiles_ps + tile] = total;
}

// one CTA per set: exclusive scan of its (<= 1024) tile sums in place; the set total goes to offsets[set][nb]
static __global__ void __launch_bounds__(SCAN_THREADS) k_scan_tops(uint32_t *__restrict__ tile_sums, uint32_t tiles_ps, uint32_t nb, uint32_t *__restrict__ offsets) {
    __shared__ uint32_t warp_tot[32];
    const uint32_t set = blockIdx.x, tid = threadIdx.x;
    uint32_t *ts = tile_sums + (size_t)set * tiles_ps;
    uint32_t v4[4], v = 0;                                      // thread owns 4 consecutive tile sums
#pragma unroll
    for (int k = 0; k < 4; k++) { const uint32_t idx = tid * 4 + k; v4[k] = idx < tiles_ps ? ts[idx] : 0; v += v4[k]; }
    uint32_t total;
    uint32_t excl = scan_block_incl(v, warp_tot, &total) - v;
#pragma unroll
    for (int k = 0; k < 4; k++) { const uint32_t idx = tid * 4 + k; if (idx < tiles_ps) ts[idx] = excl; excl += v4[k]; }
    if (tid == 0) offsets[(size_t)set * (nb + 1) + nb] = total;
}

static __global__ void __launch_bounds__(SCAN_THREADS) k_scan_apply(const uint32_t *__restrict__ counts, const uint32_t *__restrict__ tile_sums, uint32_t nb,
                                                                    uint32_t tiles_ps, uint32_t L, uint32_t *__restrict__ offsets, uint32_t *__restrict__ cursor,
                                                                    uint32_t *__restrict__ big_count, uint32_t *__restrict__ big_list) {
    __shared__ uint32_t warp_tot[32];
    const uint32_t set = blockIdx.y, tile = blockIdx.x, tid = threadIdx.x;
    const uint32_t *cw = counts + (size_t)set * nb;
    uint32_t *ow = offsets + (size_t)set * (nb + 1);
    uint32_t *kw = cursor + (size_t)set * nb;
    // thread owns SCAN_PER consecutive counts
    const uint32_t base = tile * SCAN_TILE + tid * SCAN_PER;
    uint32_t cN[SCAN_PER], v = 0;
#pragma unroll
    for (uint32_t k = 0; k < SCAN_PER; k += 4) {
        if (base + k + 3 < nb) { const uint4 q = *reinterpret_cast<const uint4 *>(cw + base + k); cN[k] = q.x; cN[k + 1] = q.y; cN[k + 2] = q.z; cN[k + 3] = q.w; }
        else { for (uint32_t j = 0; j < 4; j++) cN[k + j] = base + k + j < nb ? cw[base + k + j] : 0; }
    }
#pragma unroll
    for (uint32_t k = 0; k < SCAN_PER; k++) v += cN[k];
    uint32_t total;
    uint32_t excl = tile_sums[(size_t)set * tiles_ps + tile] + scan_block_incl(v, warp_tot, &total) - v;
#pragma unroll
    for (uint32_t k = 0; k < SCAN_PER; k++) {
        const uint32_t idx = base + k;
        if (idx < nb) {
            ow[idx] = excl;
            if (cursor) kw[idx] = excl;
            if (big_count && cN[k] && (excl + cN[k] - 1) / L - excl / L + 1 > BIG_SPAN) big_list[atomicAdd(big_count, 1u)] = set * nb + idx;
        }
        excl += cN[k];
    }
}

// K3 (windowed): scatter point indices (digit sign in bit 31) into their bucket's range.  blockIdx.y = window, so one window's
// 4n-byte output range is being filled at a time and stays in L2 while it is written.
static __global__ void __launch_bounds__(256) k_scatter(const uint32_t *__restrict__ codes, uint32_t n, uint32_t nb, uint32_t w0,
                                                        uint32_t *__restrict__ cursor, uint32_t *__restrict__ sorted) {
    const uint32_t w = w0 + blockIdx.y;             // the pipelined driver launches one window at a time (gridDim.y = 1)
    const uint32_t *dw = codes + (size_t)w * n;
    uint32_t *kw = cursor + (size_t)w * nb;
    uint32_t *sw = sorted + (size_t)w * n;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t code = dw[i];
        if (code == CODE_SKIP32) continue;
        PB_CHECK((code & 0x7FFFFFFFu) < nb, "k_scatter: bucket");
        const uint32_t pos = atomicAdd(&kw[code & 0x7FFFFFFFu], 1u);
        PB_CHECK(pos < n, "k_scatter: position in the sorted list");
        sw[pos] = i | (code & 0x80000000u);
    }
}

// K1 / K3 of the table plan (one bucket set, `ranges` bucket ranges of 2^log2_span buckets).  The scatter used to be a filter: every range pass
// read ALL codes and kept one in `ranges` -- 8 x 201 M code inspections at 2^24, which cost the overlapped accumulation kernel 3-12 % of its issue
// slots once the two were pipelined.  Now K1 writes the codes of each 256-scalar tile already GROUPED BY RANGE (a counting sort of at most
// 256 * W codes in shared memory, no atomics: per-thread counts in a conflict-free [range][thread] matrix, one row scan per range), with the
// group boundaries in a small header, so the scatter of range r reads exactly its own codes: contiguous slices, every code a hit.
//   code = sign << 31 | (w * 256 + scalar's index in the tile) << 18 | (bucket & (2^log2_span - 1)),  log2_span <= 18, W <= 32
static constexpr uint32_t TILE_PTS = 256;
static constexpr uint32_t TILE_ID_SHIFT = 18;
template <class C>
__global__ void __launch_bounds__(TILE_PTS) k_digits_tiled(const uint32_t *__restrict__ scalars, uint32_t n, uint32_t c, uint32_t W, uint32_t wide,
                                                           uint32_t log2_span, uint32_t ranges, uint32_t class_log2, uint32_t class_index,
                                                           uint32_t *__restrict__ codes, uint16_t *__restrict__ heads, uint32_t *__restrict__ counts) {
    using Fr = typename C::Fr;
    extern __shared__ __align__(16) uint32_t sh_dyn[];
    // every array is [row][thread]: bank = thread index, so all accesses but the final placement are conflict-free
    uint32_t *cnt = sh_dyn;                               // [ranges][256] digits per range and thread, then their exclusive scan
    uint32_t *base = cnt + ranges * TILE_PTS;             // [ranges + 1] range starts within the tile
    uint32_t *stage = base + 40;                          // [W][256] sign << 31 | bucket, or all-ones: the digits, extracted once
    uint32_t *out = stage + W * TILE_PTS;                 // [256 * W] the tile's codes grouped by range
    uint32_t *limbs = out;                                // [9][256] canonical scalar + a zero row (dead before `out` is written; W >= 9 for these curves)
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t span_mask = (1u << log2_span) - 1, class_mask = (1u << class_log2) - 1;
    const uint32_t tiles = (n + TILE_PTS - 1) / TILE_PTS;
    for (uint32_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const uint32_t i = tile * TILE_PTS + tid;
        for (uint32_t r = 0; r < ranges; r++) cnt[r * TILE_PTS + tid] = 0;
        {
            Fr s = Fr::zero();
            if (i < n) s = Fr::load(scalars + (size_t)i * Fr::N).from_mont();     // canonical integer; the input is left untouched (zero: no digits)
#pragma unroll
            for (int l = 0; l < Fr::N; l++) limbs[l * TILE_PTS + tid] = s.l[l];
            limbs[Fr::N * TILE_PTS + tid] = 0;
        }
        // signed-digit recoding (see for_each_digit): the window is cut out of two adjacent limbs read back from shared memory -- a dynamic limb
        // index without local memory, ~25 instructions per digit where the bit-buffer loop took ~75 (ncu: the kernel was issue-bound)
        uint32_t o = 0, carry = 0;
#pragma unroll 1
        for (uint32_t w = 0; w < W; w++) {
            const uint32_t cw = w < wide ? c : c - 1, limb = o >> 5;
            const uint32_t bits = __funnelshift_r(limbs[limb * TILE_PTS + tid], limbs[(limb + 1) * TILE_PTS + tid], o & 31) & ((1u << cw) - 1);
            o += cw;
            const uint32_t v = bits + carry;
            uint32_t mag = v, neg = 0;
            carry = 0;
            if (w + 1 < W && v > (1u << (cw - 1))) { mag = (1u << cw) - v; neg = 1; carry = 1; }   // the top window is never recoded
            uint32_t code = CODE_SKIP32;
            if (mag && ((mag - 1) & class_mask) == class_index) {      // bucket-class shard: only the buckets congruent to class_index survive, renumbered
                const uint32_t b = (mag - 1) >> class_log2;
                PB_CHECK((b >> log2_span) < ranges, "k_digits_tiled: bucket");
                atomicAdd(&counts[b], 1u);
                cnt[(b >> log2_span) * TILE_PTS + tid]++;
                code = (neg << 31) | b;
            }
            stage[w * TILE_PTS + tid] = code;
        }
        __syncthreads();
        // exclusive scan of every range's row over the 256 threads: a warp per row, 8 cells per lane
        for (uint32_t r = warp; r < ranges; r += TILE_PTS / 32) {
            uint4 *row = reinterpret_cast<uint4 *>(cnt + r * TILE_PTS) + lane * 2;
            uint4 a = row[0], b4 = row[1];
            const uint32_t v[8] = {a.x, a.y, a.z, a.w, b4.x, b4.y, b4.z, b4.w};
            uint32_t sum = 0, ex[8];
#pragma unroll
            for (int k = 0; k < 8; k++) { ex[k] = sum; sum += v[k]; }
            uint32_t incl = sum;
#pragma unroll
            for (int o2 = 1; o2 < 32; o2 <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o2); if ((int)lane >= o2) incl += t; }
            const uint32_t pre = incl - sum;
            row[0] = make_uint4(ex[0] + pre, ex[1] + pre, ex[2] + pre, ex[3] + pre);
            row[1] = make_uint4(ex[4] + pre, ex[5] + pre, ex[6] + pre, ex[7] + pre);
            if (lane == 31) base[r + 1] = incl;           // the row's total for now
        }
        __syncthreads();
        if (warp == 0) {                                   // range starts within the tile (ranges <= 32)
            const uint32_t v = lane < ranges ? base[lane + 1] : 0;
            uint32_t incl = v;
#pragma unroll
            for (int o2 = 1; o2 < 32; o2 <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o2); if ((int)lane >= o2) incl += t; }
            if (lane < ranges) base[lane + 1] = incl;
            if (lane == 0) base[0] = 0;
        }
        __syncthreads();
#pragma unroll 1
        for (uint32_t w = 0; w < W; w++) {
            const uint32_t code = stage[w * TILE_PTS + tid];
            if (code == CODE_SKIP32) continue;
            const uint32_t b = code & 0x7FFFFFFFu, r = b >> log2_span;
            const uint32_t pos = base[r] + cnt[r * TILE_PTS + tid]++;
            PB_CHECK(pos < TILE_PTS * W && r < ranges, "k_digits_tiled: tile position");
            out[pos] = (code & 0x80000000u) | ((w * TILE_PTS + tid) << TILE_ID_SHIFT) | (b & span_mask);
        }
        __syncthreads();
        const uint32_t total = base[ranges];
        uint32_t *dst = codes + (size_t)tile * TILE_PTS * W;
        for (uint32_t k = tid; k < total; k += TILE_PTS) dst[k] = out[k];
        if (tid <= ranges) heads[(size_t)tile * (ranges + 1) + tid] = (uint16_t)base[tid];
        __syncthreads();
    }
}

// K3 (table plan): bucket range(s) first_range + blockIdx.y; a warp per tile walks the tile's slice for the range -- every code is placed.
static __global__ void __launch_bounds__(256) k_scatter_tiled(const uint32_t *__restrict__ codes, const uint16_t *__restrict__ heads, uint32_t nq, uint32_t W,
                                                              uint32_t n_total, uint32_t point0, uint32_t log2_span, uint32_t ranges, uint32_t first_range,
                                                              uint32_t capacity, uint32_t *__restrict__ cursor, uint32_t *__restrict__ sorted) {
    const uint32_t r = first_range + blockIdx.y;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warp_g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t tiles = (nq + TILE_PTS - 1) / TILE_PTS;
    const uint32_t span_mask = (1u << log2_span) - 1, bucket0 = r << log2_span;
    for (uint32_t tile = warp_g; tile < tiles; tile += nwarps) {
        const uint16_t *h = heads + (size_t)tile * (ranges + 1) + r;
        const uint32_t h0 = h[0], h1 = h[1];
        PB_CHECK(h0 <= h1 && h1 <= TILE_PTS * W && r < ranges, "k_scatter_tiled: tile header");
        const uint32_t *src = codes + (size_t)tile * TILE_PTS * W;
        const uint32_t entry0 = point0 + tile * TILE_PTS;
        auto place = [&](uint32_t code) {
            const uint32_t id = (code >> TILE_ID_SHIFT) & 0x1FFFu;
            const uint32_t entry = (id >> 8) * n_total + entry0 + (id & 255u);
            PB_CHECK((id >> 8) < W && entry0 + (id & 255u) < point0 + nq, "k_scatter_tiled: code");
            const uint32_t pos = atomicAdd(&cursor[bucket0 | (code & span_mask)], 1u);
            PB_CHECK(pos < capacity, "k_scatter_tiled: position in the sorted list");
            sorted[pos] = entry | (code & 0x80000000u);
        };
        uint32_t k = h0 + lane;
        for (; k + 96 < h1; k += 128) {                    // four independent returning atomics in flight per lane
            const uint32_t c0 = __ldg(src + k), c1 = __ldg(src + k + 32), c2 = __ldg(src + k + 64), c3 = __ldg(src + k + 96);
            place(c0); place(c1); place(c2); place(c3);
        }
        for (; k < h1; k += 32) place(__ldg(src + k));
    }
}

// K4: bucket accumulation.  Thread (w, s) owns sorted entries [s*L, (s+1)*L) of bucket set w -- a fixed amount
// of work whatever the bucket sizes are -- and emits one partial sum per bucket it touches into slot
// (s + bucket), which is unique and makes a bucket's partials contiguous.
// IMAD-bound: 8 products + 2 squares = 1314 IMAD-class instructions per entry (8 limbs); 4 B index + 64 B (96 B) gathered point per entry.
template <class C>
PB_DEV void accumulate_segment(const uint8_t *__restrict__ bases, const uint32_t *__restrict__ sw, const uint32_t *__restrict__ ow,
                               uint32_t b_lo, uint32_t b_hi, uint32_t start, uint32_t end, uint8_t *__restrict__ slot_s) {
    using Fq = typename C::Fq;
    using Pt = Xyzz<Fq>;
    using Af = Affine<Fq>;
    uint32_t lo = b_lo, hi = b_hi;     // largest b in [b_lo, b_hi) with offsets[b] <= start
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (__ldg(ow + mid) <= start) lo = mid; else hi = mid;
    }
    uint32_t b = lo;
    uint32_t next_bd = __ldg(ow + b + 1);

    Pt acc = Pt::identity();
    uint32_t e_next = __ldg(sw + start);
    Af p_next = Af::load_gather(bases + (size_t)(e_next & 0x7FFFFFFFu) * Af::BYTES);
#pragma unroll 1
    for (uint32_t pos = start; pos < end; pos++) {
        const uint32_t e = e_next;
        Af p = p_next;
        if (pos + 1 < end) {           // prefetch the next point while this one is being added
            e_next = __ldg(sw + pos + 1);
            p_next = Af::load_gather(bases + (size_t)(e_next & 0x7FFFFFFFu) * Af::BYTES);
        }
        if (pos >= next_bd) {          // bucket boundary: emit the finished partial, skip empty buckets
            acc.store(slot_s + (size_t)b * Pt::BYTES);
            acc = Pt::identity();
            do { b++; PB_CHECK(b < b_hi, "accumulate: bucket walk"); next_bd = __ldg(ow + b + 1); } while (pos >= next_bd);
        }
        if (p.is_identity()) continue; // affine identity <=> x == 0 (affine.cuh:72-75)
        if (e >> 31) p.y = p.y.neg();
        acc.madd(p.x, p.y);
    }
    acc.store(slot_s + (size_t)b * Pt::BYTES);
}

template <class C>
PB_DEV void accumulate_body(const uint8_t *__restrict__ bases, const uint32_t *__restrict__ sorted,
                            const uint32_t *__restrict__ offsets, uint32_t stride, uint32_t nb, uint32_t L,
                            uint32_t segs_pw, uint32_t W, uint32_t w0, uint8_t *__restrict__ slots) {
    using Pt = Xyzz<typename C::Fq>;
    const uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (uint64_t)W * segs_pw) return;
    const uint32_t w = w0 + (uint32_t)(gid / segs_pw), s = (uint32_t)(gid % segs_pw);      // bucket sets w0 .. w0 + W - 1
    const uint32_t *ow = offsets + (size_t)w * (nb + 1);
    const uint32_t cnt = __ldg(ow + nb);
    const uint32_t start = s * L;
    if (start >= cnt) return;
    accumulate_segment<C>(bases, sorted + (size_t)w * stride, ow, 0, nb, start, min(start + L, cnt),
                          slots + ((size_t)w * ((size_t)segs_pw + nb) + s) * Pt::BYTES);
}

// K4, one bucket range of a single (folded) bucket set at a time -- the pipelined driver scatters the next range (L2-atomic bound, few
// registers, higher-priority stream) while this one is being accumulated (integer-pipe bound).  The ranges are processed from the TOP bucket
// range downwards and range [b_lo, b_hi) owns the segments whose FIRST entry lies in it: a segment that runs on into higher ranges finds them
// scattered already.  Grid-stride, so the grid is only a hint.
template <class C>
PB_DEV void accumulate_range_body(const uint8_t *__restrict__ bases, const uint32_t *__restrict__ sorted, const uint32_t *__restrict__ offsets,
                                  uint32_t nb, uint32_t L, uint32_t segs_ps, uint32_t b_lo, uint32_t b_hi, uint8_t *__restrict__ slots) {
    using Pt = Xyzz<typename C::Fq>;
    const uint32_t cnt = __ldg(offsets + nb);
    const uint32_t lo = __ldg(offsets + b_lo), hi = min(__ldg(offsets + b_hi), cnt);
#pragma unroll 1
    for (uint64_t s = (uint64_t)((lo + L - 1) / L) + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; s < segs_ps; s += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t start = (uint32_t)s * L;
        if (start >= hi) break;
        accumulate_segment<C>(bases, sorted, offsets, b_lo, nb, start, min(start + L, cnt), slots + (size_t)s * Pt::BYTES);
    }
}

template <class C>
__global__ void __launch_bounds__(ACC_THREADS) k_accumulate(const uint8_t *__restrict__ bases, const uint32_t *__restrict__ sorted,
                                                           const uint32_t *__restrict__ offsets, uint32_t stride, uint32_t nb, uint32_t L,
                                                           uint32_t segs_pw, uint32_t W, uint32_t w0, uint8_t *__restrict__ slots) {
    accumulate_body<C>(bases, sorted, offsets, stride, nb, L, segs_pw, W, w0, slots);
}
// 12-limb fields: 64-thread CTAs, five per SM (204 registers) -- ten warps per SM instead of the eight that 222 registers allow
template <class C>
__global__ void __launch_bounds__(64, 5) k_accumulate_wide(const uint8_t *__restrict__ bases, const uint32_t *__restrict__ sorted,
                                                          const uint32_t *__restrict__ offsets, uint32_t stride, uint32_t nb, uint32_t L,
                                                          uint32_t segs_pw, uint32_t W, uint32_t w0, uint8_t *__restrict__ slots) {
    accumulate_body<C>(bases, sorted, offsets, stride, nb, L, segs_pw, W, w0, slots);
}

template <class C>
__global__ void __maxnreg__(PANDA_ACC_MAXNREG) k_accumulate_range(const uint8_t *__restrict__ bases, const uint32_t *__restrict__ sorted,
                                                                 const uint32_t *__restrict__ offsets, uint32_t nb, uint32_t L, uint32_t segs_ps,
                                                                 uint32_t b_lo, uint32_t b_hi, uint8_t *__restrict__ slots) {
    accumulate_range_body<C>(bases, sorted, offsets, nb, L, segs_ps, b_lo, b_hi, slots);
}
template <class C>
__global__ void __launch_bounds__(64, 5) k_accumulate_range_wide(const uint8_t *__restrict__ bases, const uint32_t *__restrict__ sorted,
                                                                const uint32_t *__restrict__ offsets, uint32_t nb, uint32_t L, uint32_t segs_ps,
                                                                uint32_t b_lo, uint32_t b_hi, uint8_t *__restrict__ slots) {
    accumulate_range_body<C>(bases, sorted, offsets, nb, L, segs_ps, b_lo, b_hi, slots);
}

template <class F>
__device__ __noinline__ void add_cold(Xyzz<F> &a, const Xyzz<F> &b);      // defined with the stitching kernels below

// K5: bucket combine + first level of the running-sum reduction.  Thread (w, t) folds buckets
// [t*m, (t+1)*m) of logical set w from the top: B_j = sum of its partial slots over all `merge` chunks (physical set
// q * W + w holds chunk q's partials), run += B_j, tri += run.
// Emits Lc = sum_i (i+1) * B_{t*m+i} and Rc = sum_i B_{t*m+i}.
// The chain per thread is: two offsets -> the bucket's partial slots (128 B each, DRAM) -> a point addition.  With 8 warps per SM (200+ registers)
// nothing hides those dependent loads, so the walk is software-pipelined without spending registers on it: the offsets of the next
// (bucket, chunk) pair are loaded while the current pair is being added, and the next slot is pulled into L2 / L1 with a prefetch.
PB_DEV void prefetch_line(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

template <class C>
__global__ void __launch_bounds__(RED_THREADS) k_bucket_reduce(const uint8_t *__restrict__ slots, const uint32_t *__restrict__ offsets,
                                                              uint32_t nb, uint32_t L, uint32_t segs_pw, uint32_t W, uint32_t merge, uint32_t m,
                                                              uint32_t chunks_pw, uint8_t *__restrict__ chunks) {
    using Fq = typename C::Fq;
    using Pt = Xyzz<Fq>;
    const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= W * chunks_pw) return;
    const uint32_t w = gid / chunks_pw, t = gid % chunks_pw;
    const size_t set_stride = ((size_t)segs_pw + nb) * Pt::BYTES;
    // (bucket, chunk) pair -> its slots [s0, s1] (s0 > s1: none) and the address of slot 0 of that bucket
    auto bounds = [&](uint32_t i, uint32_t q, uint32_t &s0, uint32_t &s1, const uint8_t *&base) {
        const uint32_t j = t * m + i;
        const size_t set = (size_t)q * W + w;
        const uint32_t *ow = offsets + set * (nb + 1);
        const uint32_t o0 = __ldg(ow + j), o1 = __ldg(ow + j + 1);
        base = slots + set * set_stride + (size_t)j * Pt::BYTES;
        if (o1 > o0) {
            s0 = o0 / L;
            s1 = (o1 - 1) / L;
            if (s1 - s0 + 1 > BIG_SPAN) s1 = s0;           // already folded into its first slot by k_reduce_big
        } else { s0 = 1; s1 = 0; }
    };
    Pt run = Pt::identity(), tri = Pt::identity();
    uint32_t i = m - 1, q = 0, s0, s1;
    const uint8_t *base;
    bounds(i, q, s0, s1, base);
    if (s0 <= s1) prefetch_line(base + (size_t)s0 * Pt::BYTES);
#pragma unroll 1
    for (uint32_t step = m * merge; step-- > 0;) {
        // the pair after this one (its offsets travel while this pair is being added)
        uint32_t ni = i, nq = q + 1, ns0 = 1, ns1 = 0;
        const uint8_t *nbase = base;
        if (nq == merge) { nq = 0; ni = i - 1; }
        if (step) bounds(ni, nq, ns0, ns1, nbase);
#pragma unroll 1
        for (uint32_t sg = s0; sg <= s1; sg++) {
            if (sg < s1) prefetch_line(base + (size_t)(sg + 1) * Pt::BYTES);
            else if (ns0 <= ns1) prefetch_line(nbase + (size_t)ns0 * Pt::BYTES);
            Pt part = Pt::load(base + (size_t)sg * Pt::BYTES);
            if constexpr (Fq::N > 8) add_cold(run, part); else run.add(part);   // 12 limbs: out of line (255 registers + spills otherwise)
        }
        if (s0 > s1 && ns0 <= ns1) prefetch_line(nbase + (size_t)ns0 * Pt::BYTES);
        if (q + 1 == merge) { if constexpr (Fq::N > 8) add_cold(tri, run); else tri.add(run); }
        i = ni; q = nq; s0 = ns0; s1 = ns1; base = nbase;
    }
    uint8_t *out = chunks + (size_t)gid * 2 * Pt::BYTES;
    tri.store(out);
    run.store(out + Pt::BYTES);
}

// ---- cold-path wrappers: the stitching kernels below are latency-bound one-offs; keeping the group law out of
// line there keeps code size (and compile time) down without touching the hot accumulate / bucket kernels.
// 8-limb fields: the latency-oriented forms (ec.cuh add_ilp / dbl_ilp); with 12 limbs six products side by side do not fit the register file.
template <class F>
__device__ __noinline__ void add_cold(Xyzz<F> &a, const Xyzz<F> &b) { if constexpr (F::N <= 8) a.add_ilp(b); else a.add(b); }
template <class F>
__device__ __noinline__ void dbl_cold(Xyzz<F> &a) { if constexpr (F::N <= 8) a = a.dbl_ilp(); else a = a.dbl(); }

// ---- warp-shuffle helpers -------------------------------------------------------------------------

template <class F>
PB_DEV Xyzz<F> shfl_down_pt(const Xyzz<F> &p, int delta) {
    Xyzz<F> r;
#pragma unroll
    for (int i = 0; i < F::N; i++) {
        r.x.l[i] = __shfl_down_sync(0xffffffffu, p.x.l[i], delta);
        r.y.l[i] = __shfl_down_sync(0xffffffffu, p.y.l[i], delta);
        r.zz.l[i] = __shfl_down_sync(0xffffffffu, p.zz.l[i], delta);
        r.zzz.l[i] = __shfl_down_sync(0xffffffffu, p.zzz.l[i], delta);
    }
    return r;
}

// K4b: oversized buckets (more than BIG_SPAN partial slots: skewed scalars, or a short top window of the windowed plan) are folded
// into their first slot before the bucket reduction.  One warp per bucket (lanes stride over the slots, shuffle tree);
// buckets with more than BIG_WARP_SPAN slots get the whole CTA (strided serial sums, a shuffle tree per warp, one across warps).
static constexpr uint32_t BIG_WARP_SPAN = 2048;
template <class C>
__global__ void __launch_bounds__(BIG_THREADS) k_reduce_big(uint8_t *__restrict__ slots, const uint32_t *__restrict__ offsets,
                                                           const uint32_t *__restrict__ big_count, const uint32_t *__restrict__ big_list,
                                                           uint32_t nb, uint32_t L, uint32_t segs_pw) {
    using Fq = typename C::Fq;
    using Pt = Xyzz<Fq>;
    constexpr uint32_t WARPS = BIG_THREADS / 32;
    __shared__ uint4 sh_raw[WARPS * Pt::BYTES / 16];
    uint8_t *sh = reinterpret_cast<uint8_t *>(sh_raw);
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t count = *big_count;
    for (uint32_t e0 = blockIdx.x * WARPS; e0 < count; e0 += gridDim.x * WARPS) {
        // ---- warp per bucket
        {
            const uint32_t e = e0 + warp;
            if (e < count) {
                const uint32_t id = big_list[e], w = id / nb, j = id % nb;
                const uint32_t *ow = offsets + (size_t)w * (nb + 1);
                uint8_t *slot_w = slots + (size_t)w * ((size_t)segs_pw + nb) * Pt::BYTES;
                const uint32_t s0 = ow[j] / L, s1 = (ow[j + 1] - 1) / L;
                if (s1 - s0 + 1 <= BIG_WARP_SPAN) {
                    Pt acc = Pt::identity();
#pragma unroll 1
                    for (uint32_t s = s0 + lane; s <= s1; s += 32) {
                        Pt part = Pt::load(slot_w + ((size_t)s + j) * Pt::BYTES);
                        add_cold(acc, part);
                    }
                    __syncwarp();                          // every lane has read its slots (slot s0 + j is overwritten below)
#pragma unroll 1
                    for (int o = 16; o > 0; o >>= 1) {
                        Pt t = shfl_down_pt(acc, o);
                        if (lane + o < 32) add_cold(acc, t);
                    }
                    if (lane == 0) acc.store(slot_w + ((size_t)s0 + j) * Pt::BYTES);
                }
            }
        }
        // ---- whole CTA for the giants (block-uniform conditions)
#pragma unroll 1
        for (uint32_t k = 0; k < WARPS; k++) {
            const uint32_t e = e0 + k;
            if (e >= count) break;
            const uint32_t id = big_list[e], w = id / nb, j = id % nb;
            const uint32_t *ow = offsets + (size_t)w * (nb + 1);
            uint8_t *slot_w = slots + (size_t)w * ((size_t)segs_pw + nb) * Pt::BYTES;
            const uint32_t s0 = ow[j] / L, s1 = (ow[j + 1] - 1) / L;
            if (s1 - s0 + 1 <= BIG_WARP_SPAN) continue;
            Pt acc = Pt::identity();
#pragma unroll 1
            for (uint32_t s = s0 + tid; s <= s1; s += BIG_THREADS) {
                Pt part = Pt::load(slot_w + ((size_t)s + j) * Pt::BYTES);
                add_cold(acc, part);
            }
#pragma unroll 1
            for (int o = 16; o > 0; o >>= 1) {
                Pt t = shfl_down_pt(acc, o);
                if (lane + o < 32) add_cold(acc, t);
            }
            __syncthreads();                               // every read of slot s0 + j above is done; smem is free
            if (lane == 0) acc.store(sh + (size_t)warp * Pt::BYTES);
            __syncthreads();
            if (warp == 0) {
                Pt v = lane < WARPS ? Pt::load(sh + (size_t)lane * Pt::BYTES) : Pt::identity();
#pragma unroll 1
                for (int o = 4; o > 0; o >>= 1) {
                    Pt t = shfl_down_pt(v, o);
                    if (lane + o < WARPS) add_cold(v, t);
                }
                if (lane == 0) v.store(slot_w + ((size_t)s0 + j) * Pt::BYTES);
            }
        }
    }
}

template <class F>
PB_DEV Xyzz<F> mul_pow2(Xyzz<F> p, uint32_t log2k) {
#pragma unroll 1
    for (uint32_t i = 0; i < log2k; i++) dbl_cold(p);
    return p;
}

// Over the 32 lanes of a warp, with lane l holding (P_l, R_l):
//   lane 0 returns  sumP = sum_l P_l,  sumR = sum_l R_l,  wR = sum_l l * R_l
// (suffix running sums by shuffle: sum_l l*R_l = sum_{l>=1} sum_{j>=l} R_j).
template <class F>
PB_DEV void warp_running_sums(Xyzz<F> &P, Xyzz<F> &R, Xyzz<F> &wR, uint32_t lane, uint32_t active = 32) {
    // `active` (a power of two): lanes at and above it hold the identity, so the scans and trees can stop early -- each step is a
    // dependent point addition, the unit these latency-bound kernels are made of
#pragma unroll 1
    for (uint32_t o = 1; o < active; o <<= 1) {         // inclusive suffix scan of R
        Xyzz<F> t = shfl_down_pt(R, (int)o);
        if (lane + o < active) add_cold(R, t);
    }
    wR = lane ? R : Xyzz<F>::identity();
#pragma unroll 1
    for (uint32_t o = active >> 1; o > 0; o >>= 1) {    // tree sums
        Xyzz<F> t = shfl_down_pt(wR, (int)o);
        if (lane + o < active) add_cold(wR, t);
        Xyzz<F> u = shfl_down_pt(P, (int)o);
        if (lane + o < active) add_cold(P, u);
    }
}

// K6: group reduce.  CTA (g, set) stitches the chunk sums [g*cpg, (g+1)*cpg) of one bucket set:
//   S_g = sum_t Lc_t + m * sum_t t_local * Rc_t,   Rtot_g = sum_t Rc_t
// Thread-serial running sums over q items, then warp-shuffle suffix scans across lanes and across warps.
template <class C>
__global__ void __launch_bounds__(WIN_THREADS) k_group_reduce(const uint8_t *__restrict__ chunks, uint32_t chunks_ps, uint32_t cpg, uint32_t log2m,
                                                             uint8_t *__restrict__ gsums) {
    using Fq = typename C::Fq;
    using Pt = Xyzz<Fq>;
    __shared__ uint4 sh_raw[(WIN_THREADS / 32) * 2 * Pt::BYTES / 16];
    uint8_t *sh = reinterpret_cast<uint8_t *>(sh_raw);
    const uint32_t g = blockIdx.x, set = blockIdx.y, groups = gridDim.x;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t q = cpg > WIN_THREADS ? cpg / WIN_THREADS : 1;     // items per thread (power of two)
    uint32_t log2q = 0; while ((1u << log2q) < q) log2q++;
    const uint8_t *cw = chunks + ((size_t)set * chunks_ps + (size_t)g * cpg) * 2 * Pt::BYTES;

    // thread-serial part: P = sum Lc + m * sum_i i*Rc_i,  R = sum Rc   over this thread's q items
    Pt A = Pt::identity(), run = Pt::identity(), tri = Pt::identity();
#pragma unroll 1
    for (uint32_t i = q; i-- > 0;) {
        const uint32_t t = tid * q + i;
        if (t < cpg) {
            Pt lc = Pt::load(cw + (size_t)t * 2 * Pt::BYTES);
            Pt rc = Pt::load(cw + (size_t)t * 2 * Pt::BYTES + Pt::BYTES);
            add_cold(A, lc);
            add_cold(tri, run);
            add_cold(run, rc);
        }
    }
    Pt P = A;
    if (q > 1) add_cold(P, mul_pow2(tri, log2m));
    Pt R = run, wR;
    // S_g = sum_j P_j + (m*q) * sum_j j * R_j        (j = thread index)
    warp_running_sums(P, R, wR, lane);
    if (lane == 0) {
        // sum_j j*R_j over the block = sum_warp ( wR_warp + 32*warp*R_warp )
        add_cold(P, mul_pow2(wR, log2m + log2q));
        P.store(sh + (size_t)warp * 2 * Pt::BYTES);
        R.store(sh + (size_t)warp * 2 * Pt::BYTES + Pt::BYTES);
    }
    __syncthreads();
    if (warp == 0) {
        Pt P2 = Pt::identity(), R2 = Pt::identity();
        if (lane < WIN_THREADS / 32) {
            P2 = Pt::load(sh + (size_t)lane * 2 * Pt::BYTES);
            R2 = Pt::load(sh + (size_t)lane * 2 * Pt::BYTES + Pt::BYTES);
        }
        Pt wR2;
        warp_running_sums(P2, R2, wR2, lane, WIN_THREADS / 32);
        if (lane == 0) {
            add_cold(P2, mul_pow2(wR2, log2m + log2q + 5));
            uint8_t *out = gsums + ((size_t)set * groups + g) * 2 * Pt::BYTES;
            P2.store(out);
            R2.store(out + Pt::BYTES);
        }
    }
}

// K7: one warp.  Per bucket set: S_set = sum_g S_g + (m * cpg) * sum_g g * Rtot_g (one more shuffle stitch over <= 32 groups; with
// several sets or more groups a second k_group_reduce level has done this already and groups == 1 here);
// then Horner over the sets (windowed mode: (W-1)*c dependent Jacobian doublings, inherent to the window method; folded
// mode: a single set, no doublings), conversion to the reference's result coordinates, canonical store.
template <class C>
__global__ void __launch_bounds__(32) k_final(const uint8_t *__restrict__ gsums, uint32_t sets, uint32_t groups, uint32_t log2_group_unit, uint32_t c,
                                              uint32_t class_log2, uint32_t class_index, int coord, uint8_t *__restrict__ result) {
    using Fq = typename C::Fq;
    using Pt = Xyzz<Fq>;
    const uint32_t lane = threadIdx.x;
    Pt acc = Pt::identity();
#pragma unroll 1
    for (uint32_t set = sets; set-- > 0;) {
        Pt P = Pt::identity(), R = Pt::identity(), wR;
        if (lane < groups) {
            const uint8_t *in = gsums + ((size_t)set * groups + lane) * 2 * Pt::BYTES;
            P = Pt::load(in);
            R = Pt::load(in + Pt::BYTES);
        }
        if (groups > 1) {
            uint32_t active = 1; while (active < groups) active <<= 1;
            warp_running_sums(P, R, wR, lane, active);
            if (lane == 0) add_cold(P, mul_pow2(wR, log2_group_unit));
        }
        if (lane == 0 && class_log2) {
            // bucket-class shard: local bucket q stands for bucket b = q * K + g (K = 2^class_log2, g = class_index), whose weight is
            // b + 1 = K * (q + 1) - (K - 1 - g):   S_set = K * P - (K - 1 - g) * R     (P = sum (q+1) B_q, R = sum B_q)
            P = mul_pow2(P, class_log2);
            const uint32_t k = (1u << class_log2) - 1 - class_index;
            if (k) {
                Pt kr = Pt::identity();
#pragma unroll 1
                for (int bit = (int)class_log2 - 1; bit >= 0; bit--) {
                    dbl_cold(kr);
                    if ((k >> bit) & 1) add_cold(kr, R);
                }
                kr.y = kr.y.neg();
                add_cold(P, kr);
            }
        }
        if (lane == 0) {
            if (set + 1 < sets) {          // acc = 2^c * acc + S_set
                Jacobian<Fq> j = Jacobian<Fq>::from_xyzz(acc);
#pragma unroll 1
                for (uint32_t i = 0; i < c; i++) j = j.dbl();
                acc = j.to_xyzz();
            }
            add_cold(acc, P);
        }
    }
    if (lane == 0) {
        Jacobian<Fq> j = Jacobian<Fq>::from_xyzz(acc);
        if (coord == COORD_PROJECTIVE) j = j.to_homogeneous();
        j.store_canonical(result);
    }
}

// ---- precomputed tables for reused bases ------------------------------------------------------------------------
// table[j*n + i] = 2^(o_j) * P_i (affine), j < W, o_j = total width of windows 0 .. j-1 (c*j for uniform windows): every window then
// feeds ONE bucket set and the final Horner disappears.
// One thread per point: (W-1)*c Jacobian doublings, the W-1 intermediate points kept in local memory, one shared inversion
// (Montgomery's trick) to normalise them.  Run once per cached base set (~10 MSMs worth of arithmetic).
template <class C>
__global__ void __launch_bounds__(128) k_build_table(const uint8_t *__restrict__ bases, uint32_t n, uint32_t c, uint32_t W, uint32_t wide,
                                                     uint8_t *__restrict__ table) {
    using Fq = typename C::Fq;
    using Af = Affine<Fq>;
    constexpr int MAXW = 32;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Af p = Af::load(bases + (size_t)i * Af::BYTES);
    uint32_t *t0 = reinterpret_cast<uint32_t *>(table + (size_t)i * Af::BYTES);
    p.x.store(t0); p.y.store(t0 + Fq::N);
    if (p.is_identity()) {                 // identity stays identity in every table row
        for (uint32_t j = 1; j < W; j++) {
            uint32_t *t = reinterpret_cast<uint32_t *>(table + ((size_t)j * n + i) * Af::BYTES);
            Fq::zero().store(t); Fq::zero().store(t + Fq::N);
        }
        return;
    }
    Jacobian<Fq> q; q.x = p.x; q.y = p.y; q.z = Fq::one();
    Jacobian<Fq> pts[MAXW];
    Fq prefix[MAXW];
    Fq run = Fq::one();
#pragma unroll 1
    for (uint32_t j = 1; j < W; j++) {
        const uint32_t cw = j - 1 < wide ? c : c - 1;          // width of window j-1: row j = 2^(width of the windows below) * P
#pragma unroll 1
        for (uint32_t k = 0; k < cw; k++) q = q.dbl();
        pts[j] = q;
        prefix[j] = run;                   // product of z_1 .. z_{j-1}
        run = run * q.z;
    }
    Fq inv = fe_inverse(run);              // z is never 0 here: P has odd prime order, so 2^k * P is not the identity
#pragma unroll 1
    for (uint32_t j = W - 1; j >= 1; j--) {
        const Fq zinv = inv * prefix[j];
        inv = inv * pts[j].z;
        const Fq zi2 = zinv.sqr();
        const Fq x = pts[j].x * zi2;
        const Fq y = pts[j].y * (zi2 * zinv);
        uint32_t *t = reinterpret_cast<uint32_t *>(table + ((size_t)j * n + i) * Af::BYTES);
        x.canon().store(t); y.canon().store(t + Fq::N);
    }
}

// 64-bit fingerprint of a device buffer (order-sensitive multiply-xorshift mix, combined by addition): guards the table
// cache against a caller that reuses a device pointer for different bases.
static __global__ void __launch_bounds__(256) k_fingerprint(const uint4 *__restrict__ data, size_t count16, unsigned long long *__restrict__ out) {
    unsigned long long h = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count16; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 v = __ldg(data + i);
        unsigned long long x = ((unsigned long long)v.x | ((unsigned long long)v.y << 32)) ^ (0x9E3779B97F4A7C15ull * (i + 1));
        unsigned long long y = ((unsigned long long)v.z | ((unsigned long long)v.w << 32)) + 0xD1B54A32D192ED03ull * (i + 7);
        x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 32;
        y ^= y >> 31; y *= 0x94D049BB133111EBull; y ^= y >> 29;
        h += x ^ (y * 0x2545F4914F6CDD1Dull);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) h += __shfl_down_sync(0xffffffffu, h, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, h);
}

// sum of Jacobian partials (sharded MSM): one thread, a handful of additions
template <class C>
__global__ void k_combine(const uint8_t *__restrict__ partials, uint32_t count, int coord, uint8_t *__restrict__ result) {
    using Fq = typename C::Fq;
    using Pt = Xyzz<Fq>;
    if (threadIdx.x || blockIdx.x) return;
    Pt acc = Pt::identity();
#pragma unroll 1
    for (uint32_t i = 0; i < count; i++) {
        const uint32_t *q = reinterpret_cast<const uint32_t *>(partials + (size_t)i * Jacobian<Fq>::BYTES);
        Jacobian<Fq> j;
        j.x = Fq::load_plain(q); j.y = Fq::load_plain(q + Fq::N); j.z = Fq::load_plain(q + 2 * Fq::N);
        add_cold(acc, j.to_xyzz());
    }
    Jacobian<Fq> j = Jacobian<Fq>::from_xyzz(acc);
    if (coord == COORD_PROJECTIVE) j = j.to_homogeneous();
    j.store_canonical(result);
}

// ----------------------------------------------------------------------------------------------------
// host driver

#define PB_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "[panda-b200] CUDA error %d (%s) at %s:%d\n", (int)e_, cudaGetErrorString(e_), __FILE__, __LINE__); return e_; } } while (0)

struct StageTimer {
    cudaEvent_t ev[8];
    cudaStream_t stream;
    int k = 0;
    bool on;
    StageTimer(bool enable, cudaStream_t s) : stream(s), on(enable) { if (on) for (auto &e : ev) cudaEventCreate(&e); }
    ~StageTimer() { if (on) for (auto &e : ev) cudaEventDestroy(e); }
    void mark() { if (on && k < 8) cudaEventRecord(ev[k++], stream); }
    float ms(int i) { float t = 0; cudaEventElapsedTime(&t, ev[i], ev[i + 1]); return t; }
};

// PANDA_MSM_TRACE=1: timeline of a chunked / pipelined run (events on every stream involved, printed relative to the start of the call; the
// call then synchronises -- a diagnostic, never set on a measured path)
struct TraceLog {
    struct Mark { const char *what; uint32_t q; cudaEvent_t ev; };
    std::vector<Mark> marks;
    bool on;
    TraceLog() { static const bool enabled = [] { const char *e = getenv("PANDA_MSM_TRACE"); return e && atoi(e) != 0; }(); on = enabled; }
    void mark(const char *what, uint32_t q, cudaStream_t s) {
        if (!on) return;
        cudaEvent_t ev;
        cudaEventCreate(&ev);
        cudaEventRecord(ev, s);
        marks.push_back({what, q, ev});
    }
    void report(cudaStream_t s) {
        if (!on || marks.empty()) return;
        cudaStreamSynchronize(s);
        for (auto &m : marks) { cudaEventSynchronize(m.ev); float t = 0; cudaEventElapsedTime(&t, marks[0].ev, m.ev); fprintf(stderr, "[trace] %-14s %u  %8.3f ms\n", m.what, m.q, t); }
        for (auto &m : marks) cudaEventDestroy(m.ev);
        marks.clear();
    }
};

// Runs the pipeline described by `p`.  `points` is the caller's bases (windowed) or the precomputed table (folded).
// feed == nullptr: everything on `stream`, one launch per stage.  feed != nullptr brings the library's side streams (see the comment on the
// stream roles below) and, with feed->host_scalars, scalars that are still in host memory: chunk q is uploaded on feed->copy_stream in
// sub-chunks while the chunks before it are being sorted and accumulated.
template <class C>
cudaError_t msm_pipeline_t(const MsmPlan &p, const void *points, const void *scalars, void *result, CoordType coord, cudaMemPool_t pool,
                           cudaStream_t stream, MsmStageTimes *timings, const MsmFeed *feed) {
    using Pt = Xyzz<typename C::Fq>;
    uint8_t *ws = nullptr;
    if (pool) PB_CUDA(cudaMallocFromPoolAsync((void **)&ws, p.bytes, pool, stream));
    else PB_CUDA(cudaMallocAsync((void **)&ws, p.bytes, stream));
    const size_t phys = (size_t)p.sets * p.chunks;
    uint32_t *counts = (uint32_t *)(ws + p.off_counts), *offsets = (uint32_t *)(ws + p.off_offsets), *cursor = (uint32_t *)(ws + p.off_cursor);
    uint32_t *big_counts = counts + phys * p.nb, *big_list = (uint32_t *)(ws + p.off_biglist);   // the counters are zeroed with the counts
    uint32_t *tile_sums = (uint32_t *)(ws + p.off_tiles);
    uint8_t *digits = ws + p.off_digits;
    uint32_t *sorted = (uint32_t *)(ws + p.off_sorted);
    uint8_t *slots = ws + p.off_slots, *chunks = ws + p.off_chunks, *gsums = ws + p.off_gsums;
    const uint32_t tiles_ps = (p.nb + SCAN_TILE - 1) / SCAN_TILE;
    const size_t slot_stride = ((size_t)p.segs_ps + p.nb) * Pt::BYTES;       // per physical set

    StageTimer tm(timings != nullptr && p.chunks == 1, stream);
    TraceLog trace;
    trace.mark("start", 0, stream);
    // Streams by role (feed != nullptr carries the library's per-device side streams).  The sort of a chunk (digits, scan, scatter -- L2-atomic
    // bound) runs on the high-priority stream `side[0]` (digits and scan of a single resident chunk stay on the caller's stream: nothing could
    // overlap them), the accumulations (integer-pipe bound) rotate over the caller's stream and side[1..3], so that one launch's tail overlaps
    // the next one's head and a sort never queues behind an accumulation.  Table plan with several bucket ranges: scatter and accumulation go
    // range by range (the scatter of one range beside the accumulation of another); windowed plan: window by window; several chunks
    // (streamed scalars): chunk after chunk, every chunk uploaded and recoded in sub-chunks so that its digits trail its upload closely.
    // events: `fed` orders the copy stream and the side streams against `stream` / a sub-chunk's upload against its digit kernel; `sorted_ev`
    // hands a scattered range / window / chunk to its accumulation; `aux_done` joins the accumulation streams back into `stream`
    cudaEvent_t fed = nullptr, sorted_ev = nullptr, aux_done = nullptr;
    const bool uploading = feed && feed->host_scalars;
    const bool have_side = feed && feed->aux_stream && feed->aux2_stream && feed->aux3_stream && feed->aux4_stream;
    const bool chunked = have_side && p.chunks > 1;
    const bool ranged = have_side && p.folded && p.phases > 1;                    // range by range
    const bool windowed_pipe = have_side && !p.folded && p.windows > 1 && p.chunks == 1;   // window by window
    const bool multi = chunked || ranged || windowed_pipe;                        // more than the caller's stream is in play
    cudaStream_t side0 = have_side ? feed->aux_stream : nullptr;
    cudaStream_t sdig = have_side && feed->dig_stream ? feed->dig_stream : side0;   // digits + scan of a chunked run (scatters stay on side0)
    cudaStream_t acc_streams[4] = {stream, have_side ? feed->aux2_stream : stream, have_side ? feed->aux3_stream : stream, have_side ? feed->aux4_stream : stream};
    uint32_t acc_launch = 0;                                                      // rotates the accumulation launches over acc_streams
    cudaError_t err = cudaSuccess;
    do {
        if ((err = cudaMemsetAsync(counts, 0, (phys * p.nb + p.chunks) * 4, stream)) != cudaSuccess) break;
        if (multi || uploading) {
            if ((err = cudaEventCreateWithFlags(&fed, cudaEventDisableTiming)) != cudaSuccess) break;
            if ((err = cudaEventCreateWithFlags(&sorted_ev, cudaEventDisableTiming)) != cudaSuccess) break;
            if ((err = cudaEventCreateWithFlags(&aux_done, cudaEventDisableTiming)) != cudaSuccess) break;
            // the workspace / staging buffer were allocated in stream order on `stream`: other streams may only touch them from here on
            if ((err = cudaEventRecord(fed, stream)) != cudaSuccess) break;
            if (uploading && (err = cudaStreamWaitEvent(feed->copy_stream, fed, 0)) != cudaSuccess) break;
            if (multi) {
                if ((err = cudaStreamWaitEvent(side0, fed, 0)) != cudaSuccess) break;
                if (sdig != side0 && (err = cudaStreamWaitEvent(sdig, fed, 0)) != cudaSuccess) break;
                for (int j = 1; j < 4 && err == cudaSuccess; j++) err = cudaStreamWaitEvent(acc_streams[j], fed, 0);
                if (err != cudaSuccess) break;
            }
        }
        // kernels that run beside an accumulation (every chunk but the first; every scatter range but the first) get small grids: their
        // warps mostly wait for L2 atomics, and each CTA that becomes resident displaces accumulation warps that would keep the integer pipe
        // busy.  Too few, and the sort of the next chunk is late (a CTA beside the accumulation CTAs gets a small share of the issue
        // slots).  Measured at 2^24 (profiles/r2_msm_pipeline.md): scatter 2 CTAs per SM, digits 4; the range pipeline's scatters 1.
        static const uint32_t side_ctas = [] { const char *e = getenv("PANDA_MSM_SIDE_CTAS"); const int v = e ? atoi(e) : 0; return v > 0 ? (uint32_t)v : 296u; }();
        static const uint32_t side_digit_ctas = [] { const char *e = getenv("PANDA_MSM_SIDE_DIGITS"); const int v = e ? atoi(e) : 0; return v > 0 ? (uint32_t)v : 592u; }();
        uint32_t log2_span = 0; while ((p.nb >> log2_span) > p.phases) log2_span++;        // table plan: buckets per scatter range
        const uint32_t acc_threads = C::Fq::N > 8 ? 64 : ACC_THREADS;
        for (uint32_t q = 0; q < p.chunks && err == cudaSuccess; q++) {
            cudaStream_t sq = chunked ? sdig : stream;                        // digits and scan of this chunk
            const uint32_t point0 = p.chunk_begin[q], nq = p.chunk_begin[q + 1] - point0;
            const size_t set0 = (size_t)q * p.sets;                           // first physical set of this chunk
            const uint32_t *sc = (const uint32_t *)scalars + (size_t)point0 * 8;
            uint32_t *counts_q = counts + set0 * p.nb, *offsets_q = offsets + set0 * (p.nb + 1), *cursor_q = cursor + set0 * p.nb;
            uint32_t *big_count_q = big_counts + q, *big_list_q = big_list + set0 * p.nb, *tiles_q = tile_sums + set0 * tiles_ps;
            uint8_t *codes_q = digits + (size_t)q * p.codes_stride * 4;      // folded: 32-bit codes of this chunk, tile by tile (chunks == 1 otherwise)
            uint16_t *heads_q = (uint16_t *)(ws + p.off_heads) + (size_t)q * p.heads_stride;
            uint32_t *sorted_q = sorted + set0 * p.stride;
            uint8_t *slots_q = slots + set0 * slot_stride;
            const bool beside = chunked && q > 0;                             // an earlier chunk is accumulating while this one sorts
            const uint32_t digit_cap = beside ? side_digit_ctas : 148 * 8, cta_cap = beside ? side_ctas : 148 * 8;
            const uint32_t sblocks = std::min<uint32_t>((nq + 255) / 256, cta_cap);
            const uint32_t scat_blocks = std::min<uint32_t>(((nq + TILE_PTS - 1) / TILE_PTS + 7) / 8, cta_cap);
            tm.mark();
            // ---- digits (+ upload): a streamed chunk arrives in up to eight sub-chunks of whole tiles, each recoded as soon as it is there
            const uint32_t subs = (uploading && p.folded) ? (nq >= (1u << 23) ? 8 : nq >= (1u << 20) ? 4 : 1) : 1;
            for (uint32_t j = 0; j < subs && err == cudaSuccess; j++) {
                const uint32_t t0 = (uint32_t)(((uint64_t)((nq + TILE_PTS - 1) / TILE_PTS) * j / subs) * TILE_PTS);
                const uint32_t t1 = j + 1 == subs ? nq : (uint32_t)(((uint64_t)((nq + TILE_PTS - 1) / TILE_PTS) * (j + 1) / subs) * TILE_PTS);
                if (t1 <= t0) continue;
                if (uploading) {
                    if ((err = cudaMemcpyAsync((uint8_t *)feed->dev_scalars + (size_t)(point0 + t0) * 32, (const uint8_t *)feed->host_scalars + (size_t)(point0 + t0) * 32,
                                               (size_t)(t1 - t0) * 32, cudaMemcpyHostToDevice, feed->copy_stream)) != cudaSuccess) break;
                    if ((err = cudaEventRecord(fed, feed->copy_stream)) != cudaSuccess) break;
                    trace.mark("uploaded", q, feed->copy_stream);
                    if ((err = cudaStreamWaitEvent(sq, fed, 0)) != cudaSuccess) break;
                }
                if (j == 0) trace.mark("sort begins", q, sq);
                if (p.folded) {
                    const size_t sh_bytes = ((size_t)(p.phases + 2 * p.windows) * TILE_PTS + 40) * 4;
                    if (sh_bytes > 48 * 1024 && (err = cudaFuncSetAttribute(k_digits_tiled<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh_bytes)) != cudaSuccess) break;
                    const uint32_t tile0 = t0 / TILE_PTS;
                    const uint32_t tblocks = std::min<uint32_t>((t1 - t0 + TILE_PTS - 1) / TILE_PTS, digit_cap);
                    k_digits_tiled<C><<<tblocks, TILE_PTS, sh_bytes, sq>>>(sc + (size_t)t0 * 8, t1 - t0, p.c, p.windows, p.wide, log2_span, p.phases, p.class_log2, p.class_index,
                                                                          (uint32_t *)codes_q + (size_t)tile0 * TILE_PTS * p.windows, heads_q + (size_t)tile0 * (p.phases + 1), counts_q);
                } else k_digits<C><<<sblocks, 256, 0, sq>>>(sc, nq, p.c, p.windows, p.wide, p.nb, p.class_log2, p.class_index, (uint32_t *)codes_q, counts_q);
            }
            if (err != cudaSuccess) break;
            tm.mark();
            trace.mark("digits done", q, sq);
            k_scan_tiles<<<dim3(tiles_ps, p.sets), SCAN_THREADS, 0, sq>>>(counts_q, p.nb, tiles_ps, tiles_q);
            k_scan_tops<<<p.sets, SCAN_THREADS, 0, sq>>>(tiles_q, tiles_ps, p.nb, offsets_q);
            k_scan_apply<<<dim3(tiles_ps, p.sets), SCAN_THREADS, 0, sq>>>(counts_q, tiles_q, p.nb, tiles_ps, p.seg_len, offsets_q, cursor_q, big_count_q, big_list_q);
            tm.mark();
            trace.mark("scanned", q, sq);
            // ---- scatter + accumulation
            cudaStream_t ss = multi ? side0 : stream;                          // scatters
            if (multi && ss != sq) {                                           // (single resident chunk: digits and scan ran on the caller's stream)
                if ((err = cudaEventRecord(fed, sq)) != cudaSuccess) break;
                if ((err = cudaStreamWaitEvent(ss, fed, 0)) != cudaSuccess) break;
            }
            auto hand_over = [&](cudaStream_t sa) -> cudaError_t {             // what `ss` has scattered so far may be accumulated on `sa`
                if (sa == ss) return cudaSuccess;
                cudaError_t e = cudaEventRecord(sorted_ev, ss);
                return e != cudaSuccess ? e : cudaStreamWaitEvent(sa, sorted_ev, 0);
            };
            if (ranged) {
                // The balanced windows make the ranges uneven: the (W - wide) narrow windows only reach the lower half of the buckets, so with
                // uniform scalars a lower-half range holds (2 W - wide) / wide times the entries of an upper-half one (7 x at 2^24).  The ranges
                // are processed from the top one downwards (k_accumulate_range's ownership rule): the small ones go first, the exposed scatter is
                // a small one, and every launch gets the grid its expected share asks for -- the kernel is grid-stride, so other digit
                // distributions only cost balance, not correctness.
                const uint32_t span = 1u << log2_span, ranges = p.phases;
                const uint32_t chunk_segs = ((uint32_t)(((uint64_t)nq * p.windows) >> p.class_log2) + p.seg_len - 1) / p.seg_len;   // expected segments of this chunk
                const uint32_t all_blocks = (chunk_segs + acc_threads - 1) / acc_threads + 1;
                for (uint32_t k = 0; k < ranges && err == cudaSuccess; k++) {
                    const uint32_t r = ranges - 1 - k;
                    const double share = ((double)p.wide + (r < ranges / 2 || ranges == 1 ? 2.0 * (p.windows - p.wide) : 0.0)) / ((double)p.windows * ranges);
                    const uint32_t blocks = std::min<uint32_t>(all_blocks, (uint32_t)((double)all_blocks * share * 1.03) + 8);
                    const uint32_t grid = (k == 0 && q == 0) ? scat_blocks : std::min(scat_blocks, chunked ? side_ctas : side_ctas / 2);
                    k_scatter_tiled<<<dim3(grid, 1), 256, 0, ss>>>((const uint32_t *)codes_q, heads_q, nq, p.windows, p.table_n, point0, log2_span, p.phases, r, p.stride, cursor_q, sorted_q);
                    cudaStream_t sa = acc_streams[acc_launch++ & 3];
                    if ((err = hand_over(sa)) != cudaSuccess) break;
                    if (k == 0) { tm.mark(); trace.mark("first range", q, sa); }   // "scatter" = the exposed first range
                    if (C::Fq::N > 8) k_accumulate_range_wide<C><<<blocks, acc_threads, 0, sa>>>((const uint8_t *)points, sorted_q, offsets_q, p.nb, p.seg_len, p.segs_ps, r * span, (r + 1) * span, slots_q);
                    else k_accumulate_range<C><<<blocks, acc_threads, 0, sa>>>((const uint8_t *)points, sorted_q, offsets_q, p.nb, p.seg_len, p.segs_ps, r * span, (r + 1) * span, slots_q);
                }
                trace.mark("sorted", q, ss);
            } else if (windowed_pipe) {
                // windowed plan: window w+1 is scattered while window w is accumulated
                const uint32_t blocks = (p.segs_ps + acc_threads - 1) / acc_threads;
                for (uint32_t w = 0; w < p.windows && err == cudaSuccess; w++) {
                    k_scatter<<<dim3(w ? std::min(sblocks, side_ctas / 2) : sblocks, 1), 256, 0, ss>>>((const uint32_t *)codes_q, nq, p.nb, w, cursor_q, sorted_q);
                    cudaStream_t sa = acc_streams[acc_launch++ & 3];
                    if ((err = hand_over(sa)) != cudaSuccess) break;
                    if (w == 0) tm.mark();
                    if (C::Fq::N > 8) k_accumulate_wide<C><<<blocks, acc_threads, 0, sa>>>((const uint8_t *)points, sorted_q, offsets_q, p.stride, p.nb, p.seg_len, p.segs_ps, 1, w, slots_q);
                    else k_accumulate<C><<<blocks, acc_threads, 0, sa>>>((const uint8_t *)points, sorted_q, offsets_q, p.stride, p.nb, p.seg_len, p.segs_ps, 1, w, slots_q);
                }
            } else {
                if (p.folded) k_scatter_tiled<<<dim3(beside ? std::max<uint32_t>(1, scat_blocks / p.phases) : scat_blocks, p.phases), 256, 0, ss>>>((const uint32_t *)codes_q, heads_q, nq, p.windows, p.table_n, point0, log2_span, p.phases, 0, p.stride, cursor_q, sorted_q);
                else k_scatter<<<dim3(sblocks, p.windows), 256, 0, ss>>>((const uint32_t *)codes_q, nq, p.nb, 0, cursor_q, sorted_q);
                tm.mark();
                trace.mark("sorted", q, ss);
                cudaStream_t sa = multi ? acc_streams[acc_launch++ & 3] : stream;
                if ((err = hand_over(sa)) != cudaSuccess) break;
                const uint64_t threads = (uint64_t)p.sets * p.segs_ps;
                const uint32_t blocks = (uint32_t)((threads + acc_threads - 1) / acc_threads);
                // 12-limb fields: 64-thread CTAs fit 5-6 per SM (10-12 warps) where 128-thread ones fit 2 (8 warps)
                if (C::Fq::N > 8) k_accumulate_wide<C><<<blocks, acc_threads, 0, sa>>>((const uint8_t *)points, sorted_q, offsets_q, p.stride, p.nb, p.seg_len, p.segs_ps, p.sets, 0, slots_q);
                else k_accumulate<C><<<blocks, acc_threads, 0, sa>>>((const uint8_t *)points, sorted_q, offsets_q, p.stride, p.nb, p.seg_len, p.segs_ps, p.sets, 0, slots_q);
                trace.mark("accumulated", q, sa);
            }
            if (err == cudaSuccess) err = cudaGetLastError();
        }
        if (err != cudaSuccess) break;
        if (multi) {      // join the accumulation streams (every scatter was waited for by an accumulation, so this joins side0 as well)
            for (int j = 1; j < 4 && err == cudaSuccess; j++) {
                if ((err = cudaEventRecord(aux_done, acc_streams[j])) != cudaSuccess) break;
                err = cudaStreamWaitEvent(stream, aux_done, 0);
            }
            if (err != cudaSuccess) break;
        }
        for (uint32_t q = 0; q < p.chunks; q++) {
            const size_t set0 = (size_t)q * p.sets;
            k_reduce_big<C><<<148 * 2, BIG_THREADS, 0, stream>>>(slots + set0 * slot_stride, offsets + set0 * (p.nb + 1), big_counts + q, big_list + set0 * p.nb, p.nb, p.seg_len, p.segs_ps);
        }
        tm.mark();
        trace.mark("accumulated", p.chunks, stream);
        uint32_t log2m = 0; while ((1u << log2m) < p.chunk) log2m++;
        {
            const uint32_t threads = p.sets * p.chunks_ps;
            k_bucket_reduce<C><<<(threads + RED_THREADS - 1) / RED_THREADS, RED_THREADS, 0, stream>>>(slots, offsets, p.nb, p.seg_len, p.segs_ps,
                                                                                                   p.sets, p.chunks, p.chunk, p.chunks_ps, chunks);
        }
        tm.mark();
        trace.mark("buckets done", 0, stream);
        const uint32_t cpg = p.chunks_ps / p.groups;
        uint32_t log2cpg = 0; while ((1u << log2cpg) < cpg) log2cpg++;
        k_group_reduce<C><<<dim3(p.groups, p.sets), WIN_THREADS, 0, stream>>>(chunks, p.chunks_ps, cpg, log2m, gsums);
        const uint8_t *final_in = gsums;
        uint32_t final_groups = p.groups, final_unit = log2m + log2cpg;
        // second stitch level: the group sums of a set are the items of ONE more CTA (unit = m * cpg).  Needed beyond 32 groups, and worth it
        // whenever there are several bucket sets (windowed plan): the sets are stitched side by side instead of one after the other by
        // the single warp of k_final
        if (p.groups > 32 || (p.sets > 1 && p.groups > 1)) {
            uint8_t *gsums2 = gsums + (size_t)p.sets * p.groups * 2 * Pt::BYTES;
            k_group_reduce<C><<<dim3(1, p.sets), WIN_THREADS, 0, stream>>>(gsums, p.groups, p.groups, final_unit, gsums2);
            final_in = gsums2; final_groups = 1;
        }
        tm.mark();
        k_final<C><<<1, 32, 0, stream>>>(final_in, p.sets, final_groups, final_unit, p.c, p.class_log2, p.class_index, (int)coord, (uint8_t *)result);
        tm.mark();
        trace.mark("end", 0, stream);
        err = cudaGetLastError();
    } while (0);
    if (err == cudaSuccess) trace.report(stream);
    if (err != cudaSuccess && (multi || uploading)) {
        // error after work was queued on the side streams: they may still touch the workspace / the staged scalars, so `stream` (on which
        // both are freed in stream order) has to wait for them first.  On the success path the joins above did that.
        cudaGetLastError();
        cudaEvent_t join = nullptr;
        if (cudaEventCreateWithFlags(&join, cudaEventDisableTiming) == cudaSuccess) {
            if (multi && cudaEventRecord(join, side0) == cudaSuccess) cudaStreamWaitEvent(stream, join, 0);
            if (multi && sdig != side0 && cudaEventRecord(join, sdig) == cudaSuccess) cudaStreamWaitEvent(stream, join, 0);
            for (int j = 1; j < 4 && multi; j++) if (cudaEventRecord(join, acc_streams[j]) == cudaSuccess) cudaStreamWaitEvent(stream, join, 0);
            if (uploading && cudaEventRecord(join, feed->copy_stream) == cudaSuccess) cudaStreamWaitEvent(stream, join, 0);
            cudaEventDestroy(join);
        }
    }
    if (fed) cudaEventDestroy(fed);
    if (sorted_ev) cudaEventDestroy(sorted_ev);
    if (aux_done) cudaEventDestroy(aux_done);
    cudaError_t ferr = cudaFreeAsync(ws, stream);
    if (err == cudaSuccess) err = ferr;
    if (err != cudaSuccess) {
        fprintf(stderr, "[panda-b200] msm pipeline failed: %s\n", cudaGetErrorString(err));
        return err;
    }
    if (timings) {
        PB_CUDA(cudaStreamSynchronize(stream));
        if (tm.on) {
            timings->digits = tm.ms(0); timings->scan = tm.ms(1); timings->scatter = tm.ms(2); timings->accumulate = tm.ms(3);
            timings->bucket_reduce = tm.ms(4); timings->window_reduce = tm.ms(5); timings->final = tm.ms(6);
        }
    }
    return cudaSuccess;
}

template <class C>
cudaError_t msm_build_table_t(const void *bases, uint32_t n, uint32_t c, uint32_t W, uint32_t wide, void *table, cudaStream_t stream) {
    k_build_table<C><<<(n + 127) / 128, 128, 0, stream>>>((const uint8_t *)bases, n, c, W, wide, (uint8_t *)table);
    return cudaGetLastError();
}

inline cudaError_t msm_fingerprint(const void *data, size_t bytes, unsigned long long *d_out, cudaStream_t stream) {
    cudaError_t e = cudaMemsetAsync(d_out, 0, 8, stream);
    if (e != cudaSuccess) return e;
    k_fingerprint<<<148 * 8, 256, 0, stream>>>((const uint4 *)data, bytes / 16, d_out);
    return cudaGetLastError();
}

template <class C>
cudaError_t msm_combine_t(const void *partials, uint32_t count, void *result, CoordType coord, cudaStream_t stream) {
    k_combine<C><<<1, 32, 0, stream>>>((const uint8_t *)partials, count, (int)coord, (uint8_t *)result);
    return cudaGetLastError();
}

}  // namespace pb
