// msm.cu -- Pippenger MSM for sm_100a: signed-digit windows, counting sort of point indices by bucket,
// load-balanced bucket accumulation (XYZZ mixed additions), parallel running-sum bucket reduction with
// warp-shuffle stitching, on-device window combination; for bases that are reused across calls, a table of
// 2^(c*j) * P multiples folds all windows into one bucket set.  The sort is off the critical path: the digit codes leave the first
// kernel grouped by bucket range, and the scatter of one range (L2-atomic bound) runs on a side stream beside the accumulation of
// another (integer-pipe bound); the windowed plan pipelines window by window the same way.  See msm.cuh / DESIGN.md.
//
// Replaces src/cuda/core/unit/msm/msm_cuda.cuh:552-769 of the reference (kernels :148-282, :373-497 and
// the host-side Horner :59-77).  Written from scratch for B200; nothing here is derived from that code.
// This file holds the plan (window width, segment length, workspace layout), the table cache and the per-curve
// dispatch; the kernels live in msm_impl.cuh.
#include "msm.cuh"

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <mutex>
#include <vector>

namespace pb {

// per-curve instantiations (msm_bn254.cu / msm_bls12_377.cu / msm_bls12_381.cu)
cudaError_t msm_pipeline_bn254(const MsmPlan &p, const void *points, const void *scalars, void *result, CoordType coord, cudaMemPool_t pool,
                               cudaStream_t stream, MsmStageTimes *timings, const MsmFeed *feed);
cudaError_t msm_pipeline_bls12_377(const MsmPlan &p, const void *points, const void *scalars, void *result, CoordType coord, cudaMemPool_t pool,
                                   cudaStream_t stream, MsmStageTimes *timings, const MsmFeed *feed);
cudaError_t msm_pipeline_bls12_381(const MsmPlan &p, const void *points, const void *scalars, void *result, CoordType coord, cudaMemPool_t pool,
                                   cudaStream_t stream, MsmStageTimes *timings, const MsmFeed *feed);
cudaError_t msm_build_table_bls12_381(const void *bases, uint32_t n, uint32_t c, uint32_t W, uint32_t wide, void *table, cudaStream_t stream);
cudaError_t msm_combine_bls12_381(const void *partials, uint32_t count, void *result, CoordType coord, cudaStream_t stream);
cudaError_t msm_build_table_bn254(const void *bases, uint32_t n, uint32_t c, uint32_t W, uint32_t wide, void *table, cudaStream_t stream);
cudaError_t msm_build_table_bls12_377(const void *bases, uint32_t n, uint32_t c, uint32_t W, uint32_t wide, void *table, cudaStream_t stream);
cudaError_t msm_fingerprint_launch(const void *data, size_t bytes, unsigned long long *d_out, cudaStream_t stream);
cudaError_t msm_combine_bn254(const void *partials, uint32_t count, void *result, CoordType coord, cudaStream_t stream);
cudaError_t msm_combine_bls12_377(const void *partials, uint32_t count, void *result, CoordType coord, cudaStream_t stream);

// ----------------------------------------------------------------------------------------------------
// plan

static uint32_t windows_for(uint32_t bits, uint32_t c) {
    // signed digits need the top window to stay <= 2^(c-1) after the incoming carry: its bit width t must be <= c-1
    uint32_t W = (bits + c - 1) / c;
    int t = (int)bits - (int)(W - 1) * (int)c;
    if (t > (int)c - 1) W++;
    return W;
}

// Table plan: W = ceil((bits + 1) / c) windows whose widths add up to exactly bits + 1 (one spare bit for the recoding carry): the low
// `wide` windows are c bits, the others c - 1.  No window is short, so no bucket range collects a whole window's digits.
// Returns false when c - 1 would do (wide <= 0): that plan is the one of c - 1.
static bool balanced_windows(uint32_t bits, uint32_t c, uint32_t *W, uint32_t *wide) {
    const uint32_t total = bits + 1;
    *W = (total + c - 1) / c;
    const int w = (int)total - (int)(*W) * (int)(c - 1);
    if (w <= 0) return false;
    *wide = (uint32_t)w;
    return true;
}

static uint32_t pow2_floor(uint64_t v) { uint32_t r = 1; while ((uint64_t)r * 2 <= v) r *= 2; return r; }

MsmPlan msm_make_plan(CurveId curve, uint32_t n, bool folded, uint32_t c_override, uint32_t seg_override, size_t table_budget, uint32_t chunks,
                      uint32_t class_log2) {
    const uint32_t bits = curve_scalar_bits(curve);
    const size_t fq_bytes = curve_fq_bytes(curve);
    MsmPlan p{};
    p.n = n;
    p.table_n = n;
    p.folded = folded ? 1 : 0;
    p.class_log2 = class_log2;
    const uint32_t c_lo = 8, c_hi = folded ? 23 : 20;
    uint32_t best_c = 0;
    double best = 1e300;
    for (uint32_t c = c_lo; c <= c_hi; c++) {
        if (c_override && c != c_override) continue;
        uint32_t Wi = windows_for(bits, c), widei = Wi;
        if (folded && !balanced_windows(bits, c, &Wi, &widei)) continue;
        const double W = Wi, nb = (double)(1u << (c - 1));
        if (W > 32) continue;
        if (folded) {
            if ((double)n * W >= 2147483648.0) continue;                            // table index + sign must fit 32 bits
            if ((double)n * W * 2 * fq_bytes > (double)table_budget) continue;
        }
        // modmul counts: mixed add 10; bucket combine + running sums ~3.5 full adds of 14 per bucket (the reduction kernels
        // run at about half the accumulate kernel's rate, hence the factor 2); folded mode has a single bucket set
        // A narrow top window (t bits) sends n digits to 2^t buckets: their counters are hot L2-atomic addresses in the histogram and
        // scatter kernels (measured at n = 2^21, t = 7: +0.6 ms, i.e. ~20 modmul-equivalents per scalar), ~2500 / 2^t per scalar.
        const int t = (int)bits - (int)(W - 1) * (int)c;
        const double hot = !folded && t > 0 && t < 20 ? (double)n * 2500.0 / (double)(1u << t) : 0.0;   // the table plan has no short window
        // Table plan, measured on the final kernels (profiles/r2_plan_sweep.md): bucket reduction + stitching cost ~0.55 ms + 1.07 ns per bucket
        // beyond the first 2^17, i.e. 72 modmul-equivalents per bucket at the accumulation's 14.8 ps per product (2^21 points: c = 20 / W = 13
        // 5.36 ms against 5.52 for c = 19 / W = 14, which the flat 98 per bucket preferred)
        const double reduce_cost = folded ? 98.0 * std::min(nb, 131072.0) + 72.0 * std::max(0.0, nb - 131072.0) : W * nb * 3.5 * 14.0 * 2.0;
        const double cost = (double)n * W * 10.0 + reduce_cost + hot;
        if (cost < best) { best = cost; best_c = c; }
    }
    if (!best_c) { p.c = 0; return p; }                         // no feasible plan (folded table would not fit)
    p.c = best_c;
    p.windows = windows_for(bits, p.c);
    p.wide = p.windows;
    if (folded) balanced_windows(bits, p.c, &p.windows, &p.wide);
    p.nb = (1u << (p.c - 1)) >> class_log2;                     // a bucket-class shard reduces its own residue class only (both cost terms
    p.sets = folded ? 1 : p.windows;                            // above shrink by the class count, so the window width is the whole job's)
    // chunks (folded only): every chunk is a physical bucket set of its own (counts, sorted list, partial slots); the bucket
    // reduction merges the chunks of a logical set
    p.chunks = folded ? std::max<uint32_t>(1, std::min<uint32_t>(chunks, std::max<uint32_t>(1, n / 4096))) : 1;
    // Chunk sizes grow geometrically: every chunk is (growth - 1) times the sum of its predecessors (1 : 3 : 12 for three chunks).  The first
    // chunk is short (its upload is exposed: 0.58 us per 1000 scalars over PCIe 5); a later chunk is uploaded and recoded in sub-chunks while
    // its predecessors accumulate (1.7 us per 1000 points), so it may be about three times as long as everything before it before the
    // integer pipe would have to wait for it.  Measured at 2^24 (profiles/r2_msm_pipeline.md): growth 2.43 / 3.2 / 4.0 37.1 / 37.0 / 36.6 ms
    // end to end with three chunks.  PANDA_MSM_CHUNK_GROWTH overrides growth = 4.
    p.chunk_begin[0] = 0;
    if (p.chunks > 1) {
        static const double growth = [] { const char *e = getenv("PANDA_MSM_CHUNK_GROWTH"); const double v = e ? atof(e) : 0.0; return v >= 1.0 ? v : 4.0; }();
        p.chunks = std::min<uint32_t>(p.chunks, MSM_MAX_CHUNKS);
        double total = 0, wq = 1.0, acc = 0;
        for (uint32_t q = 0; q < p.chunks; q++) { total += q == 0 ? 1.0 : wq; if (q == 0) wq = growth - 1.0; else wq *= growth; }
        wq = 1.0;
        uint32_t used = 0;
        for (uint32_t q = 0; q < p.chunks; q++) {
            acc += q == 0 ? 1.0 : wq;
            if (q == 0) wq = growth - 1.0; else wq *= growth;
            uint32_t end = q + 1 == p.chunks ? n : (uint32_t)std::min<double>((double)n, (double)n * acc / total);
            end = std::min<uint32_t>(n, (end + 255) & ~255u);
            if (end <= p.chunk_begin[used]) continue;              // (tiny jobs: empty chunks are dropped)
            p.chunk_begin[++used] = end;
            if (end == n) break;
        }
        p.chunks = used;
    } else {
        p.chunk_begin[1] = n;
    }
    uint32_t max_chunk = 0;
    for (uint32_t q = 0; q < p.chunks; q++) max_chunk = std::max(max_chunk, p.chunk_begin[q + 1] - p.chunk_begin[q]);
    p.stride = folded ? max_chunk * p.windows : n;
    p.table_bytes = folded ? (size_t)n * p.windows * 2 * fq_bytes : 0;
    // segment length: about one average bucket, so that most buckets end up with one or two partial sums, but
    // never so long that the accumulation kernel has fewer than ~4 waves of threads (148 SMs x 384 threads)
    const uint64_t entries = ((uint64_t)n * p.windows) >> class_log2;           // expected (uniform digits)
    uint32_t L = pow2_floor(std::max<uint64_t>(1, entries / ((uint64_t)p.nb * p.sets)));
    L = std::min<uint32_t>(L, pow2_floor(std::max<uint64_t>(1, entries / (148ull * 384 * 4))));
    L = std::min<uint32_t>(std::max<uint32_t>(L, 8), 512);
    if (seg_override) L = seg_override;
    p.seg_len = L;
    p.segs_ps = p.stride ? (p.stride + L - 1) / L : 0;
    // reduction: ~2^18 threads fold m buckets each (about 7 waves of 2 x 128-thread CTAs per SM at the kernel's 200+ registers),
    // then CTAs of 256 threads stitch 1024 chunk sums each; more than 32 such groups get one more stitch level
    uint32_t m = pow2_floor(std::max<uint64_t>(1, ((uint64_t)p.sets * p.nb) / 262144));
    m = std::min<uint32_t>(std::min<uint32_t>(m, 32), p.nb);
    // small bucket sets: the stitching is a chain of ~60 dependent point additions per level (0.55 ms), so folding up to 8 buckets per
    // thread is worth it when it brings a set down to 32 groups of 1024 chunk sums (one stitch level instead of two; 16 per thread measured worse: 2^22 points, 1.77 vs 1.60 ms)
    if (p.nb / 8 <= 32768) m = std::max<uint32_t>(m, std::max<uint32_t>(1, p.nb / 32768));
    {   // PANDA_MSM_REDUCE_CHUNK (tuning): buckets folded per thread of the bucket reduction
        static const uint32_t forced_m = [] { const char *e = getenv("PANDA_MSM_REDUCE_CHUNK"); return e ? (uint32_t)atoi(e) : 0u; }();
        if (forced_m) m = std::min<uint32_t>(pow2_floor(forced_m), p.nb);
    }
    p.chunk = m;
    p.chunks_ps = p.nb / m;
    // at most one stitching CTA per SM over all sets (128 for one set): with 256 two of them share an SM and the latency-bound chains slow each other down
    // (measured at 2^24: 0.84 ms for 256 CTAs of 1024 chunk sums, against 0.42 ms for the single CTA of the next level)
    p.groups = std::min<uint32_t>(pow2_floor(std::max<uint32_t>(1, 148 / p.sets)), std::max<uint32_t>(1, p.chunks_ps / 1024));
    // folded scatter ranges: the slice of the sorted list written for one bucket range should stay in L2 (126 MB) while its 4-byte writes land
    // (<= 128 MiB, up to 32 ranges), and the scatter of one range hides behind the accumulation of the range before it, so only the first
    // scatter is exposed.  A range should still be a good part of a wave of accumulation segments (56 832 resident threads x L = 64 entries; up
    // to four range launches run side by side), hence >= 3 M entries per range and at most 8 ranges (16 from 512 MiB lists).  Measured
    // (profiles/r2_msm_pipeline.md): 2^24 (768 MiB list) 8 / 16 / 32 ranges 35.7 / 35.2 / 36.0 ms per MSM against 36.6 unpipelined; 2^22 1 / 4 / 8 / 16
    // ranges 10.10 / 10.01 / 9.99 / 10.12; 2^21 6.11 / 5.78 / 5.80 / 5.98; 2^20 (1 / 2 / 4 / 8) 3.46 / 3.39 / 3.33 / 3.45.
    p.phases = 1;
    if (folded) {
        static const uint32_t forced = [] { const char *e = getenv("PANDA_MSM_PHASES"); return e ? (uint32_t)atoi(e) : 0u; }();
        const uint64_t list_bytes = ((uint64_t)p.stride * 4) >> class_log2;
        const uint32_t max_ranges = list_bytes >= ((uint64_t)512 << 20) ? 16 : 8;
        while (p.phases < max_ranges && p.phases < p.nb && list_bytes / (p.phases * 2) >= ((uint64_t)12 << 20)) p.phases *= 2;   // >= 3 M entries per range
        while (p.phases < p.nb && list_bytes / p.phases > ((uint64_t)128 << 20)) p.phases *= 2;
        if (forced) p.phases = std::min<uint32_t>(pow2_floor(forced), p.nb);
        // tile codes keep 18 bits of the bucket index (the rest is the range) and the tile headers one lane per range
        p.phases = std::min<uint32_t>(p.phases, 32);
        while ((p.nb / p.phases) > (1u << 18)) p.phases *= 2;
    }
    const uint32_t chunk_tiles = (max_chunk + 255) / 256;
    p.codes_stride = folded ? chunk_tiles * 256 * p.windows : 0;
    p.heads_stride = folded ? chunk_tiles * (p.phases + 1) : 0;

    auto align = [](size_t v) { return (v + 255) & ~(size_t)255; };
    size_t off = 0;
    const size_t phys = (size_t)p.sets * p.chunks;                              // physical bucket sets
    p.off_counts = off;  off = align(off + phys * (p.nb + 1) * 4);              // per set: nb counts + its big-bucket counter
    p.off_offsets = off; off = align(off + phys * (p.nb + 1) * 4);
    p.off_cursor = off;  off = align(off + phys * p.nb * 4);
    p.off_biglist = off; off = align(off + phys * p.nb * 4);
    p.off_tiles = off;   off = align(off + phys * ((p.nb + 4095) / 4096) * 4);
    p.off_digits = off;  off = align(off + (folded ? (size_t)p.chunks * p.codes_stride * 4 : (size_t)p.windows * n * 4));
    p.off_heads = off;   off = align(off + (size_t)p.chunks * p.heads_stride * 2);
    p.off_sorted = off;  off = align(off + phys * p.stride * 4);
    p.off_slots = off;   off = align(off + phys * ((size_t)p.segs_ps + p.nb) * 4 * fq_bytes);
    p.off_chunks = off;  off = align(off + (size_t)p.sets * p.chunks_ps * 2 * 4 * fq_bytes);
    p.off_gsums = off;   off = align(off + (size_t)p.sets * (p.groups + 1) * 2 * 4 * fq_bytes);   // + the second stitch level
    p.bytes = off;
    return p;
}

// ----------------------------------------------------------------------------------------------------
// table cache for reused bases

// The table itself is reference counted: an MSM holds a reference from the lookup until its kernels are queued, so that a concurrent
// unregister / tear_down / eviction on another thread cannot free it in between (cudaFree waits for work that is already queued).
struct TableBuf {
    int device = 0;
    void *ptr = nullptr;
    cudaEvent_t ready = nullptr;    // recorded after the build; other streams wait on it
    ~TableBuf() {
        int cur = 0;
        cudaGetDevice(&cur);
        if (cur != device) cudaSetDevice(device);
        if (ptr) cudaFree(ptr);     // synchronises with outstanding work on the buffer
        if (ready) cudaEventDestroy(ready);
        if (cur != device) cudaSetDevice(cur);
    }
};

struct TableEntry {
    int device;
    CurveId curve;
    const void *bases;
    uint32_t n;
    unsigned long long fingerprint;
    unsigned sightings;
    bool registered;        // announced by msm_register_bases: immutable until unregistered, no fingerprint needed
    std::shared_ptr<TableBuf> table;     // empty until built
    uint32_t c, windows;
    size_t bytes;
    unsigned long long last_use;
};

static std::mutex g_table_mutex;
static std::vector<TableEntry> g_tables;
static unsigned long long g_use_clock = 0;
static constexpr size_t MAX_TABLES = 4;

static void drop_table(TableEntry &e) { e.table.reset(); }

cudaError_t msm_release_tables() {     // the current device's tables (panda_msm_tear_down: one MSM unit per device)
    std::lock_guard<std::mutex> lock(g_table_mutex);
    int cur = 0;
    cudaGetDevice(&cur);
    for (size_t i = 0; i < g_tables.size();) {
        if (g_tables[i].device == cur) { drop_table(g_tables[i]); g_tables.erase(g_tables.begin() + i); }
        else i++;
    }
    return cudaSuccess;
}

// PANDA_MSM_PRECOMPUTE: 0 = never build tables (not even for registered sets); 1 = also for UNANNOUNCED pointers, at their second
// sighting with an identical content fingerprint (costs a fingerprint pass and an 8-byte read-back, i.e. a host synchronisation, per
// call); 2 = unannounced pointers get a table at first sight; unset / 3 = tables only for sets announced with
// panda_msm_register_bases_* (the default: an execute call never blocks the host and never allocates a table behind the caller's back).
static int default_table_mode() {
    static int mode = [] {
        const char *e = getenv("PANDA_MSM_PRECOMPUTE");
        if (!e || !*e) return (int)MSM_TABLE_REGISTERED;
        int v = atoi(e);
        return v < 0 || v > 3 ? (int)MSM_TABLE_REGISTERED : v;
    }();
    return mode;
}

// upper bound for ONE table in bytes: PANDA_MSM_TABLE_BUDGET (GiB, fractions allowed) or half of what is free beyond a 4 GiB reserve
static size_t table_budget_bytes() {
    static const double forced_gib = [] { const char *e = getenv("PANDA_MSM_TABLE_BUDGET"); return e && *e ? atof(e) : -1.0; }();
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); return 0; }
    const size_t automatic = free_b > ((size_t)6 << 30) ? (free_b - ((size_t)4 << 30)) / 2 : 0;
    if (forced_gib >= 0) return std::min<size_t>((size_t)(forced_gib * 1073741824.0), free_b > ((size_t)1 << 30) ? free_b - ((size_t)1 << 30) : 0);
    return automatic;
}

#define PB_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "[panda-b200] CUDA error %d (%s) at %s:%d\n", (int)e_, cudaGetErrorString(e_), __FILE__, __LINE__); return e_; } } while (0)

// PANDA_L2_FETCH = 32 | 64 | 128: cudaLimitMaxL2FetchGranularity for the devices this library runs on (tuning experiment: the 64-byte point
// gathers arrive as 128-byte lines by default)
static void apply_l2_fetch_limit() {
    static const int bytes = [] { const char *e = getenv("PANDA_L2_FETCH"); return e ? atoi(e) : 0; }();
    static std::atomic<unsigned long long> done{0};
    if (!bytes) return;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev >= 64) return;
    const unsigned long long bit = 1ull << dev;
    if (done.fetch_or(bit) & bit) return;
    if (cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)bytes) != cudaSuccess) cudaGetLastError();
}

static cudaError_t run_pipeline(CurveId curve, const MsmPlan &p, const void *points, const void *scalars, void *result, CoordType coord,
                                cudaMemPool_t pool, cudaStream_t stream, MsmStageTimes *timings, const MsmFeed *feed = nullptr) {
    if (timings) { timings->folded = (int)p.folded; timings->c = p.c; timings->windows = p.windows; }
    apply_l2_fetch_limit();
    if (curve == CURVE_BLS12_377) return msm_pipeline_bls12_377(p, points, scalars, result, coord, pool, stream, timings, feed);
    if (curve == CURVE_BLS12_381) return msm_pipeline_bls12_381(p, points, scalars, result, coord, pool, stream, timings, feed);
    return msm_pipeline_bn254(p, points, scalars, result, coord, pool, stream, timings, feed);
}

// Builds the table of `hit` on `stream` (caller holds g_table_mutex).  Not fitting in memory is not an error: the entry stays table-less.
static cudaError_t build_table_locked(TableEntry *hit, CurveId curve, const void *bases, uint32_t n, uint32_t c_override, cudaStream_t stream) {
    const size_t budget = table_budget_bytes();
    MsmPlan fp_plan = msm_make_plan(curve, n, true, c_override > 16 ? c_override : 0, 0, budget);
    if (!fp_plan.c) return cudaSuccess;                // no table fits: stay on the windowed path
    auto buf = std::make_shared<TableBuf>();
    PB_CUDA(cudaGetDevice(&buf->device));
    if (cudaMalloc(&buf->ptr, fp_plan.table_bytes) != cudaSuccess) { cudaGetLastError(); buf->ptr = nullptr; return cudaSuccess; }
    cudaError_t be = curve == CURVE_BLS12_381 ? msm_build_table_bls12_381(bases, n, fp_plan.c, fp_plan.windows, fp_plan.wide, buf->ptr, stream)
                   : curve == CURVE_BLS12_377 ? msm_build_table_bls12_377(bases, n, fp_plan.c, fp_plan.windows, fp_plan.wide, buf->ptr, stream)
                                              : msm_build_table_bn254(bases, n, fp_plan.c, fp_plan.windows, fp_plan.wide, buf->ptr, stream);
    if (be == cudaSuccess) be = cudaEventCreateWithFlags(&buf->ready, cudaEventDisableTiming);
    if (be == cudaSuccess) be = cudaEventRecord(buf->ready, stream);
    if (be != cudaSuccess) return be;                  // buf's destructor frees
    hit->table = buf; hit->c = fp_plan.c; hit->windows = fp_plan.windows; hit->bytes = fp_plan.table_bytes;
    return cudaSuccess;
}

static void evict_if_full_locked() {
    if (g_tables.size() < MAX_TABLES) return;
    size_t victim = g_tables.size();                   // least recently used entry that is not registered
    for (size_t i = 0; i < g_tables.size(); i++)
        if (!g_tables[i].registered && (victim == g_tables.size() || g_tables[i].last_use < g_tables[victim].last_use)) victim = i;
    if (victim == g_tables.size()) return;             // only registered sets: let the vector grow (the caller owns their lifetime)
    drop_table(g_tables[victim]);
    g_tables.erase(g_tables.begin() + victim);
}

cudaError_t msm_register_bases(CurveId curve, const void *bases, uint32_t n, cudaStream_t stream) {
    if (!bases || n == 0) return cudaErrorInvalidValue;
    int dev = 0;
    PB_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_table_mutex);
    TableEntry *hit = nullptr;
    for (auto &t : g_tables) if (t.device == dev && t.bases == bases) { hit = &t; break; }
    if (hit) { drop_table(*hit); *hit = TableEntry{dev, curve, bases, n, 0, 0, true, nullptr, 0, 0, 0, ++g_use_clock}; }
    else {
        evict_if_full_locked();
        g_tables.push_back(TableEntry{dev, curve, bases, n, 0, 0, true, nullptr, 0, 0, 0, ++g_use_clock});
        hit = &g_tables.back();
    }
    if (default_table_mode() == MSM_TABLE_OFF || n < 1024) return cudaSuccess;      // PANDA_MSM_PRECOMPUTE=0: windowed plan everywhere
    return build_table_locked(hit, curve, bases, n, 0, stream);
}

cudaError_t msm_unregister_bases(const void *bases) {
    int dev = 0;
    PB_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_table_mutex);
    for (size_t i = 0; i < g_tables.size(); i++)
        if (g_tables[i].device == dev && g_tables[i].bases == bases) {
            drop_table(g_tables[i]);
            g_tables.erase(g_tables.begin() + i);
            return cudaSuccess;
        }
    return cudaSuccess;                                // unknown pointer: nothing to do (idempotent)
}

// Looks `bases` up in the table cache (building the table when the mode asks for it).  On return *table holds a reference to the
// table to use (with its window width in *tc, its row length in *table_n and `stream` already waiting for the build) or is empty:
// stay on the windowed path.  The caller keeps the reference until its kernels are queued.
static cudaError_t acquire_table(CurveId curve, const void *bases, uint32_t n, cudaStream_t stream, int table_mode, uint32_t c_override,
                                 std::shared_ptr<TableBuf> *table, uint32_t *tc, uint32_t *table_n) {
    table->reset();
    *table_n = n;
    const size_t fq_bytes = curve_fq_bytes(curve);
    if (table_mode == MSM_TABLE_DEFAULT) table_mode = default_table_mode();
    if (table_mode == MSM_TABLE_OFF || n < 1024) return cudaSuccess;
    int dev = 0;
    PB_CUDA(cudaGetDevice(&dev));
    {   // registered base sets (init_msm): no fingerprint, no host synchronisation; a prefix of the set may be used
        std::unique_lock<std::mutex> lock(g_table_mutex);
        for (auto &t : g_tables)
            if (t.registered && t.device == dev && t.curve == curve && t.bases == bases && n <= t.n) {
                t.last_use = ++g_use_clock;
                if (!t.table) return cudaSuccess;
                *table = t.table; *tc = t.c; *table_n = t.n;
                lock.unlock();
                PB_CUDA(cudaStreamWaitEvent(stream, (*table)->ready, 0));
                return cudaSuccess;
            }
    }
    if (table_mode == MSM_TABLE_REGISTERED) return cudaSuccess;        // unannounced pointer: windowed plan, nothing blocks, nothing is allocated
    // opt-in modes (PANDA_MSM_PRECOMPUTE = 1 | 2, or the debug entry point):
    // 1. content fingerprint of the bases (one pass over n * 64 bytes at HBM speed, 8 bytes read back -- a host synchronisation)
    unsigned long long *d_fp = nullptr, fp = 0;
    PB_CUDA(cudaMallocAsync((void **)&d_fp, 8, stream));
    cudaError_t e = msm_fingerprint_launch(bases, (size_t)n * 2 * fq_bytes, d_fp, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(&fp, d_fp, 8, cudaMemcpyDeviceToHost, stream);
    cudaError_t f = cudaFreeAsync(d_fp, stream);
    if (e == cudaSuccess) e = f;
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    PB_CUDA(e);

    std::unique_lock<std::mutex> lock(g_table_mutex);
    TableEntry *hit = nullptr;
    for (auto &t : g_tables)
        if (!t.registered && t.device == dev && t.curve == curve && t.bases == bases && t.n == n) { hit = &t; break; }
    if (hit && hit->fingerprint != fp) {      // same pointer, different points: forget what we knew
        drop_table(*hit);
        hit->fingerprint = fp; hit->sightings = 0;
    }
    if (!hit) {
        evict_if_full_locked();
        g_tables.push_back(TableEntry{dev, curve, bases, n, fp, 0, false, nullptr, 0, 0, 0, 0});
        hit = &g_tables.back();
    }
    hit->sightings++;
    hit->last_use = ++g_use_clock;
    const bool want = table_mode == MSM_TABLE_EAGER || hit->sightings >= 2;
    if (!hit->table && want) PB_CUDA(build_table_locked(hit, curve, bases, n, c_override, stream));
    if (hit->table) {
        *table = hit->table;
        *tc = hit->c;
        lock.unlock();
        PB_CUDA(cudaStreamWaitEvent(stream, (*table)->ready, 0));
    }
    return cudaSuccess;
}

static cudaError_t side_stream_for_current_device(int which, cudaStream_t *out);
static uint32_t resident_chunks(uint32_t n);
static cudaError_t side_streams(MsmFeed *feed);      // the current device's four side streams into feed->aux*_stream

cudaError_t msm_run(CurveId curve, const void *bases, const void *scalars, uint32_t n, void *result, CoordType coord,
                    cudaMemPool_t pool, cudaStream_t stream, uint32_t c_override, uint32_t seg_override, MsmStageTimes *timings, int table_mode,
                    uint32_t class_count, uint32_t class_index) {
    const size_t fq_bytes = curve_fq_bytes(curve);
    if (class_count == 0 || class_count > 64 || (class_count & (class_count - 1)) || class_index >= class_count) return cudaErrorInvalidValue;
    uint32_t class_log2 = 0;
    while ((1u << class_log2) < class_count) class_log2++;
    if (n == 0) {   // empty sum: the identity, all-zero like the reference (msm_cuda.cuh:395,405)
        PB_CUDA(cudaMemsetAsync(result, 0, 3 * fq_bytes, stream));
        return cudaSuccess;
    }
    std::shared_ptr<TableBuf> table_ref;           // held until the kernels are queued
    uint32_t tc = 0;
    uint32_t table_n = n;
    PB_CUDA(acquire_table(curve, bases, n, stream, table_mode, c_override, &table_ref, &tc, &table_n));
    const void *table = table_ref ? table_ref->ptr : nullptr;
    if (table) {
        const uint32_t chunks = (timings || class_log2) ? 1 : resident_chunks(n);
        MsmPlan p = msm_make_plan(curve, n, true, tc, seg_override, ~(size_t)0, chunks, class_log2);
        if (!p.c) return cudaErrorInvalidValue;
        p.table_n = table_n;
        p.class_index = class_index;
        if (p.chunks > 1) {
            MsmFeed feed{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
            PB_CUDA(side_streams(&feed));
            return run_pipeline(curve, p, table, scalars, result, coord, pool, stream, nullptr, &feed);
        }
        // one chunk, several scatter ranges: range r+1 is scattered while range r is accumulated (PANDA_MSM_PIPELINE=0: one launch each)
        static const bool pipeline_on = [] { const char *v = getenv("PANDA_MSM_PIPELINE"); return !v || atoi(v) != 0; }();
        if (pipeline_on && p.phases > 1) {
            MsmFeed feed{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
            PB_CUDA(side_streams(&feed));
            return run_pipeline(curve, p, table, scalars, result, coord, pool, stream, timings, &feed);
        }
        return run_pipeline(curve, p, table, scalars, result, coord, pool, stream, timings);
    }
    MsmPlan p = msm_make_plan(curve, n, false, c_override <= 16 ? c_override : 0, seg_override, ~(size_t)0, 1, class_log2);
    if (!p.c) return cudaErrorInvalidValue;
    p.class_index = class_index;
    // windowed plan: window w+1 is scattered while window w is accumulated, from 2^22 sorted entries (a window is then a wave of segments or more)
    static const bool pipeline_on = [] { const char *v = getenv("PANDA_MSM_PIPELINE"); return !v || atoi(v) != 0; }();
    if (pipeline_on && ((uint64_t)n * p.windows >> class_log2) >= ((uint64_t)1 << 22)) {
        MsmFeed feed{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
        PB_CUDA(side_streams(&feed));
        return run_pipeline(curve, p, bases, scalars, result, coord, pool, stream, timings, &feed);
    }
    return run_pipeline(curve, p, bases, scalars, result, coord, pool, stream, timings);
}

// per device: one copy stream (streamed scalars), one high-priority auxiliary compute stream (chunk / range overlap) and one more at the
// caller's priority, created on first use
static cudaError_t side_stream_for_current_device(int which, cudaStream_t *out) {
    static std::mutex m;
    static cudaStream_t streams[6][64] = {};
    int dev = 0;
    PB_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    std::lock_guard<std::mutex> lock(m);
    if (!streams[which][dev]) {
        int least = 0, greatest = 0;                       // the auxiliary compute stream outranks the caller's stream, so that its
        PB_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));   // (small) sort CTAs slot in between the accumulation CTAs
        PB_CUDA(cudaStreamCreateWithPriority(&streams[which][dev], cudaStreamNonBlocking, (which == 1 || which == 5) ? greatest : least));
    }
    *out = streams[which][dev];
    return cudaSuccess;
}
static cudaError_t copy_stream_for_current_device(cudaStream_t *out) { return side_stream_for_current_device(0, out); }
static cudaError_t side_streams(MsmFeed *feed) {
    PB_CUDA(side_stream_for_current_device(1, &feed->aux_stream));
    PB_CUDA(side_stream_for_current_device(2, &feed->aux2_stream));
    PB_CUDA(side_stream_for_current_device(3, &feed->aux3_stream));
    PB_CUDA(side_stream_for_current_device(4, &feed->aux4_stream));
    PB_CUDA(side_stream_for_current_device(5, &feed->dig_stream));
    return cudaSuccess;
}

// Point-range chunks of a device-resident table-plan MSM (chunk q+1 sorts while chunk q accumulates).  Round-2 history: 2 chunks lost at 2^24 / 2^25 (the
// second set of partial sums costs more than the overlap wins) and 4 chunks won at 2^26 (143.8 ms against 151.8: 32 filter passes over 3.2 GB of codes);
// the range pipeline now overlaps the sort WITHOUT a second set of partial sums (2^26: 125.5 ms), so resident jobs run as one chunk.
// PANDA_MSM_SPLIT = q forces q chunks (tests, tuning).
static uint32_t resident_chunks(uint32_t n) {
    static const int forced = [] { const char *v = getenv("PANDA_MSM_SPLIT"); return v ? atoi(v) : 0; }();
    if (forced > 0) return (uint32_t)forced;
    (void)n;
    return 1;
}

cudaError_t msm_run_streamed(CurveId curve, const void *bases, const void *host_scalars, uint32_t n, void *result, CoordType coord,
                             cudaMemPool_t pool, cudaStream_t stream, int table_mode, uint32_t chunks_override) {
    const size_t fq_bytes = curve_fq_bytes(curve);
    if (n == 0) {
        PB_CUDA(cudaMemsetAsync(result, 0, 3 * fq_bytes, stream));
        return cudaSuccess;
    }
    std::shared_ptr<TableBuf> table_ref;
    uint32_t tc = 0;
    uint32_t table_n = n;
    PB_CUDA(acquire_table(curve, bases, n, stream, table_mode, 0, &table_ref, &tc, &table_n));
    const void *table = table_ref ? table_ref->ptr : nullptr;
    uint8_t *d_scal = nullptr;
    if (pool) PB_CUDA(cudaMallocFromPoolAsync((void **)&d_scal, (size_t)n * 32, pool, stream));
    else PB_CUDA(cudaMallocAsync((void **)&d_scal, (size_t)n * 32, stream));
    cudaError_t e;
    if (table) {
        static const uint32_t forced = [] { const char *v = getenv("PANDA_MSM_CHUNKS"); return v ? (uint32_t)atoi(v) : 0u; }();
        uint32_t chunks = chunks_override ? chunks_override : forced ? forced : (n >= (1u << 19) ? 3 : 1);
        MsmPlan p = msm_make_plan(curve, n, true, tc, 0, ~(size_t)0, chunks);
        p.table_n = table_n;
        MsmFeed feed{host_scalars, d_scal, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
        e = copy_stream_for_current_device(&feed.copy_stream);
        if (e == cudaSuccess) e = side_streams(&feed);
        if (e == cudaSuccess) e = run_pipeline(curve, p, table, d_scal, result, coord, pool, stream, nullptr, &feed);
    } else {
        e = cudaMemcpyAsync(d_scal, host_scalars, (size_t)n * 32, cudaMemcpyHostToDevice, stream);
        if (e == cudaSuccess) {
            MsmPlan p = msm_make_plan(curve, n, false, 0, 0);
            e = run_pipeline(curve, p, bases, d_scal, result, coord, pool, stream, nullptr);
        }
    }
    cudaError_t f = cudaFreeAsync(d_scal, stream);
    return e != cudaSuccess ? e : f;
}

cudaError_t msm_combine(CurveId curve, const void *partials, uint32_t count, void *result, CoordType coord, cudaStream_t stream) {
    if (curve == CURVE_BLS12_377) return msm_combine_bls12_377(partials, count, result, coord, stream);
    if (curve == CURVE_BLS12_381) return msm_combine_bls12_381(partials, count, result, coord, stream);
    return msm_combine_bn254(partials, count, result, coord, stream);
}

}  // namespace pb
