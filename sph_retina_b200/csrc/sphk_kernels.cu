// sphk_kernels.cu -- sm_100a kernels and the C ABI (include/sphk.h) of the spherical-box IoU path.
//
// Kernels (all fp32 CUDA-core work; no tensor cores -- nothing here is a contraction):
//   k_iou_aligned2     aligned Sph2Pob IoU: stage 1 per pair, stage 2 + clipper warp-compacted        (config #1)
//   k_iou_aligned      aligned Sph-IoU / FoV-IoU, one thread per pair                                 (siblings)
//   k_box_pre          per-box records (BoxRec + BoxCull) of both operands of an N x M call
//   k_iou_pairwise2    N x M Sph2Pob IoU: lane <-> column tiles, 7-FMA prefilter, ballot compaction into
//                      per-warp rings, fused row/column max+argmax keys, tie pass, image batches  (configs #2, #5)
//   k_iou_pairwise     N x M Sph-IoU / FoV-IoU (plain tile kernel)
//   k_iou_project      rbb_angle = 'project' ablation variant, double precision
//   k_assign_targets, k_assign_epilogue   MaxIoUAssigner thresholds / low-quality matching on the packed keys
//   k_loss_fwd_bwd, k_loss_reduce         IoU + both gradients in one launch (elementwise / reduced)   (config #3)
//   k_obb_fwd/bwd, k_riou_fwd_bwd         the two loss stages exposed separately (GIoU/DIoU/CIoU epilogues)
//   k_nms              one CTA per (image, class) segment, 32-pivot rounds of warp-ballot suppression
//                      words + in-warp greedy scan                                                    (config #4)
// The per-pair arithmetic lives in sphk_math.cuh (reference-order path, clipper), sphk_fast.cuh (records,
// prefilter, fast path) and sphk_grad.cuh (analytic backward).
#include <cuda_runtime.h>

#include <atomic>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/sphk.h"
#include "sphk_coder.cuh"
#include "sphk_format.cuh"
#include "sphk_fast.cuh"
#include "sphk_grad.cuh"
#include "sphk_math.cuh"
#include "sphk_obbloss.cuh"

using namespace sphk;

namespace {

// -DSPHK_CHECKED (python -m sph_retina_b200.build --checked; tools/checked_run.sh): the invariants of the shared-memory
// rings and the index ranges of every compacted store become traps.  compute-sanitizer is closed on the GPU pool this
// was developed on, so these are the bounds checks "of our own": a violated invariant kills the launch with an error
// instead of silently overwriting a neighbour's entry.  The product build compiles them away.
#ifdef SPHK_CHECKED
#define SPHK_CHECK(cond) do { if (!(cond)) { printf("SPHK_CHECK failed: %s (%s:%d) block %d thread %d\n", #cond, __FILE__, __LINE__, (int)blockIdx.x, (int)threadIdx.x); __trap(); } } while (0)
#else
#define SPHK_CHECK(cond) do { } while (0)
#endif

thread_local char g_err[256] = "";
int g_dense = 0;   // sphk_set_dense: 1 = no disjoint-pair early-outs (measurement only)
#ifdef SPHK_TUNING
// A/B and timing hooks of the tools/ scripts: compiled ONLY into the instrumented build (build.py --tuning ->
// _lib/libsphk_tuning.so, loaded through SPHK_PROBE_LIB); the shipped library reads no environment variable.
int tune_env(const char* name) { const char* e = getenv(name); return e ? atoi(e) : 0; }
const int g_force_ctas = tune_env("SPHK_ALIGNED_CTAS");
const int g_force_minb = tune_env("SPHK_ALIGNED_MINB");
const int g_force_tr = tune_env("SPHK_TR");
const int g_no_rows32 = tune_env("SPHK_NO_ROWS32");
const int g_no_approx4 = tune_env("SPHK_NO_APPROX4");
const int g_no_pdl = tune_env("SPHK_NO_PDL");          // launch k_iou_rows32 / k_iou_pairwise2 without programmatic serialization
const int g_no_boxcull = tune_env("SPHK_NO_BOXCULL");  // prefilter: circle test only
const int g_no_sat = tune_env("SPHK_NO_SAT");          // no separating-axis stage (calls that would get it run the circle test alone)
const int g_no_rows_inline = tune_env("SPHK_NO_ROWS_INLINE");   // long-row calls: row records through the workspace, as for short-row calls
// bit 0 = do not launch k_box_pre (stale records)
const int g_probe = tune_env("SPHK_PROBE");
#else
constexpr int g_force_ctas = 0, g_force_minb = 0, g_force_tr = 0, g_no_rows32 = 0, g_no_approx4 = 0, g_no_pdl = 0,
              g_no_boxcull = 0, g_no_sat = 0, g_no_rows_inline = 0, g_probe = 0;
#endif

int fail(int code, const char* what) {
    snprintf(g_err, sizeof(g_err), "%s", what);
    return code;
}
int cuda_fail(cudaError_t e, const char* where) {
    snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
    return SPHK_ERR_CUDA;
}

// Dynamic shared memory above 48 KB needs an opt-in per kernel (and device).  It is raised ONCE to the architecture's
// maximum instead of to the size of the call at hand: two host threads launching the same kernel with different sizes
// can then never lower each other's limit between the attribute call and the launch.  TAG tells instances of one
// function-pointer type apart.
constexpr int kMaxDynamicSmem = 200 * 1024;       // what the callers size their requests against
template <int TAG, typename K>
cudaError_t allow_max_dynamic_smem(K kernel) {
    static std::atomic<unsigned long long> done{0ull};
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = 1ull << (dev & 63);
    if (done.load(std::memory_order_acquire) & bit) return cudaSuccess;
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, kernel);
    int optin = 0;
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (e != cudaSuccess) return e;
    // (the opt-in limit covers static + dynamic shared memory of the kernel)
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - (int)fa.sharedSizeBytes);
    if (e == cudaSuccess) done.fetch_or(bit, std::memory_order_release);
    return e;
}
#define SPHK_LAUNCH_CHECK(where)                                  \
    do {                                                          \
        cudaError_t e_ = cudaGetLastError();                      \
        if (e_ != cudaSuccess) return cuda_fail(e_, where);       \
    } while (0)

constexpr int kThreads = 256;
#ifndef SPHK_LOSS_SCALAR
#define SPHK_LOSS_SCALAR double    // arithmetic type of the GD / KF row losses (dual numbers): their cancellations need it
#endif

// ---- box loads --------------------------------------------------------------------------------
template <int D>
__device__ __forceinline__ RawBox load_box(const float* __restrict__ b, int64_t i, bool vec_ok) {
    RawBox r;
    if (D == 4) {
        if (vec_ok) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(b) + i);
            r.t = v.x; r.p = v.y; r.a = v.z; r.b = v.w;
        } else {
            const float* q = b + i * 4;
            r.t = __ldg(q); r.p = __ldg(q + 1); r.a = __ldg(q + 2); r.b = __ldg(q + 3);
        }
        r.g = 0.0f;
    } else {
        const float* q = b + i * 5;
        r.t = __ldg(q); r.p = __ldg(q + 1); r.a = __ldg(q + 2); r.b = __ldg(q + 3); r.g = __ldg(q + 4);
    }
    return r;
}

template <int D>
__device__ __forceinline__ void store_grad(float* __restrict__ g, int64_t i, const float* v, bool vec_ok) {
    if (D == 4 && vec_ok) {
        reinterpret_cast<float4*>(g)[i] = make_float4(v[0], v[1], v[2], v[3]);
    } else {
        float* q = g + i * D;
#pragma unroll
        for (int k = 0; k < D; ++k) q[k] = v[k];
    }
}

// The reference-order path of one pair, out of line: it is rare on the N x M / aligned workloads, and
// keeping ONE copy of its ~1300 instructions per library keeps the hot kernels inside the instruction cache.
__device__ __noinline__ float slow_pair_iou(const float* __restrict__ b1, int64_t i1, const float* __restrict__ b2,
                                            int64_t i2, int D, int kind, int mode, int edge, bool dense) {
    RawBox x, y;
    const float* q = b1 + i1 * D;
    x.t = __ldg(q); x.p = __ldg(q + 1); x.a = __ldg(q + 2); x.b = __ldg(q + 3); x.g = (D == 5) ? __ldg(q + 4) : 0.0f;
    q = b2 + i2 * D;
    y.t = __ldg(q); y.p = __ldg(q + 1); y.a = __ldg(q + 2); y.b = __ldg(q + 3); y.g = (D == 5) ? __ldg(q + 4) : 0.0f;
    return sph2pob_iou_pair(x, y, D, kind, mode, edge, dense);
}

// One pair, cheapest applicable route: exact 0 if provably disjoint, the two-stage fast path, else the
// out-of-line reference-order path (indices i1 / i2 into b1 / b2).
__device__ __forceinline__ float pair_iou_any(const float* __restrict__ b1, int64_t i1, const float* __restrict__ b2,
                                              int64_t i2, const RawBox& x, const RawBox& y, int D, int kind, int mode,
                                              int edge) {
    PairS1 s;
    ClipJob j;
    int st = pair_stage1(x, y, D, edge, true, &s);
    if (st == JOB_SLOW) st = pair_stage1_general(x, y, D, edge, true, &s);
    if (st == JOB_READY) st = pair_stage2(s, D, kind, &j);
    if (st == JOB_DEAD) return 0.0f;
    if (st == JOB_READY) return clip_job_iou(j, mode);
    return slow_pair_iou(b1, i1, b2, i2, D, kind, mode, edge, false);
}

// ---- aligned -----------------------------------------------------------------------------------
template <int KIND, int D>
__global__ void __launch_bounds__(kThreads) k_iou_aligned(const float* __restrict__ b1, const float* __restrict__ b2,
                                                          int64_t P, int mode, int edge, float* __restrict__ out,
                                                          bool vec_ok, bool dense) {
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= P) return;
    const RawBox x = load_box<D>(b1, i, vec_ok), y = load_box<D>(b2, i, vec_ok);
    float v;
    if (KIND == KIND_SPH || KIND == KIND_FOV) v = approx_iou_pair(x, y, KIND);
    else if (KIND == KIND_NAIVE) v = naive_iou_pair(x, y, D, mode);
    else if (KIND == KIND_UNBIASED) v = unbiased_iou_pair(x, y, D);
    else if (KIND == KIND_SPH2POB_LEGACY) v = sph2pob_legacy_iou_pair(x, y, mode, edge);
    else v = sph2pob_iou_pair(x, y, D, KIND, mode, edge, dense);
    out[i] = v;
}

// Sph-IoU / FoV-IoU, aligned: ~100 instructions on 36 bytes per pair, i.e. bound by how many bytes are in flight.
// Four pairs per thread, all eight 16-byte loads issued before the first use (a CTA covers 1024 consecutive pairs, the
// four passes are each coalesced): 16 M pairs run at 5.86 TB/s (89.5 % of the measured copy bandwidth; 5.4 TB/s before the
// jitter_1 identity shortcut of approx_iou_pair) against 4.7 TB/s
// with one pair per thread; the one-pair kernel is kept for small launches, where thread count matters more.
template <int KIND>
__global__ void __launch_bounds__(kThreads) k_approx_aligned4(const float4* __restrict__ b1, const float4* __restrict__ b2,
                                                               int64_t P, float* __restrict__ out) {
    const int64_t i0 = (int64_t)blockIdx.x * (kThreads * 4) + threadIdx.x;
    float4 x[4], y[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int64_t i = i0 + k * kThreads;
        if (i < P) { x[k] = __ldg(b1 + i); y[k] = __ldg(b2 + i); }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int64_t i = i0 + k * kThreads;
        if (i < P) {
            RawBox a, b;
            a.t = x[k].x; a.p = x[k].y; a.a = x[k].z; a.b = x[k].w; a.g = 0.0f;
            b.t = y[k].x; b.p = y[k].y; b.a = y[k].z; b.b = y[k].w; b.g = 0.0f;
            out[i] = approx_iou_pair(a, b, KIND);
        }
    }
}

// ---- aligned, Sph2Pob kinds: cheap conservative cull for all pairs, everything exact warp-compacted ---------
// Stage 0 (pair_far_apart) runs for every pair: ~40 instructions (four MUFU sines, a lowered haversine against the
// circumradii) prove about 58 % of random pairs disjoint -- an exact 0.  The raw boxes of the survivors are
// ballot-compacted into a per-warp ring in shared memory; jitter_1, the degree-domain sincos, the transform
// (pair_stage1 / pair_stage2) and the clipper then run 32 survivors at a time with full warps.  Pairs on which a
// reference quirk may be active are queued once more for the out-of-line reference-order path.
// Ring capacities: survivors <= 31 pending + 32 pushed per round; slow pairs <= 31 pending + 32 from one batch.
constexpr int kJobRing = 64, kSlowRing = 64, kJobWords = 12;   // 48-byte entries: conflict-free LDS.128 / STS.128
#ifdef SPHK_TIMELINE
// instrumented build (tools/timeline_aligned.py): per warp of k_iou_aligned2 -- start, end of its last scan round, end
// (globaltimer ns), and how many pairs it sent to the reference-order path
__device__ unsigned long long g_tlw[8192 * 3];
__device__ unsigned g_tlw_slow[8192];
__device__ __forceinline__ unsigned long long tl_now() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#endif

struct AlignedTile {
    // entry = (t1, p1, a1, b1 | t2, p2, a2, b2 | g1, g2, round << 5 | lane, -)
    __align__(16) float job[kThreads / 32][kJobRing][kJobWords];
    int sidx[kThreads / 32][kSlowRing];
    // boxes of the next round, fetched with cp.async one round ahead: [warp][operand][5 x 32 floats]
    // (vectorised BFoV: lane-major float4; otherwise component-major, conflict-free either way)
    __align__(16) float pre[kThreads / 32][2][160];
};

__device__ __forceinline__ void cp_async4(float* s, const float* g) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(s)), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async16(float* s, const float* g) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(s)), "l"(g) : "memory");
}
// Fetch the 32 boxes of chunk c (box index c * 32 + lane) of one operand into the warp's staging slot.
// VEC (16-byte aligned base): a whole chunk is 32 * D contiguous floats, copied as 16-byte pieces (D = 4: one per
// lane; D = 5: 40 pieces, lanes 0-7 take a second one) into a box-major slot; the last, partial chunk of the operand
// (index `full`, `tail` boxes) uses 4-byte copies into the same layout.  Not VEC: 4-byte copies, component-major.
template <int D, bool VEC>
__device__ __forceinline__ void prefetch_chunk(float* slot, const float* __restrict__ b, int c, int full, int tail, int lane) {
    const float* src = b + (int64_t)c * (32 * D);
    if (VEC && c < full) {
        cp_async16(slot + 4 * lane, src + 4 * lane);
        if (D == 5 && lane < 8) cp_async16(slot + 128 + 4 * lane, src + 128 + 4 * lane);
    } else if (c < full || lane < tail) {
#pragma unroll
        for (int k = 0; k < D; ++k) cp_async4(VEC ? slot + lane * D + k : slot + k * 32 + lane, src + lane * D + k);
    }
}
template <int D, bool VEC>
__device__ __forceinline__ RawBox fetched_box(const float* slot, int lane) {
    RawBox r;
    if (VEC && D == 4) {
        const float4 v = *reinterpret_cast<const float4*>(slot + lane * 4);
        r.t = v.x; r.p = v.y; r.a = v.z; r.b = v.w; r.g = 0.0f;
    } else if (VEC) {       // box-major, stride 5 words: conflict-free
        const float* q = slot + lane * 5;
        r.t = q[0]; r.p = q[1]; r.a = q[2]; r.b = q[3]; r.g = q[4];
    } else {
        r.t = slot[lane]; r.p = slot[32 + lane]; r.a = slot[64 + lane]; r.b = slot[96 + lane];
        r.g = (D == 5) ? slot[128 + lane] : 0.0f;
    }
    return r;
}

template <int D, bool VEC, int MINB>
__global__ void __launch_bounds__(kThreads, MINB)
k_iou_aligned2(const float* __restrict__ b1, const float* __restrict__ b2, int64_t P, int kind, int mode, int edge,
               float* __restrict__ out, bool dense) {
    __shared__ AlignedTile T;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // Persistent warps: the grid is at most one resident wave; warp wg takes the 32-pair chunks wg, wg + nw,
    // wg + 2 nw, ... so that at any moment the whole GPU streams one contiguous window of the inputs.
    // A queued pair is remembered as (round << 5 | lane): the host keeps rounds < 2^26 (and chunks < 2^31).
    const int nw = (int)gridDim.x * (kThreads / 32), wg = (int)blockIdx.x * (kThreads / 32) + warp;
    const int full = (int)(P >> 5), tail = (int)(P & 31), chunks = full + (tail ? 1 : 0);
    int hj = 0, tj = 0, hs = 0, ts = 0;
    const unsigned lt = (1u << lane) - 1u;
#ifdef SPHK_TIMELINE
    const unsigned long long tl_start = tl_now();
    unsigned long long tl_scan = tl_start;
#endif
    float* const slot1 = T.pre[warp][0];
    float* const slot2 = T.pre[warp][1];
    // the boxes of the next round are fetched one round ahead (cp.async into shared memory: no registers are
    // held across the arithmetic, and the DRAM latency hides behind a round or a whole batch)
    int c = wg, k = 0;
    if (c < chunks) {
        prefetch_chunk<D, VEC>(slot1, b1, c, full, tail, lane);
        prefetch_chunk<D, VEC>(slot2, b2, c, full, tail, lane);
    }
#pragma unroll 1
    for (;;) {
        // ---- scan: stage 0 on whole chunks until 32 survivors are queued or the chunks are used up
#pragma unroll 1
        while (c < chunks && tj - hj < 32) {
            asm volatile("cp.async.wait_all;" ::: "memory");
            if (VEC) __syncwarp();                      // pieces were fetched by other lanes
            const int64_t i = (int64_t)c * 32 + lane;
            const bool ok = c < full || lane < tail;
            RawBox x, y;
            x.t = x.p = x.a = x.b = x.g = 0.0f;
            y = x;
            if (ok) { x = fetched_box<D, VEC>(slot1, lane); y = fetched_box<D, VEC>(slot2, lane); }
            if (VEC) __syncwarp();                      // everybody has read the slot before it is refilled
            c += nw;
            if (c < chunks) {
                prefetch_chunk<D, VEC>(slot1, b1, c, full, tail, lane);
                prefetch_chunk<D, VEC>(slot2, b2, c, full, tail, lane);
            }
            bool live = false;
            if (ok) {
                live = dense || !pair_far_apart(x, y, edge);
                if (!live) out[i] = 0.0f;
            }
            const unsigned mj = __ballot_sync(0xFFFFFFFFu, live);
            if (live) {
                float4* e = reinterpret_cast<float4*>(T.job[warp][(tj + __popc(mj & lt)) & (kJobRing - 1)]);
                e[0] = make_float4(x.t, x.p, x.a, x.b);
                e[1] = make_float4(y.t, y.p, y.a, y.b);
                e[2] = make_float4(x.g, y.g, __int_as_float((k << 5) | lane), 0.0f);
            }
            tj += __popc(mj);
            SPHK_CHECK(tj - hj <= kJobRing);
            ++k;
#ifdef SPHK_TIMELINE
            tl_scan = tl_now();
#endif
        }
        // ---- one batch of up to 32 survivors: jitter_1, transform, clipper (the expensive code exists once)
        if (tj > hj) {
            const int cnt = min(tj - hj, 32);
            __syncwarp();
            const float4* e = reinterpret_cast<const float4*>(T.job[warp][(hj + lane) & (kJobRing - 1)]);
            const float4 e0 = e[0], e1 = e[1], e2 = e[2];
            __syncwarp();
            const int o2 = __float_as_int(e2.z);
            bool slow = false;
            if (lane < cnt) {
                RawBox x, y;
                x.t = e0.x; x.p = e0.y; x.a = e0.z; x.b = e0.w; x.g = e2.x;
                y.t = e1.x; y.p = e1.y; y.a = e1.z; y.b = e1.w; y.g = e2.y;
                PairS1 q;
                ClipJob job;
                int st = pair_stage1(x, y, D, edge, !dense, &q);
                if (st == JOB_SLOW) st = pair_stage1_general(x, y, D, edge, !dense, &q);   // similarity mask / upper clamp: hi + lo form
                if (st == JOB_READY) st = pair_stage2(q, D, kind, &job);
                slow = st == JOB_SLOW;
                SPHK_CHECK((((int64_t)(o2 >> 5) * nw + wg) << 5) + (o2 & 31) < P && o2 >= 0);
                if (!slow) out[(((int64_t)(o2 >> 5) * nw + wg) << 5) + (o2 & 31)] = (st == JOB_DEAD) ? 0.0f : clip_job_iou(job, mode);
            }
            const unsigned ms = __ballot_sync(0xFFFFFFFFu, slow);
            if (slow) T.sidx[warp][(ts + __popc(ms & lt)) & (kSlowRing - 1)] = o2;
            ts += __popc(ms);
            SPHK_CHECK(ts - hs <= kSlowRing);
            hj += cnt;
        }
        const bool drained = c >= chunks && tj == hj;
        const int need = drained ? 1 : 32;             // no more fast work can come: flush the reference-order ring
        while (ts - hs >= need) {
            const int cnt = min(ts - hs, 32);
            __syncwarp();
            const int o2 = T.sidx[warp][(hs + lane) & (kSlowRing - 1)];
            __syncwarp();
            if (lane < cnt) {
                const int64_t p = (((int64_t)(o2 >> 5) * nw + wg) << 5) + (o2 & 31);
                SPHK_CHECK(p >= 0 && p < P);
                out[p] = slow_pair_iou(b1, p, b2, p, D, kind, mode, edge, dense);
            }
            hs += cnt;
        }
        if (drained) break;
    }
#ifdef SPHK_TIMELINE
    if (lane == 0 && wg < 8192) {
        g_tlw[3 * wg] = tl_start; g_tlw[3 * wg + 1] = tl_scan; g_tlw[3 * wg + 2] = tl_now();
        g_tlw_slow[wg] = (unsigned)ts;
    }
#endif
}

// ---- rbb_angle = 'project': one thread per pair, double-precision transform (sphk_math.cuh) ------------
template <int D>
__global__ void __launch_bounds__(kThreads)
k_iou_project(const float* __restrict__ b1, int64_t n1, const float* __restrict__ b2, int64_t n2, bool aligned, int kind,
              int mode, int edge, float* __restrict__ out, int64_t ld) {
    const int64_t p = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    const int64_t total = aligned ? n1 : n1 * n2;
    if (p >= total) return;
    const int64_t i = aligned ? p : p / n2, j = aligned ? p : p - i * n2;
    const RawBox x = load_box<D>(b1, i, false), y = load_box<D>(b2, j, false);
    out[aligned ? p : i * ld + j] = sph2pob_iou_pair_project(x, y, D, kind, mode, edge);
}

// ---- pairwise ----------------------------------------------------------------------------------
// key = (float bits << 32) | ~index : u64 max == (max value, lowest index); IoU >= 0 so the bit
// pattern of the float is monotone.
__device__ __forceinline__ unsigned long long pack_key(float v, uint32_t idx) {
    return ((unsigned long long)__float_as_uint(v) << 32) | (unsigned long long)(0xFFFFFFFFu - idx);
}

constexpr int kTileRows = 32;

template <int KIND, int D>
__global__ void __launch_bounds__(kThreads)
k_iou_pairwise(const float* __restrict__ rows, int64_t R, const float* __restrict__ cols, int64_t C, int mode, int edge,
               float* __restrict__ out, int64_t ld, unsigned long long* __restrict__ row_key,
               unsigned long long* __restrict__ col_key, uint32_t row_base, uint32_t col_base, int64_t col_tiles,
               bool vec_ok, bool dense) {
    __shared__ float s_row[kTileRows * 5];
    __shared__ unsigned long long s_rkey[kTileRows];
    const int64_t rt = blockIdx.x / col_tiles, ct = blockIdx.x % col_tiles;
    const int64_t r0 = rt * kTileRows;
    const int nr = (int)min((int64_t)kTileRows, R - r0);
    const int64_t col = ct * kThreads + threadIdx.x;
    const bool col_ok = col < C;
    for (int k = threadIdx.x; k < nr * D; k += kThreads) s_row[k] = __ldg(rows + r0 * D + k);
    if (threadIdx.x < kTileRows) s_rkey[threadIdx.x] = 0ull;
    __syncthreads();
    RawBox y;
    y.t = y.p = y.a = y.b = y.g = 0.0f;
    if (col_ok) y = load_box<D>(cols, col, vec_ok);
    const int lane = threadIdx.x & 31;
    float best_v = -1.0f;
    uint32_t best_r = 0;
    for (int r = 0; r < nr; ++r) {
        RawBox x;
        x.t = s_row[r * D + 0]; x.p = s_row[r * D + 1]; x.a = s_row[r * D + 2]; x.b = s_row[r * D + 3];
        x.g = (D == 5) ? s_row[r * D + 4] : 0.0f;
        float v = 0.0f;
        if (col_ok) {
            if (KIND == KIND_SPH || KIND == KIND_FOV) v = approx_iou_pair(x, y, KIND);
            else if (KIND == KIND_NAIVE) v = naive_iou_pair(x, y, D, mode);
            else if (KIND == KIND_UNBIASED) v = unbiased_iou_pair(x, y, D);
            else if (KIND == KIND_SPH2POB_LEGACY) v = sph2pob_legacy_iou_pair(x, y, mode, edge);
            else v = sph2pob_iou_pair(x, y, D, KIND, mode, edge, dense);
            if (out) out[(r0 + r) * ld + col] = v;
            if (v > best_v) { best_v = v; best_r = (uint32_t)r; }
        }
        if (row_key) {
            // warp max over the 32 columns of this warp, lowest column on ties
            const uint32_t bits = col_ok ? __float_as_uint(v) : 0u;
            const uint32_t mx = __reduce_max_sync(0xFFFFFFFFu, bits);
            const uint32_t who = __ballot_sync(0xFFFFFFFFu, col_ok && bits == mx);
            if (who != 0u && lane == __ffs(who) - 1)
                atomicMax(&s_rkey[r], pack_key(v, col_base + (uint32_t)col));
        }
    }
    if (col_key && col_ok && best_v >= 0.0f) {
        const unsigned long long key = pack_key(best_v, row_base + (uint32_t)r0 + best_r);
        if (key > col_key[col]) atomicMax(&col_key[col], key);
    }
    if (row_key) {
        __syncthreads();
        if (threadIdx.x < nr) {
            const unsigned long long key = s_rkey[threadIdx.x];
            if (key > row_key[r0 + threadIdx.x]) atomicMax(&row_key[r0 + threadIdx.x], key);
        }
    }
}

// ---- pairwise, Sph2Pob kinds: precompute + prefilter + warp-compacted evaluation -------------------
// k_box_pre: BoxRec (64 B) + BoxCull (64 B) of every row and column box, once per call, into the workspace.
// k_iou_pairwise2: CTA tile = TR rows x 256 columns, lane <-> column.  For each row the 7-FMA prefilter
// decides "exactly 0" (written straight away, coalesced) or "live"; live (row, column) pairs are
// ballot-compacted into a per-warp ring buffer in shared memory and evaluated 32 at a time, so the
// expensive clipping code always runs with full warps although only ~20-30 % of the pairs need it.
// Pairs that need the reference-order path (rare) are compacted once more into a second ring.
// Row / column max+argmax are reduced in shared memory and merged into the global packed keys with
// one atomicMax per row / column and CTA.
constexpr int kTC = 256, kRing = 64, kCand = 128, kRecStride = 20;   // record stride 20 words: conflict-free LDS.128 across lanes

template <int TR, bool SAT>
struct PairTile {
    float crec[kTC * kRecStride];
    float rrec[TR * kRecStride];
    float4 rcull[TR][4];
    float4 ccull[SAT ? kTC : 1][3];      // separating-axis stage: centre, width axis + half width, height axis + half height per column
    unsigned long long rkey[TR];
    unsigned long long ckey[kTC];
    unsigned short ring[kThreads / 32][2][kRing];             // 0: pairs for the clipper, 1: reference-order path
    unsigned short cand[SAT ? kThreads / 32 : 1][kCand];      // separating-axis stage: the circle test's survivors (two rows per scan step)
    float rtgt[TR];      // tie pass: the row maxima to compare with
    int ctie[kTC];       // tie pass: per column, the largest (row index + 1) that ties its row maximum
    unsigned long long* peer[16];   // push route: the ranks' buffers (fetched during phase 0, used in the epilogue)
};

__device__ __forceinline__ void put_rec(float* base, int i, const BoxRec& b) {
    float4* d = reinterpret_cast<float4*>(base + i * kRecStride);
    d[0] = make_float4(b.t, b.p, b.tj, b.pj);
    d[1] = make_float4(b.sp, b.cp, b.w, b.h);
    d[2] = make_float4(b.sg, b.cg, b.a, b.b);
    d[3] = make_float4(b.g, b.flag, 0.0f, 0.0f);
}

__device__ __forceinline__ void put_cull(float4* u, const BoxCull& c) {
    u[0] = make_float4(c.ux, c.uy, c.uz, c.rc);
    u[1] = make_float4(c.rs, c.bias, c.r, 0.0f);
    u[2] = make_float4(c.ex, c.ey, c.ez, c.hwm);
    u[3] = make_float4(c.fx, c.fy, c.fz, c.hhm);
}
// operands that fail every prefilter test (out-of-range rows)
__device__ __forceinline__ void put_cull_never(float4* u) {
    u[0] = make_float4(0.f, 0.f, 1.f, 0.f);
    u[1] = make_float4(0.f, -10.f, 10.f, 0.f);
    u[2] = make_float4(0.f, 0.f, 0.f, 10.f);
    u[3] = make_float4(0.f, 0.f, 0.f, 10.f);
}

template <int D>
__global__ void __launch_bounds__(kThreads)
k_box_pre(const float* __restrict__ rows, int64_t R, const float* __restrict__ cols, int64_t C, int edge,
          float4* __restrict__ rec, float4* __restrict__ cull, bool rows_vec, bool cols_vec,
          unsigned long long* __restrict__ zero_rkey, unsigned long long* __restrict__ zero_ckey, bool rows_inline,
          int64_t first) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // k_iou_pairwise2 may start its prologue
    // (first = R when the rows need nothing from this launch -- records computed in the tile kernel, keys pushed, not
    //  merged: the grid then covers the columns only)
    const int64_t i = first + (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= R + C) return;
    const bool is_row = i < R;
    // the packed max / argmax keys of the same call start from "nothing seen" (one key per box: no launch of its own)
    if (is_row) { if (zero_rkey) zero_rkey[i] = 0ull; }
    else if (zero_ckey) zero_ckey[i - R] = 0ull;
    if (is_row && rows_inline) return;      // long-row calls: k_iou_pairwise2 computes the records of its row tile itself
    const RawBox x = is_row ? load_box<D>(rows, i, rows_vec) : load_box<D>(cols, i - R, cols_vec);
    BoxRec b;
    BoxCull c;
    box_pre(x, is_row ? 1 : 2, D, edge, &b, &c);
    float4* q = rec + i * 4;
    q[0] = make_float4(b.t, b.p, b.tj, b.pj);
    q[1] = make_float4(b.sp, b.cp, b.w, b.h);
    q[2] = make_float4(b.sg, b.cg, b.a, b.b);
    q[3] = make_float4(b.g, b.flag, 0.0f, 0.0f);
    put_cull(cull + i * 4, c);
}

__device__ __forceinline__ BoxRec load_rec(const float* base, int i) {
    const float4* q = reinterpret_cast<const float4*>(base + i * kRecStride);
    const float4 q0 = q[0], q1 = q[1], q2 = q[2], q3 = q[3];
    BoxRec b;
    b.t = q0.x; b.p = q0.y; b.tj = q0.z; b.pj = q0.w;
    b.sp = q1.x; b.cp = q1.y; b.w = q1.z; b.h = q1.w;
    b.sg = q2.x; b.cg = q2.y; b.a = q2.z; b.b = q2.w;
    b.g = q3.x; b.flag = q3.y; b.pad0 = 0.0f; b.pad1 = 0.0f;
    return b;
}
__device__ __forceinline__ void stage_rec(float* base, int i, const float4* __restrict__ rec, int64_t idx, bool ok) {
    // out-of-range slots get a flagged record (never evaluated anyway)
    float4 q0 = make_float4(0.f, 90.f, 0.f, 90.f), q1 = make_float4(1.f, 0.f, 1e-2f, 1e-2f);
    float4 q2 = make_float4(0.f, 1.f, 1.f, 1.f), q3 = make_float4(0.f, 1.f, 0.f, 0.f);
    if (ok) {
        const float4* q = rec + idx * 4;
        q0 = __ldg(q); q1 = __ldg(q + 1); q2 = __ldg(q + 2); q3 = __ldg(q + 3);
    }
    float4* d = reinterpret_cast<float4*>(base + i * kRecStride);
    d[0] = q0; d[1] = q1; d[2] = q2; d[3] = q3;
}

// Prefilter of one (row, column) pair in the scan loops of the N x M kernels: the row's operands come from shared memory
// (4 x float4, BoxCull layout), the column's live in registers.
//   Test 1 (sphk_fast.cuh: pre_disjoint): cos(arc) < cos(r_g + r_p) - margin, the circumscribed circles cannot touch.
//   Test 2 (pre_outside_box, BOX): the column's centre lies outside the row box grown by the column's circumradius, along
//   one of the row box's own axes; evaluated by the lanes test 1 left alive.  It prunes the part of test 1's acceptance
//   region (a disc of radius R + r around the row box) that lies outside the grown box: ~1/3 of it for small columns
//   against large rows (RetinaNet anchors x ground truths: 69 -> 76 % of all pairs proven disjoint, 80 % have IoU = 0;
//   +9.5 % throughput), a few per cent for operands of similar size, where its ~19 instructions per pair cost more than
//   they save (1 M x 1,024 random boxes: 5.5 % of test 1's survivors pruned, +2.6 % instructions, -1.2 % throughput).
//   BOX is therefore a compile-time property of the kernel instance and the HOST picks the instance per call
//   (box_test_pays): ground truths (few rows) against anchors (many columns) get it, long-row sweeps do not.
// Out-of-range columns and the dense (measurement) mode are folded into pbias / pr.
template <bool BOX>
__device__ __forceinline__ bool prefilter_live(const float4* __restrict__ rc, const float4& pc0, float prs, float pbias, float pr) {
    const float4 g0 = rc[0];
    const float2 g1 = *reinterpret_cast<const float2*>(&rc[1]);
    const float dot = fmaf(g0.x, pc0.x, fmaf(g0.y, pc0.y, g0.z * pc0.z));
    const float thr = fmaf(g0.w, pc0.w, fmaf(-g1.x, prs, g1.y + pbias));
    bool live = !(dot < thr);
    if (BOX) {
        const float4 e = rc[2], f = rc[3];
        live = live && !pre_outside_box(e.x, e.y, e.z, e.w, f.x, f.y, f.z, f.w, pc0.x, pc0.y, pc0.z, pr);
    }
    return live;
}

struct PairOut {
    float* out;
    int64_t ld, r0, c0;
    bool want_row, want_col, tie;
    uint32_t row_base, col_base;
};

template <int TR, bool SAT>
__device__ __forceinline__ void emit_pair(PairTile<TR, SAT>& T, const PairOut& o, int r, int c, float v) {
    if (o.out) o.out[(o.r0 + r) * o.ld + o.c0 + c] = v;
    if (v > 0.0f) {
        if (o.want_row) atomicMax(&T.rkey[r], pack_key(v, o.col_base + (uint32_t)(o.c0 + c)));
        if (o.want_col) atomicMax(&T.ckey[c], pack_key(v, o.row_base + (uint32_t)(o.r0 + r)));
        // MaxIoUAssigner's gt_max_assign_all scan (max_iou_assigner.py:203-205): overlaps[i, :] == gt_max[i]
        if (o.tie && v == T.rtgt[r]) atomicMax(&T.ctie[c], (int)(o.row_base + (uint32_t)(o.r0 + r)) + 1);
    }
}

#ifdef SPHK_TIMELINE
// instrumented build (tools/timeline_probe.py): start / end globaltimer and SM id of every CTA of k_iou_pairwise2
__device__ unsigned long long g_tl[16384 * 2];
__device__ unsigned g_tl_sm[16384];
#endif

// Sharded sweeps (sph_retina_b200/sharded.py, route 'peer'): the row keys leave for the peers from inside the compute
// kernel.  Instead of merging the column tiles' maxima of a row into one global key (atomicMax), every CTA stores the 32
// keys of ITS tile -- the maximum over its 256 columns, global argmax index included -- as one of ceil(C / 256) partial
// arrays, into the buffer of EVERY rank (its own included): plain 8-byte stores, one per thread for up to 8 ranks, over
// NVLink for the peers, fire-and-forget.  No atomics, no zero-fill, no arrival counting, no fence: the transfer runs
// under the rest of the launch, and the exchange step that follows takes the maximum over the partials from local memory.
struct KeyPush {
    unsigned long long* const* bufs;   // device array of the ranks' buffer base pointers; nullptr: nothing is pushed
    long long off;                     // element offset, in EVERY rank's buffer, of partial array 0 of this rank's rows
    long long part_stride;             // elements between the partial arrays of successive column tiles
    int world;
};

// CULL: what proves pairs disjoint before the clipper sees them.  0: the circle test of the scan loop alone; 1: + the
// box-frame test in the scan loop (ground truths x anchors); 2: + the separating-axis test (pre_sat_disjoint) as a stage
// of its own between the scan and the clipper -- the circle test's survivors are queued, tested 32 at a time with full
// warps, and only what is left goes on (operands of similar size: a third of the survivors end there).
template <int D, int TR, int CULL>
__global__ void __launch_bounds__(kThreads, 4)      // 4 CTAs per SM (64 registers): 3 cost 22 % of the throughput
k_iou_pairwise2(const float* __restrict__ rows, int64_t R, const float* __restrict__ cols, int64_t C,
                const float4* __restrict__ rec, const float4* __restrict__ cull, int kind, int mode, int edge,
                float* __restrict__ out, int64_t ld, unsigned long long* __restrict__ row_key,
                unsigned long long* __restrict__ col_key, uint32_t row_base, uint32_t col_base, int flags,
                const float* __restrict__ row_target, int* __restrict__ col_tie,
                const int32_t* __restrict__ row_offsets, int64_t col_stride, float* __restrict__ tile_rmax, const KeyPush push) {
    constexpr bool BOX = CULL == 1, SAT = CULL == 2;
    __shared__ __align__(16) PairTile<TR, SAT> T;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool dense = (flags & 1) != 0;                               // sphk_set_dense (measurement)
#ifdef SPHK_TIMELINE
    unsigned long long tl0 = 0;
    if (tid == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tl0));
#endif
    // grid = (row tiles, column tiles).  Column tiles are taken heaviest-first: RetinaNet-style anchor lists
    // end with the coarse pyramid levels, whose huge anchors overlap every GT, so the tail of the launch is
    // made of the light tiles.
    const uint32_t rt = blockIdx.x, ct = gridDim.y - 1u - blockIdx.y;
    // Batch mode (blockIdx.z = image): the rows are the concatenated GT lists of all images, image b owning rows
    // [row_offsets[b], row_offsets[b+1]); row tiles never straddle images, every column-side result (max/argmax
    // keys, ties) is per image at offset b * col_stride, and reported row indices are local to the image.
    int64_t r_begin = 0, r_end = R;
    if (row_offsets) {
        r_begin = row_offsets[blockIdx.z];
        r_end = row_offsets[blockIdx.z + 1];
        if ((int64_t)rt * TR >= r_end - r_begin) return;
        if (col_key) col_key += blockIdx.z * col_stride;
        if (col_tie) col_tie += blockIdx.z * col_stride;
        row_base -= (uint32_t)r_begin;
    }
    PairOut o;
    o.out = out; o.ld = ld; o.r0 = r_begin + (int64_t)rt * TR; o.c0 = (int64_t)ct * kTC;
    o.want_row = row_key != nullptr || push.bufs != nullptr; o.want_col = col_key != nullptr; o.tie = col_tie != nullptr;
    o.row_base = row_base; o.col_base = col_base;
    const int nr = (int)min((int64_t)TR, r_end - o.r0);
    const bool col_ok = o.c0 + tid < C;
    if (out) {
        // zero-fill this warp's [nr x 32] part of the matrix; live pairs overwrite their entry later
        // (ordered by the __syncthreads() below and the __syncwarp() in front of every batch)
        float* base = out + o.r0 * ld + o.c0 + warp * 32;
        const bool vec = ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15u) == 0) && (o.c0 + warp * 32 + 32 <= C);
        if (vec) {
            for (int rr = lane >> 3; rr < nr; rr += 4)
                *reinterpret_cast<float4*>(base + rr * ld + (lane & 7) * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
        } else if (col_ok) {
            for (int rr = 0; rr < nr; ++rr) base[rr * ld + lane] = 0.0f;
        }
    }
    // Programmatic dependent launch: everything above overlaps the tail of k_box_pre; its records are
    // needed from here on.  A dependent of THIS launch (the exchange step of the sharded sweep) may take the slots the
    // last CTAs free; it waits for the completion of this grid itself before it touches anything.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (o.tie && tile_rmax) {
        // Tie pass of the assigner: an entry can equal its row's maximum only inside a tile whose own maximum of that
        // row IS the maximum.  The max / argmax pass left every tile's row maxima in tile_rmax[column tile][row]: all
        // but a handful of tiles leave here (a disabled row has target -1, a tile maximum is >= 0).
        const bool mine = tid < nr && tile_rmax[(int64_t)ct * R + o.r0 + tid] == __ldg(row_target + o.r0 + tid);
        if (!__syncthreads_or(mine ? 1 : 0)) return;
    }
    // ---- phase 0: stage the tile's records (columns: records -> shared memory, cull operands -> registers)
    stage_rec(T.crec, tid, rec + R * 4, o.c0 + tid, col_ok);
    T.ckey[tid] = 0ull;
    float4 pc0 = make_float4(0.f, 0.f, 1.f, 0.f), pc1 = make_float4(0.f, -10.f, 10.f, 0.f);
    if (col_ok) {
        const float4* u = cull + (R + o.c0 + tid) * 4;
        pc0 = __ldg(u); pc1 = __ldg(u + 1);
        if (SAT) {
            // (dense, the measurement mode: half sizes of 10 switch the separating-axis test off, as for flagged boxes)
            float4 e2 = __ldg(u + 2), f2 = __ldg(u + 3);
            if (dense) { e2.w = 10.0f; f2.w = 10.0f; }
            T.ccull[tid][0] = pc0; T.ccull[tid][1] = e2; T.ccull[tid][2] = f2;
        }
    }
    if (tid < TR) {
        const bool ok = tid < nr;
        if ((flags & 4) && ok) {
            // long-row calls (sweeps: R >> C): the records of the tile's rows are computed here, straight from the raw
            // boxes into shared memory (0.4 % more instructions per CTA) instead of being written by k_box_pre and read
            // back once per column tile -- DRAM traffic of the launch = the boxes in + the keys out
            const RawBox x = load_box<D>(rows, o.r0 + tid, false);
            BoxRec b;
            BoxCull c;
            box_pre(x, 1, D, edge, &b, &c);
            put_rec(T.rrec, tid, b);
            put_cull(T.rcull[tid], c);
        } else {
            stage_rec(T.rrec, tid, rec, o.r0 + tid, ok);
            if (ok) {
                const float4* u = cull + (o.r0 + tid) * 4;
                T.rcull[tid][0] = __ldg(u); T.rcull[tid][1] = __ldg(u + 1);
                if (BOX || SAT) { T.rcull[tid][2] = __ldg(u + 2); T.rcull[tid][3] = __ldg(u + 3); }
            } else {
                put_cull_never(T.rcull[tid]);
            }
        }
        T.rkey[tid] = 0ull;
        T.rtgt[tid] = (o.tie && ok) ? __ldg(row_target + o.r0 + tid) : -1.0f;
    }
    T.ctie[tid] = 0;
    if (push.bufs && tid >= kThreads - 16 && tid - (kThreads - 16) < push.world) T.peer[tid - (kThreads - 16)] = push.bufs[tid - (kThreads - 16)];
    __syncthreads();
    // ---- phase 1: prefilter + compaction.  This thread's column is tid (= warp * 32 + lane).
    // Out-of-range columns and the dense (measurement) mode are folded into the bias term of the test.
    const float pbias = col_ok ? (dense ? -1e30f : pc1.y) : 1e30f;
    const float pr = dense ? 1e30f : pc1.z;      // dense (measurement) mode: the box-frame test never fires either
    (void)pr;
    int hf = 0, tf = 0, hs = 0, ts = 0;          // ring head / tail counters (fast, slow)
    int ha = 0, ta = 0;                          // SAT: ring of the circle test's survivors
    const unsigned lt = (1u << lane) - 1u;
    // Alternate between a tight scan phase (prefilter rows until 32 live pairs are queued or the rows are used up)
    // and ONE batch site per stage (so that the expensive code exists once); the rings are flushed when the rows are done.
    int r = 0;
#pragma unroll 1
    for (;;) {
        if (SAT) {
            unsigned short* const cand = T.cand[SAT ? warp : 0];
            // two rows per step (<= 31 queued + 64 pushed <= kCand), then the odd last row
#pragma unroll 1
            while (r + 1 < nr && ta - ha < 32) {
                const bool l0 = prefilter_live<false>(T.rcull[r], pc0, pc1.x, pbias, pr);
                const bool l1 = prefilter_live<false>(T.rcull[r + 1], pc0, pc1.x, pbias, pr);
                const unsigned m0 = __ballot_sync(0xFFFFFFFFu, l0), m1 = __ballot_sync(0xFFFFFFFFu, l1);
                const int n0 = __popc(m0);
                if (l0) cand[(ta + __popc(m0 & lt)) & (kCand - 1)] = (unsigned short)((r << 5) | lane);
                if (l1) cand[(ta + n0 + __popc(m1 & lt)) & (kCand - 1)] = (unsigned short)(((r + 1) << 5) | lane);
                ta += n0 + __popc(m1);
                SPHK_CHECK(ta - ha <= kCand);
                r += 2;
            }
            if (r + 1 == nr && ta - ha < 32) {
                const bool live = prefilter_live<false>(T.rcull[r], pc0, pc1.x, pbias, pr);
                const unsigned m = __ballot_sync(0xFFFFFFFFu, live);
                if (live) cand[(ta + __popc(m & lt)) & (kCand - 1)] = (unsigned short)((r << 5) | lane);
                ta += __popc(m);
                ++r;
            }
            if (ta > ha) {                         // here: >= 32 candidates, or the rows are done
                const int cnt = min(ta - ha, 32);
                __syncwarp();
                const int e = cand[(ha + lane) & (kCand - 1)];
                __syncwarp();
                bool live = false;
                if (lane < cnt) {
                    SPHK_CHECK((e >> 5) < nr);
                    const float4* g = T.rcull[e >> 5];
                    const float4* q = T.ccull[warp * 32 + (e & 31)];
                    const float4 g0 = g[0], g2 = g[2], g3 = g[3], q0 = q[0], q1 = q[1], q2 = q[2];
                    live = !pre_sat_disjoint(g0.x, g0.y, g0.z, g2.x, g2.y, g2.z, g2.w, g3.x, g3.y, g3.z, g3.w,
                                                      q0.x, q0.y, q0.z, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w);
                }
                const unsigned m = __ballot_sync(0xFFFFFFFFu, live);
                if (live) T.ring[warp][0][(tf + __popc(m & lt)) & (kRing - 1)] = (unsigned short)e;
                tf += __popc(m);
                SPHK_CHECK(tf - hf <= kRing);
                ha += cnt;
            }
        } else {
#pragma unroll 1
            while (r < nr && tf - hf < 32) {
                const bool live = prefilter_live<BOX>(T.rcull[r], pc0, pc1.x, pbias, pr);
                const unsigned m = __ballot_sync(0xFFFFFFFFu, live);
                if (live) T.ring[warp][0][(tf + __popc(m & lt)) & (kRing - 1)] = (unsigned short)((r << 5) | lane);
                tf += __popc(m);
                SPHK_CHECK(tf - hf <= kRing);
                ++r;
            }
        }
        const bool rows_done = r >= nr && ta == ha;        // nothing more can enter the fast ring
        if (tf - hf >= 32 || (rows_done && tf > hf)) {
            const int cnt = min(tf - hf, 32);
            __syncwarp();
            const int e = T.ring[warp][0][(hf + lane) & (kRing - 1)];
            __syncwarp();
            bool slow = false;
            // (a row whose box is not finite fails no test of the scan -- NaN compares false -- and queues all 256 columns of
            //  the tile, those past the end of the operand included: they are dropped here, not evaluated)
            if (lane < cnt && o.c0 + warp * 32 + (e & 31) < C) {
                const int rr = e >> 5, c = warp * 32 + (e & 31);
                SPHK_CHECK(rr < nr && o.c0 + c < C);
                float v;
                slow = !pair_fast(load_rec(T.rrec, rr), load_rec(T.crec, c), D, kind, mode, &v);
                if (!slow) emit_pair(T, o, rr, c, v);
            }
            const unsigned ms = __ballot_sync(0xFFFFFFFFu, slow);
            if (ms) {                                   // (1 pair in 10^5 on random boxes: keep the push out of the common path)
                if (slow) T.ring[warp][1][(ts + __popc(ms & lt)) & (kRing - 1)] = (unsigned short)e;
                ts += __popc(ms);
                SPHK_CHECK(ts - hs <= kRing);
            }
            hf += cnt;
        }
        const bool drained = rows_done && tf == hf;
        const int need = drained ? 1 : 32;          // no more fast work can come: flush the reference-order ring
        while (ts - hs >= need) {
            const int cnt = min(ts - hs, 32);
            __syncwarp();
            const int e = T.ring[warp][1][(hs + lane) & (kRing - 1)];
            __syncwarp();
            if (lane < cnt) {
                const int rr = e >> 5, c = warp * 32 + (e & 31);
                SPHK_CHECK(rr < nr && o.c0 + c < C);
                emit_pair(T, o, rr, c, slow_pair_iou(rows, o.r0 + rr, cols, o.c0 + c, D, kind, mode, edge, dense));
            }
            hs += cnt;
        }
        if (drained) break;
    }
#ifdef SPHK_TIMELINE
    __syncthreads();
    if (tid == 0) {
        unsigned long long tl1; unsigned smid;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tl1));
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        const unsigned b = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
        if (b < 16384) { g_tl[2 * b] = tl0; g_tl[2 * b + 1] = tl1; g_tl_sm[b] = smid; }
    }
#endif
    // ---- merge the tile's max/argmax into the global keys
    if (o.want_row || o.want_col || o.tie) {
        __syncthreads();
        if (o.tie && col_ok && T.ctie[tid] > 0) atomicMax(&col_tie[o.c0 + tid], T.ctie[tid]);
        if (push.bufs) {
            // this tile's row keys as partial array `ct`, into every rank's buffer (thread -> (rank, row))
            const long long at = push.off + (long long)ct * push.part_stride + o.r0;
            for (int j = tid; j < TR * push.world; j += kThreads) {
                const int d = j / TR, i = j % TR;
                SPHK_CHECK(d < push.world && (int64_t)ct * push.part_stride + o.r0 + i >= 0);
                if (i < nr) T.peer[d][at + i] = T.rkey[i];
            }
        } else if (o.want_row && tid < nr) {
            const unsigned long long key = T.rkey[tid];
            if (key != 0ull && key > row_key[o.r0 + tid]) atomicMax(&row_key[o.r0 + tid], key);
            if (tile_rmax) tile_rmax[(int64_t)ct * R + o.r0 + tid] = __uint_as_float((uint32_t)(key >> 32));
        }
        if (o.want_col && col_ok) {
            const unsigned long long key = T.ckey[tid];
            if (key != 0ull && key > col_key[o.c0 + tid]) atomicMax(&col_key[o.c0 + tid], key);
        }
    }
}

// ---- measurement: how many pairs of an N x M call survive the prefilter of the scan loops (bench.py reports the
// early-out rate next to the throughput).  Same records, same prefilter_live() as k_iou_pairwise2.
__global__ void __launch_bounds__(kThreads)
k_prefilter_count(int64_t R, int64_t C, const float4* __restrict__ cull, unsigned long long* __restrict__ live_count, int cull_kind) {
    __shared__ float4 s_rc[32][4];
    const int64_t r0 = (int64_t)blockIdx.x * 32, col = (int64_t)blockIdx.y * kThreads + threadIdx.x;
    const int nr = (int)min((int64_t)32, R - r0);
    if (threadIdx.x < 32 * 4) {
        const int rr = threadIdx.x >> 2, q = threadIdx.x & 3;
        if (rr < nr) s_rc[rr][q] = __ldg(cull + (r0 + rr) * 4 + q);
    }
    __syncthreads();
    unsigned n = 0;
    if (col < C) {
        const float4* u = cull + (R + col) * 4;
        const float4 pc0 = __ldg(u), pc1 = __ldg(u + 1), pc2 = __ldg(u + 2), pc3 = __ldg(u + 3);
        for (int r = 0; r < nr; ++r) {
            bool live = cull_kind == 1 ? prefilter_live<true>(s_rc[r], pc0, pc1.x, pc1.y, pc1.z) : prefilter_live<false>(s_rc[r], pc0, pc1.x, pc1.y, pc1.z);
            if (live && cull_kind == 2) {
                const float4 g0 = s_rc[r][0], g2 = s_rc[r][2], g3 = s_rc[r][3];
                live = !pre_sat_disjoint(g0.x, g0.y, g0.z, g2.x, g2.y, g2.z, g2.w, g3.x, g3.y, g3.z, g3.w,
                                         pc0.x, pc0.y, pc0.z, pc2.x, pc2.y, pc2.z, pc2.w, pc3.x, pc3.y, pc3.z, pc3.w);
            }
            n += live ? 1u : 0u;
        }
    }
    n = __reduce_add_sync(0xFFFFFFFFu, n);
    if ((threadIdx.x & 31) == 0 && n) atomicAdd(live_count, (unsigned long long)n);
}

// ---- few rows x many columns in ONE launch ----------------------------------------------------------------------
// SphOverlaps2D(gt[R <= 32], anchors[C]) -> [R, C] is the call MaxIoUAssigner makes once per image
// (mmdet/core/bbox/assigners/max_iou_assigner.py:113).  At this size the general path spends a tenth of the call in
// k_box_pre and the hand-over to k_iou_pairwise2; here a CTA owns ALL rows and 64 columns, so it can compute the
// per-box records itself (every column once, the <= 32 rows once per CTA: +0.2 % of the instructions), straight into
// shared memory -- no workspace, no second kernel.  Warp w works on column group w & 1 (32 columns, lane <-> column,
// coalesced row segments) and row group w >> 1 (a quarter of the rows); scan, compaction rings, fast path and the
// out-of-line reference-order path are those of k_iou_pairwise2.
constexpr int kFC = 64;
struct RowsTile {
    float crec[kFC * kRecStride];
    float rrec[32 * kRecStride];
    float4 ccull[kFC][2];
    float4 rcull[32][4];
    unsigned short ring[kThreads / 32][2][kRing];
};

template <int D>
__global__ void __launch_bounds__(kThreads)
k_iou_rows32(const float* __restrict__ rows, int R, const float* __restrict__ cols, int64_t C, int kind, int mode, int edge,
             float* __restrict__ out, int64_t ld, int flags, bool rows_vec, bool cols_vec) {
    __shared__ __align__(16) RowsTile T;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool dense = (flags & 1) != 0;
#ifdef SPHK_TIMELINE
    unsigned long long tl0 = 0;
    if (tid == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tl0));
#endif
    // column tiles heaviest-first (anchor lists end with the coarse pyramid levels), as in k_iou_pairwise2
    const int64_t c0 = (int64_t)(gridDim.x - 1u - blockIdx.x) * kFC;
    const int cg = warp & 1, rpw = (R + 3) >> 2;
    const int r_lo = min(R, (warp >> 1) * rpw), r_hi = min(R, r_lo + rpw);
    const int cl = cg * 32 + lane;
    const bool col_ok = c0 + cl < C;
    // Programmatic dependent launch (launch_rows32): the CTAs of this launch become resident while the previous kernel
    // of the stream drains and start the moment it has completed -- no launch gap between the per-image calls of
    // MaxIoUAssigner.  Nothing is read or written before griddepcontrol.wait: whatever the previous kernel produced
    // (operands included) is complete and visible, and an output buffer the allocator recycled is not touched early.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    {
        // zero-fill this warp's [rows x 32] part of the matrix; live pairs overwrite their entry later (same warp:
        // ordered by the __syncthreads() below and the __syncwarp() in front of every batch)
        float* base = out + c0 + cg * 32;
        const bool vec = ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15u) == 0) && (c0 + cg * 32 + 32 <= C);
        if (vec) {
            for (int rr = r_lo + (lane >> 3); rr < r_hi; rr += 4)
                *reinterpret_cast<float4*>(base + rr * ld + (lane & 7) * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
        } else if (col_ok) {
            for (int rr = r_lo; rr < r_hi; ++rr) base[rr * ld + lane] = 0.0f;
        }
    }
    // ---- phase 0: per-box records (sphk_fast.cuh: box_pre) by threads 0..63 (columns) and 64..95 (rows)
    if (tid < kFC + 32) {
        const bool is_col = tid < kFC;
        const int i = is_col ? tid : tid - kFC;
        const bool ok = is_col ? (c0 + i < C) : (i < R);
        float* recs = is_col ? T.crec : T.rrec;
        float4* cu = is_col ? T.ccull[i] : T.rcull[i];
        if (ok) {
            const RawBox x = is_col ? load_box<D>(cols, c0 + i, cols_vec) : load_box<D>(rows, i, rows_vec);
            BoxRec b;
            BoxCull c;
            box_pre(x, is_col ? 2 : 1, D, edge, &b, &c);
            put_rec(recs, i, b);
            cu[0] = make_float4(c.ux, c.uy, c.uz, c.rc);
            cu[1] = make_float4(c.rs, c.bias, c.r, 0.0f);
            if (!is_col) {
                cu[2] = make_float4(c.ex, c.ey, c.ez, c.hwm);
                cu[3] = make_float4(c.fx, c.fy, c.fz, c.hhm);
            }
        } else {      // never evaluated: a flagged record and operands that fail every test
            float4* d = reinterpret_cast<float4*>(recs + i * kRecStride);
            d[0] = make_float4(0.f, 90.f, 0.f, 90.f); d[1] = make_float4(1.f, 0.f, 1e-2f, 1e-2f);
            d[2] = make_float4(0.f, 1.f, 1.f, 1.f);   d[3] = make_float4(0.f, 1.f, 0.f, 0.f);
            cu[0] = make_float4(0.f, 0.f, 1.f, 0.f);
            cu[1] = make_float4(0.f, -10.f, 10.f, 0.f);
            if (!is_col) {
                cu[2] = make_float4(0.f, 0.f, 0.f, 10.f);
                cu[3] = make_float4(0.f, 0.f, 0.f, 10.f);
            }
        }
    }
    __syncthreads();
    const float4 pc0 = T.ccull[cl][0], pc1 = T.ccull[cl][1];
    // ---- phase 1: prefilter + compaction + batches, exactly as k_iou_pairwise2 (this thread's column is cl)
    const float pbias = col_ok ? (dense ? -1e30f : pc1.y) : 1e30f;
    const float pr = (dense || (flags & 2)) ? 1e30f : pc1.z;     // (bit 1: tuning build only -- test 2 never fires)
    int hf = 0, tf = 0, hs = 0, ts = 0;
    const unsigned lt = (1u << lane) - 1u;
    int r = r_lo;
#pragma unroll 1
    for (;;) {
#pragma unroll 1
        while (r < r_hi && tf - hf < 32) {
            const bool live = prefilter_live<true>(T.rcull[r], pc0, pc1.x, pbias, pr);     // few GT rows x many anchors: test 2 pays
            const unsigned m = __ballot_sync(0xFFFFFFFFu, live);
            if (live) T.ring[warp][0][(tf + __popc(m & lt)) & (kRing - 1)] = (unsigned short)((r << 5) | lane);
            tf += __popc(m);
            SPHK_CHECK(tf - hf <= kRing);
            ++r;
        }
        const bool rows_done = r >= r_hi;
        if (tf > hf) {
            const int cnt = min(tf - hf, 32);
            __syncwarp();
            const int e = T.ring[warp][0][(hf + lane) & (kRing - 1)];
            __syncwarp();
            bool slow = false;
            if (lane < cnt && c0 + cg * 32 + (e & 31) < C) {      // (columns past the end, queued by a non-finite row, are dropped)
                const int rr = e >> 5, c = cg * 32 + (e & 31);
                SPHK_CHECK(rr >= r_lo && rr < r_hi && c0 + c < C);
                float v;
                slow = !pair_fast(load_rec(T.rrec, rr), load_rec(T.crec, c), D, kind, mode, &v);
                if (!slow) out[rr * ld + c0 + c] = v;
            }
            const unsigned ms = __ballot_sync(0xFFFFFFFFu, slow);
            if (ms) {                                   // (1 pair in 10^5 on random boxes: keep the push out of the common path)
                if (slow) T.ring[warp][1][(ts + __popc(ms & lt)) & (kRing - 1)] = (unsigned short)e;
                ts += __popc(ms);
                SPHK_CHECK(ts - hs <= kRing);
            }
            hf += cnt;
        }
        const bool drained = rows_done && tf == hf;
        const int need = drained ? 1 : 32;
        while (ts - hs >= need) {
            const int cnt = min(ts - hs, 32);
            __syncwarp();
            const int e = T.ring[warp][1][(hs + lane) & (kRing - 1)];
            __syncwarp();
            if (lane < cnt) {
                const int rr = e >> 5, c = cg * 32 + (e & 31);
                SPHK_CHECK(rr >= r_lo && rr < r_hi && c0 + c < C);
                out[rr * ld + c0 + c] = slow_pair_iou(rows, rr, cols, c0 + c, D, kind, mode, edge, dense);
            }
            hs += cnt;
        }
        if (drained) break;
    }
#ifdef SPHK_TIMELINE
    __syncthreads();
    if (tid == 0) {
        unsigned long long tl1; unsigned smid;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tl1));
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        if (blockIdx.x < 16384) { g_tl[2 * blockIdx.x] = tl0; g_tl[2 * blockIdx.x + 1] = tl1; g_tl_sm[blockIdx.x] = smid; }
    }
#endif
}

// ---- MaxIoUAssigner without the matrix (mmdet/core/bbox/assigners/max_iou_assigner.py:135-220) ------------
// k_assign_targets: per GT, from the packed row key: its best overlap; for gt_max_assign_all the value the
//   tie pass compares with (-1 disables the row) and, for a GT whose best overlap is exactly 0, its claim on
//   every anchor of the image; otherwise its claim on its argmax anchor.  Later GTs override earlier ones.
__global__ void __launch_bounds__(kThreads)
k_assign_targets(const unsigned long long* __restrict__ row_key, int64_t sumK, const int32_t* __restrict__ offsets,
                 int batch, int64_t N, float min_pos_iou, bool assign_all, float* __restrict__ target,
                 int* __restrict__ zero_row, int* __restrict__ last) {
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= sumK) return;
    int b = 0;
    while (b + 1 < batch && i >= offsets[b + 1]) ++b;
    const int local = (int)(i - offsets[b]);
    const unsigned long long k = row_key[i];
    const float gmax = (k == 0ull) ? 0.0f : __uint_as_float((uint32_t)(k >> 32));
    const int64_t garg = (k == 0ull) ? 0 : (int64_t)(0xFFFFFFFFu - (uint32_t)(k & 0xFFFFFFFFull));
    const bool valid = gmax >= min_pos_iou;
    if (assign_all) {
        target[i] = (valid && gmax > 0.0f) ? gmax : -1.0f;
        if (valid && !(gmax > 0.0f)) atomicMax(&zero_row[b], local + 1);
    } else if (valid) {
        atomicMax(&last[(int64_t)b * N + garg], local + 1);
    }
}

// k_assign_epilogue: per (image, anchor): thresholds, low-quality override, labels.
__global__ void __launch_bounds__(kThreads)
k_assign_epilogue(const unsigned long long* __restrict__ col_key, const int* __restrict__ last,
                  const int* __restrict__ zero_row, const int32_t* __restrict__ offsets, int batch, int64_t N,
                  float pos_thr, float neg_lo, float neg_hi, bool match_low_quality, bool assign_all,
                  const int64_t* __restrict__ gt_labels, int64_t* __restrict__ gt_inds, float* __restrict__ max_overlaps,
                  int64_t* __restrict__ labels) {
    const int64_t p = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (p >= (int64_t)batch * N) return;
    const int b = (int)(p / N);
    const int num_gts = offsets[b + 1] - offsets[b];
    const unsigned long long k = col_key[p];
    const float m = (k == 0ull) ? 0.0f : __uint_as_float((uint32_t)(k >> 32));
    const int arg = (k == 0ull) ? 0 : (int)(0xFFFFFFFFu - (uint32_t)(k & 0xFFFFFFFFull));
    int64_t a = -1;
    if (num_gts == 0) {
        a = 0;                                            // :158-160 no ground truth: everything is background
    } else {
        if (m >= neg_lo && m < neg_hi) a = 0;             // :178-184
        if (m >= pos_thr) a = arg + 1;                    // :187-188
        if (match_low_quality) {                          // :190-207
            int l = last[p];
            if (assign_all) l = max(l, zero_row[b]);
            if (l > 0) a = l;
        }
    }
    gt_inds[p] = a;
    max_overlaps[p] = m;
    if (labels) labels[p] = (a > 0 && gt_labels) ? gt_labels[offsets[b] + a - 1] : -1;
}

__global__ void k_fill_keys(unsigned long long* keys, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keys[i] = 0ull;
}

// unpack (max, argmax); a key that was never raised (no column/row at all) reports (0, base).
// rows [0, nr) from keys_a, then columns [0, nc) from keys_b, in one launch
__global__ void k_unpack_keys2(const unsigned long long* __restrict__ keys_a, int64_t na, float* __restrict__ vmax_a,
                               int32_t* __restrict__ arg_a, uint32_t base_a, const unsigned long long* __restrict__ keys_b,
                               int64_t nb, float* __restrict__ vmax_b, int32_t* __restrict__ arg_b, uint32_t base_b) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= na + nb) return;
    const bool second = i >= na;
    if (second) i -= na;
    const unsigned long long k = second ? keys_b[i] : keys_a[i];
    float* vmax = second ? vmax_b : vmax_a;
    int32_t* arg = second ? arg_b : arg_a;
    const uint32_t base = second ? base_b : base_a;
    if (vmax) vmax[i] = (k == 0ull) ? 0.0f : __uint_as_float((uint32_t)(k >> 32));
    if (arg) arg[i] = (k == 0ull) ? (int32_t)base : (int32_t)(0xFFFFFFFFu - (uint32_t)(k & 0xFFFFFFFFull));
}

// ---- row-sharded N x M: what follows the all-gather of the shards' packed keys (sph_retina_b200/sharded.py) ---------
// gathered = [world][cap + n_short] keys: shard s holds the keys of its slice of the LONG operand (global rows
// [lo_s, hi_s), balanced contiguous split: the first n_long % world shards own one row more; entries past its slice
// are padding) followed by its view of the SHORT operand's keys.  One launch writes, for the whole long operand,
// (max, argmax over the short set) in global order, and for the short operand the maximum over the shards
// (max value, then lowest global index: integer max of the packed keys).  argmax is written as int64
// (torch.max's index type, mmdet/core/bbox/assigners/max_iou_assigner.py:173-176); key 0 = "no positive overlap"
// reads as (0.0, index 0).
__global__ void __launch_bounds__(kThreads)
k_unpack_gathered(const unsigned long long* __restrict__ gathered, int world, int64_t n_long, int64_t n_short, int64_t cap,
                  float* __restrict__ long_max, int64_t* __restrict__ long_arg, float* __restrict__ short_max,
                  int64_t* __restrict__ short_arg) {
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    const int64_t stride = cap + n_short;
    unsigned long long k = 0ull;
    if (i < n_long) {
        const int64_t base = n_long / world, extra = n_long % world;
        const int64_t cut = extra * (base + 1);                   // rows owned by the shards that hold base + 1 rows
        const int64_t s = (i < cut) ? i / (base + 1) : extra + (i - cut) / (base > 0 ? base : 1);
        const int64_t lo = s * base + (s < extra ? s : extra);
        SPHK_CHECK(s >= 0 && s < world && i - lo >= 0 && i - lo < cap);
        k = gathered[s * stride + (i - lo)];
        long_max[i] = (k == 0ull) ? 0.0f : __uint_as_float((uint32_t)(k >> 32));
        long_arg[i] = (k == 0ull) ? 0 : (int64_t)(0xFFFFFFFFu - (uint32_t)(k & 0xFFFFFFFFull));
    } else if (i < n_long + n_short) {
        const int64_t j = i - n_long;
        for (int s = 0; s < world; ++s) {
            const unsigned long long v = gathered[s * stride + cap + j];
            k = v > k ? v : k;
        }
        short_max[j] = (k == 0ull) ? 0.0f : __uint_as_float((uint32_t)(k >> 32));
        short_arg[j] = (k == 0ull) ? 0 : (int64_t)(0xFFFFFFFFu - (uint32_t)(k & 0xFFFFFFFFull));
    }
}

// ---- the same, with the gather fused into the compute and unpack launches: peer stores / loads over NVLink, no NCCL ------
// Every rank keeps the key blocks of ALL ranks in a SYMMETRIC buffer (same layout on every GPU of the node, mapped into
// every process: torch.distributed._symmetric_memory):
//     [ parity 0: world slots | parity 1: world slots | flags[world] ],   slot s = the block of rank s =
//     [ `parts` arrays of cap anchor keys | n_short ground-truth keys ]
// A step is
//   compute kernel of rank r:  'push' (sphk_iou_pairwise_keys_push, parts = column tiles): every CTA stores the keys of its
//       tile into slot r of EVERY rank's buffer while the launch is still running (KeyPush in k_iou_pairwise2);
//       'pull' (sphk_iou_pairwise_keys into slot r of the rank's OWN buffer, parts = 1): nothing leaves the GPU yet;
//   ->  this kernel:
//     1. block 0 tells every peer "my step s is complete" (a release store of s into flags[rank] of the peer's buffer:
//        the compute kernel precedes this launch in the stream, so its local and remote writes are visible before the
//        flag is);
//     2. every CTA waits until the flags of all peers in the LOCAL buffer have reached s (acquire loads);
//     3. the anchors' keys are read from the local slots (push: maximum over the partial arrays) or straight from their
//        owners' buffers (pull: coalesced 8-byte peer loads), the ground truths' keys (8 B x n_short per rank, final only
//        when the owner's kernel has ended) from the owners' slots, and everything is unpacked as k_unpack_gathered does.
// No collective kernel, no staging copy, no second launch.  The parities alternate: a rank that runs ahead by one step
// writes (and pushes into) the other parity, and it cannot run ahead by two (its step s + 1 waits for every peer's flag
// s + 1, which a peer raises only after its own step-s reads).  A rank that never shows up would leave the others
// spinning: after ~5 s the kernel traps, so a broken job fails instead of hanging the GPUs.
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__global__ void __launch_bounds__(kThreads)
k_unpack_peers(unsigned long long* const* __restrict__ peer_bufs, int rank, int world, unsigned long long step,
               int64_t block_offset, int64_t flag_offset, int64_t n_long, int64_t n_short, int64_t cap, int parts, bool long_pushed,
               float* __restrict__ long_max, int64_t* __restrict__ long_arg, float* __restrict__ short_max,
               int64_t* __restrict__ short_arg) {
    __shared__ const unsigned long long* s_blk[16];     // slot s in the buffer of rank s
    const int64_t stride = parts * cap + n_short;
    // launched with programmatic stream serialization: the CTAs may be resident while the compute kernel drains; nothing
    // of this step is read, written or signalled before that kernel and its memory operations are complete
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (threadIdx.x < world) {
        unsigned long long* const peer = peer_bufs[threadIdx.x];
        s_blk[threadIdx.x] = peer + block_offset + threadIdx.x * stride;
        if (blockIdx.x == 0) {
            __threadfence_system();
            st_release_sys(peer + flag_offset + rank, step);
        }
        const unsigned long long* mine = peer_bufs[rank] + flag_offset + threadIdx.x;
        unsigned long long t0 = 0;
        unsigned spins = 0;
        while (ld_acquire_sys(mine) < step) {
            if ((++spins & 1023u) == 0u) {
                unsigned long long now;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                if (t0 == 0) t0 = now;
                else if (now - t0 > 5000000000ull) __trap();      // a peer never arrived: fail, do not hang
            }
            __nanosleep(64);
        }
    }
    __syncthreads();
    // bounded grid (the flags are polled once per CTA, not once per 256 keys), grid-stride over the keys
    const int64_t base = n_long / world, extra = n_long % world;
    const int64_t cut = extra * (base + 1);
    const unsigned long long* const local = peer_bufs[rank] + block_offset;
#pragma unroll 4
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n_long + n_short; i += (int64_t)gridDim.x * kThreads) {
        unsigned long long k = 0ull;
        if (i < n_long) {
            const int64_t s = (i < cut) ? i / (base + 1) : extra + (i - cut) / (base > 0 ? base : 1);
            const int64_t lo = s * base + (s < extra ? s : extra);
            SPHK_CHECK(s >= 0 && s < world && i - lo >= 0 && i - lo < cap);
            const unsigned long long* src = (long_pushed ? local + s * stride : s_blk[s]) + (i - lo);
            k = __ldcg(src);
            for (int t = 1; t < parts; ++t) {                       // push: one partial array per column tile of the compute kernel
                const unsigned long long v = __ldcg(src + t * cap);
                k = v > k ? v : k;
            }
            long_max[i] = (k == 0ull) ? 0.0f : __uint_as_float((uint32_t)(k >> 32));
            long_arg[i] = (k == 0ull) ? 0 : (int64_t)(0xFFFFFFFFu - (uint32_t)(k & 0xFFFFFFFFull));
        } else {
            const int64_t j = i - n_long;
            for (int s = 0; s < world; ++s) {
                const unsigned long long v = __ldcg(s_blk[s] + parts * cap + j);
                k = v > k ? v : k;
            }
            short_max[j] = (k == 0ull) ? 0.0f : __uint_as_float((uint32_t)(k >> 32));
            short_arg[j] = (k == 0ull) ? 0 : (int64_t)(0xFFFFFFFFu - (uint32_t)(k & 0xFFFFFFFFull));
        }
    }
}

// ---- loss --------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(kThreads)
k_loss_fwd_bwd(const float* __restrict__ pred, const float* __restrict__ target, int64_t n, float* __restrict__ iou,
               const float* __restrict__ grad_iou, float* __restrict__ grad_pred, float* __restrict__ grad_target,
               bool vec_ok) {
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= n) return;
    const RawBox x = load_box<D>(pred, i, vec_ok), y = load_box<D>(target, i, vec_ok);
    if (grad_pred == nullptr && grad_target == nullptr) {
        if (iou) iou[i] = sph2pob_iou_pair(x, y, D, KIND_SPH2POB_STANDARD, MODE_IOU, EDGE_ARC);
        return;
    }
    float g1[5], g2[5];
    const float up = grad_iou ? __ldg(grad_iou + i) : 1.0f;     // no upstream gradient given: d(iou)/d(box) itself
    const float v = sph2pob_iou_pair_grad(x, y, D, KIND_SPH2POB_STANDARD, EDGE_ARC, up, g1, g2);
    if (iou) iou[i] = v;
    if (grad_pred) store_grad<D>(grad_pred, i, g1, vec_ok);
    if (grad_target) store_grad<D>(grad_target, i, g2, vec_ok);
}

// Reduced loss: sum_i w_i (1 - iou_i) as per-block partial sums (deterministic: fixed tree inside a block, the
// caller adds the few hundred partials) and the gradients already scaled by -w_i * scale, in ONE launch.
// Replaces the elementwise loss, weight multiply, reduction and the autograd tape of
// sph2pob_iou_loss.py:104-140 + mmdet/models/losses/utils.py (weight_reduce_loss) for mode 'iou'.
template <int D>
__global__ void __launch_bounds__(kThreads)
k_loss_reduce(const float* __restrict__ pred, const float* __restrict__ target, const float* __restrict__ weight, int64_t n,
              float scale, float* __restrict__ partial, float* __restrict__ grad_pred, float* __restrict__ grad_target,
              bool vec_ok, float* __restrict__ total, unsigned* __restrict__ ticket) {
    __shared__ float s_sum[kThreads / 32];
    __shared__ bool s_last;
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    float li = 0.0f;
    if (i < n) {
        const RawBox x = load_box<D>(pred, i, vec_ok), y = load_box<D>(target, i, vec_ok);
        const float w = weight ? __ldg(weight + i) : 1.0f;
        float v;
        if (grad_pred || grad_target) {
            float g1[5], g2[5];
            v = sph2pob_iou_pair_grad(x, y, D, KIND_SPH2POB_STANDARD, EDGE_ARC, -w * scale, g1, g2);
            if (grad_pred) store_grad<D>(grad_pred, i, g1, vec_ok);
            if (grad_target) store_grad<D>(grad_target, i, g2, vec_ok);
        } else {
            v = sph2pob_iou_pair(x, y, D, KIND_SPH2POB_STANDARD, MODE_IOU, EDGE_ARC);
        }
        li = w * (1.0f - v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) li += __shfl_xor_sync(0xFFFFFFFFu, li, o);
    if ((threadIdx.x & 31) == 0) s_sum[threadIdx.x >> 5] = li;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.0f;
#pragma unroll
        for (int k = 0; k < kThreads / 32; ++k) t += s_sum[k];
        partial[blockIdx.x] = t;
        s_last = false;
        if (total) {
            // the block that draws the last ticket adds the partials up -- always in index order, so the sum does not
            // depend on which block that is -- and hands the ticket counter back at zero for the next call
            __threadfence();
            s_last = atomicAdd(ticket, 1u) == gridDim.x - 1u;
        }
    }
    if (!total) return;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    float acc = 0.0f;
    for (unsigned k = threadIdx.x; k < gridDim.x; k += kThreads) acc += __ldcg(partial + k);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_sum[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.0f;
#pragma unroll
        for (int k = 0; k < kThreads / 32; ++k) t += s_sum[k];
        *total = t * scale;
        *ticket = 0u;
    }
}

__device__ __forceinline__ void store_obb(float* __restrict__ o, int64_t i, float x, float y, float w, float h, float a) {
    float* q = o + i * 5;
    q[0] = x; q[1] = y; q[2] = w; q[3] = h; q[4] = a;
}

template <int D>
__global__ void __launch_bounds__(kThreads)
k_obb_fwd(int kind, const float* __restrict__ b1, const float* __restrict__ b2, int64_t n, int edge,
          float* __restrict__ obb1, float* __restrict__ obb2, bool vec_ok) {
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= n) return;
    const RawBox x = load_box<D>(b1, i, vec_ok), y = load_box<D>(b2, i, vec_ok);
    const bool m = jitter1_mask(x, y, D);
    const JitBox g = jitter1_role1(x, m, D), p = jitter1_role2(y, m, D);
    XformAux aux;
    ObbPair o = (kind == KIND_SPH2POB_STANDARD) ? sph2pob_standard(g, p, D, edge, &aux)
                                                : sph2pob_efficient(g, p, D, edge, &aux);
    jitter2(o);
    store_obb(obb1, i, o.x1, o.y1, o.w1, o.h1, o.a1);
    store_obb(obb2, i, o.x2, o.y2, o.w2, o.h2, o.a2);
}

template <int D>
__global__ void __launch_bounds__(kThreads)
k_obb_bwd(int kind, const float* __restrict__ b1, const float* __restrict__ b2, int64_t n, int edge,
          const float* __restrict__ grad_obb1, const float* __restrict__ grad_obb2, float* __restrict__ grad_b1,
          float* __restrict__ grad_b2, bool vec_ok) {
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= n) return;
    const RawBox x = load_box<D>(b1, i, vec_ok), y = load_box<D>(b2, i, vec_ok);
    const bool m = jitter1_mask(x, y, D);
    const JitBox g = jitter1_role1(x, m, D), p = jitter1_role2(y, m, D);
    XformAux aux;
    ObbPair o = (kind == KIND_SPH2POB_STANDARD) ? sph2pob_standard(g, p, D, edge, &aux)
                                                : sph2pob_efficient(g, p, D, edge, &aux);
    const uint32_t pass2 = jitter2(o);
    float go1[5], go2[5], gb1[5], gb2[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        go1[k] = grad_obb1 ? __ldg(grad_obb1 + i * 5 + k) : 0.0f;
        go2[k] = grad_obb2 ? __ldg(grad_obb2 + i * 5 + k) : 0.0f;
    }
    jitter2_grad(pass2, go1, go2);
    xform_grad(kind, g, p, D, edge, aux, go1, go2, gb1, gb2);
    if (grad_b1) store_grad<D>(grad_b1, i, gb1, vec_ok);
    if (grad_b2) store_grad<D>(grad_b2, i, gb2, vec_ok);
}

__global__ void __launch_bounds__(kThreads)
k_riou_fwd_bwd(const float* __restrict__ obb1, const float* __restrict__ obb2, int64_t n, float* __restrict__ iou,
               const float* __restrict__ grad_iou, float* __restrict__ grad_obb1, float* __restrict__ grad_obb2) {
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= n) return;
    ObbPair o;
    const float* a = obb1 + i * 5;
    const float* b = obb2 + i * 5;
    o.x1 = __ldg(a); o.y1 = __ldg(a + 1); o.w1 = __ldg(a + 2); o.h1 = __ldg(a + 3); o.a1 = __ldg(a + 4);
    o.x2 = __ldg(b); o.y2 = __ldg(b + 1); o.w2 = __ldg(b + 2); o.h2 = __ldg(b + 3); o.a2 = __ldg(b + 4);
    if (grad_obb1 == nullptr && grad_obb2 == nullptr) {
        if (iou) iou[i] = obb_disjoint(o) ? 0.0f : riou_value(o, MODE_IOU);
        return;
    }
    float g1[5], g2[5];
    float v = 0.0f;
    if (obb_disjoint(o)) {
#pragma unroll
        for (int k = 0; k < 5; ++k) { g1[k] = 0.0f; g2[k] = 0.0f; }
    } else {
        v = riou_grad(o, MODE_IOU, grad_iou ? __ldg(grad_iou + i) : 1.0f, g1, g2);
    }
    if (iou) iou[i] = v;
    if (grad_obb1) store_grad<5>(grad_obb1, i, g1, false);
    if (grad_obb2) store_grad<5>(grad_obb2, i, g2, false);
}

// ---- the other Sph2Pob losses (GD / KF / L1 on the transformed OBBs), forward + backward in one launch ----------
// One thread per row: jitter_1 -> transform(kind) -> jitter_2 -> row loss on dual numbers (sphk_obbloss.cuh) ->
// jitter_2 / transform backward (sphk_grad.cuh).  `up` = upstream weight / gradient per row ([n]), per element
// ([n, L]) or NULL (= 1), always multiplied by `scale`.  Rows whose upstream is entirely zero are skipped when no
// elementwise loss is requested: their gradient is exactly zero and they add nothing to the reduced loss.
//   loss    [n, L]  unweighted elementwise loss (NULL to skip)
//   partial [grid]  per-block sums of sum_j up[i, j] * loss[i, j] (without `scale`; NULL to skip)
//   grad_b1 / grad_b2 [n, D]  scale * sum_j up[i, j] * d(loss[i, j]) / d(box)
template <int D, typename T>
__global__ void __launch_bounds__(kThreads, 2)
k_obb_loss(LossParams lp, int xkind, const float* __restrict__ b1, const float* __restrict__ b2, int64_t n,
           const float* __restrict__ up, int up_cols, float scale, float* __restrict__ loss, float* __restrict__ partial,
           float* __restrict__ grad_b1, float* __restrict__ grad_b2, bool vec_ok, float* __restrict__ total,
           unsigned* __restrict__ ticket) {
    __shared__ float s_sum[kThreads / 32];
    __shared__ bool s_last;
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    const int L = loss_columns(lp.kind);
    float li = 0.0f;
    if (i < n) {
        float u[5] = {1.0f, 1.0f, 1.0f, 1.0f, 1.0f};
        if (up) {
#pragma unroll
            for (int k = 0; k < 5; ++k) u[k] = (k < L) ? __ldg(up + i * up_cols + (up_cols > 1 ? k : 0)) : 0.0f;
        }
        bool any = false;
#pragma unroll
        for (int k = 0; k < 5; ++k) any = any || (k < L && u[k] != 0.0f);
        float gb1[5] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f}, gb2[5] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
        if (any || loss) {
            const RawBox x = load_box<D>(b1, i, vec_ok), y = load_box<D>(b2, i, vec_ok);
            const bool m = jitter1_mask(x, y, D);
            const JitBox g = jitter1_role1(x, m, D), p = jitter1_role2(y, m, D);
            XformAux aux;
            ObbPair o = (xkind == KIND_SPH2POB_STANDARD) ? sph2pob_standard(g, p, D, EDGE_ARC, &aux)
                                                         : sph2pob_efficient(g, p, D, EDGE_ARC, &aux);
            const uint32_t pass2 = jitter2(o);
            float go1[5] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f}, go2[5] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
            float us[5];
#pragma unroll
            for (int k = 0; k < 5; ++k) us[k] = u[k] * scale;
            if (lp.kind == LOSS_L1) {
                float ell[5];
                obb_l1_loss_row(o, lp, us, ell, go1, go2);
#pragma unroll
                for (int k = 0; k < 5; ++k) {
                    if (loss) loss[i * 5 + k] = ell[k];
                    li += u[k] * ell[k];
                }
            } else {
                const float v = obb_scalar_loss_row<T>(o, lp, us[0], go1, go2);
                if (loss) loss[i] = v;
                li = u[0] * v;
            }
            if ((grad_b1 || grad_b2) && any) {
                jitter2_grad(pass2, go1, go2);
                xform_grad(xkind, g, p, D, EDGE_ARC, aux, go1, go2, gb1, gb2);
            }
            if (!any) li = 0.0f;
        }
        if (grad_b1) store_grad<D>(grad_b1, i, gb1, vec_ok);
        if (grad_b2) store_grad<D>(grad_b2, i, gb2, vec_ok);
    }
    if (partial == nullptr) return;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) li += __shfl_xor_sync(0xFFFFFFFFu, li, o);
    if ((threadIdx.x & 31) == 0) s_sum[threadIdx.x >> 5] = li;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.0f;
#pragma unroll
        for (int k = 0; k < kThreads / 32; ++k) t += s_sum[k];
        partial[blockIdx.x] = t;
        s_last = false;
        if (total) {
            // as k_loss_reduce: the block that draws the last ticket adds the partials up in index order (the sum does not
            // depend on which block that is) and hands the ticket counter back at zero for the next call
            __threadfence();
            s_last = atomicAdd(ticket, 1u) == gridDim.x - 1u;
        }
    }
    if (!total) return;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    float acc = 0.0f;
    for (unsigned k = threadIdx.x; k < gridDim.x; k += kThreads) acc += __ldcg(partial + k);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_sum[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.0f;
#pragma unroll
        for (int k = 0; k < kThreads / 32; ++k) t += s_sum[k];
        *total = t * scale;
        *ticket = 0u;
    }
}

// ---- box coders + the decode -> loss step of the head ---------------------------------------------
template <int D>
__global__ void __launch_bounds__(kThreads)
k_coder_decode(const float* __restrict__ rois, const float* __restrict__ deltas, const float* __restrict__ grad_out, int64_t n,
               CoderParams cp, float* __restrict__ out, bool vec_ok) {
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= n) return;
    const RawBox roi = load_box<D>(rois, i, vec_ok), dl = load_box<D>(deltas, i, vec_ok);
    const float d[5] = {dl.t, dl.p, dl.a, dl.b, dl.g};
    uint32_t pass;
    float jac[5], v[5];
    const RawBox b = coder_decode(roi, d, D, cp, &pass, jac);
    if (grad_out) {       // backward: out = d(total)/d(deltas)
        const RawBox go = load_box<D>(grad_out, i, vec_ok);
        const float g[5] = {go.t, go.p, go.a, go.b, go.g};
        coder_decode_grad(g, pass, jac, D, v);
    } else {
        v[0] = b.t; v[1] = b.p; v[2] = b.a; v[3] = b.b; v[4] = b.g;
    }
    store_grad<D>(out, i, v, vec_ok);
}

// ---- box format conversions (sphdet/bbox/box_formator.py): one thread per row ---------------------------------------
__global__ void __launch_bounds__(kThreads) k_box_format(const float* __restrict__ in, int64_t n, int fmt, int d_in, int d_out,
                                                          float img_h, float img_w, float* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= n) return;
    float a[5], b[5];
    for (int k = 0; k < d_in; ++k) a[k] = in[i * d_in + k];
    box_format_row(fmt, a, d_in, b, d_out, img_h, img_w);
    for (int k = 0; k < d_out; ++k) out[i * d_out + k] = b[k];
}

// four-column rows in and out (every format between planar and spherical BFoV boxes): 16-byte loads and stores, four rows
// per thread with the four loads issued before the first use (HBM-bound: what keeps enough bytes in flight per SM)
__global__ void __launch_bounds__(kThreads) k_box_format4(const float4* __restrict__ in, int64_t n, int fmt, float img_h, float img_w,
                                                           float4* __restrict__ out) {
    const int64_t base = (int64_t)blockIdx.x * (kThreads * 4) + threadIdx.x;
    float4 v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int64_t i = base + (int64_t)k * kThreads;
        v[k] = (i < n) ? __ldg(in + i) : make_float4(0.0f, 0.0f, 1.0f, 1.0f);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int64_t i = base + (int64_t)k * kThreads;
        const float a[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
        float b[4];
        box_format_row(fmt, a, 4, b, 4, img_h, img_w);
        if (i < n) out[i] = make_float4(b[0], b[1], b[2], b[3]);
    }
}

template <int D>
__global__ void __launch_bounds__(kThreads)
k_coder_encode(const float* __restrict__ proposals, const float* __restrict__ gt, int64_t n, CoderParams cp,
               float* __restrict__ out, bool vec_ok) {
    const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= n) return;
    float v[5];
    coder_encode(load_box<D>(proposals, i, vec_ok), load_box<D>(gt, i, vec_ok), D, cp, v);
    store_grad<D>(out, i, v, vec_ok);
}

// ---- training targets of the anchor head (mmdet/models/dense_heads/anchor_head.py:254-285, PseudoSampler) -------
// From the assignment (gt_inds: 0 negative, -1 ignored, k + 1 = GT k of the image) straight to what loss_single
// consumes: labels, label_weights, bbox_targets (the GT box itself with reg_decoded_bbox, else bbox_coder.encode(anchor,
// GT)), bbox_weights and the per-image positive / negative counts.  One thread per (image, anchor); grid.y = image.
template <int D>
__global__ void __launch_bounds__(kThreads)
k_anchor_targets(const int64_t* __restrict__ gt_inds, int64_t N, const float* __restrict__ anchors, const float* __restrict__ gts,
                 const int64_t* __restrict__ gt_labels, const int32_t* __restrict__ offsets, int64_t num_classes, float pos_w,
                 bool decoded, CoderParams cp, int64_t* __restrict__ labels, float* __restrict__ label_weights,
                 float* __restrict__ bbox_targets, float* __restrict__ bbox_weights, int32_t* __restrict__ counts) {
    __shared__ int s_cnt[2];
    if (threadIdx.x < 2) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const int b = blockIdx.y;
    const int64_t a = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    bool pos = false, neg = false;
    if (a < N) {
        const int64_t i = (int64_t)b * N + a;
        const int64_t gi = gt_inds[i];
        pos = gi > 0; neg = gi == 0;
        float t[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
        int64_t lab = num_classes;
        if (pos) {
            const int64_t g = (int64_t)offsets[b] + gi - 1;
            const RawBox gt = load_box<D>(gts, g, false);
            if (decoded) { t[0] = gt.t; t[1] = gt.p; t[2] = gt.a; t[3] = gt.b; t[4] = gt.g; }
            else coder_encode(load_box<D>(anchors, a, false), gt, D, cp, t);
            lab = gt_labels ? gt_labels[g] : 0;            // only an RPN passes gt_labels = None: foreground = class 0
        }
        labels[i] = lab;
        label_weights[i] = pos ? (pos_w <= 0.0f ? 1.0f : pos_w) : (neg ? 1.0f : 0.0f);
        const float w = pos ? 1.0f : 0.0f;
#pragma unroll
        for (int k = 0; k < D; ++k) { bbox_targets[i * D + k] = t[k]; bbox_weights[i * D + k] = w; }
    }
    const unsigned mp = __ballot_sync(0xFFFFFFFFu, pos), mn = __ballot_sync(0xFFFFFFFFu, neg);
    if ((threadIdx.x & 31) == 0) {
        if (mp) atomicAdd(&s_cnt[0], __popc(mp));
        if (mn) atomicAdd(&s_cnt[1], __popc(mn));
    }
    __syncthreads();
    if (threadIdx.x < 2 && s_cnt[threadIdx.x]) atomicAdd(&counts[b * 2 + threadIdx.x], s_cnt[threadIdx.x]);
}

// Persistent warps scan the rows 32 at a time: read the weight(s), zero the row's gradient, ballot-compact the rows
// with a non-zero weight into a per-warp ring; 32 queued rows at a time go through decode + loss + backward.
// The loss is accumulated per lane in a fixed order and reduced per CTA: deterministic for a given n.
constexpr int kLossRing = 64;

template <int D>
__global__ void __launch_bounds__(kThreads)
k_decode_loss(const float* __restrict__ anchors, const float* __restrict__ deltas, const float* __restrict__ target,
              const float* __restrict__ weight, int wcols, int64_t n, CoderParams cp, float scale,
              float* __restrict__ partial, float* __restrict__ grad) {
    __shared__ int s_ring[kThreads / 32][kLossRing];
    __shared__ float s_w[kThreads / 32][kLossRing];
    __shared__ float s_sum[kThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nw = (int)gridDim.x * (kThreads / 32), wg = (int)blockIdx.x * (kThreads / 32) + warp;
    const int chunks = (int)((n + 31) >> 5);
    const unsigned lt = (1u << lane) - 1u;
    int head = 0, tail = 0;
    float acc = 0.0f;
    int c = wg;
#pragma unroll 1
    for (;;) {
#pragma unroll 1
        while (c < chunks && tail - head < 32) {
            const int64_t i = (int64_t)c * 32 + lane;
            float w = 0.0f;
            if (i < n) {
                if (wcols == 0) {
                    w = 1.0f;
                } else {
                    const float* q = weight + i * wcols;
                    float sum = 0.0f;
                    for (int k = 0; k < wcols; ++k) sum += __ldg(q + k);
                    w = (wcols == 1) ? sum : sum / (float)wcols;
                }
            }
            if (grad) {       // the chunk's 32 * D gradient floats, coalesced
                const int64_t f0 = (int64_t)c * (32 * D), fend = n * D;
#pragma unroll
                for (int k = 0; k < D; ++k)
                    if (f0 + k * 32 + lane < fend) grad[f0 + k * 32 + lane] = 0.0f;
            }
            const bool live = (i < n) && (w != 0.0f);
            const unsigned m = __ballot_sync(0xFFFFFFFFu, live);
            if (live) {
                const int slot = (tail + __popc(m & lt)) & (kLossRing - 1);
                s_ring[warp][slot] = (int)i;
                s_w[warp][slot] = w;
            }
            tail += __popc(m);
            SPHK_CHECK(tail - head <= kLossRing);
            c += nw;
        }
        if (tail == head) break;                 // chunks used up and nothing queued
        const int cnt = min(tail - head, 32);
        __syncwarp();
        const int row = s_ring[warp][(head + lane) & (kLossRing - 1)];
        const float w = s_w[warp][(head + lane) & (kLossRing - 1)];
        __syncwarp();
        if (lane < cnt) {
            SPHK_CHECK(row >= 0 && row < n);
            const RawBox roi = load_box<D>(anchors, row, false), dl = load_box<D>(deltas, row, false);
            const RawBox tg = load_box<D>(target, row, false);
            const float d[5] = {dl.t, dl.p, dl.a, dl.b, dl.g};
            uint32_t pass;
            float jac[5], g1[5], g2[5], gd[5];
            const RawBox pred = coder_decode(roi, d, D, cp, &pass, jac);
            float iou;
            if (grad) {
                iou = sph2pob_iou_pair_grad(pred, tg, D, KIND_SPH2POB_STANDARD, EDGE_ARC, -w * scale, g1, g2);
                coder_decode_grad(g1, pass, jac, D, gd);
                store_grad<D>(grad, row, gd, false);
            } else {
                iou = sph2pob_iou_pair(pred, tg, D, KIND_SPH2POB_STANDARD, MODE_IOU, EDGE_ARC);
            }
            acc += w * (1.0f - iou);
        }
        head += cnt;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, o);
    if (lane == 0) s_sum[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.0f;
#pragma unroll
        for (int k = 0; k < kThreads / 32; ++k) t += s_sum[k];
        partial[blockIdx.x] = t;
    }
}

// ---- NMS ---------------------------------------------------------------------------------------
// One CTA per segment (the boxes of one (image, class) group, score-descending through `order`).
// Pivots are taken 32 at a time.  For a block of 32 pivots every warp builds suppression words
// word(pivot i, column word jw) = ballot_j [ IoU(box_i as bboxes1, box_j as bboxes2) > thr ] for the
// columns j > i that are still alive; warp 0 then replays the greedy order inside the block with
// plain register/shared-memory bit operations (no IoU, no block barrier per pivot).  Pivots and
// columns already removed by earlier blocks are never evaluated.
__device__ __forceinline__ uint32_t desc_score_bits(float sc) {
    // ascending order of the result == descending order of the float (NaN ends up first, like torch.sort(descending=True))
    const uint32_t b = __float_as_uint(sc);
    const uint32_t asc = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
    return ~asc;
}

__device__ __forceinline__ void bitonic_sort_u64(unsigned long long* s, int n) {     // n: a power of two
    for (int k = 2; k <= n; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                const int l = i ^ j;
                if (l > i) {
                    const unsigned long long a = s[i], b = s[l];
                    if ((a > b) == ((i & k) == 0)) { s[i] = b; s[l] = a; }
                }
            }
            __syncthreads();
        }
    }
}

// The pivot test of sph_nms_op (sph_nms.py:70-73): the candidate survives iff IoU(pivot, candidate) <= thr (so a NaN IoU,
// which only the unclamped planar `naive_iou` can produce, suppresses).  kind: Sph2Pob-efficient (SphNMS's default) or naive.
__device__ __noinline__ float naive_pair_outofline(const RawBox& x, const RawBox& y, int D) { return naive_iou_pair(x, y, D, MODE_IOU); }
__device__ __noinline__ float unbiased_pair_outofline(const RawBox& x, const RawBox& y, int D) { return unbiased_iou_pair(x, y, D); }
__device__ __forceinline__ bool nms_suppresses(const float* __restrict__ boxes, int bi, int bj, const RawBox& x, const RawBox& y,
                                               int D, int kind, float thr) {
    // SphNMS keeps what passes `ious <= thr` (sph_nms.py:70): a NaN IoU suppresses.  mmcv's nms -- PlanarNMS, planar_nms.py:16 --
    // suppresses on `inter > thr * union`: a NaN IoU (two zero-area boxes) keeps the box (SPHK_NMS_RULE_GT).
    if ((kind & 0xFF) == KIND_NAIVE) {
        const float v = naive_pair_outofline(x, y, D);
        return (kind & SPHK_NMS_RULE_GT) ? v > thr : !(v <= thr);
    }
    if (kind == KIND_UNBIASED) return !(unbiased_pair_outofline(x, y, D) <= thr);
    return pair_iou_any(boxes, bi, boxes, bj, x, y, D, KIND_SPH2POB_EFFICIENT, MODE_IOU, EDGE_ARC) > thr;
}

// `scores` != NULL: the segment arrives unordered (k_img_scatter) and is first put into (score descending, index
// ascending) order by its own CTA, in shared memory (`sort_cap` keys; the box indices then stay there).
template <int D>
__global__ void __launch_bounds__(1024)
k_nms(const float* __restrict__ boxes, int32_t* order, const int32_t* __restrict__ seg_offsets,
      const int32_t* __restrict__ seg_len, const float* __restrict__ scores, int sort_cap, int32_t* __restrict__ bad_image,
      int segs_per_image, float thr, uint8_t* __restrict__ keep, int max_words, bool vec_ok, int kind) {
    extern __shared__ uint32_t s_mem[];
    const int seg = blockIdx.x;
    const int start = seg_offsets[seg];
    const int k = seg_len ? seg_len[seg] : seg_offsets[seg + 1] - start;     // seg_len: segments need not be adjacent
    if (k <= 0) return;
    const int W = (k + 31) >> 5;
    if (W > max_words) {   // caller under-sized max_seg_len: refuse rather than overrun shared memory
        for (int q = threadIdx.x; q < k; q += blockDim.x) keep[start + q] = 0xFF;
        return;
    }
    uint32_t* removed = s_mem;             // [W]
    uint32_t* mask = s_mem + max_words;    // [32][W]
    unsigned long long* s_keys = reinterpret_cast<unsigned long long*>(s_mem + ((33 * max_words + 1) & ~1));   // [sort_cap]
    bool idx_in_smem = false;
    if (scores) {
        int kp = 32;
        while (kp < k) kp <<= 1;
        if (kp == 32) {
            // up to 32 candidates: one warp sorts them in registers (shuffle network), no block barrier per stage
            if (threadIdx.x < 32) {
                unsigned long long key = ~0ull;
                if ((int)threadIdx.x < k) {
                    const int32_t idx = order[start + threadIdx.x];
                    key = ((unsigned long long)desc_score_bits(scores[idx]) << 32) | (unsigned long long)(uint32_t)idx;
                }
#pragma unroll
                for (int kk = 2; kk <= 32; kk <<= 1) {
#pragma unroll
                    for (int j = kk >> 1; j > 0; j >>= 1) {
                        const unsigned long long other = __shfl_xor_sync(0xFFFFFFFFu, key, j);
                        const bool lower = (threadIdx.x & j) == 0, up = (threadIdx.x & kk) == 0;
                        key = ((key > other) == (lower == up)) ? other : key;
                    }
                }
                s_keys[threadIdx.x] = key;
                if ((int)threadIdx.x < k) order[start + threadIdx.x] = (int32_t)(key & 0xFFFFFFFFull);
            }
            __syncthreads();
            idx_in_smem = true;
        } else if (kp <= sort_cap) {
            for (int q = threadIdx.x; q < kp; q += blockDim.x) {
                unsigned long long key = ~0ull;
                if (q < k) {
                    const int32_t idx = order[start + q];
                    key = ((unsigned long long)desc_score_bits(scores[idx]) << 32) | (unsigned long long)(uint32_t)idx;
                }
                s_keys[q] = key;
            }
            __syncthreads();
            bitonic_sort_u64(s_keys, kp);
            for (int q = threadIdx.x; q < k; q += blockDim.x) order[start + q] = (int32_t)(s_keys[q] & 0xFFFFFFFFull);
            idx_in_smem = true;
        } else {
            // longer than the sorting buffer (> 4096 candidates of one class in one image): not handled here; the image is
            // reported as unusable (out_count = -1) and the caller takes the general path (sphk_nms_batched)
            if (threadIdx.x == 0) bad_image[seg / segs_per_image] = 1;
            return;
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (int w = threadIdx.x; w < W; w += blockDim.x) removed[w] = 0u;
    __syncthreads();
    for (int i0 = 0; i0 < k; i0 += 32) {
        const int nb = min(32, k - i0);
        const int w0 = i0 >> 5;
        const int nw = W - w0;
        const uint32_t dead_pivots = removed[w0];
        // Work units of this round, one warp each.  The diagonal word (the block's own 32 columns) only has the pairs
        // pi < pj: they are enumerated as a flattened triangle, 32 pairs per unit, so that short segments -- where
        // the diagonal word is all there is -- still run full warps.  Every other word is one (pivot, word) unit
        // with lane = column.
        const int tri = nb * (nb - 1) / 2, tri_units = (tri + 31) >> 5, units = tri_units + nb * (nw - 1);
        for (int q = threadIdx.x; q < nb; q += blockDim.x) mask[q * W + w0] = 0u;
        __syncthreads();
        for (int unit = warp; unit < units; unit += nwarps) {
            if (unit < tri_units) {
                const int p = (unit << 5) + lane;
                bool sup = false;
                int pi = 0, pj = 0;
                if (p < tri) {
                    // row pi holds the pairs (pi, pi+1 .. nb-1); rows before it hold pi * (2 nb - pi - 1) / 2 pairs
                    pi = (int)(((float)(2 * nb - 1) - sqrtf((float)((2 * nb - 1) * (2 * nb - 1) - 8 * p))) * 0.5f);
                    while (pi > 0 && pi * (2 * nb - pi - 1) / 2 > p) --pi;
                    while ((pi + 1) * (2 * nb - pi - 2) / 2 <= p) ++pi;
                    pj = pi + 1 + (p - pi * (2 * nb - pi - 1) / 2);
                    if (!((dead_pivots >> pi) & 1u) && !((dead_pivots >> pj) & 1u)) {
                        const int bi = idx_in_smem ? (int)(uint32_t)s_keys[i0 + pi] : order[start + i0 + pi];
                        const int bj = idx_in_smem ? (int)(uint32_t)s_keys[i0 + pj] : order[start + i0 + pj];
                        const RawBox x = load_box<D>(boxes, bi, vec_ok), y = load_box<D>(boxes, bj, vec_ok);
                        sup = nms_suppresses(boxes, bi, bj, x, y, D, kind, thr);
                    }
                }
                if (sup) atomicOr(&mask[pi * W + w0], 1u << pj);
            } else {
                const int u = unit - tri_units;
                const int pi = u / (nw - 1), jw = w0 + 1 + u % (nw - 1);
                const int i = i0 + pi, j = (jw << 5) + lane;
                bool sup = false;
                if (!((dead_pivots >> pi) & 1u) && j < k && !((removed[jw] >> lane) & 1u)) {
                    const int bi = idx_in_smem ? (int)(uint32_t)s_keys[i] : order[start + i];
                    const int bj = idx_in_smem ? (int)(uint32_t)s_keys[j] : order[start + j];
                    const RawBox x = load_box<D>(boxes, bi, vec_ok), y = load_box<D>(boxes, bj, vec_ok);
                    sup = nms_suppresses(boxes, bi, bj, x, y, D, kind, thr);
                }
                const uint32_t word = __ballot_sync(0xFFFFFFFFu, sup);
                if (lane == 0) mask[pi * W + jw] = word;
            }
        }
        __syncthreads();
        if (warp == 0) {
            for (int pi = 0; pi < nb; ++pi) {
                const bool alive = !((removed[w0] >> pi) & 1u);
                __syncwarp();
                if (alive)
                    for (int w = w0 + lane; w < W; w += 32) removed[w] |= mask[pi * W + w];
                __syncwarp();
            }
        }
        __syncthreads();
    }
    for (int q = threadIdx.x; q < k; q += blockDim.x) keep[start + q] = ((removed[q >> 5] >> (q & 31)) & 1u) ? 0 : 1;
}

// ---- NMS of a batch laid out as equal blocks of K candidates per image: everything on the device ----------
// k_img_scatter one CTA per image: counts the candidates of every class (shared-memory atomics), scans the counts and
//               scatters the candidates into their (image, class) segment of `order` -- a counting sort by label;
//               the order inside a segment is left to the segment's own CTA
// k_nms         one CTA per (image, class) segment: sorts its candidates by descending score, then suppresses
// k_img_collect one CTA per image: the survivors of all its classes, compacted and sorted by descending score
//               (bitonic in shared memory), the first max_out of them
__global__ void __launch_bounds__(1024)
k_img_scatter(const int64_t* __restrict__ labels, const uint8_t* __restrict__ valid, int K, int C, int32_t* __restrict__ order,
              int32_t* __restrict__ seg_start, int32_t* __restrict__ seg_len, int32_t* __restrict__ bad_label) {
    extern __shared__ int s_cnt[];           // [C] counts, then cursors
    const int b = blockIdx.x;
    const int64_t base = (int64_t)b * K;
    for (int c = threadIdx.x; c < C; c += blockDim.x) s_cnt[c] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < K; i += blockDim.x) {
        const int64_t lab = labels[base + i];
        const bool wanted = !valid || valid[base + i];
        if (wanted && lab >= 0 && lab < C) atomicAdd(&s_cnt[lab], 1);
        else if (wanted) bad_label[b] = 1;      // reported through out_count = -1: the caller must not trust this image
    }
    __syncthreads();
    if (threadIdx.x < 32) {      // exclusive scan of the class counts by one warp; the counts become cursors
        int run = 0;
        for (int c0 = 0; c0 < C; c0 += 32) {
            const int c = c0 + threadIdx.x;
            const int v = (c < C) ? s_cnt[c] : 0;
            int inc = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xFFFFFFFFu, inc, o);
                if ((int)threadIdx.x >= o) inc += t;
            }
            if (c < C) {
                seg_start[(int64_t)b * C + c] = (int32_t)(base + run + inc - v);
                seg_len[(int64_t)b * C + c] = v;
                s_cnt[c] = run + inc - v;
            }
            run += __shfl_sync(0xFFFFFFFFu, inc, 31);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < K; i += blockDim.x) {
        const int64_t lab = labels[base + i];
        if ((!valid || valid[base + i]) && lab >= 0 && lab < C) order[base + atomicAdd(&s_cnt[lab], 1)] = (int32_t)(base + i);
    }
}

constexpr int kSelBins = 2048;

__global__ void __launch_bounds__(1024)
k_img_collect(const float* __restrict__ scores, const int32_t* __restrict__ order, const uint8_t* __restrict__ keep, int K, int Kp,
              int max_out, const int32_t* __restrict__ bad_label, int32_t* __restrict__ out_idx, int32_t* __restrict__ out_count) {
    extern __shared__ unsigned long long s_keys[];
    __shared__ int s_hist[kSelBins];
    __shared__ int s_n, s_total, s_bin;
    __shared__ unsigned s_kmin, s_kmax;
    const int b = blockIdx.x;
    const int64_t base = (int64_t)b * K;
    const int lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { s_n = 0; s_total = 0; s_bin = kSelBins; s_kmin = 0xFFFFFFFFu; s_kmax = 0u; }
    for (int q = threadIdx.x; q < kSelBins; q += blockDim.x) s_hist[q] = 0;
    __syncthreads();
    // ---- pass A: how many survivors, and the range of their (descending-score) keys
    for (int i0 = 0; i0 < K; i0 += blockDim.x) {
        const int i = i0 + threadIdx.x;
        const bool kept = i < K && keep[base + i] == 1;           // (positions no segment covers keep their initial 0)
        const unsigned key = kept ? desc_score_bits(scores[order[base + i]]) : 0u;
        const unsigned m = __ballot_sync(0xFFFFFFFFu, kept);
        const unsigned lo = __reduce_min_sync(0xFFFFFFFFu, kept ? key : 0xFFFFFFFFu), hi = __reduce_max_sync(0xFFFFFFFFu, key);
        if (lane == 0 && m) { atomicAdd(&s_total, __popc(m)); atomicMin(&s_kmin, lo); atomicMax(&s_kmax, hi); }
    }
    __syncthreads();
    const int total = s_total;
    const unsigned kmin = s_kmin, span = s_kmax - s_kmin;
    // ---- selection (only when it pays): a histogram over the key range finds the bin that holds the max_out-th best
    // score; only the survivors up to that bin are sorted
    const bool select = total > max_out && total > 512 && span > 0u;
    if (select) {
        for (int i0 = 0; i0 < K; i0 += blockDim.x) {
            const int i = i0 + threadIdx.x;
            if (i < K && keep[base + i] == 1) {
                const unsigned key = desc_score_bits(scores[order[base + i]]);
                atomicAdd(&s_hist[(int)(((unsigned long long)(key - kmin) * (kSelBins - 1)) / span)], 1);
            }
        }
        __syncthreads();
        if (threadIdx.x < 32) {          // first bin at which the running count reaches max_out
            constexpr int per = kSelBins / 32;
            int sum = 0;
            for (int q = 0; q < per; ++q) sum += s_hist[lane * per + q];
            int inc = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xFFFFFFFFu, inc, o);
                if (lane >= o) inc += t;
            }
            int run = inc - sum;
            if (run < max_out && inc >= max_out) {
                for (int q = 0; q < per; ++q) {
                    run += s_hist[lane * per + q];
                    if (run >= max_out) { s_bin = lane * per + q; break; }
                }
            }
        }
        __syncthreads();
    }
    const int bin_hi = s_bin;
    // ---- compaction (one shared-memory atomic per warp) of the survivors [up to the boundary bin]
    for (int i0 = 0; i0 < K; i0 += blockDim.x) {
        const int i = i0 + threadIdx.x;
        bool take = i < K && keep[base + i] == 1;
        int32_t idx = 0;
        unsigned key = 0u;
        if (take) {
            idx = order[base + i];
            key = desc_score_bits(scores[idx]);
            if (select) take = (int)(((unsigned long long)(key - kmin) * (kSelBins - 1)) / span) <= bin_hi;
        }
        const unsigned m = __ballot_sync(0xFFFFFFFFu, take);
        int slot = 0;
        if (lane == 0 && m) slot = atomicAdd(&s_n, __popc(m));
        slot = __shfl_sync(0xFFFFFFFFu, slot, 0) + __popc(m & ((1u << lane) - 1u));
        if (take) s_keys[slot] = ((unsigned long long)key << 32) | (unsigned long long)(uint32_t)idx;
    }
    __syncthreads();
    const int cnt = s_n;
    int kp = 32;
    while (kp < cnt) kp <<= 1;
    for (int i = cnt + threadIdx.x; i < kp; i += blockDim.x) s_keys[i] = ~0ull;
    __syncthreads();
    bitonic_sort_u64(s_keys, kp);
    const int n = min(total, max_out);
    for (int i = threadIdx.x; i < max_out; i += blockDim.x)
        out_idx[(int64_t)b * max_out + i] = (i < n) ? (int32_t)(s_keys[i] & 0xFFFFFFFFull) : -1;
    if (threadIdx.x == 0) out_count[b] = bad_label[b] ? -1 : n;
}

__global__ void __launch_bounds__(kThreads) k_probe_fp32(int iters, float* __restrict__ sink) {
    float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.0f, a2 = a0 + 2.0f, a3 = a0 + 3.0f;
    float a4 = a0 + 4.0f, a5 = a0 + 5.0f, a6 = a0 + 6.0f, a7 = a0 + 7.0f;
    const float m = 0.999f + blockIdx.x * 1e-9f, c = 1e-3f;
#pragma unroll 4
    for (int i = 0; i < iters; ++i) {
        a0 = fmaf(a0, m, c); a1 = fmaf(a1, m, c); a2 = fmaf(a2, m, c); a3 = fmaf(a3, m, c);
        a4 = fmaf(a4, m, c); a5 = fmaf(a5, m, c); a6 = fmaf(a6, m, c); a7 = fmaf(a7, m, c);
    }
    sink[(size_t)blockIdx.x * kThreads + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

inline int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline unsigned blocks_for(int64_t n) { return (unsigned)((n + kThreads - 1) / kThreads); }

}  // namespace

// ================================================================================================
extern "C" {

int sphk_abi_version(void) { return SPHK_ABI_VERSION; }
const char* sphk_last_error_string(void) { return g_err; }

int sphk_device_info(int* sm_count, int* cc_major, int* cc_minor) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
    cudaDeviceProp p;
    e = cudaGetDeviceProperties(&p, dev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDeviceProperties");
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    return SPHK_OK;
}

int sphk_iou_aligned(int kind, const float* b1, const float* b2, int64_t P, int D, int mode, int edge, int angle,
                     float* out, void* stream) {
    if (P < 0 || (D != 4 && D != 5)) return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_iou_aligned: P < 0 or D not in {4,5}");
    if (kind < 0 || kind > SPHK_KIND_SPH2POB_LEGACY) return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_iou_aligned: unknown kind");
    if (kind == SPHK_KIND_SPH2POB_LEGACY && D != 4)
        return fail(SPHK_ERR_UNSUPPORTED, "sph2pob_legacy_iou takes BFoV boxes (D = 4) only (torch.chunk(., 4), sph2pob_legacy.py:52)");
    if ((kind == SPHK_KIND_NAIVE || kind == SPHK_KIND_UNBIASED) && mode != SPHK_MODE_IOU)
        return fail(SPHK_ERR_UNSUPPORTED, "naive_iou / unbiased_iou support mode 'iou' only (sph_iou_api.py:104,182)");
    if (mode != SPHK_MODE_IOU && mode != SPHK_MODE_IOF) return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_iou_aligned: unknown mode");
    if (edge < 0 || edge > 2) return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_iou_aligned: unknown edge");
    if (angle != SPHK_ANGLE_EQUATOR && angle != SPHK_ANGLE_PROJECT) return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_iou_aligned: unknown angle");
    if ((kind == SPHK_KIND_SPH || kind == SPHK_KIND_FOV) && (D != 4 || mode != SPHK_MODE_IOU))
        return fail(SPHK_ERR_UNSUPPORTED, "sph_iou / fov_iou take BFoV boxes (D = 4) and mode 'iou' only");
    if (P == 0) return SPHK_OK;
    if (!b1 || !b2 || !out) return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_iou_aligned: null pointer");
    cudaStream_t s = (cudaStream_t)stream;
    const bool v = aligned16(b1) && aligned16(b2);
    if (angle == SPHK_ANGLE_PROJECT && kind <= SPHK_KIND_SPH2POB_STANDARD) {
        if (D == 4) k_iou_project<4><<<blocks_for(P), kThreads, 0, s>>>(b1, P, b2, P, true, kind, mode, edge, out, 0);
        else k_iou_project<5><<<blocks_for(P), kThreads, 0, s>>>(b1, P, b2, P, true, kind, mode, edge, out, 0);
    } else if (kind == SPHK_KIND_NAIVE) {
        if (D == 4) k_iou_aligned<KIND_NAIVE, 4><<<blocks_for(P), kThreads, 0, s>>>(b1, b2, P, mode, edge, out, v, false);
        else k_iou_aligned<KIND_NAIVE, 5><<<blocks_for(P), kThreads, 0, s>>>(b1, b2, P, mode, edge, out, false, false);
    } else if (kind == SPHK_KIND_UNBIASED) {
        if (D == 4) k_iou_aligned<KIND_UNBIASED, 4><<<blocks_for(P), kThreads, 0, s>>>(b1, b2, P, mode, edge, out, v, false);
        else k_iou_aligned<KIND_UNBIASED, 5><<<blocks_for(P), kThreads, 0, s>>>(b1, b2, P, mode, edge, out, false, false);
    } else if (kind == SPHK_KIND_SPH2POB_LEGACY) {
        k_iou_aligned<KIND_SPH2POB_LEGACY, 4><<<blocks_for(P), kThreads, 0, s>>>(b1, b2, P, mode, edge, out, v, false);
    } else if (kind == SPHK_KIND_SPH || kind == SPHK_KIND_FOV) {
        const unsigned g = blocks_for(P);
        if (v && !g_no_approx4 && P >= (int64_t)3 << 20) {      // below ~3 M pairs one pair per thread keeps more loads in flight
            const unsigned g4 = (unsigned)((P + kThreads * 4 - 1) / (kThreads * 4));
            if (kind == SPHK_KIND_SPH) k_approx_aligned4<KIND_SPH><<<g4, kThreads, 0, s>>>((const float4*)b1, (const float4*)b2, P, out);
            else k_approx_aligned4<KIND_FOV><<<g4, kThreads, 0, s>>>((const float4*)b1, (const float4*)b2, P, out);
        } else if (kind == SPHK_KIND_SPH) k_iou_aligned<KIND_SPH, 4><<<g, kThreads, 0, s>>>(b1, b2, P, mode, edge, out, v, g_dense != 0);
        else k_iou_aligned<KIND_FOV, 4><<<g, kThreads, 0, s>>>(b1, b2, P, mode, edge, out, v, g_dense != 0);
    } else {
        // persistent warps: at most one resident wave (MINB CTAs of 8 warps per SM), at least 4 rounds of 32 pairs per
        // warp (so that the survivor batches fill)
        const int64_t chunks = (P + 31) / 32, wpc = kThreads / 32;
        if (chunks > 0x7FFFFFFFll) return fail(SPHK_ERR_UNSUPPORTED, "sphk_iou_aligned: P too large; split the call");
        const int minb = (g_force_minb == 3 || g_force_minb == 4) ? g_force_minb : 4;
        int64_t g64 = (chunks + wpc * 4 - 1) / (wpc * 4);
        const int64_t wave = (g_force_ctas > 0) ? g_force_ctas : (int64_t)minb * sm_count();
        if (g64 > wave) g64 = wave;
        const unsigned g = (unsigned)g64;
        const bool dn = g_dense != 0;
#define SPHK_AL2(DD, VV, MB) k_iou_aligned2<DD, VV, MB><<<g, kThreads, 0, s>>>(b1, b2, P, kind, mode, edge, out, dn)
        if (minb == 3) {
            if (D == 4) { if (v) SPHK_AL2(4, true, 3); else SPHK_AL2(4, false, 3); }
            else { if (v) SPHK_AL2(5, true, 3); else SPHK_AL2(5, false, 3); }
        } else {
            if (D == 4) { if (v) SPHK_AL2(4, true, 4); else SPHK_AL2(4, false, 4); }
            else { if (v) SPHK_AL2(5, true, 4); else SPHK_AL2(5, false, 4); }
        }
#undef SPHK_AL2
    }
    SPHK_LAUNCH_CHECK("k_iou_aligned");
    return SPHK_OK;
}

static inline int64_t keys_bytes(int64_t R, int64_t C) { return ((R + C) * (int64_t)sizeof(unsigned long long) + 15) & ~15ll; }

int64_t sphk_iou_pairwise_workspace_bytes(int64_t R, int64_t C) {
    if (R < 0 || C < 0) return 0;
    // packed max/argmax keys of rows and columns, then the per-box precompute (BoxRec + BoxCull per box)
    return keys_bytes(R, C) + (R + C) * (int64_t)((kBoxRecFloats + kBoxCullFloats) * sizeof(float));
}

// Which calls get the kernel instance with the box-frame prefilter test (prefilter_live<true>)?  The test takes the ROW
// box's frame and the COLUMN's circumradius, so it pays when the rows are the large boxes and the columns the small
// ones: MaxIoUAssigner's orientation, ground truths (few) x anchors (many).  The host cannot see the boxes; it goes by
// the shape of the call: rows at least 4 x fewer than columns.  A misjudged call loses ~1 % (see prefilter_live).
static inline bool box_test_pays(int64_t rows_per_image, int64_t C) { return rows_per_image * 4 <= C; }

// k_box_pre + k_iou_pairwise2 for rows[R] x cols[C] (rows = concatenation of `batch` GT lists when row_offsets is
// given; max_rows = the longest list).  rec / cull: [R + C] records in the workspace.
#ifdef SPHK_TUNING
static void pw2_occupancy_note(const void* fn, const char* what) {      // tools/: resident CTAs per SM of the instance, once
    static int told = 0;
    if (told++ > 2) return;
    int n = -1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, fn, kThreads, 0);
    fprintf(stderr, "[sphk tuning] k_iou_pairwise2<%s>: %d CTAs / SM\n", what, n);
}
#else
static inline void pw2_occupancy_note(const void*, const char*) {}
#endif

static int launch_pairwise2(int kind, const float* rows, int64_t R, const float* cols, int64_t C, int D, int mode, int edge,
                            float4* rec, float4* cull, float* out, int64_t ld, unsigned long long* rkey,
                            unsigned long long* ckey, int32_t row_base, int32_t col_base, const float* row_target,
                            int* col_tie, const int32_t* row_offsets, int64_t col_stride, int batch, int64_t max_rows,
                            cudaStream_t s, bool zero_keys = false, float* tile_rmax = nullptr, bool records_ready = false,
                            int keep_keys = 0, const KeyPush* push = nullptr) {
    const int64_t col_tiles = (C + kThreads - 1) / kThreads;
    // long-row calls (sweeps: the long operand is bboxes1): the row records are computed inside k_iou_pairwise2
    const bool rows_inline = !g_no_rows_inline && batch == 1 && row_offsets == nullptr && R >= 4 * C;
    // single-image calls: the key arrays have one entry per box and are zeroed by k_box_pre (zero_keys)
    // (keep_keys bit 0 / 1: the caller accumulates into the row / column keys of an earlier call -- chunked sweeps)
    unsigned long long* zr = (zero_keys && batch == 1 && !(keep_keys & 1)) ? rkey : nullptr;
    unsigned long long* zc = (zero_keys && batch == 1 && !(keep_keys & 2)) ? ckey : nullptr;
    if ((g_probe & 1) || records_ready) {      // (records_ready: a second pass over the operands of the previous launch)
    } else {
        const int64_t first = (rows_inline && zr == nullptr) ? R : 0;
        if (D == 4) k_box_pre<4><<<blocks_for(R + C - first), kThreads, 0, s>>>(rows, R, cols, C, edge, rec, cull, aligned16(rows), aligned16(cols), zr, zc, rows_inline, first);
        else k_box_pre<5><<<blocks_for(R + C - first), kThreads, 0, s>>>(rows, R, cols, C, edge, rec, cull, false, false, zr, zc, rows_inline, first);
    }
    // row-tile height: 32 when that already yields many CTAs per SM, else 8 so that the heavy
    // (mostly-live) tiles are spread over more warps and the tail of the launch stays short
    const int64_t tiles32 = col_tiles * ((max_rows + 31) / 32) * batch;
    int tr = (tiles32 >= 16ll * sm_count()) ? 32 : 8;
    if (g_force_tr == 8 || g_force_tr == 32) tr = g_force_tr;   // tuning hook (SPHK_TR)
    const int64_t row_tiles = (max_rows + tr - 1) / tr;
    if (row_tiles > 0x7FFFFFFFll || col_tiles > 65535 || batch > 65535)
        return fail(SPHK_ERR_UNSUPPORTED, "pairwise: more than 16.7 M columns, 2^31 row tiles or 65535 images; shard the call");
    // launched with programmatic stream serialization: the prologue (zero-fill of the matrix) overlaps
    // k_box_pre, the kernel itself waits (griddepcontrol.wait) before it reads the records
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)row_tiles, (unsigned)col_tiles, (unsigned)batch);
    cfg.blockDim = dim3(kThreads, 1, 1);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = g_no_pdl ? 0 : 1;
    const float4* crec = rec;
    const float4* ccull = cull;
    const int dn = (g_dense != 0 ? 1 : 0) | (rows_inline ? 4 : 0);
    // which instance: with the box-frame prefilter test in the scan loop (rows = the short, large-box operand: ground
    // truths x anchors), or with the separating-axis stage behind the circle test (everything else: operands of similar size)
    // (the measurement mode -- sphk_set_dense: no early-out anywhere -- runs the instance without either stage)
    const int cull_kind = g_dense ? 0 : (box_test_pays(max_rows, C) ? (g_no_boxcull ? 0 : 1) : (g_no_sat ? 0 : 2));
    const KeyPush kp = push ? *push : KeyPush{nullptr, 0, 0, 1};
    cudaError_t le;
    auto go = [&](auto kernel, const char* what) {
        // every instance wants four resident CTAs per SM (the separating-axis one holds 43 KB of shared memory per CTA):
        // ask for the largest shared-memory carveout instead of leaving the split to the driver's heuristic
        // (once per instance and process: the value never changes)
        static const cudaError_t carve = cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        (void)carve;
        pw2_occupancy_note((const void*)kernel, what);
        le = cudaLaunchKernelEx(&cfg, kernel, rows, R, cols, C, crec, ccull, kind, mode, edge, out, ld, rkey, ckey, (uint32_t)row_base,
                                (uint32_t)col_base, dn, row_target, col_tie, row_offsets, col_stride, tile_rmax, kp);
    };
#define SPHK_PW2C(CK)                                                                   \
    do {                                                                                \
        if (D == 4 && tr == 32) go(k_iou_pairwise2<4, 32, CK>, "4,32," #CK);            \
        else if (D == 4) go(k_iou_pairwise2<4, 8, CK>, "4,8," #CK);                     \
        else if (tr == 32) go(k_iou_pairwise2<5, 32, CK>, "5,32," #CK);                 \
        else go(k_iou_pairwise2<5, 8, CK>, "5,8," #CK);                                 \
    } while (0)
    if (cull_kind == 1) SPHK_PW2C(1);
    else if (cull_kind == 2) SPHK_PW2C(2);
    else SPHK_PW2C(0);
#undef SPHK_PW2C
    if (le != cudaSuccess) return cuda_fail(le, "cudaLaunchKernelEx(k_iou_pairwise2)");
    return SPHK_OK;
}

// k_iou_rows32 with programmatic stream serialization: its CTAs may become resident while the previous kernel of the
// stream drains (the kernel itself waits, griddepcontrol.wait, before it touches global memory).
static int launch_rows32(int kind, const float* rows, int R, const float* cols, int64_t C, int D, int mode, int edge, float* out,
                         int64_t ld, cudaStream_t s) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)((C + kFC - 1) / kFC), 1, 1);
    cfg.blockDim = dim3(kThreads, 1, 1);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = g_no_pdl ? 0 : 1;
    const int fl = (g_dense != 0 ? 1 : 0) | (g_no_boxcull ? 2 : 0);     // (bit 1 exists in the tuning build only)
    const bool rv = D == 4 && aligned16(rows), cv = D == 4 && aligned16(cols);
    cudaError_t le;
    if (D == 4) le = cudaLaunchKernelEx(&cfg, k_iou_rows32<4>, rows, R, cols, C, kind, mode, edge, out, ld, fl, rv, cv);
    else le = cudaLaunchKernelEx(&cfg, k_iou_rows32<5>, rows, R, cols, C, kind, mode, edge, out, ld, fl, rv, cv);
    if (le != cudaSuccess) return cuda_fail(le, "cudaLaunchKernelEx(k_iou_rows32)");
    return SPHK_OK;
}

static int pairwise_impl(int kind, const float* rows, int64_t R, const float* cols, int64_t C, int D, int mode, int edge,
                         int angle, float* out, int64_t ld, float* row_max, int32_t* row_arg, float* col_max,
                         int32_t* col_arg, int32_t row_base, int32_t col_base, void* workspace, void* stream,
                         const float* row_target, int* col_tie, unsigned long long* ext_rkey = nullptr,
                         unsigned long long* ext_ckey = nullptr, int keep_keys = 0, const KeyPush* push = nullptr) {
    if (R < 0 || C < 0 || (D != 4 && D != 5)) return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_iou_pairwise: bad R, C or D");
    if (kind < 0 || kind > SPHK_KIND_SPH2POB_LEGACY) return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_iou_pairwise: unknown kind");
    if (kind == SPHK_KIND_SPH2POB_LEGACY && D != 4)
        return fail(SPHK_ERR_UNSUPPORTED, "sph2pob_legacy_iou takes BFoV boxes (D = 4) only (torch.chunk(., 4), sph2pob_legacy.py:52)");
    if ((kind == SPHK_KIND_NAIVE || kind == SPHK_KIND_UNBIASED) && mode != SPHK_MODE_IOU)
        return fail(SPHK_ERR_UNSUPPORTED, "naive_iou / unbiased_iou support mode 'iou' only (sph_iou_api.py:104,182)");
    if (mode != SPHK_MODE_IOU && mode != SPHK_MODE_IOF) return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_iou_pairwise: unknown mode");
    if (edge < 0 || edge > 2) return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_iou_pairwise: unknown edge");
    if (angle != SPHK_ANGLE_EQUATOR && angle != SPHK_ANGLE_PROJECT) return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_iou_pairwise: unknown angle");
    // "approx": the kinds evaluated one pair per thread by the generic tile kernel (no per-box records, no prefilter)
    const bool approx = (kind == SPHK_KIND_SPH || kind == SPHK_KIND_FOV || kind == SPHK_KIND_NAIVE || kind == SPHK_KIND_UNBIASED ||
                         kind == SPHK_KIND_SPH2POB_LEGACY);
    if ((kind == SPHK_KIND_SPH || kind == SPHK_KIND_FOV) && (D != 4 || mode != SPHK_MODE_IOU))
        return fail(SPHK_ERR_UNSUPPORTED, "sph_iou / fov_iou take BFoV boxes (D = 4) and mode 'iou' only");
    if (out && ld < C) return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_iou_pairwise: ld < C");
    if (R + (int64_t)(uint32_t)row_base > 0xFFFFFFFFll || C + (int64_t)(uint32_t)col_base > 0xFFFFFFFFll)
        return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_iou_pairwise: indices do not fit 32 bits");
    const bool want_row = row_max || row_arg || ext_rkey, want_col = col_max || col_arg || ext_ckey;
    if (angle == SPHK_ANGLE_PROJECT && !approx) {
        if (want_row || want_col) return fail(SPHK_ERR_UNSUPPORTED, "sphk_iou_pairwise: rbb_angle='project' supports the matrix output only");
        if (R == 0 || C == 0 || !out) return SPHK_OK;
        if (!rows || !cols) return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_iou_pairwise: null box pointer");
        if (R * C > 0x7FFFFFFFll * (int64_t)kThreads) return fail(SPHK_ERR_UNSUPPORTED, "sphk_iou_pairwise: grid too large; shard the call");
        if (D == 4) k_iou_project<4><<<blocks_for(R * C), kThreads, 0, (cudaStream_t)stream>>>(rows, R, cols, C, false, kind, mode, edge, out, ld);
        else k_iou_project<5><<<blocks_for(R * C), kThreads, 0, (cudaStream_t)stream>>>(rows, R, cols, C, false, kind, mode, edge, out, ld);
        SPHK_LAUNCH_CHECK("k_iou_project");
        return SPHK_OK;
    }
    if ((want_row || want_col || !approx) && !workspace && (R + C) > 0)
        return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_iou_pairwise: workspace of sphk_iou_pairwise_workspace_bytes(R, C) required");
    if (workspace && !aligned16(workspace)) return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_iou_pairwise: workspace must be 16-byte aligned");
    cudaStream_t s = (cudaStream_t)stream;
    unsigned long long* rkey = ext_rkey ? ext_rkey : (want_row ? (unsigned long long*)workspace : nullptr);
    unsigned long long* ckey = ext_ckey ? ext_ckey : (want_col ? (unsigned long long*)workspace + R : nullptr);
    const bool pre_zeroes_keys = !approx && R > 0 && C > 0;      // k_box_pre of the same call does it
    if (want_row && R > 0 && !pre_zeroes_keys && !(keep_keys & 1)) k_fill_keys<<<blocks_for(R), kThreads, 0, s>>>(rkey, R);
    if (want_col && C > 0 && !pre_zeroes_keys && !(keep_keys & 2)) k_fill_keys<<<blocks_for(C), kThreads, 0, s>>>(ckey, C);
    if (R > 0 && C > 0) {
        if (!rows || !cols) return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_iou_pairwise: null box pointer");
        const int64_t col_tiles = (C + kThreads - 1) / kThreads;
        const bool v = aligned16(cols);
        if (approx) {
            const int64_t row_tiles = (R + kTileRows - 1) / kTileRows;
            if (col_tiles * row_tiles > 0x7FFFFFFFll) return fail(SPHK_ERR_UNSUPPORTED, "sphk_iou_pairwise: grid too large; shard the call");
            const unsigned g = (unsigned)(col_tiles * row_tiles);
            if (kind == SPHK_KIND_SPH2POB_LEGACY)
                k_iou_pairwise<KIND_SPH2POB_LEGACY, 4><<<g, kThreads, 0, s>>>(rows, R, cols, C, mode, edge, out, ld, rkey, ckey, (uint32_t)row_base,
                                                                               (uint32_t)col_base, col_tiles, v, false);
            else if (kind == SPHK_KIND_UNBIASED && D == 4)
                k_iou_pairwise<KIND_UNBIASED, 4><<<g, kThreads, 0, s>>>(rows, R, cols, C, mode, edge, out, ld, rkey, ckey, (uint32_t)row_base,
                                                                         (uint32_t)col_base, col_tiles, v, false);
            else if (kind == SPHK_KIND_UNBIASED)
                k_iou_pairwise<KIND_UNBIASED, 5><<<g, kThreads, 0, s>>>(rows, R, cols, C, mode, edge, out, ld, rkey, ckey, (uint32_t)row_base,
                                                                         (uint32_t)col_base, col_tiles, false, false);
            else if (kind == SPHK_KIND_NAIVE && D == 4)
                k_iou_pairwise<KIND_NAIVE, 4><<<g, kThreads, 0, s>>>(rows, R, cols, C, mode, edge, out, ld, rkey, ckey, (uint32_t)row_base,
                                                                      (uint32_t)col_base, col_tiles, v, false);
            else if (kind == SPHK_KIND_NAIVE)
                k_iou_pairwise<KIND_NAIVE, 5><<<g, kThreads, 0, s>>>(rows, R, cols, C, mode, edge, out, ld, rkey, ckey, (uint32_t)row_base,
                                                                      (uint32_t)col_base, col_tiles, false, false);
            else if (kind == SPHK_KIND_SPH)
                k_iou_pairwise<KIND_SPH, 4><<<g, kThreads, 0, s>>>(rows, R, cols, C, mode, edge, out, ld, rkey, ckey, (uint32_t)row_base,
                                                                    (uint32_t)col_base, col_tiles, v, g_dense != 0);
            else
                k_iou_pairwise<KIND_FOV, 4><<<g, kThreads, 0, s>>>(rows, R, cols, C, mode, edge, out, ld, rkey, ckey, (uint32_t)row_base,
                                                                    (uint32_t)col_base, col_tiles, v, g_dense != 0);
        } else {
            if (out && !want_row && !want_col && !row_target && !col_tie && R <= 32 && !g_no_rows32 &&
                (C + kFC - 1) / kFC <= 0x7FFFFFFFll) {
                // the per-image call of MaxIoUAssigner: one launch, records computed inside the CTAs
                return launch_rows32(kind, rows, (int)R, cols, C, D, mode, edge, out, ld, s);
            }
            float4* rec = (float4*)((char*)workspace + keys_bytes(R, C));       // [R + C][4] rows first
            float4* cull = rec + (R + C) * 4;                                    // [R + C][4]
            const int rc = launch_pairwise2(kind, rows, R, cols, C, D, mode, edge, rec, cull, out, ld, rkey, ckey, row_base,
                                            col_base, row_target, col_tie, nullptr, 0, 1, R, s, true, nullptr, false, keep_keys, push);
            if (rc != SPHK_OK) return rc;
        }
        SPHK_LAUNCH_CHECK("k_iou_pairwise");
    }
    {   // one launch unpacks the row keys and the column keys
        const int64_t nr_ = (row_max || row_arg) ? R : 0, nc_ = (col_max || col_arg) ? C : 0;
        if (nr_ + nc_ > 0)
            k_unpack_keys2<<<blocks_for(nr_ + nc_), kThreads, 0, s>>>(rkey, nr_, row_max, row_arg, (uint32_t)col_base, ckey, nc_, col_max,
                                                                    col_arg, (uint32_t)row_base);
    }
    SPHK_LAUNCH_CHECK("k_unpack_keys");
    return SPHK_OK;
}

int sphk_iou_pairwise(int kind, const float* rows, int64_t R, const float* cols, int64_t C, int D, int mode, int edge,
                      int angle, float* out, int64_t ld, float* row_max, int32_t* row_arg, float* col_max,
                      int32_t* col_arg, int32_t row_base, int32_t col_base, void* workspace, void* stream) {
    return pairwise_impl(kind, rows, R, cols, C, D, mode, edge, angle, out, ld, row_max, row_arg, col_max, col_arg, row_base,
                         col_base, workspace, stream, nullptr, nullptr);
}

int sphk_iou_pairwise_keys(int kind, const float* rows, int64_t R, const float* cols, int64_t C, int D, int mode, int edge,
                           uint64_t* row_keys, uint64_t* col_keys, int32_t row_base, int32_t col_base, int keep, void* workspace,
                           void* stream) {
    if (kind != SPHK_KIND_SPH2POB_EFFICIENT && kind != SPHK_KIND_SPH2POB_STANDARD)
        return fail(SPHK_ERR_UNSUPPORTED, "sphk_iou_pairwise_keys: kind must be a Sph2Pob transform");
    if (!row_keys || !col_keys) return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_iou_pairwise_keys: null key pointer");
    if (keep < 0 || keep > 3) return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_iou_pairwise_keys: keep must be 0..3");
    return pairwise_impl(kind, rows, R, cols, C, D, mode, edge, SPHK_ANGLE_EQUATOR, nullptr, C, nullptr, nullptr, nullptr, nullptr,
                         row_base, col_base, workspace, stream, nullptr, nullptr, (unsigned long long*)row_keys,
                         (unsigned long long*)col_keys, keep);
}

int32_t sphk_key_push_parts(int64_t C) { return C <= 0 ? 1 : (int32_t)((C + kThreads - 1) / kThreads); }

int sphk_iou_pairwise_keys_push(int kind, const float* rows, int64_t R, const float* cols, int64_t C, int D, int mode, int edge,
                                uint64_t* col_keys, int32_t row_base, int32_t col_base, uint64_t* const* peer_bufs, int32_t world,
                                int64_t push_offset, int64_t part_stride, void* workspace, void* stream) {
    if (kind != SPHK_KIND_SPH2POB_EFFICIENT && kind != SPHK_KIND_SPH2POB_STANDARD)
        return fail(SPHK_ERR_UNSUPPORTED, "sphk_iou_pairwise_keys_push: kind must be a Sph2Pob transform");
    if (!col_keys || !peer_bufs) return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_iou_pairwise_keys_push: null pointer");
    if (world < 1 || world > 16 || push_offset < 0 || part_stride < R)
        return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_iou_pairwise_keys_push: bad world (1..16), offset or part_stride < R");
    // (R == 0 or C == 0 launches no tile kernel: nothing is stored, the callers' buffers start zeroed = "no overlap")
    const KeyPush kp = {(unsigned long long* const*)peer_bufs, (long long)push_offset, (long long)part_stride, world};
    return pairwise_impl(kind, rows, R, cols, C, D, mode, edge, SPHK_ANGLE_EQUATOR, nullptr, C, nullptr, nullptr, nullptr, nullptr,
                         row_base, col_base, workspace, stream, nullptr, nullptr, nullptr, (unsigned long long*)col_keys, 0, &kp);
}

int sphk_iou_pairwise_ties(int kind, const float* rows, int64_t R, const float* cols, int64_t C, int D, int mode, int edge,
                           const float* row_target, int32_t* col_tie, int32_t row_base, void* workspace, void* stream) {
    if (kind != SPHK_KIND_SPH2POB_EFFICIENT && kind != SPHK_KIND_SPH2POB_STANDARD)
        return fail(SPHK_ERR_UNSUPPORTED, "sphk_iou_pairwise_ties: kind must be a Sph2Pob transform");
    if (R < 0 || C < 0) return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_iou_pairwise_ties: bad R or C");
    if (C == 0) return SPHK_OK;
    if (!col_tie || (R > 0 && !row_target)) return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_iou_pairwise_ties: null pointer");
    cudaError_t e = cudaMemsetAsync(col_tie, 0, (size_t)C * sizeof(int32_t), (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(col_tie)");
    if (R == 0) return SPHK_OK;
    return pairwise_impl(kind, rows, R, cols, C, D, mode, edge, SPHK_ANGLE_EQUATOR, nullptr, C, nullptr, nullptr, nullptr, nullptr,
                         row_base, 0, workspace, stream, row_target, col_tie);
}

int sphk_unpack_gathered_keys(const uint64_t* gathered, int32_t world, int64_t n_long, int64_t n_short, int64_t cap,
                              float* long_max, int64_t* long_arg, float* short_max, int64_t* short_arg, void* stream) {
    if (world < 1 || n_long < 0 || n_short < 0 || cap < 0) return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_unpack_gathered_keys: bad sizes");
    if (cap < (n_long + world - 1) / world) return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_unpack_gathered_keys: cap < ceil(n_long / world)");
    if (n_long + n_short == 0) return SPHK_OK;
    if (!gathered || (n_long > 0 && (!long_max || !long_arg)) || (n_short > 0 && (!short_max || !short_arg)))
        return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_unpack_gathered_keys: null pointer");
    k_unpack_gathered<<<blocks_for(n_long + n_short), kThreads, 0, (cudaStream_t)stream>>>(
        (const unsigned long long*)gathered, world, n_long, n_short, cap, long_max, long_arg, short_max, short_arg);
    SPHK_LAUNCH_CHECK("k_unpack_gathered");
    return SPHK_OK;
}

int sphk_unpack_peer_keys(const uint64_t* const* peer_bufs, int32_t rank, int32_t world, uint64_t step, int64_t block_offset,
                          int64_t flag_offset, int64_t n_long, int64_t n_short, int64_t cap, int32_t long_parts, int32_t long_pushed,
                          float* long_max, int64_t* long_arg, float* short_max, int64_t* short_arg, void* stream) {
    if (long_parts < 1 || long_parts > 64 || (!long_pushed && long_parts != 1))
        return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_unpack_peer_keys: long_parts must be 1..64, and 1 unless the keys were pushed");
    if (world < 1 || world > 16 || rank < 0 || rank >= world || n_long < 0 || n_short < 0 || cap < 0 || block_offset < 0 || flag_offset < 0)
        return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_unpack_peer_keys: bad rank, world (1..16) or sizes");
    if (cap < (n_long + world - 1) / world) return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_unpack_peer_keys: cap < ceil(n_long / world)");
    if (step == 0) return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_unpack_peer_keys: steps count from 1 (the flags start at 0)");
    if (!peer_bufs || (n_long > 0 && (!long_max || !long_arg)) || (n_short > 0 && (!short_max || !short_arg)))
        return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_unpack_peer_keys: null pointer");
    // (at least one CTA even when there is nothing to unpack: the flags must be exchanged every step)
    const int64_t need = blocks_for(n_long + n_short > 0 ? n_long + n_short : 1), cap_g = 8ll * sm_count();
    const unsigned g = (unsigned)(need < cap_g ? need : cap_g);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(g, 1, 1);
    cfg.blockDim = dim3(kThreads, 1, 1);
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = g_no_pdl ? 0 : 1;
    const cudaError_t le = cudaLaunchKernelEx(&cfg, k_unpack_peers, (unsigned long long* const*)peer_bufs, (int)rank, (int)world,
                                              (unsigned long long)step, block_offset, flag_offset, n_long, n_short, cap, (int)long_parts,
                                              long_pushed != 0, long_max, long_arg, short_max, short_arg);
    if (le != cudaSuccess) return cuda_fail(le, "cudaLaunchKernelEx(k_unpack_peers)");
    return SPHK_OK;
}

static inline int64_t align16(int64_t x) { return (x + 15) & ~15ll; }

int64_t sphk_max_iou_assign_workspace_bytes(int64_t sumK, int64_t N, int32_t batch) {
    if (sumK < 0 || N < 0 || batch < 0) return 0;
    // row keys, column keys per image, records, tie targets, per-image zero rows, per (image, anchor) claims, offsets,
    // per (column tile, GT) maxima for the tie pass
    return align16(sumK * 8) + align16((int64_t)batch * N * 8) + (sumK + N) * (int64_t)((kBoxRecFloats + kBoxCullFloats) * sizeof(float)) +
           align16(sumK * 4) + align16((int64_t)batch * 4) + align16((int64_t)batch * N * 4) + align16(((int64_t)batch + 1) * 4) +
           align16(sumK * ((N + kThreads - 1) / kThreads) * 4);
}

int sphk_max_iou_assign(int kind, const float* gts, const int32_t* gt_offsets_host, int32_t batch, const float* boxes,
                        int64_t N, int D, float pos_iou_thr, float neg_iou_lo, float neg_iou_hi, float min_pos_iou,
                        int gt_max_assign_all, int match_low_quality, const int64_t* gt_labels, int64_t* gt_inds,
                        float* max_overlaps, int64_t* labels, void* workspace, void* stream) {
    if (kind != SPHK_KIND_SPH2POB_EFFICIENT && kind != SPHK_KIND_SPH2POB_STANDARD)
        return fail(SPHK_ERR_UNSUPPORTED, "sphk_max_iou_assign: kind must be a Sph2Pob transform");
    if (batch < 0 || N < 0 || (D != 4 && D != 5) || !gt_offsets_host) return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_max_iou_assign: bad batch, N, D or offsets");
    if (batch == 0 || N == 0) return SPHK_OK;
    if (gt_offsets_host[0] != 0) return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_max_iou_assign: gt_offsets[0] must be 0");
    int64_t max_rows = 0;
    for (int b = 0; b < batch; ++b) {
        const int64_t k = (int64_t)gt_offsets_host[b + 1] - gt_offsets_host[b];
        if (k < 0) return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_max_iou_assign: gt_offsets must be non-decreasing");
        max_rows = k > max_rows ? k : max_rows;
    }
    const int64_t sumK = gt_offsets_host[batch];
    if (!boxes || !gt_inds || !max_overlaps || !workspace || (sumK > 0 && !gts)) return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_max_iou_assign: null pointer");
    if (!aligned16(workspace)) return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_max_iou_assign: workspace must be 16-byte aligned");
    if ((int64_t)batch * N > 0x7FFFFFFFll * (int64_t)kThreads) return fail(SPHK_ERR_UNSUPPORTED, "sphk_max_iou_assign: batch * N too large");
    cudaStream_t s = (cudaStream_t)stream;
    char* w = (char*)workspace;
    unsigned long long* rkey = (unsigned long long*)w;            w += align16(sumK * 8);
    unsigned long long* ckey = (unsigned long long*)w;            w += align16((int64_t)batch * N * 8);
    float4* rec = (float4*)w;                                      w += (sumK + N) * (int64_t)(kBoxRecFloats * sizeof(float));
    float4* cull = (float4*)w;                                     w += (sumK + N) * (int64_t)(kBoxCullFloats * sizeof(float));
    float* target = (float*)w;                                     w += align16(sumK * 4);
    int* zero_row = (int*)w;                                       w += align16((int64_t)batch * 4);
    int* last = (int*)w;                                           w += align16((int64_t)batch * N * 4);
    int32_t* offsets = (int32_t*)w;                                w += align16(((int64_t)batch + 1) * 4);
    float* tile_rmax = (float*)w;      // [column tiles][sumK], written by every tile of pass 1 before pass 2 reads it
    cudaError_t e = cudaMemcpyAsync(offsets, gt_offsets_host, ((size_t)batch + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpyAsync(gt_offsets)");
    // one memset covers row keys + column keys; another one the zero rows + claims
    e = cudaMemsetAsync(rkey, 0, (size_t)(align16(sumK * 8) + align16((int64_t)batch * N * 8)), s);
    if (e == cudaSuccess) e = cudaMemsetAsync(zero_row, 0, (size_t)(align16((int64_t)batch * 4) + align16((int64_t)batch * N * 4)), s);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(assign workspace)");
    if (sumK > 0) {
        // pass 1: per-GT and per-(image, anchor) max / argmax
        int rc = launch_pairwise2(kind, gts, sumK, boxes, N, D, SPHK_MODE_IOU, SPHK_EDGE_ARC, rec, cull, nullptr, N, rkey, ckey, 0, 0,
                                  nullptr, nullptr, offsets, N, batch, max_rows, s, false, tile_rmax);
        if (rc != SPHK_OK) return rc;
        if (match_low_quality) {
            k_assign_targets<<<blocks_for(sumK), kThreads, 0, s>>>(rkey, sumK, offsets, batch, N, min_pos_iou, gt_max_assign_all != 0,
                                                                   target, zero_row, last);
            if (gt_max_assign_all) {
                // pass 2: which anchors tie the row maxima (same kernel, same operands: bit-identical overlaps)
                rc = launch_pairwise2(kind, gts, sumK, boxes, N, D, SPHK_MODE_IOU, SPHK_EDGE_ARC, rec, cull, nullptr, N, nullptr, nullptr,
                                      0, 0, target, last, offsets, N, batch, max_rows, s, false, tile_rmax, true);
                if (rc != SPHK_OK) return rc;
            }
        }
    }
    k_assign_epilogue<<<blocks_for((int64_t)batch * N), kThreads, 0, s>>>(ckey, last, zero_row, offsets, batch, N, pos_iou_thr, neg_iou_lo,
                                                                         neg_iou_hi, match_low_quality != 0, gt_max_assign_all != 0,
                                                                         gt_labels, gt_inds, max_overlaps, labels);
    SPHK_LAUNCH_CHECK("k_assign_epilogue");
    return SPHK_OK;
}

int sphk_loss_fwd_bwd(const float* pred, const float* target, int64_t n, int D, float* iou, const float* grad_iou,
                      float* grad_pred, float* grad_target, void* stream) {
    if (n < 0 || (D != 4 && D != 5)) return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_loss_fwd_bwd: n < 0 or D not in {4,5}");
    if (n == 0) return SPHK_OK;
    if (!pred || !target) return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_loss_fwd_bwd: null box pointer");
    cudaStream_t s = (cudaStream_t)stream;
    const bool v = aligned16(pred) && aligned16(target) && (!grad_pred || aligned16(grad_pred)) &&
                   (!grad_target || aligned16(grad_target));
    if (D == 4) k_loss_fwd_bwd<4><<<blocks_for(n), kThreads, 0, s>>>(pred, target, n, iou, grad_iou, grad_pred, grad_target, v);
    else k_loss_fwd_bwd<5><<<blocks_for(n), kThreads, 0, s>>>(pred, target, n, iou, grad_iou, grad_pred, grad_target, v);
    SPHK_LAUNCH_CHECK("k_loss_fwd_bwd");
    return SPHK_OK;
}

int64_t sphk_loss_reduce_partials(int64_t n) { return n <= 0 ? 0 : (n + kThreads - 1) / kThreads; }

int sphk_loss_reduce(const float* pred, const float* target, const float* weight, int64_t n, int D, float scale,
                     float* partial, float* grad_pred, float* grad_target, void* stream) {
    if (n < 0 || (D != 4 && D != 5)) return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_loss_reduce: n < 0 or D not in {4,5}");
    if (n == 0) return SPHK_OK;
    if (!pred || !target || !partial) return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_loss_reduce: null pointer");
    cudaStream_t s = (cudaStream_t)stream;
    const bool v = aligned16(pred) && aligned16(target) && (!grad_pred || aligned16(grad_pred)) &&
                   (!grad_target || aligned16(grad_target));
    if (D == 4) k_loss_reduce<4><<<blocks_for(n), kThreads, 0, s>>>(pred, target, weight, n, scale, partial, grad_pred, grad_target, v, nullptr, nullptr);
    else k_loss_reduce<5><<<blocks_for(n), kThreads, 0, s>>>(pred, target, weight, n, scale, partial, grad_pred, grad_target, v, nullptr, nullptr);
    SPHK_LAUNCH_CHECK("k_loss_reduce");
    return SPHK_OK;
}

int64_t sphk_loss_total_scratch_bytes(int64_t n) { return n <= 0 ? 16 : (((n + kThreads - 1) / kThreads * 4 + 15) & ~15ll) + 16; }

int sphk_loss_reduce_total(const float* pred, const float* target, const float* weight, int64_t n, int D, float scale,
                           float* total, void* scratch, float* grad_pred, float* grad_target, void* stream) {
    if (n < 0 || (D != 4 && D != 5)) return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_loss_reduce_total: n < 0 or D not in {4,5}");
    if (!total) return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_loss_reduce_total: null pointer");
    cudaStream_t s = (cudaStream_t)stream;
    if (n == 0) {
        const cudaError_t e = cudaMemsetAsync(total, 0, sizeof(float), s);
        return e == cudaSuccess ? SPHK_OK : cuda_fail(e, "cudaMemsetAsync(total)");
    }
    if (!pred || !target || !scratch || !aligned16(scratch)) return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_loss_reduce_total: null or unaligned pointer");
    unsigned* ticket = (unsigned*)scratch;                   // first 16 bytes: the ticket counter (zero between calls)
    float* partial = (float*)((char*)scratch + 16);
    const bool v = aligned16(pred) && aligned16(target) && (!grad_pred || aligned16(grad_pred)) &&
                   (!grad_target || aligned16(grad_target));
    if (D == 4) k_loss_reduce<4><<<blocks_for(n), kThreads, 0, s>>>(pred, target, weight, n, scale, partial, grad_pred, grad_target, v, total, ticket);
    else k_loss_reduce<5><<<blocks_for(n), kThreads, 0, s>>>(pred, target, weight, n, scale, partial, grad_pred, grad_target, v, total, ticket);
    SPHK_LAUNCH_CHECK("k_loss_reduce");
    return SPHK_OK;
}

int sphk_obb_fwd(int kind, const float* b1, const float* b2, int64_t n, int D, int edge, float* obb1, float* obb2,
                 void* stream) {
    if (n < 0 || (D != 4 && D != 5)) return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_obb_fwd: n < 0 or D not in {4,5}");
    if (kind != SPHK_KIND_SPH2POB_EFFICIENT && kind != SPHK_KIND_SPH2POB_STANDARD)
        return fail(SPHK_ERR_UNSUPPORTED, "sphk_obb_fwd: kind must be a Sph2Pob transform");
    if (edge < 0 || edge > 2) return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_obb_fwd: unknown edge");
    if (n == 0) return SPHK_OK;
    if (!b1 || !b2 || !obb1 || !obb2) return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_obb_fwd: null pointer");
    cudaStream_t s = (cudaStream_t)stream;
    const bool v = aligned16(b1) && aligned16(b2);
    if (D == 4) k_obb_fwd<4><<<blocks_for(n), kThreads, 0, s>>>(kind, b1, b2, n, edge, obb1, obb2, v);
    else k_obb_fwd<5><<<blocks_for(n), kThreads, 0, s>>>(kind, b1, b2, n, edge, obb1, obb2, v);
    SPHK_LAUNCH_CHECK("k_obb_fwd");
    return SPHK_OK;
}

int sphk_obb_bwd(int kind, const float* b1, const float* b2, int64_t n, int D, int edge, const float* grad_obb1,
                 const float* grad_obb2, float* grad_b1, float* grad_b2, void* stream) {
    if (n < 0 || (D != 4 && D != 5)) return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_obb_bwd: n < 0 or D not in {4,5}");
    if (kind != SPHK_KIND_SPH2POB_EFFICIENT && kind != SPHK_KIND_SPH2POB_STANDARD)
        return fail(SPHK_ERR_UNSUPPORTED, "sphk_obb_bwd: kind must be a Sph2Pob transform");
    if (edge < 0 || edge > 2) return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_obb_bwd: unknown edge");
    if (n == 0) return SPHK_OK;
    if (!b1 || !b2) return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_obb_bwd: null box pointer");
    cudaStream_t s = (cudaStream_t)stream;
    const bool v = aligned16(b1) && aligned16(b2) && (!grad_b1 || aligned16(grad_b1)) && (!grad_b2 || aligned16(grad_b2));
    if (D == 4) k_obb_bwd<4><<<blocks_for(n), kThreads, 0, s>>>(kind, b1, b2, n, edge, grad_obb1, grad_obb2, grad_b1, grad_b2, v);
    else k_obb_bwd<5><<<blocks_for(n), kThreads, 0, s>>>(kind, b1, b2, n, edge, grad_obb1, grad_obb2, grad_b1, grad_b2, v);
    SPHK_LAUNCH_CHECK("k_obb_bwd");
    return SPHK_OK;
}

int sphk_riou_fwd_bwd(const float* obb1, const float* obb2, int64_t n, float* iou, const float* grad_iou,
                      float* grad_obb1, float* grad_obb2, void* stream) {
    if (n < 0) return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_riou_fwd_bwd: n < 0");
    if (n == 0) return SPHK_OK;
    if (!obb1 || !obb2) return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_riou_fwd_bwd: null OBB pointer");
    k_riou_fwd_bwd<<<blocks_for(n), kThreads, 0, (cudaStream_t)stream>>>(obb1, obb2, n, iou, grad_iou, grad_obb1, grad_obb2);
    SPHK_LAUNCH_CHECK("k_riou_fwd_bwd");
    return SPHK_OK;
}

static int obb_loss_impl(int loss_kind, int fun, int flags, float tau, float alpha, float beta, float eps, int transform,
                         const float* pred, const float* target, int64_t n, int D, const float* upstream, int up_cols, float scale,
                         float* loss, float* partial, float* grad_pred, float* grad_target, float* total, unsigned* ticket,
                         void* stream) {
    if (n < 0 || (D != 4 && D != 5)) return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_obb_loss: n < 0 or D not in {4,5}");
    if (loss_kind < SPHK_LOSS_GWD || loss_kind > SPHK_LOSS_L1) return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_obb_loss: unknown loss kind");
    if (fun < 0 || fun > 2) return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_obb_loss: unknown fun");
    if (transform != SPHK_KIND_SPH2POB_EFFICIENT && transform != SPHK_KIND_SPH2POB_STANDARD)
        return fail(SPHK_ERR_UNSUPPORTED, "sphk_obb_loss: transform must be a Sph2Pob transform");
    const int L = loss_kind == SPHK_LOSS_L1 ? 5 : 1;
    if (upstream && up_cols != 1 && up_cols != L) return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_obb_loss: up_cols must be 1 or the loss's column count");
    if (loss_kind != SPHK_LOSS_L1 && loss_kind != SPHK_LOSS_KFIOU && !(alpha != 0.0f))
        return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_obb_loss: alpha must be non-zero");
    if (loss_kind == SPHK_LOSS_KFIOU && !(beta > 0.0f)) return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_obb_loss: beta must be positive");
    if (n == 0) return SPHK_OK;
    if (!pred || !target) return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_obb_loss: null box pointer");
    LossParams lp;
    lp.kind = loss_kind; lp.fun = fun; lp.flags = flags; lp.tau = tau; lp.alpha = alpha; lp.beta = beta; lp.eps = eps;
    cudaStream_t s = (cudaStream_t)stream;
    const bool v = aligned16(pred) && aligned16(target) && (!grad_pred || aligned16(grad_pred)) &&
                   (!grad_target || aligned16(grad_target));
    if (D == 4) k_obb_loss<4, SPHK_LOSS_SCALAR><<<blocks_for(n), kThreads, 0, s>>>(lp, transform, pred, target, n, upstream, up_cols, scale, loss, partial, grad_pred, grad_target, v, total, ticket);
    else k_obb_loss<5, SPHK_LOSS_SCALAR><<<blocks_for(n), kThreads, 0, s>>>(lp, transform, pred, target, n, upstream, up_cols, scale, loss, partial, grad_pred, grad_target, v, total, ticket);
    SPHK_LAUNCH_CHECK("k_obb_loss");
    return SPHK_OK;
}

int sphk_obb_loss(int loss_kind, int fun, int flags, float tau, float alpha, float beta, float eps, int transform,
                  const float* pred, const float* target, int64_t n, int D, const float* upstream, int up_cols, float scale,
                  float* loss, float* partial, float* grad_pred, float* grad_target, void* stream) {
    return obb_loss_impl(loss_kind, fun, flags, tau, alpha, beta, eps, transform, pred, target, n, D, upstream, up_cols, scale, loss,
                         partial, grad_pred, grad_target, nullptr, nullptr, stream);
}

int sphk_obb_loss_total(int loss_kind, int fun, int flags, float tau, float alpha, float beta, float eps, int transform,
                        const float* pred, const float* target, int64_t n, int D, const float* upstream, int up_cols, float scale,
                        float* total, void* scratch, float* grad_pred, float* grad_target, void* stream) {
    if (!total) return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_obb_loss_total: null pointer");
    if (n == 0) {
        const cudaError_t e = cudaMemsetAsync(total, 0, sizeof(float), (cudaStream_t)stream);
        return e == cudaSuccess ? SPHK_OK : cuda_fail(e, "cudaMemsetAsync(total)");
    }
    if (!scratch || !aligned16(scratch)) return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_obb_loss_total: null or unaligned scratch");
    // scratch as for sphk_loss_reduce_total: the ticket counter (zero between calls) in its first 16 bytes, then the per-block sums
    return obb_loss_impl(loss_kind, fun, flags, tau, alpha, beta, eps, transform, pred, target, n, D, upstream, up_cols, scale, nullptr,
                         (float*)((char*)scratch + 16), grad_pred, grad_target, total, (unsigned*)scratch, stream);
}

static int coder_params(const char* who, int D, const float* means, const float* stds, float wh_ratio_clip, int clip_border,
                        int add_ctr_clamp, float ctr_clamp, CoderParams* cp) {
    if (D != 4 && D != 5) return fail(SPHK_ERR_INVALID_ARGUMENT, "coder: D not in {4,5}");
    if (!(wh_ratio_clip > 0.0f)) return fail(SPHK_ERR_INVALID_ARGUMENT, "coder: wh_ratio_clip must be positive");
    for (int k = 0; k < 5; ++k) {
        cp->mean[k] = (means && k < D) ? means[k] : 0.0f;
        cp->stdv[k] = (stds && k < D) ? stds[k] : 1.0f;
    }
    cp->max_ratio = fabsf(logf(wh_ratio_clip));
    cp->ctr_clamp = ctr_clamp;
    cp->clip_border = clip_border ? 1 : 0;
    cp->add_ctr_clamp = add_ctr_clamp ? 1 : 0;
    (void)who;
    return SPHK_OK;
}

int sphk_anchor_targets(const int64_t* gt_inds, int32_t batch, int64_t N, int D, const float* anchors, const float* gts,
                        const int64_t* gt_labels, const int32_t* gt_offsets, int64_t num_classes, float pos_weight,
                        int reg_decoded_bbox, const float* means, const float* stds, int64_t* labels, float* label_weights,
                        float* bbox_targets, float* bbox_weights, int32_t* counts, void* stream) {
    if (batch < 0 || N < 0 || (D != 4 && D != 5)) return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_anchor_targets: bad batch, N or D");
    if (batch > 65535) return fail(SPHK_ERR_UNSUPPORTED, "sphk_anchor_targets: more than 65535 images");
    if (batch == 0) return SPHK_OK;
    if (!counts) return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_anchor_targets: null counts");
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(counts, 0, (size_t)batch * 2 * sizeof(int32_t), s);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(counts)");
    if (N == 0) return SPHK_OK;
    if (!gt_inds || !anchors || !gt_offsets || !labels || !label_weights || !bbox_targets || !bbox_weights)
        return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_anchor_targets: null pointer");
    CoderParams cp;
    const int rc = coder_params("sphk_anchor_targets", D, means, stds, 1.0f, 0, 0, 0.0f, &cp);
    if (rc != SPHK_OK) return rc;
    const dim3 grid(blocks_for(N), (unsigned)batch);
    if (D == 4) k_anchor_targets<4><<<grid, kThreads, 0, s>>>(gt_inds, N, anchors, gts, gt_labels, gt_offsets, num_classes, pos_weight, reg_decoded_bbox != 0, cp, labels, label_weights, bbox_targets, bbox_weights, counts);
    else k_anchor_targets<5><<<grid, kThreads, 0, s>>>(gt_inds, N, anchors, gts, gt_labels, gt_offsets, num_classes, pos_weight, reg_decoded_bbox != 0, cp, labels, label_weights, bbox_targets, bbox_weights, counts);
    SPHK_LAUNCH_CHECK("k_anchor_targets");
    return SPHK_OK;
}

static int coder_decode_impl(const float* rois, const float* deltas, const float* grad_out, int64_t n, int D, const float* means,
                             const float* stds, float wh_ratio_clip, int clip_border, int add_ctr_clamp, float ctr_clamp,
                             float* out, void* stream) {
    CoderParams cp;
    const int rc = coder_params("decode", D, means, stds, wh_ratio_clip, clip_border, add_ctr_clamp, ctr_clamp, &cp);
    if (rc != SPHK_OK) return rc;
    if (n < 0) return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_coder_decode: n < 0");
    if (n == 0) return SPHK_OK;
    if (!rois || !deltas || !out) return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_coder_decode: null pointer");
    const bool v = aligned16(rois) && aligned16(deltas) && aligned16(out) && (!grad_out || aligned16(grad_out));
    cudaStream_t s = (cudaStream_t)stream;
    if (D == 4) k_coder_decode<4><<<blocks_for(n), kThreads, 0, s>>>(rois, deltas, grad_out, n, cp, out, v);
    else k_coder_decode<5><<<blocks_for(n), kThreads, 0, s>>>(rois, deltas, grad_out, n, cp, out, v);
    SPHK_LAUNCH_CHECK("k_coder_decode");
    return SPHK_OK;
}

int sphk_coder_decode(const float* rois, const float* deltas, int64_t n, int D, const float* means, const float* stds,
                      float wh_ratio_clip, int clip_border, int add_ctr_clamp, float ctr_clamp, float* out, void* stream) {
    return coder_decode_impl(rois, deltas, nullptr, n, D, means, stds, wh_ratio_clip, clip_border, add_ctr_clamp, ctr_clamp, out, stream);
}

int sphk_coder_decode_bwd(const float* rois, const float* deltas, const float* grad_out, int64_t n, int D, const float* means,
                          const float* stds, float wh_ratio_clip, int clip_border, int add_ctr_clamp, float ctr_clamp,
                          float* grad_deltas, void* stream) {
    if (n > 0 && !grad_out) return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_coder_decode_bwd: null grad_out");
    return coder_decode_impl(rois, deltas, grad_out, n, D, means, stds, wh_ratio_clip, clip_border, add_ctr_clamp, ctr_clamp,
                             grad_deltas, stream);
}

int sphk_box_format(int fmt, const float* in, int64_t n, int d_in, int d_out, float img_h, float img_w, float* out, void* stream) {
    if (fmt < 0 || fmt >= FMT_COUNT) return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_box_format: unknown format");
    if (n < 0) return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_box_format: n < 0");
    const bool planar2sph = (fmt == FMT_PLANAR2SPH_PIX || fmt == FMT_PLANAR2SPH_TAN);
    const bool keeps = (fmt == FMT_GEO2SPH || fmt == FMT_SPH2GEO || fmt == FMT_SPH2PLANAR_PIX || fmt == FMT_SPH2PLANAR_TAN);
    if ((d_in != 4 && d_in != 5) || (d_out != 4 && d_out != 5)) return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_box_format: columns must be 4 or 5");
    if (keeps ? (d_out != d_in) : (d_in != box_format_cols_in(fmt, d_out) || (!planar2sph && d_out != box_format_cols_out(fmt, d_in))))
        return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_box_format: column counts do not fit the format");
    if (!(img_h > 0.0f) || !(img_w > 0.0f)) return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_box_format: image size must be positive");
    if (n == 0) return SPHK_OK;
    if (!in || !out) return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_box_format: null pointer");
    if (d_in == 4 && d_out == 4 && aligned16(in) && aligned16(out))
        k_box_format4<<<blocks_for((n + 3) / 4), kThreads, 0, (cudaStream_t)stream>>>((const float4*)in, n, fmt, img_h, img_w, (float4*)out);
    else
        k_box_format<<<blocks_for(n), kThreads, 0, (cudaStream_t)stream>>>(in, n, fmt, d_in, d_out, img_h, img_w, out);
    SPHK_LAUNCH_CHECK("k_box_format");
    return SPHK_OK;
}

int sphk_coder_encode(const float* proposals, const float* gt, int64_t n, int D, const float* means, const float* stds,
                      float* out, void* stream) {
    CoderParams cp;
    const int rc = coder_params("encode", D, means, stds, 1.0f, 0, 0, 0.0f, &cp);
    if (rc != SPHK_OK) return rc;
    if (n < 0) return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_coder_encode: n < 0");
    if (n == 0) return SPHK_OK;
    if (!proposals || !gt || !out) return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_coder_encode: null pointer");
    const bool v = aligned16(proposals) && aligned16(gt) && aligned16(out);
    cudaStream_t s = (cudaStream_t)stream;
    if (D == 4) k_coder_encode<4><<<blocks_for(n), kThreads, 0, s>>>(proposals, gt, n, cp, out, v);
    else k_coder_encode<5><<<blocks_for(n), kThreads, 0, s>>>(proposals, gt, n, cp, out, v);
    SPHK_LAUNCH_CHECK("k_coder_encode");
    return SPHK_OK;
}

// grid of the fused decode + loss kernel: persistent warps, at most 2 CTAs per SM (~105 registers per thread),
// at least 8 chunks of 32 rows per warp
static int64_t decode_loss_grid(int64_t n) {
    if (n <= 0) return 0;
    const int64_t chunks = (n + 31) / 32, wpc = kThreads / 32;
    int64_t g = (chunks + wpc * 8 - 1) / (wpc * 8);
    const int64_t wave = 2ll * sm_count();
    return g > wave ? wave : g;
}

int64_t sphk_decode_loss_partials(int64_t n) { return decode_loss_grid(n); }

int sphk_decode_loss_reduce(const float* anchors, const float* deltas, const float* target, const float* weight,
                            int weight_cols, int64_t n, int D, const float* means, const float* stds, float wh_ratio_clip,
                            int clip_border, int add_ctr_clamp, float ctr_clamp, float scale, float* partial,
                            float* grad_deltas, void* stream) {
    CoderParams cp;
    const int rc = coder_params("decode_loss", D, means, stds, wh_ratio_clip, clip_border, add_ctr_clamp, ctr_clamp, &cp);
    if (rc != SPHK_OK) return rc;
    if (n < 0 || n > 0x7FFFFFFFll) return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_decode_loss_reduce: n < 0 or n >= 2^31");
    if (n == 0) return SPHK_OK;
    if (!anchors || !deltas || !target || !partial) return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_decode_loss_reduce: null pointer");
    if (weight_cols < 0 || weight_cols > 8 || (weight_cols > 0 && !weight) || (weight_cols == 0 && weight))
        return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_decode_loss_reduce: weight / weight_cols mismatch (0 with NULL, else 1..8)");
    const unsigned g = (unsigned)decode_loss_grid(n);
    cudaStream_t s = (cudaStream_t)stream;
    if (D == 4) k_decode_loss<4><<<g, kThreads, 0, s>>>(anchors, deltas, target, weight, weight_cols, n, cp, scale, partial, grad_deltas);
    else k_decode_loss<5><<<g, kThreads, 0, s>>>(anchors, deltas, target, weight, weight_cols, n, cp, scale, partial, grad_deltas);
    SPHK_LAUNCH_CHECK("k_decode_loss");
    return SPHK_OK;
}

int sphk_nms_batched(const float* boxes, const int32_t* order, const int32_t* seg_offsets, int32_t S, int32_t max_seg_len,
                     int32_t typical_seg_len, int D, int kind, float iou_threshold, uint8_t* keep, void* stream) {
    if (S < 0 || max_seg_len < 0 || (D != 4 && D != 5)) return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_nms_batched: bad S, max_seg_len or D");
    if (kind != SPHK_KIND_SPH2POB_EFFICIENT && kind != SPHK_KIND_NAIVE && kind != SPHK_KIND_UNBIASED && kind != (SPHK_KIND_NAIVE | SPHK_NMS_RULE_GT))
        return fail(SPHK_ERR_UNSUPPORTED, "sphk_nms_batched: kind must be sph2pob_efficient, naive[ | SPHK_NMS_RULE_GT] or unbiased (SphNMS, sph_nms.py:8-16)");
    if (S == 0 || max_seg_len == 0) return SPHK_OK;
    if (!boxes || !order || !seg_offsets || !keep) return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_nms_batched: null pointer");
    const int max_words = (max_seg_len + 31) / 32;
    const size_t smem = (size_t)max_words * 33u * sizeof(uint32_t);
    if (smem > 200u * 1024u) return fail(SPHK_ERR_UNSUPPORTED, "sphk_nms_batched: segment longer than 49,000 boxes");
    cudaStream_t s = (cudaStream_t)stream;
    const bool v = aligned16(boxes);
    // one CTA per segment; its width follows the typical segment length (the longest one when no hint is given):
    // a 32-pivot round has up to 32 * k/32 (pivot, word) units, one warp each
    const int tl = (typical_seg_len > 0 && typical_seg_len < max_seg_len) ? typical_seg_len : max_seg_len;
    const int nt = tl <= 64 ? 128 : (tl <= 256 ? 256 : (tl <= 512 ? 512 : 1024));
    cudaError_t e;
    if (D == 4) {
        e = allow_max_dynamic_smem<4>(k_nms<4>);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(k_nms)");
        k_nms<4><<<S, nt, smem, s>>>(boxes, const_cast<int32_t*>(order), seg_offsets, nullptr, nullptr, 0, nullptr, 1, iou_threshold, keep, max_words, v, kind);
    } else {
        e = allow_max_dynamic_smem<5>(k_nms<5>);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(k_nms)");
        k_nms<5><<<S, nt, smem, s>>>(boxes, const_cast<int32_t*>(order), seg_offsets, nullptr, nullptr, 0, nullptr, 1, iou_threshold, keep, max_words, v, kind);
    }
    SPHK_LAUNCH_CHECK("k_nms");
    return SPHK_OK;
}

int64_t sphk_nms_images_workspace_bytes(int32_t num_images, int32_t per_image, int32_t num_classes) {
    if (num_images < 0 || per_image < 0 || num_classes < 0) return 0;
    const int64_t M = (int64_t)num_images * per_image, S = (int64_t)num_images * num_classes;
    // order, segment starts, lengths, keep flags, per-image "label out of range" flags
    return align16(M * 4) + align16(S * 4) + align16(S * 4) + align16(M) + align16((int64_t)num_images * 4);
}

int sphk_nms_images(const float* boxes, const float* scores, const int64_t* labels, const uint8_t* valid, int32_t num_images,
                    int32_t per_image, int32_t num_classes, int D, int kind, float iou_threshold, int32_t max_out, int32_t* out_idx,
                    int32_t* out_count, void* workspace, void* stream) {
    if (num_images < 0 || per_image < 0 || num_classes <= 0 || max_out < 0 || (D != 4 && D != 5))
        return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_nms_images: bad sizes or D");
    if (kind != SPHK_KIND_SPH2POB_EFFICIENT && kind != SPHK_KIND_NAIVE && kind != SPHK_KIND_UNBIASED && kind != (SPHK_KIND_NAIVE | SPHK_NMS_RULE_GT))
        return fail(SPHK_ERR_UNSUPPORTED, "sphk_nms_images: kind must be sph2pob_efficient, naive[ | SPHK_NMS_RULE_GT] or unbiased (SphNMS, sph_nms.py:8-16)");
    if (num_images == 0) return SPHK_OK;
    if (!out_idx || !out_count) return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_nms_images: null output pointer");
    cudaStream_t s = (cudaStream_t)stream;
    if (per_image == 0) {
        cudaError_t e = cudaMemsetAsync(out_count, 0, (size_t)num_images * sizeof(int32_t), s);
        if (e == cudaSuccess && max_out > 0) e = cudaMemsetAsync(out_idx, 0xFF, (size_t)num_images * max_out * sizeof(int32_t), s);
        return e == cudaSuccess ? SPHK_OK : cuda_fail(e, "cudaMemsetAsync(nms outputs)");
    }
    if (!boxes || !scores || !labels || !workspace) return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_nms_images: null pointer");
    if (!aligned16(workspace)) return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_nms_images: workspace must be 16-byte aligned");
    if (per_image > 16384 || num_classes > 0xFFF0)
        return fail(SPHK_ERR_UNSUPPORTED, "sphk_nms_images: more than 16384 candidates per image or 65520 classes; use sphk_nms_batched");
    int Kp = 32;
    while (Kp < per_image) Kp <<= 1;
    const int64_t M = (int64_t)num_images * per_image, S = (int64_t)num_images * num_classes;
    if (M > 0x7FFFFFFFll || S > 0x7FFFFFFFll) return fail(SPHK_ERR_UNSUPPORTED, "sphk_nms_images: batch too large");
    char* w = (char*)workspace;
    int32_t* order = (int32_t*)w;      w += align16(M * 4);
    int32_t* seg_start = (int32_t*)w;  w += align16(S * 4);
    int32_t* seg_len = (int32_t*)w;    w += align16(S * 4);
    uint8_t* keep = (uint8_t*)w;       w += align16(M);
    int32_t* bad = (int32_t*)w;
    cudaError_t e = cudaMemsetAsync(keep, 0, (size_t)(align16(M) + align16((int64_t)num_images * 4)), s);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(keep)");
    const int nt_img = Kp >= 1024 ? 1024 : Kp;
    const size_t smem_sc = (size_t)num_classes * 4, smem_col = (size_t)Kp * 8;
    if (smem_sc > (size_t)kMaxDynamicSmem || smem_col > (size_t)kMaxDynamicSmem)
        return fail(SPHK_ERR_UNSUPPORTED, "sphk_nms_images: more than 51,200 classes; use sphk_nms_batched");
    e = allow_max_dynamic_smem<0>(k_img_scatter);
    if (e == cudaSuccess) e = allow_max_dynamic_smem<1>(k_img_collect);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(k_img_scatter)");
    k_img_scatter<<<num_images, nt_img, smem_sc, s>>>(labels, valid, per_image, num_classes, order, seg_start, seg_len, bad);
    SPHK_LAUNCH_CHECK("k_img_scatter");
    // k_nms: suppression words sized for the longest possible segment (a whole image); segments of up to sort_cap
    // candidates are sorted in shared memory, longer ones in place in global memory
    const int max_words = (per_image + 31) / 32;
    int sort_cap = 32;
    while (sort_cap < per_image && sort_cap < 4096) sort_cap <<= 1;
    const size_t smem = (size_t)((33 * max_words + 1) & ~1) * sizeof(uint32_t) + (size_t)sort_cap * 8;
    // a 32-pivot round has 32 * k/32 (pivot, word) units, one warp each: CTA width for twice the mean segment length
    const int tl = (int)((2 * (int64_t)per_image) / num_classes) + 1;
    const int nt = num_classes >= 512 ? 512 : (tl <= 16 ? 128 : (tl <= 64 ? 256 : (tl <= 256 ? 512 : 1024)));   // >= 512: a cap, not a count
    const bool v = aligned16(boxes);
    if (D == 4) {
        e = allow_max_dynamic_smem<4>(k_nms<4>);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(k_nms)");
        k_nms<4><<<(unsigned)S, nt, smem, s>>>(boxes, order, seg_start, seg_len, scores, sort_cap, bad, num_classes, iou_threshold, keep, max_words, v, kind);
    } else {
        e = allow_max_dynamic_smem<5>(k_nms<5>);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(k_nms)");
        k_nms<5><<<(unsigned)S, nt, smem, s>>>(boxes, order, seg_start, seg_len, scores, sort_cap, bad, num_classes, iou_threshold, keep, max_words, v, kind);
    }
    SPHK_LAUNCH_CHECK("k_nms");
    k_img_collect<<<num_images, nt_img, smem_col, s>>>(scores, order, keep, per_image, Kp, max_out, bad, out_idx, out_count);
    SPHK_LAUNCH_CHECK("k_img_collect");
    return SPHK_OK;
}

int sphk_probe_fp32(int32_t blocks, int32_t iters, float* sink, void* stream) {
    if (blocks <= 0 || iters <= 0 || !sink) return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_probe_fp32: bad arguments");
    k_probe_fp32<<<blocks, kThreads, 0, (cudaStream_t)stream>>>(iters, sink);
    SPHK_LAUNCH_CHECK("k_probe_fp32");
    return SPHK_OK;
}

#ifdef SPHK_TIMELINE
int sphk_debug_timeline_warps(unsigned long long* t, unsigned* slow, int n) {
    if (cudaDeviceSynchronize() != cudaSuccess) return -1;
    if (cudaMemcpyFromSymbol(t, g_tlw, sizeof(unsigned long long) * 3 * n) != cudaSuccess) return -2;
    if (cudaMemcpyFromSymbol(slow, g_tlw_slow, sizeof(unsigned) * n) != cudaSuccess) return -3;
    return 0;
}
int sphk_debug_timeline(unsigned long long* t, unsigned* sm, int n) {
    if (cudaDeviceSynchronize() != cudaSuccess) return -1;
    if (cudaMemcpyFromSymbol(t, g_tl, sizeof(unsigned long long) * 2 * n) != cudaSuccess) return -2;
    if (cudaMemcpyFromSymbol(sm, g_tl_sm, sizeof(unsigned) * n) != cudaSuccess) return -3;
    return 0;
}
#endif

int sphk_prefilter_count(const float* rows, int64_t R, const float* cols, int64_t C, int D, int edge, uint64_t* live_count,
                         void* workspace, void* stream) {
    if (R < 0 || C < 0 || (D != 4 && D != 5) || edge < 0 || edge > 2) return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_prefilter_count: bad R, C, D or edge");
    if (!live_count) return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_prefilter_count: null pointer");
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(live_count, 0, sizeof(uint64_t), s);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(live_count)");
    if (R == 0 || C == 0) return SPHK_OK;
    if (!rows || !cols || !workspace || !aligned16(workspace)) return fail(SPHK_ERR_INVALID_ARGUMENT, "sphk_prefilter_count: null or unaligned pointer");
    const int64_t row_tiles = (R + 31) / 32, col_tiles = (C + kThreads - 1) / kThreads;
    if (row_tiles > 0x7FFFFFFFll || col_tiles > 65535) return fail(SPHK_ERR_UNSUPPORTED, "sphk_prefilter_count: grid too large; shard the call");
    float4* rec = (float4*)((char*)workspace + keys_bytes(R, C));
    float4* cull = rec + (R + C) * 4;
    if (D == 4) k_box_pre<4><<<blocks_for(R + C), kThreads, 0, s>>>(rows, R, cols, C, edge, rec, cull, aligned16(rows), aligned16(cols), nullptr, nullptr, false, 0);
    else k_box_pre<5><<<blocks_for(R + C), kThreads, 0, s>>>(rows, R, cols, C, edge, rec, cull, false, false, nullptr, nullptr, false, 0);
    k_prefilter_count<<<dim3((unsigned)row_tiles, (unsigned)col_tiles), kThreads, 0, s>>>(R, C, cull, (unsigned long long*)live_count,
                                                                                          box_test_pays(R, C) ? (g_no_boxcull ? 0 : 1) : (g_no_sat ? 0 : 2));
    SPHK_LAUNCH_CHECK("k_prefilter_count");
    return SPHK_OK;
}

int sphk_set_dense(int on) {
    const int prev = g_dense;
    g_dense = on ? 1 : 0;
    return prev;
}

}  // extern "C"
