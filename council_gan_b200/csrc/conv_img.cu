// Image-side convolutions (<= 8 input lanes, 64 output channels) on tcgen05 with the im2col patches built IN SHARED MEMORY.
//
// Layers: the council discriminator's first layer (3x3 stride 1 on cat(x, x_input) = 8 lanes, networks.py:138), the
// discriminator's first layer (4x4 stride 2 on 4 lanes, networks.py:40).  Their K = taps x lanes is tiny (72 / 64), so they
// are bound by HBM: the 64-channel full-resolution output (2.15 GB at (1+U)*B = 32 images x 4 members) dwarfs the input.
// What the other paths cost (ncu, profiles/r02_ncu_dc0_*.txt):
//   forward through TMA im2col with 32-byte rows: request-bound, 21 % of DRAM bandwidth (1.41 ms vs 0.37 ms of traffic);
//   weight gradient through an explicit patch matrix: im2col_small writes 3.2 GB, the GEMM reads it back (2.1 ms + 0.4 ms of
//   bias-gradient column sums vs 0.37 ms for one read of dy).
// Here warps gather each pixel's taps from the (L1-resident) image with plain loads, round them to TF32 and store them
// straight into the swizzled shared-memory operand layout; nothing but the image and dy / y ever crosses HBM.
//
//   forward:  D[128 px x 64 co] = A[128 x K] * W^T      A, W K-major 128-byte-swizzled; K padded to 32-float chunks
//   wgrad  :  D[128 k x 64 co] += P^T[px x k] * dY[px x co]   both MN-major (pixels are the reduction dimension); row K of P
//             is all ones, so row K of D is the bias gradient (no separate column-sum pass over dy); one work unit per CTA,
//             deterministic two-phase reduction over the CTAs of a member.
#include "common.cuh"
#include "tc_ptx.cuh"

namespace cg {

thread_local int g_img_path = 1;

constexpr int IMG_BUILD_THREADS = 256;                 // two builder groups of 4 warps, alternating pipeline stages
constexpr int IMG_MMA_WARP = 8;
constexpr int IMG_THREADS = IMG_BUILD_THREADS + 32 + 128;  // + MMA warp + 4 epilogue warps (warp % 4 = TMEM lane quadrant)

struct ImgP {
    CUtensorMap dymap;   // (wgrad) dy [G*Mpix][64] as 32-channel x kp-pixel MN-major boxes; (forward) y [G*Mpix][64] store map, 32 x 128 boxes
    const float* x; const float* w; const float* bias; float* y; float* part;
    int G, xg_images, B, H, W, L, Ho, Wo, KH, KW, stride, pad;
    int K, KC, k8, stages, act, cpg, kp, nsbuf;
    float slope;
    long Mpix, chunk;
};

__device__ __forceinline__ uint8_t* align1k(uint8_t* p) { return reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(p) + 1023) & ~(uintptr_t)1023); }
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// =====================================================================================================================
// forward
// =====================================================================================================================
template <int L, int KW>
__global__ void __launch_bounds__(IMG_THREADS, 1) img_conv_fwd_kernel(const __grid_constant__ ImgP p) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = align1k(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int a_stage = p.KC * 16384;                 // KC chunks of [128 rows x 128 B]
    uint8_t* sB = smem + (size_t)p.stages * a_stage;  // KC chunks of [64 rows x 128 B]
    uint8_t* sStage = sB + (size_t)p.KC * 8192;       // nsbuf output tiles: 2 chunks of [128 rows x 128 B], 128-byte swizzled
    uint64_t* a_full = reinterpret_cast<uint64_t*>(sStage + (size_t)p.nsbuf * 32768);
    uint64_t* a_empty = a_full + p.stages;
    uint64_t* tfull = a_empty + p.stages;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int g = blockIdx.x % p.G, cidx = blockIdx.x / p.G;
    const int pq = p.Ho * p.Wo;
    const int tiles = (int)(((long)p.B * pq) >> 7);  // 128-pixel tiles of this member

    // one-time: zero the A ring (the K-padding units are never written again) and stage this member's weights, TF32-rounded
    for (int i = threadIdx.x; i < p.stages * a_stage / 16; i += blockDim.x) reinterpret_cast<float4*>(smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    const float* wg = p.w + (long)g * 64 * p.K;
    for (int i = threadIdx.x; i < p.KC * 64 * 8; i += blockDim.x) {
        const int u = i & 7, co = (i >> 3) & 63, j = i >> 9;
        const int k = j * 32 + u * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k < p.K) v = to_tf32(ldg4(wg + (long)co * p.K + k));
        *reinterpret_cast<float4*>(sB + j * 8192 + (co >> 3) * 1024 + (co & 7) * 128 + ((u ^ (co & 7)) << 4)) = v;
    }
    if (warp == IMG_MMA_WARP) {
        if (lane == 0) {
            for (int s = 0; s < p.stages; s++) {
                mbar_init(&a_full[s], 128);
                mbar_init(&a_empty[s], 1);
            }
            for (int a = 0; a < 2; a++) {
                mbar_init(&tfull[a], 1);
                mbar_init(&tempty[a], 128);
            }
            fence_barrier_init();
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(128));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    fence_proxy_async();  // the generic-proxy stores above must be visible to the tensor core's shared-memory reads
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < IMG_BUILD_THREADS / 32) {
        // ===================== patch builders: thread = one pixel row of the tile; group gq builds tiles gq, gq+2, ... ===========
        const int gq = warp >> 2, t = threadIdx.x & 127;
        constexpr int TAPS = KW * KW, NV = TAPS * (L / 4);  // float4 loads per pixel: all issued before the first use
        int it = gq;
        for (int tile = cidx + gq * p.cpg; tile < tiles; tile += 2 * p.cpg, it += 2) {
            const int stage = it % p.stages;
            const uint32_t phase = (uint32_t)((it / p.stages) & 1);
            const int m = tile * 128 + t;
            const int img = m / pq;
            const int rem = m - img * pq;
            const int oy = rem / p.Wo, ox = rem - oy * p.Wo;
            const float* xb = p.x + ((long)(g * p.xg_images + img) * p.H * p.W) * L;
            const int iy0 = oy * p.stride - p.pad, ix0 = ox * p.stride - p.pad;
            float4 v[NV];
#pragma unroll
            for (int tap = 0; tap < TAPS; tap++) {
                const int iy = iy0 + tap / KW, ix = ix0 + tap % KW;
                const bool in = iy >= 0 && iy < p.H && ix >= 0 && ix < p.W;
                const float* px = xb + ((long)iy * p.W + ix) * L;
#pragma unroll
                for (int h = 0; h < L / 4; h++) v[tap * (L / 4) + h] = in ? ldg4(px + 4 * h) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            mbar_wait(&a_empty[stage], phase ^ 1);
            uint8_t* sa = smem + (size_t)stage * a_stage + (t >> 3) * 1024 + (t & 7) * 128;
            const int sw = t & 7;
#pragma unroll
            for (int q = 0; q < NV; q++) {  // 16-byte unit q of the pixel's K row: chunk q / 8, swizzled slot (q % 8) ^ (row % 8)
                *reinterpret_cast<float4*>(sa + (q >> 3) * 16384 + (((q & 7) ^ sw) << 4)) = to_tf32(v[q]);
            }
            fence_proxy_async();
            mbar_arrive(&a_full[stage]);
        }
    } else if (warp == IMG_MMA_WARP) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            const uint32_t idesc = make_idesc_tf32(64);
            const uint64_t desc_hi = make_kmajor_sw128_desc(0);
            const uint32_t sb = smem_u32(sB);
            int it = 0, acc = 0;
            uint32_t acc_phase = 0;
            // tiles in issue order: group 0 takes cidx, cidx + 2*cpg, ...; group 1 takes cidx + cpg, ...  -> tile(it) = cidx + it*cpg
            for (int tile = cidx; tile < tiles; tile += p.cpg, it++) {
                const int stage = it % p.stages;
                const uint32_t phase = (uint32_t)((it / p.stages) & 1);
                mbar_wait(&tempty[acc], acc_phase ^ 1);
                tc_fence_after();
                mbar_wait(&a_full[stage], phase);
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + (size_t)stage * a_stage);
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 64);
                for (int q = 0; q < p.k8; q++) {
                    const int j = q >> 2, kk = q & 3;
                    const uint64_t adesc = (desc_hi | (uint64_t)(((sa + j * 16384) & 0x3FFFF) >> 4)) + (uint64_t)(kk * 2);
                    const uint64_t bdesc = (desc_hi | (uint64_t)(((sb + j * 8192) & 0x3FFFF) >> 4)) + (uint64_t)(kk * 2);
                    umma_tf32(d_tmem, adesc, bdesc, idesc, q != 0 ? 1u : 0u);
                }
                umma_commit(&a_empty[stage]);
                umma_commit(&tfull[acc]);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ===================== epilogue: bias + activation -> swizzled shared-memory tile -> TMA tensor store =====================
        // Direct global stores from the accumulator layout (thread = pixel row, 16 bytes per lane at a 256-byte stride) touch 32
        // half-written sectors per instruction and cap at ~2 TB/s (ncu, profiles/r02_runD_*): the tile goes through shared memory
        // in the 128-byte-swizzled layout (conflict-free for row-per-thread writes) and leaves as two bulk tensor stores.
        const int quad = warp & 3;
        const int row = quad * 32 + lane;
        const bool issuer = warp == IMG_MMA_WARP + 1 && lane == 0;
        int acc = 0, sbuf = 0;
        uint32_t acc_phase = 0;
        const float* bp = p.bias ? p.bias + (long)g * 64 : nullptr;
        const int sw = row & 7;
        for (int tile = cidx; tile < tiles; tile += p.cpg) {
            mbar_wait(&tfull[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * 64);
            // the staging buffer about to be overwritten must have been read out by its previous store
            if (issuer) {
                if (p.nsbuf == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            uint8_t* stg = sStage + (size_t)sbuf * 32768 + (row >> 3) * 1024 + (row & 7) * 128;
#pragma unroll
            for (int c0 = 0; c0 < 64; c0 += 32) {
                float v[32];
                tmem_ld32(taddr + (uint32_t)c0, v);
                if (bp) {
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        float4 a = ldg4(bp + c0 + 4 * j);
                        v[4 * j] += a.x; v[4 * j + 1] += a.y; v[4 * j + 2] += a.z; v[4 * j + 3] += a.w;
                    }
                }
                if (p.act == CG_ACT_RELU) {
#pragma unroll
                    for (int j = 0; j < 32; j++) v[j] = fmaxf(v[j], 0.f);
                } else if (p.act == CG_ACT_LRELU) {
#pragma unroll
                    for (int j = 0; j < 32; j++) v[j] = v[j] > 0.f ? v[j] : v[j] * p.slope;
                }
#pragma unroll
                for (int j = 0; j < 8; j++)
                    *reinterpret_cast<float4*>(stg + (c0 >> 5) * 16384 + ((j ^ sw) << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            }
            tc_fence_before();
            mbar_arrive(&tempty[acc]);  // accumulator drained: the MMA warp may start the tile after next
            fence_proxy_async();
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (issuer) {
                const int grow = (int)((long)g * p.B * pq + (long)tile * 128);  // first output row (pixel) of the tile
                const uint32_t src = smem_u32(sStage + (size_t)sbuf * 32768);
                asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(&p.dymap), "r"(0), "r"(grow), "r"(src) : "memory");
                asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(&p.dymap), "r"(32), "r"(grow), "r"(src + 16384) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            if (++sbuf == p.nsbuf) sbuf = 0;
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if (issuer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // all stores complete before the CTA exits
    }
    tc_fence_before();
    __syncthreads();
    if (warp == IMG_MMA_WARP) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128));
    }
}

// =====================================================================================================================
// weight (+ bias) gradient
// =====================================================================================================================
// stage = [4 M groups of 32 k-rows][kp pixels x 128 B] (patches, built here) + [2 N groups of 32 couts][kp x 128 B] (dy, TMA)
template <int L, int KW, int KP>
__global__ void __launch_bounds__(IMG_THREADS, 1) img_conv_wgrad_kernel(const __grid_constant__ ImgP p) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = align1k(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int box = p.kp * 128;
    const int stage_bytes = 6 * box;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
    uint64_t* empty = full + p.stages;
    uint64_t* tfull = empty + p.stages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 2);

    const int g = blockIdx.x % p.G, split = blockIdx.x / p.G;
    const long mbeg = (long)split * p.chunk;
    const long mend = mbeg + p.chunk < p.Mpix ? mbeg + p.chunk : p.Mpix;
    const int nst = mend > mbeg ? (int)((mend - mbeg) / p.kp) : 0;
    const int pq = p.Ho * p.Wo;

    for (int i = threadIdx.x; i < p.stages * stage_bytes / 16; i += blockDim.x) reinterpret_cast<float4*>(smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (warp == IMG_MMA_WARP) {
        if (lane == 0) {
            prefetch_tmap(&p.dymap);
            for (int s = 0; s < p.stages; s++) {
                mbar_init(&full[s], 128);
                mbar_init(&empty[s], 1);
            }
            mbar_init(&tfull[0], 1);
            fence_barrier_init();
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(64));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < IMG_BUILD_THREADS / 32) {
        // ===================== builders: 128 threads per stage = (pixel, half of the taps); groups alternate stages ===============
        const int gq = warp >> 2, t = threadIdx.x & 127;
        constexpr int TAPS = KW * KW;
        constexpr int TPP = 128 / KP;                     // threads per pixel (2 at 64 pixels per stage)
        constexpr int MYT = (TAPS + TPP - 1) / TPP;       // taps per thread: sub, sub + TPP, ...
        const int px_i = t / TPP, sub = t - px_i * TPP;
        const int sw = px_i & 3;
        const uint32_t row_off = (uint32_t)((px_i >> 2) * 512 + (px_i & 3) * 128);
        for (int it = gq; it < nst; it += 2) {
            const int stage = it % p.stages;
            const uint32_t phase = (uint32_t)((it / p.stages) & 1);
            const long m = mbeg + (long)it * KP + px_i;
            const int img = (int)(m / pq);
            const int rem = (int)(m - (long)img * pq);
            const int oy = rem / p.Wo, ox = rem - oy * p.Wo;
            const float* xb = p.x + ((long)(g * p.xg_images + img) * p.H * p.W) * L;
            const int iy0 = oy * p.stride - p.pad, ix0 = ox * p.stride - p.pad;
            float4 v[MYT * (L / 4)];
#pragma unroll
            for (int j = 0; j < MYT; j++) {  // all loads of this thread's taps in flight before the first store
                const int tap = sub + j * TPP;
                const int iy = iy0 + tap / KW, ix = ix0 + tap % KW;
                const bool in = tap < TAPS && iy >= 0 && iy < p.H && ix >= 0 && ix < p.W;
                const float* px = xb + ((long)iy * p.W + ix) * L;
#pragma unroll
                for (int h = 0; h < L / 4; h++) v[j * (L / 4) + h] = in ? ldg4(px + 4 * h) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            mbar_wait(&empty[stage], phase ^ 1);
            uint8_t* sa = smem + (size_t)stage * stage_bytes;
            if (t == 0) {  // dy boxes of this stage; the expect_tx arrival comes last (below), so the phase cannot complete early
                const int row = (int)((long)g * p.Mpix + mbeg + (long)it * KP);
                tma_load_2d(&p.dymap, &full[stage], sa + 4 * box, 0, row);
                tma_load_2d(&p.dymap, &full[stage], sa + 5 * box, 32, row);
            }
#pragma unroll
            for (int j = 0; j < MYT; j++) {
                const int tap = sub + j * TPP;
                if (tap < TAPS) {
                    if (L == 8) {  // k = 8 tap: M group tap / 4, 32-byte chunk tap % 4
                        uint8_t* dst = sa + (tap >> 2) * box + row_off + (((tap & 3) ^ sw) << 5);
                        *reinterpret_cast<float4*>(dst) = to_tf32(v[j * (L / 4)]);
                        *reinterpret_cast<float4*>(dst + 16) = to_tf32(v[j * (L / 4) + (L / 4 - 1)]);
                    } else {       // k = 4 tap: M group tap / 8, 32-byte chunk (tap % 8) / 2, half tap % 2
                        *reinterpret_cast<float4*>(sa + (tap >> 3) * box + row_off + ((((tap & 7) >> 1) ^ sw) << 5) + (tap & 1) * 16) = to_tf32(v[j]);
                    }
                }
            }
            if (sub == TPP - 1) {  // row K of the patch matrix = 1: its accumulator row is the bias gradient
                const int kk = p.K & 31;
                *reinterpret_cast<float*>(sa + (p.K >> 5) * box + row_off + (((kk >> 3) ^ sw) << 5) + (kk & 7) * 4) = 1.0f;
            }
            fence_proxy_async();
            if (t == 0) mbar_expect_tx(&full[stage], (uint32_t)(2 * box));
            else mbar_arrive(&full[stage]);
        }
    } else if (warp == IMG_MMA_WARP) {
        if (lane == 0) {
            // kind::tf32, D = F32, A (patches) and B (dy) MN-major (bits 15, 16), M = 128, N = 64
            const uint32_t idesc = make_idesc_tf32(64) | (1u << 15) | (1u << 16);
            for (int it = 0; it < nst; it++) {
                const int stage = it % p.stages;
                const uint32_t phase = (uint32_t)((it / p.stages) & 1);
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
                const uint64_t mdesc = make_mnmajor_sw128_desc(sa, (uint32_t)box);
                const uint64_t ndesc = make_mnmajor_sw128_desc(sa + 4 * box, (uint32_t)box);
                const int nkk = p.kp / 8;
                for (int kk = 0; kk < nkk; kk++)  // next 8 pixels = next two 512-byte K atoms: +64 in the 16-byte address field
                    umma_tf32(tmem_base, mdesc + (uint64_t)(kk * 64), ndesc + (uint64_t)(kk * 64), idesc, (it | kk) ? 1u : 0u);
                umma_commit(&empty[stage]);
            }
            umma_commit(&tfull[0]);
        }
    } else {
        // ===================== epilogue: part[split][g][k][co], rows k <= K =====================
        const int quad = warp & 3;
        const int row = quad * 32 + lane;
        if (nst > 0) {
            mbar_wait(&tfull[0], 0);
            tc_fence_after();
        }
        float* op = p.part + (((long)split * p.G + g) * (p.K + 1) + row) * 64;
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16);
#pragma unroll
        for (int c0 = 0; c0 < 64; c0 += 32) {
            float v[32];
            if (nst > 0) {
                tmem_ld32(taddr + (uint32_t)c0, v);
            } else {
#pragma unroll
                for (int j = 0; j < 32; j++) v[j] = 0.f;
            }
            if (row <= p.K) {
#pragma unroll
                for (int j = 0; j < 8; j++) reinterpret_cast<float4*>(op + c0)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == IMG_MMA_WARP) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(64));
    }
}

// dw[g][co][k] = sum_split part[split][g][k][co];  db[g][co] = sum_split part[split][g][K][co]   (fixed order: deterministic)
__global__ void img_wgrad_reduce_kernel(const float* __restrict__ part, float* __restrict__ dw, float* __restrict__ db, int G, int K,
                                        int splits) {
    pdl_trigger();
    pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;  // over G * (K+1) * 64
    const int total = G * (K + 1) * 64;
    if (i >= total) return;
    const int co = i & 63;
    const int k = (i >> 6) % (K + 1);
    const int g = (i >> 6) / (K + 1);
    float s = 0.f;
    for (int sp = 0; sp < splits; sp++) s += __ldg(part + (((long)sp * G + g) * (K + 1) + k) * 64 + co);
    if (k < K) dw[((long)g * 64 + co) * K + k] = s;
    else if (db) db[g * 64 + co] = s;
}

// =====================================================================================================================
// host
// =====================================================================================================================
static bool img_geom_ok(const cg_conv_geom& g) {
    if (!g_img_path) return false;
    if (g.Cout != 64 || (g.Cin != 4 && g.Cin != 8) || g.ups || g.KH != g.KW) return false;
    if (g.stride != 1 && g.stride != 2) return false;
    if (!((g.Cin == 8 && g.KH == 3) || (g.Cin == 4 && g.KH == 4))) return false;  // the two instantiated layer types
    const int K = g.KH * g.KW * g.Cin;
    if (K > 96) return false;
    if (((long)g.B * g.Ho * g.Wo) % 128 != 0) return false;
    if (g.x_groups != 1 && g.x_groups != g.G) return false;
    const int sms = tc_sm_count();
    return sms > 0 && g.G <= sms;
}
bool img_fwd_supported(const cg_conv_geom& g, int act) { return img_geom_ok(g) && act != CG_ACT_TANH; }
bool img_wgrad_supported(const cg_conv_geom& g) { return img_geom_ok(g) && g.KH * g.KW * g.Cin + 1 <= 128; }

static void fill_common(ImgP& p, const cg_conv_geom& g) {
    p.G = g.G; p.xg_images = g.x_groups == 1 ? 0 : g.B; p.B = g.B; p.H = g.H; p.W = g.W; p.L = g.Cin;
    p.Ho = g.Ho; p.Wo = g.Wo; p.KH = g.KH; p.KW = g.KW; p.stride = g.stride; p.pad = g.pad;
    p.K = g.KH * g.KW * g.Cin;
    p.KC = (p.K + 31) / 32;
    p.k8 = (p.K + 7) / 8;
    p.cpg = tc_sm_count() / g.G;
    p.Mpix = (long)g.B * g.Ho * g.Wo;
}

int img_conv_fwd(const cg_conv_geom& g, const float* x, const float* w, const float* bias, float* y, int act, float slope, cudaStream_t st) {
    ImgP p{};
    fill_common(p, g);
    p.x = x; p.w = w; p.bias = bias; p.y = y; p.act = act; p.slope = slope;
    const int a_stage = p.KC * 16384;
    // output staging: two 32 KB tiles when at least three operand stages still fit, else one
    p.nsbuf = (226 * 1024 - p.KC * 8192 - 2048 - 2 * 32768) / a_stage >= 3 ? 2 : 1;
    int stages = (226 * 1024 - p.KC * 8192 - 2048 - p.nsbuf * 32768) / a_stage;
    if (stages > 6) stages = 6;
    p.stages = stages;
    const long tiles = p.Mpix / 128;
    if (p.cpg > tiles) p.cpg = (int)tiles;
    if (int rc = tc_encode_store_map(&p.dymap, y, (long)g.G * p.Mpix, 64, 128)) return rc;
    size_t smem = (size_t)stages * a_stage + (size_t)p.KC * 8192 + (size_t)p.nsbuf * 32768 + (2 * stages + 4) * 8 + 16 + 1024;
    auto kern = p.L == 8 ? img_conv_fwd_kernel<8, 3> : img_conv_fwd_kernel<4, 4>;
    static PerDeviceOnce attr_once[2];
    if (attr_once[p.L == 8].first()) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) {
            set_error("cudaFuncSetAttribute(img_conv_fwd_kernel): %s", cudaGetErrorString(e));
            return CG_ERR_CUDA;
        }
    }
    launch_k(kern, p.G * p.cpg, IMG_THREADS, smem, st, p);
    return check_launch("img_conv_fwd");
}

size_t img_wgrad_ws(const cg_conv_geom& g) {
    const int K = g.KH * g.KW * g.Cin;
    const int cpg = tc_sm_count() / (g.G > 0 ? g.G : 1);
    return (size_t)cpg * g.G * (K + 1) * 64 * sizeof(float);
}

int img_conv_wgrad(const cg_conv_geom& g, const float* x, const float* dy, float* dw, float* db, void* ws, size_t ws_bytes, cudaStream_t st) {
    ImgP p{};
    fill_common(p, g);
    size_t need = img_wgrad_ws(g);
    if (need > ws_bytes) {
        set_error("conv_wgrad(image path): workspace %zu < %zu bytes", ws_bytes, need);
        return CG_ERR_WORKSPACE;
    }
    p.kp = 64;  // img_wgrad_supported() guarantees Mpix % 128 == 0
    if (int rc = tc_encode_mn_map(&p.dymap, dy, (long)g.G * p.Mpix, 64, p.kp)) return rc;
    p.x = x; p.part = (float*)ws;
    const long nstage_total = p.Mpix / p.kp;
    if (p.cpg > nstage_total) p.cpg = (int)nstage_total;
    p.chunk = ((nstage_total + p.cpg - 1) / p.cpg) * p.kp;
    const int stage_bytes = 6 * p.kp * 128;
    int stages = (226 * 1024 - 2048) / stage_bytes;
    if (stages > 6) stages = 6;
    p.stages = stages;
    size_t smem = (size_t)stages * stage_bytes + (2 * stages + 2) * 8 + 16 + 1024;
    auto kern = p.L == 8 ? img_conv_wgrad_kernel<8, 3, 64> : img_conv_wgrad_kernel<4, 4, 64>;
    static PerDeviceOnce attr_once[2];
    if (attr_once[p.L == 8].first()) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) {
            set_error("cudaFuncSetAttribute(img_conv_wgrad_kernel): %s", cudaGetErrorString(e));
            return CG_ERR_CUDA;
        }
    }
    launch_k(kern, p.G * p.cpg, IMG_THREADS, smem, st, p);
    if (int rc = check_launch("img_conv_wgrad")) return rc;
    const int total = g.G * (p.K + 1) * 64;
    launch_k(img_wgrad_reduce_kernel, cdiv(total, 256), 256, 0, st, p.part, dw, db, g.G, p.K, p.cpg);
    return check_launch("img_wgrad_reduce");
}

}  // namespace cg
