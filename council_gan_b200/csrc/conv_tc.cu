// tcgen05 TF32 implicit-GEMM convolution for sm_100a.
//
//   D[128 pixels x BN couts] (fp32, TMEM) += A[128 x 32] (im2col tile of x, smem) * B[BN x 32] (weights, smem)
//
// * A tiles are gathered by TMA in IM2COL mode straight from the channels-last activation tensor
//   [G*B][H][W][Cin]: the zero padding of nn.ZeroPad2d (networks.py:473-474) is the TMA out-of-bound
//   fill, the stride is the TMA traversal stride -- there is no pad kernel and no im2col buffer.
// * B tiles are 2-D TMA boxes of the OHWI weight matrix [G*Cout][KH*KW*Cin] (K-major).
// * both land in shared memory in the 128-byte swizzled K-major layout tcgen05.mma reads; operands are
//   fp32 in HBM, converted to TF32 by the tensor map data type; accumulation is fp32 in tensor memory.
// * warp-specialised persistent CTAs (one per SM): warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM
//   allocator), warps 2..5 = epilogue (tcgen05.ld -> bias / activation / addend / mask -> global).
//   smem ring of STAGES (A,B) tiles with full/empty mbarriers; two TMEM accumulator stages so the epilogue
//   of tile i overlaps the mainloop of tile i+1.
// * all N council members are one launch: the member index is folded into the tensor-map coordinates
//   (image index g*B+n for A, row g*Cout+co for B), i.e. a grouped GEMM over the council.
//
// The same kernel serves the data gradient: the host passes dy as the activation, a transposed /
// flipped copy of the weights, and (for stride-2 layers) one launch "class" per output parity with its
// own im2col bounding box and a strided output mapping (see tc_conv_dgrad).
//
// Kernels in this file (host dispatch at the bottom of each section):
//   conv_tc_kernel<32|8>  forward / data gradient, one CTA per SM (tcgen05.mma.cta_group::1)
//   conv_tc2_kernel       the same for 128/256-wide tiles as CTA pairs (cta_group::2, M = 256 over two SMs)
//   wgrad_tc_kernel       weight gradient, cout on M, MN-major operands straight from the NHWC tensors
//   wgrad_xm_kernel       weight gradient with x on M for <= 64 output channels
//   wgrad_tc2_kernel      CTA-pair weight gradient (measured slower; behind a cg_set_tensor_core_mode switch)
//
// Reference call sites replaced: nn.Conv2d forward (networks.py:513,516) and cuDNN dgrad via autograd.
#include "common.cuh"
#include "tc_ptx.cuh"
#include <cstdlib>
#include <cstdio>
#include <cuda.h>
#include <mutex>
#include <string.h>

namespace cg {

// ------------------------------------------------------------------------------------------------
// driver entry points for tensor-map encoding (no link-time dependency on libcuda)
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode_tiled = nullptr;
static EncodeIm2colFn g_encode_im2col = nullptr;
static int g_sm_count_dev[CG_MAX_DEVICES] = {0};  // multiprocessors per device ordinal
static int sm_count_now() {
    int dev = current_device();
    if (!g_sm_count_dev[dev]) cudaDeviceGetAttribute(&g_sm_count_dev[dev], cudaDevAttrMultiProcessorCount, dev);
    return g_sm_count_dev[dev];
}
thread_local int g_pair_cap = 0;
thread_local int g_wgrad_xm = 1;  // x-on-M weight gradient for <= 64 output channels
thread_local int g_wgrad_2cta = 1;  // two co-resident weight-gradient CTAs per SM (run 44: -14..-33 % on the >= 128-channel layers)
thread_local int g_wgrad_xm2 = 1;   // ... also for the x-on-M kernel
thread_local int g_small_bn = 1;  // narrower N tiles when a launch has fewer tiles than SMs (mode bit 23 clears it)
thread_local int g_epi_coalesce = 1;  // epilogue stores through the per-warp patch (full sectors); bit 20 of the mode clears it
thread_local int g_fwd_2cta = 1;    // two co-resident forward / data-gradient CTAs per SM for tiles <= 64 channels wide (run 46: -1.4 ms/step)
thread_local int g_pair_mode = 3;  // bit 0: 256-wide tiles, bit 1: 128-wide, bit 2: 64-wide (measured slower than single CTAs: off),
                      // bit 3: weight gradient (MN-major operands: measured 15-20 % slower than single CTAs: off)  // cta_group::2 kernels for 256-wide layers (cg_set_tensor_core_mode bit 8 clears it for A/B runs)
static int g_driver_version = 0;
static std::once_flag g_once;

// ------------------------------------------------------------------------------------------------
// tensor-map cache ("lazily-created TMA descriptor cache keyed by (ptr, shape), guarded by a mutex"): a training
// step issues ~1500 cuTensorMapEncode* calls, almost all of them for (address, geometry) pairs seen in the previous step (the
// caching allocator hands the same blocks out again); a direct-mapped table of encoded descriptors replaces the driver call by a
// 48-byte key comparison and a 128-byte copy.  The descriptor depends on nothing but the key, so a stale entry is impossible.
// ------------------------------------------------------------------------------------------------
struct MapKey {
    const void* ptr;
    int64_t a, b;      // leading sizes (rows / images, ktot, ...)
    int32_t v[8];      // remaining geometry: sizes, box, corners, stride, kind
    bool operator==(const MapKey& o) const { return ptr == o.ptr && a == o.a && b == o.b && memcmp(v, o.v, sizeof(v)) == 0; }
};
struct MapSlot {
    MapKey key;
    CUtensorMap map;
    bool used;
};
constexpr int MAP_SLOTS = 8192;
static MapSlot* g_map_cache = nullptr;
static std::mutex g_map_mutex;
static std::atomic<uint64_t> g_map_hits{0}, g_map_misses{0};
static inline uint32_t map_hash(const MapKey& k) {
    uint64_t h = 1469598103934665603ull;
    const unsigned char* p = reinterpret_cast<const unsigned char*>(&k);
    for (size_t i = 0; i < sizeof(MapKey); i++) h = (h ^ p[i]) * 1099511628211ull;
    return (uint32_t)(h ^ (h >> 32)) & (MAP_SLOTS - 1);
}
// encode(map) is called on a miss; returns its status
template <class F>
static int cached_map(CUtensorMap* out, const MapKey& key, F encode) {
    const uint32_t slot = map_hash(key);
    {
        std::lock_guard<std::mutex> lock(g_map_mutex);
        if (!g_map_cache) g_map_cache = new MapSlot[MAP_SLOTS]();
        MapSlot& s = g_map_cache[slot];
        if (s.used && s.key == key) {
            *out = s.map;
            g_map_hits.fetch_add(1, std::memory_order_relaxed);
            return CG_OK;
        }
    }
    if (int rc = encode(out)) return rc;
    g_map_misses.fetch_add(1, std::memory_order_relaxed);
    std::lock_guard<std::mutex> lock(g_map_mutex);
    MapSlot& s = g_map_cache[slot];
    s.key = key;
    s.map = *out;
    s.used = true;
    return CG_OK;
}
void tc_map_cache_stats(uint64_t* hits, uint64_t* misses) {
    *hits = g_map_hits.load();
    *misses = g_map_misses.load();
}

static void init_driver() {
    std::call_once(g_once, [] {
        cudaDriverEntryPointQueryResult q;
        void* fn = nullptr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            g_encode_tiled = (EncodeTiledFn)fn;
        fn = nullptr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            g_encode_im2col = (EncodeIm2colFn)fn;
        cudaDriverGetVersion(&g_driver_version);
    });
}

// ------------------------------------------------------------------------------------------------
// kernel
// ------------------------------------------------------------------------------------------------
constexpr int TC_BM = 128;          // pixels per tile (TMEM lanes)
constexpr int TC_BK = 32;           // fp32 channels per stage = one 128-byte swizzle row
constexpr int TC_THREADS = 192;     // 6 warps: 0 TMA producer, 1 MMA issuer, 2..5 epilogue (TMEM lane quadrant = warp id % 4)
// Measured both ways: giving the two single-thread roles the HIGHEST warp ids (4, 5; epilogue 0..3) is ~4 % slower
// (102.2 vs 98.0 ms per step), so they keep ids 0 and 1.
constexpr int TC_PRODUCER_WARP = 0;
constexpr int TC_MMA_WARP = 1;
constexpr int TC_MAX_CLASSES = 4;   // stride-2 dgrad: one im2col map per output parity class

struct TcClass {
    CUtensorMap amap;     // im2col map of the activation for this class
    int w0, h0;           // base coordinate of output pixel (0,0): lower corner
    int out_h0, out_w0;   // output pixel (p,q) is written to (p*out_sh + out_h0, q*out_sw + out_w0)
    int wrow_off;         // row offset into the weight matrix for this class
    int pad_;
};
struct TcParams {
    CUtensorMap bmap;               // weights [rows][Ktot_class] 2-D
    const float* w_base;            // (host) what bmap was encoded over, so the CTA-pair launch can re-encode it with 128-row boxes
    long w_rows, w_ktot;
    TcClass cls[TC_MAX_CLASSES];
    int ncls;
    int G, xg_images;               // groups; images per group in the activation map (0: shared input)
    int B, P, Q;                    // output grid per class: P x Q pixels per image
    int Cin, Cout, KH, KW, stride;  // KH,KW: taps per class
    int bn;                         // N tile
    int bk;                         // K elements per stage: 32 (128-byte swizzled rows) or 8 (32-byte rows, Cin = 8)
    int cps;                        // K chunks per pipeline stage (small-N layers batch several: the MMA issue loop is latency bound)
    int n_store;                    // output channels actually stored per N tile (== bn except the 8-lane image gradient)
    int out_H, out_W, out_sh, out_sw;  // full output spatial size and class strides
    int w_rows_per_group;           // weight rows per group (ncls * Cout for dgrad classes)
    float* y; const float* bias; const float* addend; const float* mask_src;
    float* stats;                   // optional [chunks][G*B*Cout][2] partial (sum, sum of squares) of the raw output, or NULL
    long stats_gbc;                 // G*B*Cout
    int act; float slope;
    int stages;
    int coalesce;                   // epilogue leaves through the per-warp shared-memory patch with full-sector global accesses
};

// Epilogue store of one 32-row x 32-column accumulator block (lane = pixel row, v = its 32 channels, bias already added) with
// FULL-sector global accesses.  In the accumulator layout a warp-wide float4 store touches 32 different pixels, 16 bytes each:
// 32 half-written sectors per instruction, which capped the store-heavy layers near 2 TB/s (profiles/r02_summary.md).  Here the
// block goes through a 2 KB per-warp shared-memory patch, 16 columns at a time (64-byte rows, float4 index XOR-swizzled by
// (row >> 1) & 3: conflict-free for the row-wise writes and for the reads below), after which four consecutive lanes own 64
// contiguous bytes of one pixel -- for the stores and for the residual (addend) / activation-mask loads of the data gradient alike.
// out_off: element offset of this lane's pixel row at the block's first column, or -1 when the row is not stored.
__device__ __forceinline__ void epilogue_store_coalesced(const float (&v)[32], float4* patch, int lane, long out_off, float* __restrict__ y,
                                                         const float* __restrict__ addend, const float* __restrict__ mask_src, int act,
                                                         float slope) {
    const int wsw = (lane >> 1) & 3, c4 = lane & 3;
    long offs[4];
#pragma unroll
    for (int i = 0; i < 4; i++) offs[i] = __shfl_sync(0xffffffffu, out_off, i * 8 + (lane >> 2));
    // every residual / mask load of the block is issued before the shared-memory round trip (one exposed latency, not eight)
    float4 av[2][4], mv[2][4];
    if (addend) {
#pragma unroll
        for (int half = 0; half < 2; half++)
#pragma unroll
            for (int i = 0; i < 4; i++)
                av[half][i] = offs[i] >= 0 ? __ldg(reinterpret_cast<const float4*>(addend + offs[i] + half * 16 + c4 * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (mask_src) {
#pragma unroll
        for (int half = 0; half < 2; half++)
#pragma unroll
            for (int i = 0; i < 4; i++)
                mv[half][i] = offs[i] >= 0 ? __ldg(reinterpret_cast<const float4*>(mask_src + offs[i] + half * 16 + c4 * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int half = 0; half < 2; half++) {
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 4; j++)
            patch[lane * 4 + (j ^ wsw)] = make_float4(v[16 * half + 4 * j], v[16 * half + 4 * j + 1], v[16 * half + 4 * j + 2], v[16 * half + 4 * j + 3]);
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int r = i * 8 + (lane >> 2);
            float4 o = patch[r * 4 + (c4 ^ ((r >> 1) & 3))];
            if (offs[i] < 0) continue;
            if (addend) {
                const float4 a = av[half][i];
                o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
            }
            if (mask_src) {
                const float4 a = mv[half][i];
                o.x *= a.x > 0.f ? 1.f : slope; o.y *= a.y > 0.f ? 1.f : slope;
                o.z *= a.z > 0.f ? 1.f : slope; o.w *= a.w > 0.f ? 1.f : slope;
            } else if (act == CG_ACT_RELU) {
                o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
            } else if (act == CG_ACT_LRELU) {
                o.x = o.x > 0.f ? o.x : o.x * slope; o.y = o.y > 0.f ? o.y : o.y * slope;
                o.z = o.z > 0.f ? o.z : o.z * slope; o.w = o.w > 0.f ? o.w : o.w * slope;
            }
            *reinterpret_cast<float4*>(y + offs[i] + half * 16 + c4 * 4) = o;
        }
    }
}

template <int BK>
__global__ void __launch_bounds__(TC_THREADS, 2) conv_tc_kernel(const __grid_constant__ TcParams p) {
    pdl_trigger();
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int a_bytes = TC_BM * BK * 4;                       // one K chunk of A
    const int b_bytes = ((p.bn * BK * 4) + 1023) & ~1023;    // one K chunk of B (slot size)
    const int stage_bytes = p.cps * (a_bytes + b_bytes);       // [A_0..A_cps-1][B_0..B_cps-1]
    const int tx_chunk = a_bytes + p.bn * BK * 4;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
    uint64_t* empty_bar = full_bar + p.stages;
    uint64_t* tfull_bar = empty_bar + p.stages;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
    float* stat_smem = reinterpret_cast<float*>(tmem_slot + 4);  // 4 x (32 x 36) floats, only carved when p.stats
    float4* store_patch = reinterpret_cast<float4*>(stat_smem + (p.stats ? 4 * 32 * 36 : 0));  // 4 x 2 KB, only carved when p.coalesce

    const int MT = (p.B * p.P * p.Q + TC_BM - 1) / TC_BM;  // pixel tiles per (group, class)
    const int NT = (p.Cout + p.bn - 1) / p.bn;
    const int tiles = p.G * p.ncls * NT * MT;
    const int kchunks = (p.Cin + BK - 1) / BK;  // a partial last chunk reads zero-filled channels
    const int kiters = p.KH * p.KW * kchunks;
    const int tmem_cols = 2 * p.bn < 32 ? 32 : 2 * p.bn;

    if (warp == TC_PRODUCER_WARP && lane == 0) {
        prefetch_tmap(&p.bmap);
        for (int c = 0; c < p.ncls; c++) prefetch_tmap(&p.cls[c].amap);
    }
    if (warp == TC_MMA_WARP) {
        if (lane == 0) {
            for (int s = 0; s < p.stages; s++) {
                mbar_init(&full_bar[s], 1);
                mbar_init(&empty_bar[s], 1);
            }
            for (int a = 0; a < 2; a++) {
                mbar_init(&tfull_bar[a], 1);
                mbar_init(&tempty_bar[a], 128);
            }
            fence_barrier_init();
        }
        __syncwarp();
        // TMEM allocation (power of two >= 32 columns), whole warp
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();  // on-chip prologue done (barriers, TMEM): from here on the kernel reads what its predecessors in the stream wrote

    if (warp == TC_PRODUCER_WARP) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
                int mt = t % MT;
                int r = t / MT;
                int nt = r % NT;
                r /= NT;
                int c = r % p.ncls;
                int g = r / p.ncls;
                const TcClass& cl = p.cls[c];
                long m0 = (long)mt * TC_BM;
                int img = (int)(m0 / (p.P * p.Q));
                int rem = (int)(m0 - (long)img * p.P * p.Q);
                int pp = rem / p.Q, qq = rem - pp * p.Q;
                int n_coord = g * p.xg_images + img;
                int w_coord = cl.w0 + qq * p.stride;
                int h_coord = cl.h0 + pp * p.stride;
                int wrow = g * p.w_rows_per_group + cl.wrow_off + nt * p.bn;
                int kh = 0, kw = 0, kc = 0;
                for (int k0 = 0; k0 < kiters; k0 += p.cps) {
                    const int n = kiters - k0 < p.cps ? kiters - k0 : p.cps;
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + (size_t)stage * stage_bytes;
                    uint8_t* sb = sa + p.cps * a_bytes;
                    mbar_expect_tx(&full_bar[stage], (uint32_t)(n * tx_chunk));
                    for (int j = 0; j < n; j++) {
                        tma_load_im2col_4d(&cl.amap, &full_bar[stage], sa + j * a_bytes, kc * BK, w_coord, h_coord, n_coord, (uint16_t)kw,
                                           (uint16_t)kh);
                        tma_load_2d(&p.bmap, &full_bar[stage], sb + j * b_bytes, (kh * p.KW + kw) * p.Cin + kc * BK, wrow);
                        if (++kc == kchunks) { kc = 0; if (++kw == p.KW) { kw = 0; ++kh; } }
                    }
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == TC_MMA_WARP) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            const uint32_t idesc = make_idesc_tf32(p.bn);
            // descriptor without the start address (same for every chunk of this launch)
            const uint64_t desc_hi = (BK == 32 ? make_kmajor_sw128_desc(0) : make_kmajor_sw32_desc(0));
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
                mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.bn);
                for (int k0 = 0; k0 < kiters; k0 += p.cps) {
                    const int n = kiters - k0 < p.cps ? kiters - k0 : p.cps;
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
                    const uint32_t sb = sa + p.cps * a_bytes;
                    for (int j = 0; j < n; j++) {
                        const uint64_t adesc = desc_hi | (uint64_t)(((sa + j * a_bytes) & 0x3FFFF) >> 4);
                        const uint64_t bdesc = desc_hi | (uint64_t)(((sb + j * b_bytes) & 0x3FFFF) >> 4);
#pragma unroll
                        for (int kk = 0; kk < BK / 8; kk++) {
                            // advance 8 tf32 (32 bytes) along K inside the swizzled row: +2 in the 16-byte address field
                            umma_tf32(d_tmem, adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), idesc, (k0 | j | kk) != 0 ? 1u : 0u);
                        }
                    }
                    umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
                umma_commit(&tfull_bar[acc]);  // accumulator complete
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ===================== epilogue (warps 2..5 -> TMEM lane quadrants 2,3,0,1) =====================
        const int quad = warp & 3;
        const int row = quad * 32 + lane;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
            int mt = t % MT;
            int r = t / MT;
            int nt = r % NT;
            r /= NT;
            int c = r % p.ncls;
            int g = r / p.ncls;
            const TcClass& cl = p.cls[c];
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            const int m = mt * TC_BM + row;      // < 2^31: B*P*Q pixels per member
            const int pq = p.P * p.Q;
            bool valid = m < p.B * pq;
            long out_off = 0;
            if (valid) {
                int img = m / pq;
                int rem = m - img * pq;
                int pp = rem / p.Q, qq = rem - pp * p.Q;
                long pix = ((long)(g * p.B + img) * p.out_H + (pp * p.out_sh + cl.out_h0)) * p.out_W + (qq * p.out_sw + cl.out_w0);
                out_off = pix * p.Cout + nt * p.n_store;
            }
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * p.bn);
            if (p.n_store < 32) {
                // narrow outputs (image-lane gradients, the 12-channel head): 16 accumulator columns, n_store stored
                float v[32];
                tmem_ld16(taddr, v);
                if (valid) {
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        if (j >= p.n_store) break;
                        float4 o = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                        if (p.bias) {
                            const float* bp = p.bias + (long)g * p.Cout + j;
                            o.x += __ldg(bp); o.y += __ldg(bp + 1); o.z += __ldg(bp + 2); o.w += __ldg(bp + 3);
                        }
                        if (p.addend) {
                            float4 a = __ldg(reinterpret_cast<const float4*>(p.addend + out_off + j));
                            o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
                        }
                        if (p.mask_src) {
                            float4 a = __ldg(reinterpret_cast<const float4*>(p.mask_src + out_off + j));
                            o.x *= a.x > 0.f ? 1.f : p.slope; o.y *= a.y > 0.f ? 1.f : p.slope;
                            o.z *= a.z > 0.f ? 1.f : p.slope; o.w *= a.w > 0.f ? 1.f : p.slope;
                        } else if (p.act != CG_ACT_NONE) {
                            o.x = apply_act(o.x, p.act, p.slope); o.y = apply_act(o.y, p.act, p.slope);
                            o.z = apply_act(o.z, p.act, p.slope); o.w = apply_act(o.w, p.act, p.slope);
                        }
                        *reinterpret_cast<float4*>(p.y + out_off + j) = o;
                    }
                }
            } else
            for (int c0 = 0; c0 < p.bn; c0 += 32) {
                float v[32];
                tmem_ld32(taddr + (uint32_t)c0, v);
                if (p.stats) {
                    // instance-norm statistics of the raw convolution output, fused here instead of a second pass over y.
                    // Every tile lies inside one image (P*Q % 128 == 0) and this warp owns 32 of its rows: transpose the
                    // 32x32 block through a private shared-memory patch (row stride 36 floats: conflict-free both ways) and
                    // let lane j sum column j.
                    float* patch = stat_smem + quad * (32 * 36);
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 8; j++)
                        *reinterpret_cast<float4*>(patch + lane * 36 + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    __syncwarp();
                    float cs = 0.f, cq = 0.f;
#pragma unroll
                    for (int r = 0; r < 32; r++) {
                        float e = patch[r * 36 + lane];
                        cs += e;
                        cq = fmaf(e, e, cq);
                    }
                    const int tiles_per_img = pq / TC_BM;
                    const int img = (mt * TC_BM) / pq;
                    const int chunk = ((c * tiles_per_img + (mt - img * tiles_per_img)) << 2) + quad;
                    const long col = (long)(g * p.B + img) * p.Cout + nt * p.bn + c0 + lane;
                    *reinterpret_cast<float2*>(p.stats + ((long)chunk * p.stats_gbc + col) * 2) = make_float2(cs, cq);
                }
                if (p.coalesce) {
                    if (p.bias) {
                        const float4* bp = reinterpret_cast<const float4*>(p.bias + (long)g * p.Cout + nt * p.bn + c0);
#pragma unroll
                        for (int j = 0; j < 8; j++) {
                            float4 a = __ldg(bp + j);
                            v[4 * j] += a.x; v[4 * j + 1] += a.y; v[4 * j + 2] += a.z; v[4 * j + 3] += a.w;
                        }
                    }
                    epilogue_store_coalesced(v, store_patch + quad * 128, lane, valid ? out_off + c0 : -1, p.y, p.addend, p.mask_src, p.act, p.slope);
                } else
                if (valid) {
                    if (p.bias) {
                        const float4* bp = reinterpret_cast<const float4*>(p.bias + (long)g * p.Cout + nt * p.bn + c0);
#pragma unroll
                        for (int j = 0; j < 8; j++) {
                            float4 a = __ldg(bp + j);
                            v[4 * j] += a.x; v[4 * j + 1] += a.y; v[4 * j + 2] += a.z; v[4 * j + 3] += a.w;
                        }
                    }
                    if (p.addend) {
                        const float4* ap = reinterpret_cast<const float4*>(p.addend + out_off + c0);
#pragma unroll
                        for (int j = 0; j < 8; j++) {
                            float4 a = __ldg(ap + j);
                            v[4 * j] += a.x; v[4 * j + 1] += a.y; v[4 * j + 2] += a.z; v[4 * j + 3] += a.w;
                        }
                    }
                    if (p.mask_src) {
                        const float4* mp = reinterpret_cast<const float4*>(p.mask_src + out_off + c0);
#pragma unroll
                        for (int j = 0; j < 8; j++) {
                            float4 a = __ldg(mp + j);
                            v[4 * j] *= a.x > 0.f ? 1.f : p.slope; v[4 * j + 1] *= a.y > 0.f ? 1.f : p.slope;
                            v[4 * j + 2] *= a.z > 0.f ? 1.f : p.slope; v[4 * j + 3] *= a.w > 0.f ? 1.f : p.slope;
                        }
                    } else if (p.act == CG_ACT_RELU) {
#pragma unroll
                        for (int j = 0; j < 32; j++) v[j] = fmaxf(v[j], 0.f);
                    } else if (p.act == CG_ACT_LRELU) {
#pragma unroll
                        for (int j = 0; j < 32; j++) v[j] = v[j] > 0.f ? v[j] : v[j] * p.slope;
                    }  // tanh only occurs on the narrow (<= 16 channel) path; a rolled loop here would push v[] into local memory
                    float4* yp = reinterpret_cast<float4*>(p.y + out_off + c0);
#pragma unroll
                    for (int j = 0; j < 8; j++) yp[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                }
            }
            tc_fence_before();
            mbar_arrive(&tempty_bar[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == TC_MMA_WARP) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols));
    }
}

// ------------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2) for the 256-wide layers
//
// One MMA instruction spans two SMs: D[256 px x 256 co]; each CTA of the pair stages its own 128 pixel rows of A
// and HALF of the weight tile (128 of the 256 N rows), so the shared-memory -> tensor-core operand traffic per SM
// drops from 12 KB to 8 KB per 128x256x8 step -- the binding resource of the single-CTA kernel (ncu:
// sm__mem_tensor_cycles_active 72 %).  The leader CTA (cluster rank 0) issues the MMAs; both CTAs issue TMA into
// their own shared memory but signal the leader's "full" barrier; tcgen05.commit multicasts "stage free" /
// "accumulator ready" to both CTAs; both epilogues arrive on the leader's "accumulator drained" barrier.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local` in the leader CTA (rank 0) of the pair
__device__ __forceinline__ uint32_t leader_addr(uint32_t local) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(0));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    // default semantics (release at CTA scope): the .release.cluster form costs a MEMBAR.ALL.GPU per arrival
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma2_load_2d(const CUtensorMap* map, uint32_t bar_cluster, void* dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma2_load_im2col_4d(const CUtensorMap* map, uint32_t bar_cluster, void* dst, int c, int w, int h, int n,
                                                    uint16_t off_w, uint16_t off_h) {
    asm volatile(
        "cp.async.bulk.tensor.4d.im2col.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2], "
        "{%7, %8};" ::"r"(smem_u32(dst)),
        "l"(map), "r"(bar_cluster), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h)
        : "memory");
}
__device__ __forceinline__ void umma2_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0)
        : "memory");
}
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar) {  // arrive on `bar` in BOTH CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
}

constexpr int TC2_MAX_STAGES = 12;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1) conv_tc2_kernel(const __grid_constant__ TcParams p) {
    pdl_trigger();
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    constexpr int BK = TC_BK;
    constexpr int a_bytes = TC_BM * BK * 4;      // 16 KB: one K chunk of this CTA's 128 pixel rows
    const int b_bytes = (p.bn / 2) * BK * 4;     // 4 / 8 / 16 KB: one K chunk of this CTA's half of the bn weight rows
    const int stage_bytes = p.cps * (a_bytes + b_bytes);  // [A_0..A_cps-1][B_0..B_cps-1]
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
    uint64_t* empty_bar = full_bar + p.stages;
    uint64_t* tfull_bar = empty_bar + p.stages;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
    float* stat_smem = reinterpret_cast<float*>(tmem_slot + 4);  // 4 x (32 x 36) floats, only carved when p.stats
    float4* store_patch = reinterpret_cast<float4*>(stat_smem + (p.stats ? 4 * 32 * 36 : 0));  // 4 x 2 KB, only carved when p.coalesce

    const int pq = p.P * p.Q;
    const int PT = (p.B * pq + 2 * TC_BM - 1) / (2 * TC_BM);  // pair tiles (256 pixels) per (group, class)
    const int NT = p.Cout / p.bn;
    const int tiles = p.G * p.ncls * NT * PT;
    const int kchunks = p.Cin / BK;
    const int kiters = p.KH * p.KW * kchunks;
    const int cluster_id = blockIdx.x >> 1, nclusters = gridDim.x >> 1;

    if (warp == TC_PRODUCER_WARP && lane == 0) {
        prefetch_tmap(&p.bmap);
        for (int c = 0; c < p.ncls; c++) prefetch_tmap(&p.cls[c].amap);
    }
    if (warp == TC_MMA_WARP) {
        if (lane == 0) {
            for (int s = 0; s < p.stages; s++) {
                mbar_init(&full_bar[s], 1);   // leader's arrive.expect_tx covers the bytes of BOTH CTAs (the peer only issues TMA)
                mbar_init(&empty_bar[s], 1);  // multicast commit
            }
            for (int a = 0; a < 2; a++) {
                mbar_init(&tfull_bar[a], 1);     // multicast commit
                mbar_init(&tempty_bar[a], 256);  // leader: 128 epilogue threads of each CTA
            }
            fence_barrier_init();
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();  // on-chip prologue done (barriers, TMEM): from here on the kernel reads what its predecessors in the stream wrote

    if (warp == TC_PRODUCER_WARP) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int t = cluster_id; t < tiles; t += nclusters) {
                int pt = t % PT;
                int r = t / PT;
                int nt = r % NT;
                r /= NT;
                int c = r % p.ncls;
                int g = r / p.ncls;
                const TcClass& cl = p.cls[c];
                int m0 = pt * 2 * TC_BM + (int)rank * TC_BM;
                int img = m0 / pq;
                int rem = m0 - img * pq;
                int pp = rem / p.Q, qq = rem - pp * p.Q;
                int n_coord = g * p.xg_images + img;
                int w_coord = cl.w0 + qq * p.stride;
                int h_coord = cl.h0 + pp * p.stride;
                int wrow = g * p.w_rows_per_group + cl.wrow_off + nt * p.bn + (int)rank * (p.bn >> 1);
                int kh = 0, kw = 0, kc = 0;
                for (int k0 = 0; k0 < kiters; k0 += p.cps) {
                    const int n = kiters - k0 < p.cps ? kiters - k0 : p.cps;
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + (size_t)stage * stage_bytes;
                    uint8_t* sb = sa + p.cps * a_bytes;
                    const uint32_t full_leader = leader_addr(smem_u32(&full_bar[stage]));
                    // both CTAs' A + half-B land on the leader's barrier; a peer box that lands before this expect_tx only drives
                    // the transaction count negative for a moment (the phase cannot complete before the leader's arrival)
                    if (rank == 0) mbar_expect_tx(&full_bar[stage], (uint32_t)(2 * n * (a_bytes + b_bytes)));
                    for (int j = 0; j < n; j++) {
                        tma2_load_im2col_4d(&cl.amap, full_leader, sa + j * a_bytes, kc * BK, w_coord, h_coord, n_coord, (uint16_t)kw,
                                            (uint16_t)kh);
                        tma2_load_2d(&p.bmap, full_leader, sb + j * b_bytes, (kh * p.KW + kw) * p.Cin + kc * BK, wrow);
                        if (++kc == kchunks) { kc = 0; if (++kw == p.KW) { kw = 0; ++kh; } }
                    }
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == TC_MMA_WARP) {
        if (lane == 0 && rank == 0) {
            // kind::tf32, D=F32, K-major A and B, M = 256 (both CTAs), N = bn
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.bn >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
            const uint64_t desc_hi = make_kmajor_sw128_desc(0);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int t = cluster_id; t < tiles; t += nclusters) {
                mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.bn);
                for (int k0 = 0; k0 < kiters; k0 += p.cps) {
                    const int n = kiters - k0 < p.cps ? kiters - k0 : p.cps;
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
                    const uint32_t sb = sa + p.cps * a_bytes;
                    for (int j = 0; j < n; j++) {
                        const uint64_t adesc = desc_hi | (uint64_t)(((sa + j * a_bytes) & 0x3FFFF) >> 4);
                        const uint64_t bdesc = desc_hi | (uint64_t)(((sb + j * b_bytes) & 0x3FFFF) >> 4);
#pragma unroll
                        for (int kk = 0; kk < BK / 8; kk++)
                            umma2_tf32(d_tmem, adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), idesc, (k0 | j | kk) != 0 ? 1u : 0u);
                    }
                    umma2_commit_mc(&empty_bar[stage]);
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
                umma2_commit_mc(&tfull_bar[acc]);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        const int quad = warp & 3;
        const int row = quad * 32 + lane;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int t = cluster_id; t < tiles; t += nclusters) {
            int pt = t % PT;
            int r = t / PT;
            int nt = r % NT;
            r /= NT;
            int c = r % p.ncls;
            int g = r / p.ncls;
            const TcClass& cl = p.cls[c];
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            const int m = pt * 2 * TC_BM + (int)rank * TC_BM + row;
            bool valid = m < p.B * pq;
            long out_off = 0;
            if (valid) {
                int img = m / pq;
                int rem = m - img * pq;
                int pp = rem / p.Q, qq = rem - pp * p.Q;
                long pix = ((long)(g * p.B + img) * p.out_H + (pp * p.out_sh + cl.out_h0)) * p.out_W + (qq * p.out_sw + cl.out_w0);
                out_off = pix * p.Cout + nt * p.bn;
            }
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * p.bn);
            for (int c0 = 0; c0 < p.bn; c0 += 32) {
                float v[32];
                tmem_ld32(taddr + (uint32_t)c0, v);
                if (p.stats) {
                    // instance-norm statistics of the raw output in the epilogue (Conv2d -> InstanceNorm2d / AdaIN, networks.py:516-518):
                    // this CTA's 128 rows are one 128-pixel tile inside one image (P*Q % 256 == 0); same scheme as conv_tc_kernel.
                    // On the wide layers served by this kernel the main loop of the next tile (tens of thousands of cycles) hides it.
                    float* patch = stat_smem + quad * (32 * 36);
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 8; j++)
                        *reinterpret_cast<float4*>(patch + lane * 36 + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    __syncwarp();
                    float cs = 0.f, cq = 0.f;
#pragma unroll
                    for (int r2 = 0; r2 < 32; r2++) {
                        float e = patch[r2 * 36 + lane];
                        cs += e;
                        cq = fmaf(e, e, cq);
                    }
                    const int tiles_per_img = pq / TC_BM;
                    const int mt = pt * 2 + (int)rank;
                    const int img = (mt * TC_BM) / pq;
                    const int chunk = ((c * tiles_per_img + (mt - img * tiles_per_img)) << 2) + quad;
                    const long col = (long)(g * p.B + img) * p.Cout + nt * p.bn + c0 + lane;
                    if (mt * TC_BM < p.B * pq) *reinterpret_cast<float2*>(p.stats + ((long)chunk * p.stats_gbc + col) * 2) = make_float2(cs, cq);
                }
                if (p.coalesce) {
                    if (p.bias) {
                        const float4* bp = reinterpret_cast<const float4*>(p.bias + (long)g * p.Cout + nt * p.bn + c0);
#pragma unroll
                        for (int j = 0; j < 8; j++) {
                            float4 a = __ldg(bp + j);
                            v[4 * j] += a.x; v[4 * j + 1] += a.y; v[4 * j + 2] += a.z; v[4 * j + 3] += a.w;
                        }
                    }
                    epilogue_store_coalesced(v, store_patch + quad * 128, lane, valid ? out_off + c0 : -1, p.y, p.addend, p.mask_src, p.act, p.slope);
                } else
                if (valid) {
                    if (p.bias) {
                        const float4* bp = reinterpret_cast<const float4*>(p.bias + (long)g * p.Cout + nt * p.bn + c0);
#pragma unroll
                        for (int j = 0; j < 8; j++) {
                            float4 a = __ldg(bp + j);
                            v[4 * j] += a.x; v[4 * j + 1] += a.y; v[4 * j + 2] += a.z; v[4 * j + 3] += a.w;
                        }
                    }
                    if (p.addend) {
                        const float4* ap = reinterpret_cast<const float4*>(p.addend + out_off + c0);
#pragma unroll
                        for (int j = 0; j < 8; j++) {
                            float4 a = __ldg(ap + j);
                            v[4 * j] += a.x; v[4 * j + 1] += a.y; v[4 * j + 2] += a.z; v[4 * j + 3] += a.w;
                        }
                    }
                    if (p.mask_src) {
                        const float4* mp = reinterpret_cast<const float4*>(p.mask_src + out_off + c0);
#pragma unroll
                        for (int j = 0; j < 8; j++) {
                            float4 a = __ldg(mp + j);
                            v[4 * j] *= a.x > 0.f ? 1.f : p.slope; v[4 * j + 1] *= a.y > 0.f ? 1.f : p.slope;
                            v[4 * j + 2] *= a.z > 0.f ? 1.f : p.slope; v[4 * j + 3] *= a.w > 0.f ? 1.f : p.slope;
                        }
                    } else if (p.act == CG_ACT_RELU) {
#pragma unroll
                        for (int j = 0; j < 32; j++) v[j] = fmaxf(v[j], 0.f);
                    } else if (p.act == CG_ACT_LRELU) {
#pragma unroll
                        for (int j = 0; j < 32; j++) v[j] = v[j] > 0.f ? v[j] : v[j] * p.slope;
                    }
                    float4* yp = reinterpret_cast<float4*>(p.y + out_off + c0);
#pragma unroll
                    for (int j = 0; j < 8; j++) yp[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                }
            }
            tc_fence_before();
            mbar_arrive_cluster(leader_addr(smem_u32(&tempty_bar[acc])));
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == TC_MMA_WARP) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int encode_weights_map_raw(CUtensorMap* map, const float* w, long rows, long ktot, int bn, int bk) {
    cuuint64_t dims[2] = {(cuuint64_t)ktot, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ktot * 4};
    cuuint32_t box[2] = {(cuuint32_t)bk, (cuuint32_t)bn};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode_tiled(map, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 2, (void*)w, dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, bk == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_32B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(weights rows=%ld ktot=%ld bn=%d) failed: %d", rows, ktot, bn, (int)r);
        return CG_ERR_CUDA;
    }
    return CG_OK;
}

static int encode_weights_map(CUtensorMap* map, const float* w, long rows, long ktot, int bn, int bk = TC_BK) {
    MapKey k{};
    k.ptr = w; k.a = rows; k.b = ktot; k.v[0] = bn; k.v[1] = bk; k.v[7] = 1;
    return cached_map(map, k, [&](CUtensorMap* m) { return encode_weights_map_raw(m, w, rows, ktot, bn, bk); });
}

static int encode_weights_map_p(TcParams& p, const float* w, long rows, long ktot) {
    p.w_base = w;
    p.w_rows = rows;
    p.w_ktot = ktot;
    return encode_weights_map(&p.bmap, w, rows, ktot, p.bn, p.bk);
}

// activation [N][H][W][C] viewed by TMA as (C, W, H, N); bounding box corners as in CUTLASS
// (cutlass/conv/collective/detail.hpp compute_lower/upper_corner_whd): lower = -pad_lo, upper = pad_hi - (K-1).
static int encode_act_map_raw(CUtensorMap* map, const float* x, long N, int H, int W, int C, int lo_w, int lo_h, int up_w, int up_h,
                              int stride, int bk) {
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
    int lower[2] = {lo_w, lo_h};
    int upper[2] = {up_w, up_h};
    cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
    CUresult r = g_encode_im2col(map, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 4, (void*)x, dims, strides, lower, upper, bk, TC_BM, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, bk == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_32B,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeIm2col(N=%ld H=%d W=%d C=%d corners %d,%d / %d,%d stride %d) failed: %d", N, H, W, C, lo_w, lo_h,
                  up_w, up_h, stride, (int)r);
        return CG_ERR_CUDA;
    }
    // driver <= 13.1: small tensors need bit 21 of the second descriptor word cleared (same workaround as CUTLASS
    // cute/atom/copy_traits_sm90_im2col.hpp:477-483)
    if (g_driver_version <= 13010 && (long)N * H * W * C * 4 < 131072) reinterpret_cast<uint64_t*>(map)[1] &= ~(1ull << 21);
    return CG_OK;
}

static int encode_act_map(CUtensorMap* map, const float* x, long N, int H, int W, int C, int lo_w, int lo_h, int up_w, int up_h,
                          int stride, int bk = TC_BK) {
    MapKey k{};
    k.ptr = x; k.a = N; k.b = ((int64_t)H << 32) | (uint32_t)W;
    k.v[0] = C; k.v[1] = lo_w; k.v[2] = lo_h; k.v[3] = up_w; k.v[4] = up_h; k.v[5] = stride; k.v[6] = bk; k.v[7] = 2;
    return cached_map(map, k, [&](CUtensorMap* m) { return encode_act_map_raw(m, x, N, H, W, C, lo_w, lo_h, up_w, up_h, stride, bk); });
}

static int pick_bn(int cout) {
    if (cout % 256 == 0) return 256;
    if (cout == 128) return 128;
    if (cout == 64) return 64;
    if (cout <= 16 && cout % 4 == 0) return 16;  // narrow outputs: 16 accumulator columns, `cout` stored
    return 0;
}
// K elements per pipeline stage: 32-channel (128-byte) rows, or 8-channel (32-byte) rows for <= 16 channels
static int pick_bk(int cin) {
    if (cin % TC_BK == 0) return TC_BK;
    if (cin <= 16 && cin % 4 == 0) return 8;
    return 0;
}

// Few-tile launches (small maps: the 32x32 bottleneck of the 128x128 configuration is 8 pixel tiles per member): narrower N tiles put more
// SMs to work.  Every extra N tile re-reads the same activation tile, from L2, which is cheap at these sizes; a 3x3 256->256 layer on
// 1024 pixels x 2 members goes from 8 CTA pairs running 288 K-steps of 256-wide MMAs to 64 CTAs running the same steps 64 wide.
static int fill_bn(int bn, long mpix, int groups_classes, int cout) {
    if (!g_small_bn) return bn;
    const int sms = sm_count_now();
    while (bn > 64 && cout % (bn / 2) == 0 && (long)groups_classes * cdiv(mpix, TC_BM) * (cout / bn) < sms) bn >>= 1;
    return bn;
}

bool tc_fwd_supported(const cg_conv_geom& g) {
    init_driver();
    if (!g_encode_tiled || !g_encode_im2col) return false;
    if (g.ups && !(g.KH == 3 && g.KW == 3 && g.stride == 1 && g.pad == 1 && g.Cin % TC_BK == 0)) return false;
    if (pick_bk(g.Cin) == 0) return false;
    if (pick_bn(g.Cout) == 0) return false;
    if (g.pad > 120 || g.KH > 120) return false;
    if ((long)g.B * (g.ups ? g.H * g.W : g.Ho * g.Wo) < TC_BM) return false;  // tiny maps: the SIMT kernel is fine
    return true;
}

static int launch_tc(TcParams& p, cudaStream_t st) {
    int a_bytes = TC_BM * p.bk * 4, b_bytes = ((p.bn * p.bk * 4) + 1023) & ~1023;
    int chunk_bytes = a_bytes + b_bytes;
    // several K chunks per stage when a chunk carries little tensor work (N <= 64 or 32-byte rows): the single MMA-issuing
    // thread pays ~300 cycles of barrier / descriptor latency per stage
    int kiters = p.KH * p.KW * ((p.Cin + p.bk - 1) / p.bk);
    int cps = p.bn <= 128 ? 2 : 1;   // 48 KB (N=64) / 64 KB (N=128) stages, 4 / 3 deep
    if (p.bk == 8) cps = 8;          // 6 KB chunks
    while (cps > 1 && (cps > kiters || cps * chunk_bytes > 64 * 1024)) cps >>= 1;
    p.cps = cps;
    int stage_bytes = cps * chunk_bytes;
    // narrow tiles (<= 64 output channels) are the store-heavy ones; on the wide compute-bound tiles the patch traffic competes with the
    // MMA operand reads for shared-memory bandwidth and measured 1-3 % slower (visit I): those keep the accumulator-layout stores
    p.coalesce = g_epi_coalesce && p.bn >= 32 && (p.bn <= 64 || g_epi_coalesce == 2) && p.act != CG_ACT_TANH ? 1 : 0;
    const int stat_bytes = (p.stats ? 4 * 32 * 36 * 4 : 0) + (p.coalesce ? 4 * 2048 : 0);
    // narrow tiles are bound by the single MMA-issuing thread's per-stage latency: two co-resident CTAs per SM interleave their MMAs
    const bool two = g_fwd_2cta && p.bn <= 64 && !p.stats && 2 * stage_bytes <= 100 * 1024;
    int stages = ((two ? 112 : 226) * 1024 - 1024 - 512 - stat_bytes) / stage_bytes;  // two co-resident CTAs: 2 x (112 KB + 1 KB reserved) <= 228 KB  // 227 KB dynamic shared memory per SM
    if (stages > 12) stages = 12;
    if (stages > 4 && stage_bytes >= 48 * 1024) stages = 4;
    if (p.n_store == 0) p.n_store = p.bn;
    p.stages = stages;
    size_t smem = (size_t)stages * stage_bytes + 1024 /*align slack*/ + (2 * stages + 4) * 8 + 32 + stat_bytes;
    static PerDeviceOnce attr_set;
    if (attr_set.first()) {
        cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_tc_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) {
            set_error("cudaFuncSetAttribute(conv_tc_kernel): %s", cudaGetErrorString(e));
            return CG_ERR_CUDA;
        }
    }
    int MT = cdiv((long)p.B * p.P * p.Q, TC_BM);
    long tiles = (long)p.G * p.ncls * ((p.Cout + p.bn - 1) / p.bn) * MT;
    // CTA pairs (cta_group::2): one MMA spans two SMs, each staging its own 128 pixel rows and HALF of the weight rows -> less
    // shared-memory operand traffic per SM (wide layers) and half the MMA instructions per pixel (narrow layers, issue-bound)
    if (((p.bn == 256 && (g_pair_mode & 1)) || (p.bn == 128 && (g_pair_mode & 2)) || (p.bn == 64 && (g_pair_mode & 4))) && p.bk == 32 && p.Cout % p.bn == 0 && p.Cin % 32 == 0 &&
        (!p.stats || (p.P * p.Q) % (2 * TC_BM) == 0) && p.act != CG_ACT_TANH && (long)p.B * p.P * p.Q >= 512) {
        static PerDeviceOnce attr2_set;
        if (attr2_set.first()) {
            cudaError_t e = cudaFuncSetAttribute(conv_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
            if (e != cudaSuccess) {
                set_error("cudaFuncSetAttribute(conv_tc2_kernel): %s", cudaGetErrorString(e));
                return CG_ERR_CUDA;
            }
        }
        const int stage2 = p.cps * (TC_BM * TC_BK * 4 + (p.bn / 2) * TC_BK * 4);
        int stages2 = (226 * 1024 - 1536 - stat_bytes) / stage2;
        if (stages2 > TC2_MAX_STAGES) stages2 = TC2_MAX_STAGES;
        p.stages = stages2;
        size_t smem2 = (size_t)stages2 * stage2 + 1024 + (2 * TC2_MAX_STAGES + 4) * 8 + 32 + stat_bytes;
        static int max_pairs_dev[CG_MAX_DEVICES] = {0};  // co-resident CTA pairs (GPCs with an odd SM count strand one SM each)
        int& max_pairs = max_pairs_dev[current_device()];
        if (!max_pairs) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(2 * (sm_count_now() / 2));
            cfg.blockDim = dim3(TC_THREADS);
            cfg.dynamicSmemBytes = 220 * 1024;
            cudaLaunchAttribute at;
            at.id = cudaLaunchAttributeClusterDimension;
            at.val.clusterDim.x = 2;
            at.val.clusterDim.y = 1;
            at.val.clusterDim.z = 1;
            cfg.attrs = &at;
            cfg.numAttrs = 1;
            int n = 0;
            cudaError_t e = cudaOccupancyMaxActiveClusters(&n, conv_tc2_kernel, &cfg);
            if (e != cudaSuccess || n < 1) {
                (void)cudaGetLastError();
                n = sm_count_now() / 2 - 4;
            }
            max_pairs = n < sm_count_now() / 2 ? n : sm_count_now() / 2;
        }
        int pairs = g_pair_cap > 0 && g_pair_cap < max_pairs ? g_pair_cap : max_pairs;
        if (int rc = encode_weights_map(&p.bmap, p.w_base, p.w_rows, p.w_ktot, p.bn / 2, 32)) return rc;  // each CTA stages half of the rows
        long ptiles = (long)p.G * p.ncls * (p.Cout / p.bn) * cdiv((long)p.B * p.P * p.Q, 2 * TC_BM);
        int nclusters = (int)(ptiles < pairs ? ptiles : pairs);
        launch_k(conv_tc2_kernel, 2 * nclusters, TC_THREADS, smem2, st, p);
        return check_launch("conv_tc2_kernel");
    }
    const long slots = (long)sm_count_now() * (two ? 2 : 1);
    int grid = (int)(tiles < slots ? tiles : slots);
    if (p.bk == 32) launch_k(conv_tc_kernel<32>, grid, TC_THREADS, smem, st, p);
    else launch_k(conv_tc_kernel<8>, grid, TC_THREADS, smem, st, p);
    return check_launch("conv_tc_kernel");
}

// nearest-upsample x2 followed by a 3x3 pad-1 convolution == four 2x2 convolutions on the ORIGINAL tensor, one per
// output parity (a,b): output row 2i+a reads source rows {i-1,i} (a=0) or {i,i+1} (a=1), and the 3 filter rows
// collapse onto those 2 source rows: a=0 -> {w0, w1+w2}, a=1 -> {w0+w1, w2} (same along columns).  2.25x fewer
// FLOPs than convolving the materialised upsampled tensor, and the upsampled tensor never exists.
__global__ void ups_weight_transform_kernel(const float* __restrict__ w, float* __restrict__ wc, long total, int Cout, int Cin) {
    pdl_trigger();
    pdl_wait();
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;  // wc[g][cls][co][r][s][ci]
    if (i >= total) return;
    int ci = (int)(i % Cin);
    long t = i / Cin;
    int ss = (int)(t % 2); t /= 2;
    int r = (int)(t % 2); t /= 2;
    int co = (int)(t % Cout); t /= Cout;
    int cls = (int)(t % 4);
    int g = (int)(t / 4);
    int a = cls >> 1, b = cls & 1;
    // taps of the 3-wide filter that land on source offset r for parity a
    int kh0 = a == 0 ? (r == 0 ? 0 : 1) : (r == 0 ? 0 : 2), kh1 = a == 0 ? (r == 0 ? 0 : 2) : (r == 0 ? 1 : 2);
    int kw0 = b == 0 ? (ss == 0 ? 0 : 1) : (ss == 0 ? 0 : 2), kw1 = b == 0 ? (ss == 0 ? 0 : 2) : (ss == 0 ? 1 : 2);
    const float* wp = w + ((long)g * Cout + co) * 9 * Cin + ci;
    float acc = 0.f;
    for (int kh = kh0; kh <= kh1; kh++)
        for (int kw = kw0; kw <= kw1; kw++) acc += __ldg(wp + (kh * 3 + kw) * Cin);
    wc[i] = acc;
}

size_t tc_fwd_ws(const cg_conv_geom& g) {
    return g.ups ? (size_t)g.G * 4 * g.Cout * 4 * g.Cin * sizeof(float) : 0;
}

// number of partial-statistics chunks the fused epilogue writes per (image, channel); 0: fusion not possible
int tc_fwd_stats_chunks(const cg_conv_geom& g) {
    if (pick_bn(g.Cout) < 32) return 0;
    long pq = g.ups ? (long)g.H * g.W : (long)g.Ho * g.Wo;
    if (pq % TC_BM != 0) return 0;
    return (int)((g.ups ? 4 : 1) * (pq / TC_BM) * 4);
}

int tc_conv_fwd(const cg_conv_geom& g, const float* x, const float* w, const float* bias, float* y, int act, float slope, void* ws,
                size_t ws_bytes, cudaStream_t st, float* stats_part) {
    init_driver();
    TcParams p{};
    p.stats = stats_part;
    p.stats_gbc = (long)g.G * g.B * g.Cout;
    p.bn = pick_bn(g.Cout);
    if (p.bn > 64) p.bn = fill_bn(p.bn, (long)g.B * (g.ups ? g.H * g.W : g.Ho * g.Wo), g.G * (g.ups ? 4 : 1), g.Cout);
    p.n_store = p.bn == 16 ? g.Cout : p.bn;
    p.bk = pick_bk(g.Cin);
    long nimg = (long)(g.x_groups == 1 ? 1 : g.G) * g.B;
    p.G = g.G; p.xg_images = g.x_groups == 1 ? 0 : g.B;
    p.B = g.B; p.Cin = g.Cin; p.Cout = g.Cout;
    p.y = y; p.bias = bias; p.addend = nullptr; p.mask_src = nullptr; p.act = act; p.slope = slope;
    if (g.ups) {
        size_t need = tc_fwd_ws(g);
        if (need > ws_bytes) {
            set_error("conv_fwd(tc, upsample classes): workspace %zu < %zu bytes", ws_bytes, need);
            return CG_ERR_WORKSPACE;
        }
        float* wc = (float*)ws;
        long total = (long)g.G * 4 * g.Cout * 4 * g.Cin;
        launch_k(ups_weight_transform_kernel, cdiv(total, 256), 256, 0, st, w, wc, total, g.Cout, g.Cin);
        if (int rc = check_launch("ups_weight_transform")) return rc;
        if (int rc = encode_weights_map_p(p, wc, (long)g.G * 4 * g.Cout, 4L * g.Cin)) return rc;
        for (int c = 0; c < 4; c++) {
            int a = c >> 1, b = c & 1;
            int lo_h = a == 0 ? -1 : 0, lo_w = b == 0 ? -1 : 0;  // lower corner = -(left pad); upper = right pad - (K-1), K = 2
            int up_h = a == 0 ? -1 : 0, up_w = b == 0 ? -1 : 0;
            if (int rc = encode_act_map(&p.cls[c].amap, x, nimg, g.H, g.W, g.Cin, lo_w, lo_h, up_w, up_h, 1, p.bk)) return rc;
            p.cls[c].w0 = lo_w; p.cls[c].h0 = lo_h; p.cls[c].out_h0 = a; p.cls[c].out_w0 = b; p.cls[c].wrow_off = c * g.Cout;
        }
        p.ncls = 4;
        p.P = g.H; p.Q = g.W; p.KH = 2; p.KW = 2; p.stride = 1;
        p.out_H = g.Ho; p.out_W = g.Wo; p.out_sh = 2; p.out_sw = 2;
        p.w_rows_per_group = 4 * g.Cout;
        return launch_tc(p, st);
    }
    long ktot = (long)g.KH * g.KW * g.Cin;
    if (int rc = encode_weights_map_p(p, w, (long)g.G * g.Cout, ktot)) return rc;
    if (int rc = encode_act_map(&p.cls[0].amap, x, nimg, g.H, g.W, g.Cin, -g.pad, -g.pad, g.pad - (g.KW - 1), g.pad - (g.KH - 1), g.stride,
                                p.bk))
        return rc;
    p.cls[0].w0 = -g.pad; p.cls[0].h0 = -g.pad; p.cls[0].out_h0 = 0; p.cls[0].out_w0 = 0; p.cls[0].wrow_off = 0;
    p.ncls = 1;
    p.P = g.Ho; p.Q = g.Wo;
    p.KH = g.KH; p.KW = g.KW; p.stride = g.stride;
    p.out_H = g.Ho; p.out_W = g.Wo; p.out_sh = 1; p.out_sw = 1;
    p.w_rows_per_group = g.Cout;
    return launch_tc(p, st);
}

// ------------------------------------------------------------------------------------------------
// data gradient on the same kernel
//
// dx[ih] = sum_kh dy[(ih + pad - kh)/s] * w[kh]  over the kh with (ih + pad - kh) % s == 0.
// For the input-parity class ph = (ih + pad) % s only kh = ph + s*t, t < T = K/s, contribute, and with
// ih = ihf + s*i the sum is a T-tap stride-1 convolution over dy:
//     dx[ihf + s*i] = sum_r dy[i - padl + r] * w[ph + s*(T-1-r)],   padl = (T-1) - (ihf + pad - ph)/s
// so each class is a forward convolution of dy with transposed+flipped weights wt[g][class][ci][r][s][co],
// an im2col box with lower corner -padl, and an output written with stride s at offset ihf.
// ------------------------------------------------------------------------------------------------
// CinP >= Cin: rows ci >= Cin are zero (pads the 8 image lanes to the minimum UMMA N of 16)
// 32 (co) x 32 (ci) tiles through shared memory: reads run along ci (contiguous in OHWI), writes along co (contiguous in the
// transposed copy); grid = (co tiles x ci tiles, KH*KW, G)
__global__ void __launch_bounds__(256) dgrad_weight_transform_kernel(const float* __restrict__ w, float* __restrict__ wt, int Cout, int Cin,
                                                                     int CinP, int KH, int KW, int s) {
    pdl_trigger();
    pdl_wait();
    __shared__ float tile[32][33];
    const int TH = KH / s, TW = KW / s;
    const int cit = (CinP + 31) / 32;
    const int co0 = (blockIdx.x / cit) * 32, ci0 = (blockIdx.x % cit) * 32;
    int t = blockIdx.y;
    const int ss = t % TW; t /= TW;
    const int r = t % TH; t /= TH;
    const int cls = t;
    const int g = blockIdx.z;
    const int ph = cls / s, pw = cls - ph * s;
    const int kh = ph + s * (TH - 1 - r), kw = pw + s * (TW - 1 - ss);
    const int tx = threadIdx.x, ty = threadIdx.y;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int co = co0 + ty + 8 * j, ci = ci0 + tx;
        tile[ty + 8 * j][tx] = (co < Cout && ci < Cin) ? __ldg(w + ((((long)g * Cout + co) * KH + kh) * KW + kw) * Cin + ci) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int ci = ci0 + ty + 8 * j, co = co0 + tx;
        if (ci < CinP && co < Cout) wt[(((((long)g * s * s + cls) * CinP + ci) * TH + r) * TW + ss) * Cout + co] = tile[tx][ty + 8 * j];
    }
}

bool tc_dgrad_supported(const cg_conv_geom& g) {
    init_driver();
    if (!g_encode_tiled || !g_encode_im2col) return false;
    int s = g.stride;
    if (s > 2 || g.KH % s || g.KW % s) return false;
    if (pick_bk(g.Cout) == 0) return false;   // K dimension of the dgrad GEMM
    if (pick_bn(g.Cin) == 0) return false;    // N dimension (<= 16 image / head lanes are padded to 16)
    int Hin = g.ups ? 2 * g.H : g.H, Win = g.ups ? 2 * g.W : g.W;
    if (Hin % s || Win % s) return false;
    if ((long)g.B * (Hin / s) * (Win / s) < TC_BM) return false;
    return true;
}

size_t tc_dgrad_ws(const cg_conv_geom& g) {
    size_t wt = (size_t)g.G * g.Cout * g.KH * g.KW * (g.Cin <= 16 ? 16 : g.Cin) * sizeof(float);
    wt = (wt + 1023) & ~(size_t)1023;
    size_t up = g.ups ? (size_t)g.G * g.B * 4 * g.H * g.W * g.Cin * sizeof(float) : 0;
    return wt + up;
}

int tc_conv_dgrad(const cg_conv_geom& g, const float* dy, const float* w, float* dx, const float* addend, const float* mask_src,
                  float mask_slope, void* ws, size_t ws_bytes, cudaStream_t st) {
    init_driver();
    size_t need = tc_dgrad_ws(g);
    if (need > ws_bytes) {
        set_error("conv_dgrad(tc): workspace %zu < %zu bytes", ws_bytes, need);
        return CG_ERR_WORKSPACE;
    }
    const int s = g.stride, TH = g.KH / s, TW = g.KW / s, ncls = s * s;
    const int Hin = g.ups ? 2 * g.H : g.H, Win = g.ups ? 2 * g.W : g.W;
    float* wt = (float*)ws;
    const int CinP = g.Cin <= 16 ? 16 : g.Cin;
    size_t wt_bytes = ((size_t)g.G * g.Cout * g.KH * g.KW * CinP * sizeof(float) + 1023) & ~(size_t)1023;
    float* d_seen = g.ups ? (float*)((uint8_t*)ws + wt_bytes) : dx;
    long total = (long)g.G * g.Cout * g.KH * g.KW * CinP;
    (void)total;
    launch_k(dgrad_weight_transform_kernel, dim3(cdiv(g.Cout, 32) * cdiv(CinP, 32), g.KH * g.KW, g.G), dim3(32, 8), 0, st, w, wt, g.Cout, g.Cin, CinP,
                                                                                                                  g.KH, g.KW, s);
    if (int rc = check_launch("dgrad_weight_transform")) return rc;

    TcParams p{};
    p.bn = pick_bn(g.Cin);
    if (p.bn > 64) p.bn = fill_bn(p.bn, (long)g.B * ((g.ups ? 2 * g.H : g.H) / g.stride) * ((g.ups ? 2 * g.W : g.W) / g.stride), g.G * g.stride * g.stride, g.Cin);
    p.n_store = p.bn == 16 ? g.Cin : p.bn;
    p.bk = pick_bk(g.Cout);
    long ktot = (long)TH * TW * g.Cout;
    if (int rc = encode_weights_map_p(p, wt, (long)g.G * ncls * CinP, ktot)) return rc;
    for (int c = 0; c < ncls; c++) {
        int ph = c / s, pw = c - ph * s;
        int ihf = ((ph - g.pad) % s + s) % s, iwf = ((pw - g.pad) % s + s) % s;
        int padl_h = (TH - 1) - (ihf + g.pad - ph) / s, padl_w = (TW - 1) - (iwf + g.pad - pw) / s;
        int Hc = Hin / s, Wc = Win / s;
        int up_h = Hc - g.Ho - padl_h, up_w = Wc - g.Wo - padl_w;
        if (int rc = encode_act_map(&p.cls[c].amap, dy, (long)g.G * g.B, g.Ho, g.Wo, g.Cout, -padl_w, -padl_h, up_w, up_h, 1, p.bk)) return rc;
        p.cls[c].w0 = -padl_w; p.cls[c].h0 = -padl_h;
        p.cls[c].out_h0 = ihf; p.cls[c].out_w0 = iwf;
        p.cls[c].wrow_off = c * CinP;
    }
    p.ncls = ncls;
    p.G = g.G; p.xg_images = g.B;
    p.B = g.B; p.P = Hin / s; p.Q = Win / s;
    p.Cin = g.Cout; p.Cout = g.Cin; p.KH = TH; p.KW = TW; p.stride = 1;
    p.out_H = Hin; p.out_W = Win; p.out_sh = s; p.out_sw = s;
    p.w_rows_per_group = ncls * CinP;
    p.y = d_seen; p.bias = nullptr;
    p.addend = g.ups ? nullptr : addend; p.mask_src = g.ups ? nullptr : mask_src;
    p.act = CG_ACT_NONE; p.slope = mask_slope;
    if (int rc = launch_tc(p, st)) return rc;
    if (g.ups) return pool2x2_sum(d_seen, dx, addend, mask_src, mask_slope, (long)g.G * g.B, g.H, g.W, g.Cin, st);
    return CG_OK;
}


// ------------------------------------------------------------------------------------------------
// weight gradient on tensor cores
//
//   dW[g][co][kh][kw][ci] = sum over pixels m of dy[g][m][co] * x[g][n][p*s-pad+kh][q*s-pad+kw][ci]
//
// GEMM per (group, 128-cout tile, filter tap, ci tile):  D[128 co][bn ci] += A^T[pixels][co] * B[pixels][ci]
// with the reduction (K) dimension = pixels.  Both operands are "MN-major" in shared memory: the 32-channel
// x KP-pixel boxes TMA delivers from the channels-last tensors (dy as a plain 2-D matrix, x through the same
// im2col map as the forward pass, at the tap's offsets, with the 32-byte-atom 128-byte swizzle) ARE the canonical
// MN-major TF32 layout (cute::UMMA Layout_MN_SW128_32B: 32 contiguous MN elements per row, 4 K rows per 512-byte
// atom), so no transposition is ever materialised.  The pixel range is split across CTAs (split-K) and reduced in a fixed
// order by reduce_splits_kernel, keeping the result deterministic.
// ------------------------------------------------------------------------------------------------
constexpr int WG_KP = 32;      // pixel granularity (tensor-map box rows are kp = 32 or 64)
constexpr int WG_NCOLS = 256;  // accumulator columns per TMEM stage = taps-per-unit * bn

struct WgParams {
    CUtensorMap amap;  // dy  [G*Mpix][Cout]  2-D, box 32 x KP
    CUtensorMap bmap;  // x   im2col, box 32 channels x KP pixels
    int G, xg_images, B, P, Q, Cin, Cout, KH, KW, stride, pad, bn, splits, stages;
    int kp;            // pixels per pipeline stage: 64 when the pixel count allows (fewer, larger TMA boxes), else 32
    int T;             // filter taps accumulated per work unit (they share the dy tile): T * bn <= 256
    int nacc;          // TMEM accumulator stages: 2 (one CTA per SM) or 1 (two co-resident CTAs per SM share the 512 columns)
    int Tm;            // (x-on-M variant) 128-row tiles of (tap, 32-channel block) rows per work unit: Tm * Cout <= 256
    long Mpix, chunk;  // pixels per group; pixels per split (multiple of WG_KP)
    float* out;        // [splits][G][Cout][KH*KW][Cin]
};

// MN-major TF32 operands admit exactly one shared-memory layout (cutlass sm100_common.inl:92 "for mn-major tf32
// operands, SW128_32B is the only available smem layout"): rows of 32 contiguous MN elements (128 B), swizzle atom =
// 4 K-rows x 128 B with 32-byte chunks XOR-ed by the row index (Swizzle<2,5,2>) = TMA CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
// descriptor layout type SWIZZLE_128B_BASE32B (1).  LBO = stride between 32-element MN groups, SBO = stride
// between 4-row K groups (512 B); one K=8 MMA consumes two K groups.

__global__ void __launch_bounds__(TC_THREADS, 2) wgrad_tc_kernel(const __grid_constant__ WgParams p) {
    pdl_trigger();
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int box_bytes = p.kp * 128;
    const int a_bytes = 4 * box_bytes;
    const int nb = p.bn / 32;                     // 32-channel boxes per tap
    const int b_bytes = (WG_NCOLS / 32) * box_bytes;
    const int stage_bytes = a_bytes + b_bytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
    uint64_t* empty_bar = full_bar + p.stages;
    uint64_t* tfull_bar = empty_bar + p.stages;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int taps = p.KH * p.KW;
    const int TG = (taps + p.T - 1) / p.T;        // tap groups
    const int COT = (p.Cout + 127) / 128;
    const int CIT = p.Cin / p.bn;
    const int units = p.G * COT * p.splits * CIT * TG;
    const int tmem_cols = p.nacc * WG_NCOLS;

    if (warp == TC_PRODUCER_WARP && lane == 0) {
        prefetch_tmap(&p.amap);
        prefetch_tmap(&p.bmap);
    }
    if (warp == TC_MMA_WARP) {
        if (lane == 0) {
            for (int s = 0; s < p.stages; s++) {
                mbar_init(&full_bar[s], 1);
                mbar_init(&empty_bar[s], 1);
            }
            for (int a = 0; a < 2; a++) {
                mbar_init(&tfull_bar[a], 1);
                mbar_init(&tempty_bar[a], 128);
            }
            fence_barrier_init();
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();  // on-chip prologue done (barriers, TMEM): from here on the kernel reads what its predecessors in the stream wrote

    // unit -> (g, cot, split, cit, tap group), tap group fastest so CTAs running together share the dy tile in L2
    auto decode = [&](int u, int& g, int& cot, int& sp, int& cit, int& tg) {
        tg = u % TG; u /= TG;
        cit = u % CIT; u /= CIT;
        sp = u % p.splits; u /= p.splits;
        cot = u % COT;
        g = u / COT;
    };

    if (warp == TC_PRODUCER_WARP) {
        // ===================== TMA producer: the whole warp issues, one box per lane =====================
        int stage = 0;
        uint32_t phase = 0;
        for (int u = blockIdx.x; u < units; u += gridDim.x) {
            int g, cot, sp, cit, tg;
            decode(u, g, cot, sp, cit, tg);
            const int tap0 = tg * p.T;
            const int tcount = taps - tap0 < p.T ? taps - tap0 : p.T;
            const long mbeg = (long)sp * p.chunk;
            const long mend = mbeg + p.chunk < p.Mpix ? mbeg + p.chunk : p.Mpix;
            // lane roles: 0..3 -> dy boxes; 4..4+tcount*nb -> x boxes (tap t = q / nb, channel box j = q % nb)
            const int q = lane - 4;
            const bool is_a = lane < 4;
            const bool is_b = q >= 0 && q < tcount * nb;
            const int bt = is_b ? q / nb : 0, bj = is_b ? q - bt * nb : 0;
            const int tap = tap0 + bt;
            const int kh = tap / p.KW, kw = tap - kh * p.KW;
            const uint32_t tx = (uint32_t)((4 + tcount * nb) * box_bytes);
            for (long m = mbeg; m < mend; m += p.kp) {
                int img = (int)(m / (p.P * p.Q));
                int rem = (int)(m - (long)img * p.P * p.Q);
                int pp = rem / p.Q, qq = rem - pp * p.Q;
                mbar_wait(&empty_bar[stage], phase ^ 1);
                uint8_t* sa = smem + (size_t)stage * stage_bytes;
                uint8_t* sb = sa + a_bytes;
                if (lane == 0) mbar_expect_tx(&full_bar[stage], tx);
                __syncwarp();
                if (is_a)
                    tma_load_2d(&p.amap, &full_bar[stage], sa + lane * box_bytes, cot * 128 + lane * 32, (int)((long)g * p.Mpix + m));
                if (is_b)
                    tma_load_im2col_4d(&p.bmap, &full_bar[stage], sb + q * box_bytes, cit * p.bn + bj * 32, -p.pad + qq * p.stride,
                                       -p.pad + pp * p.stride, g * p.xg_images + img, (uint16_t)kw, (uint16_t)kh);
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == TC_MMA_WARP) {
        if (lane == 0) {
            // kind::tf32, D=F32, A and B MN-major (bits 15, 16), M=128, N=bn
            const uint32_t idesc = make_idesc_tf32(p.bn) | (1u << 15) | (1u << 16);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int u = blockIdx.x; u < units; u += gridDim.x) {
                int g, cot, sp, cit, tg;
                decode(u, g, cot, sp, cit, tg);
                const int tap0 = tg * p.T;
                const int tcount = taps - tap0 < p.T ? taps - tap0 : p.T;
                const long mbeg = (long)sp * p.chunk;
                const long mend = mbeg + p.chunk < p.Mpix ? mbeg + p.chunk : p.Mpix;
                mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * WG_NCOLS);
                uint32_t accum = 0;
                for (long m = mbeg; m < mend; m += p.kp) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
                    uint32_t sb = sa + a_bytes;
                    uint64_t adesc = make_mnmajor_sw128_desc(sa, box_bytes);
                    const int nkk = p.kp / 8;
                    for (int kk = 0; kk < nkk; kk++) {
                        // next 8 pixels = next two 512-byte K atoms: +1024 B = +64 in the 16-byte address field
                        for (int t = 0; t < tcount; t++) {
                            uint64_t bdesc = make_mnmajor_sw128_desc(sb + t * nb * box_bytes, box_bytes);
                            umma_tf32(d_tmem + (uint32_t)(t * p.bn), adesc + (uint64_t)(kk * 64), bdesc + (uint64_t)(kk * 64), idesc,
                                      (accum | kk) ? 1u : 0u);
                        }
                    }
                    accum = 1;
                    umma_commit(&empty_bar[stage]);
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
                umma_commit(&tfull_bar[acc]);
                if (++acc == p.nacc) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        const int quad = warp & 3;
        const int row = quad * 32 + lane;
        int acc = 0;
        uint32_t acc_phase = 0;
        const long ktot = (long)taps * p.Cin;
        for (int u = blockIdx.x; u < units; u += gridDim.x) {
            int g, cot, sp, cit, tg;
            decode(u, g, cot, sp, cit, tg);
            const int tap0 = tg * p.T;
            const int tcount = taps - tap0 < p.T ? taps - tap0 : p.T;
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            const int co = cot * 128 + row;
            const bool valid = co < p.Cout;
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * WG_NCOLS);
            for (int t = 0; t < tcount; t++) {
                float* op = p.out + (((long)sp * p.G + g) * p.Cout + co) * ktot + (long)(tap0 + t) * p.Cin + cit * p.bn;
                for (int c0 = 0; c0 < p.bn; c0 += 32) {
                    float v[32];
                    tmem_ld32(taddr + (uint32_t)(t * p.bn + c0), v);
                    if (valid) {
                        float4* yp = reinterpret_cast<float4*>(op + c0);
#pragma unroll
                        for (int j = 0; j < 8; j++) yp[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&tempty_bar[acc]);
            if (++acc == p.nacc) { acc = 0; acc_phase ^= 1; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == TC_MMA_WARP) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols));
    }
}

// CTA-pair variant of the weight gradient: one 256-cout x N tile per pair.  Each CTA stages dy for its own 128 output
// channels and HALF of the x columns of every tap (N/2), so a 64-pixel stage is 64 KB instead of 96 KB (3 stages deep
// instead of 2, 8 TMA boxes instead of 12) and the x tile crosses L2 -> shared memory once per 256 output channels.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1) wgrad_tc2_kernel(const __grid_constant__ WgParams p) {
    pdl_trigger();
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int box_bytes = p.kp * 128;
    const int a_bytes = 4 * box_bytes;
    const int nb2 = p.bn / 64;                    // 32-channel boxes per tap staged by THIS CTA (half of the tap's N)
    const int b_bytes = (WG_NCOLS / 64) * box_bytes;
    const int stage_bytes = a_bytes + b_bytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
    uint64_t* empty_bar = full_bar + p.stages;
    uint64_t* tfull_bar = empty_bar + p.stages;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int taps = p.KH * p.KW;
    const int TG = (taps + p.T - 1) / p.T;
    const int COT = p.Cout / 256;
    const int CIT = p.Cin / p.bn;
    const int units = p.G * COT * p.splits * CIT * TG;
    const int cluster_id = blockIdx.x >> 1, nclusters = gridDim.x >> 1;

    if (warp == TC_PRODUCER_WARP && lane == 0) {
        prefetch_tmap(&p.amap);
        prefetch_tmap(&p.bmap);
    }
    if (warp == TC_MMA_WARP) {
        if (lane == 0) {
            for (int s = 0; s < p.stages; s++) {
                mbar_init(&full_bar[s], 1);
                mbar_init(&empty_bar[s], 1);
            }
            for (int a = 0; a < 2; a++) {
                mbar_init(&tfull_bar[a], 1);
                mbar_init(&tempty_bar[a], 256);
            }
            fence_barrier_init();
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(2 * WG_NCOLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();  // on-chip prologue done (barriers, TMEM): from here on the kernel reads what its predecessors in the stream wrote

    auto decode = [&](int u, int& g, int& cot, int& sp, int& cit, int& tg) {
        tg = u % TG; u /= TG;
        cit = u % CIT; u /= CIT;
        sp = u % p.splits; u /= p.splits;
        cot = u % COT;
        g = u / COT;
    };

    if (warp == TC_PRODUCER_WARP) {
        int stage = 0;
        uint32_t phase = 0;
        for (int u = cluster_id; u < units; u += nclusters) {
            int g, cot, sp, cit, tg;
            decode(u, g, cot, sp, cit, tg);
            const int tap0 = tg * p.T;
            const int tcount = taps - tap0 < p.T ? taps - tap0 : p.T;
            const long mbeg = (long)sp * p.chunk;
            const long mend = mbeg + p.chunk < p.Mpix ? mbeg + p.chunk : p.Mpix;
            // lane roles: 0..3 -> dy boxes of this CTA's 128 couts; 4..4+tcount*nb2 -> this CTA's half of each tap's x boxes
            const int q = lane - 4;
            const bool is_a = lane < 4;
            const bool is_b = q >= 0 && q < tcount * nb2;
            const int bt = is_b ? q / nb2 : 0, bj = is_b ? q - bt * nb2 : 0;
            const int tap = tap0 + bt;
            const int kh = tap / p.KW, kw = tap - kh * p.KW;
            const uint32_t tx = (uint32_t)(2 * (4 + tcount * nb2) * box_bytes);  // both CTAs' boxes land on the leader's barrier
            for (long m = mbeg; m < mend; m += p.kp) {
                int img = (int)(m / (p.P * p.Q));
                int rem = (int)(m - (long)img * p.P * p.Q);
                int pp = rem / p.Q, qq = rem - pp * p.Q;
                mbar_wait(&empty_bar[stage], phase ^ 1);
                uint8_t* sa = smem + (size_t)stage * stage_bytes;
                uint8_t* sb = sa + a_bytes;
                const uint32_t full_leader = leader_addr(smem_u32(&full_bar[stage]));
                if (lane == 0 && rank == 0) mbar_expect_tx(&full_bar[stage], tx);
                __syncwarp();
                if (is_a)
                    tma2_load_2d(&p.amap, full_leader, sa + lane * box_bytes, cot * 256 + (int)rank * 128 + lane * 32,
                                 (int)((long)g * p.Mpix + m));
                if (is_b)
                    tma2_load_im2col_4d(&p.bmap, full_leader, sb + q * box_bytes, cit * p.bn + (int)rank * (p.bn >> 1) + bj * 32,
                                        -p.pad + qq * p.stride, -p.pad + pp * p.stride, g * p.xg_images + img, (uint16_t)kw, (uint16_t)kh);
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == TC_MMA_WARP) {
        if (lane == 0 && rank == 0) {
            // kind::tf32, D=F32, A and B MN-major (bits 15, 16), M = 256 (both CTAs), N = bn
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(p.bn >> 3) << 17) |
                                   ((uint32_t)(256 >> 4) << 24);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int u = cluster_id; u < units; u += nclusters) {
                int g, cot, sp, cit, tg;
                decode(u, g, cot, sp, cit, tg);
                const int tap0 = tg * p.T;
                const int tcount = taps - tap0 < p.T ? taps - tap0 : p.T;
                const long mbeg = (long)sp * p.chunk;
                const long mend = mbeg + p.chunk < p.Mpix ? mbeg + p.chunk : p.Mpix;
                mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * WG_NCOLS);
                uint32_t accum = 0;
                for (long m = mbeg; m < mend; m += p.kp) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
                    uint32_t sb = sa + a_bytes;
                    uint64_t adesc = make_mnmajor_sw128_desc(sa, box_bytes);
                    const int nkk = p.kp / 8;
                    for (int kk = 0; kk < nkk; kk++) {
                        for (int t = 0; t < tcount; t++) {
                            uint64_t bdesc = make_mnmajor_sw128_desc(sb + t * nb2 * box_bytes, box_bytes);
                            umma2_tf32(d_tmem + (uint32_t)(t * p.bn), adesc + (uint64_t)(kk * 64), bdesc + (uint64_t)(kk * 64), idesc,
                                       (accum | kk) ? 1u : 0u);
                        }
                    }
                    accum = 1;
                    umma2_commit_mc(&empty_bar[stage]);
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
                umma2_commit_mc(&tfull_bar[acc]);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        const int quad = warp & 3;
        const int row = quad * 32 + lane;
        int acc = 0;
        uint32_t acc_phase = 0;
        const long ktot = (long)taps * p.Cin;
        for (int u = cluster_id; u < units; u += nclusters) {
            int g, cot, sp, cit, tg;
            decode(u, g, cot, sp, cit, tg);
            const int tap0 = tg * p.T;
            const int tcount = taps - tap0 < p.T ? taps - tap0 : p.T;
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            const int co = cot * 256 + (int)rank * 128 + row;
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * WG_NCOLS);
            for (int t = 0; t < tcount; t++) {
                float* op = p.out + (((long)sp * p.G + g) * p.Cout + co) * ktot + (long)(tap0 + t) * p.Cin + cit * p.bn;
                for (int c0 = 0; c0 < p.bn; c0 += 32) {
                    float v[32];
                    tmem_ld32(taddr + (uint32_t)(t * p.bn + c0), v);
                    float4* yp = reinterpret_cast<float4*>(op + c0);
#pragma unroll
                    for (int j = 0; j < 8; j++) yp[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                }
            }
            tc_fence_before();
            mbar_arrive_cluster(leader_addr(smem_u32(&tempty_bar[acc])));
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == TC_MMA_WARP) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * WG_NCOLS));
    }
}

// Weight gradient with the roles swapped, for layers with <= 64 output channels: the 128 MMA rows are four (tap, 32-channel
// block) row groups of x instead of output channels (which would leave half of every M = 128 instruction empty), dy is the N
// operand (N = Cout).  D[(tap, ci)][co] is written back transposed: for a fixed co the 32 lanes of a warp hold 32
// consecutive ci = one 128-byte store.  Up to Tm row tiles share the dy stage.
__global__ void __launch_bounds__(TC_THREADS, 2) wgrad_xm_kernel(const __grid_constant__ WgParams p) {
    pdl_trigger();
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int box_bytes = p.kp * 128;
    const int nbo = p.Cout / 32;                  // dy boxes (1 or 2)
    const int a_bytes = 2 * box_bytes;            // dy slot (N operand)
    const int b_bytes = p.Tm * 4 * box_bytes;     // x row groups (M operand)
    const int stage_bytes = a_bytes + b_bytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
    uint64_t* empty_bar = full_bar + p.stages;
    uint64_t* tfull_bar = empty_bar + p.stages;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int CB = p.Cin / 32;                    // 32-channel blocks per tap
    const int RG = p.KH * p.KW * CB;              // row groups in all
    const int GPU_ = p.Tm * 4;                    // row groups per unit
    const int MG = (RG + GPU_ - 1) / GPU_;
    const int units = p.G * p.splits * MG;
    const int acc_cols = p.nacc == 1 ? WG_NCOLS / 2 : WG_NCOLS;  // nacc == 1 flags two co-resident CTAs: 2 x 128 columns each
    const int tmem_cols = 2 * acc_cols;

    if (warp == TC_PRODUCER_WARP && lane == 0) {
        prefetch_tmap(&p.amap);
        prefetch_tmap(&p.bmap);
    }
    if (warp == TC_MMA_WARP) {
        if (lane == 0) {
            for (int s = 0; s < p.stages; s++) {
                mbar_init(&full_bar[s], 1);
                mbar_init(&empty_bar[s], 1);
            }
            for (int a = 0; a < 2; a++) {
                mbar_init(&tfull_bar[a], 1);
                mbar_init(&tempty_bar[a], 128);
            }
            fence_barrier_init();
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();  // on-chip prologue done (barriers, TMEM): from here on the kernel reads what its predecessors in the stream wrote

    // unit -> (g, split, row-tile group), row-tile group fastest so CTAs running together share the dy tile in L2
    auto decode = [&](int u, int& g, int& sp, int& mg) {
        mg = u % MG; u /= MG;
        sp = u % p.splits;
        g = u / p.splits;
    };

    if (warp == TC_PRODUCER_WARP) {
        int stage = 0;
        uint32_t phase = 0;
        for (int u = blockIdx.x; u < units; u += gridDim.x) {
            int g, sp, mg;
            decode(u, g, sp, mg);
            const int rg0 = mg * GPU_;
            const int ng = RG - rg0 < GPU_ ? RG - rg0 : GPU_;  // row groups of this unit
            const long mbeg = (long)sp * p.chunk;
            const long mend = mbeg + p.chunk < p.Mpix ? mbeg + p.chunk : p.Mpix;
            // lane roles: 0..nbo-1 -> dy boxes; 2..2+ng -> x row groups (tap, channel block)
            const int q = lane - 2;
            const bool is_a = lane < nbo;
            const bool is_b = q >= 0 && q < ng;
            const int rg = is_b ? rg0 + q : 0;
            const int tap = rg / CB, cb = rg - tap * CB;
            const int kh = tap / p.KW, kw = tap - kh * p.KW;
            const uint32_t tx = (uint32_t)((nbo + ng) * box_bytes);
            for (long m = mbeg; m < mend; m += p.kp) {
                int img = (int)(m / (p.P * p.Q));
                int rem = (int)(m - (long)img * p.P * p.Q);
                int pp = rem / p.Q, qq = rem - pp * p.Q;
                mbar_wait(&empty_bar[stage], phase ^ 1);
                uint8_t* sa = smem + (size_t)stage * stage_bytes;
                uint8_t* sb = sa + a_bytes;
                if (lane == 0) mbar_expect_tx(&full_bar[stage], tx);
                __syncwarp();
                if (is_a) tma_load_2d(&p.amap, &full_bar[stage], sa + lane * box_bytes, lane * 32, (int)((long)g * p.Mpix + m));
                if (is_b)
                    tma_load_im2col_4d(&p.bmap, &full_bar[stage], sb + q * box_bytes, cb * 32, -p.pad + qq * p.stride, -p.pad + pp * p.stride,
                                       g * p.xg_images + img, (uint16_t)kw, (uint16_t)kh);
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == TC_MMA_WARP) {
        if (lane == 0) {
            // kind::tf32, D=F32, A (x) and B (dy) MN-major (bits 15, 16), M=128, N=Cout
            const uint32_t idesc = make_idesc_tf32(p.Cout) | (1u << 15) | (1u << 16);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int u = blockIdx.x; u < units; u += gridDim.x) {
                int g, sp, mg;
                decode(u, g, sp, mg);
                const int rg0 = mg * GPU_;
                const int ng = RG - rg0 < GPU_ ? RG - rg0 : GPU_;
                const int ntile = (ng + 3) >> 2;
                const long mbeg = (long)sp * p.chunk;
                const long mend = mbeg + p.chunk < p.Mpix ? mbeg + p.chunk : p.Mpix;
                mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * acc_cols);
                uint32_t accum = 0;
                for (long m = mbeg; m < mend; m += p.kp) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
                    uint32_t sb = sa + a_bytes;
                    uint64_t ndesc = make_mnmajor_sw128_desc(sa, box_bytes);  // dy: N operand
                    const int nkk = p.kp / 8;
                    for (int kk = 0; kk < nkk; kk++) {
                        for (int t = 0; t < ntile; t++) {
                            // a partial last tile multiplies stale shared memory in its missing row groups: those D rows are never stored
                            uint64_t mdesc = make_mnmajor_sw128_desc(sb + t * 4 * box_bytes, box_bytes);
                            umma_tf32(d_tmem + (uint32_t)(t * p.Cout), mdesc + (uint64_t)(kk * 64), ndesc + (uint64_t)(kk * 64), idesc,
                                      (accum | kk) ? 1u : 0u);
                        }
                    }
                    accum = 1;
                    umma_commit(&empty_bar[stage]);
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
                umma_commit(&tfull_bar[acc]);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        const int quad = warp & 3;
        int acc = 0;
        uint32_t acc_phase = 0;
        const long ktot = (long)p.KH * p.KW * p.Cin;
        for (int u = blockIdx.x; u < units; u += gridDim.x) {
            int g, sp, mg;
            decode(u, g, sp, mg);
            const int rg0 = mg * GPU_;
            const int ng = RG - rg0 < GPU_ ? RG - rg0 : GPU_;
            const int ntile = (ng + 3) >> 2;
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * acc_cols);
            for (int t = 0; t < ntile; t++) {
                const int q = t * 4 + quad;          // this warp's row group inside the unit
                const bool valid = q < ng;
                const int rg = rg0 + q;
                const int tap = rg / CB, cb = rg - tap * CB;
                // dw[sp][g][co][tap][ci], ci = cb*32 + lane: one 128-byte store per output channel
                float* op = p.out + ((long)sp * p.G + g) * p.Cout * ktot + (long)tap * p.Cin + cb * 32 + lane;
                for (int c0 = 0; c0 < p.Cout; c0 += 32) {
                    float v[32];
                    tmem_ld32(taddr + (uint32_t)(t * p.Cout + c0), v);
                    if (valid) {
#pragma unroll
                        for (int j = 0; j < 32; j++) op[(long)(c0 + j) * ktot] = v[j];
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&tempty_bar[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == TC_MMA_WARP) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols));
    }
}

__global__ void reduce_splits_tc_kernel(const float* __restrict__ part, float* __restrict__ out, long n4, int splits) {
    pdl_trigger();
    pdl_wait();
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k = 0; k < splits; k++) {
        float4 v = __ldg(reinterpret_cast<const float4*>(part) + (long)k * n4 + i);
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    reinterpret_cast<float4*>(out)[i] = s;
}

static int wg_bn(int cin) {
    if (cin % 256 == 0) return 256;
    if (cin % 32 == 0 && cin < 256) return cin;  // 32 .. 224: one N tile
    return 0;
}

// CTA pairs (256 output channels per unit) when the layer has them and every tap's N splits into two 32-channel-box halves
static bool wg_pair(const cg_conv_geom& g) { return (g_pair_mode & 8) && g.Cout % 256 == 0 && wg_bn(g.Cin) % 64 == 0; }
// x-on-M variant for <= 64 output channels (a 128-row cout tile would be half empty)
static bool wg_xm(const cg_conv_geom& g) { return g_wgrad_xm && (g.Cout == 64 || g.Cout == 32) && g.Cin % 32 == 0; }
static int wg_xm_tm(const cg_conv_geom& g) {
    int rg = g.KH * g.KW * (g.Cin / 32), mt = (rg + 3) / 4;
    int tm = WG_NCOLS / g.Cout;  // TMEM columns
    if (tm > 2) tm = 2;          // 2 dy + 8 x boxes per 64-pixel stage = 80 KB, two stages (three tiles = 112 KB would leave one)
    if (tm > mt) tm = mt;
    return tm;
}
// 64-pixel stages (fewer, larger TMA boxes) for the one-CTA-per-SM kernels; the plain kernel runs two CTAs per SM with 32-pixel stages
static bool wg_two(const cg_conv_geom& g) { return g_wgrad_2cta && !wg_pair(g) && (!wg_xm(g) || g_wgrad_xm2); }
static int wg_kp(const cg_conv_geom& g) { return (!wg_two(g) && ((long)g.B * g.Ho * g.Wo) % 64 == 0) ? 64 : 32; }
static void wg_plan(const cg_conv_geom& g, int& splits, long& chunk) {
    long Mpix = (long)g.B * g.Ho * g.Wo;
    const int kp = wg_kp(g);
    int bn = wg_bn(g.Cin);
    int T = WG_NCOLS / bn;
    if (T > g.KH * g.KW) T = g.KH * g.KW;
    const bool pair = wg_pair(g);
    long base = (long)g.G * (pair ? g.Cout / 256 : (g.Cout + 127) / 128) * (g.Cin / bn) * ((g.KH * g.KW + T - 1) / T);
    if (wg_xm(g)) {
        int tm = wg_xm_tm(g);
        int rgs = g.KH * g.KW * (g.Cin / 32);
        base = (long)g.G * ((rgs + tm * 4 - 1) / (tm * 4));
    }
    init_driver();
    const int sms = (sm_count_now() > 0 ? sm_count_now() : 148) * (wg_two(g) ? 2 : 1) / (pair ? 2 : 1);
    long maxs = Mpix / (kp * 8);  // at least 8 pipeline stages of work per split
    if (maxs < 1) maxs = 1;
    if (maxs > 64) maxs = 64;
    // pick the split count whose unit count fills whole waves best (persistent grid = #SMs), preferring >= 3 waves
    int best = 1;
    double best_score = -1.0;
    for (int s = 1; s <= maxs; s++) {
        long units = base * s;
        long waves = (units + sms - 1) / sms;
        double eff = (double)units / (double)(waves * sms);
        double score = eff - 0.01 * s - (waves < 3 ? 0.15 * (3 - waves) : 0.0);
        if (score > best_score) { best_score = score; best = s; }
    }
    // Small problems (few output tiles, e.g. a 1x1 64->64 layer: 2 units per member): the score above never leaves one split because
    // each extra split fills < 1 % of a wave, and two CTAs then walk all the pixels (0.15 ms for 134 MFLOP at 128x128).  Below half a
    // wave, take as many splits as fill one wave (each still >= 8 pipeline stages), capped so the partial sums stay <= 32 MB.
    if (g_small_bn && base * best * 2 < sms) {
        long fill = sms / base;
        const long out_bytes = (long)g.G * g.Cout * g.KH * g.KW * g.Cin * 4;
        const long cap = (32l << 20) / (out_bytes > 0 ? out_bytes : 1);
        if (fill > cap) fill = cap;
        if (fill > maxs) fill = maxs;
        if (fill > best) best = (int)fill;
    }
    splits = best;
    chunk = ((Mpix + splits - 1) / splits + kp - 1) / kp * kp;
    splits = (int)((Mpix + chunk - 1) / chunk);
}

bool tc_wgrad_supported(const cg_conv_geom& g) {
    init_driver();
    if (!g_encode_tiled || !g_encode_im2col) return false;
    if (g.ups) return false;
    if (wg_bn(g.Cin) == 0) return false;
    if (g.Cout % 4 != 0) return false;
    long Mpix = (long)g.B * g.Ho * g.Wo;
    if (Mpix % WG_KP != 0 || Mpix < 256) return false;
    if (g.pad > 120 || g.KH > 120) return false;
    return true;
}

size_t tc_wgrad_ws(const cg_conv_geom& g) {
    int splits; long chunk;
    wg_plan(g, splits, chunk);
    if (splits == 1) return 0;
    return (size_t)splits * g.G * g.Cout * g.KH * g.KW * g.Cin * sizeof(float);
}

// dy [rows][C] as a 2-D tensor map with 32-channel x kp-pixel boxes in the MN-major TF32 layout (also used by conv_img.cu)
static int encode_mn_map_raw(CUtensorMap* map, const float* t, long rows, int C, int kp) {
    init_driver();
    if (!g_encode_tiled) {
        set_error("cuTensorMapEncodeTiled unavailable");
        return CG_ERR_CUDA;
    }
    cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)C * 4};
    cuuint32_t box[2] = {32, (cuuint32_t)kp};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode_tiled(map, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 2, (void*)t, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(mn-major map rows=%ld C=%d) failed: %d", rows, C, (int)r);
        return CG_ERR_CUDA;
    }
    return CG_OK;
}
int tc_encode_mn_map(CUtensorMap* map, const float* t, long rows, int C, int kp) {
    MapKey k{};
    k.ptr = t; k.a = rows; k.v[0] = C; k.v[1] = kp; k.v[7] = 3;
    return cached_map(map, k, [&](CUtensorMap* m) { return encode_mn_map_raw(m, t, rows, C, kp); });
}
// y [rows][C] fp32 as a 2-D store map with 32-channel x box_rows boxes, 128-byte swizzled shared-memory side (epilogues that stage
// their tile in shared memory and leave through cp.async.bulk.tensor stores)
static int encode_store_map_raw(CUtensorMap* map, float* t, long rows, int C, int box_rows) {
    init_driver();
    if (!g_encode_tiled) {
        set_error("cuTensorMapEncodeTiled unavailable");
        return CG_ERR_CUDA;
    }
    cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)C * 4};
    cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode_tiled(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)t, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(store map rows=%ld C=%d) failed: %d", rows, C, (int)r);
        return CG_ERR_CUDA;
    }
    return CG_OK;
}
int tc_encode_store_map(CUtensorMap* map, float* t, long rows, int C, int box_rows) {
    MapKey k{};
    k.ptr = t; k.a = rows; k.v[0] = C; k.v[1] = box_rows; k.v[7] = 4;
    return cached_map(map, k, [&](CUtensorMap* m) { return encode_store_map_raw(m, t, rows, C, box_rows); });
}
int tc_sm_count() { return sm_count_now(); }

int tc_conv_wgrad(const cg_conv_geom& g, const float* x, const float* dy, float* dw, void* ws, size_t ws_bytes, cudaStream_t st) {
    init_driver();
    WgParams p{};
    wg_plan(g, p.splits, p.chunk);
    size_t need = tc_wgrad_ws(g);
    if (need > ws_bytes) {
        set_error("conv_wgrad(tc): workspace %zu < %zu bytes", ws_bytes, need);
        return CG_ERR_WORKSPACE;
    }
    p.bn = wg_bn(g.Cin);
    p.kp = wg_kp(g);
    p.Mpix = (long)g.B * g.Ho * g.Wo;
    if (int rc = tc_encode_mn_map(&p.amap, dy, (long)g.G * p.Mpix, g.Cout, p.kp)) return rc;
    long nimg = (long)(g.x_groups == 1 ? 1 : g.G) * g.B;
    {
        MapKey k{};
        k.ptr = x; k.a = nimg; k.b = ((int64_t)g.H << 32) | (uint32_t)g.W;
        k.v[0] = g.Cin; k.v[1] = g.pad; k.v[2] = g.KW; k.v[3] = g.KH; k.v[4] = g.stride; k.v[5] = p.kp; k.v[7] = 5;
        const int kp = p.kp;
        int rc = cached_map(&p.bmap, k, [&](CUtensorMap* m) {
            cuuint64_t dims[4] = {(cuuint64_t)g.Cin, (cuuint64_t)g.W, (cuuint64_t)g.H, (cuuint64_t)nimg};
            cuuint64_t strides[3] = {(cuuint64_t)g.Cin * 4, (cuuint64_t)g.W * g.Cin * 4, (cuuint64_t)g.H * g.W * g.Cin * 4};
            int lower[2] = {-g.pad, -g.pad};
            int upper[2] = {g.pad - (g.KW - 1), g.pad - (g.KH - 1)};
            cuuint32_t estr[4] = {1, (cuuint32_t)g.stride, (cuuint32_t)g.stride, 1};
            CUresult r = g_encode_im2col(m, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 4, (void*)x, dims, strides, lower, upper, 32, kp, estr,
                                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) {
                set_error("cuTensorMapEncodeIm2col(x for wgrad) failed: %d", (int)r);
                return (int)CG_ERR_CUDA;
            }
            if (g_driver_version <= 13010 && nimg * g.H * g.W * g.Cin * 4 < 131072) reinterpret_cast<uint64_t*>(m)[1] &= ~(1ull << 21);
            return (int)CG_OK;
        });
        if (rc) return rc;
    }
    p.G = g.G; p.xg_images = g.x_groups == 1 ? 0 : g.B;
    p.B = g.B; p.P = g.Ho; p.Q = g.Wo; p.Cin = g.Cin; p.Cout = g.Cout; p.KH = g.KH; p.KW = g.KW;
    p.stride = g.stride; p.pad = g.pad;
    p.out = p.splits == 1 ? dw : (float*)ws;
    p.T = WG_NCOLS / p.bn;
    if (p.T > g.KH * g.KW) p.T = g.KH * g.KW;
    if (wg_xm(g)) {
        p.Tm = wg_xm_tm(g);
        const bool two = wg_two(g) && p.Tm * g.Cout <= WG_NCOLS / 2;  // two co-resident CTAs: 2 x 128 TMEM columns each
        p.nacc = two ? 1 : 2;
        int stage_bytes = p.kp * 128 * (2 + 4 * p.Tm);
        int stages = ((two ? 100 : 200) * 1024) / stage_bytes;
        if (stages > 8) stages = 8;
        p.stages = stages;
        size_t smem = (size_t)stages * stage_bytes + 1024 + (2 * stages + 4) * 8 + 16;
        static PerDeviceOnce attr3_set;
        if (attr3_set.first()) {
            cudaError_t e = cudaFuncSetAttribute(wgrad_xm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
            if (e != cudaSuccess) {
                set_error("cudaFuncSetAttribute(wgrad_xm_kernel): %s", cudaGetErrorString(e));
                return CG_ERR_CUDA;
            }
        }
        int rgs = g.KH * g.KW * (g.Cin / 32);
        long units = (long)g.G * p.splits * ((rgs + p.Tm * 4 - 1) / (p.Tm * 4));
        const int slots = sm_count_now() * (two ? 2 : 1);
        int grid = (int)(units < slots ? units : slots);
        if (getenv("COUNCIL_DEBUG")) fprintf(stderr, "wgrad_xm: units=%ld splits=%d chunk=%ld stages=%d Tm=%d kp=%d\n", units, p.splits, p.chunk, stages, p.Tm, p.kp);
        launch_k(wgrad_xm_kernel, grid, TC_THREADS, smem, st, p);
        if (int rc = check_launch("wgrad_xm_kernel")) return rc;
        if (p.splits > 1) {
            long n4 = (long)g.G * g.Cout * g.KH * g.KW * g.Cin / 4;
            launch_k(reduce_splits_tc_kernel, cdiv(n4, 256), 256, 0, st, (const float*)ws, dw, n4, p.splits);
            return check_launch("reduce_splits_tc");
        }
        return CG_OK;
    }
    if (wg_pair(g)) {
        int stage_bytes = p.kp * 128 * (4 + WG_NCOLS / 64);
        int stages = (200 * 1024) / stage_bytes;
        if (stages > 8) stages = 8;
        p.stages = stages;
        size_t smem = (size_t)stages * stage_bytes + 1024 + (2 * stages + 4) * 8 + 16;
        static PerDeviceOnce attr2_set;
        if (attr2_set.first()) {
            cudaError_t e = cudaFuncSetAttribute(wgrad_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
            if (e != cudaSuccess) {
                set_error("cudaFuncSetAttribute(wgrad_tc2_kernel): %s", cudaGetErrorString(e));
                return CG_ERR_CUDA;
            }
        }
        long units = (long)g.G * (g.Cout / 256) * p.splits * (g.Cin / p.bn) * ((g.KH * g.KW + p.T - 1) / p.T);
        int pairs = sm_count_now() / 2;
        int nclusters = (int)(units < pairs ? units : pairs);
        if (getenv("COUNCIL_DEBUG")) fprintf(stderr, "wgrad_tc2: units=%ld splits=%d chunk=%ld stages=%d T=%d bn=%d kp=%d\n", units, p.splits, p.chunk, stages, p.T, p.bn, p.kp);
        launch_k(wgrad_tc2_kernel, 2 * nclusters, TC_THREADS, smem, st, p);
        if (int rc = check_launch("wgrad_tc2_kernel")) return rc;
        if (p.splits > 1) {
            long n4 = (long)g.G * g.Cout * g.KH * g.KW * g.Cin / 4;
            launch_k(reduce_splits_tc_kernel, cdiv(n4, 256), 256, 0, st, (const float*)ws, dw, n4, p.splits);
            return check_launch("reduce_splits_tc");
        }
        return CG_OK;
    }
    // two co-resident CTAs per SM (g_wgrad_2cta): 32-pixel stages, two per CTA, one TMEM accumulator stage each -- the MMAs of two
    // independent units interleave on the tensor pipe
    const bool two = wg_two(g);
    p.nacc = two ? 1 : 2;
    int stage_bytes = p.kp * 128 * (4 + WG_NCOLS / 32);
    int stages = ((two ? 100 : 200) * 1024) / stage_bytes;
    p.stages = stages;
    size_t smem = (size_t)stages * stage_bytes + 1024 + (2 * stages + 4) * 8 + 16;
    static PerDeviceOnce attr_set;
    if (attr_set.first()) {
        cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) {
            set_error("cudaFuncSetAttribute(wgrad_tc_kernel): %s", cudaGetErrorString(e));
            return CG_ERR_CUDA;
        }
    }
    long units = (long)g.G * ((g.Cout + 127) / 128) * p.splits * (g.Cin / p.bn) * ((g.KH * g.KW + p.T - 1) / p.T);
    const int slots = sm_count_now() * (two ? 2 : 1);
    int grid = (int)(units < slots ? units : slots);
    launch_k(wgrad_tc_kernel, grid, TC_THREADS, smem, st, p);
    if (int rc = check_launch("wgrad_tc_kernel")) return rc;
    if (p.splits > 1) {
        long n4 = (long)g.G * g.Cout * g.KH * g.KW * g.Cin / 4;
        launch_k(reduce_splits_tc_kernel, cdiv(n4, 256), 256, 0, st, (const float*)ws, dw, n4, p.splits);
        return check_launch("reduce_splits_tc");
    }
    return CG_OK;
}

}  // namespace cg
