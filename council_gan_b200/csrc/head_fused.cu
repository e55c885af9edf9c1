// The decoder tail of a no-grad generator pass as ONE kernel (Decoder_V2_atten, networks.py:391-407):
//
//   AdaIN normalise + ReLU of the last 3x3 block  ->  1x1 64->64 + ReLU  ->  1x1 64->64 + ReLU  ->  1x1 64->12 + tanh
//   ->  attention-mask compositing with the input image  ->  x_fake, mask
//
// As separate launches every arrow is a round trip of a [G][B][256][256][64] fp32 map (537 MB at the BASELINE shape) through
// HBM: norm_act_fwd 0.39 + two 1x1 convolutions 0.24 each + the 12-channel head 0.13 + mask head 0.05 ms = 1.05 ms per decoder
// pass, three such passes per iteration (dis_update, dis_council_update x2; the pass of gen_update keeps its activations for the
// backward and stays unfused).  Fused, the 64-channel map is read once and 8 floats per pixel are written.
//
// One CTA = 128 threads = 128 pixels per tile, three co-resident CTAs per SM hide each other's latencies.  All threads
// gather / transform the tile into the K-major 128-byte-swizzled shared-memory operand (TF32-rounded), one thread issues the
// tcgen05 MMAs (accumulators in TMEM), all threads read the accumulator back (tcgen05.ld), apply bias + ReLU and write the
// next layer's operand into the same shared-memory tile.  The three weight matrices of the CTA's council member stay resident.
#include "common.cuh"
#include "tc_ptx.cuh"

namespace cg {

struct HeadP {
    const float* y; const float* mean; const float* rstd; const float* adain;
    const float* w1; const float* b1; const float* w2; const float* b2; const float* w3; const float* b3;
    const float* x_img; float* x_fake; float* mask;
    int G, B, HW, P, off, cpg;
};

constexpr int HD_THREADS = 128;
constexpr int HD_A = 32768;   // activation tile: 2 chunks of [128 rows x 128 B]
constexpr int HD_W = 16384;   // 64x64 weight matrix: 2 chunks of [64 rows x 128 B]
constexpr int HD_W3 = 4096;   // 16x64 (12 used): 2 chunks of [16 rows x 128 B]

__device__ __forceinline__ float4 hd_ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// rows x 64 K-major matrix (row stride 64 floats in global) -> chunk j = k / 32: [rows][128 B], 16-byte unit u at u ^ (row % 8)
__device__ __forceinline__ void stage_weights(uint8_t* dst, const float* w, int rows, int rows_pad) {
    for (int i = threadIdx.x; i < 2 * rows_pad * 8; i += HD_THREADS) {
        const int u = i & 7, r = (i >> 3) % rows_pad, j = i / (8 * rows_pad);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < rows) v = to_tf32(hd_ldg4(w + (long)r * 64 + j * 32 + u * 4));
        *reinterpret_cast<float4*>(dst + j * (rows_pad * 128) + (r >> 3) * 1024 + (r & 7) * 128 + ((u ^ (r & 7)) << 4)) = v;
    }
}

__global__ void __launch_bounds__(HD_THREADS, 3) head_fused_kernel(const HeadP p) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;
    uint8_t* sW1 = sA + HD_A;
    uint8_t* sW2 = sW1 + HD_W;
    uint8_t* sW3 = sW2 + HD_W;
    float* s_b1 = reinterpret_cast<float*>(sW3 + HD_W3);  // 64
    float* s_b2 = s_b1 + 64;                               // 64
    float* s_b3 = s_b2 + 64;                               // 16
    uint64_t* bar = reinterpret_cast<uint64_t*>(s_b3 + 16);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = blockIdx.x % p.G, cidx = blockIdx.x / p.G;
    const int tiles = (int)(((long)p.B * p.HW) >> 7);

    stage_weights(sW1, p.w1 + (long)g * 64 * 64, 64, 64);
    stage_weights(sW2, p.w2 + (long)g * 64 * 64, 64, 64);
    stage_weights(sW3, p.w3 + (long)g * 12 * 64, 12, 16);
    if (threadIdx.x < 64) {
        s_b1[threadIdx.x] = __ldg(p.b1 + g * 64 + threadIdx.x);
        s_b2[threadIdx.x] = __ldg(p.b2 + g * 64 + threadIdx.x);
    }
    if (threadIdx.x < 16) s_b3[threadIdx.x] = threadIdx.x < 12 ? __ldg(p.b3 + g * 12 + threadIdx.x) : 0.f;
    if (warp == 0) {
        if (lane == 0) {
            mbar_init(bar, 1);
            fence_barrier_init();
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(128));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t d_main = tmem_base, d_head = tmem_base + 64;
    const uint32_t idesc64 = make_idesc_tf32(64), idesc16 = make_idesc_tf32(16);
    const uint64_t desc_hi = make_kmajor_sw128_desc(0);
    const uint32_t a_addr = smem_u32(sA);
    uint32_t phase = 0;

    // loader mapping (coalesced: 8 threads read one 128-byte line): 16-byte unit u of rows r0, r0 + 16, ...
    const int u = threadIdx.x & 7, r0 = threadIdx.x >> 3;
    // accumulator mapping: thread = pixel row of the tile (TMEM lane), warp w reads lanes 32 w .. 32 w + 31
    const int row = threadIdx.x;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;

    auto issue = [&](uint32_t w_addr, int w_chunk_bytes, uint32_t d_tmem, uint32_t idesc) {
        // D[128 x N] = A[128 x 64] * W^T: K = 64 = 2 chunks x 4 steps of 8
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        if (threadIdx.x == 0) {
            tc_fence_after();
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const int j = q >> 2, kk = q & 3;
                const uint64_t adesc = (desc_hi | (uint64_t)(((a_addr + j * 16384) & 0x3FFFF) >> 4)) + (uint64_t)(kk * 2);
                const uint64_t bdesc = (desc_hi | (uint64_t)(((w_addr + j * w_chunk_bytes) & 0x3FFFF) >> 4)) + (uint64_t)(kk * 2);
                umma_tf32(d_tmem, adesc, bdesc, idesc, q != 0 ? 1u : 0u);
            }
            umma_commit(bar);
        }
        mbar_wait(bar, phase);
        phase ^= 1;
        tc_fence_after();
    };
    // accumulator (64 columns) + bias, ReLU, TF32 -> the shared-memory operand of the next layer (row = this thread)
    auto relayer = [&](const float* bias) {
        uint8_t* dst = sA + (row >> 3) * 1024 + (row & 7) * 128;
        const int sw = row & 7;
#pragma unroll
        for (int c0 = 0; c0 < 64; c0 += 32) {
            float v[32];
            tmem_ld32(d_main + lane_base + (uint32_t)c0, v);
#pragma unroll
            for (int q = 0; q < 8; q++) {
                float4 o;
                o.x = fmaxf(v[4 * q] + bias[c0 + 4 * q], 0.f);
                o.y = fmaxf(v[4 * q + 1] + bias[c0 + 4 * q + 1], 0.f);
                o.z = fmaxf(v[4 * q + 2] + bias[c0 + 4 * q + 2], 0.f);
                o.w = fmaxf(v[4 * q + 3] + bias[c0 + 4 * q + 3], 0.f);
                *reinterpret_cast<float4*>(dst + (c0 >> 5) * 16384 + ((q ^ sw) << 4)) = to_tf32(o);
            }
        }
    };

    for (int tile = cidx; tile < tiles; tile += p.cpg) {
        const long m0 = (long)tile * 128;        // first pixel of the tile within the member (B * HW pixels)
        const int img = (int)(m0 / p.HW);        // HW % 128 == 0: a tile lies inside one image
        const long gb = (long)g * p.B + img;
        // ---- AdaIN affine of this thread's 8 channels (c = j * 32 + u * 4 .. + 3), then gather + transform the tile
        float4 a[2], b[2];
#pragma unroll
        for (int j = 0; j < 2; j++) {
            const int c = j * 32 + u * 4;
            const float4 mu = hd_ldg4(p.mean + gb * 64 + c), rs = hd_ldg4(p.rstd + gb * 64 + c);
            float4 ga = make_float4(1.f, 1.f, 1.f, 1.f), be = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.adain) {
                const float* ap = p.adain + gb * p.P + p.off;
                be = hd_ldg4(ap + c);
                ga = hd_ldg4(ap + 64 + c);
            }
            a[j] = make_float4(ga.x * rs.x, ga.y * rs.y, ga.z * rs.z, ga.w * rs.w);
            b[j] = make_float4(be.x - mu.x * a[j].x, be.y - mu.y * a[j].y, be.z - mu.z * a[j].z, be.w - mu.w * a[j].w);
        }
        const float* yb = p.y + ((long)g * p.B * p.HW + m0) * 64 + u * 4;
        float4 v[16];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            v[2 * i] = hd_ldg4(yb + (long)(r0 + 16 * i) * 64);
            v[2 * i + 1] = hd_ldg4(yb + (long)(r0 + 16 * i) * 64 + 32);
        }
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int r = r0 + 16 * i;
            uint8_t* dst = sA + (r >> 3) * 1024 + (r & 7) * 128 + ((u ^ (r & 7)) << 4);
#pragma unroll
            for (int j = 0; j < 2; j++) {
                const float4 x = v[2 * i + j];
                float4 o;
                o.x = fmaxf(fmaf(x.x, a[j].x, b[j].x), 0.f);
                o.y = fmaxf(fmaf(x.y, a[j].y, b[j].y), 0.f);
                o.z = fmaxf(fmaf(x.z, a[j].z, b[j].z), 0.f);
                o.w = fmaxf(fmaf(x.w, a[j].w, b[j].w), 0.f);
                *reinterpret_cast<float4*>(dst + j * 16384) = to_tf32(o);
            }
        }
        issue(smem_u32(sW1), 8192, d_main, idesc64);   // dec.model.7: 1x1 64->64
        relayer(s_b1);
        issue(smem_u32(sW2), 8192, d_main, idesc64);   // dec.model.8: 1x1 64->64
        relayer(s_b2);
        issue(smem_u32(sW3), 2048, d_head, idesc16);   // dec.model.9: 1x1 64->12 (N padded to 16), tanh below
        // ---- mask head (networks.py:398-407): h = tanh(.), [o0 rgb | o1 rgb | o2 rgb | m0 m1 m2]
        float h[16];
        tmem_ld16(d_head + lane_base, h);
#pragma unroll
        for (int j = 0; j < 12; j++) h[j] = tanhf(h[j] + s_b3[j]);
        const long pix = m0 + row;                       // pixel within the member
        const float4 xi = hd_ldg4(p.x_img + (pix % ((long)p.B * p.HW)) * 4);
        float im[3] = {xi.x, xi.y, xi.z};
        float mk[3];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            mk[k] = (tanhf(10.f * h[9 + k]) + 1.f) * 0.5f;
#pragma unroll
            for (int ch = 0; ch < 3; ch++) im[ch] = (1.f - mk[k]) * im[ch] + mk[k] * h[3 * k + ch];
        }
        const long o = ((long)g * p.B * p.HW + pix) * 4;
        *reinterpret_cast<float4*>(p.x_fake + o) = make_float4(im[0], im[1], im[2], 0.f);
        *reinterpret_cast<float4*>(p.mask + o) = make_float4(mk[0], mk[1], mk[2], 0.f);
        tc_fence_before();  // the next tile's first MMA overwrites d_main / the shared tile only after the barrier inside issue()
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128));
    }
}

}  // namespace cg

using namespace cg;

extern "C" int cg_head_fused(const float* y, const float* mean, const float* rstd, const float* adain, int P, int off, const float* w1,
                             const float* b1, const float* w2, const float* b2, const float* w3, const float* b3, const float* x_img,
                             float* x_fake, float* mask, int G, int B, int HW, void* stream) {
    CG_REQUIRE(G >= 1 && B >= 1 && HW % 128 == 0, "head_fused: needs H*W %% 128 == 0 (got %d)", HW);
    const int sms = tc_sm_count();
    CG_REQUIRE(sms > 0 && G <= 3 * sms, "head_fused: no device / too many groups");
    HeadP p{};
    p.y = y; p.mean = mean; p.rstd = rstd; p.adain = adain; p.w1 = w1; p.b1 = b1; p.w2 = w2; p.b2 = b2; p.w3 = w3; p.b3 = b3;
    p.x_img = x_img; p.x_fake = x_fake; p.mask = mask; p.G = G; p.B = B; p.HW = HW; p.P = P; p.off = off;
    p.cpg = 3 * sms / G;
    const long tiles = (long)B * HW / 128;
    if (p.cpg > tiles) p.cpg = (int)tiles;
    const size_t smem = HD_A + 2 * HD_W + HD_W3 + (64 + 64 + 16) * 4 + 8 + 16 + 1024;
    static PerDeviceOnce attr_once;
    if (attr_once.first()) {
        cudaError_t e = cudaFuncSetAttribute(head_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            set_error("cudaFuncSetAttribute(head_fused_kernel): %s", cudaGetErrorString(e));
            return CG_ERR_CUDA;
        }
    }
    launch_k(head_fused_kernel, G * p.cpg, HD_THREADS, smem, (cudaStream_t)stream, p);
    return check_launch("head_fused");
}
