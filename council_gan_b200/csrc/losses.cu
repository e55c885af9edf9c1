// Fused loss kernels of the three updates: LSGAN over every discriminator scale, the focus
// losses, the loss-history matching and all their gradients -- 1 launch per discriminator update and direction,
// 2 launches per gen_update and direction (a reduction pass, then -- after the data-parallel all-reduce of the 6N
// scalars -- one pass that finalises the loss values ON THE DEVICE and writes every gradient).  No host round trip.
//
// Reference semantics (paths relative to the reference tree):
//   MsImageDis.calc_dis_loss / calc_gen_loss                networks.py:56-64, 84-90   (lsgan)
//   MsImageDisCouncil.calc_dis_loss / calc_gen_loss         networks.py:158-166, 188-194
//   mask_zero_one_criterion / mask_small_criterion(_square) / mask_criterion_TV   trainer_council.py:230-250
//   loss-history matching                                   trainer_council.py:518-524, 576-586
//   total-loss assembly in gen_update                       trainer_council.py:392-451, 497-529, 559-634
#include "common.cuh"

namespace cg {

__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float sgnf(float x) { return (float)((x > 0.f) - (x < 0.f)); }

template <int N>
__device__ __forceinline__ void block_sum(float (&v)[N], float* smem /* >= N*32 */) {
#pragma unroll
    for (int i = 0; i < N; i++)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    __syncthreads();
    if (lane == 0)
#pragma unroll
        for (int i = 0; i < N; i++) smem[i * 32 + warp] = v[i];
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int i = 0; i < N; i++) {
            float t = lane < nw ? smem[i * 32 + lane] : 0.f;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
            v[i] = t;
        }
    }
}

// "last block done" ticket: returns true in exactly one block (all threads), after every other block's partials are visible.
__device__ __forceinline__ bool last_block_done(unsigned int* counter, unsigned int nblocks) {
    __shared__ unsigned int s_ticket;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_ticket = atomicAdd(counter, 1u);
    __syncthreads();
    bool last = s_ticket == nblocks - 1;
    if (last) __threadfence();
    return last;
}

// ---------------------------------------------------------------------------------------------------------------
// discriminator updates: LSGAN loss of every scale + its gradient, one launch
// ---------------------------------------------------------------------------------------------------------------
// block b -> (map m, member g, segment s); ws: float part[nmaps][G][nseg] then the ticket counter
__global__ void __launch_bounds__(256) lsgan_fused_kernel(const cg_lsgan_desc d, float* __restrict__ loss_total, int accumulate,
                                                          float* __restrict__ loss_plain, float* __restrict__ part,
                                                          unsigned int* __restrict__ counter) {
    pdl_trigger();
    pdl_wait();
    __shared__ float sm[32];
    const int nseg = d.nseg, G = d.G;
    int b = blockIdx.x;
    const int m = b / (G * nseg);
    b -= m * G * nseg;
    const int g = b / nseg, s = b - g * nseg;
    const int n = d.n_per_seg[m];
    const float t = d.target[s];
    const float coef = d.grad_scale * d.weight[g][s] * 2.0f / (float)n;
    const float* p = d.out[m] + ((long)g * nseg + s) * n;
    float* q = d.dout[m] ? d.dout[m] + ((long)g * nseg + s) * n : nullptr;
    float v[1] = {0.f};
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        float df = __ldg(p + i) - t;
        v[0] += df * df;
        if (q) q[i] = coef * df;
    }
    block_sum<1>(v, sm);
    if (threadIdx.x == 0) part[(m * G + g) * nseg + s] = v[0];
    if (!last_block_done(counter, gridDim.x)) return;
    if (threadIdx.x < G) {
        const int gg = threadIdx.x;
        const volatile float* vp = part;
        float tot = 0.f, plain = 0.f;
        for (int mm = 0; mm < d.nmaps; mm++) {
            float tm = 0.f, pm = 0.f;
            for (int ss = 0; ss < nseg; ss++) {
                float mean = vp[(mm * G + gg) * nseg + ss] / (float)d.n_per_seg[mm];
                tm += d.weight[gg][ss] * mean;
                pm += mean;
            }
            tot += tm;
            plain += pm;
        }
        tot *= d.loss_scale;
        loss_total[gg] = accumulate ? loss_total[gg] + tot : tot;
        if (loss_plain) loss_plain[gg] = plain * d.loss_scale;
    }
    if (threadIdx.x == 0) *counter = 0u;
}

// ---------------------------------------------------------------------------------------------------------------
// gen_update, pass 1: every reduction of the generator loss in one launch
// ---------------------------------------------------------------------------------------------------------------
constexpr int GL_PIX = 2048;  // mask pixels per focus block

// blocks [0, nmaps*G): LSGAN maps (adversarial D maps first, then council-D maps); then nchunks*G focus blocks.
// ws: float part_map[nmaps][G]; float part_focus[nchunks][G][4]; ticket counter.
__global__ void __launch_bounds__(256) gen_loss_fwd_kernel(const cg_gen_loss_desc d, float* __restrict__ scal, float* __restrict__ part_map,
                                                           float* __restrict__ part_focus, unsigned int* __restrict__ counter, int nchunks) {
    pdl_trigger();
    pdl_wait();
    __shared__ float sm[4 * 32];
    const int G = d.G, nmaps = d.n_adv + d.n_cl;
    const int b = blockIdx.x;
    if (b < nmaps * G) {
        const int m = b / G, g = b - m * G;
        const bool adv = m < d.n_adv;
        const int n = adv ? d.adv_n[m] : d.cl_n[m - d.n_adv];
        const float* p = (adv ? d.adv_out[m] : d.cl_out[m - d.n_adv]) + (long)g * n;
        float* q = adv && d.adv_dout[m] ? d.adv_dout[m] + (long)g * n : nullptr;
        const float coef = d.adv_grad_scale * 2.0f / (float)n;
        float v[1] = {0.f};
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            float df = __ldg(p + i) - 1.0f;  // calc_gen_loss: target 1 (networks.py:90,194)
            v[0] += df * df;
            if (q) q[i] = coef * df;
        }
        block_sum<1>(v, sm);
        if (threadIdx.x == 0) part_map[m * G + g] = v[0];
    } else {
        const int fb = b - nmaps * G;
        const int g = fb % G, chunk = fb / G;
        const int H = d.H, W = d.W;
        const long npix = (long)d.B * H * W;
        const long p0 = (long)chunk * GL_PIX, p1 = min(npix, p0 + GL_PIX);
        const float* mb = d.mask + (long)g * npix * 4;
        const float center = d.center, eps = d.eps;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        for (long px = p0 + threadIdx.x; px < p1; px += blockDim.x) {
            int w = (int)(px % W);
            int h = (int)((px / W) % H);
            float4 mk = ld4(mb + px * 4);
            v[0] += 1.f / (fabsf(mk.x - center) + eps) + 1.f / (fabsf(mk.y - center) + eps) + 1.f / (fabsf(mk.z - center) + eps);
            v[1] += mk.x + mk.y + mk.z;
            if (h + 1 < H) {
                float4 q = ld4(mb + (px + W) * 4);
                v[2] += fabsf(q.x - mk.x) + fabsf(q.y - mk.y) + fabsf(q.z - mk.z);
            }
            if (w + 1 < W) {
                float4 r = ld4(mb + (px + 1) * 4);
                v[3] += fabsf(r.x - mk.x) + fabsf(r.y - mk.y) + fabsf(r.z - mk.z);
            }
        }
        block_sum<4>(v, sm);
        if (threadIdx.x == 0) {
            float* o = part_focus + ((long)chunk * G + g) * 4;
            o[0] = v[0]; o[1] = v[1]; o[2] = v[2]; o[3] = v[3];
        }
    }
    if (!last_block_done(counter, gridDim.x)) return;
    // scal[g] = { sum_scales mean (D(x)-1)^2, sum_scales mean (DC(x)-1)^2, focus sums[4] }   (local to this rank)
    const int t = threadIdx.x;
    if (t < G * 6) {
        const int g = t / 6, k = t - g * 6;
        float r = 0.f;
        if (k < 2) {
            const volatile float* vp = part_map;
            int m0 = k == 0 ? 0 : d.n_adv, m1 = k == 0 ? d.n_adv : nmaps;
            for (int m = m0; m < m1; m++) {
                int n = m < d.n_adv ? d.adv_n[m] : d.cl_n[m - d.n_adv];
                r += vp[m * G + g] / (float)n;
            }
        } else if (d.mask) {
            const volatile float* vp = part_focus;
            double s = 0.0;
            for (int c = 0; c < nchunks; c++) s += (double)vp[((long)c * G + g) * 4 + (k - 2)];
            r = (float)s;
        }
        scal[g * 6 + k] = r;
    }
    if (t == 0) *counter = 0u;
}

// ---------------------------------------------------------------------------------------------------------------
// gen_update, pass 2: finalise the loss values on the device, then every remaining gradient, one launch
// ---------------------------------------------------------------------------------------------------------------
struct GenCoef {      // per member, computed redundantly by every block
    float cdis;       // council-loss weight of this member: w_match * council_w          (gradient of DC maps)
    float c01, csum, ctv;  // focus-loss gradient coefficients
};

// hist rings: double [G][hist+1]; the live window is positions (head+k) % (hist+1), k = 0..hist-1; an append writes
// position (head+hist) % (hist+1) (not in anybody's read set) and the HOST advances head afterwards.
__device__ __forceinline__ double ring_mean_after_append(const double* ring, int R, int head, int hist, double v, int lane) {
    double s = 0.0;
    for (int k = 1 + lane; k < hist; k += 32) s += ring[(head + k) % R];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    return (s + v) / (double)hist;
}
__device__ __forceinline__ double ring_mean(const double* ring, int R, int head, int hist, int lane) {
    double s = 0.0;
    for (int k = lane; k < hist; k += 32) s += ring[(head + k) % R];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    return s / (double)hist;
}

__global__ void __launch_bounds__(256) gen_loss_bwd_kernel(const cg_gen_loss_desc d, const cg_gen_loss_hp hp, const float* __restrict__ scal,
                                                           double* __restrict__ hist_gan, double* __restrict__ hist_council,
                                                           float* __restrict__ total, double* __restrict__ total64, int accumulate,
                                                           float* __restrict__ pub, float* __restrict__ d_mask, int map_blocks) {
    pdl_trigger();
    pdl_wait();
    __shared__ GenCoef sc[CG_LOSS_MAX_G];
    const int G = d.G;
    const int R = hp.hist_size + 1;
    // ---- finalise (warp 0 of every block; block 0 also publishes and appends to the histories) --------------------
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        for (int g = 0; g < G; g++) {
            const double adv = (double)scal[g * 6 + 0] / (double)hp.world;  // mean over the GLOBAL minibatch (equal shards)
            const double cl = (double)scal[g * 6 + 1] / (double)hp.world;
            const double l01 = (double)scal[g * 6 + 2] / hp.numel;
            const double msum = (double)scal[g * 6 + 3] / hp.numel;
            const double ltv = ((double)scal[g * 6 + 4] + (double)scal[g * 6 + 5]) / hp.numel;
            double tot = 0.0, ltot = 0.0, c01 = 0.0, csum = 0.0, ctv = 0.0;
            if (hp.focus_on) {
                if (hp.w01 != 0.0) { tot += hp.w01 * l01; c01 = hp.w01 / hp.numel; }            // trainer_council.py:392-415
                if (hp.wtv != 0.0) { tot += hp.wtv * ltv; ctv = hp.wtv / hp.numel; }            // :425-431
                if (hp.wtot != 0.0) {                                                            // :418-422, :447-451
                    if (hp.small_abs) { ltot += fabs(msum); csum += hp.wtot * (double)((msum > 0.0) - (msum < 0.0)) / hp.numel; }
                    if (hp.small_square) { ltot += msum * msum; csum += hp.wtot * 2.0 * msum / hp.numel; }
                    tot += hp.wtot * ltot;
                }
            }
            double mean_gan;
            const double* rg = hist_gan + (long)g * R;
            const double adv32 = (double)(float)adv;  // the history stores the float32 loss value (:520)
            if (hp.gan_on) {
                mean_gan = hp.matching ? ring_mean_after_append(rg, R, hp.head_gan, hp.hist_size, adv32, lane)
                                       : ring_mean(rg, R, hp.head_gan, hp.hist_size, lane);
                tot += hp.gan_w * adv;
            } else {
                mean_gan = ring_mean(rg, R, hp.head_gan, hp.hist_size, lane);
            }
            double w = 1.0, closs = 0.0, cdis = 0.0;
            if (hp.council_on) {
                if (hp.matching) {  // :576-586
                    const double cl32 = (double)(float)cl;
                    double mean_c = ring_mean_after_append(hist_council + (long)g * R, R, hp.head_council, hp.hist_size, cl32, lane);
                    w = mean_gan / mean_c;
                }
                closs = cl * (double)(float)w * hp.council_w;  // float32 tensor * python float (cast to float32) * council_w
                tot += closs;
                cdis = w * hp.council_w;
            }
            if (lane == 0) {
                sc[g].cdis = (float)cdis;
                sc[g].c01 = (float)c01;
                sc[g].csum = (float)csum;
                sc[g].ctv = (float)ctv;
                if (blockIdx.x == 0) {
                    double t64 = accumulate ? total64[g] + tot : tot;  // directions are summed in double, published as float32
                    total64[g] = t64;
                    total[g] = (float)t64;
                    float* o = pub + g * 8;
                    o[0] = (float)tot; o[1] = (float)adv; o[2] = (float)l01; o[3] = (float)ltot; o[4] = (float)ltv;
                    o[5] = (float)closs; o[6] = (float)w; o[7] = (float)cl;
                    if (hp.gan_on && hp.matching) hist_gan[(long)g * R + (hp.head_gan + hp.hist_size) % R] = adv32;
                    if (hp.council_on && hp.matching)
                        hist_council[(long)g * R + (hp.head_council + hp.hist_size) % R] = (double)(float)cl;
                }
            }
        }
    }
    __syncthreads();
    // ---- gradients ------------------------------------------------------------------------------------------------
    if ((int)blockIdx.x < map_blocks) {
        // council-D patch maps: d out = cdis[g] * 2 / (n * world) * (out - 1), one block per (map, member)
        const int m = blockIdx.x / G, g = blockIdx.x - m * G;
        const int n = d.cl_n[m];
        const float coef = sc[g].cdis * (2.0f / ((float)n * (float)hp.world));
        const float* p = d.cl_out[m] + (long)g * n;
        float* q = d.cl_dout[m] + (long)g * n;
        for (int i = threadIdx.x; i < n; i += blockDim.x) q[i] = coef * (__ldg(p + i) - 1.0f);
        return;
    }
    if (!d_mask) return;
    const int H = d.H, W = d.W;
    const long npix = (long)d.B * H * W, total_px = npix * G;
    const float center = d.center, eps = d.eps;
    const long stride = (long)(gridDim.x - map_blocks) * blockDim.x;
    for (long i = (long)(blockIdx.x - map_blocks) * blockDim.x + threadIdx.x; i < total_px; i += stride) {
        const int g = (int)(i / npix);
        const long px = i - (long)g * npix;
        const int w = (int)(px % W);
        const int h = (int)((px / W) % H);
        const float c01 = sc[g].c01, cs = sc[g].csum, ctv = sc[g].ctv;
        const float* mp = d.mask + i * 4;
        float4 mk = ld4(mp);
        float mm[3] = {mk.x, mk.y, mk.z}, o[3];
#pragma unroll
        for (int c = 0; c < 3; c++) {
            float df = mm[c] - center;
            float den = fabsf(df) + eps;
            o[c] = cs - c01 * sgnf(df) / (den * den);
        }
        if (ctv != 0.f) {
            float tv[3] = {0.f, 0.f, 0.f};
            if (h + 1 < H) { float4 q = ld4(mp + (long)W * 4); tv[0] -= sgnf(q.x - mk.x); tv[1] -= sgnf(q.y - mk.y); tv[2] -= sgnf(q.z - mk.z); }
            if (h > 0)     { float4 q = ld4(mp - (long)W * 4); tv[0] += sgnf(mk.x - q.x); tv[1] += sgnf(mk.y - q.y); tv[2] += sgnf(mk.z - q.z); }
            if (w + 1 < W) { float4 q = ld4(mp + 4);           tv[0] -= sgnf(q.x - mk.x); tv[1] -= sgnf(q.y - mk.y); tv[2] -= sgnf(q.z - mk.z); }
            if (w > 0)     { float4 q = ld4(mp - 4);           tv[0] += sgnf(mk.x - q.x); tv[1] += sgnf(mk.y - q.y); tv[2] += sgnf(mk.z - q.z); }
#pragma unroll
            for (int c = 0; c < 3; c++) o[c] += ctv * tv[c];
        }
        reinterpret_cast<float4*>(d_mask)[i] = make_float4(o[0], o[1], o[2], 0.f);
    }
}

}  // namespace cg

using namespace cg;
#define ST ((cudaStream_t)stream)

extern "C" size_t cg_loss_workspace_bytes(int G, int B, int H, int W) {
    long npix = (long)B * H * W;
    size_t nchunks = (size_t)cdiv(npix > 0 ? npix : 1, GL_PIX);
    // maps partials (<= 4 maps x G x 8 segments) + focus partials + ticket + the double accumulators of pass 2
    return 16 + (size_t)CG_LOSS_MAX_MAPS * CG_LOSS_MAX_G * CG_LOSS_MAX_SEG * 4 + nchunks * (size_t)G * 16 + CG_LOSS_MAX_G * 8;
}

// workspace layout (all three calls): [0,16) ticket counter (must be zero on entry; left zero), then double total64[MAX_G],
// then float partials
static inline unsigned int* ws_counter(void* ws) { return (unsigned int*)ws; }
static inline double* ws_total64(void* ws) { return (double*)((char*)ws + 16); }
static inline float* ws_part(void* ws) { return (float*)((char*)ws + 16 + CG_LOSS_MAX_G * 8); }

extern "C" int cg_lsgan_fused(const cg_lsgan_desc* d, float* loss_total, int accumulate, float* loss_plain, void* ws,
                              size_t ws_bytes, void* stream) {
    CG_REQUIRE(d && d->nmaps >= 1 && d->nmaps <= CG_LOSS_MAX_MAPS && d->G >= 1 && d->G <= CG_LOSS_MAX_G && d->nseg >= 1 &&
               d->nseg <= CG_LOSS_MAX_SEG, "lsgan_fused: nmaps=%d G=%d nseg=%d out of range", d ? d->nmaps : -1, d ? d->G : -1,
               d ? d->nseg : -1);
    size_t need = 16 + CG_LOSS_MAX_G * 8 + (size_t)d->nmaps * d->G * d->nseg * 4;
    if (need > ws_bytes) {
        set_error("lsgan_fused: workspace %zu < %zu bytes", ws_bytes, need);
        return CG_ERR_WORKSPACE;
    }
    for (int m = 0; m < d->nmaps; m++) CG_REQUIRE(d->out[m] && d->n_per_seg[m] > 0, "lsgan_fused: map %d is empty", m);
    int blocks = d->nmaps * d->G * d->nseg;
    launch_k(lsgan_fused_kernel, blocks, 256, 0, ST, *d, loss_total, accumulate, loss_plain, ws_part(ws), ws_counter(ws));
    return check_launch("lsgan_fused");
}

static int check_gen_desc(const cg_gen_loss_desc* d, const char* who) {
    CG_REQUIRE(d && d->G >= 1 && d->G <= CG_LOSS_MAX_G && d->n_adv >= 0 && d->n_adv <= 2 && d->n_cl >= 0 && d->n_cl <= 2,
               "%s: G / map counts out of range", who);
    CG_REQUIRE(d->G * 6 <= 256, "%s: G too large", who);
    return CG_OK;
}

extern "C" int cg_gen_loss_fwd(const cg_gen_loss_desc* d, float* scal, void* ws, size_t ws_bytes, void* stream) {
    if (int rc = check_gen_desc(d, "gen_loss_fwd")) return rc;
    long npix = (long)d->B * d->H * d->W;
    int nchunks = d->mask ? cdiv(npix, GL_PIX) : 0;
    int nmaps = d->n_adv + d->n_cl;
    size_t need = 16 + CG_LOSS_MAX_G * 8 + ((size_t)nmaps * d->G + (size_t)nchunks * d->G * 4) * 4;
    if (need > ws_bytes) {
        set_error("gen_loss_fwd: workspace %zu < %zu bytes", ws_bytes, need);
        return CG_ERR_WORKSPACE;
    }
    int blocks = nmaps * d->G + nchunks * d->G;
    if (blocks == 0) {  // nothing to reduce: all six scalars are zero
        cudaError_t e = cudaMemsetAsync(scal, 0, (size_t)d->G * 6 * 4, ST);
        if (e != cudaSuccess) {
            set_error("gen_loss_fwd: %s", cudaGetErrorString(e));
            return CG_ERR_CUDA;
        }
        return CG_OK;
    }
    float* part_map = ws_part(ws);
    float* part_focus = part_map + (size_t)nmaps * d->G;
    launch_k(gen_loss_fwd_kernel, blocks, 256, 0, ST, *d, scal, part_map, part_focus, ws_counter(ws), nchunks);
    return check_launch("gen_loss_fwd");
}

extern "C" int cg_gen_loss_bwd(const cg_gen_loss_desc* d, const cg_gen_loss_hp* hp, const float* scal, double* hist_gan,
                               double* hist_council, float* total, int accumulate, float* pub, float* d_mask, void* ws,
                               size_t ws_bytes, void* stream) {
    if (int rc = check_gen_desc(d, "gen_loss_bwd")) return rc;
    CG_REQUIRE(hp && hp->world >= 1 && hp->hist_size >= 1 && hp->numel > 0, "gen_loss_bwd: bad hyper-parameters");
    CG_REQUIRE(ws_bytes >= 16 + CG_LOSS_MAX_G * 8, "gen_loss_bwd: workspace too small");
    CG_REQUIRE(!d_mask || d->mask, "gen_loss_bwd: d_mask without mask");
    for (int m = 0; m < d->n_cl; m++) CG_REQUIRE(d->cl_dout[m] && d->cl_out[m], "gen_loss_bwd: council map %d without buffers", m);
    int map_blocks = d->n_cl * d->G;
    long total_px = d_mask ? (long)d->G * d->B * d->H * d->W : 0;
    int px_blocks = d_mask ? cdiv(total_px, 256) : 0;
    if (px_blocks > 148 * 8) px_blocks = 148 * 8;  // grid-stride: each block pays the finalise prologue once
    int blocks = map_blocks + px_blocks;
    if (blocks == 0) blocks = 1;  // the finalise / publish step always runs
    launch_k(gen_loss_bwd_kernel, blocks, 256, 0, ST, *d, *hp, scal, hist_gan, hist_council, total, ws_total64(ws), accumulate, pub, d_mask,
                                                map_blocks);
    return check_launch("gen_loss_bwd");
}
