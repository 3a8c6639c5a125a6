// gemm_bcsr_ring.cu -- BCSR GEMM with X and W staged through shared memory (the default BCSR kernel; TSG_BCSR_RING=0 or
// tsg_bcsr_set_kernel(1) selects the plain kernel of gemm_bcsr.cu).
//
// Same math and the same sequence of fp32 roundings as k_bcsr_gemm (gemm_bcsr.cu; reference sparse/bcsr.c:141-175):
// every output starts from its bias and takes one FFMA per stored block element in ascending k.  What changes is where
// the operands come from.  The plain kernel re-reads a 512-byte row of the X tile from L2 for every block of every
// block-column (4096^2, 1x8 blocks: 34 GB of L2 reads per call -> L2-bandwidth bound at ~22 % of the FFMA peak).  Here a
// persistent CTA per SM walks (128 rows of X) x (256 output columns) units like the TCSC kernel does: a producer thread
// streams kc-row chunks of the K-major X tile AND the matching run of W block rows (local k + c values each, a private
// chunk-major copy of the matrix: BStream; an r x c block is r consecutive entries) into a two-stage shared-memory ring with bulk async copies (TMA engine)
// completing on mbarriers; 16 compute warps own 16 output columns each (16/c block-columns), lane l holds rows
// l, l+32, l+64, l+96 -> 64 accumulators per thread.  Per block row: one conflict-free LDS.128 of X, c/4 uniform
// LDS.128 of values, 4c FFMA -> FFMA-issue bound for c >= 8 instead of L2 bound.
// Measured on B200 (profiles/bcsr_ring_check_r01.json): 4096^3, 1x8 blocks, 50 % sparsity 7.69 -> 3.19 ms, bit-identical.
#include "tsg_f32x2.cuh"
#include "tsg_internal.h"
#include "tsg_ptx.cuh"

namespace tsg {

constexpr int BR_TM = 128;                        // rows of X per tile (XT layout of gemm_tcsc.cu)
constexpr int BR_NWARP = 16;                      // compute warps
constexpr int BR_THREADS = (BR_NWARP + 4) * 32;   // + one producer warpgroup (register re-balancing is per warpgroup)
constexpr int BR_REGS_COMPUTE = 112, BR_REGS_PRODUCER = 24;  // same balance as k_tcsc_gemm: 4*128*(112-96) <= 128*(96-24)
constexpr int BR_TN = 256;                        // output columns per unit
constexpr int BR_CW = BR_TN / BR_NWARP;           // output columns per warp
constexpr int BR_CNT_BYTES = 256, BR_WSTART_BYTES = 64;
constexpr size_t BR_SMEM_MAX = 232448;
#ifndef BR_UNROLL
#define BR_UNROLL 8
#endif
constexpr int BR_UNROLL_N = BR_UNROLL;  // entries of a block-column in flight per warp

// ---- BStream builder ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ int lower_bound_i32(const int *__restrict__ a, int lo, int hi, int key) {
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(a + mid) < key) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

// one CTA per run (tile, chunk), thread t = block-column t of the tile: how many of its blocks fall into the chunk
// (crow is ascending inside a block-column), their position inside the run, and the run's padded length
__global__ void __launch_bounds__(256) k_bs_count(const int *__restrict__ cptr, const int *__restrict__ crow, int bc, int tbc, int bpw, int kcb,
                                                  int r, int nchunk, uint8_t *__restrict__ cnt, uint32_t *__restrict__ colpos,
                                                  uint32_t *__restrict__ colsrc, uint32_t *__restrict__ wstart,
                                                  uint32_t *__restrict__ run_total, int *__restrict__ max_run) {
    __shared__ uint32_t wsum[8];
    const int run = blockIdx.x, tile = run / nchunk, chunk = run % nchunk;
    const int t = threadIdx.x, col = tile * tbc + t;
    int n = 0, lo = 0;
    if (t < tbc && col < bc) {
        const int b = __ldg(cptr + col), e = __ldg(cptr + col + 1);
        lo = lower_bound_i32(crow, b, e, chunk * kcb);
        n = (lower_bound_i32(crow, lo, e, (chunk + 1) * kcb) - lo) * r;  // <= kcb * r <= 224 block rows
    }
    const int lane = t & 31, w = t >> 5;
    uint32_t v = (uint32_t)n;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t up = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += up;
    }
    if (lane == 31) wsum[w] = v;
    __syncthreads();
    uint32_t base = 0, total = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        if (i < w) base += wsum[i];
        total += wsum[i];
    }
    const uint32_t excl = base + v - (uint32_t)n;
    const size_t o = (size_t)run * 256 + t;
    cnt[o] = (uint8_t)n;
    colpos[o] = excl;
    colsrc[o] = (uint32_t)lo;
    if (t < tbc && (t % bpw) == 0) wstart[(size_t)run * BR_NWARP + t / bpw] = excl;
    if (t == 0) {
        const uint32_t padded = (total + 15u) & ~15u;
        run_total[run] = padded;
        atomicMax(max_run, (int)padded);
    }
}

// one CTA per run: warp per block-column, lanes copy the column's local block-rows and its values (coalesced writes)
__global__ void __launch_bounds__(256) k_bs_fill(const int *__restrict__ crow, const int *__restrict__ cblk, const float *__restrict__ values,
                                                 int r, int c, int tbc, int kcb, int nchunk, const uint8_t *__restrict__ cnt,
                                                 const uint32_t *__restrict__ colpos, const uint32_t *__restrict__ colsrc,
                                                 const uint32_t *__restrict__ eoff, uint8_t *__restrict__ hdr, float *__restrict__ val) {
    const int run = blockIdx.x, chunk = run % nchunk;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t e0 = __ldg(eoff + run);
    for (int t = w; t < tbc; t += 8) {
        const size_t o = (size_t)run * 256 + t;
        const int n = cnt[o];
        if (n == 0) continue;
        const uint32_t dst = e0 + colpos[o], src = colsrc[o];
        for (int i = lane; i < n; i += 32) {  // row of X inside the chunk
            const int blk = i / r;
            hdr[dst + i] = (uint8_t)((__ldg(crow + src + blk) - chunk * kcb) * r + (i - blk * r));
        }
        const int rc = r * c, nf = n * c;  // a block's r x c values are row-major: r consecutive entries
        for (int q = lane; q < nf; q += 32) {
            const int blk = q / rc, j = q - blk * rc;
            val[(size_t)dst * c + q] = __ldg(values + (size_t)__ldg(cblk + src + blk) * rc + j);
        }
    }
}

static size_t ring_stage_bytes(int kcb, int r, int c, int max_run) {
    return (size_t)kcb * r * BR_TM * 4 + (size_t)max_run * c * 4 + (size_t)max_run + BR_CNT_BYTES + BR_WSTART_BYTES;
}

static int build_bstream(tsg_bcsr *W) {
    std::lock_guard<std::recursive_mutex> lk(W->mu);
    BStream &bs = W->bs;
    if (bs.built || bs.unsupported) return TSG_OK;
    const int c = W->c, r = W->r;
    if (!(c == 1 || c == 2 || c == 4 || c == 8 || c == 16) || r > 224 || W->br <= 0 || W->bc <= 0) {
        bs.unsupported = true;
        return TSG_OK;
    }
    TSG_TRY(bcsr_build_cols(W));
    cudaStream_t st = stream();
    bs.tbc = BR_TN / c;
    bs.ntile = (W->bc + bs.tbc - 1) / bs.tbc;
    const int bpw = BR_CW / c;
    uint32_t *colpos = nullptr, *colsrc = nullptr, *run_total = nullptr, *total = nullptr;
    int *max_run_dev = nullptr;
    auto drop_scratch = [&]() {
        dev_free(colpos); dev_free(colsrc); dev_free(run_total); dev_free(total); dev_free(max_run_dev);
        colpos = colsrc = run_total = total = nullptr;
        max_run_dev = nullptr;
    };
    auto drop_stream = [&]() {
        dev_free(bs.cnt); dev_free(bs.wstart); dev_free(bs.eoff); dev_free(bs.hdr); dev_free(bs.val);
        bs.cnt = nullptr; bs.wstart = nullptr; bs.eoff = nullptr; bs.hdr = nullptr; bs.val = nullptr;
    };
    struct Cleanup {  // any early error return leaves neither scratch nor a half-built stream behind
        decltype(drop_scratch) &scratch;
        decltype(drop_stream) &stream;
        BStream &bs;
        ~Cleanup() {
            scratch();
            if (!bs.built) stream();
        }
    } cleanup{drop_scratch, drop_stream, bs};
    // largest chunk whose worst run still fits two stages: try a descending ladder of chunk heights (rows of X)
    const int ladder[8] = {224, 160, 112, 72, 56, 40, 24, 8};
    int nrun = 0, prev_kcb = -1;
    bool ok = false;
    for (int li = 0; li < 8 && !ok; ++li) {
        int kcb = ladder[li] / r;
        if (kcb < 1) kcb = 1;
        if (kcb > W->br) kcb = W->br;
        if (kcb == prev_kcb) continue;
        prev_kcb = kcb;
        const int nchunk = (W->br + kcb - 1) / kcb;
        nrun = bs.ntile * nchunk;
        TSG_TRY(dev_alloc_t(&bs.cnt, (size_t)nrun * 256));
        TSG_TRY(dev_alloc_t(&bs.wstart, (size_t)nrun * BR_NWARP));
        TSG_TRY(dev_alloc_t(&bs.eoff, (size_t)nrun + 2));
        TSG_TRY(dev_alloc_t(&colpos, (size_t)nrun * 256));
        TSG_TRY(dev_alloc_t(&colsrc, (size_t)nrun * 256));
        TSG_TRY(dev_alloc_t(&run_total, (size_t)nrun + 2));
        TSG_TRY(dev_alloc_t(&total, 1));
        TSG_TRY(dev_alloc_t(&max_run_dev, 1));
        TSG_CUDA(cudaMemsetAsync(max_run_dev, 0, sizeof(int), st));
        TSG_CUDA(cudaMemsetAsync(bs.wstart, 0, (size_t)nrun * BR_NWARP * 4, st));
        k_bs_count<<<nrun, 256, 0, st>>>(W->cptr, W->crow, W->bc, bs.tbc, bpw, kcb, r, nchunk, bs.cnt, colpos, colsrc, bs.wstart, run_total,
                                         max_run_dev);
        TSG_KERNEL_CHECK("k_bs_count");
        int max_run = 0;
        TSG_CUDA(cudaMemcpyAsync(&max_run, max_run_dev, sizeof(int), cudaMemcpyDeviceToHost, st));
        TSG_CUDA(cudaStreamSynchronize(st));
        if (2 * ring_stage_bytes(kcb, r, c, max_run) + 64 <= BR_SMEM_MAX) {
            ok = true;
            bs.kcb = kcb;
            bs.nchunk = nchunk;
            bs.max_run = max_run;
        } else {
            drop_scratch();
            drop_stream();
        }
    }
    if (!ok) {
        bs.unsupported = true;
        return TSG_OK;
    }
    TSG_TRY(scan_exclusive_u32(run_total, bs.eoff, nrun, total));
    TSG_CUDA(cudaMemcpyAsync(bs.eoff + nrun, total, 4, cudaMemcpyDeviceToDevice, st));
    uint32_t h_total = 0;
    TSG_CUDA(cudaMemcpyAsync(&h_total, total, 4, cudaMemcpyDeviceToHost, st));
    TSG_CUDA(cudaStreamSynchronize(st));
    bs.entries = h_total;
    TSG_TRY(dev_alloc_t(&bs.hdr, (size_t)h_total + 16));
    TSG_TRY(dev_alloc_t(&bs.val, ((size_t)h_total + 16) * c));
    TSG_CUDA(cudaMemsetAsync(bs.hdr, 0, (size_t)h_total + 16, st));
    TSG_CUDA(cudaMemsetAsync(bs.val, 0, ((size_t)h_total + 16) * c * sizeof(float), st));
    k_bs_fill<<<nrun, 256, 0, st>>>(W->crow, W->cblk, W->values, r, c, bs.tbc, bs.kcb, bs.nchunk, bs.cnt, colpos, colsrc, bs.eoff, bs.hdr, bs.val);
    TSG_KERNEL_CHECK("k_bs_fill");
    bs.built = true;
    return TSG_OK;
}

// ---- the kernel ---------------------------------------------------------------------------------------------------------
struct BcsrRingParams {
    const float *XT;
    const uint8_t *cnt;
    const uint32_t *wstart;
    const uint32_t *eoff;
    const uint8_t *hdr;
    const float *val;
    const float *B;
    float *Y;
    long long ldy;
    int M, N, K, r, br, ncov;  // ncov = bc * c: columns W covers
    int kcb, nchunk, ntile, units;
    int full_units, sub;  // units [full_units, units) are dealt as `sub` column slices each (balances the last round over the SMs)
    float a;
    int use_prelu;
    uint32_t xstage_bytes, val_stage_bytes, hdr_stage_bytes;
};

// the block rows of this warp's block-columns inside the staged run, in (block-column, ascending k) order
template <int C, int NB>
__device__ __forceinline__ void bcsr_chunk(float (&acc)[NB * C][4], const float *xs, const float *val_s, const uint8_t *hdr_s,
                                           const uint8_t *cn, uint32_t ei) {
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        const int n = cn[b];
#pragma unroll BR_UNROLL_N
        for (int t = 0; t < n; ++t, ++ei) {
            const int kl = hdr_s[ei];
            const float4 x = *reinterpret_cast<const float4 *>(xs + (size_t)kl * BR_TM);  // rows l, l+32, l+64, l+96 of X row kl
            const float *wv = val_s + (size_t)ei * C;
            float w[C];
            if constexpr (C >= 4) {
#pragma unroll
                for (int q = 0; q < C / 4; ++q) {
                    const float4 t4 = *reinterpret_cast<const float4 *>(wv + 4 * q);
                    w[(4 * q) % C] = t4.x; w[(4 * q + 1) % C] = t4.y; w[(4 * q + 2) % C] = t4.z; w[(4 * q + 3) % C] = t4.w;
                }
            } else {
#pragma unroll
                for (int j = 0; j < C; ++j) w[j] = wv[j];
            }
            // bcsr.c:160-170: y += x * w, ascending k, one rounding per step
            if constexpr (C % 2 == 0) {
                // the same 4C fmas as C*2 packed pairs (tsg_f32x2.cuh): a pair couples (row v, column j) with (row v+1, column
                // j+1), so that both the X pair and the W pair are register pairs as the LDS.128 delivered them (the crossed
                // combination takes the X pair swapped: 4 MOVs per block row instead of 4C/2 more issue slots)
                const float2 x01 = make_float2(x.x, x.y), x10 = make_float2(x.y, x.x), x23 = make_float2(x.z, x.w), x32 = make_float2(x.w, x.z);
#pragma unroll
                for (int j = 0; j < C; j += 2) {
                    float(&a0)[4] = acc[b * C + j];
                    float(&a1)[4] = acc[b * C + j + 1];
                    const float2 wp = make_float2(w[j], w[j + 1]);
                    float2 d;
                    d = ffma2(x01, wp, make_float2(a0[0], a1[1])); a0[0] = d.x; a1[1] = d.y;
                    d = ffma2(x10, wp, make_float2(a0[1], a1[0])); a0[1] = d.x; a1[0] = d.y;
                    d = ffma2(x23, wp, make_float2(a0[2], a1[3])); a0[2] = d.x; a1[3] = d.y;
                    d = ffma2(x32, wp, make_float2(a0[3], a1[2])); a0[3] = d.x; a1[2] = d.y;
                }
            } else {
#pragma unroll
                for (int j = 0; j < C; ++j) {
                    acc[b * C + j][0] = fmaf(x.x, w[j], acc[b * C + j][0]);
                    acc[b * C + j][1] = fmaf(x.y, w[j], acc[b * C + j][1]);
                    acc[b * C + j][2] = fmaf(x.z, w[j], acc[b * C + j][2]);
                    acc[b * C + j][3] = fmaf(x.w, w[j], acc[b * C + j][3]);
                }
            }
        }
    }
}

// one unit (or column slice `part` of one) on the consumer side: NB block-columns = NB*C output columns per warp, 4 rows per lane
template <int C, int NB>
__device__ __forceinline__ void ring_unit(const BcsrRingParams &p, uint8_t *smem, uint32_t stage_bytes, uint64_t *full, uint64_t *empty,
                                          uint32_t &it, int u, int part, int warp, int lane) {
    constexpr int BPW = BR_CW / C, W = NB * C;
    const int mt = u / p.ntile, tile = u % p.ntile;
    const int q0 = (part * BR_NWARP + warp) * NB;  // first block-column (inside the tile) of this warp
    const int nbase = tile * BR_TN + q0 * C;
    const int mbase = mt * BR_TM + lane;
    const bool vec_ok = ((p.ldy & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.Y) & 15) == 0) && (W % 4 == 0);
    float acc[W][4];
#pragma unroll
    for (int j = 0; j < W; ++j) {
        const float b = (nbase + j < p.ncov) ? __ldg(p.B + nbase + j) : 0.f;  // bcsr.c:146-150: Y starts as the bias
        acc[j][0] = b; acc[j][1] = b; acc[j][2] = b; acc[j][3] = b;
    }
    for (int c = 0; c < p.nchunk; ++c, ++it) {
        const uint32_t s = it & 1u;
        mbar_wait(&full[s], (it >> 1) & 1u);
        const uint8_t *st = smem + (size_t)s * stage_bytes;
        const float *xs = reinterpret_cast<const float *>(st) + lane * 4;
        const float *val_s = reinterpret_cast<const float *>(st + p.xstage_bytes);
        const uint8_t *hdr_s = st + p.xstage_bytes + p.val_stage_bytes;
        const uint8_t *cnt_s = hdr_s + p.hdr_stage_bytes;
        const uint32_t *wstart_s = reinterpret_cast<const uint32_t *>(cnt_s + BR_CNT_BYTES);
        uint32_t ei = wstart_s[q0 / BPW];  // entries are stored per block-column in tile order; wstart marks every BPW-th column
        if constexpr (NB < BPW) {
            for (int j = (q0 / BPW) * BPW; j < q0; ++j) ei += cnt_s[j];
        }
        bcsr_chunk<C, NB>(acc, xs, val_s, hdr_s, cnt_s + q0, ei);
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
    }
    const bool full_vec = vec_ok && (nbase + W <= p.ncov);
#pragma unroll
    for (int v = 0; v < 4; ++v) {
        const int m = mbase + 32 * v;
        if (m >= p.M) continue;
        float *yrow = p.Y + (size_t)m * p.ldy + nbase;
        float out[W];
#pragma unroll
        for (int j = 0; j < W; ++j) {
            float y = acc[j][v];
            if (p.use_prelu) y = (y < 0.0f) ? p.a * y : y;
            out[j] = y;
        }
        if (full_vec) {
            if constexpr (W % 4 == 0) {
#pragma unroll
                for (int j = 0; j < W; j += 4) *reinterpret_cast<float4 *>(yrow + j) = make_float4(out[j], out[j + 1], out[j + 2], out[j + 3]);
            }
        } else {
#pragma unroll
            for (int j = 0; j < W; ++j)
                if (nbase + j < p.ncov) yrow[j] = out[j];
        }
    }
}

template <int C>
__global__ void __launch_bounds__(BR_THREADS, 1) k_bcsr_gemm_ring(const BcsrRingParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t stage_bytes = p.xstage_bytes + p.val_stage_bytes + p.hdr_stage_bytes + BR_CNT_BYTES + BR_WSTART_BYTES;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + 2 * (size_t)stage_bytes);
    uint64_t *empty = full + 2;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        mbar_init(&empty[0], BR_NWARP);
        mbar_init(&empty[1], BR_NWARP);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp >= BR_NWARP) {
        // ===== producer warpgroup: hands its registers to the compute warpgroups; one thread feeds the two-stage ring =====
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(BR_REGS_PRODUCER));
        if (warp == BR_NWARP && lane == 0) {
            uint32_t it = 0;
            const int nvirt = p.full_units + (p.units - p.full_units) * p.sub;
            for (int v = blockIdx.x; v < nvirt; v += gridDim.x) {
                const int u = (v < p.full_units) ? v : p.full_units + (v - p.full_units) / p.sub;  // a slice stages the whole run of its unit
                const int mt = u / p.ntile, tile = u % p.ntile;
                for (int c = 0; c < p.nchunk; ++c, ++it) {
                    const uint32_t s = it & 1u;
                    mbar_wait(&empty[s], ((it >> 1) & 1u) ^ 1u);
                    uint8_t *st = smem + (size_t)s * stage_bytes;
                    const int run = tile * p.nchunk + c;
                    const uint32_t e0 = __ldg(p.eoff + run), e1 = __ldg(p.eoff + run + 1), ne = e1 - e0;
                    const int brows = min(p.kcb, p.br - c * p.kcb);
                    const uint32_t xbytes = (uint32_t)brows * p.r * (BR_TM * 4);
                    mbar_arrive_expect_tx(&full[s], xbytes + ne * (uint32_t)C * 4u + ne + BR_CNT_BYTES + BR_WSTART_BYTES);
                    bulk_g2s(st, p.XT + ((size_t)mt * p.K + (size_t)c * p.kcb * p.r) * BR_TM, xbytes, &full[s]);
                    if (ne) {
                        bulk_g2s(st + p.xstage_bytes, p.val + (size_t)e0 * C, ne * (uint32_t)C * 4u, &full[s]);
                        bulk_g2s(st + p.xstage_bytes + p.val_stage_bytes, p.hdr + e0, ne, &full[s]);
                    }
                    bulk_g2s(st + p.xstage_bytes + p.val_stage_bytes + p.hdr_stage_bytes, p.cnt + (size_t)run * 256, BR_CNT_BYTES, &full[s]);
                    bulk_g2s(st + p.xstage_bytes + p.val_stage_bytes + p.hdr_stage_bytes + BR_CNT_BYTES, p.wstart + (size_t)run * BR_NWARP,
                             BR_WSTART_BYTES, &full[s]);
                }
            }
        }
        return;
    }

    // ===== consumers: 16 warps x 16 output columns, 4 rows per lane =====
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(BR_REGS_COMPUTE));
    constexpr int BPW = BR_CW / C;
    uint32_t it = 0;
    const int nvirt = p.full_units + (p.units - p.full_units) * p.sub;
    for (int v = blockIdx.x; v < nvirt; v += gridDim.x) {
        if (v < p.full_units) {
            ring_unit<C, BPW>(p, smem, stage_bytes, full, empty, it, v, 0, warp, lane);
        } else {
            const int q = v - p.full_units;
            if constexpr (BPW >= 2) {
                if (p.sub == 2) {
                    ring_unit<C, BPW / 2>(p, smem, stage_bytes, full, empty, it, p.full_units + (q >> 1), q & 1, warp, lane);
                    continue;
                }
            }
            ring_unit<C, BPW>(p, smem, stage_bytes, full, empty, it, p.full_units + q, 0, warp, lane);
        }
    }
}

// Units are dealt round-robin over the persistent CTAs, so `units % sms` left-over units cost a whole extra round on a few SMs
// while the rest idle (4096^3: 512 units on 148 SMs = 3.46 rounds, ncu: SMs active 85 % of the kernel).  Dealing the left-over
// units as two column slices each (8 instead of 16 columns per warp; per-column arithmetic and its order unchanged) turns that
// round into a half-length one whenever the slices still fit one round -- and doubles the CTAs of a problem smaller than the GPU.
void bcsr_ring_plan(int units, int sms, int bpw, int *full_units, int *sub) {
    const int rem = units % sms;
    static const bool off = getenv("TSG_BCSR_NO_SPLIT") != nullptr;
    if (off || bpw < 2 || rem == 0 || 2 * rem > sms) {
        *full_units = units;
        *sub = 1;
    } else {
        *full_units = units - rem;
        *sub = 2;
    }
}

template <int C>
static int launch_ring(const BcsrRingParams &p, size_t smem_bytes) {
    static std::atomic<unsigned long long> attr_done{0};  // one per instantiation
    TSG_TRY(once_per_device(attr_done, [] {
        TSG_CUDA(cudaFuncSetAttribute(k_bcsr_gemm_ring<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BR_SMEM_MAX));
        return (int)TSG_OK;
    }));
    const int nvirt = p.full_units + (p.units - p.full_units) * p.sub;
    const int grid = nvirt < num_sms() ? nvirt : num_sms();
    k_bcsr_gemm_ring<C><<<grid, BR_THREADS, smem_bytes, stream()>>>(p);
    TSG_KERNEL_CHECK("k_bcsr_gemm_ring");
    return TSG_OK;
}

int bcsr_gemm_ring(tsg_bcsr *W, const float *XT, const float *B, float a, int use_prelu, float *Y, int M, int N, int K, long long ldy,
                   int *handled) {
    *handled = 0;
    TSG_TRY(build_bstream(W));
    const BStream &bs = W->bs;
    if (!bs.built) return TSG_OK;  // outside the ring kernel's limits: the caller uses the plain kernel
    BcsrRingParams p;
    p.XT = XT; p.cnt = bs.cnt; p.wstart = bs.wstart; p.eoff = bs.eoff; p.hdr = bs.hdr; p.val = bs.val;
    p.B = B; p.Y = Y; p.ldy = ldy;
    p.M = M; p.N = N; p.K = K; p.r = W->r; p.br = W->br; p.ncov = W->bc * W->c;
    p.kcb = bs.kcb; p.nchunk = bs.nchunk; p.ntile = bs.ntile;
    p.units = ((M + BR_TM - 1) / BR_TM) * bs.ntile;
    bcsr_ring_plan(p.units, num_sms(), BR_CW / W->c, &p.full_units, &p.sub);
    p.a = a; p.use_prelu = use_prelu;
    p.xstage_bytes = (uint32_t)bs.kcb * W->r * BR_TM * 4;
    p.val_stage_bytes = (uint32_t)bs.max_run * W->c * 4;
    p.hdr_stage_bytes = (uint32_t)bs.max_run;
    const size_t smem_bytes = 2 * ring_stage_bytes(bs.kcb, W->r, W->c, bs.max_run) + 64;
    int rc;
    switch (W->c) {
        case 1: rc = launch_ring<1>(p, smem_bytes); break;
        case 2: rc = launch_ring<2>(p, smem_bytes); break;
        case 4: rc = launch_ring<4>(p, smem_bytes); break;
        case 8: rc = launch_ring<8>(p, smem_bytes); break;
        default: rc = launch_ring<16>(p, smem_bytes); break;
    }
    if (rc == TSG_OK) *handled = 1;
    return rc;
}

}  // namespace tsg

// diagnostic (tests/test_plan.py): how the ring kernel deals `units` over `sms` persistent CTAs
extern "C" void tsg_dbg_bcsr_ring_plan(int units, int sms, int bpw, int *full_units, int *sub) {
    tsg::bcsr_ring_plan(units, sms, bpw, full_units, sub);
}
