// gemm_tcsc.cu -- the multiplication-free sparse ternary GEMM  Y = [PReLU](X*W + B)  on sm_100a.
//
// Replaces tcsc_sgemm_basic / _optimized / _prelu_basic / _prelu_optimized_separate / _prelu_optimized_onthego
// (sparse/tcsc.c:69-165,179-275) and sparseGEMM<float> / sparseGEMM_PReLU<float> (SparseGEMM.h:104-119,151-168).
//
// Tiled kernel (M >= TSG_SKINNY_M).  One persistent CTA per SM walks work units (128 rows of X) x (TN columns of W):
//   * lanes own rows: lane l of every warp holds rows l, l+32, l+64, l+96 of the 128-row tile, a warp owns CW columns, so each
//     thread keeps CW x 4 accumulators in registers and the k index of every non-zero is warp-uniform;
//   * X is pre-transposed once per call into K-major 128-row tiles XT[mtile][k][128] (transpose_x_tiles), so that a
//     kc-row chunk of a tile is ONE contiguous block: a producer thread streams chunk after chunk into a two-stage
//     shared-memory ring with bulk async copies (cp.async.bulk, TMA engine) completing on mbarriers, together with
//     the tile's slice of the private gather stream (ktformat.cu);
//   * per non-zero a warp issues one conflict-free LDS.128 (32 lanes x 16 B = the tile's row k) and four FADDs per
//     lane -- no multiplies, no tensor cores.  The binding resource is the shared-memory crossbar (128 B/clk/SM =
//     one warp-wide add per clock per SM = 25 % of the FP32 add peak; profiles/microbench/);
//   * bias, PReLU and the store of Y are fused into the epilogue; Y may be a column slab of a wider matrix (ldy).
// Summation order: a unit makes two sweeps over the K chunks, first the +1 entries, then the -1 entries, each in
// ascending k, with a single accumulator per output -- the exact sequence of fp32 roundings of the reference
// function selected by `order` (DESIGN.md "Summation order").
//
// Skinny kernel (M < TSG_SKINNY_M, the decode shape): HBM/L2-bound on the index stream.  One warp per column, lanes
// stride over the column's non-zeros (coalesced index loads), gather up to 8 rows of X per index from a K-major copy
// of X, tree-reduce across the warp.
#include <vector>

#include "tsg_internal.h"
#include "tsg_ptx.cuh"

namespace tsg {

constexpr int TM = 128;     // rows of X per tile
constexpr int NWARP = 16;   // compute warps per CTA
constexpr int NTHREADS = (NWARP + 4) * 32;  // 4 compute warpgroups + 1 producer warpgroup (register re-balancing is per warpgroup)
// register re-balancing happens inside the CTA's own pool: 4 x 128 x (112 - 96) <= 128 x (96 - 24), else setmaxnreg.inc never returns
constexpr int REGS_COMPUTE = 112, REGS_PRODUCER = 24;
constexpr int WOFF_BYTES = 160;  // (256/8 + 1) offsets, rounded up to 16 B
constexpr int CNT_BYTES = 256;
constexpr int TILE_PITCH = 260;  // words per row of the staged output tile (fused TMA epilogue)
// multicast epilogue (fused_tma == 3): the spare warps of the producer warpgroup re-read every finished tile from the local Y
// (L2) and write it once to the multicast mapping
constexpr int MC_DRAIN_WARPS = 3;
constexpr int REGS_PRODUCER_MC = 32;  // the drain loop needs a few more registers than the TMA producer (still inside the re-balancing budget)

struct GemmParams {
    const float *XT;
    const uint8_t *cnt;
    const uint32_t *woff;
    const uint32_t *body;
    const float *B;
    float *Y;
    long long ldy;
    int M, N, K;
    int kc, nchunk, ncols_pad, ngroup;
    int nstage;                  // depth of the shared-memory ring (2 or 3)
    int mtiles, ntiles;          // ntiles = number of 256-column tiles
    int units_full, units_total, sub;  // unit decomposition, see decode_unit()
    float a;
    int use_prelu, order;
    uint32_t xstage_bytes, body_stage_bytes;
    uint32_t bar_off;  // byte offset of the mbarriers: behind the stage ring (and behind the staged output tile if that is larger)
    // column-partitioned multi-GPU path: the epilogue additionally stores the finished slab into every peer's Y
    // (NVLink peer mappings, already offset to this rank's first column); 0 = single GPU
    int npeer;
    float *peerY[TSG_MAX_PEERS];
    // progress counters, one per 128-row tile: every compute warp bumps done[mt] once its part of a unit is stored, so
    // that stream-ordered peer copies (dist.cu, mode 2) can start on finished row blocks while the kernel still runs
    // fused all-gather through the TMA engine (dist.cu mode 3): the finished tile is staged in shared memory and every
    // 128-row x tn-column row segment goes to the local Y and to every peer with one bulk async store per row
    int fused_tma;
    unsigned int *done;
    int ngroups;       // row tiles [gbound[g], gbound[g+1]) form progress group g (at most 8 groups)
    int gbound[9];
    uint32_t tile_off;  // TILE_SEP instantiation: byte offset of the output tile (behind the stage ring)
};

// ---- gather-add over one chunk for the columns of a warp ---------------------------------------------------------------
// Accumulators are kept as packed fp32x2 pairs (rows l,l+32 and l+64,l+96 of the tile) so that one FADD2 (add.f32x2,
// sm_100) retires two adds per issue slot; each component is an ordinary IEEE fp32 add, so the roundings are unchanged.
// One 32-bit word of the gather stream holds up to four non-zeros (bytes k+1); a 0 byte is padding: its load and adds
// are predicated off, so it costs issue slots but no shared-memory bandwidth.  Adds are applied in stream order.
template <bool NEG>
__device__ __forceinline__ void gather_word(float2 &a01, float2 &a23, uint32_t word, uint32_t xbase) {
    // Entry bytes hold k+1 (1..kc), 0 is padding.  Masking the byte in place gives the value and the "not padding"
    // predicate in one LOP3; the row's byte offset (k+1)*512 is then one IMAD / IMAD.HI on the still-shifted field
    // (xbase already has the -512 folded in), so an entry costs 5 instructions: LOP3, IMAD, LDS.128, 2 x FADD2.
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const uint32_t f = word & (0xFFu << (8 * e));
        if (f != 0u) {  // warp-uniform: compiles to predicated LDS.128 + FADD2, no branch
            uint32_t addr;
            if (e == 0) addr = f * (TM * 4) + xbase;
            else if (e == 1) addr = f * (TM * 4 / 256) + xbase;
            else if (e == 2) addr = __umulhi(f, 1u << 25) + xbase;   // (f >> 16) * 512 = f >> 7
            else addr = __umulhi(f, 1u << 17) + xbase;               // (f >> 24) * 512 = f >> 15
            float4 x;  // 32-bit shared-window address: no generic->shared conversion per load
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w) : "r"(addr));
            if (NEG) {
                a01 = __fadd2_rn(a01, make_float2(-x.x, -x.y));
                a23 = __fadd2_rn(a23, make_float2(-x.z, -x.w));
            } else {
                a01 = __fadd2_rn(a01, make_float2(x.x, x.y));
                a23 = __fadd2_rn(a23, make_float2(x.z, x.w));
            }
        }
    }
}

// ---- thread-block cluster helpers (mid-size M: the CTAs that share a row tile of X receive each chunk by ONE multicast copy) ----
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// global -> the same shared-memory offset of every CTA in `mask`, completing `bytes` on the mbarrier at the same offset of each
__device__ __forceinline__ void bulk_g2s_multicast(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar, uint16_t mask) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
                     smem_addr(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_addr(bar)), "h"(mask)
                 : "memory");
}
// arrive on the mbarrier at the same shared-memory offset of CTA `rank` of this cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t *bar, uint32_t rank) {
    asm volatile("{ .reg .b32 ra; mapa.shared::cluster.u32 ra, %0, %1; mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra]; }" ::"r"(
                     smem_addr(bar)),
                 "r"(rank)
                 : "memory");
}

constexpr int CWMAX = 16;  // columns per warp in a full (256-column) tile; tail units use 8 or 4 (runtime `cw`)

template <bool NEG>
__device__ __forceinline__ void gather_chunk(float2 (&acc)[CWMAX][2], uint32_t xbase, const uint8_t *__restrict__ cnt_s,
                                             const uint32_t *__restrict__ woff_s, const uint32_t *__restrict__ body_s, int warp, int cw) {
    // word offset of this warp's first column inside the staged body slice
    const int col0 = warp * cw;
    uint32_t off = woff_s[col0 >> 3] - woff_s[0];
    if (const int r = col0 & 7) {  // cw < 8: the warp starts inside an 8-column group -> skip the lists of the r columns before it
        const uint2 prev = *reinterpret_cast<const uint2 *>(cnt_s + (col0 & ~7));
        const uint32_t lo = (r >= 4) ? prev.x : (prev.x & ((1u << (8 * r)) - 1u));
        const uint32_t hi = (r > 4) ? (prev.y & ((1u << (8 * (r - 4))) - 1u)) : 0u;
        off += 4u * (__vsadu4(lo, 0u) + __vsadu4(hi, 0u));
    }
    uint32_t cwd[CWMAX / 4];
#pragma unroll
    for (int i = 0; i < CWMAX / 4; ++i) cwd[i] = (4 * i < cw) ? *reinterpret_cast<const uint32_t *>(cnt_s + ((col0 + 4 * i) & ~3)) : 0u;
    if (cw < 4) cwd[0] = (cwd[0] >> (8 * (col0 & 3))) & ((1u << (8 * cw)) - 1u);  // cw == 2: the warp's two counts sit inside a word
    const uint4 *qp = reinterpret_cast<const uint4 *>(body_s + off);  // every list starts on a 16-byte boundary
    // one quad of look-ahead: the lists of a warp are contiguous in the stream, so the next quad is fetched while the
    // current one is being gathered (reading one quad past the warp's region stays inside the stage buffer)
    uint4 nxt = *qp;
#pragma unroll
    for (int j = 0; j < CWMAX; ++j) {
        const int nq = (cwd[j >> 2] >> (8 * (j & 3))) & 0xFF;  // 0 for j >= cw
#pragma unroll 1
        for (int i = 0; i < nq; ++i) {
            const uint4 w = nxt;  // one uniform-address LDS.128 = up to 16 non-zeros
            nxt = *++qp;
            gather_word<NEG>(acc[j][0], acc[j][1], w.x, xbase);
            if (w.y != 0u) {  // entries are packed from the front: an all-padding word ends the list
                gather_word<NEG>(acc[j][0], acc[j][1], w.y, xbase);
                if (w.z != 0u) {
                    gather_word<NEG>(acc[j][0], acc[j][1], w.z, xbase);
                    if (w.w != 0u) gather_word<NEG>(acc[j][0], acc[j][1], w.w, xbase);
                }
            }
        }
    }
}

// work unit u -> (row tile, first column, columns per warp).  Units [0, units_full) are full 256-column tiles; the
// remaining full tiles are cut into `sub` narrower units each so that the last round of the persistent grid is short.
struct Unit {
    int mt, n0, cw;
};
__host__ __device__ __forceinline__ Unit decode_unit_raw(int units_full, int sub, int ntiles, int u) {
    int fu = u, part = 0, cw = CWMAX;
    if (u >= units_full) {
        const int v = u - units_full;
        fu = units_full + v / sub;
        part = v % sub;
        cw = CWMAX / sub;
    }
    Unit r;
    r.mt = fu / ntiles;
    r.cw = cw;
    r.n0 = (fu % ntiles) * (CWMAX * NWARP) + part * (cw * NWARP);
    return r;
}
__device__ __forceinline__ Unit decode_unit(const GemmParams &p, int u) { return decode_unit_raw(p.units_full, p.sub, p.ntiles, u); }

// TILE_SEP (dist mode 4): the staged output tile has shared memory of its own instead of overlaying the stage ring, so the
// producer never waits for the bulk stores and a unit's stores drain while the next unit is being gathered.
// MC (dist mode 5): the epilogue hands 32-row slices of the finished tile to three drain warps through two staging buffers of
// their own; nothing overlays the stage ring and no CTA-wide barrier is left in the unit loop.
// FAST (TSG_ORDER_FAST): a stage carries the +1 AND the -1 slice of a chunk and one sweep over K does both; a template
// parameter so that the exact-order instantiations carry none of its code.
// CL > 1 (mid-size M, fewer tiles than SMs): a thread-block cluster of CL CTAs shares one 128-row tile of X and splits its 256
// columns.  The cluster's leader fetches every X chunk ONCE with a multicast bulk copy into all CL shared memories (the chunk
// is what every unit re-streams: L2 -> SM traffic falls by CL); the leader waits for all CL x 16 consumer warps (remote
// mbarrier arrivals) before it overwrites a stage.  Stream slices stay per CTA.
template <bool TILE_SEP, bool MC = false, bool FAST = false, int CL = 1>
__global__ void __launch_bounds__(NTHREADS, 1) k_tcsc_gemm(const GemmParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    // a stage = X chunk + one stream area (exact orders: the +1 OR the -1 lists of the chunk) or two (TSG_ORDER_FAST: both)
    const uint32_t area_bytes = p.body_stage_bytes + CNT_BYTES + WOFF_BYTES;
    constexpr bool fast = FAST;
    const uint32_t stage_bytes = p.xstage_bytes + (fast ? 2u : 1u) * area_bytes;
    constexpr int npass = fast ? 1 : 2;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + p.bar_off);
    uint64_t *empty = full + 3;
    uint64_t *epi = full + 6;  // fused epilogue: consumers -> producer "the output tile has left shared memory"
    unsigned int *tiles_done = reinterpret_cast<unsigned int *>(full + 7);  // multicast epilogue: compute-warp arrivals, monotonic
    uint64_t *xempty = full + 8;  // cluster variant, meaningful on the leader: stage s released by the consumers of ALL CTAs
    const uint32_t nstage = (uint32_t)p.nstage;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // work distribution: CL == 1 walks the unit list; a cluster walks the tiles and its CTAs take the CL column parts of each
    const int crank = (CL > 1) ? (int)(blockIdx.x % CL) : 0;
    const int u_first = (CL > 1) ? (int)(blockIdx.x / CL) : (int)blockIdx.x;
    const int u_step = (CL > 1) ? (int)(gridDim.x / CL) : (int)gridDim.x;
    const int u_end = (CL > 1) ? p.mtiles * p.ntiles : p.units_total;
    auto unit_of = [&](int u) {
        if constexpr (CL > 1) {
            Unit r;
            r.mt = u / p.ntiles;
            r.cw = CWMAX / CL;
            r.n0 = (u % p.ntiles) * (CWMAX * NWARP) + crank * (r.cw * NWARP);
            return r;
        } else {
            return decode_unit(p, u);
        }
    };

    if (tid == 0) {
        for (int i = 0; i < 3; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], NWARP);
            mbar_init(&xempty[i], NWARP * CL);
        }
        mbar_init(epi, NWARP);
        *tiles_done = 0u;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if constexpr (CL > 1) cluster_sync_all();  // every CTA's barriers exist before anybody arrives on them remotely

    if (warp >= NWARP) {
        // ===== producer warpgroup: hands its registers to the compute warpgroups; one thread feeds the two-stage ring =====
        if (MC || CL > 1) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_PRODUCER_MC));
        else asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_PRODUCER));
        if (MC && warp > NWARP) {
            // ===== drain warps (multicast epilogue, dist.cu mode 5): once the 16 compute warps have stored a unit's tile into the
            // local Y (and fenced), re-read it from L2 and write it ONCE with multimem.st to the NVSwitch multicast mapping of Y
            // (peerY[0]); the switch replicates every store into the Y of every rank.  A warp moves 512 contiguous bytes of a
            // row per instruction.  Fabric back-pressure stalls these three warps, never the gather warps, and nothing of the
            // stage ring is borrowed: the gather loop runs exactly as on one GPU =====
            const int dwarp = warp - NWARP - 1;
            uint32_t uidx = 0;
            for (int u = blockIdx.x; u < p.units_total; u += gridDim.x, ++uidx) {
                const Unit un = decode_unit(p, u);
                const int ncol = min(un.cw * NWARP, p.N - un.n0);
                const int nvec = ncol > 0 ? (ncol >> 2) : 0;
                const int m0 = un.mt * TM;
                const int rows = min(TM, p.M - m0);
                while (*reinterpret_cast<volatile unsigned int *>(tiles_done) < (uidx + 1u) * NWARP) __nanosleep(256);
                __threadfence();  // acquire side of the compute warps' fence + arrival
                const float *src = p.Y + (size_t)m0 * p.ldy + un.n0;
                float *mc = p.peerY[0] + (size_t)m0 * p.ldy + un.n0;
                for (int r = dwarp; r < rows; r += MC_DRAIN_WARPS) {
                    const size_t ro = (size_t)r * p.ldy;
                    for (int c = lane; c < nvec; c += 32) {
                        float4 x;
                        asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w) : "l"(src + ro + 4 * c));
                        asm volatile("multimem.st.weak.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc + ro + 4 * c), "f"(x.x), "f"(x.y), "f"(x.z), "f"(x.w) : "memory");
                    }
                }
            }
            return;
        }
        if (warp == NWARP && lane == 0) {
            uint32_t s = 0, ph = 0, uidx = 0;  // ring position: stage s, phase parity ph of its barriers
            for (int u = u_first; u < u_end; u += u_step, ++uidx) {
                const Unit un = unit_of(u);
                const int tn = un.cw * NWARP;
                // fused epilogue: the previous unit's output tile overlays the stage ring until its bulk stores have read it
                if (!TILE_SEP && !MC && p.fused_tma && uidx > 0) mbar_wait_relaxed(epi, (uidx - 1) & 1u);
                const uint32_t woff_copy = (uint32_t)(((tn / 8 + 1) * 4 + 15) & ~15);
                for (int pass = 0; pass < npass; ++pass) {
                    for (int c = 0; c < p.nchunk; ++c) {
                        mbar_wait_relaxed(&empty[s], ph ^ 1u);
                        if (CL > 1 && crank == 0) mbar_wait_relaxed(&xempty[s], ph ^ 1u);  // ... and by every other CTA of the cluster
                        uint8_t *st = smem + (size_t)s * stage_bytes;
                        const int rows = min(p.kc, p.K - c * p.kc);
                        const uint32_t xbytes = (uint32_t)rows * (TM * 4);
                        // stream slices of this chunk: plane `pass*nchunk + c` (exact orders), or the +1 and the -1 plane (fast order)
                        size_t gidx[2];
                        uint32_t w0[2], bbytes[2], total = xbytes;
                        const int nareas = fast ? 2 : 1;
                        for (int q = 0; q < nareas; ++q) {
                            const int plane = (fast ? q : pass) * p.nchunk + c;
                            gidx[q] = (size_t)plane * p.ngroup + (un.n0 >> 3);
                            w0[q] = __ldg(p.woff + gidx[q]);
                            bbytes[q] = (__ldg(p.woff + gidx[q] + tn / 8) - w0[q]) * 4u;
                            total += bbytes[q] + tn + woff_copy;
                        }
                        mbar_arrive_expect_tx(&full[s], total);
                        if constexpr (CL > 1) {  // the leader's copy lands in every CTA and completes xbytes on every CTA's full[s]
                            if (crank == 0) bulk_g2s_multicast(st, p.XT + ((size_t)un.mt * p.K + (size_t)c * p.kc) * TM, xbytes, &full[s], (uint16_t)((1u << CL) - 1u));
                        } else {
                            bulk_g2s(st, p.XT + ((size_t)un.mt * p.K + (size_t)c * p.kc) * TM, xbytes, &full[s]);
                        }
                        for (int q = 0; q < nareas; ++q) {
                            const int plane = (fast ? q : pass) * p.nchunk + c;
                            uint8_t *area = st + p.xstage_bytes + (size_t)q * area_bytes;
                            if (bbytes[q]) bulk_g2s(area, p.body + w0[q], bbytes[q], &full[s]);
                            bulk_g2s(area + p.body_stage_bytes, p.cnt + (size_t)plane * p.ncols_pad + un.n0, tn, &full[s]);
                            bulk_g2s(area + p.body_stage_bytes + CNT_BYTES, p.woff + gidx[q], woff_copy, &full[s]);
                        }
                        if (++s == nstage) { s = 0; ph ^= 1u; }
                    }
                }
            }
        }
        if constexpr (CL > 1) cluster_sync_all();  // nobody leaves while a neighbour may still arrive on its barriers
        return;
    }

    // ===== consumers: 16 warps x cw columns, 4 rows per lane =====
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_COMPUTE));
    float2 acc[CWMAX][2];
    uint32_t s = 0, ph = 0;  // ring position of this warp
    const bool vec_ok = ((p.ldy & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.Y) & 15) == 0);
    for (int u = u_first; u < u_end; u += u_step) {
        const Unit un = unit_of(u);
        const int cw = un.cw;
        const int nbase = un.n0 + warp * cw;
        const int mbase = un.mt * TM + lane;  // lane l holds rows l, l+32, l+64, l+96 of the tile (XT position 4l+v)
        // the bias is re-read (L1/L2 hit) where it is needed instead of living in registers across the gather loops
        auto bias_of = [&](int j) { return (nbase + j < p.N) ? __ldg(p.B + nbase + j) : 0.f; };
#pragma unroll
        for (int j = 0; j < CWMAX; ++j) {
            const float init = (p.order == TSG_ORDER_BIAS_FIRST && j < cw) ? bias_of(j) : 0.f;  // tcsc.c:84 vs tcsc.c:149
            acc[j][0] = make_float2(init, init);
            acc[j][1] = acc[j][0];
        }
        for (int pass = 0; pass < npass; ++pass) {
            if (pass == 1 && p.order == TSG_ORDER_SPLIT) {
                // tcsc.c:125,138: Y = B + acc_pos is rounded and parked in Y, acc_neg starts from 0
#pragma unroll
                for (int j = 0; j < CWMAX; ++j) {
                    if (j < cw && nbase + j < p.N) {
                        const float b = bias_of(j);
                        const float t4[4] = {b + acc[j][0].x, b + acc[j][0].y, b + acc[j][1].x, b + acc[j][1].y};
#pragma unroll
                        for (int v = 0; v < 4; ++v)
                            if (mbase + 32 * v < p.M) p.Y[(size_t)(mbase + 32 * v) * p.ldy + nbase + j] = t4[v];
                    }
                    acc[j][0] = make_float2(0.f, 0.f);
                    acc[j][1] = acc[j][0];
                }
            }
            const bool neg = (pass == 1) && (p.order != TSG_ORDER_SPLIT);
            for (int c = 0; c < p.nchunk; ++c) {
                mbar_wait(&full[s], ph);
                const uint8_t *st = smem + (size_t)s * stage_bytes;
                const uint32_t xbase = smem_addr(st) + lane * 16 - TM * 4;  // entries are k+1: fold the -512 in here
                const uint32_t *body_s = reinterpret_cast<const uint32_t *>(st + p.xstage_bytes);
                const uint8_t *cnt_s = st + p.xstage_bytes + p.body_stage_bytes;
                const uint32_t *woff_s = reinterpret_cast<const uint32_t *>(cnt_s + CNT_BYTES);
                if (neg) gather_chunk<true>(acc, xbase, cnt_s, woff_s, body_s, warp, cw);
                else gather_chunk<false>(acc, xbase, cnt_s, woff_s, body_s, warp, cw);
                if constexpr (FAST) {  // the chunk's -1 entries while the chunk is still resident: X streams through shared memory once
                    const uint8_t *area = st + p.xstage_bytes + area_bytes;
                    gather_chunk<true>(acc, xbase, area + p.body_stage_bytes, reinterpret_cast<const uint32_t *>(area + p.body_stage_bytes + CNT_BYTES),
                                       reinterpret_cast<const uint32_t *>(area), warp, cw);
                }
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(&empty[s]);
                    if constexpr (CL > 1) mbar_arrive_remote(&xempty[s], 0u);
                }
                if (++s == nstage) { s = 0; ph ^= 1u; }
            }
        }
        // ---- fused epilogue: bias, PReLU, store (and, on the multi-GPU path, the same store into every peer's Y) ----
        if constexpr (TILE_SEP) {
            // the previous unit's bulk stores must have read the tile before it is overwritten
            if (tid < TM) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            asm volatile("bar.sync 2, 512;" ::: "memory");
        } else {
            if (p.fused_tma) asm volatile("bar.sync 2, 512;" ::: "memory");  // every warp is done reading the stage ring
        }
        uint8_t *const tile_base = TILE_SEP ? smem + p.tile_off : smem;
        const bool full_vec = vec_ok && (cw % 4 == 0) && (nbase + cw <= p.N);
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            const int m = mbase + 32 * v;
            if (m >= p.M && !p.fused_tma) continue;
            float *yrow = p.Y + (size_t)m * p.ldy + nbase;
            float out[CWMAX];
#pragma unroll
            for (int j = 0; j < CWMAX; ++j) {
                const float2 pr = acc[j][v >> 1];
                const float av = (v & 1) ? pr.y : pr.x;
                float y = av;
                if (j < cw) {
                    if (p.order == TSG_ORDER_BIAS_LAST || p.order == TSG_ORDER_FAST) y = av + bias_of(j);      // tcsc.c:161
                    else if (p.order == TSG_ORDER_SPLIT) y = ((nbase + j < p.N && m < p.M) ? yrow[j] : 0.f) - av;          // tcsc.c:138
                    if (p.use_prelu) y = (y < 0.0f) ? p.a * y : y;                                              // tcsc.c:162
                }
                out[j] = y;
            }
            if (p.fused_tma) {
                // stage this lane's cw values of row (lane + 32 v) into the output tile (row pitch 260 words: consecutive
                // lanes = consecutive rows land 4 banks apart, so each quarter-warp float4 store covers all 32 banks)
                float *trow = reinterpret_cast<float *>(tile_base) + (size_t)(lane + 32 * v) * TILE_PITCH + warp * cw;
#pragma unroll
                for (int j = 0; j < CWMAX; j += 4)
                    if (j < cw) *reinterpret_cast<float4 *>(trow + j) = make_float4(out[j], out[j + 1], out[j + 2], out[j + 3]);
                continue;
            }
            for (int q = -1; q < p.npeer; ++q) {  // q == -1: the local Y
                float *row = (q < 0) ? yrow : p.peerY[q] + (size_t)m * p.ldy + nbase;
                if (full_vec) {
#pragma unroll
                    for (int j = 0; j < CWMAX; j += 4)
                        if (j < cw) *reinterpret_cast<float4 *>(row + j) = make_float4(out[j], out[j + 1], out[j + 2], out[j + 3]);
                } else {
#pragma unroll
                    for (int j = 0; j < CWMAX; ++j)
                        if (j < cw && nbase + j < p.N) row[j] = out[j];
                }
            }
        }
        if (p.fused_tma) {
            // tile complete in shared memory -> one bulk async store (TMA engine) per row and destination
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("bar.sync 2, 512;" ::: "memory");
            if (tid < TM) {
                const int m = un.mt * TM + tid;
                const int ncol = min(cw * NWARP, p.N - un.n0);
                if (m < p.M && ncol > 0) {
                    const uint32_t src = smem_addr(reinterpret_cast<float *>(tile_base) + (size_t)tid * TILE_PITCH);
                    const uint32_t bytes = (uint32_t)ncol * 4u;
                    for (int q = -1; q < p.npeer; ++q) {
                        float *dst = ((q < 0) ? p.Y : p.peerY[q]) + (size_t)m * p.ldy + un.n0;
                        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
                    }
                }
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                if constexpr (!TILE_SEP) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // shared memory may be overwritten again
            }
            if constexpr (!TILE_SEP) {
                asm volatile("bar.sync 2, 512;" ::: "memory");
                if (lane == 0) mbar_arrive(epi);
            }
        }
        if constexpr (MC) {  // this warp's share of the tile is in the local Y: let the drain warps send it
            __threadfence();
            __syncwarp();
            if (lane == 0) atomicAdd(tiles_done, 1u);
        }
        if (p.done) {  // publish: this warp's share of unit (mt, n0) is in memory
            __threadfence_system();
            __syncwarp();
            if (lane == 0) {
                int g = 0;
                while (g + 1 < p.ngroups && un.mt >= p.gbound[g + 1]) ++g;
                atomicAdd(p.done + g, 1u);
            }
        }
    }
    if (!MC && p.fused_tma && tid < TM) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // all row stores have been performed
    if constexpr (CL > 1) cluster_sync_all();
}

// ---- X (M x K row-major) -> XT[mtile][k][128], rows >= M zero ----------------------------------------------------------
__global__ void __launch_bounds__(256) k_transpose_x(const float *__restrict__ X, float *__restrict__ XT, int M, int K) {
    // one CTA = one 128-row tile x 32 k: coalesced reads along k, then every thread writes a float4 holding rows
    // l, l+32, l+64, l+96 of one k (position 4l..4l+3 of the tile row) -> a warp writes 512 contiguous bytes
    __shared__ float tile[TM][33];
    const int k0 = blockIdx.x * 32, mt = blockIdx.y, m0 = mt * TM;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
    for (int i = 0; i < TM / 8; ++i) {
        const int r = ty + 8 * i, m = m0 + r, k = k0 + tx;
        tile[r][tx] = (m < M && k < K) ? __ldg(X + (size_t)m * K + k) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int kk = ty + 8 * i, k = k0 + kk;
        if (k < K) {
            const float4 v = make_float4(tile[tx][kk], tile[tx + 32][kk], tile[tx + 64][kk], tile[tx + 96][kk]);
            *reinterpret_cast<float4 *>(XT + ((size_t)mt * K + k) * TM + 4 * tx) = v;
        }
    }
}

int transpose_x_tiles(const float *X, float *XT, int M, int K, int mtiles_min) {
    int mtiles = (M + TM - 1) / TM;
    if (mtiles < mtiles_min) mtiles = mtiles_min;  // tiles beyond M are written as zeros (the kernel pads rows >= M)
    dim3 grid((K + 31) / 32, mtiles);
    k_transpose_x<<<grid, 256, 0, stream()>>>(X, XT, M, K);
    TSG_KERNEL_CHECK("k_transpose_x");
    return TSG_OK;
}

// =====================================================================================================================
// skinny kernel
// =====================================================================================================================
constexpr int SK_MT = 8;  // rows of X handled together

// X rows [m0, m0+8) -> XS[group][k][8] (zero padded)
__global__ void k_skinny_pack_x(const float *__restrict__ X, float *__restrict__ XS, int M, int K) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int g = blockIdx.y;
    if (k >= K) return;
    float v[SK_MT];
#pragma unroll
    for (int i = 0; i < SK_MT; ++i) {
        const int m = g * SK_MT + i;
        v[i] = (m < M) ? __ldg(X + (size_t)m * K + k) : 0.f;
    }
    float4 *dst = reinterpret_cast<float4 *>(XS + ((size_t)g * K + k) * SK_MT);
    dst[0] = make_float4(v[0], v[1], v[2], v[3]);
    dst[1] = make_float4(v[4], v[5], v[6], v[7]);
}

template <int MT>
__device__ __forceinline__ void skinny_add(float (&acc)[MT], const float *__restrict__ px, float sign) {
    const float4 a = __ldg(reinterpret_cast<const float4 *>(px));
    acc[0] += sign * a.x;
    if (MT > 1) acc[1 % MT] += sign * a.y;
    if (MT > 2) { acc[2 % MT] += sign * a.z; acc[3 % MT] += sign * a.w; }
    if (MT > 4) {
        const float4 b = __ldg(reinterpret_cast<const float4 *>(px) + 1);
        acc[4 % MT] += sign * b.x; acc[5 % MT] += sign * b.y; acc[6 % MT] += sign * b.z; acc[7 % MT] += sign * b.w;
    }
}

// lanes stride over the list; four independent index loads, then four independent gathers per trip (the list is short:
// latency, not bandwidth, is what a warp sees, so keep several loads in flight)
template <int MT>
__device__ __forceinline__ void skinny_accumulate(float (&acc)[MT], const float *__restrict__ xs, const int *__restrict__ idx, int lo,
                                                  int hi, int lane, float sign) {
    int t = lo + lane;
    for (; t + 96 < hi; t += 128) {
        const int k0 = __ldg(idx + t), k1 = __ldg(idx + t + 32), k2 = __ldg(idx + t + 64), k3 = __ldg(idx + t + 96);
        skinny_add<MT>(acc, xs + (size_t)k0 * SK_MT, sign);
        skinny_add<MT>(acc, xs + (size_t)k1 * SK_MT, sign);
        skinny_add<MT>(acc, xs + (size_t)k2 * SK_MT, sign);
        skinny_add<MT>(acc, xs + (size_t)k3 * SK_MT, sign);
    }
    int kk[3];
    int n = 0;
#pragma unroll
    for (int i = 0; i < 3; ++i)
        if (t + 32 * i < hi) { kk[i] = __ldg(idx + t + 32 * i); n = i + 1; }
#pragma unroll
    for (int i = 0; i < 3; ++i)
        if (i < n) skinny_add<MT>(acc, xs + (size_t)kk[i] * SK_MT, sign);
}

// M <= 2: no repacking, gather straight from the row-major X (one row = one K-vector)
__global__ void __launch_bounds__(256) k_tcsc_skinny_direct(const float *__restrict__ X, const int *__restrict__ csp, const int *__restrict__ csn,
                                                            const int *__restrict__ rip, const int *__restrict__ rin, const float *__restrict__ B,
                                                            float a, int use_prelu, float *__restrict__ Y, long long ldy, int M, int N, int K) {
    const int lane = threadIdx.x & 31;
    const int wglobal = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const float *x1 = X + (M > 1 ? (size_t)K : 0);
    for (int n = wglobal; n < N; n += nwarps) {
        float a0 = 0.f, a1 = 0.f;
        const int p0 = __ldg(csp + n), p1 = __ldg(csp + n + 1), q0 = __ldg(csn + n), q1 = __ldg(csn + n + 1);
        for (int t = p0 + lane; t < p1; t += 32) {
            const int k = __ldg(rip + t);
            a0 += __ldg(X + k);
            a1 += __ldg(x1 + k);
        }
        for (int t = q0 + lane; t < q1; t += 32) {
            const int k = __ldg(rin + t);
            a0 -= __ldg(X + k);
            a1 -= __ldg(x1 + k);
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            a0 += __shfl_xor_sync(0xffffffffu, a0, d);
            a1 += __shfl_xor_sync(0xffffffffu, a1, d);
        }
        if (lane == 0) {
            const float b = __ldg(B + n);
            float y = a0 + b;
            if (use_prelu) y = (y < 0.0f) ? a * y : y;
            Y[n] = y;
            if (M > 1) {
                y = a1 + b;
                if (use_prelu) y = (y < 0.0f) ? a * y : y;
                Y[(size_t)ldy + n] = y;
            }
        }
    }
}

template <int MT>
__global__ void __launch_bounds__(256) k_tcsc_skinny(const float *__restrict__ XS, const int *__restrict__ csp, const int *__restrict__ csn,
                                                     const int *__restrict__ rip, const int *__restrict__ rin, const float *__restrict__ B,
                                                     float a, int use_prelu, float *__restrict__ Y, long long ldy, int M, int N, int K) {
    const int lane = threadIdx.x & 31;
    const int wglobal = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const int g = blockIdx.y;
    const float *xs = XS + (size_t)g * K * SK_MT;
    for (int n = wglobal; n < N; n += nwarps) {
        float acc[MT];
#pragma unroll
        for (int i = 0; i < MT; ++i) acc[i] = 0.f;
        skinny_accumulate<MT>(acc, xs, rip, __ldg(csp + n), __ldg(csp + n + 1), lane, 1.0f);
        skinny_accumulate<MT>(acc, xs, rin, __ldg(csn + n), __ldg(csn + n + 1), lane, -1.0f);
#pragma unroll
        for (int i = 0; i < MT; ++i) {
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], d);
        }
        if (lane == 0) {
            const float b = __ldg(B + n);
#pragma unroll
            for (int i = 0; i < MT; ++i) {
                const int m = g * SK_MT + i;
                if (m < M) {
                    float y = acc[i] + b;
                    if (use_prelu) y = (y < 0.0f) ? a * y : y;
                    Y[(size_t)m * ldy + n] = y;
                }
            }
        }
    }
}

static thread_local int g_force_kernel = 0;
static thread_local int g_profile = 0;
static thread_local std::vector<cudaEvent_t> g_prof_events;

void profile_mark(bool begin) {
    if (!g_profile) return;
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    cudaEventRecord(e, stream());
    g_prof_events.push_back(e);
    (void)begin;
}

static int launch_skinny(tsg_tcsc *W, const float *X, const float *B, float a, int use_prelu, float *Y, int M, int N, int K, long long ldy) {
    static const int env_decode = getenv("TSG_DECODE") ? atoi(getenv("TSG_DECODE")) : 1;  // 0: the first skinny kernels (A/B runs)
    if (env_decode) {  // X rows in shared memory (decode_tcsc.cu); not handled when a row of X does not fit
        int handled = 0;
        TSG_TRY(tcsc_decode(W, X, B, a, use_prelu, Y, M, N, K, ldy, &handled));
        if (handled) return TSG_OK;
    }
    cudaStream_t st = stream();
    const int groups = (M + SK_MT - 1) / SK_MT;
    const int warps_per_cta = 8;
    int ctas = (N + warps_per_cta - 1) / warps_per_cta;
    const int max_ctas = num_sms() * 8;
    if (ctas > max_ctas) ctas = max_ctas;
    if (ctas < 1) ctas = 1;
    if (M <= 2) {
        k_tcsc_skinny_direct<<<ctas, 256, 0, st>>>(X, W->csp, W->csn, W->rip, W->rin, B, a, use_prelu, Y, ldy, M, N, K);
        TSG_KERNEL_CHECK("k_tcsc_skinny_direct");
        return TSG_OK;
    }
    float *XS = nullptr;
    WsHold ws(1);
    TSG_TRY(ws.acquire((size_t)groups * K * SK_MT * sizeof(float), reinterpret_cast<void **>(&XS)));
    k_skinny_pack_x<<<dim3((K + 255) / 256, groups), 256, 0, st>>>(X, XS, M, K);
    TSG_KERNEL_CHECK("k_skinny_pack_x");
    dim3 grid(ctas, groups);
    const int mt = (M >= 5) ? 8 : (M >= 3 ? 4 : (M == 2 ? 2 : 1));
    switch (mt) {
        case 1: k_tcsc_skinny<1><<<grid, 256, 0, st>>>(XS, W->csp, W->csn, W->rip, W->rin, B, a, use_prelu, Y, ldy, M, N, K); break;
        case 2: k_tcsc_skinny<2><<<grid, 256, 0, st>>>(XS, W->csp, W->csn, W->rip, W->rin, B, a, use_prelu, Y, ldy, M, N, K); break;
        case 4: k_tcsc_skinny<4><<<grid, 256, 0, st>>>(XS, W->csp, W->csn, W->rip, W->rin, B, a, use_prelu, Y, ldy, M, N, K); break;
        default: k_tcsc_skinny<8><<<grid, 256, 0, st>>>(XS, W->csp, W->csn, W->rip, W->rin, B, a, use_prelu, Y, ldy, M, N, K); break;
    }
    TSG_KERNEL_CHECK("k_tcsc_skinny");
    return ws.release();
}

static thread_local int g_plan_sub_all = 0;
void set_plan_sub_all(int sub) { g_plan_sub_all = sub; }

// ---- host-side planning (pure arithmetic; also exported for the CPU tests: tsg_plan_*) ---------------------------------
struct UnitPlan {
    int mtiles, ntiles, units_full, sub, units_total;
};

// unit decomposition: full 256-column tiles for as many complete rounds of the persistent grid as there are, the
// left-over tiles cut into 2 or 4 narrower units each so that the last round is short (tail balancing)
static UnitPlan plan_units(int M, int N, int sms, int max_sub = 8) {
    UnitPlan u;
    u.mtiles = (M + TM - 1) / TM;
    u.ntiles = (N + CWMAX * NWARP - 1) / (CWMAX * NWARP);
    const int U = u.mtiles * u.ntiles;
    u.units_full = (U / sms) * sms;
    // every tile cut into 2 or 4 units: the multi-GPU path at 8 ranks asks for it (smaller, more frequent tiles keep the fabric
    // busy evenly while the kernel runs, dist.cu mode 5); TSG_FORCE_SUB does the same for experiments
    static const int env_sub = getenv("TSG_FORCE_SUB") ? atoi(getenv("TSG_FORCE_SUB")) : 0;
    const int force_sub = g_plan_sub_all ? g_plan_sub_all : env_sub;
    if ((force_sub == 2 || force_sub == 4) && U >= sms) {
        u.units_full = 0;
        u.sub = force_sub;
        u.units_total = U * force_sub;
        return u;
    }
    if (U < sms) {
        // fewer tiles than SMs (mid-size M, narrow W): cut EVERY tile into `sub` units of 256/sub columns.  A narrower unit
        // still streams the whole X tile, so its time does not fall below the fill-bound floor (about 1/8 of a full unit)
        int best_sub = 1;
        double best = 1e30;
        for (int sub = 1; sub <= max_sub; sub *= 2) {
            const double unit = 1.0 / sub + 0.06;
            const double cost = (double)((U * sub + sms - 1) / sms) * (unit > 0.125 ? unit : 0.125);
            if (cost < best - 1e-9) { best = cost; best_sub = sub; }
        }
        u.units_full = 0;
        u.sub = best_sub;
        u.units_total = U * best_sub;
        return u;
    }
    const int R = U - u.units_full;
    u.sub = 1;
    if (R > 0) {
        // cost of the tail in full-unit times for sub = 1, 2, 4 (narrower units re-stream X, so prefer the smaller sub on ties)
        double best = (double)((R + sms - 1) / sms);
        for (int sub = 2; sub <= 4; sub *= 2) {
            const double cost = (double)((R * sub + sms - 1) / sms) / sub * (1.0 + 0.04 * sub);
            if (cost < best - 1e-9) { best = cost; u.sub = sub; }
        }
    }
    u.units_total = u.units_full + R * u.sub;
    return u;
}

// progress groups (dist.cu mode 2): three quarters of the row tiles in three big groups (they complete round by round of
// the persistent grid anyway), then ever smaller ones so that little is left to push once the kernel retires; target[g] =
// arrivals group g will see = one per compute warp per unit covering its row tiles
static void plan_progress(const UnitPlan &u, Progress *prog) {
    const double frac[8] = {0.25, 0.25, 0.25, 0.125, 0.0625, 0.03125, 0.015625, 1.0};
    int b = 0, g = 0;
    for (int i = 0; i < 9; ++i) prog->gbound[i] = 0;
    while (b < u.mtiles && g < 8) {
        int sz = (g == 7) ? u.mtiles - b : (int)(u.mtiles * frac[g] + 0.5);
        if (sz < 1) sz = 1;
        if (b + sz > u.mtiles) sz = u.mtiles - b;
        b += sz;
        prog->gbound[++g] = b;
    }
    prog->ngroups = g;
    for (int i = 0; i < 8; ++i) prog->target[i] = 0;
    for (int i = 0; i < g; ++i) {
        unsigned int n = 0;
        for (int mt = prog->gbound[i]; mt < prog->gbound[i + 1]; ++mt)
            for (int nt = 0; nt < u.ntiles; ++nt) n += (mt * u.ntiles + nt < u.units_full) ? 1u : (unsigned int)u.sub;
        prog->target[i] = n * NWARP;
    }
}

template <bool FAST, int CL>
static int launch_cluster(const GemmParams &p, size_t smem_bytes, int tiles, cudaStream_t st) {
    // as many clusters as tiles, capped by what the device can co-schedule (a cluster of 4 cannot use every SM: GPC granularity)
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(NTHREADS);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CL;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    cfg.gridDim = dim3(CL * (tiles > 0 ? tiles : 1));
    int max_clusters = 0;
    TSG_CUDA(cudaOccupancyMaxActiveClusters(&max_clusters, k_tcsc_gemm<false, false, FAST, CL>, &cfg));
    if (max_clusters < 1) return set_error(TSG_EUNSUPPORTED, "no cluster of %d CTAs fits this device", CL);
    const int nclusters = tiles < max_clusters ? tiles : max_clusters;
    cfg.gridDim = dim3(CL * nclusters);
    TSG_CUDA(cudaLaunchKernelEx(&cfg, k_tcsc_gemm<false, false, FAST, CL>, p));
    return TSG_OK;
}

// cl: 1 = one CTA per unit; 2 / 4 = thread-block clusters that share a row tile of X (mid-size M)
static int launch_tiled(const GemmParams &p, size_t smem_bytes, bool tile_sep, bool mc = false, int cl = 1) {
    const bool fast = (p.order == TSG_ORDER_FAST);
    if (fast && tile_sep) return set_error(TSG_EUNSUPPORTED, "TSG_ORDER_FAST is not available with the separate-tile epilogue (dist mode 4)");
    static std::atomic<unsigned long long> attr_done{0};
    TSG_TRY(once_per_device(attr_done, [] {
        TSG_CUDA(cudaFuncSetAttribute(k_tcsc_gemm<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
        TSG_CUDA(cudaFuncSetAttribute(k_tcsc_gemm<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
        TSG_CUDA(cudaFuncSetAttribute(k_tcsc_gemm<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
        TSG_CUDA(cudaFuncSetAttribute(k_tcsc_gemm<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
        TSG_CUDA(cudaFuncSetAttribute(k_tcsc_gemm<false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
        TSG_CUDA(cudaFuncSetAttribute(k_tcsc_gemm<false, false, false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
        TSG_CUDA(cudaFuncSetAttribute(k_tcsc_gemm<false, false, false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
        TSG_CUDA(cudaFuncSetAttribute(k_tcsc_gemm<false, false, true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
        TSG_CUDA(cudaFuncSetAttribute(k_tcsc_gemm<false, false, true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
        return (int)TSG_OK;
    }));
    const int grid = p.units_total < num_sms() ? p.units_total : num_sms();
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (g_profile) {  // bench.py: device time of this kernel alone, measured on the launching stream
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        cudaEventRecord(e0, stream());
    }
    if (cl > 1 && !mc && !tile_sep) {
        const int tiles = p.mtiles * p.ntiles;
        int rc;
        if (cl == 2) rc = fast ? launch_cluster<true, 2>(p, smem_bytes, tiles, stream()) : launch_cluster<false, 2>(p, smem_bytes, tiles, stream());
        else rc = fast ? launch_cluster<true, 4>(p, smem_bytes, tiles, stream()) : launch_cluster<false, 4>(p, smem_bytes, tiles, stream());
        if (rc != TSG_OK) return rc;
    } else if (mc && fast) k_tcsc_gemm<false, true, true><<<grid, NTHREADS, smem_bytes, stream()>>>(p);
    else if (mc) k_tcsc_gemm<false, true><<<grid, NTHREADS, smem_bytes, stream()>>>(p);
    else if (tile_sep) k_tcsc_gemm<true><<<grid, NTHREADS, smem_bytes, stream()>>>(p);
    else if (fast) k_tcsc_gemm<false, false, true><<<grid, NTHREADS, smem_bytes, stream()>>>(p);
    else k_tcsc_gemm<false><<<grid, NTHREADS, smem_bytes, stream()>>>(p);
    TSG_KERNEL_CHECK("k_tcsc_gemm");
    if (g_profile) {
        cudaEventRecord(e1, stream());
        g_prof_events.push_back(e0);
        g_prof_events.push_back(e1);
    }
    return TSG_OK;
}

}  // namespace tsg
using namespace tsg;

namespace tsg {
int tcsc_gemm_peers(tsg_tcsc *W, const float *X, const float *B, float a, int use_prelu, int order, float *Y, int M, int N, int K,
                    long long ldy, int npeer, float *const *peerY, unsigned int *done, Progress *prog, int fused_tma);
}

extern "C" {

static thread_local int g_fast_order = -1;  // -1: not decided yet (environment TSG_FAST_ORDER)
int tsg_set_fast_order(int on) {
    g_fast_order = on ? 1 : 0;
    return TSG_OK;
}
int tsg_get_fast_order(void) {
    if (g_fast_order < 0) {
        const char *e = getenv("TSG_FAST_ORDER");
        g_fast_order = (e && atoi(e) != 0) ? 1 : 0;
    }
    return g_fast_order;
}

int tsg_tcsc_set_kernel(int which) {
    g_force_kernel = which;
    return TSG_OK;
}
int tsg_tcsc_get_kernel(void) { return g_force_kernel; }

// planning hooks (pure host arithmetic, usable without a device): the unit decomposition and the progress groups
int tsg_plan_units(int M, int N, int sms, int out5[5]) {
    if (M <= 0 || N <= 0 || sms <= 0) return set_error(TSG_EINVAL, "tsg_plan_units: bad arguments");
    const UnitPlan u = plan_units(M, N, sms);
    out5[0] = u.mtiles; out5[1] = u.ntiles; out5[2] = u.units_full; out5[3] = u.sub; out5[4] = u.units_total;
    return TSG_OK;
}
int tsg_plan_unit_at(int M, int N, int sms, int u, int out3[3]) {
    if (M <= 0 || N <= 0 || sms <= 0) return set_error(TSG_EINVAL, "tsg_plan_unit_at: bad arguments");
    const UnitPlan up = plan_units(M, N, sms);
    if (u < 0 || u >= up.units_total) return set_error(TSG_EINVAL, "tsg_plan_unit_at: unit out of range");
    const Unit un = decode_unit_raw(up.units_full, up.sub, up.ntiles, u);
    out3[0] = un.mt; out3[1] = un.n0; out3[2] = un.cw;
    return TSG_OK;
}
int tsg_plan_progress(int M, int N, int sms, int *ngroups, int gbound9[9], unsigned int target8[8]) {
    if (M <= 0 || N <= 0 || sms <= 0) return set_error(TSG_EINVAL, "tsg_plan_progress: bad arguments");
    Progress pr;
    plan_progress(plan_units(M, N, sms), &pr);
    *ngroups = pr.ngroups;
    for (int i = 0; i < 9; ++i) gbound9[i] = pr.gbound[i];
    for (int i = 0; i < 8; ++i) target8[i] = pr.target[i];
    return TSG_OK;
}

// per-launch device timing of the tiled GEMM kernel (CUDA events on the launching stream)
int tsg_profile_enable(int on) {
    g_profile = on;
    return TSG_OK;
}
// synchronises, sums the elapsed time of every tiled-kernel launch recorded since the last call, clears the list
int tsg_profile_read(double *total_ms, int *launches) {
    double tot = 0.0;
    int n = 0;
    for (size_t i = 0; i + 1 < g_prof_events.size(); i += 2) {
        float ms = 0.f;
        cudaEventSynchronize(g_prof_events[i + 1]);
        if (cudaEventElapsedTime(&ms, g_prof_events[i], g_prof_events[i + 1]) == cudaSuccess) { tot += ms; ++n; }
        cudaEventDestroy(g_prof_events[i]);
        cudaEventDestroy(g_prof_events[i + 1]);
    }
    g_prof_events.clear();
    if (total_ms) *total_ms = tot;
    if (launches) *launches = n;
    return TSG_OK;
}

int tsg_tcsc_gemm(tsg_tcsc *W, const float *X, const float *B, float a, int use_prelu, int order, float *Y, int M, int N, int K,
                  long long ldy) {
    return tcsc_gemm_peers(W, X, B, a, use_prelu, order, Y, M, N, K, ldy, 0, nullptr, nullptr, nullptr, 0);
}

}  // extern "C"

namespace tsg {
int tcsc_gemm_peers(tsg_tcsc *W, const float *X, const float *B, float a, int use_prelu, int order, float *Y, int M, int N, int K,
                    long long ldy, int npeer, float *const *peerY, unsigned int *done, Progress *prog, int fused_tma) {
    TSG_TRY(ensure_device());
    if (npeer < 0 || npeer > TSG_MAX_PEERS) return set_error(TSG_EINVAL, "too many peers");
    if (!W || !X || !B || !Y) return set_error(TSG_EINVAL, "tsg_tcsc_gemm: null argument");
    if (N != W->cols || K != W->rows) return set_error(TSG_EINVAL, "tsg_tcsc_gemm: W is %d x %d but K=%d, N=%d", W->rows, W->cols, K, N);
    if (order < 0 || order > 3) return set_error(TSG_EINVAL, "tsg_tcsc_gemm: bad order %d", order);
    if (ldy < N) return set_error(TSG_EINVAL, "tsg_tcsc_gemm: ldy < N");
    if (!fused_tma && npeer == 0 && !done) {  // TSG_FUSED_EPILOGUE=1|2: route the single-GPU store through the TMA epilogue too (2 = separate tile)
        static const int env_fused = getenv("TSG_FUSED_EPILOGUE") ? atoi(getenv("TSG_FUSED_EPILOGUE")) : 0;
        if (env_fused && M >= TSG_SKINNY_M && !(N & 3) && !(ldy & 3) && !(reinterpret_cast<uintptr_t>(Y) & 15)) fused_tma = (env_fused == 2) ? 2 : 1;
    }
    if (fused_tma && ((N & 3) || (ldy & 3) || (reinterpret_cast<uintptr_t>(Y) & 15)))
        return set_error(TSG_EUNSUPPORTED, "fused TMA epilogue needs N, ldy multiples of 4 and a 16-byte aligned Y");
    if (M <= 0 || N <= 0) return TSG_OK;
    const bool skinny = npeer == 0 && !done && !fused_tma && ((g_force_kernel == 2) || (g_force_kernel == 0 && M < TSG_SKINNY_M));
    if (skinny) return launch_skinny(W, X, B, a, use_prelu, Y, M, N, K, ldy);

    if (order == TSG_ORDER_FAST && npeer == 0 && !done && !fused_tma && M >= TSG_SKINNY_M) {  // dense regime of the opt-in order: multiplier form, every k
        int handled = 0;
        TSG_TRY(tcsc_gemm_dense_fast(W, X, B, a, use_prelu, Y, M, N, K, ldy, &handled));
        if (handled) return TSG_OK;
    }
    const bool tile_sep = (fused_tma == 2), mc = (fused_tma == 3);
    constexpr int kTileBytes = TM * TILE_PITCH * 4;
    const int sep_bytes = tile_sep ? kTileBytes : 0;  // shared memory behind the ring that is not part of it
    const bool fast = (order == TSG_ORDER_FAST);
    // ring depth: 3 shorter stages let a warp run up to two chunks ahead of the slowest one (TSG_STAGES=2|3; the overlay epilogues keep 2)
    static const int env_stages = getenv("TSG_STAGES") ? atoi(getenv("TSG_STAGES")) : 2;
    const int nstage = (env_stages == 3 && fused_tma != 1 && fused_tma != 2) ? 3 : 2;
    TSG_TRY(build_kstream(W, sep_bytes, fast ? 2 : 1, nstage));
    const KStream &ks = fast ? W->ks_fast : W->ks;
    GemmParams p;
    p.mtiles = (M + TM - 1) / TM;
    float *XT = nullptr;
    WsHold ws(0);
    TSG_TRY(ws.acquire((size_t)p.mtiles * (K > 0 ? K : 1) * TM * sizeof(float), reinterpret_cast<void **>(&XT)));
    if (K > 0) TSG_TRY(transpose_x_tiles(X, XT, M, K));
    p.XT = XT; p.cnt = ks.cnt; p.woff = ks.woff; p.body = ks.body; p.B = B; p.Y = Y; p.ldy = ldy;
    p.M = M; p.N = N; p.K = K; p.kc = ks.kc; p.nchunk = (K > 0) ? ks.nchunk : 0; p.ncols_pad = ks.ncols_pad; p.ngroup = ks.ngroup;
    p.a = a; p.use_prelu = use_prelu; p.order = order;
    p.npeer = (fused_tma == 3) ? 0 : npeer;  // multicast: peerY[0] is the multicast mapping, used by the drain warps only
    for (int q = 0; q < TSG_MAX_PEERS; ++q) p.peerY[q] = (q < npeer) ? peerY[q] : nullptr;
    p.xstage_bytes = (uint32_t)ks.kc * TM * 4;
    p.body_stage_bytes = ((uint32_t)ks.max_tile_words * 4 + 15) & ~15u;
    p.nstage = nstage;
    size_t ring_bytes = (size_t)nstage * ((size_t)p.xstage_bytes + (fast ? 2 : 1) * (size_t)(p.body_stage_bytes + CNT_BYTES + WOFF_BYTES));
    if (fused_tma == 1 && ring_bytes < (size_t)kTileBytes) ring_bytes = (size_t)kTileBytes;  // tiny K: the tile is the larger one
    p.tile_off = sep_bytes ? (uint32_t)ring_bytes : 0u;
    p.bar_off = (uint32_t)(ring_bytes + sep_bytes);
    const size_t smem_bytes = (size_t)p.bar_off + 128;
    if (smem_bytes > 232448) return set_error(TSG_EUNSUPPORTED, "tsg_tcsc_gemm: the gather stream of this matrix leaves no room for a separate output tile");
    // the staged (TMA) epilogues write float4 pieces per warp: at least 4 columns per warp there
    const UnitPlan up = plan_units(M, N, num_sms(), (fused_tma == 1 || fused_tma == 2) ? 4 : 8);
    p.ntiles = up.ntiles;
    p.units_full = up.units_full;
    p.sub = up.sub;
    p.units_total = up.units_total;
    p.fused_tma = mc ? 0 : fused_tma;  // the multicast variant stores its tile like the single-GPU kernel; its drain warps do the rest
    p.done = done;
    p.ngroups = 0;
    for (int g = 0; g < 9; ++g) p.gbound[g] = 0;
    if (done && prog) {
        plan_progress(up, prog);
        p.ngroups = prog->ngroups;
        for (int i = 0; i <= prog->ngroups; ++i) p.gbound[i] = prog->gbound[i];
    }
    // opt-in (TSG_CLUSTER=1): with fewer tiles than SMs the parts of a tile run as one thread-block cluster that receives X by multicast
    static const int env_cluster = getenv("TSG_CLUSTER") ? atoi(getenv("TSG_CLUSTER")) : 0;  // measured: no faster (profiles/), L2 traffic / CL
    int cl = 1;
    if (env_cluster && !fused_tma && npeer == 0 && !done && up.units_full == 0 && up.sub >= 2 && K > 0) cl = up.sub >= 4 ? 4 : 2;
    int rc = launch_tiled(p, smem_bytes, tile_sep, mc, cl);
    int rc2 = ws.release();
    return rc ? rc : rc2;
}
}  // namespace tsg
