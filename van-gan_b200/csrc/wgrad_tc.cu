// Conv3D weight gradient (stride 1) on the 5th-generation tensor cores.
//
// Replaces cuDNN Conv3DBackpropFilterV2 behind every stride-1 Keras Conv3D on the hot path
// (resunet_model.py:64-65,89-90,96,133-134; discriminator.py:91-103), i.e. the `minimize` calls at
// vangan.py:426-438.
//
//     dW[td,th,tw][ci][co] = sum over output voxels v of  X[v + t][ci] * dY[v][co]
//
// GEMM view: the reduction (K) dimension is the voxel index, which in NDHWC storage is the SLOW index
// of both operands -> both operands are "MN-major" for tcgen05.mma.  One MMA consumes 16 consecutive
// w-voxels (K = 16).  The two operands are staged in shared memory as planes of 16-byte cells (8 channels
// of one voxel), w-contiguous, so 8 consecutive voxels x 8 channels form one canonical 128-byte
// no-swizzle core matrix, and a filter tap is nothing but a different descriptor start address.
//
// Channel counts on this path are small (16..64 at the expensive resolutions), far below the MMA's
// M = 64/128 and a poor match for N, so two folds turn unused MMA area into useful taps:
//   * th-fold (M side): the X planes are laid out [d][chunk][h][plane][w]; M-groups then run over
//     (row offset j, plane) at one constant stride, so an M = 128 (or 64) operand holds R = M/CC
//     consecutive h-rows of a CC-channel chunk: D rows (j, ci) accumulate the taps th = th_base + j.
//   * td-fold (N side): the dY planes are laid out [h][d][plane][w] with K-1 zero slices on either side,
//     so an N = K*CO operand holds K consecutive d-slices: D columns (r, co) accumulate td = K-1-r.
//   * tw-fold (N side, 16-wide co blocks: K*K*CO <= 256): every dY slice is staged K times, copy c shifted by c voxels along w
//     ([h][d][copy][plane][w]), so D columns (r, c, co) accumulate (td, tw) = (K-1-r, c) and ONE MMA of N = K*K*CO covers what
//     took K accumulators before: an MMA costs (operand fetch) ~ M/4 + N/4 cycles whatever part of it is useful.
//     The X brick then needs no w halo; bricks tile the X w-axis (OW + K - 1 voxels).
// Folded slices that fall outside the brick are not staged as zeros: the edge X slices issue a narrower MMA (fewer r blocks).
// The remaining tap index (tw, and th_base / td when not folded) selects the TMEM accumulator.
// A CTA owns a set of accumulators (<= 512 TMEM columns) for one (ci-block, co-block, tap-set) tile and a
// split-K share of the voxel bricks; it finishes with fp32 red.global.add into dW (Keras layout).
//
// Roles (416 threads): warps 0-11 gather the X halo brick and the dY brick (16-byte cp.async, zero fill
// outside the tensors), warp 12 issues tcgen05.mma (one elected lane) and owns TMEM; when the bricks are
// done warps 0-3 drain the accumulators.  Stages form a full/empty mbarrier ring released by tcgen05.commit.
// ncu (profiles/r02_ncu_full_wg_16-16_*_call8.txt) showed the kernel bound by the PRODUCERS' instruction streams (8 warps x ~1 200
// dependent instructions per brick, tensor pipe 24-35 % busy), hence: 12 producer warps (the drain warps used to spin on the final
// barrier for the whole kernel), an incremental row decode and a bounds-free inner loop for rows that lie inside the tensor.
#include <stdio.h>
#include <stdlib.h>

#include "tc_ptx.cuh"

namespace {

using namespace tcp;

constexpr int WG_NPROD = 384;                 // producer threads: warps 0..11 (warps 0..3 also drain TMEM at the end)
constexpr int WG_MMA_WARP = WG_NPROD / 32;    // warp 12 issues the MMAs
constexpr int WG_THREADS = WG_NPROD + 32;
constexpr int WG_MAXACC = 16;
constexpr int WG_KW = 16;   // voxels along w per MMA (K of the bf16 MMA)
// Plane pitches (in 16-byte cells) are ODD: for a fixed k the MMA reads one 16-byte row from every M/N-group, i.e. addresses
// at a stride of one plane pitch; an odd pitch spreads 8 consecutive groups over all 8 bank groups (no conflicts), whereas a
// pitch that is a multiple of 8 cells serialises them (measured: 16-29 B/cycle of operand fetch).
constexpr int WG_XW = 19;   // X plane pitch: >= 16 + K - 1 for K <= 4
constexpr int WG_YP = 17;   // dY plane pitch

struct WgParams {
    const bf16* x;
    const bf16* dy;
    float* dw;
    int Nb, XD, XH, XW, Cx, OD, OH, OW, Cy, K, st;   // st: convolution stride (1 or 2)
    int M, CC, R, PLC, NCH, CB;     // MMA M; channels per chunk; rows folded; planes per chunk; chunks per CTA; CB = NCH*CC
    int CO, NF, NW, KW, Nmma, NPLy; // co block; slices folded (1 or K); w-shifted dY copies (1 or K); KW = taps along w that select an
                                    // accumulator (K, or 1 with the tw-fold); MMA N = NF*NW*CO; dY planes = CO/8
    int nth, NTA, TPC, nsets;       // th bases; tap-accumulators in total / per CTA; tap sets
    int n_co_blocks, tiles, ksplit;
    int BDo, BHo, bd_tiles, bh_tiles, bw_tiles, nbricks;
    int XDb, XHb, XHu, XWb, YDb, nB;         // XHu: rows that carry useful taps (<= XHb)
    int x_sd, x_sc, x_spar, x_sh, x_nw;      // X stage strides (cells): slice, chunk, w-parity plane set (stride 2), row; voxels per row
    uint32_t x_bytes, stage_bytes, tmem_cols;
    int stages, ca;
};

// One gathered region: cells of 16 bytes (8 channels of one voxel).  A region row is (slice d, channel chunk c, row h);
// inside a row the cells are (voxel w, plane pl) with P = 2^lgp planes per chunk.
struct Region {
    const bf16* g;        // sample base + channel offset of the CTA's channel block
    int gd0, gh0, gw0;    // global origin of the region
    int GD, GH, GW, C;    // tensor bounds and channels per voxel
    int nd, nc, nh, nw;   // extents: slices, chunks, rows, voxels per row
    int lgp;              // log2(planes per chunk)
    int s2, spar;         // stride-2 source: voxel w goes to parity plane set (w & 1) at index w >> 1
    int ca;               // cp.async.ca (through L1: the second 16-byte half of a 32-byte sector hits) instead of .cg
    int sd, sc, sh, sp;   // shared-memory strides in cells: slice, chunk, row, plane (voxel stride = 1)
    FastDiv by_cnh, by_nh;
};

// 16-byte cp.async per cell (zero fill outside the tensor): no register staging, so a whole stage is in flight per SM.
// A warp owns whole region rows (row = warp, warp + 12, ...): (d, c, h) are carried incrementally, every lane keeps a fixed
// (plane, voxel phase), and a row that lies inside the tensor runs a loop of copy + two adds per cell.
__device__ __forceinline__ void gather_region(const Region& r, uint32_t dst_base, int warp, int lane) {
    constexpr int NWARP = WG_NPROD / 32;
    const int P = 1 << r.lgp;
    const int pl = lane & (P - 1), wl = lane >> r.lgp, wps = 32 >> r.lgp;
    const int rows = r.nd * r.nc * r.nh;
    const uint32_t lane_dst = (uint32_t)(pl * r.sp + (r.s2 ? (wl & 1) * r.spar + (wl >> 1) : wl)) * 16u;
    const uint32_t dst_step = (uint32_t)(r.s2 ? wps >> 1 : wps) * 16u;
    const int lane_src = wl * r.C + pl * 8;
    const int wstep_src = wps * r.C;
    const bool w_inside = r.gw0 >= 0 && r.gw0 + r.nw <= r.GW;
    int d = (int)r.by_cnh.div(warp), rem = warp - d * (r.nc * r.nh);
    int c = (int)r.by_nh.div(rem), h = rem - c * r.nh;
    for (int row = warp; row < rows; row += NWARP) {
        const int gd = r.gd0 + d, gh = r.gh0 + h;
        const bool rowok = (unsigned)gd < (unsigned)r.GD && (unsigned)gh < (unsigned)r.GH;
        const bf16* src = r.g + (ptrdiff_t)(((rowok ? gd : 0) * r.GH + (rowok ? gh : 0)) * r.GW + r.gw0) * r.C + ((c << (r.lgp + 3)) + lane_src);
        uint32_t dst = dst_base + (uint32_t)(d * r.sd + c * r.sc + h * r.sh) * 16u + lane_dst;
        if (rowok && w_inside) {
#pragma unroll 2
            for (int w = wl; w < r.nw; w += wps) {
                if (r.ca) asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(src) : "memory");
                else asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(src) : "memory");
                src += wstep_src;
                dst += dst_step;
            }
        } else {
            for (int w = wl; w < r.nw; w += wps) {
                const bool ok = rowok && (unsigned)(r.gw0 + w) < (unsigned)r.GW;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(ok ? src : r.g), "r"(ok ? 16 : 0) : "memory");
                src += wstep_src;
                dst += dst_step;
            }
        }
        h += NWARP;
        while (h >= r.nh) {
            h -= r.nh;
            if (++c == r.nc) { c = 0; d++; }
        }
    }
}

struct IssueCtx {
    uint64_t a_desc0, b_desc0;
    uint32_t idesc[5], leader, tmem_base, sbase16, stage16, x16, full0, empty0, done_bar;   // idesc[n]: n folded slices in N
    int a_sd, a_sh, b_sd, b_sh, nB, BHo, BDo, NF, colr, Nmma, stages, first, nbricks, step;   // colr: D columns per folded slice
};

// The MMA warp's whole life: for every brick of this CTA wait for the stage, issue nB x BHo x NACC MMAs, release the stage.
// Descriptors advance by plain 64-bit adds on the 14-bit start-address field (shared memory is < 256 KB, so no carry).
template <int NACC>
__device__ __forceinline__ void issue_bricks(const IssueCtx& c, const int* s_aoff) {
    uint64_t adesc[NACC];
    uint32_t tm[NACC];
#pragma unroll
    for (int a = 0; a < NACC; a++) {
        adesc[a] = c.a_desc0 + (uint32_t)s_aoff[a];
        tm[a] = c.tmem_base + (uint32_t)(a * c.Nmma);
    }
    int stage = 0;
    uint32_t phase = 0;
    for (int brick = c.first; brick < c.nbricks; brick += c.step) {
        mbar_wait(c.full0 + 8 * stage, phase);
        tc_fence_after();
        const uint32_t xs16 = c.sbase16 + stage * c.stage16;
        uint32_t a_row = xs16;
        for (int bs = 0; bs < c.nB; bs++) {
            // td-fold: X slice bs meets the dY slices bs-(NF-1)+r, r = 0..NF-1, that exist in the brick -> r_lo..r_hi
            const int r_lo = max(0, c.NF - 1 - bs), r_hi = min(c.NF - 1, c.BDo + c.NF - 2 - bs), cnt = r_hi - r_lo + 1;
            const uint32_t idesc = cnt == 1 ? c.idesc[1] : (cnt == 2 ? c.idesc[2] : (cnt == 3 ? c.idesc[3] : c.idesc[4]));
            const uint32_t dcol = (uint32_t)(r_lo * c.colr);
            uint32_t a_pos = a_row, b_pos = xs16 + c.x16 + (uint32_t)((bs - (c.NF - 1) + r_lo) * c.b_sd);
            for (int h = 0; h < c.BHo; h++) {
                if (c.leader) {
                    const uint64_t bdesc = c.b_desc0 + b_pos;
#pragma unroll
                    for (int a = 0; a < NACC; a++) tc_mma(tm[a] + dcol, adesc[a] + a_pos, bdesc, idesc, 1u);
                }
                a_pos += c.a_sh; b_pos += c.b_sh;
            }
            a_row += c.a_sd;
        }
        __syncwarp();
        if (c.leader) tc_commit(c.empty0 + 8 * stage);
        if (++stage == c.stages) { stage = 0; phase ^= 1; }
    }
    if (c.leader) tc_commit(c.done_bar);
}

__global__ void __launch_bounds__(WG_THREADS, 1) wgrad_tc_kernel(const WgParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* bar_base = smem + (size_t)p.stages * p.stage_bytes;
    const uint32_t full0 = s_addr(bar_base), empty0 = full0 + 8 * p.stages, done_bar = empty0 + 8 * p.stages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_base + 16 * p.stages + 16);
    int* s_aoff = reinterpret_cast<int*>(bar_base + 16 * p.stages + 32);   // WG_MAXACC ints
    const uint32_t sbase = s_addr(smem);

    // ---- which tile / split is this CTA
    const int tile = blockIdx.x % p.tiles, split = blockIdx.x / p.tiles;
    const int set = tile % p.nsets;
    const int cob = (tile / p.nsets) % p.n_co_blocks;
    const int cib = tile / (p.nsets * p.n_co_blocks);
    const int q0 = set * p.TPC;
    const int nq = min(p.TPC, p.NTA - q0);
    const int nacc = nq * p.NCH;
    const bool fold = p.NF > 1;
    // tap-accumulator q -> (tw fastest, then th base, then td)
    int td_min = 1 << 30, thb_min = 1 << 30, tw_min = 1 << 30;
    for (int i = 0; i < nq; i++) {
        const int q = q0 + i;
        const int tw = q % p.KW, thb = (q / p.KW) % p.nth, td = q / (p.KW * p.nth);
        td_min = min(td_min, td); thb_min = min(thb_min, thb); tw_min = min(tw_min, tw);
    }
    if (fold) td_min = 0;
    const int tw_org = p.st == 2 ? (tw_min & ~1) : tw_min;   // stride 2: keep the parity of the taps

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; s++) {
            mbar_init(full0 + 8 * s, WG_NPROD);
            mbar_init(empty0 + 8 * s, 1);
        }
        mbar_init(done_bar, 1);
        mbar_init_fence();
        for (int i = 0; i < nq; i++) {
            const int q = q0 + i;
            const int tw = q % p.KW, thb = (q / p.KW) % p.nth, td = fold ? 0 : q / (p.KW * p.nth);
            for (int c = 0; c < p.NCH; c++)
                s_aoff[i * p.NCH + c] = (td - td_min) * p.x_sd + c * p.x_sc + (thb - thb_min) * p.R * p.x_sh +
                                        (p.st == 2 ? ((tw - tw_org) & 1) * p.x_spar + ((tw - tw_org) >> 1) : tw - tw_org);
        }
    }
    // zero the stages once (pad cells of the odd plane pitches are never written)
    for (uint32_t o = threadIdx.x * 16u; o < (uint32_t)p.stages * p.stage_bytes; o += WG_THREADS * 16u)
        asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};\n" ::"r"(sbase + o), "r"(0) : "memory");
    fence_async_smem();
    if (warp == WG_MMA_WARP) tmem_alloc(s_addr(tmem_slot), p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // every MMA accumulates (the edge slices of a td-folded brick touch only part of the columns, so "first MMA overwrites" does
    // not cover them): the drain warps hand the accumulators over zeroed
    if (warp < 4) {
        for (uint32_t col = 0; col < p.tmem_cols; col += 16) tc_st16_zero(tmem_base + ((uint32_t)(warp * 32) << 16) + col);
        tc_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    if (warp < WG_MMA_WARP) {
        // ------------------------------------------------------------------ gather producers
        Region rx, ry;
        int lgx = 0, lgy = 0;
        while ((1 << lgx) < p.PLC) lgx++;
        while ((1 << lgy) < p.NPLy) lgy++;
        rx.ca = ry.ca = p.ca;
        rx.GD = p.XD; rx.GH = p.XH; rx.GW = p.XW; rx.C = p.Cx;
        rx.nd = p.XDb; rx.nc = p.NCH; rx.nh = p.XHu; rx.nw = p.x_nw; rx.lgp = lgx;
        rx.sd = p.x_sd; rx.sc = p.x_sc; rx.sh = p.x_sh; rx.sp = p.XWb; rx.s2 = p.st == 2; rx.spar = p.x_spar;
        rx.by_cnh = FastDiv(rx.nc * rx.nh); rx.by_nh = FastDiv(rx.nh);
        ry.GD = p.OD; ry.GH = p.OH; ry.GW = p.OW; ry.C = p.Cy;
        ry.nd = p.BDo; ry.nc = 1; ry.nh = p.BHo; ry.nw = WG_KW; ry.lgp = lgy;
        ry.s2 = 0; ry.spar = 0;
        ry.sd = p.NW * p.NPLy * WG_YP; ry.sc = 0; ry.sh = p.YDb * ry.sd; ry.sp = WG_YP;
        ry.by_cnh = FastDiv(ry.nh); ry.by_nh = FastDiv(ry.nh);
        int stage = 0, prev_stage = -1;
        uint32_t phase = 0;
        for (int brick = split; brick < p.nbricks; brick += p.ksplit) {
            int b = brick;
            const int bw = b % p.bw_tiles; b /= p.bw_tiles;
            const int bh = b % p.bh_tiles; b /= p.bh_tiles;
            const int bd = b % p.bd_tiles;
            const int n = b / p.bd_tiles;
            rx.g = p.x + (size_t)n * p.XD * p.XH * p.XW * p.Cx + (size_t)cib * p.CB;
            rx.gd0 = bd * p.BDo * p.st + td_min; rx.gh0 = bh * p.BHo * p.st + thb_min * p.R; rx.gw0 = bw * WG_KW * p.st + tw_org;
            ry.g = p.dy + (size_t)n * p.OD * p.OH * p.OW * p.Cy + (size_t)cob * p.CO;
            ry.gd0 = bd * p.BDo; ry.gh0 = bh * p.BHo;
            mbar_wait(empty0 + 8 * stage, phase ^ 1);
            const uint32_t xs = sbase + stage * p.stage_bytes;
            gather_region(rx, xs, warp, lane);
            for (int c = 0; c < p.NW; c++) {   // tw-fold: copy c holds dY shifted by c voxels (zero fill outside the tensor)
                ry.gw0 = bw * WG_KW - c;
                gather_region(ry, xs + p.x_bytes + (uint32_t)(c * p.NPLy * WG_YP) * 16u, warp, lane);
            }
            asm volatile("cp.async.commit_group;\n" ::: "memory");
            if (p.stages >= 3) {
                // lagged publish: hand over the previous brick while this one is in flight
                if (prev_stage >= 0) {
                    asm volatile("cp.async.wait_group 1;\n" ::: "memory");
                    fence_async_smem();   // cp.async wrote through the generic proxy; the MMA reads through the async proxy
                    mbar_arrive(full0 + 8 * prev_stage);
                }
                prev_stage = stage;
            } else {
                // two stages: a lagged publish would wait for the MMA warp to drain the other stage first -> publish at once
                asm volatile("cp.async.wait_group 0;\n" ::: "memory");
                fence_async_smem();
                mbar_arrive(full0 + 8 * stage);
            }
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        if (prev_stage >= 0) {
            asm volatile("cp.async.wait_group 0;\n" ::: "memory");
            fence_async_smem();
            mbar_arrive(full0 + 8 * prev_stage);
        }
    }
    if (warp == WG_MMA_WARP) {
        // ------------------------------------------------------------------ MMA issuer
        // One elected lane issues; the loops are warp-uniform so descriptors stay in uniform registers.  The issue loop is
        // the critical resource of this kernel (one MMA is only 24-128 tensor-pipe cycles), hence the compile-time NACC.
        IssueCtx c;
        for (int n = 1; n <= 4; n++) c.idesc[n] = make_idesc_bf16(p.M, (p.NF > 1 ? n : 1) * p.NW * p.CO, 1, 1);
        c.leader = elect_one();
        c.a_desc0 = ((uint64_t)((uint32_t)p.XWb | (1u << 14)) << 32) | (8u << 16);   // SBO = X plane pitch | version ; LBO = 128 B
        c.b_desc0 = ((uint64_t)((uint32_t)WG_YP | (1u << 14)) << 32) | (8u << 16);   // SBO = dY plane pitch
        c.a_sd = p.st * p.x_sd; c.a_sh = p.st * p.x_sh;
        c.b_sd = p.NW * p.NPLy * WG_YP; c.b_sh = p.YDb * c.b_sd;
        c.nB = p.nB; c.BHo = p.BHo; c.BDo = p.BDo; c.NF = p.NF; c.colr = p.NW * p.CO; c.Nmma = p.Nmma; c.tmem_base = tmem_base;
        c.sbase16 = sbase >> 4; c.stage16 = p.stage_bytes >> 4; c.x16 = p.x_bytes >> 4; c.stages = p.stages;
        c.full0 = full0; c.empty0 = empty0; c.done_bar = done_bar;
        c.first = split; c.nbricks = p.nbricks; c.step = p.ksplit;
        switch (nacc) {
            case 1: issue_bricks<1>(c, s_aoff); break;
            case 2: issue_bricks<2>(c, s_aoff); break;
            case 3: issue_bricks<3>(c, s_aoff); break;
            case 4: issue_bricks<4>(c, s_aoff); break;
            case 5: issue_bricks<5>(c, s_aoff); break;
            case 6: issue_bricks<6>(c, s_aoff); break;
            case 7: issue_bricks<7>(c, s_aoff); break;
            case 8: issue_bricks<8>(c, s_aoff); break;
            case 9: issue_bricks<9>(c, s_aoff); break;
            case 10: issue_bricks<10>(c, s_aoff); break;
            case 11: issue_bricks<11>(c, s_aoff); break;
            case 12: issue_bricks<12>(c, s_aoff); break;
            case 13: issue_bricks<13>(c, s_aoff); break;
            case 14: issue_bricks<14>(c, s_aoff); break;
            case 15: issue_bricks<15>(c, s_aoff); break;
            default: issue_bricks<16>(c, s_aoff); break;
        }
    } else if (warp < 4) {
        // ------------------------------------------------------------------ epilogue: TMEM -> red.global.add
        const int qd = warp;                          // TMEM lane quarter this warp may access
        mbar_wait(done_bar, 0);
        tc_fence_after();
        // lane -> MMA row m.  M = 128: m = 32*qd + lane.  M = 64: rows live in lanes 0-15 of each quarter, m = 16*qd + lane.
        const int m = p.M == 128 ? qd * 32 + lane : qd * 16 + lane;
        const bool lane_ok = p.M == 128 || lane < 16;
        const int j = m / p.CC, cil = m - j * p.CC;
        for (int a = 0; a < nacc; a++) {
            const int i = a / p.NCH, c = a - i * p.NCH;
            const int q = q0 + i;
            const int twq = q % p.KW, thb = (q / p.KW) % p.nth, tdq = q / (p.KW * p.nth);
            const int th = thb * p.R + j;
            const bool row_ok = lane_ok && th < p.K;
            const int ci = cib * p.CB + c * p.CC + cil;
            for (int n0 = 0; n0 < p.Nmma; n0 += 16) {
                uint32_t v[16];
                tc_ld16(tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(a * p.Nmma + n0), v);
                if (row_ok) {
                    const int blk = n0 / p.CO, col = n0 - blk * p.CO;
                    const int r = blk / p.NW, tw = p.NW > 1 ? blk - r * p.NW : twq;
                    const int td = fold ? p.K - 1 - r : tdq;
                    float* o = p.dw + ((size_t)((td * p.K + th) * p.K + tw) * p.Cx + ci) * p.Cy + cob * p.CO + col;
                    // 16-byte vector reductions (sm_90+): a quarter of the atomic requests -- the epilogue's ksplit x |dW| atomics are
                    // a FIXED cost per call (the deep layers at b = 1 spent most of their ~75 us there)
#pragma unroll
                    for (int e = 0; e < 16; e += 4)
                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"l"(o + e), "f"(__uint_as_float(v[e])),
                                     "f"(__uint_as_float(v[e + 1])), "f"(__uint_as_float(v[e + 2])), "f"(__uint_as_float(v[e + 3]))
                                     : "memory");
                }
            }
        }
        tc_fence_before();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == WG_MMA_WARP) {
        tc_fence_after();
        tmem_dealloc(tmem_base, p.tmem_cols);
    }
}

inline int largest_div(int v, const int* cands, int n) {
    for (int i = 0; i < n; i++)
        if (v % cands[i] == 0) return cands[i];
    return 0;
}

}  // namespace

unsigned long long g_vg_wg_tc_launches = 0;

// dw[K,K,K,Cx,Cy] += X^T * dY on tcgen05 (stride 1).  Returns VG_ERR_UNSUPPORTED when the shape does not fit.
int vg_wg_tc_launch(const bf16* x, const bf16* dy, float* dw, int Nb, int XD, int XH, int XW, int Cx, int OD, int OH, int OW, int Cy,
                    int K, int stride, cudaStream_t stream) {
    if (Cx % 16 || Cy % 16 || K < 1 || K > 4 || stride < 1 || stride > 2) return VG_ERR_UNSUPPORTED;
    if (stride == 2 && K < 2) return VG_ERR_UNSUPPORTED;   // k1 s2 would stage 8x the voxels it uses
    WgParams p{};
    p.x = x; p.dy = dy; p.dw = dw;
    p.Nb = Nb; p.XD = XD; p.XH = XH; p.XW = XW; p.Cx = Cx; p.OD = OD; p.OH = OH; p.OW = OW; p.Cy = Cy; p.K = K; p.st = stride;
    const int ccs[4] = {128, 64, 32, 16};
    p.CC = largest_div(Cx, ccs, 4);
    p.M = p.CC == 16 ? 64 : 128;
    p.R = p.M / p.CC;
    p.PLC = p.CC / 8;
    // co block: 128 at most.  A 256-wide block is never td-folded (K*CO > 256), so it only halves the taps a CTA can keep in TMEM
    // (512 columns) and every CTA then streams all of X and dY for two taps: L2 traffic per tile ~ (CB + CO), i.e. 1.5x more at
    // CO = 256 than at CO = 128, where the MMA already runs at the N >= 128 issue floor (VG_WG_CO256=1 restores the old choice).
    static int co256 = -1;
    if (co256 < 0) {
        const char* e = getenv("VG_WG_CO256");
        co256 = (e && e[0] == '1') ? 1 : 0;
    }
    const int cos[5] = {256, 128, 64, 32, 16};
    p.CO = (co256 || stride == 2) ? largest_div(Cy, cos, 5) : largest_div(Cy, cos + 1, 4);   // measured: 256->512 k4 s1 0.96 -> 0.69 ms, 128->256 k4 s2 0.51 -> 0.55 ms (kept at 256)
    p.NPLy = p.CO / 8;
    p.NF = (stride == 1 && K > 1 && K * p.CO <= 256) ? K : 1;
    // tw-fold: opt-in (VG_WG_NW=1).  Measured on B200 (profiles/r02_wgrad_twfold_call9.txt): 3x fewer MMAs but 16->16 0.77 -> 1.14 ms,
    // because the kernel is bound by the 16-byte cp.async gather (~1.5 cycles per cell per SM, one L2 sector request per cell) and
    // the K shifted dY copies add 44 % more cells per brick.
    static int nwfold = -1;
    if (nwfold < 0) {
        const char* e = getenv("VG_WG_NW");
        nwfold = (e && e[0] == '1') ? 1 : 0;
    }
    p.NW = (nwfold && p.NF > 1 && K * K * p.CO <= 256) ? K : 1;
    p.KW = p.NW > 1 ? 1 : K;
    p.Nmma = p.NF * p.NW * p.CO;
    if (p.M == 128 && p.Nmma % 16) return VG_ERR_UNSUPPORTED;
    p.nth = (K + p.R - 1) / p.R;
    p.NTA = (p.NF > 1 ? 1 : K) * p.nth * p.KW;
    int maxacc = 512 / p.Nmma;
    if (maxacc > WG_MAXACC) maxacc = WG_MAXACC;
    const int nch_tot = Cx / p.CC;
    const size_t smem_cap = 220 * 1024;
    const int cand[7][2] = {{8, 8}, {8, 4}, {4, 4}, {4, 2}, {2, 2}, {2, 1}, {1, 1}};
    bool found = false;
    for (int attempt = 0; attempt < 2 && !found; attempt++) {
        // attempt 0: all channel chunks of X in one CTA (dY staged once); attempt 1: one chunk per CTA
        if (attempt == 0) {
            if (p.NTA * nch_tot > maxacc) continue;
            p.NCH = nch_tot; p.TPC = p.NTA;
        } else {
            p.NCH = 1; p.TPC = p.NTA < maxacc ? p.NTA : maxacc;
        }
        p.CB = p.NCH * p.CC;
        p.nsets = (p.NTA + p.TPC - 1) / p.TPC;
        p.n_co_blocks = Cy / p.CO;
        p.tiles = (Cx / p.CB) * p.n_co_blocks * p.nsets;
        // spans of the tap sets
        int tdspan = 0, thspan = 0;
        for (int s = 0; s < p.nsets; s++) {
            int tdl = 1 << 30, tdh = -1, thl = 1 << 30, thh = -1;
            for (int q = s * p.TPC; q < p.NTA && q < (s + 1) * p.TPC; q++) {
                int thb = (q / p.KW) % p.nth, td = q / (p.KW * p.nth);
                tdl = td < tdl ? td : tdl; tdh = td > tdh ? td : tdh;
                thl = thb < thl ? thb : thl; thh = thb > thh ? thb : thh;
            }
            if (tdh - tdl > tdspan) tdspan = tdh - tdl;
            if (thh - thl > thspan) thspan = thh - thl;
        }
        p.XWb = p.NW > 1 ? WG_YP : WG_XW;   // tw-fold: no w halo on the X side
        const int ow_ext = p.NW > 1 ? OW + K - 1 : OW;   // w extent the bricks tile
        static int fb_d = -1, fb_h = -1;                 // VG_WG_BRICK=d,h forces the brick (tuning)
        if (fb_d < 0) {
            fb_d = fb_h = 0;
            const char* e = getenv("VG_WG_BRICK");
            if (e) sscanf(e, "%d,%d", &fb_d, &fb_h);
        }
        const int want = 148 / p.tiles > 0 ? 148 / p.tiles : 1;
        int best = -1;
        for (int ci = 0; ci < 7; ci++) {
            const int bdo = cand[ci][0], bho = cand[ci][1];
            if (fb_d > 0 && (bdo != fb_d || bho != fb_h)) continue;
            if (bdo > 1 && bdo >= 2 * OD) continue;
            if (bho > 1 && bho >= 2 * OH) continue;
            const int xdb = p.NF > 1 ? bdo + K - 1 : (bdo - 1) * stride + tdspan + 1;
            const int xhb = (bho - 1) * stride + thspan * p.R + p.R;
            const int ydb = bdo;
            const size_t xb = (size_t)xdb * p.NCH * stride * xhb * p.PLC * p.XWb * 16;
            const size_t yb = (size_t)bho * ydb * p.NW * p.NPLy * WG_YP * 16;
            if (2 * (xb + yb) + 256 > smem_cap) continue;
            const long long nbr = (long long)Nb * ((OD + bdo - 1) / bdo) * ((OH + bho - 1) / bho) * ((ow_ext + WG_KW - 1) / WG_KW);
            best = ci;
            if (nbr >= 2LL * want) break;   // enough bricks to feed every split; else keep shrinking
        }
        if (best < 0) continue;
        p.BDo = cand[best][0]; p.BHo = cand[best][1];
        p.XDb = p.NF > 1 ? p.BDo + K - 1 : (p.BDo - 1) * stride + tdspan + 1;
        p.XHb = (p.BHo - 1) * stride + thspan * p.R + p.R;
        p.XHu = (p.BHo - 1) * stride + thspan * p.R + (p.R < K ? p.R : K);
        p.x_sh = p.PLC * p.XWb; p.x_spar = p.XHb * p.x_sh; p.x_sc = stride * p.x_spar; p.x_sd = p.NCH * p.x_sc;
        p.x_nw = stride == 2 ? 2 * (WG_KW + 1) : p.XWb;
        p.YDb = p.BDo;
        p.nB = p.NF > 1 ? p.BDo + K - 1 : p.BDo;
        p.x_bytes = (uint32_t)((size_t)p.XDb * p.x_sd * 16);
        const uint32_t yb = (uint32_t)((size_t)p.BHo * p.YDb * p.NW * p.NPLy * WG_YP * 16);
        p.stage_bytes = (p.x_bytes + yb + 127) & ~127u;
        p.stages = (int)((smem_cap - 256) / p.stage_bytes);
        if (p.stages > 4) p.stages = 4;
        if (p.stages < 2) continue;
        found = true;
    }
    if (!found) return VG_ERR_UNSUPPORTED;
    p.bd_tiles = (OD + p.BDo - 1) / p.BDo; p.bh_tiles = (OH + p.BHo - 1) / p.BHo;
    p.bw_tiles = ((p.NW > 1 ? OW + K - 1 : OW) + WG_KW - 1) / WG_KW;
    const long long nbricks = (long long)Nb * p.bd_tiles * p.bh_tiles * p.bw_tiles;
    if (nbricks > 0x3fffffff) return VG_ERR_UNSUPPORTED;
    p.nbricks = (int)nbricks;
    p.ksplit = 148 / p.tiles;
    if (p.ksplit < 1) p.ksplit = 1;
    if (p.ksplit > p.nbricks) p.ksplit = p.nbricks;
    uint32_t cols = 32;
    while (cols < (uint32_t)(p.TPC * p.NCH * p.Nmma)) cols <<= 1;
    if (cols > 512) return VG_ERR_UNSUPPORTED;
    p.tmem_cols = cols;

    static int ca = -1;   // VG_CPASYNC=ca: gather through L1 (A/B testing)
    if (ca < 0) {
        const char* e = getenv("VG_CPASYNC");
        ca = (e && e[0] == 'c' && e[1] == 'a') ? 1 : 0;
    }
    p.ca = ca;
    static VgPerDevice attr_done;
    if (!attr_done.done()) {
        if (cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) return VG_ERR_CUDA;
        attr_done.mark();
    }
    const size_t smem = (size_t)p.stages * p.stage_bytes + 256;
    wgrad_tc_kernel<<<p.tiles * p.ksplit, WG_THREADS, smem, stream>>>(p);
    if (getenv("VG_DEBUG")) {
        cudaError_t e = cudaPeekAtLastError();
        fprintf(stderr, "[wgrad_tc] Cx=%d Cy=%d K=%d M=%d CC=%d NCH=%d CO=%d NF=%d NW=%d N=%d TPC=%d nsets=%d tiles=%d ksplit=%d brick=%dx%d stages=%d stage=%uB tmem=%u : %s\n",
                Cx, Cy, K, p.M, p.CC, p.NCH, p.CO, p.NF, p.NW, p.Nmma, p.TPC, p.nsets, p.tiles, p.ksplit, p.BDo, p.BHo, p.stages, p.stage_bytes,
                p.tmem_cols, cudaGetErrorString(e));
    }
    VG_LAUNCHED(1);
    g_vg_wg_tc_launches++;
    return VG_OK;
}
