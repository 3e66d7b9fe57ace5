// clDice soft-skeleton (soft_erode / soft_dilate / soft_skel), forward and backward.
//
// Replaces the Keras MaxPool3D graph of clDice_func.py:8-80 (reference).  fp32, one channel,
// volumes stored [N][D][H][W].  Level j of the forward pass fuses
//     e_{j+1} = erode(e_j);  o_j = dilate(e_{j+1});  delta_j = relu(e_j - o_j);
//     skel_j  = skel_{j-1} + relu(delta_j - skel_{j-1} * delta_j)        (skel_0 = delta_0)
// into one shared-memory-tiled stencil pass (halo 2).  The erode inside the reference's
// soft_open of iteration j is the image of iteration j+1, so iters+1 erodes suffice (bitwise
// identical to the reference's 2*iters+1).  min/max/sub/mul are evaluated with explicit
// round-to-nearest intrinsics (no FMA contraction) so the result is bit-exact.
//
// The backward pass walks the levels in reverse in gather form (no atomics): a per-window
// winner index (first extreme in scan order) is recomputed in shared memory and every voxel
// sums the upstream gradients of the windows it wins.
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int TX = 32, TY = 8, TZ = 8;          // output tile per block
constexpr int H2X = TX + 4, H2Y = TY + 4, H2Z = TZ + 4;  // halo-2 tile
constexpr int H1X = TX + 2, H1Y = TY + 2, H1Z = TZ + 2;  // halo-1 tile
constexpr int NTHREADS = 256;

struct Vol {
    int N, D, H, W;
};

__device__ __forceinline__ bool inside(const Vol& v, int z, int y, int x) {
    return (unsigned)z < (unsigned)v.D && (unsigned)y < (unsigned)v.H && (unsigned)x < (unsigned)v.W;
}

// 19-voxel erosion neighbourhood = union of the (3,3,1), (3,1,3), (1,3,3) planes:
// all offsets in {-1,0,1}^3 with at most two non-zero components.
__device__ __forceinline__ bool in_n19(int dz, int dy, int dx) { return !(dz != 0 && dy != 0 && dx != 0); }

// load e (halo 2) into shared memory; out-of-volume voxels get `oov`
__device__ __forceinline__ void load_halo2(const float* __restrict__ e, const Vol& v, int z0, int y0, int x0,
                                           float* sA, float oov) {
    for (int i = threadIdx.x; i < H2Z * H2Y * H2X; i += NTHREADS) {
        int lx = i % H2X, ly = (i / H2X) % H2Y, lz = i / (H2X * H2Y);
        int z = z0 + lz - 2, y = y0 + ly - 2, x = x0 + lx - 2;
        sA[i] = inside(v, z, y, x) ? __ldg(e + ((size_t)z * v.H + y) * v.W + x) : oov;
    }
}

__global__ void __launch_bounds__(NTHREADS)
skel_level_fwd_kernel(const float* __restrict__ e_in, float* __restrict__ e_out, const float* __restrict__ skel_in,
                      float* __restrict__ skel_out, Vol v, int tiles_x, int tiles_y, int tiles_z, int first) {
    __shared__ float sA[H2Z * H2Y * H2X];
    __shared__ float sB[H1Z * H1Y * H1X];
    int t = blockIdx.x;
    int tx = t % tiles_x, ty = (t / tiles_x) % tiles_y, tz = (t / (tiles_x * tiles_y)) % tiles_z;
    int n = t / (tiles_x * tiles_y * tiles_z);
    size_t voff = (size_t)n * v.D * v.H * v.W;
    int z0 = tz * TZ, y0 = ty * TY, x0 = tx * TX;
    load_halo2(e_in + voff, v, z0, y0, x0, sA, INFINITY);
    __syncthreads();
    // e_{j+1} on the halo-1 tile; out-of-volume -> -inf so the dilation ignores it
    for (int i = threadIdx.x; i < H1Z * H1Y * H1X; i += NTHREADS) {
        int lx = i % H1X, ly = (i / H1X) % H1Y, lz = i / (H1X * H1Y);
        int z = z0 + lz - 1, y = y0 + ly - 1, x = x0 + lx - 1;
        float m = -INFINITY;
        if (inside(v, z, y, x)) {
            m = INFINITY;
#pragma unroll
            for (int dz = -1; dz <= 1; dz++)
#pragma unroll
                for (int dy = -1; dy <= 1; dy++)
#pragma unroll
                    for (int dx = -1; dx <= 1; dx++)
                        if (in_n19(dz, dy, dx))
                            m = fminf(m, sA[((lz + 1 + dz) * H2Y + (ly + 1 + dy)) * H2X + (lx + 1 + dx)]);
        }
        sB[i] = m;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < TZ * TY * TX; i += NTHREADS) {
        int lx = i % TX, ly = (i / TX) % TY, lz = i / (TX * TY);
        int z = z0 + lz, y = y0 + ly, x = x0 + lx;
        if (!inside(v, z, y, x)) continue;
        float o = -INFINITY;
#pragma unroll
        for (int dz = 0; dz <= 2; dz++)
#pragma unroll
            for (int dy = 0; dy <= 2; dy++)
#pragma unroll
                for (int dx = 0; dx <= 2; dx++) o = fmaxf(o, sB[((lz + dz) * H1Y + (ly + dy)) * H1X + (lx + dx)]);
        float ej = sA[((lz + 2) * H2Y + (ly + 2)) * H2X + (lx + 2)];
        float delta = fmaxf(__fsub_rn(ej, o), 0.f);
        size_t g = voff + ((size_t)z * v.H + y) * v.W + x;
        float s;
        if (first) {
            s = delta;
        } else {
            float sp = skel_in[g];
            s = __fadd_rn(sp, fmaxf(__fsub_rn(delta, __fmul_rn(sp, delta)), 0.f));
        }
        skel_out[g] = s;
        e_out[g] = sB[((lz + 1) * H1Y + (ly + 1)) * H1X + (lx + 1)];
    }
}

// a_j and G_{j-1} from G_j, skel_{j-1}, e_j, e_{j+1} (see header comment / DESIGN.md)
__global__ void __launch_bounds__(NTHREADS)
skel_bwd_coeff_kernel(const float* __restrict__ G, const float* __restrict__ skel_prev, const float* __restrict__ ej,
                      const float* __restrict__ ej1, float* __restrict__ a_out, float* __restrict__ G_out, Vol v,
                      int tiles_x, int tiles_y, int tiles_z, int first) {
    __shared__ float sB[H1Z * H1Y * H1X];
    int t = blockIdx.x;
    int tx = t % tiles_x, ty = (t / tiles_x) % tiles_y, tz = (t / (tiles_x * tiles_y)) % tiles_z;
    int n = t / (tiles_x * tiles_y * tiles_z);
    size_t voff = (size_t)n * v.D * v.H * v.W;
    int z0 = tz * TZ, y0 = ty * TY, x0 = tx * TX;
    for (int i = threadIdx.x; i < H1Z * H1Y * H1X; i += NTHREADS) {
        int lx = i % H1X, ly = (i / H1X) % H1Y, lz = i / (H1X * H1Y);
        int z = z0 + lz - 1, y = y0 + ly - 1, x = x0 + lx - 1;
        sB[i] = inside(v, z, y, x) ? __ldg(ej1 + voff + ((size_t)z * v.H + y) * v.W + x) : -INFINITY;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < TZ * TY * TX; i += NTHREADS) {
        int lx = i % TX, ly = (i / TX) % TY, lz = i / (TX * TY);
        int z = z0 + lz, y = y0 + ly, x = x0 + lx;
        if (!inside(v, z, y, x)) continue;
        float o = -INFINITY;
#pragma unroll
        for (int dz = 0; dz <= 2; dz++)
#pragma unroll
            for (int dy = 0; dy <= 2; dy++)
#pragma unroll
                for (int dx = 0; dx <= 2; dx++) o = fmaxf(o, sB[((lz + dz) * H1Y + (ly + dy)) * H1X + (lx + dx)]);
        size_t g = voff + ((size_t)z * v.H + y) * v.W + x;
        float diff = __fsub_rn(ej[g], o);
        float delta = fmaxf(diff, 0.f);
        float Gj = G[g];
        float dd;
        if (first) {
            dd = Gj;
        } else {
            float sp = skel_prev[g];
            float u = __fsub_rn(delta, __fmul_rn(sp, delta));
            float m = u > 0.f ? 1.f : 0.f;
            dd = Gj * m * (1.f - sp);
            G_out[g] = Gj * (1.f - m * delta);
        }
        a_out[g] = diff > 0.f ? dd : 0.f;
    }
}

// D_j[q] = a_j[q] + sum_{p in N19(q)} [argmin_p(e_j)==q] D_{j+1}[p] - sum_{p in N27(q)} [argmax_p(e_j)==q] a_{j-1}[p]
__global__ void __launch_bounds__(NTHREADS)
skel_bwd_route_kernel(const float* __restrict__ ej, const float* __restrict__ a_j, const float* __restrict__ D_next,
                      const float* __restrict__ a_prev, float* __restrict__ D_out, Vol v, int tiles_x, int tiles_y,
                      int tiles_z) {
    extern __shared__ __align__(16) unsigned char route_smem[];
    float* sA = reinterpret_cast<float*>(route_smem);
    float* sD = sA + H2Z * H2Y * H2X;
    float* sP = sD + H1Z * H1Y * H1X;
    unsigned char* wmin = reinterpret_cast<unsigned char*>(sP + H1Z * H1Y * H1X);
    unsigned char* wmax = wmin + H1Z * H1Y * H1X;
    int t = blockIdx.x;
    int tx = t % tiles_x, ty = (t / tiles_x) % tiles_y, tz = (t / (tiles_x * tiles_y)) % tiles_z;
    int n = t / (tiles_x * tiles_y * tiles_z);
    size_t voff = (size_t)n * v.D * v.H * v.W;
    int z0 = tz * TZ, y0 = ty * TY, x0 = tx * TX;
    // NaN sentinel outside the volume: never wins a min or a max comparison below
    load_halo2(ej + voff, v, z0, y0, x0, sA, NAN);
    for (int i = threadIdx.x; i < H1Z * H1Y * H1X; i += NTHREADS) {
        int lx = i % H1X, ly = (i / H1X) % H1Y, lz = i / (H1X * H1Y);
        int z = z0 + lz - 1, y = y0 + ly - 1, x = x0 + lx - 1;
        bool in = inside(v, z, y, x);
        size_t g = voff + ((size_t)z * v.H + y) * v.W + x;
        sD[i] = (in && D_next) ? __ldg(D_next + g) : 0.f;
        sP[i] = (in && a_prev) ? __ldg(a_prev + g) : 0.f;
    }
    __syncthreads();
    // winners of every window centred in the halo-1 tile (index = (dz+1)*9+(dy+1)*3+(dx+1))
    for (int i = threadIdx.x; i < H1Z * H1Y * H1X; i += NTHREADS) {
        int lx = i % H1X, ly = (i / H1X) % H1Y, lz = i / (H1X * H1Y);
        float bmin = INFINITY, bmax = -INFINITY;
        int imin = 255, imax = 255;
#pragma unroll
        for (int dz = -1; dz <= 1; dz++)
#pragma unroll
            for (int dy = -1; dy <= 1; dy++)
#pragma unroll
                for (int dx = -1; dx <= 1; dx++) {
                    float val = sA[((lz + 1 + dz) * H2Y + (ly + 1 + dy)) * H2X + (lx + 1 + dx)];
                    int idx = (dz + 1) * 9 + (dy + 1) * 3 + (dx + 1);
                    if (val > bmax) { bmax = val; imax = idx; }
                    if (in_n19(dz, dy, dx) && val < bmin) { bmin = val; imin = idx; }
                }
        wmin[i] = (unsigned char)imin;
        wmax[i] = (unsigned char)imax;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < TZ * TY * TX; i += NTHREADS) {
        int lx = i % TX, ly = (i / TX) % TY, lz = i / (TX * TY);
        int z = z0 + lz, y = y0 + ly, x = x0 + lx;
        if (!inside(v, z, y, x)) continue;
        size_t g = voff + ((size_t)z * v.H + y) * v.W + x;
        float acc = a_j ? a_j[g] : 0.f;
#pragma unroll
        for (int dz = -1; dz <= 1; dz++)
#pragma unroll
            for (int dy = -1; dy <= 1; dy++)
#pragma unroll
                for (int dx = -1; dx <= 1; dx++) {
                    // window centre p = q + (dz,dy,dx); q sits at offset (-dz,-dy,-dx) inside p's window
                    int pi = ((lz + 1 + dz) * H1Y + (ly + 1 + dy)) * H1X + (lx + 1 + dx);
                    int self = (1 - dz) * 9 + (1 - dy) * 3 + (1 - dx);
                    if (wmin[pi] == self) acc += sD[pi];   // sD is 0 outside the volume
                    if (wmax[pi] == self) acc -= sP[pi];
                }
        D_out[g] = acc;
    }
}

// ------------------------------------------------------------------------------------------ z-marching level kernels
// The tile kernels above spend their time on shared-memory window scans (19 + 27 loads per voxel plus the halo re-computation:
// 58 LDS per output voxel, 20 % of the HBM model).  min / max are exact and associative, so the same bits come out of a
// SEPARABLE evaluation:  erode = min(Pxy, Pxz, Pyz) with  rx = min3_x(e), ry = min3_y(e):
//     Pxy = min3_y(rx),  Pxz = min3_z(rx),  Pyz = min3_z(ry);      dilate = max3_z(max3_y(max3_x(e1))).
// A thread owns one (x, y) column and marches along z with the 3-deep z windows in registers; x neighbours come from warp
// shuffles (a warp = 32 consecutive x, 28 outputs + halo 2), y neighbours from one shared-memory row exchange per plane
// (three arrays, double-buffered, ONE barrier per plane).  Per plane and thread: 1 global load, 3 STS, 6 LDS, 4 SHFL.
constexpr int MW_OUT = 28, MH_OUT = 16, M_ROWS = MH_OUT + 4, M_THREADS = 32 * M_ROWS;

__global__ void __launch_bounds__(M_THREADS, 2)
skel_level_fwd_march_kernel(const float* __restrict__ e_in, float* __restrict__ e_out, const float* __restrict__ skel_in,
                            float* __restrict__ skel_out, Vol v, int tiles_x, int tiles_y, int zchunks, int ZL, int first) {
    __shared__ float sV[2][M_ROWS + 2][32], sRX[2][M_ROWS + 2][32], sMX[2][M_ROWS + 2][32];
    const int lane = threadIdx.x & 31, wy = threadIdx.x >> 5;
    int t = blockIdx.x;
    const int tx = t % tiles_x; t /= tiles_x;
    const int ty = t % tiles_y; t /= tiles_y;
    const int zc = t % zchunks;
    const int n = t / zchunks;
    const int gx = tx * MW_OUT - 2 + lane, gy = ty * MH_OUT - 2 + wy;
    const int z0 = zc * ZL, zend = min(z0 + ZL, v.D);
    const bool col_in = (unsigned)gx < (unsigned)v.W && (unsigned)gy < (unsigned)v.H;
    const bool out_thread = col_in && lane >= 2 && lane < 30 && wy >= 2 && wy < M_ROWS - 2;
    const size_t HW = (size_t)v.H * v.W;
    const size_t col = (size_t)n * v.D * HW + (size_t)(col_in ? gy : 0) * v.W + (col_in ? gx : 0);
    if (wy == 0) {
        for (int b = 0; b < 2; b++) {
            sV[b][0][lane] = INFINITY; sV[b][M_ROWS + 1][lane] = INFINITY;
            sRX[b][0][lane] = INFINITY; sRX[b][M_ROWS + 1][lane] = INFINITY;
            sMX[b][0][lane] = -INFINITY; sMX[b][M_ROWS + 1][lane] = -INFINITY;
        }
    }
    auto loadv = [&](int pz) -> float {
        return (col_in && pz >= 0 && pz < v.D && pz < zend + 2) ? __ldg(e_in + col + (size_t)pz * HW) : INFINITY;
    };
    float vnext = loadv(z0 - 2);
    float v1 = INFINITY, v2 = INFINITY, v3 = INFINITY;          // e at planes pz-1, pz-2, pz-3
    float rx1 = INFINITY, rx2 = INFINITY, ry1 = INFINITY, ry2 = INFINITY, rxy1 = INFINITY;
    float e1_1 = -INFINITY, e1_2 = -INFINITY;                   // eroded image at planes pz-2, pz-3
    float mxy3 = -INFINITY, mxy4 = -INFINITY, mx_prev = -INFINITY;
    const unsigned FULL = 0xffffffffu;
    for (int i = 0; i < ZL + 5; i++) {
        const int pz = z0 - 2 + i, buf = i & 1, zo = pz - 3;
        const float vA = vnext;
        vnext = loadv(pz + 1);
        const bool out_ok = out_thread && zo >= z0 && zo < zend;
        float sp = 0.f;
        if (out_ok && !first) sp = skel_in[col + (size_t)zo * HW];
        const float rxA = fminf(vA, fminf(__shfl_up_sync(FULL, vA, 1), __shfl_down_sync(FULL, vA, 1)));
        sV[buf][wy + 1][lane] = vA;
        sRX[buf][wy + 1][lane] = rxA;
        sMX[buf][wy + 1][lane] = mx_prev;
        __syncthreads();
        const float ryA = fminf(vA, fminf(sV[buf][wy][lane], sV[buf][wy + 2][lane]));
        const float rxyA = fminf(rxA, fminf(sRX[buf][wy][lane], sRX[buf][wy + 2][lane]));
        const float mxyC = fmaxf(mx_prev, fmaxf(sMX[buf][wy][lane], sMX[buf][wy + 2][lane]));   // plane pz-2
        // eroded image at plane b = pz-1 (out-of-volume voxels must not take part in the dilation)
        float e1b = fminf(rxy1, fminf(fminf(rx2, fminf(rx1, rxA)), fminf(ry2, fminf(ry1, ryA))));
        if (!(col_in && pz - 1 >= 0 && pz - 1 < v.D)) e1b = -INFINITY;
        const float mxb = fmaxf(e1b, fmaxf(__shfl_up_sync(FULL, e1b, 1), __shfl_down_sync(FULL, e1b, 1)));
        if (out_ok) {
            const float o = fmaxf(mxy4, fmaxf(mxy3, mxyC));
            const float delta = fmaxf(__fsub_rn(v3, o), 0.f);
            const float sk = first ? delta : __fadd_rn(sp, fmaxf(__fsub_rn(delta, __fmul_rn(sp, delta)), 0.f));
            const size_t g = col + (size_t)zo * HW;
            skel_out[g] = sk;
            e_out[g] = e1_2;
        }
        v3 = v2; v2 = v1; v1 = vA;
        rx2 = rx1; rx1 = rxA; ry2 = ry1; ry1 = ryA; rxy1 = rxyA;
        e1_2 = e1_1; e1_1 = e1b;
        mxy4 = mxy3; mxy3 = mxyC;
        mx_prev = mxb;
    }
}

// chunk length along z for the marching kernels: fill the resident-block slots evenly (2 blocks per SM) at a small halo cost
inline int pick_zl(int D, long long columns, int halo_iters) {
    int best = D;
    double best_eff = 0.0;
    for (int zl = 8; zl <= D; zl += 4) {
        const long long blocks = columns * ((D + zl - 1) / zl);
        const long long waves = (blocks + 295) / 296;
        const double eff = ((double)blocks / (double)(waves * 296)) * ((double)zl / (double)(zl + halo_iters));
        if (eff > best_eff + 1e-9) { best_eff = eff; best = zl; }
    }
    if (D < 8) best = D;
    return best;
}
// VG_SKEL=tile selects the shared-memory tile kernels (A/B testing and cross-checks)
inline bool skel_march() {
    static int m = -1;
    if (m < 0) {
        const char* e = getenv("VG_SKEL");
        m = (e && e[0] == 't') ? 0 : 1;
    }
    return m == 1;
}

constexpr size_t ROUTE_SMEM = (size_t)(H2Z * H2Y * H2X + 2 * H1Z * H1Y * H1X) * sizeof(float) + 2 * H1Z * H1Y * H1X;

inline void tiles_of(const Vol& v, int& tx, int& ty, int& tz) {
    tx = vg_cdiv(v.W, TX);
    ty = vg_cdiv(v.H, TY);
    tz = vg_cdiv(v.D, TZ);
}

}  // namespace

extern "C" {

// E: [iters+2][N*D*H*W] erosion pyramid (E[0] is filled with a copy of x), S: [iters+1][...] skeleton history.
// The soft skeleton is S[iters].
int vg_soft_skel_fwd(const float* x, float* E, float* S, int N, int D, int H, int W, int iters, void* stream) {
    VG_REQUIRE(x && E && S && N > 0 && D > 0 && H > 0 && W > 0 && iters >= 0);
    cudaStream_t st = (cudaStream_t)stream;
    Vol v{N, D, H, W};
    size_t nv = (size_t)N * D * H * W;
    if (cudaMemcpyAsync(E, x, nv * sizeof(float), cudaMemcpyDeviceToDevice, st) != cudaSuccess) return VG_ERR_CUDA;
    int tx, ty, tz;
    tiles_of(v, tx, ty, tz);
    int blocks = tx * ty * tz * N;
    const int mtx = vg_cdiv(W, MW_OUT), mty = vg_cdiv(H, MH_OUT);
    const int ZL = pick_zl(D, (long long)N * mtx * mty, 5), zch = vg_cdiv(D, ZL);
    for (int j = 0; j <= iters; j++) {
        if (skel_march()) {
            skel_level_fwd_march_kernel<<<N * mtx * mty * zch, M_THREADS, 0, st>>>(E + (size_t)j * nv, E + (size_t)(j + 1) * nv,
                                                                                  j ? S + (size_t)(j - 1) * nv : nullptr,
                                                                                  S + (size_t)j * nv, v, mtx, mty, zch, ZL, j == 0);
        } else {
            skel_level_fwd_kernel<<<blocks, NTHREADS, 0, st>>>(E + (size_t)j * nv, E + (size_t)(j + 1) * nv,
                                                              j ? S + (size_t)(j - 1) * nv : nullptr, S + (size_t)j * nv, v,
                                                              tx, ty, tz, j == 0);
        }
        VG_LAUNCHED(1);
    }
    VG_CHECK_LAUNCH();
    return VG_OK;
}

size_t vg_soft_skel_bwd_workspace_bytes(int N, int D, int H, int W) { return (size_t)6 * N * D * H * W * sizeof(float); }

// gskel: dL/d skel (same shape as x); dx: dL/dx.  E, S as written by vg_soft_skel_fwd.
int vg_soft_skel_bwd(const float* E, const float* S, const float* gskel, float* dx, void* workspace,
                     size_t workspace_bytes, int N, int D, int H, int W, int iters, void* stream) {
    VG_REQUIRE(E && S && gskel && dx && workspace);
    if (workspace_bytes < vg_soft_skel_bwd_workspace_bytes(N, D, H, W)) return VG_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    Vol v{N, D, H, W};
    size_t nv = (size_t)N * D * H * W;
    float* ws = (float*)workspace;
    float* Gb[2] = {ws, ws + nv};
    float* Ab[2] = {ws + 2 * nv, ws + 3 * nv};
    float* Db[2] = {ws + 4 * nv, ws + 5 * nv};
    int tx, ty, tz;
    tiles_of(v, tx, ty, tz);
    int blocks = tx * ty * tz * N;
    if (cudaFuncSetAttribute(skel_bwd_route_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ROUTE_SMEM) != cudaSuccess)
        return VG_ERR_CUDA;
    const int k = iters;
    auto Ej = [&](int j) { return E + (size_t)j * nv; };
    // a_j lives in Ab[j&1]; G_{j} (for j<k) in Gb[j&1]; D_j in Db[j&1]
    const float* Gcur = gskel;
    // (a z-marching version of the coefficient kernel was measured slower than the tile kernel: 5 global accesses and one
    // barrier per plane leave nothing to amortise -- 10.8 -> 13.3 ms at 256^3, iters 15 -- and was dropped)
    auto coeff = [&](const float* Gin, const float* sprev, const float* e0, const float* e1, float* aout, float* gout, int first) {
        skel_bwd_coeff_kernel<<<blocks, NTHREADS, 0, st>>>(Gin, sprev, e0, e1, aout, gout, v, tx, ty, tz, first);
        VG_LAUNCHED(1);
    };
    coeff(Gcur, k ? S + (size_t)(k - 1) * nv : nullptr, Ej(k), Ej(k + 1), Ab[k & 1], Gb[(k + 1) & 1], k == 0);
    Gcur = Gb[(k + 1) & 1];  // now holds G_{k-1}
    skel_bwd_route_kernel<<<blocks, NTHREADS, ROUTE_SMEM, st>>>(Ej(k + 1), nullptr, nullptr, Ab[k & 1], Db[(k + 1) & 1], v, tx, ty,
                                                       tz); VG_LAUNCHED(1);
    for (int j = k; j >= 0; j--) {
        if (j >= 1) {
            int jj = j - 1;
            coeff(Gcur, jj ? S + (size_t)(jj - 1) * nv : nullptr, Ej(jj), Ej(jj + 1), Ab[jj & 1], Gb[(jj + 1) & 1], jj == 0);
            Gcur = Gb[(jj + 1) & 1];
        }
        float* out = j == 0 ? dx : Db[j & 1];
        skel_bwd_route_kernel<<<blocks, NTHREADS, ROUTE_SMEM, st>>>(Ej(j), Ab[j & 1], Db[(j + 1) & 1], j >= 1 ? Ab[(j - 1) & 1] : nullptr,
                                                           out, v, tx, ty, tz); VG_LAUNCHED(1);
    }
    VG_CHECK_LAUNCH();
    return VG_OK;
}

}  // extern "C"
