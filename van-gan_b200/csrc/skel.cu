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
    for (int j = 0; j <= iters; j++) {
        skel_level_fwd_kernel<<<blocks, NTHREADS, 0, st>>>(E + (size_t)j * nv, E + (size_t)(j + 1) * nv,
                                                          j ? S + (size_t)(j - 1) * nv : nullptr, S + (size_t)j * nv, v,
                                                          tx, ty, tz, j == 0); VG_LAUNCHED(1);
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
    skel_bwd_coeff_kernel<<<blocks, NTHREADS, 0, st>>>(Gcur, k ? S + (size_t)(k - 1) * nv : nullptr, Ej(k), Ej(k + 1),
                                                       Ab[k & 1], Gb[(k + 1) & 1], v, tx, ty, tz, k == 0); VG_LAUNCHED(1);
    Gcur = Gb[(k + 1) & 1];  // now holds G_{k-1}
    skel_bwd_route_kernel<<<blocks, NTHREADS, ROUTE_SMEM, st>>>(Ej(k + 1), nullptr, nullptr, Ab[k & 1], Db[(k + 1) & 1], v, tx, ty,
                                                       tz); VG_LAUNCHED(1);
    for (int j = k; j >= 0; j--) {
        if (j >= 1) {
            int jj = j - 1;
            skel_bwd_coeff_kernel<<<blocks, NTHREADS, 0, st>>>(Gcur, jj ? S + (size_t)(jj - 1) * nv : nullptr, Ej(jj),
                                                               Ej(jj + 1), Ab[jj & 1], Gb[(jj + 1) & 1], v, tx, ty, tz,
                                                               jj == 0); VG_LAUNCHED(1);
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
