// Layout helpers, optimizer and sliding-window stitching kernels (all HBM-bound streaming passes).
//
// Replaces: UpSampling3D(2)+concatenate (resunet_model.py:176,181) and its gradient; the
// discriminator's input ReflectionPadding3D+GaussianNoise (discriminator.py:50-52); gradient
// accumulation done by tf.GradientTape; Keras OptimizerV2 Adam with clipnorm (vangan.py:220-235,
// 426-438); the numpy accumulate / divide / min-max of GanMonitor.stitch_subvolumes
// (custom_callback.py:165-166,177-183,192,202).
#include "common.cuh"

unsigned long long g_vg_launches = 0;

namespace {

constexpr int NT = 256;

// A block walks whole output rows (n, d, h): the row decode is one 32-bit division chain per row instead of five 64-bit
// divisions per 16-byte packet (the first version was instruction-bound at 39 % of HBM bandwidth).
__global__ void __launch_bounds__(NT) upsample_concat_kernel(const bf16* __restrict__ lo, const bf16* __restrict__ skip,
                                                             bf16* __restrict__ out, int N, int D, int H, int W, int C0, int C1) {
    const int C = C0 + C1, cg = C / 8, cg0 = C0 / 8, cg1 = cg - cg0;
    const int D2 = 2 * D, H2 = 2 * H, W2 = 2 * W;
    const int rows = N * D2 * H2, per_row = W2 * cg;
    for (int row = blockIdx.x; row < rows; row += gridDim.x) {
        const int h = row % H2, r = row / H2;
        const int d = r % D2, n = r / D2;
        const uint4* lrow = reinterpret_cast<const uint4*>(lo + (((size_t)n * D + (d >> 1)) * H + (h >> 1)) * W * C0);
        const uint4* srow = reinterpret_cast<const uint4*>(skip + (size_t)row * W2 * C1);
        uint4* orow = reinterpret_cast<uint4*>(out + (size_t)row * W2 * C);
        // 4 independent 16-byte packets per thread in flight (a 128-voxel row of 48 channels is 768 packets = 3 per thread)
        for (int i0 = threadIdx.x; i0 < per_row; i0 += 4 * NT) {
            uint4 v[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int i = i0 + u * NT;
                if (i < per_row) {
                    const int w = i / cg, c8 = i - w * cg;
                    v[u] = c8 < cg0 ? __ldg(lrow + (w >> 1) * cg0 + c8) : __ldg(srow + w * cg1 + (c8 - cg0));
                }
            }
#pragma unroll
            for (int u = 0; u < 4; u++)
                if (i0 + u * NT < per_row) orow[i0 + u * NT] = v[u];
        }
    }
}

// dlo[n,d,h,w] = sum of the 8 fine voxels' first C0 channels.  A block owns coarse rows (n, d, h).
__global__ void __launch_bounds__(NT) upsample_concat_bwd_lo_kernel(const bf16* __restrict__ dcat, bf16* __restrict__ dlo, int N,
                                                                    int D, int H, int W, int C0, int C1) {
    const int C = C0 + C1, cg0 = C0 / 8;
    const int H2 = 2 * H, W2 = 2 * W, D2 = 2 * D;
    const int rows = N * D * H, per_row = W * cg0;
    for (int row = blockIdx.x; row < rows; row += gridDim.x) {
        const int h = row % H, r = row / H;
        const int d = r % D, n = r / D;
        const bf16* src = dcat + (((size_t)n * D2 + 2 * d) * H2 + 2 * h) * W2 * C;
        bf16* dst = dlo + (size_t)row * W * C0;
        for (int i = threadIdx.x; i < per_row; i += NT) {
            const int w = i / cg0, c8 = i - w * cg0;
            uint4 pk[8];
#pragma unroll
            for (int dz = 0; dz < 2; dz++)
#pragma unroll
                for (int dy = 0; dy < 2; dy++)
#pragma unroll
                    for (int dx = 0; dx < 2; dx++)
                        pk[dz * 4 + dy * 2 + dx] =
                            __ldg(reinterpret_cast<const uint4*>(src + (((size_t)dz * H2 + dy) * W2 + 2 * w + dx) * C + c8 * 8));
            float a[8];
#pragma unroll
            for (int k = 0; k < 8; k++) a[k] = 0.f;
#pragma unroll
            for (int q = 0; q < 8; q++) {
                float f[8];
                unpack8(pk[q], f);
#pragma unroll
                for (int k = 0; k < 8; k++) a[k] += f[k];
            }
            store8<bf16>(dst + (size_t)i * 8, a);
        }
    }
}

__global__ void __launch_bounds__(NT) upsample_concat_bwd_skip_kernel(const bf16* __restrict__ dcat, bf16* __restrict__ dskip,
                                                                      size_t V2, int C0, int C1, int accumulate) {
    const int C = C0 + C1, cg1 = C1 / 8;
    const size_t total = V2 * cg1;
    const bool small = total <= 0xffffffffull;   // 32-bit division on the packet index (a 64-bit one costs more than the 16-byte copy)
    for (size_t i = (size_t)blockIdx.x * NT + threadIdx.x; i < total; i += (size_t)gridDim.x * NT) {
        const size_t v = small ? (size_t)((unsigned)i / (unsigned)cg1) : i / (unsigned)cg1;
        const int c8 = (int)(i - v * cg1);
        uint4 p = __ldg(reinterpret_cast<const uint4*>(dcat + v * C + C0 + c8 * 8));
        if (accumulate) {
            float f[8], o[8];
            unpack8(p, f);
            load8<bf16>(dskip + i * 8, o);
#pragma unroll
            for (int k = 0; k < 8; k++) f[k] += o[k];
            p = pack8(f);
        }
        reinterpret_cast<uint4*>(dskip)[i] = p;
    }
}

// A warp owns padded rows (n, pd, ph): three 32-bit divisions per ROW instead of four 64-bit ones per voxel; lanes walk pw.
__global__ void __launch_bounds__(NT) pad_noise_kernel(const float* __restrict__ x, float* __restrict__ y, int N, int D, int H,
                                                       int W, const float* __restrict__ noise, float noise_std,
                                                       unsigned long long seed, const unsigned long long* __restrict__ seed_dev) {
    if (seed_dev) seed += *seed_dev;   // per-step offset kept in device memory (CUDA-graph replays draw fresh noise)
    const int PD = D + 2, PH = H + 2, PW = W + 2;
    const unsigned rows = (unsigned)N * PD * PH;   // < 2^31 (checked by the caller)
    const int lane = threadIdx.x & 31;
    const unsigned nwarps = gridDim.x * (NT / 32);
    for (unsigned row = blockIdx.x * (NT / 32) + (threadIdx.x >> 5); row < rows; row += nwarps) {
        const unsigned ph = row % PH, r = row / PH;
        const unsigned pd = r % PD, n = r / PD;
        const int d = reflect1((int)pd - 1, D), h = reflect1((int)ph - 1, H);
        const float* xr = x + (((size_t)n * D + d) * H + h) * W;
        const size_t i0 = (size_t)row * PW;
        for (int pw0 = lane; pw0 < PW; pw0 += 128) {
            float v[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {   // four independent loads in flight per lane
                const int pw = pw0 + 32 * u;
                v[u] = pw < PW ? xr[reflect1(pw - 1, W)] : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int pw = pw0 + 32 * u;
                if (pw >= PW) continue;
                const size_t i = i0 + pw;   // flat index of the padded voxel: also the Philox counter (unchanged stream)
                if (noise) {
                    v[u] += noise[i];
                } else if (noise_std > 0.f) {
                    uint4 rr = philox4x32(make_uint4((uint32_t)i, (uint32_t)(i >> 32), 2u, 0x56414e47u),
                                          make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
                    v[u] += noise_std * box_muller(rr.x, rr.y).x;
                }
                y[i] = v[u];
            }
        }
    }
}

__global__ void __launch_bounds__(NT) pad_fold_kernel(const float* __restrict__ dy, float* __restrict__ dx, int N, int D, int H,
                                                      int W, int accumulate) {
    const int PH = H + 2, PW = W + 2, PD = D + 2;
    size_t total = (size_t)N * D * H * W;
    for (size_t i = (size_t)blockIdx.x * NT + threadIdx.x; i < total; i += (size_t)gridDim.x * NT) {
        int w = (int)(i % W), h = (int)((i / W) % H), d = (int)((i / ((size_t)W * H)) % D);
        int n = (int)(i / ((size_t)W * H * D));
        int dd[3], hh[3], ww[3], nd = 1, nh = 1, nw = 1;
        dd[0] = d + 1; if (d == 1) dd[nd++] = 0; if (d == D - 2) dd[nd++] = D + 1;
        hh[0] = h + 1; if (h == 1) hh[nh++] = 0; if (h == H - 2) hh[nh++] = H + 1;
        ww[0] = w + 1; if (w == 1) ww[nw++] = 0; if (w == W - 2) ww[nw++] = W + 1;
        float s = 0.f;
        for (int a = 0; a < nd; a++)
            for (int b = 0; b < nh; b++)
                for (int c = 0; c < nw; c++) s += dy[(((size_t)n * PD + dd[a]) * PH + hh[b]) * PW + ww[c]];
        dx[i] = accumulate ? dx[i] + s : s;
    }
}

template <typename T>
__global__ void __launch_bounds__(NT) accumulate_kernel(T* __restrict__ a, const T* __restrict__ b, size_t n8, size_t n) {
    for (size_t i = (size_t)blockIdx.x * NT + threadIdx.x; i < n8; i += (size_t)gridDim.x * NT) {
        float fa[8], fb[8];
        load8<T>(a + i * 8, fa);
        load8<T>(b + i * 8, fb);
#pragma unroll
        for (int k = 0; k < 8; k++) fa[k] += fb[k];
        store8<T>(a + i * 8, fa);
    }
    if (blockIdx.x == 0)
        for (size_t i = n8 * 8 + threadIdx.x; i < n; i += NT) a[i] = (T)((float)a[i] + (float)b[i]);
}

__global__ void __launch_bounds__(NT) tanh_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                                                      float* __restrict__ out, size_t n) {
    for (size_t i = (size_t)blockIdx.x * NT + threadIdx.x; i < n; i += (size_t)gridDim.x * NT) {
        float t = y[i];
        out[i] = dy[i] * (1.f - t * t);
    }
}

// ------------------------------------------------------------------ clip-by-norm + Adam
__device__ __forceinline__ int find_seg(const long long* __restrict__ off, int nseg, long long i) {
    int lo = 0, hi = nseg - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (off[mid] <= i) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// per-variable squared norms: a warp owns runs of 4096 consecutive elements and keeps its partial sum while it stays inside one
// variable -- one double atomic per (run, variable) instead of one per 128 elements (the large variables serialised ~80 k atomics on
// a handful of addresses: 0.125 ms per network for 40 MB of reads)
__global__ void __launch_bounds__(NT) seg_sqnorm_kernel(const float* __restrict__ g, const long long* __restrict__ off, int nseg,
                                                        double* __restrict__ norms, long long total) {
    constexpr long long CH = 4096;
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * NT + threadIdx.x) >> 5, nwarps = ((long long)gridDim.x * NT) >> 5;
    for (long long chunk = warp * CH; chunk < total; chunk += nwarps * CH) {
        const long long cend = chunk + CH < total ? chunk + CH : total;
        int cur = -1;
        double acc = 0;
        for (long long base = chunk; base < cend; base += 128) {
            const int s0 = find_seg(off, nseg, base);
            const long long end = base + 128 < cend ? base + 128 : cend;
            const bool uniform = off[s0 + 1] >= end;     // warp-uniform
            if (uniform && s0 != cur) {
                acc = warp_sum_d(acc);
                if (lane == 0 && cur >= 0 && acc != 0.0) atomicAdd(norms + cur, acc);
                cur = s0;
                acc = 0;
            }
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const long long i = base + k * 32 + lane;
                if (i < end) {
                    const float v = g[i];
                    if (uniform) acc += (double)v * v;
                    else atomicAdd(norms + find_seg(off, nseg, i), (double)v * v);
                }
            }
        }
        acc = warp_sum_d(acc);
        if (lane == 0 && cur >= 0 && acc != 0.0) atomicAdd(norms + cur, acc);
    }
}

__global__ void __launch_bounds__(NT) clip_adam_kernel(float* __restrict__ w, const float* __restrict__ g, float* __restrict__ m,
                                                       float* __restrict__ v, const long long* __restrict__ off, int nseg,
                                                       const double* __restrict__ norms, float lr_t, float b1, float b2,
                                                       float eps, float clip, long long total, const float* __restrict__ lr_t_dev) {
    if (lr_t_dev) lr_t = *lr_t_dev;    // bias-corrected step size kept in device memory (changes every step under graph replay)
    // four consecutive parameters per thread: one segment search and 128-bit accesses when they belong to one variable
    const long long nq = (total + 3) / 4;
    for (long long q = (long long)blockIdx.x * NT + threadIdx.x; q < nq; q += (long long)gridDim.x * NT) {
        const long long i = 4 * q;
        const int lo = find_seg(off, nseg, i);
        if (i + 4 <= total && off[lo + 1] >= i + 4) {
            const float nrm = (float)sqrt(norms[lo]);
            const float scale = clip / fmaxf(nrm, clip);   // tf.clip_by_norm
            const float4 g4 = *reinterpret_cast<const float4*>(g + i);
            float4 m4 = *reinterpret_cast<const float4*>(m + i), v4 = *reinterpret_cast<const float4*>(v + i),
                   w4 = *reinterpret_cast<const float4*>(w + i);
            const float gi[4] = {g4.x * scale, g4.y * scale, g4.z * scale, g4.w * scale};
            float* mp = &m4.x; float* vp = &v4.x; float* wp = &w4.x;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                mp[k] = b1 * mp[k] + (1.f - b1) * gi[k];
                vp[k] = b2 * vp[k] + (1.f - b2) * gi[k] * gi[k];
                wp[k] -= lr_t * mp[k] / (sqrtf(vp[k]) + eps);
            }
            *reinterpret_cast<float4*>(m + i) = m4;
            *reinterpret_cast<float4*>(v + i) = v4;
            *reinterpret_cast<float4*>(w + i) = w4;
        } else {
            for (long long e = i; e < i + 4 && e < total; e++) {
                const int l2 = find_seg(off, nseg, e);
                const float nrm = (float)sqrt(norms[l2]);
                const float scale = clip / fmaxf(nrm, clip);
                const float gi = g[e] * scale;
                const float mi = b1 * m[e] + (1.f - b1) * gi;
                const float vi = b2 * v[e] + (1.f - b2) * gi * gi;
                m[e] = mi;
                v[e] = vi;
                w[e] -= lr_t * mi / (sqrtf(vi) + eps);
            }
        }
    }
}

// SpatialDropout3D mask: out[i] = (u_i >= rate) / (1 - rate), u_i from Philox keyed on seed (+ *seed_dev)
__global__ void dropout_mask_kernel(float* __restrict__ out, int n, float rate, unsigned long long seed,
                                    const unsigned long long* __restrict__ seed_dev) {
    if (seed_dev) seed += *seed_dev;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint4 r = philox4x32(make_uint4((uint32_t)i, 0u, 3u, 0x56414e47u), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    const float u = r.x * 2.3283064365386963e-10f;
    out[i] = u >= rate ? 1.f / (1.f - rate) : 0.f;
}

// ------------------------------------------------------------------ stitching
__global__ void __launch_bounds__(NT) stitch_accumulate_kernel(float* __restrict__ pred, float* __restrict__ cnt, int H, int W,
                                                               int D, const float* __restrict__ win, const int* __restrict__ starts,
                                                               int B, int kH, int kW, int kD, int pH, int pW, int pD) {
    const int iH = kH - 2 * pH, iW = kW - 2 * pW, iD = kD - 2 * pD;
    size_t per = (size_t)iH * iW * iD, total = per * B;
    for (size_t i = (size_t)blockIdx.x * NT + threadIdx.x; i < total; i += (size_t)gridDim.x * NT) {
        int b = (int)(i / per);
        size_t r = i % per;
        int z = (int)(r % iD) + pD, y = (int)((r / iD) % iW) + pW, x = (int)(r / ((size_t)iD * iW)) + pH;
        float v = win[(((size_t)b * kH + x) * kW + y) * kD + z];
        size_t o = (((size_t)(starts[3 * b] + x)) * W + starts[3 * b + 1] + y) * D + starts[3 * b + 2] + z;
        // windows of one batch may overlap -> atomics (fp32 add order is the only nondeterminism)
        atomicAdd(pred + o, v);
        atomicAdd(cnt + o, 1.0f);
    }
}

// win[b][x][y][z] = vol[starts[b] + (x,y,z)]   (window extraction, custom_callback.py:167-169)
__global__ void __launch_bounds__(NT) stitch_gather_kernel(const float* __restrict__ vol, int H, int W, int D, float* __restrict__ win,
                                                           const int* __restrict__ starts, int B, int kH, int kW, int kD) {
    size_t per = (size_t)kH * kW * kD, total = per * B;
    for (size_t i = (size_t)blockIdx.x * NT + threadIdx.x; i < total; i += (size_t)gridDim.x * NT) {
        int b = (int)(i / per);
        size_t r = i % per;
        int z = (int)(r % kD), y = (int)((r / kD) % kW), x = (int)(r / ((size_t)kD * kW));
        win[i] = vol[(((size_t)(starts[3 * b] + x)) * W + starts[3 * b + 1] + y) * D + starts[3 * b + 2] + z];
    }
}

// the same extraction from the UN-padded volume when the reference first pads it with np.pad(..., 'symmetric') (custom_callback.py:82-104):
// a padded coordinate p maps to sym(p - pad), sym(i) = -i-1 below 0 and 2n-1-i at or above n (pad < n), so the padded copy never exists
__device__ __forceinline__ int sym_idx(int i, int n) { return i < 0 ? -i - 1 : (i >= n ? 2 * n - 1 - i : i); }
__global__ void __launch_bounds__(NT) stitch_gather_sym_kernel(const float* __restrict__ vol, int row0, int H0, int W0, int D0, int xs, int ys, int zs,
                                                               float* __restrict__ win, const int* __restrict__ starts, int B, int kH, int kW,
                                                               int kD) {
    size_t per = (size_t)kH * kW * kD, total = per * B;
    for (size_t i = (size_t)blockIdx.x * NT + threadIdx.x; i < total; i += (size_t)gridDim.x * NT) {
        int b = (int)(i / per);
        size_t r = i % per;
        int z = (int)(r % kD), y = (int)((r / kD) % kW), x = (int)(r / ((size_t)kD * kW));
        const int sx = sym_idx(starts[3 * b] + x - xs, H0), sy = sym_idx(starts[3 * b + 1] + y - ys, W0), sz = sym_idx(starts[3 * b + 2] + z - zs, D0);
        win[i] = vol[((size_t)(sx - row0) * W0 + sy) * D0 + sz];
    }
}

// On-device input pipeline (dataset.py:205-251): tf.image.random_crop + random_spatial_augmentation on a (H, W, D, 1) volume.  The
// tf.image ops see a 4-D tensor as [batch, height, width, channels], so on a volume flip_left_right reverses axis 2 (D),
// flip_up_down reverses axis 1 (W) and rot90(k) turns the (W, D) plane k quarter turns counter-clockwise (np.rot90(axes=(1, 2))).
// out[x][y][z] = rot90_k( flip_ud( flip_lr( vol[x0 + x][y0 + .][z0 + .] ) ) ); odd k needs kW == kD.
__global__ void __launch_bounds__(NT) crop_augment_kernel(const float* __restrict__ vol, int H, int W, int D, float* __restrict__ out, int kH,
                                                          int kW, int kD, int x0, int y0, int z0, int flip_lr, int flip_ud, int rot_k) {
    const size_t total = (size_t)kH * kW * kD;
    for (size_t i = (size_t)blockIdx.x * NT + threadIdx.x; i < total; i += (size_t)gridDim.x * NT) {
        const int z = (int)(i % kD), y = (int)((i / kD) % kW), x = (int)(i / ((size_t)kD * kW));
        // invert the rotation: element (y, z) of rot90^k(m) comes from m at ...
        int sy, sz;
        switch (rot_k & 3) {
            case 0: sy = y; sz = z; break;
            case 1: sy = z; sz = kD - 1 - y; break;            // rot90(m)[y][z] = m[z][n-1-y]
            case 2: sy = kW - 1 - y; sz = kD - 1 - z; break;
            default: sy = kW - 1 - z; sz = y; break;           // rot90^3(m)[y][z] = m[n-1-z][y]
        }
        if (flip_ud) sy = kW - 1 - sy;
        if (flip_lr) sz = kD - 1 - sz;
        out[i] = vol[((size_t)(x0 + x) * W + y0 + sy) * D + z0 + sz];
    }
}

__device__ __forceinline__ uint32_t enc_f(float f) {
    uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float dec_f(uint32_t u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }

__global__ void stitch_init_kernel(uint32_t* enc) { enc[0] = 0xffffffffu; enc[1] = 0u; }

__global__ void __launch_bounds__(NT) stitch_divide_kernel(const float* __restrict__ pred, const float* __restrict__ cnt, int H,
                                                           int W, int D, int x0, int y0, int z0, int oH, int oW, int oD,
                                                           float* __restrict__ out, uint32_t* enc) {
    size_t total = (size_t)oH * oW * oD;
    float mn = INFINITY, mx = -INFINITY;
    for (size_t i = (size_t)blockIdx.x * NT + threadIdx.x; i < total; i += (size_t)gridDim.x * NT) {
        int z = (int)(i % oD), y = (int)((i / oD) % oW), x = (int)(i / ((size_t)oD * oW));
        size_t s = (((size_t)(x + x0)) * W + y + y0) * D + z + z0;
        float v = __fdiv_rn(pred[s], cnt[s]);   // np.true_divide
        out[i] = v;
        mn = fminf(mn, v);
        mx = fmaxf(mx, v);
    }
    mn = warp_min(mn);
    mx = warp_max(mx);
    if ((threadIdx.x & 31) == 0) {
        atomicMin(enc, enc_f(mn));
        atomicMax(enc + 1, enc_f(mx));
    }
}

__global__ void stitch_decode_kernel(const uint32_t* enc, float* mm) {
    mm[0] = dec_f(enc[0]);
    mm[1] = dec_f(enc[1]);
}

__global__ void __launch_bounds__(NT) stitch_scale_kernel(float* __restrict__ out, size_t n, const float* __restrict__ mm) {
    float mn = mm[0], r = __fsub_rn(mm[1], mn);
    for (size_t i = (size_t)blockIdx.x * NT + threadIdx.x; i < n; i += (size_t)gridDim.x * NT)
        out[i] = __fmul_rn(255.f, __fdiv_rn(__fsub_rn(out[i], mn), r));   // 255 * (data - dmin) / (dmax - dmin)
}


// Order-exact stitching (custom_callback.py:142-192 restated in gather form).  The reference adds the windows into `pred` one after the
// other in enumeration order (i over rows, j over columns, k over depth; the clamped last window repeats).  Here every output voxel walks
// the windows that cover it in that same order and adds their values sequentially in fp32, so the sum -- and hence pred / count, the global
// min / max and the final uint8 cast -- is bit-identical to the numpy loop, whatever the batching, the rank count or the launch order.
// Window outputs live in `wins[slot]`; `slot_of[(i * nW + j) * nD + k]` maps an enumeration index to its slot (duplicates share one).
constexpr int ST_MAXW = 128;   // windows per axis (with the duplicated clamped window)
__global__ void __launch_bounds__(NT) stitch_gather_sum_kernel(const float* __restrict__ wins, const int* __restrict__ slot_of,
                                                               const int* __restrict__ starts, int nH, int nW, int nD, int kH, int kW,
                                                               int kD, int pH, int pW, int pD, int x0, int y0, int z0, int row0, int rows,
                                                               int oW, int oD, float* __restrict__ out, uint32_t* enc) {
    __shared__ int s_start[3][ST_MAXW];   // window starts per axis, enumeration order (non-decreasing)
    for (int i = threadIdx.x; i < nH + nW + nD; i += NT) {
        if (i < nH) s_start[0][i] = starts[i];
        else if (i < nH + nW) s_start[1][i - nH] = starts[i];
        else s_start[2][i - nH - nW] = starts[i];
    }
    __syncthreads();
    const size_t total = (size_t)rows * oW * oD;
    const size_t per = (size_t)kH * kW * kD;
    float mn = INFINITY, mx = -INFINITY;
    for (size_t i = (size_t)blockIdx.x * NT + threadIdx.x; i < total; i += (size_t)gridDim.x * NT) {
        const int z = (int)(i % oD) + z0, y = (int)((i / oD) % oW) + y0, x = (int)(i / ((size_t)oD * oW)) + row0 + x0;
        float acc = 0.f, cnt = 0.f;
        for (int a = 0; a < nH; a++) {
            const int lx = x - s_start[0][a];
            if (lx < pH || lx >= kH - pH) continue;
            for (int b = 0; b < nW; b++) {
                const int ly = y - s_start[1][b];
                if (ly < pW || ly >= kW - pW) continue;
                for (int c = 0; c < nD; c++) {
                    const int lz = z - s_start[2][c];
                    if (lz < pD || lz >= kD - pD) continue;
                    const int slot = slot_of[(a * nW + b) * nD + c];
                    acc = __fadd_rn(acc, wins[(size_t)slot * per + ((size_t)lx * kW + ly) * kD + lz]);
                    cnt += 1.f;
                }
            }
        }
        const float v = __fdiv_rn(acc, cnt);   // np.true_divide(pred, pix_tracker); 0 / 0 = NaN as in numpy
        out[i] = v;
        mn = fminf(mn, v);
        mx = fmaxf(mx, v);
    }
    mn = warp_min(mn);
    mx = warp_max(mx);
    if ((threadIdx.x & 31) == 0) {
        atomicMin(enc, enc_f(mn));
        atomicMax(enc + 1, enc_f(mx));
    }
}

// 255 * min_max_norm(pred) (custom_callback.py:202) and, for complete=False, .astype('uint8') (:204-205) in the same pass
__global__ void __launch_bounds__(NT) stitch_scale_u8_kernel(const float* __restrict__ in, uint8_t* __restrict__ out, size_t n,
                                                             const float* __restrict__ mm) {
    const float mn = mm[0], r = __fsub_rn(mm[1], mn);
    for (size_t i = (size_t)blockIdx.x * NT + threadIdx.x; i < n; i += (size_t)gridDim.x * NT)
        out[i] = (uint8_t)__fmul_rn(255.f, __fdiv_rn(__fsub_rn(in[i], mn), r));
}

}  // namespace

extern "C" {

// number of kernels this library has launched so far (monotonic; bench.py reports the difference)
unsigned long long vg_launch_count(void) { return g_vg_launches; }

int vg_upsample_concat(const void* lo, const void* skip, void* out, int N, int D, int H, int W, int C0, int C1, void* stream) {
    VG_REQUIRE(lo && skip && out && C0 % 8 == 0 && C1 % 8 == 0 && N > 0);
    VG_REQUIRE((long long)N * 4 * D * H < 0x7fffffffLL && (long long)2 * W * (C0 + C1) / 8 < 0x7fffffffLL);
    upsample_concat_kernel<<<vg_grid_for((long long)N * 4 * D * H, 1, 16), NT, 0, (cudaStream_t)stream>>>((const bf16*)lo, (const bf16*)skip, (bf16*)out,
                                                                                       N, D, H, W, C0, C1); VG_LAUNCHED(1);
    VG_CHECK_LAUNCH();
    return VG_OK;
}

int vg_upsample_concat_bwd(const void* dcat, void* dlo, void* dskip, int accumulate_skip, int N, int D, int H, int W, int C0,
                           int C1, void* stream) {
    VG_REQUIRE(dcat && dlo && dskip && C0 % 8 == 0 && C1 % 8 == 0);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t V2 = (size_t)N * 8 * D * H * W;
    VG_REQUIRE((long long)N * D * H < 0x7fffffffLL);
    upsample_concat_bwd_lo_kernel<<<vg_grid_for((long long)N * D * H, 1, 16), NT, 0, st>>>((const bf16*)dcat, (bf16*)dlo, N, D, H, W, C0, C1); VG_LAUNCHED(1);
    upsample_concat_bwd_skip_kernel<<<vg_grid_for(V2 * (C1 / 8), NT, 16), NT, 0, st>>>((const bf16*)dcat, (bf16*)dskip, V2, C0, C1,
                                                                                      accumulate_skip); VG_LAUNCHED(1);
    VG_CHECK_LAUNCH();
    return VG_OK;
}

int vg_pad_noise(const float* x, float* y, int N, int D, int H, int W, const float* noise, float noise_std,
                 unsigned long long seed, const unsigned long long* seed_dev, void* stream) {
    VG_REQUIRE(x && y && D >= 2 && H >= 2 && W >= 2);
    const long long rows = (long long)N * (D + 2) * (H + 2);
    VG_REQUIRE(rows < 0x7fffffffLL);
    pad_noise_kernel<<<vg_grid_for(rows, NT / 32, 16), NT, 0, (cudaStream_t)stream>>>(x, y, N, D, H, W, noise, noise_std, seed, seed_dev); VG_LAUNCHED(1);
    VG_CHECK_LAUNCH();
    return VG_OK;
}

int vg_pad_fold(const float* dy, float* dx, int N, int D, int H, int W, int accumulate, void* stream) {
    VG_REQUIRE(dy && dx);
    size_t total = (size_t)N * D * H * W;
    pad_fold_kernel<<<vg_grid_for(total, NT, 16), NT, 0, (cudaStream_t)stream>>>(dy, dx, N, D, H, W, accumulate); VG_LAUNCHED(1);
    VG_CHECK_LAUNCH();
    return VG_OK;
}

int vg_accumulate(void* a, const void* b, size_t n, int dtype, void* stream) {
    VG_REQUIRE(a && b);
    cudaStream_t st = (cudaStream_t)stream;
    size_t n8 = n / 8;
    if (dtype == VG_BF16) {
        accumulate_kernel<bf16><<<vg_grid_for(n8 + 1, NT, 16), NT, 0, st>>>((bf16*)a, (const bf16*)b, n8, n); VG_LAUNCHED(1);
    }
    else if (dtype == VG_F32) {
        accumulate_kernel<float><<<vg_grid_for(n8 + 1, NT, 16), NT, 0, st>>>((float*)a, (const float*)b, n8, n); VG_LAUNCHED(1);
    }
    else
        return VG_ERR_INVALID;
    VG_CHECK_LAUNCH();
    return VG_OK;
}

int vg_tanh_bwd(const float* dy, const float* y, float* out, size_t n, void* stream) {
    VG_REQUIRE(dy && y && out);
    tanh_bwd_kernel<<<vg_grid_for(n, NT * 2, 8), NT, 0, (cudaStream_t)stream>>>(dy, y, out, n); VG_LAUNCHED(1);
    VG_CHECK_LAUNCH();
    return VG_OK;
}

// norm_ws: nseg doubles (zeroed here)
static int clip_adam_impl(float* w, const float* g, float* m, float* v, const long long* seg_offsets, int nseg, long long total,
                          float lr_t, const float* lr_t_dev, float beta1, float beta2, float eps, float clipnorm, double* norm_ws,
                          void* stream) {
    VG_REQUIRE(w && g && m && v && seg_offsets && nseg > 0 && total > 0 && norm_ws);
    cudaStream_t st = (cudaStream_t)stream;
    if (cudaMemsetAsync(norm_ws, 0, (size_t)nseg * sizeof(double), st) != cudaSuccess) return VG_ERR_CUDA;
    seg_sqnorm_kernel<<<vg_grid_for(total / 128 + 1, NT, 8), NT, 0, st>>>(g, seg_offsets, nseg, norm_ws, total); VG_LAUNCHED(1);
    clip_adam_kernel<<<vg_grid_for((total + 3) / 4, NT, 16), NT, 0, st>>>(w, g, m, v, seg_offsets, nseg, norm_ws, lr_t, beta1, beta2, eps,
                                                               clipnorm, total, lr_t_dev); VG_LAUNCHED(1);
    VG_CHECK_LAUNCH();
    return VG_OK;
}

int vg_clip_adam_step(float* w, const float* g, float* m, float* v, const long long* seg_offsets, int nseg, long long total,
                      float lr_t, float beta1, float beta2, float eps, float clipnorm, double* norm_ws, void* stream) {
    return clip_adam_impl(w, g, m, v, seg_offsets, nseg, total, lr_t, nullptr, beta1, beta2, eps, clipnorm, norm_ws, stream);
}

int vg_clip_adam_step_dev(float* w, const float* g, float* m, float* v, const long long* seg_offsets, int nseg, long long total,
                          const float* lr_t_dev, float beta1, float beta2, float eps, float clipnorm, double* norm_ws, void* stream) {
    VG_REQUIRE(lr_t_dev);
    return clip_adam_impl(w, g, m, v, seg_offsets, nseg, total, 0.f, lr_t_dev, beta1, beta2, eps, clipnorm, norm_ws, stream);
}

int vg_dropout_mask(float* out, int n, float rate, unsigned long long seed, const unsigned long long* seed_dev, void* stream) {
    VG_REQUIRE(out && n > 0 && rate >= 0.f && rate < 1.f);
    dropout_mask_kernel<<<vg_cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(out, n, rate, seed, seed_dev); VG_LAUNCHED(1);
    VG_CHECK_LAUNCH();
    return VG_OK;
}

int vg_stitch_gather(const float* vol, int H, int W, int D, float* win, const int* starts, int B, int kH, int kW, int kD, void* stream) {
    VG_REQUIRE(vol && win && starts && B > 0);
    stitch_gather_kernel<<<vg_grid_for((size_t)B * kH * kW * kD, NT, 16), NT, 0, (cudaStream_t)stream>>>(vol, H, W, D, win, starts, B, kH,
                                                                                                   kW, kD); VG_LAUNCHED(1);
    VG_CHECK_LAUNCH();
    return VG_OK;
}

int vg_crop_augment(const float* vol, int H, int W, int D, float* out, int kH, int kW, int kD, int x0, int y0, int z0, int flip_lr,
                    int flip_ud, int rot_k, void* stream) {
    VG_REQUIRE(vol && out && kH > 0 && kW > 0 && kD > 0 && x0 >= 0 && y0 >= 0 && z0 >= 0 && x0 + kH <= H && y0 + kW <= W && z0 + kD <= D);
    VG_REQUIRE(!(rot_k & 1) || kW == kD);
    crop_augment_kernel<<<vg_grid_for((size_t)kH * kW * kD, NT, 16), NT, 0, (cudaStream_t)stream>>>(vol, H, W, D, out, kH, kW, kD, x0, y0, z0,
                                                                                              flip_lr, flip_ud, rot_k); VG_LAUNCHED(1);
    VG_CHECK_LAUNCH();
    return VG_OK;
}

int vg_stitch_gather_sym(const float* vol, int row0, int H0, int W0, int D0, int xs, int ys, int zs, float* win, const int* starts, int B,
                         int kH, int kW, int kD, void* stream) {
    VG_REQUIRE(vol && win && starts && B > 0 && row0 >= 0 && xs >= 0 && ys >= 0 && zs >= 0 && xs < H0 && ys < W0 && zs < D0);
    stitch_gather_sym_kernel<<<vg_grid_for((size_t)B * kH * kW * kD, NT, 16), NT, 0, (cudaStream_t)stream>>>(vol, row0, H0, W0, D0, xs, ys, zs, win,
                                                                                                       starts, B, kH, kW, kD); VG_LAUNCHED(1);
    VG_CHECK_LAUNCH();
    return VG_OK;
}

int vg_stitch_accumulate(float* pred, float* cnt, int H, int W, int D, const float* win, const int* starts, int B, int kH,
                         int kW, int kD, int pH, int pW, int pD, void* stream) {
    VG_REQUIRE(pred && cnt && win && starts && B > 0);
    size_t total = (size_t)B * (kH - 2 * pH) * (kW - 2 * pW) * (kD - 2 * pD);
    stitch_accumulate_kernel<<<vg_grid_for(total, NT, 16), NT, 0, (cudaStream_t)stream>>>(pred, cnt, H, W, D, win, starts, B, kH, kW,
                                                                                         kD, pH, pW, pD); VG_LAUNCHED(1);
    VG_CHECK_LAUNCH();
    return VG_OK;
}

int vg_stitch_finalize(const float* pred, const float* cnt, int H, int W, int D, int x0, int y0, int z0, int oH, int oW, int oD,
                       float* out, float* mm, void* enc_ws, void* stream) {
    VG_REQUIRE(pred && cnt && out && mm && enc_ws);
    cudaStream_t st = (cudaStream_t)stream;
    stitch_init_kernel<<<1, 1, 0, st>>>((uint32_t*)enc_ws); VG_LAUNCHED(1);
    stitch_divide_kernel<<<vg_grid_for((size_t)oH * oW * oD, NT, 16), NT, 0, st>>>(pred, cnt, H, W, D, x0, y0, z0, oH, oW, oD, out,
                                                                                  (uint32_t*)enc_ws); VG_LAUNCHED(1);
    stitch_decode_kernel<<<1, 1, 0, st>>>((const uint32_t*)enc_ws, mm); VG_LAUNCHED(1);
    VG_CHECK_LAUNCH();
    return VG_OK;
}

int vg_stitch_scale(float* out, size_t n, const float* mm, void* stream) {
    VG_REQUIRE(out && mm);
    stitch_scale_kernel<<<vg_grid_for(n, NT, 16), NT, 0, (cudaStream_t)stream>>>(out, n, mm); VG_LAUNCHED(1);
    VG_CHECK_LAUNCH();
    return VG_OK;
}

int vg_stitch_gather_sum(const float* wins, const int* slot_of, const int* starts_dev, int nH, int nW, int nD, int kH, int kW, int kD, int pH,
                         int pW, int pD, int x0, int y0, int z0, int row0, int rows, int oW, int oD, float* out, void* enc_ws, int init_enc,
                         void* stream) {
    VG_REQUIRE(wins && slot_of && starts_dev && out && enc_ws);
    VG_REQUIRE(nH > 0 && nW > 0 && nD > 0 && nH <= ST_MAXW && nW <= ST_MAXW && nD <= ST_MAXW && rows > 0 && oW > 0 && oD > 0);
    cudaStream_t st = (cudaStream_t)stream;
    if (init_enc) { stitch_init_kernel<<<1, 1, 0, st>>>((uint32_t*)enc_ws); VG_LAUNCHED(1); }
    stitch_gather_sum_kernel<<<vg_grid_for((size_t)rows * oW * oD, NT, 16), NT, 0, st>>>(wins, slot_of, starts_dev, nH, nW, nD, kH, kW, kD, pH, pW,
                                                                                        pD, x0, y0, z0, row0, rows, oW, oD, out,
                                                                                        (uint32_t*)enc_ws); VG_LAUNCHED(1);
    VG_CHECK_LAUNCH();
    return VG_OK;
}

int vg_stitch_minmax_decode(const void* enc_ws, float* mm, void* stream) {
    VG_REQUIRE(enc_ws && mm);
    stitch_decode_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((const uint32_t*)enc_ws, mm); VG_LAUNCHED(1);
    VG_CHECK_LAUNCH();
    return VG_OK;
}

int vg_stitch_scale_u8(const float* in, unsigned char* out, size_t n, const float* mm, void* stream) {
    VG_REQUIRE(in && out && mm);
    stitch_scale_u8_kernel<<<vg_grid_for(n, NT, 16), NT, 0, (cudaStream_t)stream>>>(in, out, n, mm); VG_LAUNCHED(1);
    VG_CHECK_LAUNCH();
    return VG_OK;
}

}  // extern "C"
