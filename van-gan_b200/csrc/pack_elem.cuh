// Weight-operand packing, element-wise: the three layouts the convolution kernels read (mma.sync forward / dgrad packs of conv_mma.cu,
// tcgen05 packs of conv_tc.cu) expressed as "value of packed element i" so that ONE kernel can refresh every operand copy of a
// network after the optimizer step (vg_pack_run) instead of ~100 tiny launches per network (vg_conv3d_pack_weights per layer).
#pragma once
#include "common.cuh"

// One packing job = one contiguous packed buffer.  kind 0: mma.sync forward pack; 1: mma.sync dgrad pack of ONE stride-parity class;
// 2: tcgen05 pack (forward, dgrad class, stride-2 forward classes, fused stride-2 dgrad classes).
struct vg_pack_job {
    const float* w;      // fp32 Keras kernel (K,K,K,Cin,Cout)
    bf16* out;           // packed destination
    long long total;     // elements of this job
    int kind;
    int K, stride, Cin, Cout;
    int dgrad, ad, ah, aw, td, th, tw;   // kind 1 / 2: class offsets and taps per axis
    int ncta, nblk, dm, dsplit;          // kind 2
    int Np, T;                           // kind 0 / 1: padded row count, taps
};

// pack_src_*: index into the fp32 Keras kernel of packed element i, or -1 for a zero (padding) element
__device__ __forceinline__ long long pack_src_mma_fwd(const vg_pack_job& j, size_t i) {
    const int ci = (int)(i % j.Cin);
    const int co = (int)((i / j.Cin) % j.Np);
    const int t = (int)(i / ((size_t)j.Cin * j.Np));
    return co < j.Cout ? (long long)(((size_t)t * j.Cin + ci) * j.Cout + co) : -1;
}

__device__ __forceinline__ long long pack_src_mma_dgrad(const vg_pack_job& j, size_t i) {
    const int co = (int)(i % j.Cout);
    const int ci = (int)((i / j.Cout) % j.Np);
    const int tt = (int)(i / ((size_t)j.Cout * j.Np));
    const int w_ = tt % j.tw, h_ = (tt / j.tw) % j.th, d_ = tt / (j.tw * j.th);
    const int kd = j.ad + j.stride * d_, kh = j.ah + j.stride * h_, kw = j.aw + j.stride * w_;
    const int t = (kd * j.K + kh) * j.K + kw;
    return ci < j.Cin ? (long long)(((size_t)t * j.Cin + ci) * j.Cout + co) : -1;
}

// tensor-core layout: out[nb][c][t][kh][n][j] = src(t, k = c*16+kh*8+j, col = nb*NCTA+n)
// fwd:   src(t,k,col) = w[t][k][col]            (K = Cin, cols = Cout)
// dgrad: src(t',k,col) = w[tap(t')][col][k]      (K = Cout, cols = Cin), taps restricted to one stride-parity class
// fwd stride 2 (dgrad == 2): chunk c = (parity class a,b,c ; 16-channel chunk), taps t' in 2x2x2, src = w[2t'+a][k][col] (0 if >= K)
// dgrad == 3: fused stride-2 parity classes -- columns are (class, ci), every class padded to 2x2x2 taps
__device__ __forceinline__ long long pack_src_tc(const vg_pack_job& q, size_t i) {
    const int K = q.K, stride = q.stride, Cin = q.Cin, Cout = q.Cout, dgrad = q.dgrad, td = q.td, th = q.th, tw = q.tw, ncta = q.ncta;
    const int T = q.dsplit ? th * tw : td * th * tw;
    const int Kt = (dgrad == 1 || dgrad == 3) ? Cout : Cin, ncols = dgrad == 1 ? Cin : (dgrad == 3 ? 8 * Cin : Cout);
    const int cpc = Kt / 16;
    const int nchunks = (dgrad == 2 ? 8 : (q.dsplit ? td : 1)) * cpc;
    size_t r = i;
    int j = (int)(r % 8); r /= 8;
    int n = (int)(r % ncta); r /= ncta;
    int w_, h_, d_, kh;
    if (q.dm) {
        // d-march layout [q = (th, tw)][K half][pos][n][8]: position pos along N holds the tap that maps source slice s to output
        // tile m_lo + pos, i.e. loop index td = TD-1-pos for a forward gather (slice = m + td) and td = pos for dgrad
        int pos = (int)(r % td); r /= td;
        kh = (int)(r % 2); r /= 2;
        int qq = (int)(r % (th * tw)); r /= (th * tw);
        w_ = qq % tw; h_ = qq / tw;
        d_ = dgrad == 1 ? pos : td - 1 - pos;
    } else {
        kh = (int)(r % 2); r /= 2;
        int t = (int)(r % T); r /= T;
        w_ = t % tw; h_ = (t / tw) % th; d_ = t / (tw * th);   // d-split: t < th*tw, d_ = 0 here, set from the chunk below
    }
    int c = (int)(r % nchunks);
    int nb = (int)(r / nchunks);
    int col = nb * ncta + n;
    int k, kd, kh2, kw;
    if (dgrad == 2) {
        const int cls = c / cpc, cc = c - cls * cpc;
        k = cc * 16 + kh * 8 + j;
        kd = 2 * d_ + ((cls >> 2) & 1); kh2 = 2 * h_ + ((cls >> 1) & 1); kw = 2 * w_ + (cls & 1);
    } else {
        int cc = c;
        if (q.dsplit) { d_ = c / cpc; cc = c - d_ * cpc; }
        k = cc * 16 + kh * 8 + j;
        kd = dgrad ? q.ad + stride * d_ : d_; kh2 = dgrad ? q.ah + stride * h_ : h_; kw = dgrad ? q.aw + stride * w_ : w_;
    }
    int ci = col;
    if (dgrad == 3) {
        const int cls = col / Cin;
        ci = col - cls * Cin;
        kd = ((cls >> 2) & 1) + 2 * d_; kh2 = ((cls >> 1) & 1) + 2 * h_; kw = (cls & 1) + 2 * w_;
    }
    int tap = (kd * K + kh2) * K + kw;
    if (col < ncols && kd < K && kh2 < K && kw < K)
        return (dgrad == 1 || dgrad == 3) ? (long long)(((size_t)tap * Cin + ci) * Cout + k) : (long long)(((size_t)tap * Cin + k) * Cout + col);
    return -1;
}

__device__ __forceinline__ long long pack_src(const vg_pack_job& j, size_t i) {
    return j.kind == 0 ? pack_src_mma_fwd(j, i) : (j.kind == 1 ? pack_src_mma_dgrad(j, i) : pack_src_tc(j, i));
}
__device__ __forceinline__ float pack_elem(const vg_pack_job& j, size_t i) {
    const long long s = pack_src(j, i);
    return s < 0 ? 0.f : j.w[s];
}

// host side of the tcgen05 pack: fills the job for one (layer, class); returns false when the shape has no tcgen05 layout
bool vg_tc_pack_job(const float* w, bf16* out, int K, int stride, int Cin, int Cout, int dgrad, int ad, int ah, int aw, int td, int th, int tw,
                    vg_pack_job* job);
