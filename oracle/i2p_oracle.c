/*
 * i2p_oracle.c -- CPU restatement of the I2PNet hot-path index/gather operators.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in i2pnet_b200/ may import, link or call this
 * file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, and only as the checker / the CPU baseline.
 *
 * Each function restates one reference kernel line by line (citations are relative
 * to /root/reference).  The reference has no CPU implementation and no tests or
 * golden vectors of its own for this path, so this oracle is
 * pinned against outputs of the reference's own CUDA kernels compiled unmodified for
 * sm_100 (oracle/_ref, see oracle/build_ref.py) and run on a B200: those outputs are
 * committed under tests/golden/ref_gpu_*.npz together with the generating script
 * (tests/golden/make_ref_gpu_golden.py).
 *
 * Arithmetic: nvcc contracts every squared-distance expression of the reference to
 *     d = fmaf(dz, dz, fmaf(dx, dx, dy * dy))
 * (SASS of the -O2 sm_100 build: FMUL dy*dy; FFMA dx*dx+t; FFMA dz*dz+t; SURVEY.md
 * section 2.2).  This file spells that with fmaf() and must be compiled with
 * -ffp-contract=off so that gcc adds no contraction of its own.
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline float sqd(float dx, float dy, float dz) {
    return fmaf(dz, dz, fmaf(dx, dx, dy * dy));
}

/* ------------------------------------------------------------------------- */
/* K1 furthest point sampling: pointnet2/src/sampling_gpu.cu:86-209           */
/* block size rule: pointnet2/src/cuda_utils.h:10-14                          */
/* ------------------------------------------------------------------------- */
int orc_fps_block_size(int n) {
    const int pow_2 = (int)(log((double)n) / log(2.0));
    int t = 1 << pow_2;
    if (t > 1024) t = 1024;
    if (t < 1) t = 1;
    return t;
}

/* dataset (B,N,3), temp (B,N) pre-filled by caller (1e10), idxs (B,M) */
void orc_fps(int b, int n, int m, const float *dataset, float *temp, int32_t *idxs) {
    if (m <= 0) return;                                   /* sampling_gpu.cu:101 */
    const int bs = orc_fps_block_size(n);                 /* sampling_gpu.cu:219 */
    float *dists = (float *)malloc(sizeof(float) * bs);
    int *dists_i = (int *)malloc(sizeof(int) * bs);
    for (int bi = 0; bi < b; ++bi) {
        const float *d = dataset + (size_t)bi * n * 3;
        float *tp = temp + (size_t)bi * n;
        int32_t *out = idxs + (size_t)bi * m;
        int old = 0;                                      /* :113 */
        out[0] = old;                                     /* :115 */
        for (int j = 1; j < m; ++j) {                     /* :118 */
            const float x1 = d[old * 3 + 0], y1 = d[old * 3 + 1], z1 = d[old * 3 + 2];
            for (int tid = 0; tid < bs; ++tid) {          /* one "thread" */
                int besti = 0;
                float best = -1.f;                        /* :119-120 */
                for (int k = tid; k < n; k += bs) {       /* :124 */
                    const float x2 = d[k * 3 + 0], y2 = d[k * 3 + 1], z2 = d[k * 3 + 2];
                    const float dd = sqd(x2 - x1, y2 - y1, z2 - z1);   /* :133 */
                    const float d2 = fminf(dd, tp[k]);    /* :134 */
                    tp[k] = d2;
                    besti = d2 > best ? k : besti;        /* :136 */
                    best = d2 > best ? d2 : best;         /* :137 */
                }
                dists[tid] = best;
                dists_i[tid] = besti;
            }
            /* tree reduction :143-203, __update :86-91 (lower slot wins ties) */
            for (int s = bs / 2; s >= 1; s >>= 1) {
                for (int tid = 0; tid < s; ++tid) {
                    const float v1 = dists[tid], v2 = dists[tid + s];
                    const int i1 = dists_i[tid], i2 = dists_i[tid + s];
                    dists[tid] = fmaxf(v1, v2);
                    dists_i[tid] = v2 > v1 ? i2 : i1;
                }
            }
            old = dists_i[0];                             /* :205 */
            out[j] = old;
        }
    }
    free(dists);
    free(dists_i);
}

/* ------------------------------------------------------------------------- */
/* K2/K3 gather points (+grad): sampling_gpu.cu:8-63                           */
/* ------------------------------------------------------------------------- */
void orc_gather_points(int b, int c, int n, int m, const float *points, const int32_t *idx, float *out) {
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci)
            for (int p = 0; p < m; ++p)
                out[((size_t)bi * c + ci) * m + p] = points[((size_t)bi * c + ci) * n + idx[(size_t)bi * m + p]];
}

/* grad_points must be zero-filled by the caller (pointnet2_utils.py:98) */
void orc_gather_points_grad(int b, int c, int n, int m, const float *grad_out, const int32_t *idx, float *grad_points) {
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci)
            for (int p = 0; p < m; ++p)
                grad_points[((size_t)bi * c + ci) * n + idx[(size_t)bi * m + p]] += grad_out[((size_t)bi * c + ci) * m + p];
}

/* ------------------------------------------------------------------------- */
/* K4 ball query: ball_query_gpu.cu:9-45 (idx pre-zeroed by caller)            */
/* ------------------------------------------------------------------------- */
void orc_ball_query(int b, int n, int m, float radius, int nsample, const float *new_xyz, const float *xyz, int32_t *idx) {
    const float radius2 = radius * radius;                /* :24 */
    for (int bi = 0; bi < b; ++bi)
        for (int p = 0; p < m; ++p) {
            const float *q = new_xyz + ((size_t)bi * m + p) * 3;
            const float *x = xyz + (size_t)bi * n * 3;
            int32_t *o = idx + ((size_t)bi * m + p) * nsample;
            int cnt = 0;
            for (int k = 0; k < n; ++k) {
                const float d2 = sqd(q[0] - x[k * 3 + 0], q[1] - x[k * 3 + 1], q[2] - x[k * 3 + 2]);  /* :33 */
                if (d2 < radius2) {                       /* :34 strict */
                    if (cnt == 0)
                        for (int l = 0; l < nsample; ++l) o[l] = k;   /* :35-39 */
                    o[cnt] = k;
                    ++cnt;
                    if (cnt >= nsample) break;            /* :42 */
                }
            }
        }
}

/* ------------------------------------------------------------------------- */
/* K5/K6 group points (+grad): group_points_gpu.cu:8-66                        */
/* ------------------------------------------------------------------------- */
void orc_group_points(int b, int c, int n, int npoints, int nsample, const float *points, const int32_t *idx, float *out) {
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci)
            for (int p = 0; p < npoints; ++p)
                for (int s = 0; s < nsample; ++s)
                    out[(((size_t)bi * c + ci) * npoints + p) * nsample + s] =
                        points[((size_t)bi * c + ci) * n + idx[((size_t)bi * npoints + p) * nsample + s]];
}

void orc_group_points_grad(int b, int c, int n, int npoints, int nsample, const float *grad_out, const int32_t *idx, float *grad_points) {
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci)
            for (int p = 0; p < npoints; ++p)
                for (int s = 0; s < nsample; ++s)
                    grad_points[((size_t)bi * c + ci) * n + idx[((size_t)bi * npoints + p) * nsample + s]] +=
                        grad_out[(((size_t)bi * c + ci) * npoints + p) * nsample + s];
}

/* ------------------------------------------------------------------------- */
/* K7 three_nn: interpolate_gpu.cu:9-52 (best* are double, d is float)         */
/* unknown (B,N,3), known (B,M,3) -> dist2 (B,N,3) squared, idx (B,N,3)         */
/* ------------------------------------------------------------------------- */
void orc_three_nn(int b, int n, int m, const float *unknown, const float *known, float *dist2, int32_t *idx) {
    for (int bi = 0; bi < b; ++bi)
        for (int p = 0; p < n; ++p) {
            const float *u = unknown + ((size_t)bi * n + p) * 3;
            const float *kn = known + (size_t)bi * m * 3;
            double best1 = 1e40, best2 = 1e40, best3 = 1e40;   /* :30 */
            int besti1 = 0, besti2 = 0, besti3 = 0;
            for (int k = 0; k < m; ++k) {
                const float d = sqd(u[0] - kn[k * 3 + 0], u[1] - kn[k * 3 + 1], u[2] - kn[k * 3 + 2]);  /* :36 */
                if (d < best1) {
                    best3 = best2; besti3 = besti2;
                    best2 = best1; besti2 = besti1;
                    best1 = d; besti1 = k;
                } else if (d < best2) {
                    best3 = best2; besti3 = besti2;
                    best2 = d; besti2 = k;
                } else if (d < best3) {
                    best3 = d; besti3 = k;
                }
            }
            float *od = dist2 + ((size_t)bi * n + p) * 3;
            int32_t *oi = idx + ((size_t)bi * n + p) * 3;
            od[0] = (float)best1; od[1] = (float)best2; od[2] = (float)best3;   /* :50 double->float */
            oi[0] = besti1; oi[1] = besti2; oi[2] = besti3;
        }
}

/* ------------------------------------------------------------------------- */
/* K8/K9 three_interpolate (+grad): interpolate_gpu.cu:77-142                  */
/* nvcc contracts w0*p0 + w1*p1 + w2*p2 like the distance sites: FMUL w1*p1;    */
/* FFMA w0*p0+t; FFMA w2*p2+t (SASS of oracle/_ref/pointnet2_cuda.so)            */
/* ------------------------------------------------------------------------- */
void orc_three_interpolate(int b, int c, int m, int n, const float *points, const int32_t *idx, const float *weight, float *out) {
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci) {
            const float *pt = points + ((size_t)bi * c + ci) * m;
            for (int p = 0; p < n; ++p) {
                const float *w = weight + ((size_t)bi * n + p) * 3;
                const int32_t *id = idx + ((size_t)bi * n + p) * 3;
                out[((size_t)bi * c + ci) * n + p] = fmaf(w[2], pt[id[2]], fmaf(w[0], pt[id[0]], w[1] * pt[id[1]]));
            }
        }
}

void orc_three_interpolate_grad(int b, int c, int n, int m, const float *grad_out, const int32_t *idx, const float *weight, float *grad_points) {
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci) {
            float *gp = grad_points + ((size_t)bi * c + ci) * m;
            for (int p = 0; p < n; ++p) {
                const float *w = weight + ((size_t)bi * n + p) * 3;
                const int32_t *id = idx + ((size_t)bi * n + p) * 3;
                const float g = grad_out[((size_t)bi * c + ci) * n + p];
                gp[id[0]] += g * w[0];
                gp[id[1]] += g * w[1];
                gp[id[2]] += g * w[2];
            }
        }
}

/* ------------------------------------------------------------------------- */
/* K10 fused_conv_select_k: src/projectPN/fused_conv_select/fused_conv_go.cu:11-240
 * Outputs are pre-zeroed by the caller and only partially written (utils.py:86-94).
 * valid_idx / valid_in_dis_idx are never written by the reference (:150,:167 are
 * commented out), so they are not parameters here.                            */
/* ------------------------------------------------------------------------- */
#define ORC_MAX_KERNEL 150
int orc_fused_conv_select_k(int batch_size, int H, int W, int npoints, int kH, int kW, int K, int flag,
                            float distance, int stride_h, int stride_w, const float *xyz1, const float *xyz2,
                            const int32_t *idx_n2, const int32_t *random_hw, int64_t *sel_b, int64_t *sel_h,
                            int64_t *sel_w, float *sel_mask, int small_h, int small_w) {
    const int kernel_total = kH * kW;                     /* :23 */
    if (kernel_total > ORC_MAX_KERNEL || K > ORC_MAX_KERNEL) return -1;   /* arrays are [150] :52-53 */
    const float dist_square = distance * distance;        /* :26 */
    const int half_H = kH / 2, half_W = kW / 2;           /* :28-29 */
    for (int bi = 0; bi < batch_size; ++bi) {
        const float *x1 = xyz1 + (size_t)bi * H * W * 3;
        const float *x2 = xyz2 + (size_t)bi * small_h * small_w * 3;
        const int32_t *ctr = idx_n2 + (size_t)bi * npoints * 2;
        int64_t *ob = sel_b + (size_t)bi * npoints * K;
        int64_t *oh = sel_h + (size_t)bi * npoints * K;
        int64_t *ow = sel_w + (size_t)bi * npoints * K;
        float *om = sel_mask + (size_t)bi * npoints * K;
        for (int cn = 0; cn < npoints; ++cn) {            /* :49 */
            int idx_w[ORC_MAX_KERNEL], idx_h[ORC_MAX_KERNEL];
            float Dist[ORC_MAX_KERNEL];
            for (int ii = 0; ii < ORC_MAX_KERNEL; ++ii) { idx_w[ii] = 0; idx_h[ii] = 0; Dist[ii] = 1e10f; }
            int m_idx = 0, num_select = 0;
            const int sH = ctr[cn * 2 + 0], sW = ctr[cn * 2 + 1];          /* :65-66 */
            const float xc = x1[(sH * W + sW) * 3 + 0], yc = x1[(sH * W + sW) * 3 + 1], zc = x1[(sH * W + sW) * 3 + 2];
            const float Dist_c = fmaxf(sqd(xc, yc, zc), 1e-10f);           /* :72 */
            if (Dist_c <= 1e-10f) continue;                                /* :74-78 */
            for (int cur = 0; cur < kernel_total; ++cur) {                 /* :82 */
                const int khw = random_hw[cur];
                int kh = sH / stride_h + khw / kW - half_H;                /* :89 */
                int kw = sW / stride_w + khw % kW - half_W;                /* :91 */
                if (flag & 0x2) {                                          /* :96 */
                    if (kh < 0 || kh >= small_h) { ++m_idx; continue; }
                    if (kw < 0) kw = small_w + kw;                         /* :106 */
                    if (kw >= small_w) kw = kw - small_w;                  /* :110 */
                } else {
                    if (kh < 0 || kh >= small_h || kw < 0 || kw >= small_w) { ++m_idx; continue; }
                }
                const float xq = x2[(kh * small_w + kw) * 3 + 0], yq = x2[(kh * small_w + kw) * 3 + 1],
                            zq = x2[(kh * small_w + kw) * 3 + 2];
                const float Dist_q_0 = sqd(xq, yq, zq);                    /* :140 */
                if (Dist_q_0 <= 1e-10f) { ++m_idx; continue; }
                const float Dist_q = fmaxf(sqd(xc - xq, yc - yq, zc - zq), 1e-10f);   /* :153 */
                if (Dist_q > dist_square) { ++m_idx; continue; }           /* :156 */
                Dist[m_idx] = Dist_q; idx_h[m_idx] = kh; idx_w[m_idx] = kw;
                ++m_idx; ++num_select;
                if (num_select >= kernel_total) break;
            }
            for (int s = 0; s < K; ++s) {                                  /* :183 selection sort */
                int mi = s;
                for (int t = s + 1; t < kernel_total; ++t)
                    if (Dist[t] < Dist[mi]) mi = t;
                if (mi != s) {
                    float td = Dist[mi]; int tw = idx_w[mi], th = idx_h[mi];
                    Dist[mi] = Dist[s]; idx_w[mi] = idx_w[s]; idx_h[mi] = idx_h[s];
                    Dist[s] = td; idx_w[s] = tw; idx_h[s] = th;
                }
                if ((flag & 0x1) && s == 0)                                /* :211 COPY */
                    for (int k = 0; k < K; ++k) {
                        ob[cn * K + k] = bi; oh[cn * K + k] = idx_h[s]; ow[cn * K + k] = idx_w[s]; om[cn * K + k] = 1.0f;
                    }
                if (Dist[s] < 1e10f) {                                     /* :225 */
                    ob[cn * K + s] = bi; oh[cn * K + s] = idx_h[s]; ow[cn * K + s] = idx_w[s]; om[cn * K + s] = 1.0f;
                }
            }
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------- */
/* kNN: src/projectPN/utils.py:344-380 (square_distance + topk(sorted=False)).
 * dist = -2*(q.x) + |q|^2 + |x|^2 in f32 (torch: matmul, then two in-place adds);
 * the k smallest per query.  The order of the k outputs is unspecified in the
 * reference (sorted=False), so this oracle returns them sorted by (dist, index)
 * and parity is SET equality per query.  dist_out may be NULL.                 */
/* ------------------------------------------------------------------------- */
void orc_knn(int b, int n, int s, int k, const float *xyz, const float *new_xyz, int64_t *idx_out, float *dist_out) {
    float *dd = (float *)malloc(sizeof(float) * n);
    int *ord = (int *)malloc(sizeof(int) * n);
    for (int bi = 0; bi < b; ++bi)
        for (int q = 0; q < s; ++q) {
            const float *qq = new_xyz + ((size_t)bi * s + q) * 3;
            const float qn = (qq[0] * qq[0] + qq[1] * qq[1]) + qq[2] * qq[2];
            for (int j = 0; j < n; ++j) {
                const float *x = xyz + ((size_t)bi * n + j) * 3;
                const float dot = fmaf(qq[2], x[2], fmaf(qq[1], x[1], qq[0] * x[0]));
                const float xn = (x[0] * x[0] + x[1] * x[1]) + x[2] * x[2];
                dd[j] = (-2.0f * dot + qn) + xn;
                ord[j] = j;
            }
            /* partial selection of the k smallest by (dist, index) */
            for (int a = 0; a < k && a < n; ++a) {
                int mi = a;
                for (int t = a + 1; t < n; ++t)
                    if (dd[ord[t]] < dd[ord[mi]] || (dd[ord[t]] == dd[ord[mi]] && ord[t] < ord[mi])) mi = t;
                int tmp = ord[a]; ord[a] = ord[mi]; ord[mi] = tmp;
                idx_out[((size_t)bi * s + q) * k + a] = ord[a];
                if (dist_out) dist_out[((size_t)bi * s + q) * k + a] = dd[ord[a]];
            }
        }
    free(dd);
    free(ord);
}

/* ------------------------------------------------------------------------- */
/* gather_torch: src/projectPN/utils.py:36-60.  feature (B,HW,C) channels-last,
 * flat index idx = h*W + w (B,M) -> out (B,M,C).  Backward = scatter-add.       */
/* ------------------------------------------------------------------------- */
void orc_gather_rows(int b, int hw, int c, int m, const float *feat, const int32_t *idx, float *out) {
    for (int bi = 0; bi < b; ++bi)
        for (int p = 0; p < m; ++p)
            memcpy(out + ((size_t)bi * m + p) * c, feat + ((size_t)bi * hw + idx[(size_t)bi * m + p]) * c, sizeof(float) * c);
}

void orc_gather_rows_grad(int b, int hw, int c, int m, const float *grad_out, const int32_t *idx, float *grad_feat) {
    for (int bi = 0; bi < b; ++bi)
        for (int p = 0; p < m; ++p) {
            float *g = grad_feat + ((size_t)bi * hw + idx[(size_t)bi * m + p]) * c;
            const float *go = grad_out + ((size_t)bi * m + p) * c;
            for (int ci = 0; ci < c; ++ci) g[ci] += go[ci];
        }
}

/* ------------------------------------------------------------------------- */
/* project_seq with rank=False: src/projectPN/utils.py:111-187.
 * f32 arithmetic as torch evaluates it on CUDA: scalars are computed in double by the Python
 * code (:125-139) and cast to f32 where they meet f32 tensors; a division by a Python scalar
 * is a multiplication by its f32 reciprocal (ATen BinaryDivTrueKernel.cu, cpu-scalar fast
 * path).  Duplicate cells: the highest point index wins (index_put_ on CPU is sequential,
 * last writer wins; on CUDA the winner is unordered -- parity inputs avoid duplicates).
 * xyz (B,N,3); feats[j] (B,N,fdim[j]) -> xyz_proj (B,H,W,3), feat_proj[j] (B,H,W,fdim[j]).    */
/* ------------------------------------------------------------------------- */
static long long clampll(long long v, long long lo, long long hi) { return v < lo ? lo : (v > hi ? hi : v); }

void orc_project_seq(int b, int n, int H, int W, float fup_deg, float fdown_deg, const float *xyz, int nfeat,
                     const float *const *feats, const int *fdims, float *xyz_proj, float *const *feat_projs) {
    const double deg2rad = M_PI / 180.0;
    const double az = 360.0 / W * deg2rad;
    const double down = (double)fdown_deg * deg2rad, up = (double)fup_deg * deg2rad;
    const double vres = (up - down) / (H - 1);
    const double voff = -down / vres;
    const float pi = (float)M_PI, inv_az = 1.0f / (float)az, inv_vres = 1.0f / (float)vres, voff_f = (float)voff;
    memset(xyz_proj, 0, sizeof(float) * (size_t)b * H * W * 3);
    for (int j = 0; j < nfeat; ++j) memset(feat_projs[j], 0, sizeof(float) * (size_t)b * H * W * fdims[j]);
    for (int bi = 0; bi < b; ++bi)
        for (int i = 0; i < n; ++i) {                     /* ascending i: the last writer wins */
            const float *p = xyz + ((size_t)bi * n + i) * 3;
            const float r = sqrtf(fmaf(p[2], p[2], fmaf(p[1], p[1], p[0] * p[0])));       /* :144 */
            const long long col = (long long)((pi - atan2f(p[1], p[0])) * inv_az);        /* :147 */
            const float beta = asinf(p[2] / r);                                           /* :150 */
            const long long row = (long long)H - (long long)(beta * inv_vres + voff_f);   /* :152 */
            const size_t cell = ((size_t)bi * H + (size_t)clampll(row, 0, H - 1)) * W + (size_t)clampll(col, 0, W - 1);
            memcpy(xyz_proj + cell * 3, p, sizeof(float) * 3);
            for (int j = 0; j < nfeat; ++j)
                memcpy(feat_projs[j] + cell * fdims[j], feats[j] + ((size_t)bi * n + i) * fdims[j],
                       sizeof(float) * fdims[j]);
        }
}
