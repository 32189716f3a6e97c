/* pgk.h -- C ABI of libpgk.so: the sm_100a kernels behind the Progressive-GAN
 * G + D + WGAN-GP training step.
 *
 * The reference (deepsound-project/pggan-pytorch) has no FFI of its own: all of
 * its arithmetic is PyTorch library calls made from network.py / wgan_gp_loss.py.
 * Each entry point below replaces the library call(s) named in its comment
 * (file:line into the reference); INTEGRATION.md shows the ctypes binding a
 * maintainer of the reference would add.
 *
 * Conventions
 *  - every function ENQUEUES work on `stream` (a cudaStream_t) and returns
 *    immediately: no allocation, no hidden synchronisation, no host<->device copy.
 *  - return value 0 = ok; non-zero = error code, text via pgk_last_error().
 *  - all pointers are DEVICE pointers owned by the caller for the duration of
 *    the call.
 *  - "planes" tensors are the internal activation format: P planes (1, 2 or 3)
 *    of bfloat16 in N,H,W,C order (channels innermost, C % 8 == 0).  value =
 *    plane0 + plane1 + plane2, plane k = bf16(value - earlier planes).  P = 3 is
 *    the fp32-faithful mode (24 mantissa bits; the tensor-core kernels form the
 *    six products of planes i, j with i + j <= 2), P = 2 keeps 16 bits (three
 *    products), P = 1 is the bf16 mode.  `*_ps` arguments are the plane stride
 *    in ELEMENTS.
 *  - "image" tensors are the reference's own surface format: fp32, N,C,H,W.
 *  - LeakyReLU slope is 0.2 (network.py:27); lrelu'(v) = v > 0 ? 1 : 0.2.
 */
#ifndef PGK_H
#define PGK_H

#ifdef __cplusplus
extern "C" {
#endif

typedef void* pgk_stream_t; /* cudaStream_t */

#define PGK_OK 0
#define PGK_ERR_ARG 1
#define PGK_ERR_CUDA 2
#define PGK_ERR_ARCH 3

/* weight layouts (`kind`) understood by pgk_prep_weight / pgk_unprep_grad */
#define PGK_W_CONV 0   /* KSxKS "same" conv, KS in {1,3}   (network.py:16,34)            */
#define PGK_W_GFIRST 1 /* G's 4x4 pad-3 conv on a 1x1 input = dense GEMM (network.py:47)  */
#define PGK_W_DLAST 2  /* D's 4x4 valid conv on a 4x4 input = dense GEMM (network.py:163) */

int pgk_version(void);
const char* pgk_last_error(void);
/* 0 iff `device` is compute capability 10.x (the library carries sm_100a SASS only). */
int pgk_arch_check(int device);
/* number of kernels launched by this library since load / since the last reset (bench's gpu_launches). */
long long pgk_launch_count(void);
void pgk_reset_launch_count(void);
/* add n to the counter: kernels replayed from a captured CUDA graph do not pass through the entry points */
void pgk_count_launch(int n);
/* Programmatic dependent launch: every kernel of the library is launched with the programmatic-stream-serialization
 * attribute and runs griddepcontrol.wait before its first global read, so a kernel's prologue overlaps its predecessor's
 * tail (the reference's default-stream ordering, trainer.py:85-115, is kept: nothing is read before the predecessor has
 * completed).  set = 0 / 1 switches it off / on, set < 0 queries; returns the state.  Default: on unless PGK_PDL=0. */
int pgk_pdl_state(int set);
/* per-launch device timing of the GEMM-shaped kernels (bench.py's roofline): while enabled, pgk_conv / pgk_wgrad
 * bracket their launch with CUDA events on `stream`.  pgk_prof_read synchronises on the recorded events and returns
 * the summed algorithmic FLOPs (2*M*N*K of each launch), the summed algorithmic HBM bytes (operand planes read +
 * output planes written, each once), the summed device milliseconds and the launch count of one kernel family;
 * pgk_prof_reset drops the records. */
#define PGK_PROF_CONV 0       /* forward conv / data gradient on tcgen05 (pgk_conv)  */
#define PGK_PROF_WGRAD 1      /* weight gradient on tcgen05 (pgk_wgrad)              */
#define PGK_PROF_CONV_SIMT 2  /* pgk_conv launches served by the CUDA-core kernel    */
#define PGK_PROF_WGRAD_SIMT 3 /* pgk_wgrad launches served by the CUDA-core kernel   */
#define PGK_PROF_CONV_THIN 4  /* pgk_conv launches served by the thin-layer tcgen05 kernel (Cin 8/16/32) */
#define PGK_PROF_WGRAD_THIN 5 /* pgk_wgrad launches served by the thin-layer tcgen05 kernel              */
void pgk_prof_enable(int on);
int pgk_prof_read(int family, double* flops, double* bytes, double* ms, long long* launches);
/* the bf16 tensor-core FLOPs the recorded launches of a family ISSUED: 2*M*N*K x the number of plane products
 * (Pr = 1, 2, 3 planes read -> 1, 3, 6 products per algorithmic FLOP); 0 for the CUDA-core families.  This is what
 * the tensor pipe's own peak bounds -- in the fp32-faithful mode it is 3..6x the algorithmic figure above. */
int pgk_prof_read_products(int family, double* product_flops);
void pgk_prof_reset(void);
/* 0 routes every shape to the CUDA-core kernels (A/B comparisons; also PGK_TC=0 in the environment). */
void pgk_set_tc(int on);

/* ---- equalised-LR weights: fold c into the weight, re-lay for the kernels --------------
 * replaces `h = x * self.c` (network.py:33) + the cuDNN filter transform.
 * w: (Cout, cin_stride, KH, KW) fp32 as PyTorch stores it; only input channels [0, cin) are used.
 * wf: forward operand  [K][Cout]   K = taps*cin  (PGK_W_CONV), cin (GFIRST, Cout' = 16*cout), 16*cin (DLAST)
 * wb: backward operand [K'][cin']  the transposed / tap-flipped operand for the data gradient. */
int pgk_prep_weight(const float* w, float c, int kind, int cin, int cin_stride, int cout, int ks,
                    float* wf, float* wb, pgk_stream_t stream);
/* tensor-core operand: out[p][n][k] = bf16 plane p of w[k][n]  (w = wf or wb above, fp32 [K][Nn]; out: P planes,
 * out_ps elements apart, each [Nn][K] K-major -- what the TMA descriptors of pgk_conv read). */
int pgk_pack_operand(const float* w, int K, int Nn, void* out, long long out_ps, int P, pgk_stream_t stream);
/* the same for the thin-layer kernel (KS = 3, Cin in {8,16,32}, Cout in {8,16,32,64}): out = P planes of
 * pgk_pack_thin_plane_elems(Cin, Cout) bf16 each, in the K = 16 step order of csrc/pgk_conv_thin.cu.  For such channel
 * counts pgk_conv's `wt` must point to THIS packing (wt_ps = the plane size). */
long long pgk_pack_thin_plane_elems(int Cin, int Cout);
int pgk_pack_thin(const float* w, int Cin, int Cout, void* out, long long out_ps, int P, pgk_stream_t stream);
/* All of the above for every conv layer of a network in ONE launch (what the engine calls after each optimizer step;
 * the three entry points above stay for single layers / tools and define the result bit for bit):
 *   w, c, kind, cin, cin_stride, cout, ks   as pgk_prep_weight
 *   planes                                  bf16 planes written to F / B (1 = bf16 mode, 3 = fp32-faithful mode)
 *   wf, wb                                  the fp32 operands of pgk_prep_weight, or NULL (only the CUDA-core kernels read them)
 *   F, F_ps / B, B_ps                       pgk_pack_operand(wf) / pgk_pack_operand(wb), `planes` planes F_ps / B_ps elements
 *                                           apart; thinF / thinB = 1: pgk_pack_thin's layout instead (the buffer's padding
 *                                           entries must have been zeroed once: they are not written).  B may be NULL.
 *   F16, F16_ps                             pgk_pack_operand_fp16(wf) (two half planes), or NULL */
typedef struct PgkPrepLayer {
    const float* w;
    float c;
    int kind, cin, cin_stride, cout, ks, planes;
    float* wf;
    float* wb;
    void* F;
    long long F_ps;
    void* B;
    long long B_ps;
    void* F16;
    long long F16_ps;
    int thinF, thinB;
} PgkPrepLayer;
int pgk_prep_multi(const PgkPrepLayer* layers, int n, pgk_stream_t stream);
/* inverse map for weight gradients: dw (PyTorch layout) (+)= c * dwp (wf layout). */
int pgk_unprep_grad(const float* dwp, float c, int kind, int cin, int cin_stride, int cout, int ks,
                    float* dw, int accumulate, pgk_stream_t stream);
/* the same for every layer of a network in one launch (the arguments of pgk_unprep_grad, one entry per layer) */
typedef struct PgkUnprepLayer {
    const float* dwp;
    float c;
    int kind, cin, cin_stride, cout, ks;
    float* dw;
    int accumulate;
} PgkUnprepLayer;
int pgk_unprep_multi(const PgkUnprepLayer* layers, int n, pgk_stream_t stream);

/* ---- the convolution (network.py:34, F.conv2d through cuDNN) ---------------------------
 * out[n,y,x,co] = E( sum_{tap,ci} X[n, y+dy, x+dx, ci] * wf[tap*Cin+ci][co] ), zero padded, stride 1.
 *   ups = 1: X is the nearest-neighbour 2x upsample of x (x has H/2 x W/2)  (network.py:127,129)
 *   E(v) = v + bias[co] + pos_s[n] * posT[(y*W+x)*Cout + co]   (each term optional)
 *          then LeakyReLU if act == 1                          (network.py:36)
 *          then * lrelu'(mask_ref[n,y,x,co]) if mask_ref       (the backward of network.py:36)
 *          then * out_scale
 * The same entry point computes data gradients (x = output gradient, wf = wb of the layer,
 * mask_ref = the stored input activation of the layer) and the gradient-penalty's second chain.
 * wt / wt_ps: the same operand packed by pgk_pack_operand (3 planes).  Shapes with Cin % 64 == 0, Cout % 16 == 0,
 * power-of-two H, W and ups == 0 run on the TMA + tcgen05 kernel and read wt; KS = 3 layers with Cin in {8,16,32},
 * Cout in {8,16,32,64}, W % 128 == 0 run on the row-streaming thin-layer tcgen05 kernel and read wt in
 * pgk_pack_thin's layout; all others run on the CUDA-core implicit GEMM and read wf.  Either pointer may be NULL if the shape never takes that path.
 * pn_r (optional, fp32 per output pixel): the generator's pixel norm after the activation (network.py:37-40):
 *          out = E(v) * r, r = rsqrt(mean_co(E(v)^2) + 1e-8), r stored for the backward pass.  Requires
 *          mask_ref == NULL and out_scale == 1.  On the thin-layer kernel (Cout <= 32) it is part of the epilogue;
 *          on the other kernels pgk_conv runs pgk_pixelnorm in place right after the convolution.
 * Pr (1 <= Pr <= P): how many planes of x and wt the tensor-core kernel READS (products of planes i + j < Pr).
 * Forward passes, whose values decide the LeakyReLU masks, use Pr = P; the gradient chains use Pr = min(P, 2)
 * (16 mantissa bits, three products instead of six) -- gradients are continuous in these operands. */
int pgk_conv(const void* x, int P, int Pr, long long x_ps, int N, int H, int W, int Cin, int Cout, int KS, int ups,
             const float* wf, const void* wt, long long wt_ps, const float* bias, const float* posT,
             const float* pos_s, int act, const void* mask_ref, long long mask_ps, float out_scale, void* out,
             long long out_ps, float* pn_r, pgk_stream_t stream);

/* The row-streaming thin-layer kernel called directly (pgk_conv dispatches to it for Cin in {8, 16, 32}).  Additionally,
 * in the one-plane mode (P = Pr = 1), Cin = 64 with Cout in {32, 64}: layers that pgk_conv gives to the wide kernel, where
 * an MMA of N = 32 ... 64 runs at a quarter of the tensor rate; here the three filter rows are stacked along N.  wpack:
 * pgk_pack_thin's packing (wpack_ps = pgk_pack_thin_plane_elems(Cin, Cout)); arguments otherwise as pgk_conv's. */
int pgk_conv_thin(const void* x, int P, int Pr, long long x_ps, int N, int H, int W, int Cin, int Cout, const void* wpack,
                  long long wpack_ps, const float* bias, int act, const void* mask_ref, long long mask_ps, float out_scale,
                  void* out, long long out_ps, float* pn_r, pgk_stream_t stream);

/* ---- forward convolution on IEEE-half operand planes (experimental; fp32-faithful modes only) ----------------------
 * The fp32-faithful mode pays six bf16 products per forward FLOP (three planes, i + j <= 2) because two bf16 planes
 * (16 bits) flip too many LeakyReLU units.  Two fp16 planes carry 22 bits, so the three products hi*hi, hi*lo, lo*hi
 * do the same job at half the tensor work (tests/dev/precision_model.py: gradient error 5e-7 against 4e-3 for two
 * bf16 planes, on the same three products).  Only the OPERANDS of forward convolutions change format; everything
 * stored stays bf16 planes:
 *   pgk_cvt_fp16x2      dst = {half(v), half(v - hi)} of v = sum of the P bf16 planes of src (count elements per plane)
 *   pgk_pack_operand_fp16  like pgk_pack_operand, P <= 2 half planes of w * 2^PGK_FP16_WSHIFT (the low plane of the
 *                       small equalised-LR weights would otherwise fall into the fp16 subnormals)
 *   pgk_conv_fp16       pgk_conv for a forward pass (no mask, out_scale 1, no upsample) reading xh / wth as produced by
 *                       the two calls above and writing P (2 or 3) bf16 planes; shapes of the wide tensor-core kernel
 *                       only (error otherwise -- the caller keeps pgk_conv for the rest).  E(v) as in pgk_conv. */
#define PGK_FP16_WSHIFT 6
int pgk_cvt_fp16x2(const void* src, long long src_ps, int P, long long count, void* dst, long long dst_ps,
                   pgk_stream_t stream);
int pgk_pack_operand_fp16(const float* w, int K, int Nn, void* out, long long out_ps, int P, pgk_stream_t stream);
int pgk_conv_fp16(const void* xh, long long xh_ps, int N, int H, int W, int Cin, int Cout, int KS, const void* wth,
                  long long wth_ps, const float* bias, const float* posT, const float* pos_s, int act, void* out, int P,
                  long long out_ps, pgk_stream_t stream);
/* 1 iff pgk_conv would run this shape on the wide tensor-core kernel (what pgk_conv_fp16 accepts) */
int pgk_conv_tc_supported(int N, int H, int W, int Cin, int Cout, int KS, int ups);

/* ---- weight gradient (cuDNN convolution_backward, weight part) --------------------------
 * dwp[tap*Cin+ci][co] += sum over the listed sample groups of X[..., ci] (shifted by tap) * g[..., co].
 * Groups: ngroups (<= 4) groups of group_n samples; group i reads x samples starting at xoff[i] and g samples
 * starting at goff[i].  dwp must be zeroed by the caller (fp32, wf layout).
 * db (optional): the layer's bias gradient db[co] += sum over the pixels of the groups whose bit is set in
 * bias_groups of g[..., co] (also accumulated: the caller zeroes it).  The thin-layer kernel produces it as one more
 * accumulator row of the same launch; the other kernels are followed by pgk_bias_grad. */
int pgk_wgrad(const void* x, long long x_ps, const void* g, long long g_ps, int P, int Pr, int H, int W, int Cin,
              int Cout, int KS, int ups, int ngroups, int group_n, const int* xoff, const int* goff, float* dwp,
              float* db, unsigned bias_groups, pgk_stream_t stream);
/* db[co] (+)= scale * sum over pixels of the listed sample groups of g[..., co]. */
int pgk_bias_grad(const void* g, long long g_ps, int P, int HW, int Cout, int ngroups, int group_n, const int* goff,
                  float scale, float* db, int accumulate, pgk_stream_t stream);

/* out16 / y16 (optional, NULL = off) of pgk_from_rgb, pgk_pool2, pgk_mask_mul and pgk_pixelnorm: a second output, the
 * two IEEE-half planes (out16_ps elements apart) that pgk_cvt_fp16x2 would derive from the bf16 planes just stored --
 * bit for bit -- for the wide forward conv that reads this tensor next (pgk_conv_fp16); saves the conversion pass. */

/* ---- 1x1 convs against the image surface ------------------------------------------------
 * fromRGB (network.py:145,160): out[n,y,x,co] = E(sum_c c*w[co][c] * img[n,c,y,x]); E as in pgk_conv
 * (bias, act, mask_ref).  w is the PyTorch (Cout, C, 1, 1) tensor, scale c applied here. */
int pgk_from_rgb(const float* img, int N, int C, int H, int W, int Cout, const float* w, float c, const float* bias,
                 int act, const void* mask_ref, long long mask_ps, void* out, int P, long long out_ps, void* out16,
                 long long out16_ps, pgk_stream_t stream);
/* data gradient of fromRGB to the image: dimg[n,c,y,x] (+)= scale * sum_co c*w[co][c] * g[n,y,x,co];
 * ups = 1: g has H/2 x W/2 and is read at (y/2, x/2) (the avg-pooled low-res branch, network.py:231-232). */
int pgk_from_rgb_dgrad(const void* g, int P, long long g_ps, int N, int C, int H, int W, int Cout, const float* w,
                       float c, float scale, int ups, int accumulate, float* dimg, pgk_stream_t stream);
/* DEVICE-SIDE FADE-IN SCALARS.  DepthManager changes alpha every iteration of a transition phase (plugins.py:57-81), and
 * alpha enters five entry points as a host scalar (the lerps of network.py:131-138, 230-233 and their derivatives).  So
 * that ONE captured CUDA graph can be replayed over a whole phase, each of them also takes optional device pointers
 * d_*: when non-NULL the corresponding host scalar is multiplied by the float they point to at execution time (the
 * caller passes the alpha-free part as the host scalar and keeps {alpha, 1 - alpha} in a two-float device buffer that
 * it rewrites before every replay).  NULL = the host scalar alone, as before. */
/* toRGB (network.py:49,65) with the generator's fade-in (network.py:131-138):
 * img[n,c,y,x] = a_hi * (sum_k c_hi*w_hi[c][k]*h[n,y,x,k] + b_hi[c])
 *              + a_lo * (sum_k c_lo*w_lo[c][k]*h_lo[n,y/2,x/2,k] + b_lo[c])     (second term iff h_lo != NULL) */
int pgk_to_rgb(const void* h, int P, long long h_ps, int N, int H, int W, int Cin, const float* w_hi, float c_hi,
               const float* b_hi, float a_hi, const void* h_lo, long long hlo_ps, int Cin_lo, const float* w_lo,
               float c_lo, const float* b_lo, float a_lo, int C, float* img, const float* d_a_hi, const float* d_a_lo,
               pgk_stream_t stream);
/* data gradient of toRGB: dh[n,y,x,k] = scale * sum_c c*w[c][k] * dimg(n,c,y,x); pool = 1: dimg is summed over the
 * 2x2 block (2y..2y+1, 2x..2x+1) of a 2H x 2W image (the backward of toRGB_prev(upsample(h))). */
int pgk_to_rgb_dgrad(const float* dimg, int N, int C, int H, int W, int Cin, const float* w, float c, float scale,
                     int pool, void* dh, int P, long long dh_ps, const float* d_scale, pgk_stream_t stream);
/* weight/bias gradients of both 1x1 families in one pass over pixels:
 * dw[a*sa + k*sk] += scale_w * sum_{n,pix} IMG(n,a,pix) * t[n,pix,k];  d_colsum[k] += scale_b*sum t;  d_imgsum[a] += scale_b*sum IMG
 * IMG = img, or its 2x2 block sum when pool = 1 (img is then 2H x 2W).  Any output pointer may be NULL. */
int pgk_rgb_wgrad(const float* img, int img_n0, const void* t, int P, long long t_ps, int t_n0, int N, int C, int H,
                  int W, int K, int pool, float scale_w, float scale_b, float* dw, int sa, int sk, float* d_colsum,
                  float* d_imgsum, const float* d_scale, pgk_stream_t stream);

/* ---- elementwise / reductions on planes ------------------------------------------------- */
/* out = a * pool2x2(src) [+ b * other]; avg = 1: mean of the block (F.avg_pool2d, network.py:229,238),
 * avg = 0: sum (backward of the nearest upsample).  src is N x 2H x 2W x C, out/other N x H x W x C. */
int pgk_pool2(const void* src, long long src_ps, int P, int N, int H, int W, int C, int avg, float a,
              const void* other, long long other_ps, float b, void* out, long long out_ps, const float* d_a,
              const float* d_b, void* out16, long long out16_ps, pgk_stream_t stream);
/* out[n,y,x,c] = scale * src[n, y>>ups, x>>ups, c] * lrelu'(ref[n,y,x,c]) (ref optional) */
int pgk_mask_mul(const void* src, long long src_ps, int P, int N, int H, int W, int C, int ups, float scale,
                 const void* ref, long long ref_ps, void* out, long long out_ps, const float* d_scale, void* out16,
                 long long out16_ps, pgk_stream_t stream);
/* out = a*x + b*y (y optional) on planes with `count` elements per plane */
int pgk_axpby(const void* x, long long x_ps, float a, const void* y, long long y_ps, float b, int P, long long count,
              void* out, long long out_ps, pgk_stream_t stream);
/* pixel norm (network.py:37-40): y = h * r, r = rsqrt(mean_c(h^2) + 1e-8); in place allowed; r (fp32 per pixel) stored. */
int pgk_pixelnorm(const void* h, long long h_ps, int P, long long npix, int C, void* y, long long y_ps, float* r,
                  void* y16, long long y16_ps, pgk_stream_t stream);
/* backward of LeakyReLU -> pixel norm: da = r * (dy - y * mean_c(dy*y)) * lrelu'(y)  */
int pgk_pixelnorm_bwd(const void* dy, long long dy_ps, const void* y, long long y_ps, const float* r, int P,
                      long long npix, int C, void* da, long long da_ps, pgk_stream_t stream);
/* latent normalisation (network.py:119-123): fp32 (N, L) -> planes (N,1,1,L) */
int pgk_latent_norm(const float* z, int N, int L, int normalize, void* out, int P, long long out_ps,
                    pgk_stream_t stream);

/* ---- minibatch stddev (network.py:174-187) and its first / second derivatives ------------
 * stats[g*4 + {0,1,2,3}] = {mean, s, 1/(n*s), n} over group g (group_n samples x HWC values each);
 * svec[g*group_n + i] = s_g (the per-sample scalar pgk_conv's pos_s wants; optional). */
int pgk_stddev_stats(const void* h, long long h_ps, int P, int ngroups, long long group_count, float* stats,
                     float* svec, int group_n, pgk_stream_t stream);
/* q[g] = sum over group g of ua[n,y,x,co] * posT[(y*W+x)*C + co]   (the gradient arriving at the stddev scalar) */
int pgk_group_dot_pos(const void* ua, long long ua_ps, int P, int ngroups, int group_n, int HW, int C,
                      const float* posT, float* q, pgk_stream_t stream);
/* first-order backward: dh += q[g] * (h - mean_g) / (n s_g) */
int pgk_stddev_bwd(const void* h, long long h_ps, const float* stats, const float* q, int P, int ngroups,
                   long long group_count, void* dh, long long dh_ps, pgk_stream_t stream);
/* second-order terms for ONE group (the mixed samples): given v (cotangent of the data-gradient at h),
 *   e  = <v, (h-mean)/(n s)>                       -> ev[0..ev_n)  (value of the extra channel in the v-chain)
 *   wh = q/(n s) * (v - mean(v)) - q*e*(h-mean)/(n s^2)        (cotangent entering the forward graph at h) */
int pgk_stddev_bwd2(const void* h, long long h_ps, const void* v, long long v_ps, const float* stats,
                    const float* q, int P, long long group_count, float* ev, int ev_n, void* wh, long long wh_ps,
                    float* scratch, pgk_stream_t stream);
/* gradient of the extra (stddev) input channel's filter taps of D's last 3x3 conv:
 * dw[co][ch][ky][kx] += c * sum_{n,y,x : (y+ky-1,x+kx-1) inside} coef[n] * ua[n,y,x,co];   dw has cin_stride channels */
int pgk_posbias_wgrad(const void* ua, long long ua_ps, int P, int N, int H, int W, int Cout, const float* coef,
                      float c, int cin_stride, int ch, float* dw, pgk_stream_t stream);
/* posT[(y*W+x)*Cout+co] = c * sum_{taps inside at (y,x)} w[co][ch][ky][kx] */
int pgk_prep_posbias(const float* w, float c, int cin_stride, int ch, int Cout, int H, int W, float* posT,
                     pgk_stream_t stream);

/* ---- head: nn.Linear(512,1) (network.py:219,239) ------------------------------------------ */
int pgk_linear_fwd(const void* h, long long h_ps, int P, int N, int K, const float* w, const float* b, float* scores,
                   pgk_stream_t stream);
/* ua[n,k] = seed[n] * w[k] * lrelu'(h[n,k]);  dw[k] += sum_n wseed[n]*h[n,k];  db += sum_n wseed[n]  (wseed optional) */
int pgk_linear_bwd(const void* h, long long h_ps, int P, int N, int K, const float* w, const float* seed,
                   const float* wseed, void* ua, long long ua_ps, float* dw, float* db, pgk_stream_t stream);
/* dw[k] += sum_n v[n,k]   (gradient-penalty term of the head's weight) */
int pgk_colsum(const void* v, long long v_ps, int P, int N, int K, float scale, float* dw, pgk_stream_t stream);

/* ---- WGAN-GP algebra (wgan_gp_loss.py) ---------------------------------------------------- */
/* mixed = real*(1-eps_n) + fake*eps_n per sample (wgan_gp_loss.py:8-10,19); per = C*H*W */
int pgk_interpolate(const float* real, const float* fake, const float* eps, int N, long long per, float* mixed,
                    pgk_stream_t stream);
/* scores: [real N | fake N | mixed N].  Writes (wgan_gp_loss.py:48,55; trainer.py:98):
 *   d_real_loss[n] = -Dr + eps_drift*Dr^2 ; d_fake_loss[n] = Df
 *   seed[3N]: d cost / d score = {(-1 + 2 eps_drift Dr)/N, 1/N, 1 (the grad_outputs of wgan_gp_loss.py:21-23)}
 *   wseed[3N]: the same with 0 for the mixed samples (their scores do not enter the cost directly) */
int pgk_d_loss_seed(const float* scores, int N, float eps_drift, float* d_real_loss, float* d_fake_loss, float* seed,
                    float* wseed, pgk_stream_t stream);
/* out[0] = scale * mean(x[0..n))   (wgan_gp_loss.py:72-73: G_cost = mean(-D(G(z)))) */
int pgk_mean_scale(const float* x, int n, float scale, float* out, pgk_stream_t stream);
/* per-sample ||g||_2 of an image-shaped gradient, the penalty (wgan_gp_loss.py:29-31) and the cotangent
 *   v0 = (1/N) * 2*lambda*(nrm - T)/(T^2 * nrm) * g ; also writes cost = mean(real)+mean(fake)+mean(gp).
 * `norms` must hold 2N floats: [0,N) receives the norms, [N,2N) is scratch. */
int pgk_gp_penalty(const float* g, int N, long long per, float lambda, float target, const float* d_real_loss,
                   const float* d_fake_loss, float* norms, float* gp, float* v0, float* cost, pgk_stream_t stream);
/* generic fp32 helpers */
int pgk_fill(float* p, long long n, float v, pgk_stream_t stream);
/* out[n,c,y,x] = scale * (mean | sum) of the 2x2 block of img; H, W are the OUTPUT sizes (img is 2H x 2W) */
int pgk_pool_img(const float* img, int N, int C, int H, int W, int avg, float scale, float* out, pgk_stream_t stream);
/* dst[n,c,y,x] (+)= scale * src[n,c,y>>1,x>>1] (fp32 images; H, W are dst sizes) */
int pgk_unpool_img_add(const float* src, int N, int C, int H, int W, float scale, int accumulate, float* dst,
                       pgk_stream_t stream);

/* ---- real-image preparation (the step just before the path) -----------------------------------
 * dataset.py:60-67: alpha_fade (dataset.py:109-113: 2x2 box mean -> nearest upsample -> lerp by 1 - alpha, applied iff
 * alpha < 1) followed by utils.adjust_dynamic_range (utils.py:24-30).  src: N,C,H,W uint8 (src_is_u8) or fp32. */
int pgk_real_prep(const void* src, int src_is_u8, int N, int C, int H, int W, double alpha, double min_in,
                  double max_in, double min_out, double max_out, float* out, pgk_stream_t stream);

/* ---- optimizer (the step just after the path) ----------------------------------------------------
 * torch.optim.Adam as wired by train.py:148-149,195 for every parameter with a gradient, in ONE launch.
 * table (device): ntensors (<= 1024) rows of 8 x 64-bit words {param ptr, grad ptr, exp_avg ptr, exp_avg_sq ptr, numel,
 * float bits of lr / (1 - beta1^t), float bits of 1 / sqrt(1 - beta2^t), first block}; all tensors fp32 contiguous.
 * Block b of the one-dimensional grid updates PGK_ADAM_CHUNK elements of the tensor whose block range holds it:
 * first block of row 0 = 0, of row i+1 = first block of row i + ceil(numel_i / PGK_ADAM_CHUNK); total_blocks = the sum.
 * pgk_adam_chunk() returns PGK_ADAM_CHUNK. */
#define PGK_ADAM_CHUNK 4096
int pgk_adam_chunk(void);
int pgk_adam_multi(const void* table, int ntensors, long long total_blocks, float beta1, float beta2, float eps,
                   pgk_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PGK_H */
