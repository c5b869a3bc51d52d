/*
 * rsm.h -- C ABI of librsm_b200.so: B200 (sm_100a) kernels for the stereo-matching hot path
 * of babiking/realtime_stereo_matcher (cost-volume construction + disparity regression).
 *
 * The reference has no FFI / plugin registry: its boundary for this path is plain Python
 * call signatures.  Each entry point below names the reference function it
 * replaces (file:line relative to the reference checkout).  The Python mirror of those
 * signatures lives in realtime_stereo_matcher_b200/ and is the only caller; the binding a
 * maintainer of the reference would add is shown in INTEGRATION.md.
 *
 * Conventions
 *   - all pointers are DEVICE pointers on `device`; nothing is allocated, freed or
 *     synchronised inside the library; work is enqueued on `stream` (a cudaStream_t passed
 *     as void*; NULL = the legacy default stream) and the call returns immediately.
 *   - inputs are borrowed and never written; outputs are dense (contiguous) tensors in the
 *     reference's own layouts, fully overwritten (fill-value regions included).
 *   - feature maps may be non-contiguous views (the v4 forward passes width-cropped slices,
 *     model/mobile_stereo_net_v4.py:446): they are described by rsm_feat with strides in
 *     ELEMENTS.
 *   - return value: RSM_OK (0) or an rsm_status code; rsm_last_error(code) returns a
 *     static description (plus the CUDA error text of the calling thread's last failure).
 *   - no global mutable state; safe to call concurrently from several host threads on
 *     different devices (the reference trains under nn.DataParallel, train_stereo.py:139).
 *   - there is NO CPU fallback.
 */
#ifndef RSM_B200_H_
#define RSM_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RSM_VERSION 107 /* major*10000 + minor*100 + patch */

typedef enum rsm_dtype {
  RSM_F32 = 0,
  RSM_F16 = 1,
  RSM_BF16 = 2
} rsm_dtype;

typedef enum rsm_status {
  RSM_OK = 0,
  RSM_ERR_INVALID_SHAPE = 1,    /* negative / inconsistent sizes, C % G != 0, index overflow */
  RSM_ERR_UNSUPPORTED_DTYPE = 2,
  RSM_ERR_NULL_POINTER = 3,
  RSM_ERR_CUDA = 4,             /* a CUDA runtime call or launch failed */
  RSM_ERR_MISALIGNED = 5,       /* dense output / gradient pointer not aligned to its element */
  RSM_ERR_UNSUPPORTED_CONFIG = 6,
  RSM_ERR_IO = 7                /* a file could not be opened, read or written (rsm_pfm_*) */
} rsm_status;

/* how a correlation is normalised over the reduced channels */
typedef enum rsm_reduce {
  RSM_REDUCE_SUM = 0,  /* TorchInnerProductCost: torch.sum(dim=1), cost_volume/inner_product.py:38-40 */
  RSM_REDUCE_MEAN = 1  /* make_correlation_volume / groupwise: .mean(), mobile_disp_net_c.py:197-201 */
} rsm_reduce;

/* a (N,C,H,W) feature map on the device, strides in elements */
typedef struct rsm_feat {
  const void* data;
  int64_t stride_n, stride_c, stride_h, stride_w;
} rsm_feat;

/* optional outputs of the regression kernels; any pointer may be NULL (not produced) */
typedef struct rsm_regress_out {
  void* soft;        /* (N,H,W) expectation sum_d d*softmax_d(+cost), dtype = cost dtype   */
  int64_t* argmin;   /* (N,H,W) torch.argmin(cost, 1): first index on ties, NaN wins       */
  int64_t* argmax;   /* (N,H,W) torch.argmax(cost, 1)                                       */
  float* lse;        /* (N,H,W) fp32 log-sum-exp over D, saved for the backward pass        */
  float* expect;     /* (N,H,W) the expectation again, ALWAYS fp32 (not rounded to a 16-bit cost dtype):
                      * what the backward pass reads, and what a caller under autocast returns -- the reference's
                      * F.softmax runs in fp32 there, so its disparity is fp32 (mobile_stereo_net.py:144-147)  */
} rsm_regress_out;

int rsm_version(void);
const char* rsm_last_error(int code);

/* ---- concatenate volume: TorchConcatenateCost.forward, cost_volume/concatenate.py:11-41
 * out (N,2C,H,W,D): out[:, :C,y,x,d] = L[:,:,y,x], out[:, C:,y,x,d] = R[:,:,y,x-d] for x>=d, else 0 */
int rsm_concat_fwd(rsm_feat left, rsm_feat right, void* out, int64_t N, int64_t C, int64_t H,
                   int64_t W, int64_t D, int dtype, int device, void* stream);
/* adjoint (autograd through concatenate.py:33-39): gout (N,2C,H,W,D) -> gleft, gright (N,C,H,W) */
int rsm_concat_bwd(const void* gout, void* gleft, void* gright, int64_t N, int64_t C, int64_t H,
                   int64_t W, int64_t D, int dtype, int device, void* stream);

/* ---- interweave: TorchInterweaveCost.forward, cost_volume/interweave.py:10-22 and
 * interweave_tensors, model/mobile_stereo_net_v4.py:17-23.  out (N,2C,H,W), even ch = L, odd = R */
int rsm_interweave_fwd(rsm_feat left, rsm_feat right, void* out, int64_t N, int64_t C, int64_t H,
                       int64_t W, int dtype, int device, void* stream);
int rsm_interweave_bwd(const void* gout, void* gleft, void* gright, int64_t N, int64_t C,
                       int64_t H, int64_t W, int dtype, int device, void* stream);

/* ---- inner product / correlation: TorchInnerProductCost.forward, cost_volume/inner_product.py:11-42
 * (reduce = SUM) and make_correlation_volume, model/mobile_disp_net_c.py:188-205 (reduce = MEAN).
 * out (N,D,H,W) = s * sum_c L[c,x] R[c,x-d] for x>=d, else 0; fp32 accumulation */
int rsm_inner_fwd(rsm_feat left, rsm_feat right, void* out, int64_t N, int64_t C, int64_t H,
                  int64_t W, int64_t D, int reduce, int in_dtype, int out_dtype, int device,
                  void* stream);
/* gout (N,D,H,W) dense in out_dtype; gleft/gright (N,C,H,W) dense in in_dtype; either may be NULL */
int rsm_inner_bwd(const void* gout, rsm_feat left, rsm_feat right, void* gleft, void* gright,
                  int64_t N, int64_t C, int64_t H, int64_t W, int64_t D, int reduce, int in_dtype,
                  int out_dtype, int device, void* stream);
/* diagnostic twin: `prof` = 16 zero-initialised uint64 on the device; the tcgen05 adjoint kernel (16-bit tensors;
 * every 64-disparity launch of it) adds clock64 cycles per warp role, summed over CTAs: issuer [0] waiting for a free accumulator, [1] for a band
 * matrix, [2] for feature atoms, [3] total; builder warp 0 [4] waiting for the gradient tile, [5] for a free band matrix,
 * [6] building, [7] total; epilogue warp 0 [8] waiting for an accumulator, [9] total */
int rsm_inner_bwd_profile(const void* gout, rsm_feat left, rsm_feat right, void* gleft, void* gright,
                  int64_t N, int64_t C, int64_t H, int64_t W, int64_t D, int reduce, int in_dtype,
                  int out_dtype, int device, void* stream, uint64_t* prof);

/* ---- group-wise correlation: TorchGroupwiseCost.forward / .groupwise, cost_volume/groupwise.py:12-56
 * out (N,G,H,W,D) = (1/(C/G)) * sum_{c in group g} L[c,x] R[c,x-d] for x>=d, else 0 */
int rsm_groupwise_fwd(rsm_feat left, rsm_feat right, void* out, int64_t N, int64_t C, int64_t H,
                      int64_t W, int64_t D, int64_t G, int in_dtype, int out_dtype, int device,
                      void* stream);
int rsm_groupwise_bwd(const void* gout, rsm_feat left, rsm_feat right, void* gleft, void* gright,
                      int64_t N, int64_t C, int64_t H, int64_t W, int64_t D, int64_t G,
                      int in_dtype, int out_dtype, int device, void* stream);

/* ---- difference volume: make_cost_volume, model/mobile_stereo_net.py:8-27 (= v2 :8-27, v3 :9-28)
 * out (N,C,D,H,W) = L[c,y,x] - R[c,y,x-d] for x>=d, else `fill` (the reference uses 1.0) */
int rsm_difference_fwd(rsm_feat left, rsm_feat right, void* out, int64_t N, int64_t C, int64_t H,
                       int64_t W, int64_t D, float fill, int dtype, int device, void* stream);
int rsm_difference_bwd(const void* gout, void* gleft, void* gright, int64_t N, int64_t C,
                       int64_t H, int64_t W, int64_t D, int dtype, int device, void* stream);

/* ---- shifted interweave stack (first step): every iteration of MobileStereoNetV4's
 * per-disparity loop, model/mobile_stereo_net_v4.py:444-458, builds
 * interweave_tensors(featL[..., i:], featR[..., :-i]); this entry point builds all D of them at once,
 * full width, zero where x < d (so the zero padding the cropped convolutions saw is reproduced):
 * out (D,N,2C,H,W): out[d,n,2c,y,x] = L[n,c,y,x], out[d,n,2c+1,y,x] = R[n,c,y,x-d] for x >= d, else 0 */
int rsm_shift_interweave_fwd(rsm_feat left, rsm_feat right, void* out, int64_t N, int64_t C, int64_t H,
                             int64_t W, int64_t D, int dtype, int device, void* stream);
/* gleft[n,c,y,x] = sum_{d<=x} gout[d,n,2c,y,x];  gright[n,c,y,x'] = sum_{d, x'+d<W} gout[d,n,2c+1,y,x'+d] */
int rsm_shift_interweave_bwd(const void* gout, void* gleft, void* gright, int64_t N, int64_t C,
                             int64_t H, int64_t W, int64_t D, int dtype, int device, void* stream);

/* ---- MobileStereoNetV4's whole per-disparity learned volume in eval mode: the loop
 * model/mobile_stereo_net_v4.py:443-458 -- for every disparity i: interweave_tensors(featL[..., i:], featR[..., :-i])
 * -> self.conv3d (:317-333: Conv3d(1,16,(8,3,3),stride (8,1,1)) + BN + ReLU, Conv3d(16,32,(4,3,3),stride (4,1,1)) + BN +
 * ReLU, Conv3d(32,16,(2,3,3),stride (2,1,1)) + BN + ReLU) -> self.volume11 (:335: 1x1 conv 16->1 + BN + ReLU) ->
 * volume[:, 0, i, :, i:] -- as three kernels (layer-1 maps once per pair, layers 2 and 3 as tcgen05 implicit GEMMs with
 * 16-bit operands and fp32 accumulation); out (B,D,H,W) in in_dtype, zero where x < d.
 * The caller folds each BatchNorm (eval: y = (conv + bias - mean) * gamma / sqrt(var + eps) + beta) into the weights
 * (scale per output channel) and an additive term t, and packs layers 2 / 3 into the UMMA core-matrix layout:
 *   w1  (16,8,3,3) fp32      = conv3d[0].weight[:,0] * s1[:,None,None,None];            t1 (16) fp32
 *   w2  (9,8,32,8) op_dtype  : [tap = dy*3+dx][chunk][co][e] = conv3d[3].weight[co, ci, kd, dy, dx] * s2[co]
 *                              with kd*16 + ci = chunk*8 + e;                             t2 (32) fp32
 *   w3  (9,8,16,8) op_dtype  : conv3d[6].weight[co, ci, kd, dy, dx] * s3[co] with kd*32 + ci = chunk*8 + e;  t3 (16)
 *   w11 (16) fp32            = volume11[0][0].weight[0,:,0,0] * s11;                      t11 (1) fp32
 * all DEVICE pointers.  op_dtype = RSM_F16 (fp32 / fp16 features; 11-bit significands like TF32, saturating at
 * +-65504) or RSM_BF16.  C must be 32 (2C = 64 = the 8 x 4 x 2 depth the Conv3d kernels collapse). */
typedef struct rsm_v4_weights {
  const float* w1;
  const float* t1;
  const void* w2;
  const float* t2;
  const void* w3;
  const float* t3;
  const float* w11;
  const float* t11;
} rsm_v4_weights;
/* bytes of device scratch rsm_v4_volume_fwd needs (layer-1 maps + the 16-bit layer-2 activations); 256-byte aligned */
int64_t rsm_v4_volume_workspace(int64_t B, int64_t H, int64_t W, int64_t D);
int rsm_v4_volume_fwd(rsm_feat left, rsm_feat right, rsm_v4_weights w, void* out, void* workspace, int64_t B,
                      int64_t C, int64_t H, int64_t W, int64_t D, int in_dtype, int op_dtype, int device,
                      void* stream);

/* diagnostic twin of rsm_v4_volume_fwd: `prof` = 32 zero-initialised uint64 on the device; the two GEMM kernels add
 * clock64 cycles per warp role (waiting for rows / accumulators / ring slots, working, total), summed over CTAs */
int rsm_v4_volume_fwd_profile(rsm_feat left, rsm_feat right, rsm_v4_weights w, void* out, void* workspace, int64_t B,
                              int64_t C, int64_t H, int64_t W, int64_t D, int in_dtype, int op_dtype, int device,
                              void* stream, uint64_t* prof);

/* ---- refinement warp: warp_by_flow_map, model/mobile_stereo_net_v2.py:59-96
 * (= model/mobile_stereo_net_v3.py:60-97, tools/warp.py:5-42; call sites v2 :127, v3 :136).
 * image (N,C,H,W), flow (N,flow_channels,H,W) with flow_channels 1 or 2, all dense, same dtype.
 * out[n,c,y,x] = bilinear sample (zero padding) of image[n,c] at ix = (x - flow[n,0,y,x]) * W/(W-1) - 0.5,
 * iy = (y - flow[n,1,y,x]) * H/(H-1) - 0.5 (flow[n,1] = 0 for one channel) -- the coordinates
 * F.grid_sample(align_corners=False) sees after the reference's (size-1) normalisation */
int rsm_warp_fwd(const void* image, const void* flow, void* out, int64_t N, int64_t C, int64_t H, int64_t W,
                 int flow_channels, int dtype, int device, void* stream);
/* adjoint: gimage (N,C,H,W) is ALWAYS fp32 (cleared and accumulated with atomics inside the call, like
 * ATen's grid_sampler backward); gflow (N,flow_channels,H,W) in `dtype`; either may be NULL (not needed: no
 * memset, no atomics for a NULL gimage), not both */
int rsm_warp_bwd(const void* gout, const void* image, const void* flow, float* gimage, void* gflow, int64_t N,
                 int64_t C, int64_t H, int64_t W, int flow_channels, int dtype, int device, void* stream);

/* ---- pre / post steps either side of the path.  `planes` = N*C images of one size.
 * prepare: model/mobile_stereo_net.py:121-130 (= _v2.py:194-203, _v3.py:296-305, mobile_disp_net_c.py:339-351;
 * _v4.py:433-434 with Hp = H, Wp = W): out (planes,Hp,Wp) = 2 * (img / 255) - 1 for y < H, x < W and 0 in the
 * right / bottom pad (F.pad after the normalisation), each op rounded as the reference's tensor ops */
int rsm_prepare_fwd(const void* img, void* out, int64_t planes, int64_t H, int64_t W, int64_t Hp, int64_t Wp,
                    int dtype, int device, void* stream);
/* gimg (planes,H,W) = (gout[:, :H, :W] * 2) / 255 */
int rsm_prepare_bwd(const void* gout, void* gimg, int64_t planes, int64_t H, int64_t W, int64_t Hp, int64_t Wp,
                    int dtype, int device, void* stream);
/* finalize: model/mobile_stereo_net.py:154-159 (= _v2.py:227-232; mode 0 = F.interpolate's default 'nearest') and
 * disparity_interpolate + crop + negate, model/mobile_disp_net_c.py:223-234 + :408-411 (mode 1 = bilinear,
 * align_corners=False): out (planes,h,w) = -resize(disp * vscale, (Hp,Wp))[:h, :w], disp (planes,hs,ws);
 * the reference passes vscale = Wp / ws.  ATen's index rules: scale = (float)in/out; nearest
 * min((int)floorf(dst*scale), in-1); linear max(scale*(dst+0.5)-0.5, 0) */
int rsm_finalize_fwd(const void* disp, void* out, int64_t planes, int64_t hs, int64_t ws, int64_t Hp, int64_t Wp,
                     int64_t h, int64_t w, float vscale, int mode, int dtype, int device, void* stream);
/* adjoint: gdisp (planes,hs,ws) gathered per source pixel from gout (planes,h,w); deterministic */
int rsm_finalize_bwd(const void* gout, void* gdisp, int64_t planes, int64_t hs, int64_t ws, int64_t Hp, int64_t Wp,
                     int64_t h, int64_t w, float vscale, int mode, int dtype, int device, void* stream);

/* ---- loss and metrics on the device.  Reductions run in two fixed-shape stages through
 * `workspace` (>= RSM_REDUCE_WS_DOUBLES doubles, device memory, contents irrelevant) and are deterministic. */
#define RSM_REDUCE_MAX_BLOCKS 1184
#define RSM_REDUCE_WS_DOUBLES (RSM_REDUCE_MAX_BLOCKS * 8)
/* one term of SequenceLoss.forward, loss/loss.py:36-81: pred (N,1,hs,ws) is resized like
 * F.interpolate(pred * (W / ws), (H, W)) (nearest; identity when the sizes agree, :71-73), compared with gt (N,1,H,W)
 * by L1 (kind 0) or smooth-L1 with beta 1 (kind 1, the last prediction, :75-78) and averaged over
 * valid = (flow_valid >= 0.5) & (|gt| < max_flow) (:55-58); flow_valid (N,H,W) is fp32.
 * result[4] (device, fp64) = {mean, valid count, non-finite predictions seen, infinite valid ground truth} -- the
 * last two are the asserts of :60,:66-67, left to the caller to test with one read */
int rsm_seqloss_fwd(const void* pred, const void* gt, const float* valid, double* workspace, double* result, int64_t N,
                    int64_t hs, int64_t ws, int64_t H, int64_t W, float max_flow, int kind, int dtype, int device,
                    void* stream);
/* gpred (N,1,hs,ws) = gmean[0] / count * d mean / d pred; gmean (device, fp32 scalar), result from the forward call */
int rsm_seqloss_bwd(const float* gmean, const double* result, const void* pred, const void* gt, const float* valid,
                    void* gpred, int64_t N, int64_t hs, int64_t ws, int64_t H, int64_t W, float max_flow, int kind,
                    int dtype, int device, void* stream);
/* get_flow_map_metrics, loss/loss.py:6-22: epe = sqrt(sum_c (pred - gt)^2) over flow_valid >= 0.5;
 * result[8] (device, fp64) = {epe mean, share < 0.5 px, < 1 px, < 3 px, < 5 px, min pred[0], max pred[0], valid count} */
int rsm_flow_metrics(const void* gt, const void* pred, const float* valid, double* workspace, double* result, int64_t N,
                     int64_t C, int64_t H, int64_t W, int dtype, int device, void* stream);

/* ---- PFM files, tools/pfm_file_io.py:6-77 (writer call site test_stereo.py:133).  HOST pointers; `image` is
 * H x W x channels fp32 in C order, channels 1 ("Pf") or 3 ("PF"); the scale is written with "%f", negated on a
 * little-endian host as the reference does; flip_rows != 0 stores / returns the rows bottom-up, which is what the
 * reference's np.flipud at the call site (write) and inside read_pfm_file (:44) amount to */
int rsm_pfm_write(const char* path, const float* image, int64_t H, int64_t W, int channels, double scale, int flip_rows);
/* header only: sizes, channels, the scale as stored (negative = little-endian data), byte offset of the data */
int rsm_pfm_read_header(const char* path, int64_t* H, int64_t* W, int* channels, double* scale, int64_t* data_offset);
/* data into a caller-allocated H x W x channels buffer, converted to the host byte order */
int rsm_pfm_read(const char* path, float* image, int64_t H, int64_t W, int channels, int flip_rows);

/* ---- disparity regression over a dense (N,D,H,W) cost: softmax(+cost) expectation
 * (model/mobile_stereo_net.py:144-147, mobile_stereo_net_v4.py:10-14 + :517,
 * mobile_disp_net_c.py:208-220) and hard argmin / argmax (torch.argmin(cost, 1) semantics;
 * not in the reference) in ONE pass over the volume */
int rsm_regress_fwd(const void* cost, int64_t N, int64_t D, int64_t H, int64_t W, int dtype,
                    rsm_regress_out out, int device, void* stream);
/* gcost[d] = gout * p[d] * (d - expect);  expect/lse (fp32) from the forward call; gout (N,H,W) in dtype */
int rsm_regress_bwd(const void* gout, const void* cost, const float* expect, const float* lse,
                    void* gcost, int64_t N, int64_t D, int64_t H, int64_t W, int dtype,
                    int device, void* stream);

/* ---- expectation of given probabilities: disparity_regression(x, maxdisp),
 * model/mobile_stereo_net_v4.py:10-14 (x is already softmax-ed): out (N,H,W) = sum_d d * prob[d] */
int rsm_expect_fwd(const void* prob, void* out, int64_t N, int64_t D, int64_t H, int64_t W, int dtype,
                   int device, void* stream);
/* gprob[d] = gout * d */
int rsm_expect_bwd(const void* gout, void* gprob, int64_t N, int64_t D, int64_t H, int64_t W, int dtype,
                   int device, void* stream);

/* ---- MobileStereoNetV4 head: F.interpolate(cost[:,None], [D,H,W], 'trilinear') -> softmax ->
 * disparity_regression, model/mobile_stereo_net_v4.py:511-518 (training heads :471-506), fused:
 * cost (B,Dc,Hc,Wc) dense -> (B,H,W) without materialising the (B,D,H,W) tensor */
int rsm_upsample_regress_fwd(const void* cost, int64_t B, int64_t Dc, int64_t Hc, int64_t Wc,
                             int64_t D, int64_t H, int64_t W, int dtype, rsm_regress_out out,
                             int device, void* stream);
/* bytes of scratch the backward needs: an upper bound, (B,Dc,H,W) fp32 -- what the general two-stage form writes; the
 * x4 x4 x4 head (D = 4 Dc, H = 4 Hc, W = 4 Wc) uses (B, tiles of 32 x 8 fine pixels, Dc, 40) of it */
int64_t rsm_upsample_regress_bwd_workspace(int64_t B, int64_t Dc, int64_t H, int64_t W);
int rsm_upsample_regress_bwd(const void* gout, const void* cost, const float* expect,
                             const float* lse, void* gcost, void* workspace, int64_t B,
                             int64_t Dc, int64_t Hc, int64_t Wc, int64_t D, int64_t H, int64_t W,
                             int dtype, int device, void* stream);

/* ---- fused build + regress: inner product / correlation volume reduced on chip to the
 * soft-argmax disparity and hard argmin / argmax; the (N,D,H,W) volume never reaches HBM.
 * out.soft is fp32 here (the volume stays in the fp32 accumulator); out.lse optional */
int rsm_inner_regress_fwd(rsm_feat left, rsm_feat right, int64_t N, int64_t C, int64_t H,
                          int64_t W, int64_t D, int reduce, int in_dtype, rsm_regress_out out,
                          int device, void* stream);
/* diagnostic twin: `prof` = 16 zero-initialised uint64 on the device; the row-streaming tcgen05 kernel adds clock64
 * cycles per warp role, summed over CTAs: [0] issuer waiting for operands, [1] issuer waiting for a free TMEM block,
 * [2] issuer total, [3] TMA producer waiting for a free ring slot, [4] producer total, [5] epilogue warps waiting
 * for an accumulator group, [6] epilogue warps total, [7] scanning, [8] fence + arrive, [9] merging + storing */
int rsm_inner_regress_fwd_profile(rsm_feat left, rsm_feat right, int64_t N, int64_t C, int64_t H,
                                  int64_t W, int64_t D, int reduce, int in_dtype, rsm_regress_out out,
                                  int device, void* stream, uint64_t* prof);

#ifdef __cplusplus
}
#endif
#endif /* RSM_B200_H_ */
