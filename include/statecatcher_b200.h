/* statecatcher_b200 — C-ABI of the B200-native LucyRNN + CTC (+RNN-T) training hot path.
 *
 * This is the drop-in boundary.  Every entry point replaces a piece
 * of PyTorch/Triton work that the reference performs inside
 *   /root/reference/lucyrnn.py   LucyRNN.forward            (lines 89-191)
 *   /root/reference/lucyrnn.py   LucyRNNCell.forward        (lines 44-70)
 *   /root/reference/lucyrnn_triton.py  fused_decay_scan     (lines 158-177, retired)
 *   /root/reference/model.py     compute_loss CTC branch    (lines 68-71)
 *   /root/reference/model.py     compute_loss RNN-T branch  (lines 73-105)
 * The reference has no FFI of its own (pure Python); the binding a maintainer adds is the
 * ctypes stub shown in INTEGRATION.md (statecatcher_b200/_lib.py is that stub).
 *
 * Conventions
 *   - all pointers are raw DEVICE pointers unless a parameter says "host";
 *   - nothing is allocated and every call is re-entrant (forward thread and autograd's backward
 *     thread may call concurrently, for one or several devices of a process).  The only state the
 *     library keeps is a set of one-time, per-DEVICE caches published through atomics (SM count,
 *     'max dynamic shared memory already raised for this kernel') and environment switches
 *     (SC_*) latched once, immutable afterwards; sc_ctc_head additionally keeps a per-thread, per-device pool of
 *     timing-disabled CUDA events (created on first use) for its cross-stream ordering;
 *   - kernels are launched on the CUDA runtime's CURRENT device: the caller selects the device
 *     that owns the pointers first (the Python layer does: `torch.cuda.device_of(tensor)`);
 *   - every call takes the cudaStream_t to launch on (as void*), and only enqueues work;
 *   - sizes/strides are int64_t counted in ELEMENTS, not bytes;
 *   - dtype codes: SC_F32 = 0, SC_BF16 = 1 (storage type of activations; all arithmetic
 *     accumulates in fp32);
 *   - return value: 0 = ok, >0 = a cudaError_t from the launch, <0 = SC_E_* argument error.
 *     No C++ exception crosses this boundary.
 */
#ifndef STATECATCHER_B200_H
#define STATECATCHER_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SC_F32 0
#define SC_BF16 1

#define SC_E_BADARG (-1)   /* null pointer / negative size / unsupported flag  */
#define SC_E_ALIGN  (-2)   /* pointer or stride breaks the kernel's alignment   */
#define SC_E_DTYPE  (-3)   /* unknown dtype code                                */
#define SC_E_SHAPE  (-4)   /* shape outside what the kernel family supports     */
#define SC_E_UNSUP  (-5)   /* valid request this build cannot serve             */

/* gate column blocks inside the 5H-wide gate tensor G (r, the dead gate of
 * lucyrnn.py:50/56, is never computed): order follows lucyrnn.py:49 minus r. */
#define SC_GATE_Z 0
#define SC_GATE_K 1
#define SC_GATE_V 2
#define SC_GATE_P 3   /* 'h_pre' chunk */
#define SC_GATE_Q 4   /* decay logits  */

#define SC_SCAN_CKPT 8   /* timesteps between saved S checkpoints of the fused scan */

int sc_version(void);                    /* ABI version, bumps on any signature change */
const char* sc_error_string(int code);   /* static string for an SC_E_* / cudaError_t  */
int sc_build_info(char* buf, int64_t n); /* "sm_100a nvcc 12.9 ..." into a host buffer */

/* ---------------------------------------------------------------- K1: projections ----
 * Replaces nn.Linear input_proj / W_fused / output_proj (lucyrnn.py:15, 23, 85; 113, 116,
 * 186) and their autograd backward.
 * fwd  : Y[M,N]  = A[M,K] . W[N,K]^T + bias[N]           (bias may be NULL)
 * dgrad: dA[M,K] = dY[M,N] . W[N,K]
 * wgrad: dW[N,K] (+)= dY[M,N]^T . A[M,K]   fp32 output; accumulate!=0 adds into dW
 * lda/ldw/ldy are row strides in elements.  in_dtype applies to A, W (and dY), out_dtype
 * to the result; bias and dW are always fp32.  W is fp32 when in_dtype==SC_F32 and bf16
 * when in_dtype==SC_BF16.  impl: 0 = auto, 1 = SIMT fp32-FMA kernel, 2 = tcgen05/TMA. */
int sc_gemm_fwd(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias,
                void* Y, int64_t ldy, int64_t M, int64_t N, int64_t K,
                int in_dtype, int out_dtype, int impl, void* stream);
int sc_gemm_dgrad(const void* dY, int64_t lddy, const void* W, int64_t ldw,
                  void* dA, int64_t ldda, int64_t M, int64_t N, int64_t K,
                  int in_dtype, int out_dtype, int impl, void* stream);
int sc_gemm_wgrad(const void* dY, int64_t lddy, const void* A, int64_t lda,
                  float* dW, int64_t lddw, int64_t M, int64_t N, int64_t K,
                  int in_dtype, int accumulate, int impl, void* stream);
/* bytes of scratch the tcgen05 path wants for the given problem (0 for SIMT). */
int64_t sc_gemm_workspace_bytes(int64_t M, int64_t N, int64_t K);

/* ---------------------------------------------------------------- row-wise helpers ---
 * cast  : dst[i] = (dst_dtype) src[i] over a [rows, cols] matrix with row strides.
 * colsum: out[n] (+)= sum_m X[m,n]   (bias gradients), fp32 out.
 * layernorm fwd/bwd over the last dim H: nn.LayerNorm(H) of lucyrnn.py:17-20 (eps 1e-5);
 * the backward ADDS into dw/db (caller zeroes them). */
int sc_cast(const void* src, int64_t lds, int src_dtype, void* dst, int64_t ldd, int dst_dtype,
            int64_t rows, int64_t cols, void* stream);
/* split: dst[r, c] = bf16(src[r,c]), dst[r, cols + c] = bf16(src[r,c] - dst[r,c]); dst is
 * bf16 [rows, 2*cols] (row stride ldd).  Used to push fp32 weight-space products through
 * the bf16 tensor-core GEMM without losing the low mantissa bits (projection folding). */
int sc_split_bf16(const float* src, int64_t lds, void* dst, int64_t ldd, int64_t rows, int64_t cols,
                  void* stream);
/* fp32 [rows, cols] -> six bf16 blocks out of {hi, mid, lo} (x = hi + mid + lo to 2^-25 |x|) for the tensor-core
 * evaluation of an fp32 product — a.w = a_hi w_hi + a_hi w_mid + a_mid w_hi + a_hi w_lo + a_mid w_mid + a_lo w_hi as ONE
 * bf16 GEMM with fp32 accumulation over a six-fold reduction dimension; what replaces the fp32 cuBLAS GEMMs of nn.Linear
 * (lucyrnn.py:113-116, 186) on the fp32 path.  pattern 0 = [lo hi mid mid hi hi] (left operand), 1 = [hi lo mid hi mid hi]
 * (right operand; products in ascending magnitude, the tensor core truncates when it accumulates).  block_stride = element offset between consecutive blocks in dst: `cols` for blocks side by side
 * (ldd >= 6*cols), `rows*ldd` for blocks stacked vertically. */
int sc_split6_bf16(const float* src, int64_t lds, void* dst, int64_t ldd, int64_t rows, int64_t cols,
                   int pattern, int64_t block_stride, void* stream);

int sc_colsum(const void* X, int64_t ldx, int dtype, float* out, int64_t M, int64_t N,
              int accumulate, void* stream);
int sc_layernorm_fwd(const void* X, int64_t ldx, const float* w, const float* b,
                     void* Y, int64_t ldy, float* mean, float* rstd,
                     int64_t M, int64_t H, int dtype, void* stream);
/* dxsum (may be null): [H] fp32, the column sums of dX are ADDED to it — the bias gradient of the projection that
 * produced X, formed while dX is in registers instead of by a second pass over it (sc_colsum). */
int sc_layernorm_bwd(const void* dY, int64_t lddy, const void* X, int64_t ldx, const float* w,
                     const float* mean, const float* rstd, void* dX, int64_t lddx,
                     float* dw, float* db, float* dxsum, int64_t M, int64_t H, int dtype, void* stream);

/* ---------------------------------------------------------------- K2: fused scan -----
 * The whole recurrent part of one layer for fused_ops=True, layer_norm=False (the
 * configuration model.py:232-245 wires).  Replaces the decay-scan loop
 * (lucyrnn.py:153-158 / lucyrnn_triton.py:158-177) AND the per-timestep cell loop
 * (lucyrnn.py:160-166 -> 44-70) AND the step path (lucyrnn.py:172-184):
 *   d=sigmoid(q); S_t=d_t*S_{t-1}+k_t*v_t;
 *   train_mode=1: S_{-1}=0,  s'_t=d_t*S_t+k_t*v_t, sT = s0 is NOT written (caller aliases)
 *   train_mode=0: S_{-1}=s0, s'_t=S_t,             sT = S_{T-1}
 *   c=tanh(p+s'); zh=sigmoid(z); h_t=(1-zh)*c+zh*h_{t-1}, h_{-1}=h0; Hout[b,t,:]=h_t.
 * G    [B,T,5H] gate pre-activations, block g at columns [g*H,(g+1)*H), row stride ldg.
 * h0,s0[B,H] fp32 carried state; hT,sT [B,H] fp32 outputs (sT may be NULL in train mode).
 * Hout [B,T,H] row stride ldh.  Sckpt [B, ceil(T/SC_SCAN_CKPT), H] fp32: S before the first
 * step of each checkpoint interval (saved for the backward's recompute; may be NULL).
 * H must be a multiple of 8 (bf16) / 4 (fp32) and pointers 16-byte aligned. */
int sc_lucy_scan_fwd(const void* G, int64_t ldg, const float* h0, const float* s0,
                     void* Hout, int64_t ldh, float* hT, float* sT, float* Sckpt,
                     int64_t B, int64_t T, int64_t H, int dtype, int train_mode, void* stream);
/* Reverse-time adjoint.  dHout [B,T,H] is dL/dHout; writes dG
 * [B,T,5H] (same block order) and ADDS the per-column sums of dG into dbias[5H] (fp32,
 * the caller zeroes or pre-loads it).  No gradient is produced for h0/s0 (the carried
 * state is detached between segments, model.py:60-61). */
/* Few-streams variant of sc_lucy_scan_fwd (streaming inference with one or a few live streams):
 * the same outputs computed by a time-parallel chunked scan (both recurrences are affine in their
 * state, lucyrnn.py:153-166 / 172-184), three launches over (stream, channel block, 64-step
 * chunk).  sc_lucy_scan_chunked_work_bytes returns the fp32 workspace it needs, or 0 when the
 * sequential kernel is the better choice for (B, T, H) on this device. */
int64_t sc_lucy_scan_chunked_work_bytes(int64_t B, int64_t T, int64_t H);
int sc_lucy_scan_fwd_chunked(const void* G, int64_t ldg, const float* h0, const float* s0,
                             void* Hout, int64_t ldh, float* hT, float* sT, float* Sckpt,
                             float* work, int64_t B, int64_t T, int64_t H, int dtype, int train_mode,
                             void* stream);
int sc_lucy_scan_bwd(const void* G, int64_t ldg, const void* Hout, int64_t ldh,
                     const float* h0, const float* s0, const float* Sckpt,
                     const void* dHout, int64_t lddh, void* dG, int64_t lddg, float* dbias,
                     int64_t B, int64_t T, int64_t H, int dtype, int train_mode, void* stream);

/* ---------------------------------------------------------------- K2': split scans ---
 * General path (layer_norm=True and/or fused_ops=False and/or decay_mode='prefix_sum'):
 * the S scan and the h scan as separate kernels so LayerNorm / W_h can sit between them.
 * sscan: S scan + second application; A[b,t,:] = addend[b,t,:] + s'_t where addend is the
 *        'h_pre' chunk (fused, lucyrnn.py:54) or u (unfused, lucyrnn.py:62).
 *        decay_mode 0 = learned sigmoid(q) (lucyrnn.py:124), 1 = prefix_sum with
 *        lambda_decay (lucyrnn.py:126-142; training path only).  S_all, fp32, is what the backward needs:
 *        decay_mode 0: [B, ceil(T/SC_SCAN_CKPT), H], S entering each interval (entry 0 = the initial state; the
 *        backward recomputes S_t inside an interval); decay_mode 1: [B,T,H], every S_t.
 * hscan: c=tanh(An); zh=sigmoid(Zn); h_t=(1-zh)c+zh*h_{t-1}.
 * sscan_bwd's dsum (may be null): [3][H] fp32, the column sums of dk, dv, dq are ADDED to it (bias gradients). */
int sc_lucy_sscan_fwd(const void* k, const void* v, const void* q, int64_t ldg,
                      const void* addend, int64_t ldadd, const float* s0,
                      void* A, int64_t lda, float* S_all, float* sT,
                      int64_t B, int64_t T, int64_t H, int dtype, int train_mode,
                      int decay_mode, float lambda_decay, void* stream);
int sc_lucy_sscan_bwd(const void* k, const void* v, const void* q, int64_t ldg,
                      const float* S_all, const float* s0, const void* dA, int64_t ldda,
                      void* dk, void* dv, void* dq, int64_t lddg, float* dsum,
                      int64_t B, int64_t T, int64_t H, int dtype, int train_mode,
                      int decay_mode, float lambda_decay, void* stream);
int sc_lucy_hscan_fwd(const void* An, int64_t ldan, const void* Zn, int64_t ldzn,
                      const float* h0, void* Hout, int64_t ldh, float* hT,
                      int64_t B, int64_t T, int64_t H, int dtype, void* stream);
int sc_lucy_hscan_bwd(const void* An, int64_t ldan, const void* Zn, int64_t ldzn,
                      const void* Hout, int64_t ldh, const float* h0,
                      const void* dHout, int64_t lddh, void* dAn, int64_t lddan,
                      void* dZn, int64_t lddzn,
                      int64_t B, int64_t T, int64_t H, int dtype, void* stream);

/* ---------------------------------------------------------------- K3: CTC ------------
 * Replaces log_softmax + nn.CTCLoss(blank, zero_infinity=True) forward and backward
 * (model.py:70-71, train.py:142; ATen ctc_loss).  logits[b,t,:] at
 * logits + b*stride_b + t*stride_t (V contiguous) — so both the (B,T,V) encoder output and
 * the (T,B,V) transposed view nn.CTCLoss receives are accepted.  Rows need not be
 * normalised (log-softmax is folded in and is idempotent).
 * targets [B,Umax] int64 (row stride ldt), in_lens/tgt_lens [B] int64.
 * Lengths are validated on the device (they may be device tensors the host never sees): T_b is clamped
 * to T; an utterance with U_b < 0 or U_b > Umax, or with a label outside [0, V), is reported infeasible
 * (nll = +inf, zero gradient row) instead of reading out of bounds.
 * Workspaces (caller-allocated, 16-byte aligned): lse [B,T] fp32, cshift [B,T] fp32 (per-frame shift
 * taken out of the lattice emissions: the largest of them, log2 units), alpha/beta [B,T,S] 4-byte
 * cells with S = 2*Umax+1 rounded up to a multiple of 8 (whole 32-byte sectors), lplat [B,T,sc_ctc_lplat_pitch(Umax)] 4-byte cells
 * (>= S; wider for narrow lattices, whose emission rows keep the pitch of the recursion's shared-memory
 * rows so that a block of them is one bulk copy), and the opaque `ws` of
 * sc_ctc_workspace_bytes(B, T, Umax) bytes (per-utterance format flags, per-direction likelihoods and range
 * records).  The CONTENT of lplat/alpha/beta is private to the three passes and depends on Umax:
 *   Umax <= 255: fp64 linear-domain recursion (csrc/sc_ctc_lin64.cuh) — lplat rows hold U+1 emission
 *     probabilities, alpha/beta rows the high words of fp64 node values; utterances whose dynamic range fp64
 *     cannot hold are detected, flagged in ws and recomputed by the log-domain kernel;
 *   larger lattices: log-domain recursion, fp32 log2 values (alpha with its frame's emission, beta without).
 * The same lse / alpha / beta / nll / ws must be handed to sc_ctc_bwd.
 * nll [B] = per-utterance negative log-likelihood (+inf if infeasible);
 * loss [1] = reduction of nll: reduction 0 none (loss untouched), 1 mean
 * = mean_b(nll_b/max(U_b,1)), 2 sum; infeasible utterances contribute 0 (zero_infinity). */
int64_t sc_ctc_workspace_bytes(int64_t B, int64_t T, int64_t Umax);
int64_t sc_ctc_lplat_pitch(int64_t Umax);
int sc_ctc_fwd(const void* logits, int64_t stride_b, int64_t stride_t, int dtype,
               const int64_t* targets, int64_t ldt, const int64_t* in_lens,
               const int64_t* tgt_lens, int64_t B, int64_t T, int64_t V, int64_t Umax,
               int64_t blank, float* lse, float* lplat, float* cshift, float* alpha, float* beta,
               float* nll, float* loss, int reduction, void* ws, void* stream);
/* The two halves of sc_ctc_fwd as separate calls (what statecatcher_b200/ctc.py binds, so that
 * the bandwidth-bound emission pass and the latency-bound lattice recursion are timed apart):
 * sc_ctc_emissions: logits -> lse, lplat, cshift (log-softmax + gather of the lattice emissions);
 * sc_ctc_lattice:   lplat, cshift -> alpha, beta, nll, loss, ws (the alpha/beta recursions + reduction). */
int sc_ctc_emissions(const void* logits, int64_t stride_b, int64_t stride_t, int dtype,
                     const int64_t* targets, int64_t ldt, const int64_t* in_lens,
                     const int64_t* tgt_lens, int64_t B, int64_t T, int64_t V, int64_t Umax,
                     int64_t blank, float* lse, float* lplat, float* cshift, void* stream);
int sc_ctc_lattice(const float* lplat, const float* cshift, const int64_t* targets, int64_t ldt,
                   const int64_t* in_lens, const int64_t* tgt_lens, int64_t B, int64_t T,
                   int64_t Umax, int64_t blank, float* alpha, float* beta, float* nll,
                   float* loss, int reduction, void* ws, void* stream);
/* dlogits[b,t,:] = gout * scale_b * (softmax - occupancy) for t < T_b, exactly 0 for
 * t >= T_b and for infeasible utterances.  grad_out: device fp32, [1] for mean/sum, [B]
 * for reduction none.  dlogits strides like logits. */
int sc_ctc_bwd(const void* logits, int64_t stride_b, int64_t stride_t, int dtype,
               const int64_t* targets, int64_t ldt, const int64_t* in_lens,
               const int64_t* tgt_lens, int64_t B, int64_t T, int64_t V, int64_t Umax,
               int64_t blank, const float* lse, const float* alpha, const float* beta,
               const float* nll, const float* grad_out, int reduction,
               void* dlogits, int64_t dstride_b, int64_t dstride_t, int out_dtype,
               const void* ws, void* stream);

/* The whole head in one call — loss AND the gradient for a unit upstream gradient, with the two V-wide passes
 * running UNDER the latency-bound recursions (model.py:70-71 + the backward autograd would run later).
 * The segment is cut into sc_ctc_head_phases(T, Umax, phases) chunks of whole 64-frame emission blocks; the
 * recursions run as that many launches over frame ranges on `stream_l` (the column travels through `ws` in fp64:
 * rows are bit-identical to sc_ctc_lattice), the emission pass feeds them chunk by chunk from `stream_e` (both
 * ends inwards) and the gradient pass follows on `stream_g` (the middle outwards); ordering is by events, the side
 * streams fork from and join `stream` (safe inside a stream capture).  stream_l should have a higher priority
 * than stream_e / stream_g (the recursion blocks must find room on SMs the V-wide kernels fill); null: `stream`.  dlogits = scale_b * (softmax - occupancy)
 * as sc_ctc_bwd with grad_out = 1; the caller multiplies by the upstream gradient when autograd delivers it
 * (sc_ctc_scale_grad).  reduction 1 (mean) or 2 (sum) only.  phases <= 0: the library chooses.  stream_e /
 * stream_g null or equal to `stream`, or a single phase: the same passes one after the other on `stream`.
 * Workspaces and outputs as sc_ctc_fwd / sc_ctc_bwd. */
int64_t sc_ctc_head_phases(int64_t T, int64_t Umax, int64_t phases);
int sc_ctc_head(const void* logits, int64_t stride_b, int64_t stride_t, int dtype,
                const int64_t* targets, int64_t ldt, const int64_t* in_lens,
                const int64_t* tgt_lens, int64_t B, int64_t T, int64_t V, int64_t Umax,
                int64_t blank, float* lse, float* lplat, float* cshift, float* alpha, float* beta,
                float* nll, float* loss, int reduction, void* ws,
                void* dlogits, int64_t dstride_b, int64_t dstride_t, int out_dtype,
                int64_t phases, void* stream, void* stream_l, void* stream_e, void* stream_g);
/* x[0..n) *= *scale (device fp32 scalar); returns without touching x when *scale == 1. */
int sc_ctc_scale_grad(void* x, int dtype, int64_t n, const float* scale, void* stream);

/* ---------------------------------------------------------------- K4: RNN-T ----------
 * Replaces warp_rnnt.RNNTLoss as called at model.py:97-105 (gather=True): transducer
 * alpha/beta over the T x (U+1) lattice swept by anti-diagonals.
 * log_probs [B,T,U1,V] normalised fp32, contiguous (U1 = Umax+1 <= 1024), labels [B,Umax]
 * int64 (row stride ldl), frame_lens/label_lens [B] int64.
 * Workspaces fp32, 16-byte aligned, SKEWED layout [B, T+U1, U1p] with U1p = U1 rounded up to a
 * multiple of 4 and row d = t+u: eb (blank log-probs), el (label log-probs), alpha, beta.
 * row_offsets: NULL for the padded layout; else [B] int64 first-row index of each utterance in
 * the COMPACT packing of model.py:147-200 (log_probs is then [total_rows, V], utterance b
 * holding T_b x (U_b+1) rows in (t,u) order; T and U1 are still the batch maxima).
 * fwd: fills the workspaces and nll[B] (0 for utterances with no frames).
 * bwd: grad (same layout as log_probs) := d(sum_b grad_w[b]*nll_b)/dlog_probs — fully written
 * (zeros except the blank and label entry of each live node). */
int sc_rnnt_fwd(const float* log_probs, const int64_t* labels, int64_t ldl,
                const int64_t* frame_lens, const int64_t* label_lens,
                int64_t B, int64_t T, int64_t U1, int64_t V, int64_t blank,
                const int64_t* row_offsets,
                float* eb, float* el, float* alpha, float* beta, float* nll, void* stream);
int sc_rnnt_bwd(const int64_t* labels, int64_t ldl, const int64_t* frame_lens,
                const int64_t* label_lens, int64_t B, int64_t T, int64_t U1, int64_t V,
                int64_t blank, const int64_t* row_offsets, int64_t total_rows,
                const float* eb, const float* el, const float* alpha,
                const float* beta, const float* nll, const float* grad_w, float* grad,
                void* stream);

/* ---------------------------------------------------------------- K4': fused joint head
 * Chunked joint -> loss path that never materialises the (B,T,U+1,V) logits of
 * model.py:136-144 as a whole (host logic in rnnt.py RNNTFusedHead).
 * joint_fwd : out[b,t,u,:] = tanh(enc[b,t,:] + pred[b,u,:])  for a block of Tc frames
 *             (enc/pred strides in elements: batch, then frame / label position).
 * joint_bwd : d_pre = dJ*(1-joint^2) (joint recomputed); d_enc[b,t,:] = sum_u d_pre (written),
 *             d_pred[b,u,:] += sum_t d_pre (fp32 [B,U1,J], accumulated across blocks).
 * rnnt_lse_gather: logits [B,Tc,U1,V] of frames [t0,t0+Tc) -> lse [B,T,U1] and the skewed
 *             eb/el entries of those nodes.
 * rnnt_lattice   : alpha/beta/nll from filled eb/el (same kernel as sc_rnnt_fwd).
 * rnnt_node_grads: gb/gl [B,T,U1] = d(sum_b grad_w[b]*nll_b)/d(log-prob of blank / label).
 * rnnt_dlogits   : dlogits[b,tc,u,v] = gb*([v==blank]-p_v) + gl*([v==label]-p_v). */
int sc_joint_fwd(const void* enc, int64_t enc_sb, int64_t enc_st, const void* pred, int64_t pred_sb,
                 int64_t pred_su, void* out, int64_t B, int64_t Tc, int64_t U1, int64_t J, int dtype,
                 void* stream);
int sc_joint_bwd(const void* dJ, const void* enc, int64_t enc_sb, int64_t enc_st, const void* pred,
                 int64_t pred_sb, int64_t pred_su, void* d_enc, int64_t denc_sb, int64_t denc_st,
                 float* d_pred, int64_t B, int64_t Tc, int64_t U1, int64_t J, int dtype, void* stream);
int sc_rnnt_lse_gather(const void* logits, int dtype, const int64_t* labels, int64_t ldl,
                       const int64_t* frame_lens, const int64_t* label_lens, int64_t B, int64_t T,
                       int64_t t0, int64_t Tc, int64_t U1, int64_t V, int64_t blank, float* lse,
                       float* eb, float* el, void* stream);
int sc_rnnt_lattice(const int64_t* frame_lens, const int64_t* label_lens, int64_t B, int64_t T, int64_t U1,
                    const float* eb, const float* el, float* alpha, float* beta, float* nll, void* stream);
int sc_rnnt_node_grads(const int64_t* frame_lens, const int64_t* label_lens, int64_t B, int64_t T,
                       int64_t U1, const float* eb, const float* el, const float* alpha, const float* beta,
                       const float* nll, const float* grad_w, float* gb, float* gl, void* stream);
int sc_rnnt_dlogits(const void* logits, int dtype, const float* lse, const float* gb, const float* gl,
                    const int64_t* labels, int64_t ldl, const int64_t* label_lens, int64_t B, int64_t T,
                    int64_t t0, int64_t Tc, int64_t U1, int64_t V, int64_t blank, void* dlogits,
                    float* dbias /* nullable: [V] fp32, += column sums of dlogits */, void* stream);

/* ---------------------------------------------------------------- greedy CTC decode ---
 * Replaces decoder.py:3-30 (argmax + Python collapse loop with one .item() sync per token).
 * logits [B,T,V] (strides like sc_ctc_fwd; log-probs or raw logits — argmax is the same),
 * in_lens [B] int64.  pred [B,T] int32 workspace; out_tokens [B,T] int64 receives each
 * utterance's collapsed label sequence (blanks and repeats removed) left-aligned, out_lens [B]
 * its length.  Argmax ties resolve to the lowest index. */
int sc_ctc_greedy_decode(const void* logits, int64_t stride_b, int64_t stride_t, int dtype,
                         const int64_t* in_lens, int64_t B, int64_t T, int64_t V, int64_t blank,
                         int* pred, int64_t* out_tokens, int64_t* out_lens, void* stream);

/* ---------------------------------------------------------------- fused clip + Adam(W)
 * Replaces clip_grad_norm_ (train.py:553), the per-parameter .item() grad-norm loop
 * (train.py:555-560) and optimizer.step() for Adam/AdamW (train.py:112-137, 563-566).
 * sumsq_accum: *acc (device double, caller zeroes it) += sum(g^2) over one fp32 tensor.
 * scale_grads: g *= min(1, max_norm/(sqrt(*sumsq)+1e-6))           (standalone clip).
 * adam_step  : one Adam (decoupled=0, weight decay as L2) / AdamW (decoupled=1) update of a
 *              contiguous fp32 tensor; step is the 1-based step count (bias correction);
 *              sumsq may be NULL (no clipping), else the clip coefficient above is applied to
 *              the gradient on the fly (the gradient buffer itself is left untouched). */
int sc_sumsq_accum(const float* g, int64_t n, double* acc, void* stream);
int sc_scale_grads(float* g, int64_t n, const double* sumsq, float max_norm, void* stream);
int sc_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1,
                 float beta2, float eps, float weight_decay, int64_t step, const double* sumsq,
                 float max_norm, int decoupled, void* stream);
/* lion_step  : one Lion update (train.py:125-131 constructs `lion_pytorch.Lion`, absent upstream and
 *              unpinned; the published rule): p *= 1 - lr*wd; p -= lr*sign(beta1*m + (1-beta1)*g);
 *              m = beta2*m + (1-beta2)*g.  sumsq / max_norm as in adam_step. */
int sc_lion_step(float* p, const float* g, float* m, int64_t n, float lr, float beta1, float beta2,
                 float weight_decay, const double* sumsq, float max_norm, void* stream);
/* Multi-tensor forms of the three calls above: ONE launch per 32 tensors instead of one per tensor
 * (the 26 parameter tensors of the 6-layer model: 3 launches per optimizer step instead of 52).
 * p / g / m / v / n are HOST arrays of `count` DEVICE pointers / element counts (the only host-array
 * arguments of this ABI: the table is passed in the kernel parameters, so it follows gradients that
 * zero_grad(set_to_none=True) re-allocates every step); tensors with n == 0 are skipped; all tensors
 * of a call share the hyper-parameters and the step count.  Same arithmetic per element as the
 * single-tensor calls (bit-identical parameter updates). */
int sc_sumsq_accum_multi(const float* const* g, const int64_t* n, int64_t count, double* acc, void* stream);
int sc_adam_step_multi(float* const* p, const float* const* g, float* const* m, float* const* v,
                       const int64_t* n, int64_t count, float lr, float beta1, float beta2, float eps,
                       float weight_decay, int64_t step, const double* sumsq, float max_norm, int decoupled,
                       void* stream);
int sc_lion_step_multi(float* const* p, const float* const* g, float* const* m, const int64_t* n,
                       int64_t count, float lr, float beta1, float beta2, float weight_decay,
                       const double* sumsq, float max_norm, void* stream);
/* Data-parallel gradient exchange (reference: none — train.py is single-process): pack the fp32
 * gradients of a bucket into their slices of ONE flat communication buffer (flat_dtype SC_F32, or SC_BF16 = half
 * the NVLink payload of the all-reduce) and, after the collective, the reverse with the averaging factor
 * (g = scale * slice).  g, flat_slices and n are HOST arrays of device pointers / element counts (one launch per
 * 32 tensors, like the optimizer's multi-tensor calls); a slice must be aligned to its element size. */
int sc_grads_pack_multi(const float* const* g, void* const* flat_slices, const int64_t* n, int64_t count,
                        int flat_dtype, void* stream);
int sc_grads_unpack_multi(float* const* g, const void* const* flat_slices, const int64_t* n, int64_t count,
                          int flat_dtype, float scale, void* stream);


/* ---------------------------------------------------------------- frontend ------------
 * Replaces make_frontend (model.py:250-279): torchaudio MFCC(n_mfcc=80, dct_type=2, norm='ortho',
 * log_mels=True) / MelSpectrogram + AmplitudeToDB(top_db=80) with n_fft = win_length = 400,
 * hop_length = 160, n_mels = 80, center=False, power=2, mel_scale='htk', followed by
 * train.py:475's transpose: wav [B,S] fp32 (row stride ldw) -> out [B,T,80] fp32 (stream stride
 * out_stride_b >= T*80), T = 1 + (S-400)/160 (nothing is written when S < 400).
 * sc_frontend_tables fills a HOST buffer of sc_frontend_tables_len() floats (window, folded DFT
 * bases, sparse mel bands, DCT; pure C, no CUDA call) which the caller uploads once and passes as
 * `tables` (device).  mode 0 = mfcc; mode 1 = mel dB, gmax = device word receiving the batch
 * maximum, top_db < 0 skips AmplitudeToDB's floor. */
int64_t sc_frontend_tables_len(void);
int sc_frontend_tables(float* host_out, int64_t n, int sample_rate);
int sc_frontend(const float* wav, int64_t ldw, int64_t B, int64_t S, const float* tables, int mode,
                float top_db, const uint8_t* frame_mask /* nullable, mode 0 only: [B,T] bytes, row stride ldmask;
                frames with 0 are written as zeros = model.py:377 folded in */, int64_t ldmask,
                float* out, int64_t out_stride_b, unsigned int* gmax, void* stream);
/* compute_frame_mask (train.py:296-306) + the in_lens line (train.py:490), bit-exact:
 * sample_mask [B,S] bytes (row stride ldm); frame_mask[b,t] = any(sample_mask[b, t*sub:(t+1)*sub])
 * for t < T; in_lens[b] = (int64) min(float(sum_s sample_mask[b,s]) / subsample, nfeat) with the
 * division in fp32.  Requires T*sub <= S (the caller enforces the reference's S_trim == T*sub). */
/* dst[r,:] = src[r,:] * float(mask[r]) — `feats * mask.unsqueeze(-1).float()` (model.py:377). */
int sc_mask_rows(const void* src, int64_t lds, int dtype, const uint8_t* mask, void* dst, int64_t ldd,
                 int64_t rows, int64_t cols, void* stream);
int sc_frame_mask(const uint8_t* sample_mask, int64_t ldm, int64_t B, int64_t S, int64_t T, int64_t sub,
                  float subsample, int64_t nfeat, uint8_t* frame_mask, int64_t* in_lens, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* STATECATCHER_B200_H */
