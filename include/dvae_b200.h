/* dvae_b200 -- C-ABI of the B200 (sm_100a) Disentangled-VAE hot path.
 *
 * The reference (v-manhlt3/Disentangle-VAE-for-VC) has no FFI layer: its boundary is the Python class API of
 * model/disentangled_vae.py and model/variational_base_vae.py, which reaches the GPU through torch.nn (cuDNN/cuBLAS).
 * This header is what replaces those library calls.  Each entry point names the reference call site(s) it stands in
 * for (paths relative to the reference root).  The Python host side (disentangle-vae-for-vc_b200/dvae_b200/lib.py)
 * binds exactly these symbols with ctypes; INTEGRATION.md shows the stub.
 *
 * Conventions: plain pointers + sizes, no allocation inside, no implicit synchronisation, everything is enqueued on
 * `stream` (a cudaStream_t passed as void*).  Return 0 = ok, 1 = invalid argument, 2 = CUDA error; the message is in
 * dvae_last_error() (thread local).  `dtype` tags the activation storage: 0 = bf16 (tcgen05 kind::f16),
 * 1 = fp32 kept on the tf32 grid (tcgen05 kind::tf32), 2 = IEEE fp16 (kind::f16; tf32's 10 mantissa bits at bf16's
 * byte count and rate -- the caller keeps the activation-gradient stream scaled by a power of two: `scale` / `gscale`
 * arguments put it on the stream, `alpha` arguments take it off the parameter gradients), 3 = strict fp32: fp32 storage,
 * every contraction on the CUDA cores with fp32 FMA accumulation (the mode of north_star's 1e-5 check).  "act" means that
 * storage type.  Activations are channels-last [rows, T, C]; T = 64 (model/disentangled_vae.py:165,235).
 */
#ifndef DVAE_B200_H
#define DVAE_B200_H
#ifdef __cplusplus
extern "C" {
#endif

const char* dvae_last_error(void);
/* bytes of caller-provided scratch: op in {"bn_stats","bn_stat","bn_bwd_coef","loss","lstm_bwd_dc","segment_ids","group_acc"},
 * n0..n2 = its shape parameters (see csrc/host_common.cu); -1 for an unknown name.  dvae_lstm_bwd_workspace sizes the
 * split-K fix-up buffers of the LSTM backward. */
long dvae_workspace_bytes(const char* op, long n0, long n1, long n2);
int dvae_version(void);
int dvae_sm_arch(void);              /* 100: built for sm_100a only */
int dvae_lstm_gate_tile(int H);   /* gate-interleave tile (columns) used by the LSTM forward for hidden size H */
int dvae_debug_seq_stamps(long long* buf);   /* optional per-step SM-clock stamps [T][8] of the sequence-resident LSTM kernels (debug) */
int dvae_lstm_launches(int H, int T, int backward);   /* kernels dvae_lstm_fwd / dvae_lstm_bwd enqueue for this shape */
int dvae_lstm_launches_for(int dtype, int rows, int T, int H, int D, int backward);   /* same, for the exact call (time-resident forward kernel: one launch per 1024 rows) */
int dvae_debug_res_stamps(unsigned long long* buf);   /* optional globaltimer stamps [T][2][8] of CTA 0 of the time-resident LSTM forward kernel (debug) */
int dvae_set_lstm_resident(int on);   /* 1 / 0: time-resident forward kernel for the H = 512 / 1024 recurrences (ops_lstm_res.cu) on / off; < 0 queries; returns the previous setting */
int dvae_set_background(int on);   /* GEMMs launched while on: small-footprint kernels that co-run with a latency-critical stream */
int dvae_debug_timing(unsigned long long* buf, int capacity);   /* optional per-CTA phase stamps of the GEMM kernel (debug) */

/* ---- nn.Linear (model/disentangled_vae.py:98-100 LinearNorm.forward; :165-171, :194, :211-213, :232-233, :247) */
int dvae_linear_fwd(int dtype, const void* x, long ldx, const void* w, const float* bias, void* out, float* out_f32,
                    long ldo, int M, int N, int K, int relu, int block_n, void* stream);
/* split-precision variants for the small layers (M = rows): w = [N, 2K] = [w_hi | w_lo] when b_terms = 2; out_cat (may be
 * null) = act [M, 3N] = [hi | lo | hi] of the result, the operand of a following GEMM against [w_hi | w_hi | w_lo] */
int dvae_linear_fwd_split(int dtype, const void* x, long ldx, const void* w, const float* bias, void* out, float* out_f32,
                          long ldo, void* out_cat, int M, int N, int K, int relu, int b_terms, void* stream);
int dvae_linear_dgrad(int dtype, const void* dy, long lddy, const void* w, void* dx, float* dx_f32, const void* relu_mask,
                      long ldx, int M, int N, int K, int block_n, void* stream);
int dvae_linear_wgrad(int dtype, const void* dy, long lddy, const void* x, long ldx, float* dw, long lddw, int M, int N,
                      int K, float alpha, void* stream);

/* ---- nn.Conv1d k=5 pad=2 (ConvNorm.forward :119-121; enc :154-160/:201-202, dec :178-189/:242-243, postnet :54-78/:81-87)
 *      x [R,T,Cin] act, wk [Cout,5,Cin] act (dvae_prep_conv_weight), y [R,T,Cout] */
int dvae_conv5_fwd(int dtype, const void* x, const void* wk, const float* bias, void* y, float* y_f32, int R, int T, int Cin,
                   int Cout, void* stream);
/* conv + the statistics pass of the train-mode BatchNorm1d behind it (:154-160, :178-189, :54-78): bn_ws [halves*2*Cout + 1]
 * doubles = per-half column sums / sums of squares of y; consumed by dvae_bn_finalize_apply.
 * y_f32 != 0 (fp16 / tf32 activations): y is stored as UNROUNDED fp32 -- the rounding of the pre-BatchNorm tensor is the
 * largest single contribution to the forward error; the BatchNorm entry points take a flag for the fp16 case (y wider
 * than the activations) */
int dvae_conv5_fwd_bnstats(int dtype, const void* x, const void* wk, const float* bias, void* y, int y_f32, int R, int T, int Cin,
                           int Cout, double* bn_ws, int rows_half, int halves, void* stream);
int dvae_conv5_dgrad(int dtype, const void* dy, const void* wk, void* dx, float* dx_f32, int R, int T, int Cin, int Cout,
                     void* stream);
int dvae_conv5_wgrad(int dtype, const void* dy, const void* x, float* dwk, int R, int T, int Cin, int Cout, float alpha,
                     void* stream);

/* ---- nn.LSTM recurrence (:163/:208 enc_lstm, :172/:238 dec_lstm1, :193/:246 dec_lstm2); the input projection is a
 *      dvae_linear_fwd with gate-interleaved weights.  xg [rows,T,D*4H] in: projection, out: activated gates */
int dvae_lstm_fwd(int dtype, void* xg, const void* whh_p, void* h_all, float* c_all, int rows, int T, int H, int D,
                  void* stream);
int dvae_lstm_bwd(int dtype, const void* dh_all, const void* gates, const float* c_all, const void* whh_n, void* da_all,
                  float* dc_ws, float* splitk_ws, int* tickets, int rows, int T, int H, int D, void* stream);
int dvae_lstm_bwd_workspace(int dtype, int rows, int H, int D, long* ws_floats, int* num_tickets);
int dvae_lstm_wgrad_hh(int dtype, const void* da_all, const void* h_all, float* dwhh, int rows, int T, int H, int D,
                       float alpha, void* stream);

/* ---- parameter re-layout (replaces cuDNN's internal filter transforms / flatten_parameters :206) */
int dvae_prep_cast(int dtype, const float* src, void* dst, long n, float scale, void* stream);   /* dst = act(scale * src) */
int dvae_copy_f32(const float* src, float* dst, long n, void* stream);
int dvae_add_inplace(int dtype, void* a, const void* b, long n, void* stream);
int dvae_add_f32_act(int dtype, const float* a, const void* b, float* out, long n, void* stream);  /* residual, channels-last */
int dvae_prep_conv_weight(int dtype, const float* w, void* wk, void* wk_cat, int Co, int Ci, void* stream);   /* wk_cat (may be null): [Co,5,3Ci] = [hi | hi | lo] */
int dvae_prep_cast_split(int dtype, const float* src, void* dst, long rows, long K, int parts, void* stream);   /* [hi | lo] or [hi | hi | lo] per row */
/* every tensor-core copy of the parameters in ONE launch (engine.PreparedWeights.refresh): descs = device array of 56-byte
 * descriptors {const float* src, src2; void* dst0, dst1; int64 n; int32 kind, d0, d1, pad} (kinds: csrc/ops_pointwise.cu PrepKind),
 * blk_desc / blk_off = per block the descriptor index and the first source element of its chunk of `chunk` elements */
int dvae_prep_all(int dtype, const void* descs, const int* blk_desc, const long* blk_off, int num_blocks, int chunk, void* stream);
int dvae_conv_wgrad_unpack(const float* dwk, float* dw, int Co, int Ci, void* stream);
int dvae_prep_lstm_weight(int dtype, const float* w, void* dst, int H, int In, int tile, void* stream);
int dvae_prep_lstm_bias(const float* b_ih, const float* b_hh, float* dst, int H, int tile, void* stream);

/* ---- layout: NCL fp32 <-> channels-last act (the transposes at :204, :240, :244, :248), residual output :277-278 */
int dvae_pack_ncl_to_cl(int dtype, const float* x, void* y, void* y_cat, int R, int C, int T, void* stream);   /* y_cat (may be null): [R,T,3C] = [hi | lo | hi] */
int dvae_unpack_cl_to_ncl(int dtype, const void* a, int a_is_f32, const void* b, float* out_a, float* out_sum, int R, int C,
                          int T, void* stream);
int dvae_recon_out_bwd(int dtype, const float* g_rec, const float* g_hat, void* d_rec, void* d_post, int R, int C, int T,
                       float scale, void* stream);

/* ---- conversion front / back end for a batch of utterances of different lengths: chunking_mel
 *      (model/variational_base_vae.py:335-348; utterance u = fp32 [C, t_len[u]] at mel + mel_off[u], chunks
 *      [chunk_first[u], chunk_first[u+1]), chunk_utt[k] = utterance of chunk k) and the time-concat + residual + clamp of
 *      :288-296 (out_u = fp32 [C, n_u * T] at out + out_off[u]) */
int dvae_chunk_mel(int dtype, const float* mel, const long* mel_off, const int* t_len, const int* chunk_first,
                   const int* chunk_utt, void* x_cl, int n_chunks, int C, int T, void* stream);
int dvae_unchunk_mel(int dtype, const float* a, const void* b, const long* out_off, const int* chunk_first, const int* chunk_utt,
                     float* out, int n_chunks, int C, int T, int clamp, float lo, float hi, void* stream);

/* ---- nn.BatchNorm1d + activation (:159 / :182,:189 / :58,:69,:78 with F.relu :202,:243 and torch.tanh :83) */
int dvae_bn_train_fwd(int dtype, const void* y, int y_f32, void* out, const float* gamma, const float* beta, float* run_mean,
                      float* run_var, long long* num_batches, double* ws, float* stat, int rows_half, int halves, int C,
                      int act, float eps, float momentum, void* stream);
int dvae_bn_finalize_apply(int dtype, const void* y, int y_f32, void* out, const float* gamma, const float* beta, float* run_mean,
                           float* run_var, long long* num_batches, const double* ws, float* stat, int rows_half, int halves,
                           int C, int act, float eps, float momentum, void* stream);
int dvae_bn_eval_fwd(int dtype, const void* y, void* out, const float* gamma, const float* beta, const float* run_mean,
                     const float* run_var, float* stat, long rows, int C, int act, float eps, void* stream);
int dvae_bn_train_bwd(int dtype, const void* dout, const void* y, int y_f32, const float* stat, double* ws, float* coef, void* dy,
                      float* dgamma, float* dbeta, int rows_half, int halves, int C, int act, float alpha, void* stream);
int dvae_colsum(int dtype, const void* x, float* out, long rows, int C, long ldx, float alpha, void* stream);

/* ---- latent tail: _reparameterize :222-228, pair-mean style posterior with detach :257-261, concatenations :263-272 */
int dvae_latent_tail_fwd(int dtype, const float* heads, const float* eps_c1, const float* eps_c2, const float* eps_s,
                         void* z, float* q1_mu, float* q1_lv, float* q2_mu, float* q2_lv, float* zs_mu, float* zs_lv, int R,
                         int L, int S, int sample_content, void* stream);
int dvae_latent_tail_bwd(int dtype, const float* heads, const float* eps_c1, const float* eps_c2, const float* eps_s,
                         const float* dz, const float* dq1_mu, const float* dq1_lv, const float* dq2_mu, const float* dq2_lv,
                         const float* dzs_mu, const float* dzs_lv, void* dheads, int R, int L, int S, int sample_content,
                         float gscale, void* stream);

/* ---- ConvolutionalMulVAE.loss_functionGVAE2 :310-327 (4 x L1-sum/batch_size, 2 x KL, style KL, total) */
int dvae_loss_fwd(const float* x1, const float* x2, const float* r1, const float* r2, const float* h1, const float* h2, long n,
                  const float* q1_mu, const float* q1_lv, const float* q2_mu, const float* q2_lv, int q_rows, int L,
                  const float* s_mu, const float* s_lv, int S, float batch_size, float mse_cof, float kl_cof, void* ws,
                  float* out, void* stream);
int dvae_loss_bwd(const float* x1, const float* x2, const float* r1, const float* r2, const float* h1, const float* h2, long n,
                  const float* q1_mu, const float* q1_lv, const float* q2_mu, const float* q2_lv, int q_rows, int L,
                  const float* s_mu, const float* s_lv, int S, float batch_size, float mse_cof, float kl_cof,
                  const float* gout, float* dr1, float* dr2, float* dh1, float* dh2, float* dq1_mu, float* dq1_lv,
                  float* dq2_mu, float* dq2_lv, float* ds_mu, float* ds_lv, void* stream);

/* ---- speaker-group ops keyed by a group id per row: model/utils.py:13-75 accumulate_group_evidence (mode 0),
 *      model/variational_base_vae.py:281-282 chunk mean (mode 1), raw segmented sums (mode 2),
 *      model/utils.py:95-116 group_wise_reparameterize */
int dvae_segment_ids_sorted(const long long* labels, int* gid, int* scratch, int* num_groups, long B, void* stream);
int dvae_group_accumulate(int mode, const float* a, const float* b, const int* gid, float* acc, float* cnt, long B, int D,
                          void* stream);
int dvae_group_finalize(int mode, const float* acc, const float* cnt, const int* gid, float* table, float* out_a,
                        float* out_b, long B, long G, int D, void* stream);
int dvae_group_pog_bwd(const float* mu, const float* logvar, const int* gid, const float* acc_f, const float* acc_g,
                       float* dmu, float* dlv, long B, int D, void* stream);
int dvae_group_reparam(const float* mu, const float* logvar, const int* gid, const float* eps_group, float* z, long B, int D,
                       void* stream);

/* ---- optimizer: optim.Adam(...) (model/disentangled_vae.py:304) stepped at model/variational_base_vae.py:69.  One launch
 *      for all tensors; every pointer is DEVICE memory: arrays of tensor base pointers / sizes, and for each block the
 *      tensor index and element offset of its chunk.  `step` counts from 1 (bias corrections). */
/* found_inf (device int, may be null): set to 1 when a gradient element is not finite; such elements are left untouched */
int dvae_adam_step(float* const* params, const float* const* grads, float* const* exp_avg, float* const* exp_avg_sq,
                   const long* sizes, const int* blk_tensor, const long* blk_off, int num_blocks, int chunk, double lr,
                   double beta1, double beta2, double eps, long step, int* found_inf, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DVAE_B200_H */
