/* tdm_b200.h — C ABI of libtdm_b200.so, the B200 (sm_100a) hot path of TinyDiffusionModels.
 *
 * The reference (LiamConnell/TinyDiffusionModels) has no FFI / plugin layer: its boundary is
 * the Python surface of src/mnist.py and src/shakespeare.py.  Each entry point
 * below therefore cites the reference *Python* function (file:line under /root/reference) whose
 * device work it replaces; the ctypes binding that calls it lives in
 * tinydiffusionmodels_b200/_lib.py and the reference-side stub is shown in INTEGRATION.md.
 *
 * Conventions (all entry points):
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless named host_*;
 *   - `stream` is a cudaStream_t passed as void*; work is enqueued asynchronously, never
 *     synchronised, never allocates, and is CUDA-graph capturable;
 *   - buffers are borrowed for the duration of the enqueued work, ownership never transfers;
 *   - return value 0 = success, non-zero = error; tdm_last_error() gives a thread-local message;
 *   - there is no CPU fallback: on a machine without an sm_100 GPU every compute call fails.
 */
#ifndef TDM_B200_H
#define TDM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TDM_OK 0
#define TDM_ERR_ARG 1
#define TDM_ERR_CUDA 2
#define TDM_ERR_UNSUPPORTED 3

/* ABI version of this header (bumped on any signature change). */
int tdm_version(void);
/* Thread-local description of the last non-zero return. Never NULL. */
const char* tdm_last_error(void);
/* Number of kernel launches this library has enqueued since load (bench.py's gpu_launches). */
int64_t tdm_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * Diffusion elementwise math
 * ------------------------------------------------------------------------------------------- */

/* q_sample: out[b,i] = sqrt_acp[t[b]] * x0[b,i] + sqrt_om_acp[t[b]] * noise[b,i]
 * Replaces src/mnist.py:36-42 (inner = 784) and src/shakespeare.py:37-44 (inner = L*dim).
 * fp32, bit-exact with the reference's mul/mul/add sequence (no FMA contraction).
 * x0, noise, out: [batch, inner] contiguous fp32; t: [batch] int64 in [0, n_steps);
 * sqrt_acp, sqrt_om_acp: [n_steps] fp32 schedule tables (src/mnist.py:32-33). */
int tdm_q_sample(const float* x0, const float* noise, const int64_t* t, const float* sqrt_acp,
                 const float* sqrt_om_acp, float* out, int64_t batch, int64_t inner, int n_steps,
                 void* stream);

/* q_sample with in-kernel Philox4x32-10 noise (training fast path, src/mnist.py:155-156).
 * Writes the generated N(0,1) noise to noise_out (needed by the loss) and the diffused sample
 * to out.  The noise for element i of sample b is the (i%4)-th normal of the Philox block
 * counter = (i/4, sample_offset + b, stream_id + *stream_id_dev, 0), key = seed (oracle/philox.py).
 * stream_id_dev (nullable DEVICE int64) lets a captured CUDA graph draw fresh noise per replay:
 * the trainer points it at its on-device step counter. */
int tdm_q_sample_philox(const float* x0, const int64_t* t, const float* sqrt_acp,
                        const float* sqrt_om_acp, float* noise_out, float* out, int64_t batch,
                        int64_t inner, int n_steps, uint64_t seed, uint64_t sample_offset,
                        uint32_t stream_id, const int64_t* stream_id_dev, void* stream);

/* Reverse (ancestral) step, src/mnist.py:167-180 and src/shakespeare.py:343-352:
 *   mean = (1/sqrt(alphas[t])) * (x - betas[t]/sqrt_om_acp[t] * eps)
 *   out  = mean                       if t[0] == 0   (the reference branches on t[0] only)
 *        = mean + sqrt(betas[t]) * z  otherwise
 * z != NULL : injected-noise variant (parity tests), bit-exact with the reference op sequence.
 * z == NULL : in-kernel Philox noise, counter = (i/4, sample_offset + b, step_id + t[b], 1):
 *             the timestep itself keys the noise, so one captured launch serves every step.
 * out may alias x.  All tensors [batch, inner] fp32; t [batch] int64. */
int tdm_reverse_step(const float* x, const float* eps, const float* z, const int64_t* t,
                     const float* betas, const float* alphas, const float* sqrt_om_acp, float* out,
                     int64_t batch, int64_t inner, int n_steps, uint64_t seed,
                     uint64_t sample_offset, uint32_t step_id, void* stream);

/* t[b] += delta for b < batch (device-side loop counter of the captured sampling step;
 * replaces the per-step torch.full of src/mnist.py:192). */
int tdm_timestep_advance(int64_t* t, int64_t batch, int64_t delta, void* stream);

/* Standard-normal fill with the library's Philox stream (x_T initialisation, src/mnist.py:190,
 * src/shakespeare.py:382).  counter = (i/4, sample_offset + b, stream_id, 2). */
int tdm_randn_philox(float* out, int64_t batch, int64_t inner, uint64_t seed,
                     uint64_t sample_offset, uint32_t stream_id, void* stream);

/* (clamp(x,-1,1)+1)/2, src/mnist.py:194. n elements fp32, out may alias x. */
int tdm_to_unit_range(const float* x, float* out, int64_t n, void* stream);

/* Input pipeline: replaces the host-side DataLoader transform of src/mnist.py:139-147
 * (torchvision ToTensor = uint8 -> fp32 / 255, then Normalize = (x - mean) / std) for a dataset that is
 * resident on the device as uint8 rows of row_elems pixels (row_elems % 4 == 0).  out[i] = normalised
 * images[index[i]] for i < n; index == NULL means rows 0..n-1.  Bit-identical to the reference transform.
 * Indices must address rows of `images` (not checked on the device). */
int tdm_u8_gather_normalize(const uint8_t* images, const int64_t* index, float* out, int64_t n,
                            int64_t row_elems, float mean, float stdv, void* stream);

/* Output step: replaces torchvision.utils.save_image's pixel work at src/mnist.py:194-199 / 116-119:
 * optional (clamp(x,-1,1)+1)/2 (from_signed != 0), make_grid(nrow, padding, pad_value 0) of n single-channel
 * h x w fp32 images (channel tripled; n == 1 -> the image itself, no border) and the conversion
 * mul(255).add(0.5).clamp(0,255).to(uint8), written as an HWC uint8 array of the shape
 * tdm_image_grid_shape reports (host call, no GPU needed).  PNG encoding stays on the host. */
int tdm_image_grid_shape(int64_t n, int h, int w, int nrow, int padding, int* out_h, int* out_w);
int tdm_image_grid_u8(const float* x, uint8_t* grid_hwc, int64_t n, int h, int w, int nrow, int padding,
                      int from_signed, void* stream);

/* ---------------------------------------------------------------------------------------------
 * MNIST UNet (src/mnist.py:45-87), bf16 tcgen05 implicit-GEMM convolutions, fp32 accumulate
 * ------------------------------------------------------------------------------------------- */

/* Number of fp32 parameters in the flat parameter vector (state_dict order, 181,473). */
int64_t tdm_unet_param_count(void);
/* Bytes of the packed bf16/fp32 weight image produced by tdm_unet_pack_weights. */
int64_t tdm_unet_wpack_bytes(void);
/* Bytes of activation workspace for a batch of `batch` images (forward only / with backward). */
int64_t tdm_unet_workspace_bytes(int64_t batch, int for_backward);

/* Convert the flat fp32 parameter vector (reference state_dict order: rb1.conv1.weight,
 * rb1.conv1.bias, rb1.conv2.weight, ..., out.weight, out.bias; conv weights OIHW) into the
 * kernel-side packed image (bf16 [tap][Cin/8][Cout][8] planes + fp32 biases). */
int tdm_unet_pack_weights(const float* flat_params, void* wpack, void* stream);

/* Same, and additionally registers a HOST copy of the flat parameters (181,473 floats, copied into
 * library-owned memory) for this `wpack` buffer.  While a mirror is registered, tdm_unet_forward /
 * tdm_unet_p_sample pass the per-channel epilogue vectors (conv / skip biases, time-embedding weight
 * and bias, the 1x1 out conv) to the kernels BY VALUE as launch arguments - they are then read through
 * the constant bank instead of shared memory, which the tensor-core operand reads already saturate.
 * The caller guarantees host and device copies hold the same values.  Repacking the buffer with
 * tdm_unet_pack_weights drops the mirror (training updates parameters on the device only);
 * tdm_unet_forget_host_params drops it explicitly (call before freeing `wpack`).  Launch arguments
 * captured into a CUDA graph are frozen: re-capture after loading new weights. */
int tdm_unet_pack_weights_host(const float* flat_params, const float* flat_params_host, void* wpack,
                               void* stream);
int tdm_unet_forget_host_params(const void* wpack);

/* Sampling schedule switch (process-wide): 1 (default; TDM_UNFUSED=1 in the environment starts at 0) runs each 28x28
 * ResidualBlock as ONE kernel whose conv1 output stays in shared memory (csrc/resblock_tc.cuh; needs the host mirror
 * of tdm_unet_pack_weights_host); 0 runs the layer-by-layer kernels, which leave every intermediate (t1, t4, s4) in
 * the workspace for the per-layer parity tests.  Returns the previous setting.  Training always runs layer by layer
 * (the backward pass needs the intermediates). */
int tdm_unet_set_fused(int on);

/* The workspace must be zero-filled once after allocation (guard rows), then may be reused
 * for the same batch size; zero it again before using it with a different batch size. */

/* Test/debug aid: workspace layout for `batch` as 16 int64 written to HOST memory:
 * {tiles28, tiles14, plane_stride28, plane_stride14, off_t1, off_cat, off_p1, off_t2, off_s2,
 *  off_h2, off_t3, off_t4, off_s4, total_bytes, off_h3, GUARD28}.  Activation tensor [C][pos] lives at
 * off + (c/8)*plane_stride + (GUARD + pos)*16 + (c%8)*2 with pos(b,y,x) = b*S + (y+1)*(W+1) + x,
 * S = (W+1)^2, GUARD = GUARD28 = 168 (W=28) or 24 (W=14).  In the sampling path rb3's output stays at 14x14
 * (off_h3) and channels 0..63 of the concat buffer are not materialised; with the fused blocks
 * (tdm_unet_set_fused) t1, t4 and s4 are not materialised either. */
int tdm_unet_debug_layout(int64_t batch, int64_t* host_out16);

/* SimpleUNet.forward(x, t) (src/mnist.py:76-87).  x: [batch,1,28,28] fp32, t: [batch] int64,
 * eps_out: [batch,1,28,28] fp32. */
int tdm_unet_forward(const void* wpack, const float* x, const int64_t* t, float* eps_out,
                     void* workspace, int64_t workspace_bytes, int64_t batch, void* stream);

/* One fused p_sample (src/mnist.py:167-180): UNet forward with the reverse-step update applied
 * in the last convolution's epilogue, so eps never reaches HBM.  z / Philox semantics as in
 * tdm_reverse_step.  x_out may alias x_in: the only kernel that reads neighbours of x (rb1.conv1)
 * has retired before the last kernel overwrites it, and that kernel reads x at its own pixel only. */
int tdm_unet_p_sample(const void* wpack, const float* x_in, const int64_t* t, const float* z,
                      const float* betas, const float* alphas, const float* sqrt_om_acp,
                      float* x_out, void* workspace, int64_t workspace_bytes, int64_t batch,
                      int n_steps, uint64_t seed, uint64_t sample_offset, uint32_t step_id,
                      void* stream);

/* ---------------------------------------------------------------------------------------------
 * Training step (src/mnist.py:153-159): forward that keeps what the backward needs, backward into
 * one flat fp32 gradient vector (state_dict order), fused AdamW over the flat buffers.
 * The workspace must be sized with tdm_unet_workspace_bytes(batch, 1).
 * ------------------------------------------------------------------------------------------- */

/* SimpleUNet.forward in training mode: as tdm_unet_forward, additionally storing ReLU masks and the
 * last block's output in the workspace. */
int tdm_unet_forward_train(const void* wpack, const float* x, const int64_t* t, float* eps_out,
                           void* workspace, int64_t workspace_bytes, int64_t batch, void* stream);

/* loss = mean((eps - noise)^2) (F.mse_loss, src/mnist.py:158) and d loss / d params.
 * x, t: the inputs of the preceding tdm_unet_forward_train; eps: its output; noise: the target.
 * flat_grad [181,473] and loss_out [1] are OVERWRITTEN (zeroed, then accumulated with atomics). */
int tdm_unet_backward(const void* wpack, const float* x, const int64_t* t, const float* noise,
                      const float* eps, float* flat_grad, float* loss_out, void* workspace,
                      int64_t workspace_bytes, int64_t batch, void* stream);

/* torch.optim.AdamW update (src/mnist.py:148,159) over flat fp32 buffers of n elements.
 * grads are multiplied by grad_scale first (1/world_size after a sum all-reduce).
 * step_dev: DEVICE pointer to the 1-based index of this update (kept on the device so a captured
 * CUDA graph can be replayed; advance it with tdm_timestep_advance(step_dev, 1, +1, stream)). */
int tdm_adamw_flat(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                   float lr, float beta1, float beta2, float eps, float weight_decay,
                   float grad_scale, const int64_t* step_dev, void* stream);

/* ---- data-parallel training: gradient exchange fused into the optimizer (csrc/peer.cu, csrc/unet_bwd.cu) ----
 * No counterpart in the reference (src/mnist.py:129-164 trains on one device).  Each rank allocates one buffer
 * of tdm_peer_buffer_bytes(n) with tdm_peer_alloc, exports its 64-byte CUDA-IPC handle (host memory), exchanges
 * handles with the other ranks (e.g. torch.distributed.all_gather) and maps theirs with tdm_peer_import.  The
 * backward of step k (1-based, the value of *step_dev) writes its flat gradient at
 * base + tdm_peer_grad_offset(n, k & 1); tdm_adamw_flat_peer then waits - on the device - until every rank has
 * announced gradient k, sums the `world` gradients over NVLink in rank order and applies AdamW (grad_scale
 * = 1/world gives the mean).  host_peer_bases[r] is rank r's buffer as mapped in THIS process ([rank] = own).
 * All ranks must call it once per step, in step order; at most 8 ranks. */
int64_t tdm_peer_buffer_bytes(int64_t n);
int64_t tdm_peer_grad_offset(int64_t n, int parity);
int tdm_peer_alloc(int64_t bytes, void** out_ptr);
int tdm_peer_free(void* ptr);
int tdm_peer_export(const void* ptr, void* host_handle64);
int tdm_peer_import(const void* host_handle64, void** out_ptr);
int tdm_peer_close(void* ptr);
int tdm_adamw_flat_peer(float* params, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                        float beta2, float eps, float weight_decay, float grad_scale, const int64_t* step_dev,
                        const void* const* host_peer_bases, int world, int rank, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Shakespeare embedding-space sampler (src/shakespeare.py): TinyTransformer denoiser, reverse step,
 * rounding.  Supported shapes: seq_len 64 or 128, width 256 or 2048 (4 heads, FFN 2048, post-LN,
 * ReLU — the nn.TransformerEncoderLayer defaults the reference uses, src/shakespeare.py:108-111).
 * ------------------------------------------------------------------------------------------- */

/* Row-major fp32 weight [n][k] (nn.Linear.weight) -> bf16 planes [k/8][n_padded][8]; rows >= n are
 * zero.  n_padded must be a multiple of 256 for use as a GEMM / rounding operand. */
int tdm_pack_linear(const float* w, int n, int k, int n_padded, void* out_planes, void* stream);

/* Width 256 only: linear1.weight (2048, 256) and linear2.weight (256, 2048), fp32 row-major, in the stage order the
 * fused feed-forward kernel streams them (1 MB each) - these two replace the tdm_pack_linear images in entries 4 and 6 of
 * tdm_text_forward's per-layer pointer table when dim == 256. */
int tdm_pack_ffn_weights(const float* w1, const float* w2, void* out1, void* out2, void* stream);

int64_t tdm_text_workspace_bytes(int64_t batch, int seq_len, int dim);

/* Put x_t (batch, seq_len, dim) fp32 into the workspace as the sampler state and prepare the first
 * GEMM operand x + time_emb(t/1000) (src/shakespeare.py:116-118). */
int tdm_text_load_state(const float* x_rows, const int64_t* t, const float* time_w, const float* time_b,
                        void* workspace, int64_t workspace_bytes, int64_t batch, int seq_len, int dim,
                        void* stream);

/* Copy out of the workspace as (batch, seq_len, dim) fp32: which = 0 the state x, which = 1 the
 * denoiser output of the last tdm_text_forward. */
int tdm_text_read(const void* workspace, int64_t workspace_bytes, int which, float* out_rows,
                  int64_t batch, int seq_len, int dim, void* stream);

/* TinyTransformer.forward (src/shakespeare.py:115-120, eval mode) on the loaded state.
 * host_ptrs: HOST array of DEVICE pointers, 12 per encoder layer in this order —
 *   in_proj_weight*, in_proj_bias, out_proj.weight*, out_proj.bias, linear1.weight*, linear1.bias,
 *   linear2.weight*, linear2.bias, norm1.weight, norm1.bias, norm2.weight, norm2.bias
 * (* = packed with tdm_pack_linear - except linear1 / linear2 at dim == 256, which are the two images written by
 * tdm_pack_ffn_weights; the rest fp32 vectors) — followed by time_emb.weight [dim] and time_emb.bias [dim] (fp32). */
int tdm_text_forward(const void* const* host_ptrs, int depth, void* workspace, int64_t workspace_bytes,
                     const int64_t* t, int64_t batch, int seq_len, int dim, void* stream);

/* One reverse step (src/shakespeare.py:343-352) on the loaded state: forward as above with the
 * update x_{t-1} = (x_t - beta_t/sqrt(1-acp_t) eps)/sqrt(alpha_t) + sqrt(beta_t) z and the next
 * step's time embedding (t-1) fused into the last LayerNorm.  z_rows: injected noise (batch,
 * seq_len, dim) or NULL for in-kernel Philox keyed like tdm_reverse_step (inner = seq_len*dim).
 * After the call the state holds x_{t-1}; advance t with tdm_timestep_advance and call again. */
int tdm_text_p_sample(const void* const* host_ptrs, int depth, void* workspace, int64_t workspace_bytes,
                      const int64_t* t, const float* z_rows, const float* betas, const float* alphas,
                      const float* sqrt_om_acp, int64_t batch, int seq_len, int dim, uint64_t seed,
                      uint64_t sample_offset, uint32_t step_id, void* stream);

int64_t tdm_round_workspace_bytes(int64_t rows, int dim, int64_t vocab);

/* Rounding: out_idx[r] = argmax_v score(r, v) without materialising the (rows, vocab) logits.
 *   learned (cosine=0): score = x_r . W_v + bias_v          (src/shakespeare.py:389-390, 451-454)
 *   cosine  (cosine=1): score = x_r . E_v / (|x_r| |E_v|)   (src/shakespeare.py:398-401, 462-464);
 *                       w_planes must hold the row-normalised embedding matrix, bias = NULL
 *   guided (ar_logits != NULL, rows = batch): score = (1-alpha) ar[r][v]/T + alpha score/T
 *                                                          (src/shakespeare.py:449-467)
 * x_rows (rows, dim) fp32; w_planes from tdm_pack_linear with n_padded = vocab_padded;
 * out_val (nullable) receives the winning score.  Ties go to the lowest index (torch.argmax). */
int tdm_round_argmax(const float* x_rows, int64_t rows, int dim, const void* w_planes, int64_t vocab,
                     int64_t vocab_padded, const float* bias, int cosine, const float* ar_logits,
                     int64_t ar_ld, float alpha, float temperature, int64_t* out_idx, float* out_val,
                     void* workspace, int64_t workspace_bytes, void* stream);

/* LearnedRounding.forward (src/shakespeare.py:93-102): out_logits[r][v] = x_r . W_v + bias_v as fp32, row-major
 * with leading dimension out_ld >= vocab (cosine=1: x_r . E_v / |x_r| with a pre-normalised table, bias = NULL).
 * Same operands and workspace (tdm_round_workspace_bytes) as tdm_round_argmax; the samplers never call this -
 * they fuse the GEMM with the argmax - it exists for callers that want the logits themselves (training losses). */
int tdm_linear_logits(const float* x_rows, int64_t rows, int dim, const void* w_planes, int64_t vocab,
                      int64_t vocab_padded, const float* bias, int cosine, float* out_logits, int64_t out_ld,
                      void* workspace, int64_t workspace_bytes, void* stream);

/* LearnedEmbedding.forward (src/shakespeare.py:71-80): out_rows[i][:] = table[ids[i]][:] (fp32, dim % 4 == 0).
 * An id outside [0, vocab) writes a zero row and sets *bad_flag (device int, nullable) to 1. */
int tdm_embedding_gather(const float* table, int64_t vocab, int dim, const int64_t* ids, int64_t n,
                         float* out_rows, int* bad_flag, void* stream);

/* ---- Shakespeare training step (src/shakespeare.py:221-250): row f2 of the scope table -----------------------
 * All parameters live in ONE flat fp32 buffer; `offsets` is a HOST array of float offsets into it (and into the
 * gradient buffer of the same layout): 12 per encoder layer in tdm_text_forward's order (weights here are the plain
 * fp32 nn.Linear matrices), then time_emb.weight [dim], time_emb.bias [dim], decoder.weight [vocab][dim],
 * decoder.bias [vocab], embeddings.weight [vocab][dim] (offset -1: the embedding table is not trained and is passed
 * as emb_table instead - the reference's use_learned_embeddings=False, src/shakespeare.py:227-228). */
int64_t tdm_text_train_workspace_bytes(int64_t batch, int seq_len, int dim, int depth, int64_t vocab);
int64_t tdm_text_train_wpack_bytes(int dim, int depth, int64_t vocab);

/* bf16 operand forms of every weight matrix (forward and transposed) from the flat parameters; call after every
 * optimiser step. */
int tdm_text_train_pack(const float* flat_params, const int64_t* offsets, int dim, int depth, int64_t vocab,
                        void* wpack, int64_t wpack_bytes, void* stream);

/* One evaluation of the training objective on a batch of token ids (batch, seq_len):
 *   x0 = table[ids]; t ~ U{0..999} (t_in: injected, else Philox); noise ~ N(0,1) (noise_in: injected (batch,
 *   seq_len, dim), else Philox keyed like tdm_q_sample_philox with stream id = *step_dev);
 *   noise_pred = TinyTransformer(q_sample(x0, t, noise), t) with dropout_p applied at the module's five dropout
 *   sites when grads != NULL (train mode; masks are Philox bits keyed (seed, *step_dev, site, element), see
 *   oracle/text_train_oracle.py); losses[0] = mse(noise_pred, noise), losses[1] = cross_entropy(x0 W^T + b, ids),
 *   losses[2] = losses[0] + *rounding_weight_dev * losses[1]   (src/shakespeare.py:226-244).
 * grads != NULL: also the gradient of losses[2] w.r.t. every parameter, written (not accumulated) into grads.
 * grads == NULL: eval mode (no dropout, no backward) - the validation pass (src/shakespeare.py:268-287).
 * The (batch*seq_len, vocab) logits are never materialised. */
int tdm_text_train_step(const float* flat_params, float* grads, const int64_t* offsets, const void* wpack,
                        const float* emb_table, const int64_t* token_ids, const int64_t* t_in, const float* noise_in,
                        const float* sqrt_acp, const float* sqrt_om_acp, void* workspace, int64_t workspace_bytes,
                        int64_t batch, int seq_len, int dim, int depth, int64_t vocab, float dropout_p,
                        const float* rounding_weight_dev, uint64_t seed, uint64_t sample_offset,
                        const int64_t* step_dev, float* losses, void* stream);

/* torch.optim.AdamW's update (src/shakespeare.py:196) over a flat buffer with the learning rate read from the device
 * (the cosine / warm-up schedule of src/shakespeare.py:159-167 changes it every step of a replayed graph).  The betas
 * are doubles because torch rounds beta and 1 - beta to fp32 separately. */
int tdm_adamw_flat_lr(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                      const float* lr_dev, double beta1, double beta2, float eps, float weight_decay, float grad_scale,
                      const int64_t* step_dev, void* stream);

/* Test aid: byte offsets of the training workspace's tensors (25 values, see text_train.cu; the last one is the int32
 * flag an out-of-range token id sets - nn.Embedding would raise IndexError). */
int tdm_text_train_debug_layout(int64_t batch, int seq_len, int dim, int depth, int64_t vocab, int64_t* out);

/* Measurement aid (bench.py roofline): one fused p_sample with CUDA events recorded on `stream`
 * between its nine launches; SYNCHRONISES on the last event and writes the nine per-kernel
 * durations in milliseconds to host_ms9 (order: rb1.conv1, rb1.conv2, avgpool, rb2.conv1,
 * rb2.conv2, rb3.conv1, rb3.conv2, rb4.conv1, rb4.conv2+out+step). Philox noise. */
int tdm_unet_profile_p_sample(const void* wpack, const float* x_in, const int64_t* t,
                              const float* betas, const float* alphas, const float* sqrt_om_acp,
                              float* x_out, void* workspace, int64_t workspace_bytes, int64_t batch,
                              uint64_t seed, float* host_ms9, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TDM_B200_H */
