/* b200rt -- C ABI of the in-box B200 runtime behind the `modal` shim's .map() fan-out.
 *
 * The reference (modal-labs/modal-examples) has no native interface for this path: its fan-out is
 *   Function.map  -> cloudpickle -> gRPC -> one container per input -> gather
 * as used at 06_gpu_and_ml/embeddings/text_embeddings_inference.py:167 (`model.embed.map(...)`), and its
 * arithmetic is an HTTP call into an un-vendored TEI server (`POST /embed`, same file :100).  The entry
 * points below are what a binding for that path replaces; each cites the reference interface it stands
 * in for.  Plain pointers and sizes only; every function is thread-safe (forwards on one replica execute one
 * after the other on the device whichever stream they were enqueued on); nothing calls back into the caller.  Return 0 on success, a negative B200RT_E_* code on failure (text via b200rt_last_error()).
 */
#ifndef B200RT_H
#define B200RT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200RT_OK 0
#define B200RT_TIMEOUT 1        /* b200rt_wait: not complete within timeout_ms; b200rt_poll_any: none ready */
#define B200RT_E_INVALID (-1)   /* bad argument (shape, id out of vocabulary, unknown handle) */
#define B200RT_E_STATE (-2)     /* not initialised / already initialised / shut down */
#define B200RT_E_CUDA (-3)      /* CUDA error; the context is poisoned and every later call fails fast */
#define B200RT_E_NOMEM (-4)
#define B200RT_E_UNSUPPORTED (-5) /* e.g. device is not sm_100, or geometry is not BERT-base */

/* Geometry of a BERT encoder (HF BertConfig fields).  The sm_100a kernels are specialised for the
 * BGE-base geometry (hidden 768, 12 heads of 64, inter 3072, max_pos <= 512); `layers` is free.   */
typedef struct b200rt_bert_config {
    int32_t vocab, hidden, layers, heads, inter, max_pos, type_vocab;
    float eps;
} b200rt_bert_config;

/* Geometry of a CLIP-style ViT image tower (HF CLIPVisionConfig fields + projection_dim).  The kernels are specialised for
 * ViT-B/16: hidden 768, 12 heads of 64, inter 3072, (image/patch)^2 + 1 <= 512 tokens, 3*patch*patch a multiple of 64. */
typedef struct b200rt_vit_config {
    int32_t image, patch, hidden, layers, heads, inter, proj;
    float eps;
} b200rt_vit_config;

typedef struct b200rt_stats_t {
    uint64_t items, waves, tickets;     /* completed so far */
    uint64_t kernel_launches;           /* of this library's own kernels */
    uint64_t h2d_bytes, d2h_bytes;      /* through submit()/wait() */
    uint64_t peer_bytes;                /* scatter + fused-gather bytes that crossed NVLink */
    double stage_us, h2d_scatter_us, forward_us, d2h_us; /* summed per-wave stage times (device events; forward = root replica) */
    double gap_us;                      /* idle time of the first participating replica's compute stream between consecutive waves */
    double dispatch_us;                 /* host time the dispatcher spent forming, staging and enqueueing waves */
    double forward_max_us;              /* summed per wave: the slowest participating replica's forward (device events) */
    double gap_max_us;                  /* summed per wave: the largest idle time any participating replica's compute stream had before it */
} b200rt_stats_t;

/* Replica pool.  Stands in for `@app.cls(gpu=..., max_containers=N)` + `@modal.concurrent`
 * (text_embeddings_inference.py:79-86): n_gpus local B200s instead of N cloud containers.  Enables peer
 * access between all of them and starts the scheduler threads.  devices = NULL means 0..n_gpus-1.    */
/* flags: bits 0-15 = items of 512 tokens one replica takes per wave (0 = default: the SM count of the first device, 148 on a B200), see B200RT_INIT_WAVE_ITEMS */
#define B200RT_INIT_WAVE_ITEMS(n) ((uint32_t)(n) & 0xFFFFu)
/*        bits 16-23 = k > 0: TWO replicas per GPU; each replica's GEMM launches ask for (SMs - k) SMs and its attention launches
 *        for k, so that one replica's attention (MUFU-bound, light on L2) runs beside the other's GEMMs (L2-fill-bound).       */
#define B200RT_INIT_SPLIT_SMS(k) (((uint32_t)(k) & 0xFFu) << 16)
int b200rt_init(int n_gpus, uint32_t flags);
int b200rt_init_devices(const int* devices, int n_gpus, uint32_t flags);
int b200rt_num_gpus(void);

/* Cold start.  Stands in for `download_model` / `spawn_server` (text_embeddings_inference.py:37-56):
 * one host->root-GPU copy of the fp32 weight blob, fp16 conversion on the GPU, then a peer broadcast
 * over NVLink to the other replicas.  kind = "bert" (cfg: b200rt_bert_config; blob order: b200rt.weights.blob_layout) or
 * "vit" (cfg: b200rt_vit_config; the CLIP ViT-B/16 image tower behind 06_gpu_and_ml/embeddings/image_embeddings_infinity.py:76-77;
 * blob order: b200rt.weights.vit_blob_layout).                                                         */
int b200rt_model_load(const char* kind, const void* cfg, const void* weights, size_t nbytes, int* model_out);

/* One .map() input.  Stands in for `TextEmbeddingsInference.embed` -> `POST /embed`
 * (text_embeddings_inference.py:97-104) with token ids instead of strings: ids is [n_items, max_len]
 * int32 row-major (positions >= lens[i] are ignored; lens == NULL means every item is max_len long),
 * out receives [n_items, hidden] fp32 unit-norm embeddings.  Host buffers; `out` stays caller-owned and
 * must remain valid until the ticket completes; ids/lens may be reused as soon as submit returns.     */
int b200rt_submit(int model, const int32_t* ids, const int32_t* lens, int n_items, int max_len, float* out,
                  uint64_t* ticket_out);
/* Zero-copy variant (host wire format).  B200RT_SUBMIT_BORROW_IDS: the caller keeps `ids` valid
 * and unchanged until the ticket completes, and the library does not take a private copy; when `ids` (at the
 * wave's padded length) and/or `out` lie inside a b200rt_alloc_pinned() allocation they are DMA'd from / to where
 * they lie instead of passing through the scheduler's staging buffers.                                          */
#define B200RT_SUBMIT_BORROW_IDS 1u
int b200rt_submit_ex(int model, const int32_t* ids, const int32_t* lens, int n_items, int max_len, float* out,
                     uint32_t flags, uint64_t* ticket_out);
/* One .map() input of the image example (image_embeddings_infinity.py:330-350, `engine.image_embed(images=...)`), with
 * preprocessed pixels instead of image paths: pixels is [n_items, 3, image, image] fp32 (resized / normalised as CLIP's
 * processor does), out receives [n_items, proj] fp32 unit-norm embeddings.  Both stay caller-owned and must remain valid until
 * the ticket completes; each replica pulls its share of the pixels over its own PCIe link (no scatter through the root:
 * an image is 600 KB, not 2 KB).  Pinned memory (b200rt_alloc_pinned) makes the copies asynchronous.  model: kind "vit". */
int b200rt_submit_pixels(int model, const float* pixels, int n_items, float* out, uint64_t* ticket_out);
/* Completion, ordered (`order_outputs=True`): a ticket being waited on is never handed to b200rt_poll_any ...  */
int b200rt_wait(uint64_t ticket, int timeout_ms); /* timeout_ms < 0: forever */
/* ... and unordered (`order_outputs=False`, text_embeddings_inference.py:167): next finished ticket
 * that nobody has waited on yet; B200RT_TIMEOUT when none is ready within timeout_ms.                 */
int b200rt_poll_any(uint64_t* ticket_out, int timeout_ms);

/* Device-resident variant for callers that already hold the batch in HBM on replica `gpu` (index into
 * the pool): enqueues the forward of one batch on `stream` (a cudaStream_t; NULL = the replica's own
 * compute stream) and returns without synchronising.  n_items * max_len must fit the wave capacity.  */
int b200rt_embed_device(int model, int gpu, const int32_t* d_ids, const int32_t* d_lens, int n_items, int max_len,
                        float* d_out, void* stream);
int b200rt_device_sync(int gpu);
int b200rt_wave_capacity_items(void); /* items of 512 tokens one replica takes per wave */

void* b200rt_alloc_pinned(size_t nbytes);
void b200rt_free_pinned(void* p);

int b200rt_stats(b200rt_stats_t* out);
const char* b200rt_last_error(void); /* thread-local */
void b200rt_shutdown(void);

#ifdef __cplusplus
}
#endif
#endif /* B200RT_H */
