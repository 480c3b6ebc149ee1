/*
 * svs_b200 - C ABI of the B200 (sm_100a) block-DCT + parity-QIM frame path.
 *
 * This is the drop-in boundary for ONE function of the reference,
 *     proses_frame_qim_dct(frame, mode, delta, bit_payload_segment, ..., num_ac_coeffs_to_use)
 *     /root/reference/config_and_setup.py:106-174,
 * called by the reference at embed_process.py:117-121 (mode 'embed') and
 * extract_process.py:64-68,180 (mode 'extract').  The reference has no native code and no FFI;
 * the entry points below are what a ctypes binding of that function binds (INTEGRATION.md shows
 * the stub).  Plain pointers and sizes only - no torch / numpy types.
 *
 * Conventions
 *   - frames: uint8, `channels` = 3 (interleaved B,G,R as OpenCV delivers them) or 1 (gray),
 *     addressed as base + f*frame_stride + y*row_stride + x*channels (strides in bytes), so
 *     cropped / non-contiguous views (embed_process.py:113) need no copy.  height and width must
 *     be multiples of 8 (the reference's callers crop to that, embed_process.py:94).
 *   - bits are packed MSB-first into bytes, exactly like bytes_ke_bitstream
 *     (config_and_setup.py:22-23): stream bit p lives in byte p/8, mask 0x80 >> (p%8).
 *   - frame f of a batch carries payload bits [f*cap, min((f+1)*cap, total)), cap =
 *     (height/8)*(width/8)*min(num_ac,63): the running index of embed_process.py:115-128.
 *   - `d_` pointers are device memory, `h_` pointers host memory.  The caller owns every buffer.
 *     Device entry points only enqueue work on `stream` (a cudaStream_t, NULL = default stream):
 *     no allocation, no synchronisation, no global state.
 *   - return value: 0 = ok; < 0 = argument error (SVS_ERR_*); > 0 = a cudaError_t.
 *     svs_last_error_string() describes the last failure of the calling thread.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef SVS_B200_H
#define SVS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SVS_OK              0
#define SVS_ERR_SHAPE      -1   /* height/width not positive multiples of 8, bad channels        */
#define SVS_ERR_POINTER    -2   /* a required pointer is NULL                                    */
#define SVS_ERR_ALIGNMENT  -3   /* payload / stego / bits pointer or stride alignment            */
#define SVS_ERR_DELTA      -4   /* delta is NaN/inf, or 0 < delta < 2^-10 (quantiser range)      */
#define SVS_ERR_STRIDE     -5   /* a stride is smaller than the row / frame it must hold         */
#define SVS_ERR_CONTEXT    -6   /* invalid svs_ctx                                               */

#define SVS_MAX_AC 63

/* ABI version of this header (major*100 + minor). */
int svs_version(void);

/* Message for the last error returned to the calling thread ("" if none). */
const char* svs_last_error_string(void);

/* Bits one frame carries: (height/8)*(width/8)*clamp(num_ac,0,63)  (config_and_setup.py:138). */
int64_t svs_capacity_bits(int height, int width, int num_ac);

/* Smallest bits_frame_stride the fast store path of svs_extract_frames accepts:
 * ceil(cap/32)*4 rounded up to 16 bytes. ceil(cap/8) also works (byte-store path). */
int64_t svs_bits_row_bytes(int height, int width, int num_ac);

/*
 * mode == 'extract' for a batch (config_and_setup.py:129-163,173-174).
 * For every 8x8 block in raster order and every flat coefficient index 1..n (row-major, n =
 * min(num_ac,63)): bit = int(round(c / float32(delta))) mod 2 with c the float32 orthonormal
 * 2-D DCT-II coefficient (scipy op order).  Frame f's bits start at d_bits_out +
 * f*bits_frame_stride; ceil(cap/8) bytes are defined per frame, pad bits are 0.
 * delta <= 0 writes all-zero bits (:143-145); num_ac <= 0 writes nothing.
 * Requires bits_frame_stride >= ceil(cap/8).
 */
int svs_extract_frames(const uint8_t* d_frames, int channels, int64_t n_frames,
                       int height, int width, int64_t frame_stride, int64_t row_stride,
                       double delta, int num_ac,
                       uint8_t* d_bits_out, int64_t bits_frame_stride,
                       void* stream);

/*
 * svs_extract_frames fused with the all-gather of a frame-sharded job: besides d_bits_out the
 * kernel stores every row to n_peers (<= 15) further buffers - normally the other ranks' gathered
 * buffers, peer-mapped over NVLink (CUDA IPC / symmetric memory), each pointer already offset to
 * where THIS rank's frames belong.  Plain stores from the extract kernel: no separate collective,
 * no SMs set aside for one.  The rows are complete on a peer once this kernel has finished and
 * the ranks have synchronised (e.g. a symmetric-memory barrier).  Only the packed kernels do this
 * (16-byte friendly input, bits_frame_stride from svs_bits_row_bytes, delta >= 1/16); otherwise
 * SVS_ERR_ALIGNMENT is returned and the caller gathers with NCCL instead.
 */
int svs_extract_frames_scatter(const uint8_t* d_frames, int channels, int64_t n_frames,
                               int height, int width, int64_t frame_stride, int64_t row_stride,
                               double delta, int num_ac,
                               uint8_t* d_bits_out, int64_t bits_frame_stride,
                               uint8_t* const* peer_bits_out, int n_peers,
                               void* stream);

/*
 * The same fused all-gather through ONE NVSwitch multicast (multimem) address: mc_bits_out maps
 * the gathered buffer of every rank of the job at once (CUDA multicast object / torch symmetric
 * memory `multicast_ptr`), already offset to where THIS rank's frames belong.  Every packed word
 * leaves the GPU once (multimem.st) and the switch replicates it to all ranks, this one included:
 * d_bits_local only has to be a valid local alias of the same rows (used for argument checks;
 * the packed kernels do not write it separately).  Same restrictions as the scatter variant.
 */
int svs_extract_frames_multicast(const uint8_t* d_frames, int channels, int64_t n_frames,
                                 int height, int width, int64_t frame_stride, int64_t row_stride,
                                 double delta, int num_ac,
                                 uint8_t* mc_bits_out, uint8_t* d_bits_local, int64_t bits_frame_stride,
                                 void* stream);

/*
 * mode == 'embed' for a batch (config_and_setup.py:129-158,166-172).
 * Payload bit i of the batch is bit (payload_bit_offset + i) of d_payload (MSB-first);
 * payload_total_bits bits are available.  d_payload must be 4-byte aligned, and the buffer must be
 * readable up to the end of the 32-bit word that holds its last bit (the kernels load whole
 * words; bits past payload_bit_offset + payload_total_bits are never used).  Pad a buffer whose
 * byte length is not a multiple of 4 - svs_b200.embed_frames (Python) and svs_embed_frames_host
 * do so themselves.
 * Per block: float32 DCT-II, q = int(round(c/delta)), q' = q - (q mod 2) + bit on the first k
 * coefficients (k = bits left, at most n), c' = float32(q'*delta), DCT-III, clip to [0,255],
 * truncate.  Blocks after the payload end are copied as gray; the block in which it ends is
 * still inverse-transformed (:130-132,166-169).  delta <= 0 or num_ac <= 0 with a non-empty
 * payload: every block makes a DCT->IDCT round trip and 0 bits are embedded (:143-145).
 *
 * Outputs (all optional except d_stego_out):
 *   d_stego_out        stego frames, `stego_channels` = 1 (what the reference function returns)
 *                      or 3 (gray replicated to B,G,R = the caller's cv2.cvtColor(GRAY2BGR),
 *                      embed_process.py:126), at f*stego_frame_stride + y*stego_row_stride;
 *                      rows must be 8-byte aligned.
 *   d_gray_out         contiguous n_frames x height x width gray reference (first return value)
 *   d_bits_embedded_out  per-frame count of embedded bits (third return value)
 *   d_sse_out          per-frame sum of (stego-gray)^2 over all pixels, for PSNR
 *                      (embed_process.py:204-206); the caller must zero it beforehand.
 *                      (gray / SSE of the frames the packed kernels take come from one extra
 *                      streaming launch; the kernel launch count reflects it.)
 */
int svs_embed_frames(const uint8_t* d_frames, int channels, int64_t n_frames,
                     int height, int width, int64_t frame_stride, int64_t row_stride,
                     const uint8_t* d_payload, int64_t payload_bit_offset, int64_t payload_total_bits,
                     double delta, int num_ac,
                     uint8_t* d_stego_out, int stego_channels,
                     int64_t stego_frame_stride, int64_t stego_row_stride,
                     uint8_t* d_gray_out, int64_t* d_bits_embedded_out,
                     unsigned long long* d_sse_out,
                     void* stream);

/*
 * Host-buffer entry points: the same operations on HOST memory.  Frames are streamed through
 * device staging buffers owned by the context in chunks, with the host->device copy of chunk
 * i+1, the kernel of chunk i and the device->host copy of chunk i-1 overlapping on three
 * streams.  Pinned host buffers (cudaHostAlloc / cudaHostRegister / torch pin_memory) give
 * true overlap; pageable buffers work but serialise.  The calls return after all results are
 * in the host buffers.  A context belongs to one device and must not be used concurrently.
 */
typedef struct svs_ctx svs_ctx;

int svs_ctx_create(int device, int64_t staging_bytes_hint, svs_ctx** out);
int svs_ctx_destroy(svs_ctx* ctx);

int svs_extract_frames_host(svs_ctx* ctx, const uint8_t* h_frames, int channels, int64_t n_frames,
                            int height, int width, int64_t frame_stride, int64_t row_stride,
                            double delta, int num_ac,
                            uint8_t* h_bits_out, int64_t bits_frame_stride);

int svs_embed_frames_host(svs_ctx* ctx, const uint8_t* h_frames, int channels, int64_t n_frames,
                          int height, int width, int64_t frame_stride, int64_t row_stride,
                          const uint8_t* h_payload, int64_t payload_bit_offset, int64_t payload_total_bits,
                          double delta, int num_ac,
                          uint8_t* h_stego_out, int stego_channels,
                          uint8_t* h_gray_out, int64_t* h_bits_embedded_out,
                          unsigned long long* h_sse_out);

/* Number of kernels this library has launched in the calling process (for bench accounting). */
int64_t svs_kernel_launch_count(void);

/* The packed kernels are persistent and by default occupy every SM (one CTA per SM with the
 * whole register file).  A caller that runs a collective (NCCL all-gather of the extracted bits)
 * concurrently on another stream leaves `n` SMs free for it; n < 0 only queries.  Returns the
 * previous value.  Process-wide. */
int svs_set_reserved_sms(int n);

/* Stream-ordered device-to-device copy on the copy engines (cudaMemcpyAsync) between raw device
 * addresses.  Exists for the multi-GPU exchange of the extracted bit rows (sharding.py): `d_dst`
 * may be a peer mapping or an NVSwitch MULTICAST address of the gathered buffer - the copy engines
 * can write to it, and the switch then replicates the rows into every rank's buffer (one outbound
 * copy instead of one per peer; profiles/microbench/ce_multicast.py).  No counterpart in the
 * reference (single process). */
int svs_memcpy_d2d_async(void* d_dst, const void* d_src, int64_t bytes, void* stream);

/* Diagnostic: selects the kernel family, process-wide.  0 = automatic (default: the packed
 * one-block-per-thread kernels of svs_block.cuh whenever they apply, the scalar kernels
 * otherwise), 1 = scalar kernels only, 5 = packed block kernels.  A library built with
 * -DSVS_WITH_VARIANTS (measurement builds, profiles/build_variant.sh) also knows 2 = packed
 * lockstep (two blocks per thread), 3 = packed tile, 4 = packed row - the round-1 organisations.
 * Returns the previous setting; a negative argument only queries; -1 is returned (and nothing
 * changes) when this build does not contain the requested family.  All families produce
 * identical results; the tests use this to prove it. */
int svs_debug_kernel_family(int family_id);

#ifdef __cplusplus
}
#endif
#endif /* SVS_B200_H */
