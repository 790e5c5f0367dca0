/* gh_codec.h -- C ABI of the B200-native canonical Huffman codec (drop-in for golden-huffman's ".crs2" path).
 *
 * This is the boundary a reference-side binding talks to: plain pointers and sizes, no C++/torch types.
 * Each entry point names the reference interface it replaces (paths relative to the reference root,
 * chenghuige/golden-huffman).  Everything the reference does per byte on its hot path
 *     Compressor<CanonicalHuffEncoder<> >::compress()      include/compressor.h:62-73
 *     Decompressor<CanonicalHuffDecoder<> >::decompress()  include/compressor.h:87-92
 * runs in hand-written sm_100a kernels behind these calls; the 257-symbol code construction and the file
 * header stay on the host with the reference's exact tie-breaking.  There is NO CPU fallback: every device
 * entry point returns GH_ERR_CUDA if no CUDA device / kernel image is usable.
 *
 * Conventions
 *   - all functions return an int status (GH_OK = 0); the reference has no error channel at all
 *     (include/encoder.h:67-70), so inputs on which the reference is undefined are rejected explicitly:
 *     empty input (GH_ERR_EMPTY) and code lengths > 32 (GH_ERR_TOO_LONG)  -- SURVEY.md D7, D11;
 *   - `d_` pointers are device pointers owned by the caller; `stream` is a cudaStream_t passed as void*
 *     (NULL = the legacy default stream); device entry points are asynchronous on `stream` unless stated;
 *   - workspaces are caller-allocated device memory of at least gh_*_workspace_bytes() bytes, 256-byte aligned;
 *   - bit order is the reference's: one MSB-first bit string (utils/include/buffer.h:241-248,290-295),
 *     header words big-endian (utils/include/buffer.h:255-268).
 */
#ifndef GH_CODEC_H_
#define GH_CODEC_H_

#include <stddef.h>
#include <stdint.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GH_NSYM 257 /* include/type_traits.h:50  CharSymbolNum: bytes 0..255 + the end-of-encoding mark */
#define GH_EOF_SYMBOL 256
#define GH_MAX_CODE_LEN 32 /* include/canonical_huff_encoder.h:43-44, utils/include/buffer.h:288-295 */

enum gh_status {
  GH_OK = 0,
  GH_ERR_EMPTY = 1,     /* empty input */
  GH_ERR_TOO_LONG = 2,  /* a code length > 32 */
  GH_ERR_SPACE = 3,     /* an output / workspace buffer is too small */
  GH_ERR_FORMAT = 4,    /* malformed header or stream */
  GH_ERR_NO_EOF = 5,    /* the stream ends before an end-of-encoding mark is decoded */
  GH_ERR_ARG = 6,       /* bad argument (NULL, misaligned pointer, ...) */
  GH_ERR_CUDA = 7       /* CUDA runtime / launch failure (see gh_last_cuda_error) */
};

/* The code tables of CanonicalHuffEncoder / CanonicalHuffDecoder
 * (include/canonical_huff_encoder.h:107-120, 142-150), 257 entries wide. */
typedef struct gh_code {
  uint32_t length[GH_NSYM];   /* length_[s], 0 = symbol absent */
  uint32_t codeword[GH_NSYM]; /* codeword_[s], right-justified */
  uint32_t symbol[GH_NSYM];   /* symbol_[]: symbols bucket-sorted by (length, value); unused = 0xFFFFFFFF */
  uint32_t min_len;
  uint32_t max_len;
  uint32_t start_pos[33];  /* start_pos_[1..max_len] */
  uint32_t first_code[33]; /* first_code_[1..max_len]; 1024 sentinel below min_len */
} gh_code;

/* The kernels' lookup form of the encoder's tables: (codeword, length) per symbol, lengths as bytes. */
typedef struct gh_encode_table {
  uint32_t codeword[GH_NSYM];
  uint8_t length[GH_NSYM + 3];
} gh_encode_table;

/* What gh_build_code_device leaves in DEVICE memory: the code, its status and sizes, and the encode table. */
typedef struct gh_device_code {
  gh_code code;
  uint32_t status;        /* GH_OK, GH_ERR_EMPTY (all counters zero) or GH_ERR_TOO_LONG (a length > 32) */
  uint32_t header_bytes;  /* 1040 + 8 * max_len */
  uint32_t reserved;      /* keeps the 64-bit fields 8-byte aligned without implicit padding */
  uint64_t payload_bits;  /* sum of length * count + the end mark; the 1-padding is not included */
  uint64_t total_symbols; /* sum of the byte counters */
  gh_encode_table table;
} gh_device_code;

const char* gh_strerror(int status);
/* cudaError_t of the last failing CUDA call made by this library on the calling thread (0 = none) */
int gh_last_cuda_error(void);
/* number of kernel launches issued by this library since load (bench.py's `gpu_launches`) */
uint64_t gh_launch_count(void);
/* Per-kernel device timing for bench.py's roofline pass: while enabled every launch is bracketed by CUDA
 * events on its stream; gh_profile_fetch waits for them and writes "kernel_name launches total_ms\n" lines
 * (and resets the counters). Returns the bytes written. Off by default. */
void gh_profile_enable(int on);
size_t gh_profile_fetch(char* buf, size_t cap);
/* Test hooks (explicit calls: the library reads no environment variables). Output is identical either way.
 * gh_debug_select_writer: 0 automatic, 1 always the thread-per-subsequence pipeline; 2 (warp per 2 KiB segment)
 * exists only in builds with -DGH_EXPERIMENTS and is ignored otherwise.
 * gh_debug_disable_phase_walk: non-zero makes codes of 8 and 9 bits take the general re-walk rounds instead of
 * the phase walk. */
void gh_debug_select_writer(int pipeline);
void gh_debug_disable_phase_walk(int off);

/* ------------------------------------------------------------------------------------------------
 * Host: code construction and the file header (microseconds; stays on the host by design)
 * ---------------------------------------------------------------------------------------------- */

/* Replaces Encoder::do_init (include/encoder.h:123-129: frequency_map_[256] = 1),
 * CanonicalHuffEncoder::get_encoding_length (include/canonical_huff_encoder.cc:289-345) and
 * ::do_gen_encode (:69-141). Ties are broken exactly as the reference's
 * std::priority_queue<int, std::deque<int>, HuffNodeIndexGreater> does (include/canonical_huff_encoder.h:58-70). */
int gh_build_code(const uint64_t hist256[256], gh_code* code);

/* The same construction ON THE DEVICE, by one warp: bit-for-bit gh_build_code's result for
 * the sum of the n_hists histograms at d_hists (n_hists * 256 counters, device memory; n_hists > 1 = the shards of
 * one input), left in *d_code (device memory). d_header, when not NULL, receives the header's bytes (device memory,
 * 4-byte aligned, >= 1296 bytes): write_encode_info (include/canonical_huff_encoder.cc:210-242) without the host.
 * Nothing is synchronised: the result is consumed by later work on the same stream (gh_compress_device does so when
 * gh_ctx_set_device_code is on). */
int gh_build_code_device(const uint64_t* d_hists, int n_hists, gh_device_code* d_code, uint8_t* d_header, void* stream);

/* Header size = 1040 + 8*max_len. Replaces CanonicalHuffEncoder::write_encode_info
 * (include/canonical_huff_encoder.cc:210-242). */
size_t gh_header_bytes(const gh_code* code);
int gh_write_header(const gh_code* code, uint8_t* dst, size_t cap, size_t* written);

/* Replaces CanonicalHuffDecoder::get_encode_info (include/canonical_huff_encoder.cc:349-374).
 * Only symbol/min_len/max_len/start_pos/first_code come from the file; length/codeword are derived. */
int gh_parse_header(const uint8_t* src, size_t n, gh_code* code, size_t* header_bytes);

/* Sum over bytes of length[b]*hist[b] (+ length[256] when with_eof): payload bits before the 1-padding. */
uint64_t gh_payload_bits(const gh_code* code, const uint64_t hist256[256], int with_eof);

/* ------------------------------------------------------------------------------------------------
 * Device: the data-parallel path
 * ---------------------------------------------------------------------------------------------- */

/* K1. Replaces Encoder::do_caculate_frequency(char_tag) (include/encoder.h:136-150).
 * d_hist256[b] (+)= number of bytes equal to b in d_in[0..n). accumulate = 0 overwrites the 256 counters. */
int gh_histogram(const uint8_t* d_in, uint64_t n, uint64_t* d_hist256, int accumulate, void* stream);

/* K2-K4. Replaces CanonicalHuffEncoder::encode_file / encode_each_byte
 * (include/canonical_huff_encoder.cc:245-285) and the bit writer FixedFileBuffer::write_bits / flush_bits
 * (utils/include/buffer.h:277-295).
 *
 * Appends the codewords of d_in[0..n) to the bit string held in d_payload, starting at bit `start_bit`
 * (bit 0 = MSB of d_payload[0]). With append_eof != 0 the end-of-encoding codeword follows and the last
 * byte is padded with 1-bits, exactly as the reference ends a file.  `start_bit` exists for sharding: a
 * shard passes its global start bit modulo 128 and gets a phase-aligned slice; bits of the first 32-bit
 * word that lie before start_bit are written as 0 so the caller can OR the neighbour's tail in.
 * *d_end_bit (device, optional) receives start_bit + bits appended (end mark included, padding excluded).
 * d_payload must be 16-byte aligned with payload_cap >= gh_encode_payload_capacity(...).            */
size_t gh_encode_workspace_bytes(uint64_t n);
uint64_t gh_encode_payload_capacity(uint64_t n, const gh_code* code, uint64_t start_bit);
int gh_encode(const uint8_t* d_in, uint64_t n, const gh_code* code, uint64_t start_bit, int append_eof,
              uint8_t* d_payload, uint64_t payload_cap, uint64_t* d_end_bit, void* d_workspace,
              size_t workspace_bytes, void* stream);

/* K5-K7. Replaces CanonicalHuffDecoder::decode_file (include/canonical_huff_encoder.cc:377-419), and is
 * result-identical to FastCanonicalHuffDecoder / TableCanonicalHuffDecoder (:422-461, :519-568).
 *
 * Decodes the unindexed serial bit string d_payload[0..payload_bytes) by self-synchronising subsequence
 * decoding and stops at the first end-of-encoding mark on the true decode path.
 * Synchronous with respect to `stream` (the symbol count is returned to the host in *n_out).
 * GH_ERR_SPACE if more than out_cap symbols precede the end mark (*n_out is still set).
 * d_payload must be 16-byte aligned; 32-byte alignment lets the kernels read one whole sector per lane
 * and request (256-bit loads) instead of two 128-bit loads.                                          */
size_t gh_decode_workspace_bytes(uint64_t payload_bytes);
int gh_decode(const uint8_t* d_payload, uint64_t payload_bytes, const gh_code* code, uint8_t* d_out,
              uint64_t out_cap, uint64_t* n_out, void* d_workspace, size_t workspace_bytes, void* stream);

/* The two halves of gh_decode, exposed for sharded (multi-GPU) decoding.
 * A shard is a 16-byte aligned slice of the payload; `readable_bytes` >= slice_bytes says how far past the
 * slice the pointer may be read (halo of the next shard: 8 bytes are enough; 0 extra on the last shard).
 * gh_decode_sync finds every subsequence's first codeword given that the slice's first codeword starts at
 * bit `entry_bit`, and reports where decoding leaves the slice. It may be called again with a corrected
 * entry_bit (re-using the workspace): only the affected subsequences are re-walked.
 * gh_decode_write then emits the slice's symbols to d_out[0..n_symbols).                              */
typedef struct gh_shard_sync {
  uint64_t n_symbols;  /* symbols decoded in this slice on the path from entry_bit (up to the end mark) */
  uint32_t exit_bit;   /* offset into the next slice at which its first codeword starts */
  uint32_t eof_found;  /* 1 if the end-of-encoding mark lies in this slice on that path */
  uint32_t rounds;     /* synchronisation rounds used (diagnostic) */
  uint32_t sub_bytes;  /* subsequence size used (diagnostic) */
} gh_shard_sync;

int gh_decode_sync(const uint8_t* d_payload, uint64_t slice_bytes, uint64_t readable_bytes, const gh_code* code,
                   uint32_t entry_bit, int first_call, gh_shard_sync* result, void* d_workspace,
                   size_t workspace_bytes, void* stream);
int gh_decode_write(const uint8_t* d_payload, uint64_t slice_bytes, uint64_t readable_bytes, const gh_code* code,
                    uint8_t* d_out, uint64_t out_cap, void* d_workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Whole ".crs2" images from / to HOST buffers (what Compressor::compress / Decompressor::decompress do
 * for a file, minus the file I/O): H2D copy, kernels, D2H copy, synchronous.
 * ---------------------------------------------------------------------------------------------- */
typedef struct gh_ctx gh_ctx; /* device scratch buffers + a private stream on the current device */
int gh_ctx_create(gh_ctx** ctx);
void gh_ctx_destroy(gh_ctx* ctx);
/* run the context's work on the caller's stream (e.g. the framework's current stream) instead of its own */
int gh_ctx_set_stream(gh_ctx* ctx, void* stream);
/* on != 0: gh_compress_device / gh_compress_host build the code on the device (gh_build_code_device) and run
 * histogram -> code -> header -> packing back to back on the stream, with ONE read-back at the end; 0 (default):
 * the histogram is read back and the code is built on the host (gh_build_code). Same image either way. */
int gh_ctx_set_device_code(gh_ctx* ctx, int on);
/* gh_decompress_host moves the payload up and the decoded bytes back in chunks, on two copy engines beside the
 * kernels (chunk k+1 uploads and chunk k-1 reads back while chunk k decodes). `bytes` = payload bytes per chunk
 * (rounded up to 4 KiB; 0 = the default, 64 MiB). The decoded bytes do not depend on it. */
int gh_ctx_set_host_chunk(gh_ctx* ctx, uint64_t bytes);

uint64_t gh_compress_bound(uint64_t n); /* 1040 + 8*32 + 4*n + 32 */
int gh_compress_host(gh_ctx* ctx, const uint8_t* in, uint64_t n, uint8_t* out, uint64_t cap, uint64_t* out_bytes);
int gh_decompress_host(gh_ctx* ctx, const uint8_t* in, uint64_t n, uint8_t* out, uint64_t cap, uint64_t* out_bytes);

/* Same on N GPUs of one box from ONE process (no NCCL: the data comes from and returns to host memory, so what the shards
 * exchange -- 256 counters, one boundary byte, one bit position -- crosses on the host): the input is cut into n_shards
 * contiguous shards, shard d runs on device devices[d] (NULL: device d) in its own host thread. A device may be listed
 * more than once. Same image / same bytes as the single-GPU calls. */
int gh_compress_host_multi(int n_shards, const int* devices, const uint8_t* in, uint64_t n, uint8_t* out, uint64_t cap,
                           uint64_t* out_bytes);
int gh_decompress_host_multi(int n_shards, const int* devices, const uint8_t* in, uint64_t n, uint8_t* out, uint64_t cap,
                             uint64_t* out_bytes);

/* Same, with device-resident input and output (no PCIe traffic): the kernels-only path bench.py times. */
int gh_compress_device(gh_ctx* ctx, const uint8_t* d_in, uint64_t n, uint8_t* d_out, uint64_t cap,
                       uint64_t* out_bytes);
int gh_decompress_device(gh_ctx* ctx, const uint8_t* d_in, uint64_t n, uint8_t* d_out, uint64_t cap,
                         uint64_t* out_bytes);

/* ------------------------------------------------------------------------------------------------
 * Streaming files of any size through a fixed amount of memory: what the reference does with its 64 KiB
 * FixedFileBuffer (include/encoder.h:55,136-150; include/canonical_huff_encoder.cc:245-285, 377-419), with chunks of
 * `chunk_bytes` (0 = 64 MiB) through two pinned host buffers: the file I/O, the PCIe copies and the kernels overlap.
 * Inputs up to `resident_max_bytes` (~0 = a quarter of the free device memory) stay on the device between the two
 * passes of compress; larger ones are read twice, like the reference reads its input twice.
 *   compress   = gh_stream_histogram (pass 1) -> gh_build_code -> gh_write_header -> gh_stream_encode (pass 2)
 *   decompress = gh_parse_header -> gh_stream_decode
 * FILE* positions: histogram reads from the current position to the end and seeks back; encode reads the same range
 * and appends the payload to `out`; decode expects `in` at the first payload byte (header_bytes into the file).
 * ---------------------------------------------------------------------------------------------- */
typedef struct gh_stream gh_stream;
int gh_stream_create(gh_stream** s, uint64_t chunk_bytes, uint64_t resident_max_bytes);
void gh_stream_destroy(gh_stream* s);
int gh_stream_histogram(gh_stream* s, FILE* in, uint64_t hist256[256]);
int gh_stream_encode(gh_stream* s, FILE* in, FILE* out, const gh_code* code, uint64_t* payload_bytes);
int gh_stream_decode(gh_stream* s, FILE* in, FILE* out, const gh_code* code, uint64_t header_bytes, uint64_t* n_out);

/* ------------------------------------------------------------------------------------------------
 * Staged variants, one per step of the reference's template methods, for adapters that must keep the
 * reference's call order (Compressor::compress: caculate_frequency -> gen_encode -> write_encode_info ->
 * encode_file; Decompressor::decompress: get_encode_info -> decode_file). Device buffers live in the ctx
 * between the calls. All synchronous.
 * ---------------------------------------------------------------------------------------------- */
/* caculate_frequency(): H2D copy of the whole input + K1; the 256 counters come back to the host */
int gh_stage_input(gh_ctx* ctx, const uint8_t* in, uint64_t n, uint64_t hist256[256]);
/* encode_file(): K2-K4 over the staged input, then D2H of the payload (ceil(bits/8) bytes) */
int gh_encode_staged(gh_ctx* ctx, const gh_code* code, uint8_t* out_payload, uint64_t cap, uint64_t* payload_bytes);
/* decode_file(), first half: H2D of the payload + K5/K6; reports how many symbols precede the end mark */
int gh_stage_payload(gh_ctx* ctx, const uint8_t* payload, uint64_t nbytes, const gh_code* code, uint64_t* n_symbols);
/* decode_file(), second half: K7 + D2H of the n_symbols bytes reported by gh_stage_payload */
int gh_decode_staged(gh_ctx* ctx, uint8_t* out, uint64_t cap);

#ifdef __cplusplus
}
#endif
#endif /* GH_CODEC_H_ */
