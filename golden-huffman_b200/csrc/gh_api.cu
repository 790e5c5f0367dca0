// Whole ".crs2" images: what Compressor<CanonicalHuffEncoder<> >::compress() and
// Decompressor<CanonicalHuffDecoder<> >::decompress() (reference include/compressor.h:62-73, 87-92) do for a
// file, minus the file I/O. Device-resident variants (kernels only) and host-buffer variants (H2D + kernels + D2H).
#include <stddef.h>
#include <string.h>

#include <stdlib.h>

#include "gh_common.cuh"

struct gh_ctx {
  cudaStream_t stream;
  bool own_stream;
  uint8_t* d_in;
  size_t in_cap;
  uint8_t* d_out;
  size_t out_cap;
  void* d_ws;
  size_t ws_cap;
  uint64_t* d_small;  // 256 histogram counters + end bit
  uint8_t* h_small;   // pinned: histogram read-back, header staging
  gh_device_code* d_code;  // code built on the device (gh_ctx_set_device_code)
  bool device_code;
  // copy engines of the host entry points: uploads and read-backs run beside the kernels, in chunks
  cudaStream_t copy_in, copy_out;
  cudaEvent_t* ev_in;               // [ev_in_n] chunk k is on the device
  size_t ev_in_n;
  cudaEvent_t ev_dec[2];            // chunk k (mod 2) is decoded
  uint64_t host_chunk;              // payload bytes per pipeline step of gh_decompress_host
  // state of the staged (step-by-step) entry points
  uint64_t staged_n;         // input bytes resident in d_in (gh_stage_input)
  uint64_t staged_payload;   // payload bytes resident in d_in (gh_stage_payload)
  uint64_t staged_symbols;   // symbols gh_stage_payload found
  gh_code staged_code;
};

namespace gh {

constexpr size_t kSmallBytes = 4096;
constexpr size_t kMaxHeader = 1040 + 8 * 32;

constexpr uint64_t kHostChunk = 64ull << 20;  // default payload bytes per pipeline step of gh_decompress_host
constexpr size_t kMaxHostChunks = 4096;
constexpr uint64_t kHostHalo = 64;            // bytes of the next chunk a chunk's decode may read

// the side streams and `n_in` upload events, created on first use
static int ensure_copy_engines(gh_ctx* c, size_t n_in) {
  if (!c->copy_in) GH_CUDA_TRY(cudaStreamCreateWithFlags(&c->copy_in, cudaStreamNonBlocking));
  if (!c->copy_out) GH_CUDA_TRY(cudaStreamCreateWithFlags(&c->copy_out, cudaStreamNonBlocking));
  for (int i = 0; i < 2; ++i)
    if (!c->ev_dec[i]) GH_CUDA_TRY(cudaEventCreateWithFlags(&c->ev_dec[i], cudaEventDisableTiming));
  if (c->ev_in_n < n_in) {
    cudaEvent_t* grown = static_cast<cudaEvent_t*>(realloc(c->ev_in, n_in * sizeof(cudaEvent_t)));
    if (!grown) return GH_ERR_ARG;
    c->ev_in = grown;
    while (c->ev_in_n < n_in) {
      GH_CUDA_TRY(cudaEventCreateWithFlags(&c->ev_in[c->ev_in_n], cudaEventDisableTiming));
      ++c->ev_in_n;
    }
  }
  return GH_OK;
}

// Nothing may still be copying from or to the caller's buffers when a host entry point returns, whatever the path out.
struct CopyEngineFence {
  gh_ctx* c;
  ~CopyEngineFence() {
    if (c->copy_in) cudaStreamSynchronize(c->copy_in);
    if (c->copy_out) cudaStreamSynchronize(c->copy_out);
  }
};

static int grow(void** p, size_t* cap, size_t need) {
  if (*cap >= need) return GH_OK;
  if (*p) GH_CUDA_TRY(cudaFree(*p));
  *p = nullptr;
  *cap = 0;
  const size_t want = need + need / 8 + 4096;
  GH_CUDA_TRY(cudaMalloc(p, want));
  *cap = want;
  return GH_OK;
}

}  // namespace gh

extern "C" {

int gh_ctx_create(gh_ctx** out) {
  using namespace gh;
  if (!out) return GH_ERR_ARG;
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) return cuda_fail(e);
  gh_ctx* c = new gh_ctx();
  memset(c, 0, sizeof(*c));
  if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaMalloc(reinterpret_cast<void**>(&c->d_small), kSmallBytes) != cudaSuccess ||
      cudaMalloc(reinterpret_cast<void**>(&c->d_code), sizeof(gh_device_code)) != cudaSuccess ||
      cudaMallocHost(reinterpret_cast<void**>(&c->h_small), kSmallBytes) != cudaSuccess) {
    const int rc = cuda_fail(cudaGetLastError());
    gh_ctx_destroy(c);
    return rc;
  }
  c->own_stream = true;
  c->host_chunk = kHostChunk;
  *out = c;
  return GH_OK;
}

int gh_ctx_set_host_chunk(gh_ctx* c, uint64_t bytes) {
  if (!c) return GH_ERR_ARG;
  c->host_chunk = bytes ? (bytes + 4095) / 4096 * 4096 : gh::kHostChunk;
  return GH_OK;
}

int gh_ctx_set_stream(gh_ctx* c, void* stream) {
  if (!c) return GH_ERR_ARG;
  if (c->own_stream) cudaStreamDestroy(c->stream);
  c->own_stream = false;
  c->stream = (cudaStream_t)stream;
  return GH_OK;
}

int gh_ctx_set_device_code(gh_ctx* c, int on) {
  if (!c) return GH_ERR_ARG;
  c->device_code = on != 0;
  return GH_OK;
}

void gh_ctx_destroy(gh_ctx* c) {
  if (!c) return;
  if (c->d_in) cudaFree(c->d_in);
  if (c->d_out) cudaFree(c->d_out);
  if (c->d_ws) cudaFree(c->d_ws);
  if (c->d_small) cudaFree(c->d_small);
  if (c->d_code) cudaFree(c->d_code);
  if (c->h_small) cudaFreeHost(c->h_small);
  for (size_t i = 0; i < c->ev_in_n; ++i) cudaEventDestroy(c->ev_in[i]);
  free(c->ev_in);
  for (int i = 0; i < 2; ++i)
    if (c->ev_dec[i]) cudaEventDestroy(c->ev_dec[i]);
  if (c->copy_in) cudaStreamDestroy(c->copy_in);
  if (c->copy_out) cudaStreamDestroy(c->copy_out);
  if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

int gh_compress_device(gh_ctx* c, const uint8_t* d_in, uint64_t n, uint8_t* d_out, uint64_t cap, uint64_t* out_bytes) {
  using namespace gh;
  if (!c || !d_out || !out_bytes) return GH_ERR_ARG;
  c->staged_n = c->staged_payload = 0;  // the context's buffers are about to be reused: nothing stays staged
  if (n == 0) return GH_ERR_EMPTY;
  if (!d_in || (reinterpret_cast<uintptr_t>(d_in) & 15) || (reinterpret_cast<uintptr_t>(d_out) & 15)) return GH_ERR_ARG;
  // 1. histogram (Encoder::caculate_frequency)
  int rc = gh_histogram(d_in, n, c->d_small, 0, c->stream);
  if (rc != GH_OK) return rc;
  if (c->device_code) {
    // 2.-3. code, header and payload without the host: K1 -> build_code_kernel -> packer, one read-back at the end
    rc = gh_build_code_device(c->d_small, 1, c->d_code, d_out, c->stream);
    if (rc != GH_OK) return rc;
    rc = grow(&c->d_ws, &c->ws_cap, gh_encode_workspace_bytes(n));
    if (rc != GH_OK) return rc;
    uint64_t* d_end = c->d_small + 256;
    rc = encode_with_device_code(d_in, n, c->d_code, d_out, cap, d_end, c->d_ws, c->ws_cap, c->stream);
    if (rc != GH_OK) return rc;
    struct Tail {  // gh_device_code from `status` on
      uint32_t status, header_bytes, reserved;
      uint32_t payload_bits_lo, payload_bits_hi;  // (the struct's 64-bit fields are 8-byte aligned, `status` is not)
    };
    static_assert(offsetof(gh_device_code, payload_bits) - offsetof(gh_device_code, status) == offsetof(Tail, payload_bits_lo),
                  "Tail mirrors gh_device_code");
    Tail* h_tail = reinterpret_cast<Tail*>(c->h_small);
    GH_CUDA_TRY(cudaMemcpyAsync(h_tail, &c->d_code->status, sizeof(Tail), cudaMemcpyDeviceToHost, c->stream));
    GH_CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (h_tail->status != uint32_t(GH_OK)) return int(h_tail->status);
    const uint64_t payload_bits = (uint64_t(h_tail->payload_bits_hi) << 32) | h_tail->payload_bits_lo;
    const uint64_t bytes = uint64_t(h_tail->header_bytes) + (payload_bits + 7) / 8;
    if (cap < (bytes + 3) / 4 * 4) return GH_ERR_SPACE;  // the packer's stores were bounded by cap
    *out_bytes = bytes;
    return GH_OK;
  }
  // 2. counters back to the host
  uint64_t* h_hist = reinterpret_cast<uint64_t*>(c->h_small);
  GH_CUDA_TRY(cudaMemcpyAsync(h_hist, c->d_small, 256 * 8, cudaMemcpyDeviceToHost, c->stream));
  GH_CUDA_TRY(cudaStreamSynchronize(c->stream));
  // 2. code + header on the host (gen_encode, write_encode_info)
  gh_code code;
  rc = gh_build_code(h_hist, &code);
  if (rc != GH_OK) return rc;
  const size_t hdr = gh_header_bytes(&code);
  const uint64_t bits = gh_payload_bits(&code, h_hist, 1);
  const uint64_t payload = (bits + 7) / 8;
  const uint64_t need = (hdr + payload + 3) / 4 * 4;  // the packer stores whole 32-bit words
  if (cap < need) return GH_ERR_SPACE;
  uint8_t* h_hdr = c->h_small + 2048;
  size_t written = 0;
  rc = gh_write_header(&code, h_hdr, kMaxHeader, &written);
  if (rc != GH_OK) return rc;
  // 3. payload (encode_file): the header is 8-byte aligned, the kernels work from the 32-byte boundary below it ->
  //    the packer starts up to 192 bits into that sector
  const size_t base = hdr & ~size_t(31);
  const uint64_t start_bit = uint64_t(hdr - base) * 8;
  rc = grow(&c->d_ws, &c->ws_cap, gh_encode_workspace_bytes(n));
  if (rc != GH_OK) return rc;
  rc = encode_unchecked(d_in, n, &code, start_bit, 1, d_out + base, (cap - base) / 4 * 4, nullptr, c->d_ws, c->ws_cap,
                        c->stream);
  if (rc != GH_OK) return rc;
  GH_CUDA_TRY(cudaMemcpyAsync(d_out, h_hdr, hdr, cudaMemcpyHostToDevice, c->stream));
  GH_CUDA_TRY(cudaStreamSynchronize(c->stream));
  *out_bytes = hdr + payload;
  return GH_OK;
}

int gh_decompress_device(gh_ctx* c, const uint8_t* d_in, uint64_t n, uint8_t* d_out, uint64_t cap, uint64_t* out_bytes) {
  using namespace gh;
  if (!c || !d_in || !out_bytes || (!d_out && cap)) return GH_ERR_ARG;
  c->staged_n = c->staged_payload = 0;  // the context's buffers are about to be reused: nothing stays staged
  if (reinterpret_cast<uintptr_t>(d_in) & 15) return GH_ERR_ARG;
  // 1. header (get_encode_info)
  const size_t peek = n < kMaxHeader ? size_t(n) : kMaxHeader;
  uint8_t* h_hdr = c->h_small + 2048;
  GH_CUDA_TRY(cudaMemcpyAsync(h_hdr, d_in, peek, cudaMemcpyDeviceToHost, c->stream));
  GH_CUDA_TRY(cudaStreamSynchronize(c->stream));
  gh_code code;
  size_t hdr = 0;
  int rc = gh_parse_header(h_hdr, peek, &code, &hdr);
  if (rc != GH_OK) return rc;
  if (n <= hdr) return GH_ERR_NO_EOF;
  // 2. payload (decode_file), read from the 32-byte boundary below it (one sector per lane and request)
  const size_t base = hdr & ~size_t(31);
  const uint64_t slice = n - base;
  rc = grow(&c->d_ws, &c->ws_cap, gh_decode_workspace_bytes(slice));
  if (rc != GH_OK) return rc;
  return decode_full(d_in + base, slice, &code, uint32_t(hdr - base) * 8, d_out, cap, out_bytes, c->d_ws, c->ws_cap,
                     c->stream);
}

int gh_compress_host(gh_ctx* c, const uint8_t* in, uint64_t n, uint8_t* out, uint64_t cap, uint64_t* out_bytes) {
  using namespace gh;
  if (!c || !out || !out_bytes) return GH_ERR_ARG;
  c->staged_n = c->staged_payload = 0;  // the context's buffers are about to be reused: nothing stays staged
  if (n == 0) return GH_ERR_EMPTY;
  if (!in) return GH_ERR_ARG;
  int rc = grow(reinterpret_cast<void**>(&c->d_in), &c->in_cap, n + 16);
  if (rc != GH_OK) return rc;
  GH_CUDA_TRY(cudaMemcpyAsync(c->d_in, in, n, cudaMemcpyHostToDevice, c->stream));
  // the device image may need up to the full bound; the caller's buffer only has to hold the result
  const uint64_t bound = gh_compress_bound(n);
  const uint64_t dev_cap = cap + 16 < bound ? cap + 16 : bound;
  rc = grow(reinterpret_cast<void**>(&c->d_out), &c->out_cap, dev_cap);
  if (rc != GH_OK) return rc;
  uint64_t bytes = 0;
  rc = gh_compress_device(c, c->d_in, n, c->d_out, c->out_cap, &bytes);
  if (rc != GH_OK) return rc;
  if (bytes > cap) return GH_ERR_SPACE;
  GH_CUDA_TRY(cudaMemcpyAsync(out, c->d_out, bytes, cudaMemcpyDeviceToHost, c->stream));
  GH_CUDA_TRY(cudaStreamSynchronize(c->stream));
  *out_bytes = bytes;
  return GH_OK;
}

// Decompressor::decompress() from and to host memory. The payload goes up in chunks on one copy engine while the
// chunks already there are decoded and their bytes come back on the other: a chunk's first codeword starts where
// the previous chunk's decode ran out (gh_decode_sync's exit bit), so the chunks decode in order with nothing but
// that bit position between them, and the call takes about as long as its larger copy (the read-back) instead of
// upload + kernels + read-back (PCIe is full duplex).
int gh_decompress_host(gh_ctx* c, const uint8_t* in, uint64_t n, uint8_t* out, uint64_t cap, uint64_t* out_bytes) {
  using namespace gh;
  if (!c || !in || !out_bytes || (!out && cap)) return GH_ERR_ARG;
  *out_bytes = 0;
  c->staged_n = c->staged_payload = 0;  // the context's buffers are about to be reused: nothing stays staged
  // 1. header (get_encode_info), straight from the caller's bytes
  gh_code code;
  size_t hdr = 0;
  int rc = gh_parse_header(in, n < kMaxHeader ? size_t(n) : kMaxHeader, &code, &hdr);
  if (rc != GH_OK) return rc;
  if (n <= hdr) return GH_ERR_NO_EOF;
  // 2. payload (decode_file), read from the 32-byte boundary below it (one sector per lane and request); the bytes
  // between that boundary and the payload are skipped through chunk 0's entry bit
  const uint64_t base = hdr & ~uint64_t(31);
  const uint64_t stream_bytes = n - base;
  // chunk boundaries (stream offsets, multiples of 4 KiB): the first chunks are small and double up to the configured
  // size, so that the read-back -- the longer of the two copies -- starts after a fraction of a millisecond of upload
  const uint64_t P = c->host_chunk;
  uint64_t cuts[kMaxHostChunks + 1];
  size_t nchunks = 0;
  {
    uint64_t at = 0, step = P / 16 >= 4096 ? (P / 16 + 4095) / 4096 * 4096 : P;
    const uint64_t rest_chunks = (stream_bytes + P - 1) / P;
    if (rest_chunks + 8 > kMaxHostChunks) step = (stream_bytes / (kMaxHostChunks - 8) + 4095) / 4096 * 4096;  // huge inputs: fewer, larger chunks
    while (at < stream_bytes) {
      cuts[nchunks++] = at;
      at += step;
      if (step < P) step = step * 2 < P ? step * 2 : P;
    }
    cuts[nchunks] = stream_bytes;
  }
  uint64_t largest = 0;
  for (size_t k = 0; k < nchunks; ++k) largest = cuts[k + 1] - cuts[k] > largest ? cuts[k + 1] - cuts[k] : largest;
  rc = grow(reinterpret_cast<void**>(&c->d_in), &c->in_cap, n + 64);
  if (rc != GH_OK) return rc;
  rc = grow(reinterpret_cast<void**>(&c->d_out), &c->out_cap, cap + 64);
  if (rc != GH_OK) return rc;
  rc = grow(&c->d_ws, &c->ws_cap, gh_decode_workspace_bytes(largest + kHostHalo));
  if (rc != GH_OK) return rc;
  rc = ensure_copy_engines(c, nchunks);
  if (rc != GH_OK) return rc;
  CopyEngineFence fence{c};
  for (size_t k = 0; k < nchunks; ++k) {
    const uint64_t off = base + cuts[k];
    GH_CUDA_TRY(cudaMemcpyAsync(c->d_in + off, in + off, size_t(cuts[k + 1] - cuts[k]), cudaMemcpyHostToDevice, c->copy_in));
    GH_CUDA_TRY(cudaEventRecord(c->ev_in[k], c->copy_in));
  }
  uint32_t entry = uint32_t(hdr - base) * 8;
  uint64_t total = 0;
  bool done = false, overflow = false;
  for (size_t k = 0; k < nchunks && !done; ++k) {
    const uint64_t left = stream_bytes - cuts[k];
    const uint64_t slice = cuts[k + 1] - cuts[k];
    const uint64_t readable = left < slice + kHostHalo ? left : slice + kHostHalo;
    // the chunk and the halo it may read (the head of the next chunk) are on the device
    GH_CUDA_TRY(cudaStreamWaitEvent(c->stream, c->ev_in[k + 1 < nchunks ? k + 1 : k], 0));
    const uint8_t* d = c->d_in + base + cuts[k];
    gh_shard_sync res;
    rc = gh_decode_sync(d, slice, readable, &code, entry, 1, &res, c->d_ws, c->ws_cap, c->stream);
    if (rc != GH_OK) return rc;
    const uint64_t nsym = res.n_symbols;
    if (overflow || nsym > cap - total) overflow = true;  // as gh_decode: keep counting, write nothing more
    if (nsym && !overflow) {
      rc = gh_decode_write(d, slice, readable, &code, c->d_out + total, nsym, c->d_ws, c->ws_cap, c->stream);
      if (rc != GH_OK) return rc;
      GH_CUDA_TRY(cudaEventRecord(c->ev_dec[k & 1], c->stream));
      GH_CUDA_TRY(cudaStreamWaitEvent(c->copy_out, c->ev_dec[k & 1], 0));
      GH_CUDA_TRY(cudaMemcpyAsync(out + total, c->d_out + total, size_t(nsym), cudaMemcpyDeviceToHost, c->copy_out));
    }
    total += nsym;
    entry = res.exit_bit;
    done = res.eof_found != 0;
  }
  *out_bytes = total;
  GH_CUDA_TRY(cudaStreamSynchronize(c->copy_out));
  GH_CUDA_TRY(cudaStreamSynchronize(c->stream));
  if (!done) return GH_ERR_NO_EOF;
  return overflow ? GH_ERR_SPACE : GH_OK;
}


int gh_stage_input(gh_ctx* c, const uint8_t* in, uint64_t n, uint64_t hist256[256]) {
  using namespace gh;
  if (!c || !hist256) return GH_ERR_ARG;
  if (n == 0) return GH_ERR_EMPTY;
  if (!in) return GH_ERR_ARG;
  c->staged_n = 0;
  int rc = grow(reinterpret_cast<void**>(&c->d_in), &c->in_cap, n + 16);
  if (rc != GH_OK) return rc;
  GH_CUDA_TRY(cudaMemcpyAsync(c->d_in, in, n, cudaMemcpyHostToDevice, c->stream));
  rc = gh_histogram(c->d_in, n, c->d_small, 0, c->stream);
  if (rc != GH_OK) return rc;
  GH_CUDA_TRY(cudaMemcpyAsync(c->h_small, c->d_small, 256 * 8, cudaMemcpyDeviceToHost, c->stream));
  GH_CUDA_TRY(cudaStreamSynchronize(c->stream));
  memcpy(hist256, c->h_small, 256 * 8);
  c->staged_n = n;
  return GH_OK;
}

int gh_encode_staged(gh_ctx* c, const gh_code* code, uint8_t* out_payload, uint64_t cap, uint64_t* payload_bytes) {
  using namespace gh;
  if (!c || !code || !out_payload || !payload_bytes) return GH_ERR_ARG;
  if (c->staged_n == 0) return GH_ERR_ARG;  // gh_stage_input must come first
  const uint64_t n = c->staged_n;
  const uint64_t dev_cap = gh_encode_payload_capacity(n, code, 0);
  int rc = grow(reinterpret_cast<void**>(&c->d_out), &c->out_cap, dev_cap);
  if (rc != GH_OK) return rc;
  rc = grow(&c->d_ws, &c->ws_cap, gh_encode_workspace_bytes(n));
  if (rc != GH_OK) return rc;
  uint64_t* d_end = c->d_small + 256;
  rc = gh_encode(c->d_in, n, code, 0, 1, c->d_out, c->out_cap, d_end, c->d_ws, c->ws_cap, c->stream);
  if (rc != GH_OK) return rc;
  uint64_t* h_end = reinterpret_cast<uint64_t*>(c->h_small);
  GH_CUDA_TRY(cudaMemcpyAsync(h_end, d_end, 8, cudaMemcpyDeviceToHost, c->stream));
  GH_CUDA_TRY(cudaStreamSynchronize(c->stream));
  const uint64_t bytes = (*h_end + 7) / 8;
  *payload_bytes = bytes;
  if (bytes > cap) return GH_ERR_SPACE;
  GH_CUDA_TRY(cudaMemcpyAsync(out_payload, c->d_out, bytes, cudaMemcpyDeviceToHost, c->stream));
  GH_CUDA_TRY(cudaStreamSynchronize(c->stream));
  return GH_OK;
}

int gh_stage_payload(gh_ctx* c, const uint8_t* payload, uint64_t nbytes, const gh_code* code, uint64_t* n_symbols) {
  using namespace gh;
  if (!c || !payload || !code || !n_symbols) return GH_ERR_ARG;
  if (nbytes == 0) return GH_ERR_NO_EOF;
  c->staged_payload = 0;
  c->staged_n = 0;
  int rc = grow(reinterpret_cast<void**>(&c->d_in), &c->in_cap, nbytes + 16);
  if (rc != GH_OK) return rc;
  GH_CUDA_TRY(cudaMemcpyAsync(c->d_in, payload, nbytes, cudaMemcpyHostToDevice, c->stream));
  rc = grow(&c->d_ws, &c->ws_cap, gh_decode_workspace_bytes(nbytes) + 256);
  if (rc != GH_OK) return rc;
  gh_shard_sync res;
  rc = gh_decode_sync(c->d_in, nbytes, nbytes, code, 0, 1, &res, c->d_ws, c->ws_cap, c->stream);
  if (rc != GH_OK) return rc;
  *n_symbols = res.n_symbols;
  if (!res.eof_found) return GH_ERR_NO_EOF;
  c->staged_payload = nbytes;
  c->staged_symbols = res.n_symbols;
  c->staged_code = *code;
  return GH_OK;
}

int gh_decode_staged(gh_ctx* c, uint8_t* out, uint64_t cap) {
  using namespace gh;
  if (!c || (!out && cap)) return GH_ERR_ARG;
  if (c->staged_payload == 0) return GH_ERR_ARG;  // gh_stage_payload must come first
  const uint64_t n = c->staged_symbols;
  if (cap < n) return GH_ERR_SPACE;
  int rc = grow(reinterpret_cast<void**>(&c->d_out), &c->out_cap, n + 16);
  if (rc != GH_OK) return rc;
  rc = gh_decode_write(c->d_in, c->staged_payload, c->staged_payload, &c->staged_code, c->d_out, n, c->d_ws, c->ws_cap,
                       c->stream);
  if (rc != GH_OK) return rc;
  if (n) GH_CUDA_TRY(cudaMemcpyAsync(out, c->d_out, n, cudaMemcpyDeviceToHost, c->stream));
  GH_CUDA_TRY(cudaStreamSynchronize(c->stream));
  return GH_OK;
}

}  // extern "C"
