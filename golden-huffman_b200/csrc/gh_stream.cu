// Streaming file <-> device layer for the adapters: what the reference does with its 64 KiB
// FixedFileBuffer (include/encoder.h:55,136-150; include/canonical_huff_encoder.cc:245-285, 377-419) done with chunks
// of tens of MiB through two pinned host buffers, so that a file of any size goes through a fixed amount of host and
// device memory and the disk, the PCIe link and the kernels work at the same time:
//
//   gh_stream_histogram   pass 1 of Compressor::compress(): chunk k+1 is read from the file while chunk k is copied to
//                         the device and counted (K1 per chunk, so every chunk's bit total is known afterwards). A file
//                         that fits the device budget stays resident; a larger one is read again by pass 2, exactly as
//                         the reference reads its input twice.
//   gh_stream_encode      pass 2: per chunk H2D (unless resident) -> gh_encode with the running bit phase -> D2H, while
//                         the previous chunk's payload is written to the file. Chunks meet inside a byte: the byte that
//                         two chunks share is OR-ed together on the host (one byte per chunk).
//   gh_stream_decode      Decompressor::decompress(): payload chunks in order; chunk k+1's first codeword starts where
//                         chunk k's decode ran out (gh_decode_sync's exit bit), so nothing but that bit position crosses
//                         from one chunk to the next; the chunk that contains the end mark ends the stream.
#include <stdio.h>
#include <string.h>

#include <vector>

#include "gh_common.cuh"

struct gh_stream {
  cudaStream_t stream;
  uint64_t chunk;          // input bytes per chunk (multiple of 4096)
  uint64_t resident_max;   // inputs up to this size stay on the device between the two passes
  uint8_t* h_in[2];        // pinned
  uint8_t* h_out[2];       // pinned
  size_t h_in_cap, h_out_cap[2];
  uint8_t* d_in;           // resident input, or two chunk slots
  size_t d_in_cap;
  uint8_t* d_out[2];
  size_t d_out_cap[2];
  void* d_ws;
  size_t ws_cap;
  uint64_t* d_hists;       // per-chunk histograms
  size_t hists_cap;        // in chunks
  cudaEvent_t ev_in[2], ev_out[2];
  // state left by gh_stream_histogram for gh_stream_encode
  bool resident;
  uint64_t file_bytes;
  std::vector<uint64_t> chunk_hist;  // n_chunks x 256
};

namespace gh {

static int sgrow_dev(uint8_t** p, size_t* cap, size_t need) {
  if (*cap >= need) return GH_OK;
  if (*p) GH_CUDA_TRY(cudaFree(*p));
  *p = nullptr, *cap = 0;
  GH_CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(p), need + 4096));
  *cap = need + 4096;
  return GH_OK;
}
static int sgrow_host(uint8_t** p, size_t* cap, size_t need) {
  if (*cap >= need) return GH_OK;
  if (*p) GH_CUDA_TRY(cudaFreeHost(*p));
  *p = nullptr, *cap = 0;
  GH_CUDA_TRY(cudaMallocHost(reinterpret_cast<void**>(p), need + 4096));
  *cap = need + 4096;
  return GH_OK;
}

static uint64_t file_size_from_here(FILE* f) {
  const long at = ftell(f);
  if (at < 0 || fseek(f, 0, SEEK_END) != 0) return ~0ull;
  const long end = ftell(f);
  if (end < 0 || fseek(f, at, SEEK_SET) != 0) return ~0ull;
  return uint64_t(end - at);
}

}  // namespace gh

extern "C" {

int gh_stream_create(gh_stream** out, uint64_t chunk_bytes, uint64_t resident_max_bytes) {
  using namespace gh;
  if (!out) return GH_ERR_ARG;
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) return cuda_fail(e);
  gh_stream* s = new gh_stream();
  s->stream = nullptr;
  s->chunk = chunk_bytes ? (chunk_bytes + 4095) / 4096 * 4096 : (64ull << 20);
  s->resident_max = resident_max_bytes;
  if (resident_max_bytes == ~0ull) {  // automatic: a quarter of what the device has free now
    size_t free_b = 0, total_b = 0;
    s->resident_max = cudaMemGetInfo(&free_b, &total_b) == cudaSuccess ? uint64_t(free_b) / 4 : 0;
  }
  s->h_in[0] = s->h_in[1] = s->h_out[0] = s->h_out[1] = nullptr;
  s->h_in_cap = s->h_out_cap[0] = s->h_out_cap[1] = 0;
  s->d_in = nullptr, s->d_in_cap = 0;
  s->d_out[0] = s->d_out[1] = nullptr;
  s->d_out_cap[0] = s->d_out_cap[1] = 0;
  s->d_ws = nullptr, s->ws_cap = 0;
  s->d_hists = nullptr, s->hists_cap = 0;
  s->resident = false, s->file_bytes = 0;
  for (int i = 0; i < 2; ++i) s->ev_in[i] = s->ev_out[i] = nullptr;
  if (cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking) != cudaSuccess) {
    const int rc = cuda_fail(cudaGetLastError());
    delete s;
    return rc;
  }
  for (int i = 0; i < 2; ++i) {
    if (cudaEventCreateWithFlags(&s->ev_in[i], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&s->ev_out[i], cudaEventDisableTiming) != cudaSuccess) {
      const int rc = cuda_fail(cudaGetLastError());
      delete s;
      return rc;
    }
  }
  *out = s;
  return GH_OK;
}

void gh_stream_destroy(gh_stream* s) {
  if (!s) return;
  for (int i = 0; i < 2; ++i) {
    if (s->h_in[i]) cudaFreeHost(s->h_in[i]);
    if (s->h_out[i]) cudaFreeHost(s->h_out[i]);
    if (s->d_out[i]) cudaFree(s->d_out[i]);
    if (s->ev_in[i]) cudaEventDestroy(s->ev_in[i]);
    if (s->ev_out[i]) cudaEventDestroy(s->ev_out[i]);
  }
  if (s->d_in) cudaFree(s->d_in);
  if (s->d_ws) cudaFree(s->d_ws);
  if (s->d_hists) cudaFree(s->d_hists);
  if (s->stream) cudaStreamDestroy(s->stream);
  delete s;
}

// Pass 1 (caculate_frequency): the file from its current position to its end. hist256 receives the byte counts.
int gh_stream_histogram(gh_stream* s, FILE* in, uint64_t hist256[256]) {
  using namespace gh;
  if (!s || !in || !hist256) return GH_ERR_ARG;
  const uint64_t total = file_size_from_here(in);
  if (total == ~0ull) return GH_ERR_ARG;  // not seekable
  if (total == 0) return GH_ERR_EMPTY;
  const long start_pos = ftell(in);
  const uint64_t C = s->chunk, nchunks = (total + C - 1) / C;
  s->file_bytes = total;
  s->resident = total <= s->resident_max;
  int rc;
  for (int i = 0; i < 2; ++i) {
    uint8_t* p = s->h_in[i];
    size_t cap = s->h_in_cap;
    rc = sgrow_host(&p, &cap, size_t(C));
    if (rc != GH_OK) return rc;
    s->h_in[i] = p;
    if (i == 1) s->h_in_cap = cap;
  }
  rc = sgrow_dev(&s->d_in, &s->d_in_cap, size_t(s->resident ? nchunks * C : 2 * C) + 64);
  if (rc != GH_OK) return rc;
  if (s->hists_cap < nchunks) {
    if (s->d_hists) GH_CUDA_TRY(cudaFree(s->d_hists));
    s->d_hists = nullptr, s->hists_cap = 0;
    GH_CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&s->d_hists), size_t(nchunks) * 256 * 8));
    s->hists_cap = size_t(nchunks);
  }
  for (uint64_t k = 0; k < nchunks; ++k) {
    const int i = int(k & 1);
    const uint64_t nk = k + 1 < nchunks ? C : total - k * C;
    if (k >= 2) GH_CUDA_TRY(cudaEventSynchronize(s->ev_in[i]));  // the pinned buffer's previous copy has left it
    if (fread(s->h_in[i], 1, size_t(nk), in) != size_t(nk)) return GH_ERR_FORMAT;
    uint8_t* dst = s->d_in + (s->resident ? k * C : uint64_t(i) * C);
    GH_CUDA_TRY(cudaMemcpyAsync(dst, s->h_in[i], size_t(nk), cudaMemcpyHostToDevice, s->stream));
    GH_CUDA_TRY(cudaEventRecord(s->ev_in[i], s->stream));
    rc = gh_histogram(dst, nk, s->d_hists + k * 256, 0, s->stream);
    if (rc != GH_OK) return rc;
  }
  s->chunk_hist.resize(size_t(nchunks) * 256);
  GH_CUDA_TRY(cudaMemcpyAsync(s->chunk_hist.data(), s->d_hists, size_t(nchunks) * 256 * 8, cudaMemcpyDeviceToHost, s->stream));
  GH_CUDA_TRY(cudaStreamSynchronize(s->stream));
  for (int b = 0; b < 256; ++b) hist256[b] = 0;
  for (uint64_t k = 0; k < nchunks; ++k)
    for (int b = 0; b < 256; ++b) hist256[b] += s->chunk_hist[size_t(k) * 256 + b];
  if (fseek(in, start_pos, SEEK_SET) != 0) return GH_ERR_ARG;
  return GH_OK;
}

// Pass 2 (encode_file): the same file range again -> payload bytes appended to `out` (end mark and padding included)
int gh_stream_encode(gh_stream* s, FILE* in, FILE* out, const gh_code* code, uint64_t* payload_bytes) {
  using namespace gh;
  if (!s || !in || !out || !code) return GH_ERR_ARG;
  const uint64_t total = s->file_bytes, C = s->chunk;
  if (total == 0 || s->chunk_hist.empty()) return GH_ERR_ARG;  // gh_stream_histogram must come first
  const uint64_t nchunks = (total + C - 1) / C;
  int rc;
  const uint64_t pay_cap = gh_encode_payload_capacity(C, code, 7);
  for (int i = 0; i < 2; ++i) {
    rc = sgrow_dev(&s->d_out[i], &s->d_out_cap[i], size_t(pay_cap));
    if (rc != GH_OK) return rc;
  }
  {
    uint8_t* w = static_cast<uint8_t*>(s->d_ws);
    rc = sgrow_dev(&w, &s->ws_cap, gh_encode_workspace_bytes(C) + 256);
    s->d_ws = w;
    if (rc != GH_OK) return rc;
  }
  uint64_t S = 0, written = 0;  // payload bits before the current chunk; bytes written to the file
  uint64_t prev_bytes = 0;      // size of the previous chunk's slice in h_out[prev]
  bool prev_open = false;       // its last byte is shared with the current chunk
  uint8_t carry = 0;            // bits of the byte the previous-previous chunk left open
  bool carry_valid = false;
  auto flush_prev = [&](int i) -> int {  // write chunk k-1 (its D2H has been waited for)
    uint8_t* p = s->h_out[i];
    if (carry_valid) p[0] |= carry;
    uint64_t nbytes = prev_bytes;
    if (prev_open) {
      carry = p[nbytes - 1];
      carry_valid = true;
      nbytes -= 1;
    } else {
      carry_valid = false;
    }
    if (nbytes && fwrite(p, 1, size_t(nbytes), out) != size_t(nbytes)) return GH_ERR_SPACE;
    written += nbytes;
    return GH_OK;
  };
  for (uint64_t k = 0; k < nchunks; ++k) {
    const int i = int(k & 1);
    const bool last = k + 1 == nchunks;
    const uint64_t nk = last ? total - k * C : C;
    const uint8_t* src = s->d_in + (s->resident ? k * C : uint64_t(i) * C);
    if (!s->resident) {
      if (k >= 2) GH_CUDA_TRY(cudaEventSynchronize(s->ev_in[i]));
      if (fread(s->h_in[i], 1, size_t(nk), in) != size_t(nk)) return GH_ERR_FORMAT;
      GH_CUDA_TRY(cudaMemcpyAsync(const_cast<uint8_t*>(src), s->h_in[i], size_t(nk), cudaMemcpyHostToDevice, s->stream));
    }
    const uint64_t bits = gh_payload_bits(code, &s->chunk_hist[size_t(k) * 256], last ? 1 : 0);
    const uint32_t phase = uint32_t(S & 7);
    rc = gh_encode(src, nk, code, phase, last ? 1 : 0, s->d_out[i], s->d_out_cap[i], nullptr, s->d_ws, s->ws_cap, s->stream);
    if (rc != GH_OK) return rc;
    if (!s->resident) GH_CUDA_TRY(cudaEventRecord(s->ev_in[i], s->stream));  // the input slot may be refilled after this
    const uint64_t nbytes = (phase + bits + 7) / 8;  // the last chunk's padding completes its last byte
    rc = sgrow_host(&s->h_out[i], &s->h_out_cap[i], size_t(nbytes) + 16);
    if (rc != GH_OK) return rc;
    GH_CUDA_TRY(cudaMemcpyAsync(s->h_out[i], s->d_out[i], size_t(nbytes), cudaMemcpyDeviceToHost, s->stream));
    GH_CUDA_TRY(cudaEventRecord(s->ev_out[i], s->stream));
    // while the device works on chunk k, chunk k-1 goes to the file
    if (k > 0) {
      GH_CUDA_TRY(cudaEventSynchronize(s->ev_out[i ^ 1]));
      rc = flush_prev(i ^ 1);
      if (rc != GH_OK) return rc;
    }
    prev_bytes = nbytes;
    prev_open = !last && ((S + bits) & 7) != 0;
    S += bits;
  }
  GH_CUDA_TRY(cudaEventSynchronize(s->ev_out[int((nchunks - 1) & 1)]));
  rc = flush_prev(int((nchunks - 1) & 1));
  if (rc != GH_OK) return rc;
  fflush(out);
  if (payload_bytes) *payload_bytes = written;
  return GH_OK;
}

// decode_file: `in` is positioned at the first payload byte, header_bytes = how many bytes of the file precede it.
int gh_stream_decode(gh_stream* s, FILE* in, FILE* out, const gh_code* code, uint64_t header_bytes, uint64_t* n_out) {
  using namespace gh;
  if (!s || !in || !out || !code) return GH_ERR_ARG;
  const uint64_t remaining_file = file_size_from_here(in);
  if (remaining_file == ~0ull) return GH_ERR_ARG;
  if (remaining_file == 0) return GH_ERR_NO_EOF;
  // the kernels read from a 32-byte boundary: chunk 0 starts (header_bytes mod 32) bytes before the payload, those
  // bytes are skipped through its entry bit
  const uint64_t lead = header_bytes & 31;
  const uint64_t P = s->chunk, halo = 64;
  int rc;
  for (int i = 0; i < 2; ++i) {
    uint8_t* p = s->h_in[i];
    size_t cap = s->h_in_cap;
    rc = sgrow_host(&p, &cap, size_t(P + halo + 64));
    if (rc != GH_OK) return rc;
    s->h_in[i] = p;
    if (i == 1) s->h_in_cap = cap;
  }
  rc = sgrow_dev(&s->d_in, &s->d_in_cap, size_t(2 * (P + halo + 64)));
  if (rc != GH_OK) return rc;
  {
    uint8_t* w = static_cast<uint8_t*>(s->d_ws);
    rc = sgrow_dev(&w, &s->ws_cap, gh_decode_workspace_bytes(P + halo) + 512);
    s->d_ws = w;
    if (rc != GH_OK) return rc;
  }
  void* ws = reinterpret_cast<void*>((reinterpret_cast<uintptr_t>(s->d_ws) + 255) & ~uintptr_t(255));
  const size_t ws_bytes = s->ws_cap - 256;
  const uint64_t stream_bytes = lead + remaining_file;  // bytes from chunk 0's first byte to the end of the file
  const uint64_t nchunks = (stream_bytes + P - 1) / P;
  uint32_t entry = uint32_t(lead * 8);
  uint64_t total_out = 0;
  bool done = false;
  int pending = -1;        // h_out slot whose D2H is in flight
  uint64_t pending_n = 0;
  // chunk k = stream bytes [kP, (k+1)P) plus `halo` bytes of the next chunk; the pinned buffer of chunk k+1 starts with
  // the halo bytes already read for chunk k
  uint64_t have_next = 0;  // bytes of chunk k+1 already sitting at the front of its pinned buffer
  for (uint64_t k = 0; k < nchunks && !done; ++k) {
    const int i = int(k & 1);
    const uint64_t slice = k + 1 < nchunks ? P : stream_bytes - k * P;
    const uint64_t want = (k + 1 < nchunks ? slice + halo : slice);  // bytes of this chunk incl. halo
    uint64_t off = have_next;                                            // already there (copied below)
    if (k == 0 && lead) {
      memset(s->h_in[i], 0xff, size_t(lead));  // never decoded: skipped by the entry bit
      off = lead;
    }
    uint64_t avail = off;
    const uint64_t to_read = stream_bytes - k * P - off < want - off ? stream_bytes - k * P - off : want - off;
    if (to_read) {
      const size_t got = fread(s->h_in[i] + off, 1, size_t(to_read), in);
      avail += got;
      if (got != size_t(to_read)) return GH_ERR_FORMAT;
    }
    // the part of the next chunk that was read as this chunk's halo
    have_next = avail > slice ? avail - slice : 0;
    if (have_next) memcpy(s->h_in[i ^ 1], s->h_in[i] + slice, size_t(have_next));
    uint8_t* d = s->d_in + uint64_t(i) * (P + halo + 64);
    GH_CUDA_TRY(cudaMemcpyAsync(d, s->h_in[i], size_t(avail), cudaMemcpyHostToDevice, s->stream));
    gh_shard_sync res;
    rc = gh_decode_sync(d, slice, avail, code, entry, 1, &res, ws, ws_bytes, s->stream);
    if (rc != GH_OK) return rc;
    const uint64_t nsym = res.n_symbols;
    rc = sgrow_dev(&s->d_out[i], &s->d_out_cap[i], size_t(nsym) + 64);
    if (rc != GH_OK) return rc;
    rc = gh_decode_write(d, slice, avail, code, s->d_out[i], nsym, ws, ws_bytes, s->stream);
    if (rc != GH_OK) return rc;
    // the previous chunk's output goes to the file while this one is decoded
    if (pending >= 0) {
      GH_CUDA_TRY(cudaEventSynchronize(s->ev_out[pending]));
      if (pending_n && fwrite(s->h_out[pending], 1, size_t(pending_n), out) != size_t(pending_n)) return GH_ERR_SPACE;
      pending = -1;
    }
    rc = sgrow_host(&s->h_out[i], &s->h_out_cap[i], size_t(nsym) + 16);
    if (rc != GH_OK) return rc;
    if (nsym) GH_CUDA_TRY(cudaMemcpyAsync(s->h_out[i], s->d_out[i], size_t(nsym), cudaMemcpyDeviceToHost, s->stream));
    GH_CUDA_TRY(cudaEventRecord(s->ev_out[i], s->stream));
    pending = i;
    pending_n = nsym;
    total_out += nsym;
    done = res.eof_found != 0;
    entry = res.exit_bit;
  }
  if (pending >= 0) {
    GH_CUDA_TRY(cudaEventSynchronize(s->ev_out[pending]));
    if (pending_n && fwrite(s->h_out[pending], 1, size_t(pending_n), out) != size_t(pending_n)) return GH_ERR_SPACE;
  }
  fflush(out);
  if (n_out) *n_out = total_out;
  return done ? GH_OK : GH_ERR_NO_EOF;
}

}  // extern "C"
