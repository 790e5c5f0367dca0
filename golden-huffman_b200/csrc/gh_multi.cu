// N GPUs of one box from ONE process, behind the C ABI: whole ".crs2" images from / to host buffers, the input cut into
// contiguous shards, one host thread and one device per shard (without torch.distributed: the data
// comes from and goes back to host memory, so the only things the shards exchange -- 256 counters, one boundary byte, one
// bit position -- cross on the host). The same device may be listed more than once (the shards then share it), which is
// how the sharding logic is tested on a single GPU.
//
//   compress    per shard: H2D, K1 histogram                                    | host: sum, gh_build_code, header, every
//               shard's bit total and start bit S_d from ITS histogram          | per shard: gh_encode(start_bit = S_d mod 8,
//               end mark on the last one), D2H straight into the image; the byte two shards share is OR-ed on the host.
//   decompress  per shard: H2D of its 16-byte-aligned slice of the payload plus a 4 KiB left halo and a right halo; the
//               halo is walked from an arbitrary bit (a self-synchronising code has found the true codeword boundaries
//               long before its end), which gives the slice's first codeword; gh_decode_sync      | host: every shard's
//               entry must equal its left neighbour's exit (else that shard is re-synchronised from the true entry, in
//               order); first end mark, output offsets                                             | per shard:
//               gh_decode_write, D2H into its place of the output.
#include <string.h>

#include <thread>
#include <vector>

#include "gh_common.cuh"

namespace gh {

constexpr uint64_t kMultiLeftHalo = 4096;
constexpr uint64_t kMultiRightHalo = 64;

struct MultiShard {
  int device;
  cudaStream_t stream;
  uint8_t* d_in;
  uint8_t* d_out;
  void* d_ws;
  uint64_t* d_small;
  int rc;
  // compress
  uint64_t in_off, n;
  uint64_t hist[256];
  uint64_t start_bit, bits;
  uint8_t first_byte;
  // decompress
  uint64_t cut, slice, lead_halo, copied;
  uint32_t entry;
  gh_shard_sync res;
  uint64_t out_off;
};

static int shard_begin(MultiShard& s) {
  GH_CUDA_TRY(cudaSetDevice(s.device));
  GH_CUDA_TRY(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
  return GH_OK;
}
static void shard_end(MultiShard& s) {
  cudaSetDevice(s.device);
  if (s.d_in) cudaFree(s.d_in);
  if (s.d_out) cudaFree(s.d_out);
  if (s.d_ws) cudaFree(s.d_ws);
  if (s.d_small) cudaFree(s.d_small);
  if (s.stream) cudaStreamDestroy(s.stream);
  s.d_in = s.d_out = nullptr, s.d_ws = nullptr, s.d_small = nullptr, s.stream = nullptr;
}

// runs fn(shard) for every shard: one host thread per shard (the emulated build has one "device" and is not
// re-entrant: there the shards run one after another)
template <class F>
static void for_each_shard(std::vector<MultiShard>& sh, F fn) {
#ifdef GH_EMUL
  for (auto& s : sh) s.rc = fn(s);
#else
  std::vector<std::thread> th;
  th.reserve(sh.size());
  for (auto& s : sh) th.emplace_back([&s, &fn]() { s.rc = fn(s); });
  for (auto& t : th) t.join();
#endif
}
static int first_error(const std::vector<MultiShard>& sh) {
  for (const auto& s : sh)
    if (s.rc != GH_OK) return s.rc;
  return GH_OK;
}

static std::vector<MultiShard> make_shards(int n_shards, const int* devices) {
  std::vector<MultiShard> sh(size_t(n_shards > 0 ? n_shards : 0));
  for (int d = 0; d < n_shards; ++d) {
    memset(&sh[size_t(d)], 0, sizeof(MultiShard));
    sh[size_t(d)].device = devices ? devices[d] : d;
  }
  return sh;
}

}  // namespace gh

extern "C" {

int gh_compress_host_multi(int n_shards, const int* devices, const uint8_t* in, uint64_t n, uint8_t* out, uint64_t cap,
                           uint64_t* out_bytes) {
  using namespace gh;
  if (n_shards < 1 || !out || !out_bytes) return GH_ERR_ARG;
  if (n == 0) return GH_ERR_EMPTY;
  if (!in) return GH_ERR_ARG;
  if (uint64_t(n_shards) > n / 64 + 1) n_shards = int(n / 64 + 1);  // every shard holds at least a few bytes
  std::vector<MultiShard> sh = make_shards(n_shards, devices);
  const int D = n_shards;
  for (int d = 0; d < D; ++d) {
    const uint64_t a = d == 0 ? 0 : (n / uint64_t(D) * uint64_t(d)) / 32 * 32;
    const uint64_t b = d + 1 == D ? n : (n / uint64_t(D) * uint64_t(d + 1)) / 32 * 32;
    sh[size_t(d)].in_off = a;
    sh[size_t(d)].n = b - a;
  }
  // 1. per shard: input to its device, K1
  for_each_shard(sh, [&](MultiShard& s) -> int {
    int rc = shard_begin(s);
    if (rc != GH_OK) return rc;
    GH_CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&s.d_in), size_t(s.n) + 64));
    GH_CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&s.d_small), 4096));
    GH_CUDA_TRY(cudaMemcpyAsync(s.d_in, in + s.in_off, size_t(s.n), cudaMemcpyHostToDevice, s.stream));
    rc = gh_histogram(s.d_in, s.n, s.d_small, 0, s.stream);
    if (rc != GH_OK) return rc;
    GH_CUDA_TRY(cudaMemcpyAsync(s.hist, s.d_small, 256 * 8, cudaMemcpyDeviceToHost, s.stream));
    GH_CUDA_TRY(cudaStreamSynchronize(s.stream));
    return GH_OK;
  });
  int rc = first_error(sh);
  gh_code code;
  size_t hdr = 0;
  uint64_t total_bits = 0;
  if (rc == GH_OK) {
    // 2. the code of the whole input, the header, every shard's place in the stream
    uint64_t hist[256];
    for (int b = 0; b < 256; ++b) {
      hist[b] = 0;
      for (int d = 0; d < D; ++d) hist[b] += sh[size_t(d)].hist[b];
    }
    rc = gh_build_code(hist, &code);
    if (rc == GH_OK) {
      hdr = gh_header_bytes(&code);
      for (int d = 0; d < D; ++d) {
        sh[size_t(d)].start_bit = total_bits;
        sh[size_t(d)].bits = gh_payload_bits(&code, sh[size_t(d)].hist, d + 1 == D ? 1 : 0);
        total_bits += sh[size_t(d)].bits;
      }
      if (cap < hdr + (total_bits + 7) / 8) rc = GH_ERR_SPACE;
    }
    if (rc == GH_OK) {
      size_t written = 0;
      rc = gh_write_header(&code, out, size_t(cap), &written);
    }
  }
  if (rc == GH_OK) {
    // 3. per shard: pack with the shard's bit phase, copy the bytes it owns straight into the image
    for_each_shard(sh, [&](MultiShard& s) -> int {
      GH_CUDA_TRY(cudaSetDevice(s.device));
      const uint32_t phase = uint32_t(s.start_bit & 7);
      const bool last = &s == &sh.back();
      const uint64_t pcap = gh_encode_payload_capacity(s.n, &code, phase);
      const size_t ws_bytes = gh_encode_workspace_bytes(s.n) + 256;
      GH_CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&s.d_out), size_t(pcap)));
      GH_CUDA_TRY(cudaMalloc(&s.d_ws, ws_bytes));
      int r = gh_encode(s.d_in, s.n, &code, phase, last ? 1 : 0, s.d_out, pcap, nullptr, s.d_ws, ws_bytes, s.stream);
      if (r != GH_OK) return r;
      const uint64_t nbytes = (phase + s.bits + 7) / 8;
      uint8_t* dst = out + hdr + (s.start_bit >> 3);
      // the first byte, when shared with the left neighbour, is handed over separately (the neighbour writes that byte)
      const uint64_t skip = phase ? 1 : 0;
      if (skip) GH_CUDA_TRY(cudaMemcpyAsync(&s.first_byte, s.d_out, 1, cudaMemcpyDeviceToHost, s.stream));
      if (nbytes > skip)
        GH_CUDA_TRY(cudaMemcpyAsync(dst + skip, s.d_out + skip, size_t(nbytes - skip), cudaMemcpyDeviceToHost, s.stream));
      GH_CUDA_TRY(cudaStreamSynchronize(s.stream));
      return GH_OK;
    });
    rc = first_error(sh);
  }
  if (rc == GH_OK) {
    for (int d = 1; d < D; ++d)
      if (sh[size_t(d)].start_bit & 7) out[hdr + (sh[size_t(d)].start_bit >> 3)] |= sh[size_t(d)].first_byte;
    *out_bytes = hdr + (total_bits + 7) / 8;
  }
  for (auto& s : sh) shard_end(s);
  return rc;
}

int gh_decompress_host_multi(int n_shards, const int* devices, const uint8_t* in, uint64_t n, uint8_t* out, uint64_t cap,
                             uint64_t* out_bytes) {
  using namespace gh;
  if (n_shards < 1 || !in || !out_bytes || (!out && cap)) return GH_ERR_ARG;
  gh_code code;
  size_t hdr = 0;
  int rc = gh_parse_header(in, size_t(n < 1296 ? n : 1296), &code, &hdr);
  if (rc != GH_OK) return rc;
  if (n <= hdr) return GH_ERR_NO_EOF;
  // the stream the kernels see starts at the 32-byte boundary below the payload (its first `lead` bytes are header)
  const uint64_t base = uint64_t(hdr) & ~31ull, lead = uint64_t(hdr) - base, stream_bytes = n - base;
  uint64_t min_slice = kMultiLeftHalo + 4096;
  if (uint64_t(n_shards) > stream_bytes / min_slice + 1) n_shards = int(stream_bytes / min_slice + 1);
  const int D = n_shards;
  std::vector<MultiShard> sh = make_shards(D, devices);
  for (int d = 0; d < D; ++d) {
    const uint64_t a = d == 0 ? 0 : (stream_bytes / uint64_t(D) * uint64_t(d)) / 32 * 32;
    const uint64_t b = d + 1 == D ? stream_bytes : (stream_bytes / uint64_t(D) * uint64_t(d + 1)) / 32 * 32;
    sh[size_t(d)].cut = a;
    sh[size_t(d)].slice = b - a;
  }
  uint64_t max_slice = 0;
  for (const auto& s : sh) max_slice = s.slice > max_slice ? s.slice : max_slice;
  const size_t ws_bytes_max = gh_decode_workspace_bytes(max_slice + kMultiRightHalo + kMultiLeftHalo) + 512;
  // 1. per shard: slice + halos to the device, entry from the left halo, self-synchronisation
  for_each_shard(sh, [&](MultiShard& s) -> int {
    int r = shard_begin(s);
    if (r != GH_OK) return r;
    s.lead_halo = s.cut >= kMultiLeftHalo ? kMultiLeftHalo : 0;  // shard 0 (and tiny streams) have no left halo
    const uint64_t from = s.cut - s.lead_halo;
    const uint64_t want = s.lead_halo + s.slice + kMultiRightHalo;
    s.copied = from + want <= stream_bytes ? want : stream_bytes - from;
    GH_CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&s.d_in), size_t(want) + 64));
    GH_CUDA_TRY(cudaMalloc(&s.d_ws, ws_bytes_max));
    GH_CUDA_TRY(cudaMemcpyAsync(s.d_in, in + base + from, size_t(s.copied), cudaMemcpyHostToDevice, s.stream));
    void* ws = reinterpret_cast<void*>((reinterpret_cast<uintptr_t>(s.d_ws) + 255) & ~uintptr_t(255));
    const size_t wsb = ws_bytes_max - 256;
    s.entry = s.cut == 0 ? uint32_t(lead * 8) : 0u;
    if (s.lead_halo) {
      gh_shard_sync walk;
      r = gh_decode_sync(s.d_in, s.lead_halo, s.copied, &code, 0, 1, &walk, ws, wsb, s.stream);
      if (r != GH_OK) return r;
      s.entry = walk.exit_bit;
    }
    return gh_decode_sync(s.d_in + s.lead_halo, s.slice, s.copied - s.lead_halo, &code, s.entry, 1, &s.res, ws, wsb, s.stream);
  });
  rc = first_error(sh);
  // 2. the chain of entries: a shard that started from a wrong bit is re-synchronised from its neighbour's exit
  if (rc == GH_OK) {
    for (int d = 1; d < D && rc == GH_OK; ++d) {
      MultiShard& s = sh[size_t(d)];
      const uint32_t want = sh[size_t(d - 1)].res.exit_bit;
      if (sh[size_t(d - 1)].res.eof_found) break;  // nothing after the end mark is decoded
      if (s.entry != want) {
        if (cudaSetDevice(s.device) != cudaSuccess) {
          rc = cuda_fail(cudaGetLastError());
          break;
        }
        void* ws = reinterpret_cast<void*>((reinterpret_cast<uintptr_t>(s.d_ws) + 255) & ~uintptr_t(255));
        s.entry = want;
        rc = gh_decode_sync(s.d_in + s.lead_halo, s.slice, s.copied - s.lead_halo, &code, s.entry, 0, &s.res, ws, ws_bytes_max - 256,
                            s.stream);
      }
    }
  }
  uint64_t total = 0;
  bool eof = false;
  if (rc == GH_OK) {
    for (int d = 0; d < D; ++d) {
      MultiShard& s = sh[size_t(d)];
      s.out_off = total;
      if (eof) s.res.n_symbols = 0;
      total += s.res.n_symbols;
      eof = eof || s.res.eof_found != 0;
    }
    *out_bytes = total;
    if (!eof) rc = GH_ERR_NO_EOF;
    else if (total > cap) rc = GH_ERR_SPACE;
  }
  // 3. per shard: write, copy to its place of the output
  if (rc == GH_OK) {
    for_each_shard(sh, [&](MultiShard& s) -> int {
      if (s.res.n_symbols == 0) return GH_OK;
      GH_CUDA_TRY(cudaSetDevice(s.device));
      void* ws = reinterpret_cast<void*>((reinterpret_cast<uintptr_t>(s.d_ws) + 255) & ~uintptr_t(255));
      GH_CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&s.d_out), size_t(s.res.n_symbols) + 64));
      int r = gh_decode_write(s.d_in + s.lead_halo, s.slice, s.copied - s.lead_halo, &code, s.d_out, s.res.n_symbols, ws,
                              ws_bytes_max - 256, s.stream);
      if (r != GH_OK) return r;
      GH_CUDA_TRY(cudaMemcpyAsync(out + s.out_off, s.d_out, size_t(s.res.n_symbols), cudaMemcpyDeviceToHost, s.stream));
      GH_CUDA_TRY(cudaStreamSynchronize(s.stream));
      return GH_OK;
    });
    rc = first_error(sh);
  }
  for (auto& s : sh) shard_end(s);
  return rc;
}

}  // extern "C"
