// GpuCanonicalHuffEncoder / GpuCanonicalHuffDecoder -- host C++ adapters that satisfy the reference's
// duck-typed _Encoder / _Decoder contract (reference include/compressor.h:47-73, 84-92) so that
//     glzip::Compressor<glzip_b200::GpuCanonicalHuffEncoder>   and
//     glzip::Decompressor<glzip_b200::GpuCanonicalHuffDecoder>
// compile against the reference's UNMODIFIED compressor.h and produce / consume the same ".crs2" files as
// CanonicalHuffEncoder<> / CanonicalHuffDecoder<> (reference include/canonical_huff_encoder.{h,cc}).
//
// Same names, argument meaning and file-name behaviour as the reference:
//   encoder: E(in, out&), E(), set_file(in, out&), clear(), caculate_frequency() [sic], gen_encode(),
//            write_encode_info(), encode_file();  empty `out` becomes in + ".crs2"  (canonical_huff_encoder.cc:15-23)
//   decoder: D(in, out&), get_encode_info(), decode_file(); empty `out` becomes in + ".de"  (encoder.h:227-232)
// Error behaviour: the reference has no error channel (unchecked fopen, UB on bad input -- encoder.h:67-70);
// the template contract returns void, so these adapters throw std::runtime_error on any non-zero C-ABI status
// (empty input, code length > 32, malformed stream, no CUDA device: there is no CPU fallback).
//
// Every per-byte operation happens on the GPU through the C ABI (include/gh_codec.h). Files are STREAMED, as the
// reference streams them through its 64 KiB buffers (include/encoder.h:55,136-150): chunks of 64 MiB through two
// pinned host buffers (gh_stream_*), so a file of any size needs a fixed amount of host and device memory, and the
// disk, the PCIe link and the kernels overlap. Header-only, depends on gh_codec.h and libgh_b200.so.
#ifndef GPU_CANONICAL_HUFF_H_
#define GPU_CANONICAL_HUFF_H_

#include <stdio.h>

#include <stdexcept>
#include <string>

#include "gh_codec.h"

namespace glzip_b200 {

inline void gh_check(int status, const char* what) {
  if (status != GH_OK) throw std::runtime_error(std::string(what) + ": " + gh_strerror(status));
}

class GpuStream {  // one per encoder/decoder object; owns the pinned and device chunk buffers
 public:
  GpuStream() : s_(NULL) {}
  ~GpuStream() { reset(); }
  gh_stream* get() {
    // chunk size and resident limit: the library's defaults (64 MiB; a quarter of the free device memory), or the
    // values a test set through set_geometry()
    if (!s_) gh_check(gh_stream_create(&s_, chunk_bytes(), resident_max()), "gh_stream_create");
    return s_;
  }
  void reset() {
    if (s_) gh_stream_destroy(s_);
    s_ = NULL;
  }
  static uint64_t& chunk_bytes() {
    static uint64_t v = 0;
    return v;
  }
  static uint64_t& resident_max() {
    static uint64_t v = ~0ull;
    return v;
  }
  static void set_geometry(uint64_t chunk, uint64_t resident) { chunk_bytes() = chunk, resident_max() = resident; }

 private:
  GpuStream(const GpuStream&);
  GpuStream& operator=(const GpuStream&);
  gh_stream* s_;
};

class GpuCanonicalHuffEncoder {
 public:
  GpuCanonicalHuffEncoder(const std::string& infile_name, std::string& outfile_name) : infile_(NULL), outfile_(NULL) {
    set_file(infile_name, outfile_name);
  }
  GpuCanonicalHuffEncoder() : infile_(NULL), outfile_(NULL) {}
  ~GpuCanonicalHuffEncoder() { clear(); }

  void set_file(const std::string& infile_name, std::string& outfile_name) {
    clear();
    infile_name_ = infile_name;
    infile_ = fopen(infile_name.c_str(), "rb");
    if (outfile_name.empty()) outfile_name = infile_name + ".crs2";
    outfile_ = fopen(outfile_name.c_str(), "wb");
    if (!infile_ || !outfile_) throw std::runtime_error("GpuCanonicalHuffEncoder: cannot open " + infile_name + " / " + outfile_name);
  }

  void clear() {
    if (infile_) fclose(infile_);
    if (outfile_) fclose(outfile_);
    infile_ = outfile_ = NULL;
  }

  // step 1 of Compressor::compress(): pass 1 over the file, chunk by chunk -> K1 (reference encoder.h:136-150)
  void caculate_frequency() {
    fseek(infile_, 0, SEEK_SET);
    gh_check(gh_stream_histogram(gpu_.get(), infile_, hist_), "gh_stream_histogram");
  }

  // step 2: code lengths + canonical codewords, reference tie-breaking (canonical_huff_encoder.cc:35-42)
  void gen_encode() { gh_check(gh_build_code(hist_, &code_), "gh_build_code"); }

  // step 3: header at file offset 0 (canonical_huff_encoder.cc:210-242)
  void write_encode_info() {
    unsigned char hdr[1040 + 8 * 32];
    size_t n = 0;
    gh_check(gh_write_header(&code_, hdr, sizeof(hdr), &n), "gh_write_header");
    fseek(outfile_, 0, SEEK_SET);
    if (fwrite(hdr, 1, n, outfile_) != n) throw std::runtime_error("GpuCanonicalHuffEncoder: header write failed");
    fflush(outfile_);
  }

  // step 4: payload, end mark and 1-padding (canonical_huff_encoder.cc:245-285)
  void encode_file() {
    const uint64_t bytes = (gh_payload_bits(&code_, hist_, 1) + 7) / 8;
    uint64_t got = 0;
    fseek(infile_, 0, SEEK_SET);
    gh_check(gh_stream_encode(gpu_.get(), infile_, outfile_, &code_, &got), "gh_stream_encode");
    if (got != bytes) throw std::runtime_error("GpuCanonicalHuffEncoder: payload size mismatch");
  }

  const gh_code& code() const { return code_; }

 private:
  GpuCanonicalHuffEncoder(const GpuCanonicalHuffEncoder&);
  GpuCanonicalHuffEncoder& operator=(const GpuCanonicalHuffEncoder&);
  FILE* infile_;
  FILE* outfile_;
  std::string infile_name_;
  uint64_t hist_[256];
  gh_code code_;
  GpuStream gpu_;
};

class GpuCanonicalHuffDecoder {
 public:
  GpuCanonicalHuffDecoder(const std::string& infile_name, std::string& outfile_name) : header_bytes_(0) {
    infile_ = fopen(infile_name.c_str(), "rb");
    if (outfile_name.empty()) outfile_name = infile_name + ".de";
    outfile_ = fopen(outfile_name.c_str(), "wb");
    if (!infile_ || !outfile_) throw std::runtime_error("GpuCanonicalHuffDecoder: cannot open " + infile_name + " / " + outfile_name);
  }
  ~GpuCanonicalHuffDecoder() {
    if (infile_) fclose(infile_);
    if (outfile_) fclose(outfile_);
  }

  // step 1 of Decompressor::decompress(): header -> tables (canonical_huff_encoder.cc:349-374)
  void get_encode_info() {
    unsigned char hdr[1040 + 8 * 32];
    fseek(infile_, 0, SEEK_SET);
    const size_t got = fread(hdr, 1, sizeof(hdr), infile_);
    gh_check(gh_parse_header(hdr, got, &code_, &header_bytes_), "gh_parse_header");
  }

  // step 2: decode up to the end mark (canonical_huff_encoder.cc:377-419)
  void decode_file() {
    uint64_t n = 0;
    if (fseek(infile_, long(header_bytes_), SEEK_SET) != 0) throw std::runtime_error("GpuCanonicalHuffDecoder: seek failed");
    gh_check(gh_stream_decode(gpu_.get(), infile_, outfile_, &code_, header_bytes_, &n), "gh_stream_decode");
  }

 private:
  GpuCanonicalHuffDecoder(const GpuCanonicalHuffDecoder&);
  GpuCanonicalHuffDecoder& operator=(const GpuCanonicalHuffDecoder&);
  FILE* infile_;
  FILE* outfile_;
  size_t header_bytes_;
  gh_code code_;
  GpuStream gpu_;
};

}  // namespace glzip_b200
#endif  // GPU_CANONICAL_HUFF_H_
