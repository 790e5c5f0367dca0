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
// Every per-byte operation happens on the GPU through the C ABI (include/gh_codec.h); this file only moves
// whole files between disk and pinned-free host buffers. Header-only, depends on gh_codec.h and libgh_b200.so.
#ifndef GPU_CANONICAL_HUFF_H_
#define GPU_CANONICAL_HUFF_H_

#include <stdio.h>

#include <stdexcept>
#include <string>
#include <vector>

#include "gh_codec.h"

namespace glzip_b200 {

inline void gh_check(int status, const char* what) {
  if (status != GH_OK) throw std::runtime_error(std::string(what) + ": " + gh_strerror(status));
}

class GpuContext {  // one per encoder/decoder object; owns the device scratch buffers
 public:
  GpuContext() : ctx_(NULL) {}
  ~GpuContext() { reset(); }
  gh_ctx* get() {
    if (!ctx_) gh_check(gh_ctx_create(&ctx_), "gh_ctx_create");
    return ctx_;
  }
  void reset() {
    if (ctx_) gh_ctx_destroy(ctx_);
    ctx_ = NULL;
  }

 private:
  GpuContext(const GpuContext&);
  GpuContext& operator=(const GpuContext&);
  gh_ctx* ctx_;
};

inline bool read_whole_file(FILE* f, std::vector<unsigned char>& buf) {
  if (!f) return false;
  if (fseek(f, 0, SEEK_END) != 0) return false;
  long size = ftell(f);
  if (size < 0 || fseek(f, 0, SEEK_SET) != 0) return false;
  buf.resize(size_t(size));
  return size == 0 || fread(&buf[0], 1, size_t(size), f) == size_t(size);
}

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

  // step 1 of Compressor::compress(): whole file -> device, K1 histogram (reference encoder.h:136-150)
  void caculate_frequency() {
    if (!read_whole_file(infile_, input_)) throw std::runtime_error("GpuCanonicalHuffEncoder: read failed: " + infile_name_);
    gh_check(gh_stage_input(gpu_.get(), input_.empty() ? NULL : &input_[0], input_.size(), hist_), "gh_stage_input");
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
    std::vector<unsigned char> payload(bytes + 16);
    uint64_t got = 0;
    gh_check(gh_encode_staged(gpu_.get(), &code_, &payload[0], payload.size(), &got), "gh_encode_staged");
    if (got != bytes) throw std::runtime_error("GpuCanonicalHuffEncoder: payload size mismatch");
    if (fwrite(&payload[0], 1, got, outfile_) != got) throw std::runtime_error("GpuCanonicalHuffEncoder: payload write failed");
    fflush(outfile_);
    std::vector<unsigned char>().swap(input_);
  }

  const gh_code& code() const { return code_; }

 private:
  GpuCanonicalHuffEncoder(const GpuCanonicalHuffEncoder&);
  GpuCanonicalHuffEncoder& operator=(const GpuCanonicalHuffEncoder&);
  FILE* infile_;
  FILE* outfile_;
  std::string infile_name_;
  std::vector<unsigned char> input_;
  uint64_t hist_[256];
  gh_code code_;
  GpuContext gpu_;
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
    if (!read_whole_file(infile_, image_)) throw std::runtime_error("GpuCanonicalHuffDecoder: read failed");
    gh_check(gh_parse_header(image_.empty() ? NULL : &image_[0], image_.size(), &code_, &header_bytes_), "gh_parse_header");
  }

  // step 2: decode up to the end mark (canonical_huff_encoder.cc:377-419)
  void decode_file() {
    if (image_.size() <= header_bytes_) throw std::runtime_error("GpuCanonicalHuffDecoder: no payload");
    uint64_t n = 0;
    gh_check(gh_stage_payload(gpu_.get(), &image_[header_bytes_], image_.size() - header_bytes_, &code_, &n), "gh_stage_payload");
    std::vector<unsigned char> out(n ? n : 1);
    gh_check(gh_decode_staged(gpu_.get(), &out[0], n), "gh_decode_staged");
    if (n && fwrite(&out[0], 1, n, outfile_) != n) throw std::runtime_error("GpuCanonicalHuffDecoder: write failed");
    fflush(outfile_);
  }

 private:
  GpuCanonicalHuffDecoder(const GpuCanonicalHuffDecoder&);
  GpuCanonicalHuffDecoder& operator=(const GpuCanonicalHuffDecoder&);
  FILE* infile_;
  FILE* outfile_;
  std::vector<unsigned char> image_;
  size_t header_bytes_;
  gh_code code_;
  GpuContext gpu_;
};

}  // namespace glzip_b200
#endif  // GPU_CANONICAL_HUFF_H_
