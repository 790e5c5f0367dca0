// The reference's template-method drivers, restated for builds where the reference tree is absent (the GPU
// box): same contract as glzip::Compressor<_Encoder> / glzip::Decompressor<_Decoder>
// (reference include/compressor.h:44-95) -- four duck-typed calls on the encoder, two on the decoder, in a
// fixed order. Where /root/reference is present the adapters are also compiled against the reference's own
// compressor.h (tests/test_cpp_adapters.py) to show they drop in unchanged.
#ifndef GH_CODEC_FRAME_H_
#define GH_CODEC_FRAME_H_
#include <string>

namespace glzip_b200 {

template <typename EncoderT>
class Compressor {
 public:
  Compressor() {}
  Compressor(const std::string& in, std::string& out) : enc_(in, out) {}
  void set_file(const std::string& in, std::string& out) { enc_.set_file(in, out); }
  void clear() { enc_.clear(); }
  void compress() {
    enc_.caculate_frequency();
    enc_.gen_encode();
    enc_.write_encode_info();
    enc_.encode_file();
  }

 private:
  EncoderT enc_;
};

template <typename DecoderT>
class Decompressor {
 public:
  Decompressor(const std::string& in, std::string& out) : dec_(in, out) {}
  void decompress() {
    dec_.get_encode_info();
    dec_.decode_file();
  }

 private:
  DecoderT dec_;
};

}  // namespace glzip_b200
#endif
