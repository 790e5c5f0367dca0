// ghzip <infile> <type> [outfile] [chunk_bytes [resident_max_bytes]]
// Same command shape as the reference's test driver (reference unit_tests/test.cc:291-317):
//   type 3 = canonical compress (<infile>.crs2), 4/5/6 = canonical decompress (<infile>.de) -- the three
//   reference decoders produce identical output, so all three map to the one GPU decoder.
// Built with -DGH_USE_REFERENCE_FRAME against the reference's own compressor.h when that tree is present.
#include <stdio.h>
#include <stdlib.h>

#include <exception>
#include <string>

#include "gpu_canonical_huff.h"
#ifdef GH_USE_REFERENCE_FRAME
#include "compressor.h"  // the reference's unmodified template drivers
namespace frame = glzip;
#else
#include "codec_frame.h"
namespace frame = glzip_b200;
#endif

int main(int argc, char** argv) {
  if (argc < 3) {
    fprintf(stderr, "usage: %s <infile> <type 3|4|5|6> [outfile]\n", argv[0]);
    return 2;
  }
  const std::string in(argv[1]);
  std::string out(argc > 3 ? argv[3] : "");
  const int type = atoi(argv[2]);
  // optional streaming geometry (tests: small chunks, nothing resident -> the multi-pass path on a small file)
  if (argc > 4) glzip_b200::GpuStream::set_geometry(strtoull(argv[4], NULL, 10), argc > 5 ? strtoull(argv[5], NULL, 10) : ~0ull);
  try {
    if (type == 3) {
      frame::Compressor<glzip_b200::GpuCanonicalHuffEncoder> compressor;
      compressor.set_file(in, out);
      compressor.compress();
    } else if (type >= 4 && type <= 6) {
      frame::Decompressor<glzip_b200::GpuCanonicalHuffDecoder> decompressor(in, out);
      decompressor.decompress();
    } else {
      fprintf(stderr, "unknown type %d\n", type);
      return 2;
    }
  } catch (const std::exception& e) {
    fprintf(stderr, "ghzip: %s\n", e.what());
    return 1;
  }
  printf("%s\n", out.c_str());
  return 0;
}
